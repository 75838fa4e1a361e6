#include "model.hpp"

#include <charconv>

#include <exception>
#include <thread>

#include <cstdio>
#include <cstring>

namespace pgmb {

namespace {
using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

template <class T> std::pair<T const*, T const*> scenario_span(ComponentBuffer const& b, Idx s) {
    if (b.data == nullptr) return {nullptr, nullptr};
    auto const* p = static_cast<T const*>(b.data);
    if (b.indptr != nullptr) return {p + b.indptr[s], p + b.indptr[s + 1]};
    if (b.n < 0) throw InvalidArgument("sparse batch buffer without indptr");
    return {p + s * b.n, p + (s + 1) * b.n};
}
} // namespace

Model::Model(double system_frequency, InputData const& in) : freq_{system_frequency} {
    auto add_id = [this](ID id) {
        if (!all_ids_.emplace(id, 0).second) throw InvalidArgument("Conflicting id detected: " + std::to_string(id) + "\n");
    };
    auto const* nodes = static_cast<NodeInput const*>(in.node.data);
    for (Idx i = 0; i != in.node.n; ++i) {
        add_id(nodes[i].id);
        node_idx_[nodes[i].id] = i;
        node_.push_back(nodes[i]);
    }
    auto u_rated = [this](ID id) { return node_[node_seq(id)].u_rated; };
    auto const* lines = static_cast<LineInput const*>(in.line.data);
    for (Idx i = 0; i != in.line.n; ++i) {
        LineInput const& l = lines[i];
        add_id(l.id);
        double const u1 = u_rated(l.from_node), u2 = u_rated(l.to_node);
        if (std::abs(u1 - u2) > kNumTol) throw InvalidArgument("Conflicting voltage for line " + std::to_string(l.id) + "\n");
        line_idx_[l.id] = i;
        line_in_.push_back(l);
        line_c_.push_back(line_constants(l, freq_, u1));
        branch_st_.push_back({l.from_status != 0, l.to_status != 0});
    }
    auto const* alines = static_cast<AsymLineInput const*>(in.asym_line.data);
    for (Idx i = 0; i != in.asym_line.n; ++i) {
        AsymLineInput const& l = alines[i];
        add_id(l.id);
        double const u1 = u_rated(l.from_node), u2 = u_rated(l.to_node);
        if (std::abs(u1 - u2) > kNumTol) throw InvalidArgument("Conflicting voltage for line " + std::to_string(l.id) + "\n");
        aline_idx_[l.id] = i;
        aline_in_.push_back(l);
        aline_c_.push_back(asym_line_constants(l, freq_, u1));
        branch_st_.push_back({l.from_status != 0, l.to_status != 0});
    }
    auto const* links = static_cast<LinkInput const*>(in.link.data);
    for (Idx i = 0; i != in.link.n; ++i) {
        LinkInput const& l = links[i];
        add_id(l.id);
        link_idx_[l.id] = i;
        link_in_.push_back(l);
        link_base_i_.push_back({kBasePower3p / u_rated(l.from_node) / kSqrt3, kBasePower3p / u_rated(l.to_node) / kSqrt3});
        branch_st_.push_back({l.from_status != 0, l.to_status != 0});
    }
    auto const* gbs = static_cast<GenericBranchInput const*>(in.generic_branch.data);
    for (Idx i = 0; i != in.generic_branch.n; ++i) {
        GenericBranchInput const& g = gbs[i];
        add_id(g.id);
        gb_idx_[g.id] = i;
        gb_in_.push_back(g);
        gb_c_.push_back(generic_branch_constants(g, u_rated(g.from_node), u_rated(g.to_node)));
        branch_st_.push_back({g.from_status != 0, g.to_status != 0});
    }
    auto const* trafos = static_cast<TransformerInput const*>(in.transformer.data);
    for (Idx i = 0; i != in.transformer.n; ++i) {
        TransformerInput const& t = trafos[i];
        add_id(t.id);
        trafo_idx_[t.id] = i;
        trafo_in_.push_back(t);
        trafo_c_.push_back(transformer_constants(t, u_rated(t.from_node), u_rated(t.to_node)));
        if (!trafo_c_.back().clock_valid) throw InvalidArgument("Invalid clock for transformer " + std::to_string(t.id) + "\n");
        branch_st_.push_back({t.from_status != 0, t.to_status != 0});
        trafo_st_.push_back({trafo_c_.back().initial_tap_pos});
    }
    auto const* t3ws = static_cast<ThreeWindingTransformerInput const*>(in.three_winding_transformer.data);
    for (Idx i = 0; i != in.three_winding_transformer.n; ++i) {
        ThreeWindingTransformerInput const& t = t3ws[i];
        add_id(t.id);
        t3w_idx_[t.id] = i;
        t3w_c_.push_back(three_winding_constants(t, u_rated(t.node_1), u_rated(t.node_2), u_rated(t.node_3)));
        if (!t3w_c_.back().clock_valid) throw InvalidArgument("Invalid clock for transformer " + std::to_string(t.id) + "\n");
        t3w_st_.push_back({{t.status_1 != 0, t.status_2 != 0, t.status_3 != 0}, t3w_c_.back().initial_tap_pos});
    }
    auto const* shunts = static_cast<ShuntInput const*>(in.shunt.data);
    for (Idx i = 0; i != in.shunt.n; ++i) {
        ShuntInput const& s = shunts[i];
        add_id(s.id);
        shunt_idx_[s.id] = i;
        shunt_in_.push_back(s);
        double const u = u_rated(s.node);
        double const base_i = kBasePower3p / u / kSqrt3;
        shunt_base_y_.push_back(base_i / (u / kSqrt3));
        ShuntState st{s.status != 0, kNaN, kNaN, kNaN, kNaN, cplx{kNaN, 0.0}, cplx{kNaN, 0.0}};
        shunt_set(st, shunt_base_y_.back(), s.g1, s.b1, s.g0, s.b0);
        shunt_st_.push_back(st);
    }
    auto const* sources = static_cast<SourceInput const*>(in.source.data);
    for (Idx i = 0; i != in.source.n; ++i) {
        SourceInput const& s = sources[i];
        add_id(s.id);
        (void)u_rated(s.node);
        source_idx_[s.id] = i;
        source_in_.push_back(s);
        source_st_.push_back({s.status != 0, s.u_ref, std::isnan(s.u_ref_angle) ? 0.0 : s.u_ref_angle,
                              std::isnan(s.sk) ? 1e10 : s.sk, std::isnan(s.rx_ratio) ? 0.1 : s.rx_ratio,
                              std::isnan(s.z01_ratio) ? 1.0 : s.z01_ratio});
    }
    auto add_lg = [&](auto const* p, Idx n, int lb, double direction) {
        for (Idx i = 0; i != n; ++i) {
            add_id(p[i].id);
            Idx const node = node_seq(p[i].node);
            lg_idx_[p[i].id] = static_cast<Idx>(lg_.size());
            lg_.push_back({p[i].id, node, lb, direction, kBasePower3p / node_[node].u_rated / kSqrt3, p[i].type});
            LoadGenState st{p[i].status != 0, {cplx{kNaN, kNaN}, cplx{kNaN, kNaN}, cplx{kNaN, kNaN}}};
            lg_st_.push_back(st);
            if constexpr (std::is_same_v<std::remove_cvref_t<decltype(p[i])>, SymLoadGenInput>) {
                set_load_power(static_cast<Idx>(lg_.size()) - 1, &p[i].p_specified, &p[i].q_specified);
            } else {
                set_load_power(static_cast<Idx>(lg_.size()) - 1, p[i].p_specified, p[i].q_specified);
            }
        }
    };
    n_sym_gen_ = in.sym_gen.n;
    n_asym_gen_ = in.asym_gen.n;
    n_sym_load_ = in.sym_load.n;
    n_asym_load_ = in.asym_load.n;
    add_lg(static_cast<SymLoadGenInput const*>(in.sym_gen.data), in.sym_gen.n, 1, 1.0);
    add_lg(static_cast<AsymLoadGenInput const*>(in.asym_gen.data), in.asym_gen.n, 3, 1.0);
    add_lg(static_cast<SymLoadGenInput const*>(in.sym_load.data), in.sym_load.n, 1, -1.0);
    add_lg(static_cast<AsymLoadGenInput const*>(in.asym_load.data), in.asym_load.n, 3, -1.0);
    std::unordered_map<ID, int> regulated;
    // transformer tap regulators (main_core/input.hpp:168-214): the regulated object is a transformer or a three-winding
    // transformer; the control side names one of its terminals, whose node gives the rated voltage of the set point
    auto const* tap_regs = static_cast<TransformerTapRegulatorInput const*>(in.transformer_tap_regulator.data);
    for (Idx i = 0; i != in.transformer_tap_regulator.n; ++i) {
        TransformerTapRegulatorInput const& r = tap_regs[i];
        add_id(r.id);
        if (all_ids_.find(r.regulated_object) == all_ids_.end()) {
            throw InvalidArgument("The id cannot be found: " + std::to_string(r.regulated_object) + "\n");
        }
        TapTarget target{};
        if (auto it = trafo_idx_.find(r.regulated_object); it != trafo_idx_.end()) {
            if (r.control_side != 0 && r.control_side != 1) {
                throw InvalidArgument("transformer_tap_regulator item retrieval is not implemented for ControlSide #" +
                                      std::to_string(static_cast<int>(r.control_side)) + "!\n");
            }
            TransformerInput const& t = trafo_in_[it->second];
            target = {0, it->second, u_rated(r.control_side == 0 ? t.from_node : t.to_node)};
        } else if (auto it3 = t3w_idx_.find(r.regulated_object); it3 != t3w_idx_.end()) {
            if (r.control_side < 0 || r.control_side > 2) {
                throw InvalidArgument("transformer_tap_regulator item retrieval is not implemented for ControlSide #" +
                                      std::to_string(static_cast<int>(r.control_side)) + "!\n");
            }
            ThreeWindingTransformerInput const& t = t3w_c_[it3->second].in;
            target = {1, it3->second, u_rated(r.control_side == 0 ? t.node_1 : r.control_side == 1 ? t.node_2 : t.node_3)};
        } else {
            throw InvalidArgument("transformer_tap_regulator regulator is not supported for object with ID " +
                                  std::to_string(r.regulated_object) + "\n");
        }
        regulated.emplace(r.regulated_object, 0); // duplicates are reported after every regulator has been read
        tap_reg_idx_[r.id] = i;
        tap_reg_in_.push_back(r);
        tap_reg_target_.push_back(target);
        tap_reg_st_.push_back({r.status != 0, r.u_set, r.u_band, r.line_drop_compensation_r, r.line_drop_compensation_x});
    }
    if (regulated.size() != tap_reg_in_.size()) {
        throw InvalidArgument("There are objects regulated by more than one regulator. Maximum one regulator is allowed.\n");
    }
    // voltage regulators (main_core/input.hpp:216-241): the regulated object is a load / generator, one regulator per object
    auto const* regs = static_cast<VoltageRegulatorInput const*>(in.voltage_regulator.data);
    for (Idx i = 0; i != in.voltage_regulator.n; ++i) {
        VoltageRegulatorInput const& r = regs[i];
        add_id(r.id);
        auto it = lg_idx_.find(r.regulated_object);
        if (it == lg_idx_.end()) {
            throw InvalidArgument("Wrong type for object with id " + std::to_string(r.regulated_object) + "\n");
        }
        if (!regulated.emplace(r.regulated_object, 0).second) {
            throw InvalidArgument("There are objects regulated by more than one regulator. Maximum one regulator is allowed.\n");
        }
        reg_idx_[r.id] = i;
        reg_in_.push_back(r);
        reg_lg_.push_back(it->second);
        reg_st_.push_back({r.status != 0, r.u_ref, r.q_min, r.q_max});
    }
}

// check_state_validity (main_core/calculation_preparation.hpp:163-225) + the method restriction of the regulator
// (main_model_impl.hpp:362-366, 400-420)
template <int B> void Model::check_regulators(ModelOptions const& opt) const {
    if (reg_in_.empty()) return;
    if (!tap_reg_in_.empty()) { // check_state_validity, calculation_preparation.hpp:219-222
        throw InvalidArgument("The combination of voltage regulators and transformer tap regulators is not supported in the same model.");
    }
    if (opt.method != 1 && opt.method != -128) throw InvalidArgument("The calculation method is invalid for this calculation!\n");
    std::unordered_map<Idx, std::pair<ID, double>> node_ref; // node -> (regulator id, u_ref)
    for (size_t r = 0; r != reg_in_.size(); ++r) {
        if (!reg_st_[r].status) continue;
        Idx const node = lg_[reg_lg_[r]].node;
        auto it = node_ref.find(node);
        if (it != node_ref.end()) {
            if (it->second.second != reg_st_[r].u_ref) {
                // the reference's texts (common/exception.hpp:168-186)
                throw InvalidArgument("Conflicting u_ref values detected for voltage regulators " + std::to_string(it->second.first) + ", " +
                                      std::to_string(reg_in_[r].id) + ".");
            }
        } else {
            node_ref[node] = {reg_in_[r].id, reg_st_[r].u_ref};
        }
    }
    for (size_t r = 0; r != reg_in_.size(); ++r) {
        if (reg_st_[r].status && lg_[reg_lg_[r]].type != 0) {
            throw InvalidArgument("Unsupported load_gen type for voltage regulators " + std::to_string(reg_in_[r].id) + ".");
        }
    }
    for (size_t i = 0; i != source_in_.size(); ++i) {
        if (source_st_[i].status && node_ref.count(node_idx_.at(source_in_[i].node)) != 0) {
            throw InvalidArgument("Nodes with a source and a voltage regulated load/generator are not supported when both are enabled. Found at node with id " +
                                  std::to_string(source_in_[i].node));
        }
    }
    if constexpr (B == 3) {
        for (auto const& st : reg_st_)
            if (!std::isnan(st.q_min) || !std::isnan(st.q_max)) {
                throw InvalidArgument("Voltage Regulator with Qmin/Qmax limits for asymmetric calculations is an experimental feature\n");
            }
    }
}

Model::BranchInfo Model::branch_info(Idx b) const {
    if (b < off_aline()) {
        auto const& l = line_in_[b];
        return {l.id, node_seq(l.from_node), node_seq(l.to_node), line_c_[b].base_i, line_c_[b].base_i, -l.i_n};
    }
    if (b < off_link()) {
        Idx const i = b - off_aline();
        auto const& l = aline_in_[i];
        return {l.id, node_seq(l.from_node), node_seq(l.to_node), aline_c_[i].base_i, aline_c_[i].base_i, -l.i_n};
    }
    if (b < off_gb()) { // Link::loading = 0 (link.hpp:30)
        Idx const i = b - off_link();
        auto const& l = link_in_[i];
        return {l.id, node_seq(l.from_node), node_seq(l.to_node), link_base_i_[i][0], link_base_i_[i][1], std::numeric_limits<double>::infinity()};
    }
    if (b < off_trafo()) {
        Idx const i = b - off_gb();
        auto const& g = gb_in_[i];
        return {g.id, node_seq(g.from_node), node_seq(g.to_node), gb_c_[i].base_i_from, gb_c_[i].base_i_to,
                std::isnan(gb_c_[i].sn) ? std::numeric_limits<double>::infinity() : gb_c_[i].sn};
    }
    Idx const i = b - off_trafo();
    auto const& t = trafo_in_[i];
    return {t.id, node_seq(t.from_node), node_seq(t.to_node), trafo_c_[i].base_i_from, trafo_c_[i].base_i_to, trafo_c_[i].sn};
}

Idx Model::node_seq(ID id) const {
    auto it = node_idx_.find(id);
    if (it == node_idx_.end()) throw InvalidArgument("The id cannot be found: " + std::to_string(id) + "\n");
    return it->second;
}

// LoadGen::set_power (load_gen.hpp:86-95): NaN keeps the present value
void Model::set_load_power(Idx i, double const* p, double const* q) {
    double const scalar = lg_[i].direction / (lg_[i].lb == 1 ? kBasePower3p : kBasePower1p);
    for (int k = 0; k != lg_[i].lb; ++k) {
        double ps = lg_st_[i].s[k].real(), qs = lg_st_[i].s[k].imag();
        if (!std::isnan(p[k])) ps = scalar * p[k];
        if (!std::isnan(q[k])) qs = scalar * q[k];
        lg_st_[i].s[k] = cplx{ps, qs};
    }
}

void Model::prepare_topology() {
    if (topo_valid_) return;
    GridGraph g;
    g.n_node = static_cast<Idx>(node_.size());
    for (Idx i = 0; i != n_line(); ++i) {
        g.branch_node.push_back({node_seq(line_in_[i].from_node), node_seq(line_in_[i].to_node)});
        g.branch_status.push_back({static_cast<int8_t>(branch_st_[i].from_status), static_cast<int8_t>(branch_st_[i].to_status)});
        g.branch_shift.push_back(0.0);
    }
    for (Idx b = off_aline(); b != off_trafo(); ++b) {
        BranchInfo const info = branch_info(b);
        g.branch_node.push_back({info.from, info.to});
        g.branch_status.push_back({static_cast<int8_t>(branch_st_[b].from_status), static_cast<int8_t>(branch_st_[b].to_status)});
        g.branch_shift.push_back(b < off_gb() ? 0.0 : gb_c_[b - off_gb()].theta);
    }
    for (Idx i = 0; i != n_trafo(); ++i) {
        auto const& st = branch_st_[off_trafo() + i];
        g.branch_node.push_back({node_seq(trafo_in_[i].from_node), node_seq(trafo_in_[i].to_node)});
        g.branch_status.push_back({static_cast<int8_t>(st.from_status), static_cast<int8_t>(st.to_status)});
        g.branch_shift.push_back(trafo_c_[i].clock * kDeg30);
    }
    for (Idx i = 0; i != n_t3w(); ++i) { // phase_shift = node_k - internal node: {0, -clock_12, -clock_13} * 30 deg
        auto const& c = t3w_c_[i];
        g.branch3_node.push_back({node_seq(c.in.node_1), node_seq(c.in.node_2), node_seq(c.in.node_3)});
        g.branch3_status.push_back({static_cast<int8_t>(t3w_st_[i].status[0]), static_cast<int8_t>(t3w_st_[i].status[1]),
                                    static_cast<int8_t>(t3w_st_[i].status[2])});
        g.branch3_shift.push_back({0.0, -c.clock_12 * kDeg30, -c.clock_13 * kDeg30});
    }
    for (auto const& s : shunt_in_) g.shunt_node.push_back(node_seq(s.node));
    for (size_t i = 0; i != source_in_.size(); ++i) {
        g.source_node.push_back(node_seq(source_in_[i].node));
        g.source_status.push_back(static_cast<int8_t>(source_st_[i].status));
    }
    for (auto const& l : lg_) {
        g.load_gen_node.push_back(l.node);
        g.load_gen_type.push_back(l.type);
    }
    g.regulated_load_gen = reg_lg_;
    topo_ = build_topology(g);
    dev_.reset();
    engines_.clear();
    engines_.resize(topo_.math.size());
    param_valid_[0] = param_valid_[1] = false;
    index_cache_.clear();
    real_cache_.clear();
    topo_valid_ = true;
}

template <int B>
void Model::param_arrays(Idx group, std::vector<double>& bp, std::vector<double>& sp, std::vector<double>& srcp) const {
    constexpr int bb2 = B * B * 2;
    auto const& m = topo_.math[group];
    bp.assign(m.n_branch() * 4 * bb2, 0.0);
    sp.assign(m.n_shunt() * bb2, 0.0);
    srcp.assign(m.n_source() * 4, 0.0);
    for (Idx i = 0; i != n_line(); ++i) {
        Coupling const c = topo_.branch[i];
        if (c.group == group) line_param<B>(line_c_[i], branch_st_[i], &bp[c.pos * 4 * bb2]);
    }
    for (Idx i = 0; i != n_aline(); ++i) {
        Coupling const c = topo_.branch[off_aline() + i];
        if (c.group == group) asym_line_param<B>(aline_c_[i], branch_st_[off_aline() + i], &bp[c.pos * 4 * bb2]);
    }
    for (Idx i = 0; i != n_link(); ++i) {
        Coupling const c = topo_.branch[off_link() + i];
        if (c.group == group) line_param<B>(link_constants(1.0), branch_st_[off_link() + i], &bp[c.pos * 4 * bb2]);
    }
    for (Idx i = 0; i != n_t3w(); ++i) { // main_core/y_bus.hpp:196-204: the three branches of a Branch3
        Coupling3 const& c = topo_.branch3[i];
        if (c.group != group) continue;
        double p3[3 * 4 * bb2];
        three_winding_param<B>(t3w_c_[i], t3w_st_[i], p3);
        for (int k = 0; k != 3; ++k) std::copy_n(p3 + k * 4 * bb2, 4 * bb2, &bp[c.pos[k] * 4 * bb2]);
    }
    for (Idx i = 0; i != n_gb(); ++i) {
        Coupling const c = topo_.branch[off_gb() + i];
        if (c.group != group) continue;
        if constexpr (B == 1) {
            generic_branch_param(gb_c_[i], branch_st_[off_gb() + i], &bp[c.pos * 4 * bb2]);
        } else { // GenericBranch::asym_calc_param throws NotImplementedError (generic_branch.hpp:91)
            throw InvalidArgument("Function not yet implemented: generic_branch in an asymmetric calculation\n");
        }
    }
    for (Idx i = 0; i != n_trafo(); ++i) {
        Coupling const c = topo_.branch[off_trafo() + i];
        if (c.group == group) transformer_param<B>(trafo_c_[i], branch_st_[off_trafo() + i], trafo_st_[i].tap_pos, &bp[c.pos * 4 * bb2]);
    }
    for (size_t i = 0; i != shunt_in_.size(); ++i) {
        Coupling const c = topo_.shunt[i];
        if (c.group == group) shunt_param<B>(shunt_st_[i], &sp[c.pos * bb2]);
    }
    for (size_t i = 0; i != source_in_.size(); ++i) {
        Coupling const c = topo_.source[i];
        if (c.group == group) source_param(source_st_[i], &srcp[c.pos * 4]);
    }
}

template <int B> void Model::prepare_engines() {
    prepare_topology();
    constexpr int si = B == 1 ? 0 : 1;
    for (size_t g = 0; g != topo_.math.size(); ++g) {
        auto& e = engines_[g].engine[si];
        if (e && e->device() != device_) e.reset(); // the caller moved the model to another GPU
        bool const fresh = !e;
        if (fresh) e = std::make_unique<Engine>(topo_.math[g], B == 1, device_);
        if (fresh || !param_valid_[si]) {
            std::vector<double> bp, sp, srcp;
            param_arrays<B>(static_cast<Idx>(g), bp, sp, srcp);
            e->set_param(bp.data(), sp.data(), srcp.data());
        }
    }
    param_valid_[si] = true;
}

// PowerFlowInput of the current state: sinj[g] = [n_lg][B] complex, uref[g] = [n_src] complex (appended)
template <int B>
void Model::gather_pf_input(std::vector<std::vector<double>>& sinj, std::vector<std::vector<double>>& uref,
                            RegulatorInput* reg) const {
    if (reg != nullptr) { // load_gen status per scenario: also what the output step reports as `energized`
        reg->param.resize(topo_.math.size());
        reg->lg_status.resize(topo_.math.size());
        std::vector<size_t> base(topo_.math.size());
        for (size_t g = 0; g != topo_.math.size(); ++g) {
            reg->param[g].assign(topo_.math[g].n_voltage_regulator() * 4, 0.0);
            base[g] = reg->lg_status[g].size();
            reg->lg_status[g].resize(base[g] + topo_.math[g].n_load_gen(), 0);
        }
        for (size_t r = 0; r != reg_in_.size(); ++r) { // VoltageRegulator::calc_param (voltage_regulator.hpp:83-91)
            Coupling const c = topo_.voltage_regulator[r];
            if (c.group == -1) continue;
            double* o = &reg->param[c.group][c.pos * 4];
            o[0] = reg_st_[r].status ? 1.0 : 0.0;
            o[1] = reg_st_[r].u_ref;
            o[2] = reg_st_[r].q_min / kBasePower3p;
            o[3] = reg_st_[r].q_max / kBasePower3p;
        }
        for (size_t i = 0; i != lg_.size(); ++i) {
            Coupling const c = topo_.load_gen[i];
            if (c.group != -1) reg->lg_status[c.group][base[c.group] + c.pos] = lg_st_[i].status ? 1 : 0;
        }
    }
    for (size_t g = 0; g != topo_.math.size(); ++g) {
        size_t const off_s = sinj[g].size(), off_u = uref[g].size();
        sinj[g].resize(off_s + topo_.math[g].n_load_gen() * B * 2, 0.0);
        uref[g].resize(off_u + topo_.math[g].n_source() * 2, 0.0);
    }
    std::vector<size_t> base_s(topo_.math.size()), base_u(topo_.math.size());
    for (size_t g = 0; g != topo_.math.size(); ++g) {
        base_s[g] = sinj[g].size() - topo_.math[g].n_load_gen() * B * 2;
        base_u[g] = uref[g].size() - topo_.math[g].n_source() * 2;
    }
    for (size_t i = 0; i != lg_.size(); ++i) {
        Coupling const c = topo_.load_gen[i];
        if (c.group == -1) continue;
        cplx v[B];
        load_gen_injection<B>(lg_st_[i], lg_[i].lb, v);
        double* o = &sinj[c.group][base_s[c.group] + c.pos * B * 2];
        for (int p = 0; p != B; ++p) {
            o[2 * p] = v[p].real();
            o[2 * p + 1] = v[p].imag();
        }
    }
    for (size_t i = 0; i != source_in_.size(); ++i) {
        Coupling const c = topo_.source[i];
        if (c.group == -1) continue;
        cplx const u = source_u_ref(source_st_[i]);
        uref[c.group][base_u[c.group] + c.pos * 2] = u.real();
        uref[c.group][base_u[c.group] + c.pos * 2 + 1] = u.imag();
    }
}

std::string scenario_failure_text(int32_t status, int64_t max_iter, double max_dev, double err_tol) {
    auto shortest = [](double v) { // std::format("{}", double)
        char buf[64];
        auto const r = std::to_chars(buf, buf + sizeof(buf), v);
        return std::string(buf, r.ptr);
    };
    if (status == 1) {
        return "Iteration failed to converge after " + std::to_string(max_iter) + " iterations! Max deviation: " + shortest(max_dev) +
               ", error tolerance: " + shortest(err_tol) + ".\n";
    }
    if (status == 4) return "Unallocated Q remains after distribution on a regulated bus";
    return "Sparse matrix error, possibly singular matrix!\n"
           "If you get this error from state estimation, "
           "it might mean the system is not fully observable, i.e. not enough measurements.\n"
           "It might also mean that you are running into a corner case where PGM cannot resolve yet.\n"
           "See https://github.com/PowerGridModel/power-grid-model/issues/864.";
}

void Model::mark(bool topo, bool param, Saved* saved) {
    if (topo) {
        topo_valid_ = false;
        if (saved != nullptr) saved->topo = true;
    }
    if (topo || param) {
        param_valid_[0] = param_valid_[1] = false;
        if (saved != nullptr) saved->param = true;
    }
}

void Model::apply_scenario(UpdateData const& u, Idx s, Saved* saved) {
    auto find = [](auto const& upd, Idx pos, Idx n_in_scenario, Idx n_comp, std::unordered_map<ID, Idx> const& map, Idx offset) {
        if (upd.id == kNaID) {
            if (n_in_scenario != n_comp) throw InvalidArgument("update without ids must cover every element of the component");
            return offset + pos;
        }
        auto it = map.find(upd.id);
        if (it == map.end()) throw InvalidArgument("The id cannot be found: " + std::to_string(upd.id) + "\n");
        return it->second;
    };
    auto set_status = [](bool& st, IntS v) {
        if (v == kNaIntS || static_cast<bool>(v) == st) return false;
        st = static_cast<bool>(v);
        return true;
    };
    {
        auto [b, e] = scenario_span<BranchUpdate>(u.line, s);
        for (auto p = b; p != e; ++p) {
            Idx const i = find(*p, p - b, e - b, n_line(), line_idx_, 0);
            if (saved != nullptr) saved->branch.emplace_back(i, branch_st_[i]);
            bool changed = set_status(branch_st_[i].from_status, p->from_status);
            changed = set_status(branch_st_[i].to_status, p->to_status) || changed;
            mark(changed, changed, saved);
        }
    }
    auto upd_plain_branch = [&](ComponentBuffer const& buf, Idx count, std::unordered_map<ID, Idx> const& map, Idx offset) {
        auto [b, e] = scenario_span<BranchUpdate>(buf, s);
        for (auto p = b; p != e; ++p) {
            Idx const i = offset + find(*p, p - b, e - b, count, map, 0);
            if (saved != nullptr) saved->branch.emplace_back(i, branch_st_[i]);
            bool changed = set_status(branch_st_[i].from_status, p->from_status);
            changed = set_status(branch_st_[i].to_status, p->to_status) || changed;
            mark(changed, changed, saved);
        }
    };
    upd_plain_branch(u.asym_line, n_aline(), aline_idx_, off_aline());
    upd_plain_branch(u.link, n_link(), link_idx_, off_link());
    upd_plain_branch(u.generic_branch, n_gb(), gb_idx_, off_gb());
    {
        // ThreeWindingTransformer::update (three_winding_transformer.hpp:158-163)
        auto [b, e] = scenario_span<ThreeWindingTransformerUpdate>(u.three_winding_transformer, s);
        for (auto p = b; p != e; ++p) {
            Idx const i = find(*p, p - b, e - b, n_t3w(), t3w_idx_, 0);
            if (saved != nullptr) saved->t3w.emplace_back(i, t3w_st_[i]);
            bool topo = set_status(t3w_st_[i].status[0], p->status_1);
            topo = set_status(t3w_st_[i].status[1], p->status_2) || topo;
            topo = set_status(t3w_st_[i].status[2], p->status_3) || topo;
            bool tap = false;
            if (p->tap_pos != kNaIntS && p->tap_pos != t3w_st_[i].tap_pos) {
                t3w_st_[i].tap_pos = tap_limit(t3w_c_[i], p->tap_pos);
                tap = true;
            }
            mark(topo, tap || topo, saved);
        }
    }
    {
        auto [b, e] = scenario_span<TransformerUpdate>(u.transformer, s);
        for (auto p = b; p != e; ++p) {
            Idx const i = find(*p, p - b, e - b, n_trafo(), trafo_idx_, 0);
            Idx const bi = off_trafo() + i;
            if (saved != nullptr) {
                saved->branch.emplace_back(bi, branch_st_[bi]);
                saved->trafo.emplace_back(i, trafo_st_[i]);
            }
            bool topo = set_status(branch_st_[bi].from_status, p->from_status);
            topo = set_status(branch_st_[bi].to_status, p->to_status) || topo;
            bool tap = false;
            if (p->tap_pos != kNaIntS && p->tap_pos != trafo_st_[i].tap_pos) {
                trafo_st_[i].tap_pos = tap_limit(trafo_c_[i], p->tap_pos);
                tap = true;
            }
            mark(topo, tap || topo, saved);
        }
    }
    {
        auto [b, e] = scenario_span<ShuntUpdate>(u.shunt, s);
        for (auto p = b; p != e; ++p) {
            Idx const i = find(*p, p - b, e - b, static_cast<Idx>(shunt_in_.size()), shunt_idx_, 0);
            if (saved != nullptr) saved->shunt.emplace_back(i, shunt_st_[i]);
            bool changed = set_status(shunt_st_[i].status, p->status);
            changed = shunt_set(shunt_st_[i], shunt_base_y_[i], p->g1, p->b1, p->g0, p->b0) || changed;
            mark(false, changed, saved);
        }
    }
    {
        auto [b, e] = scenario_span<SourceUpdate>(u.source, s);
        for (auto p = b; p != e; ++p) {
            Idx const i = find(*p, p - b, e - b, static_cast<Idx>(source_in_.size()), source_idx_, 0);
            if (saved != nullptr) saved->source.emplace_back(i, source_st_[i]);
            auto& st = source_st_[i];
            bool const topo = set_status(st.status, p->status);
            if (!std::isnan(p->u_ref)) st.u_ref = p->u_ref;
            if (!std::isnan(p->u_ref_angle)) st.u_ref_angle = p->u_ref_angle;
            bool param = false;
            if (!std::isnan(p->sk)) st.sk = p->sk, param = true;
            if (!std::isnan(p->rx_ratio)) st.rx_ratio = p->rx_ratio, param = true;
            if (!std::isnan(p->z01_ratio)) st.z01_ratio = p->z01_ratio, param = true;
            mark(topo, param || topo, saved);
        }
    }
    auto upd_lg = [&](auto tag, ComponentBuffer const& buf, Idx offset, Idx count) {
        using U = decltype(tag);
        auto [b, e] = scenario_span<U>(buf, s);
        for (auto p = b; p != e; ++p) {
            Idx const i = find(*p, p - b, e - b, count, lg_idx_, offset);
            if (i < offset || i >= offset + count) throw InvalidArgument("The id cannot be found: " + std::to_string(p->id) + "\n");
            if (saved != nullptr) saved->lg.emplace_back(i, lg_st_[i]);
            set_status(lg_st_[i].status, p->status);
            if constexpr (std::is_same_v<U, SymLoadGenUpdate>) {
                set_load_power(i, &p->p_specified, &p->q_specified);
            } else {
                set_load_power(i, p->p_specified, p->q_specified);
            }
        }
    };
    upd_lg(SymLoadGenUpdate{}, u.sym_gen, 0, n_sym_gen_);
    upd_lg(AsymLoadGenUpdate{}, u.asym_gen, n_sym_gen_, n_asym_gen_);
    upd_lg(SymLoadGenUpdate{}, u.sym_load, n_sym_gen_ + n_asym_gen_, n_sym_load_);
    upd_lg(AsymLoadGenUpdate{}, u.asym_load, n_sym_gen_ + n_asym_gen_ + n_sym_load_, n_asym_load_);
    {
        // VoltageRegulator::update (voltage_regulator.hpp:35-41): neither topology nor parameters change
        auto [b, e] = scenario_span<VoltageRegulatorUpdate>(u.voltage_regulator, s);
        for (auto p = b; p != e; ++p) {
            Idx const i = find(*p, p - b, e - b, static_cast<Idx>(reg_in_.size()), reg_idx_, 0);
            if (saved != nullptr) saved->reg.emplace_back(i, reg_st_[i]);
            set_status(reg_st_[i].status, p->status);
            if (!std::isnan(p->u_ref)) reg_st_[i].u_ref = p->u_ref;
            if (!std::isnan(p->q_min)) reg_st_[i].q_min = p->q_min;
            if (!std::isnan(p->q_max)) reg_st_[i].q_max = p->q_max;
        }
    }
    {
        // TransformerTapRegulator::update (transformer_tap_regulator.hpp:41-50): neither topology nor parameters change
        auto [b, e] = scenario_span<TransformerTapRegulatorUpdate>(u.transformer_tap_regulator, s);
        for (auto p = b; p != e; ++p) {
            Idx const i = find(*p, p - b, e - b, static_cast<Idx>(tap_reg_in_.size()), tap_reg_idx_, 0);
            if (saved != nullptr) saved->tap_reg.emplace_back(i, tap_reg_st_[i]);
            // Regulator::set_status takes the value as it is (regulator.hpp:31): "not given" reads as on
            tap_reg_st_[i].status = static_cast<bool>(p->status);
            if (!std::isnan(p->u_set)) tap_reg_st_[i].u_set = p->u_set;
            if (!std::isnan(p->u_band)) tap_reg_st_[i].u_band = p->u_band;
            if (!std::isnan(p->line_drop_compensation_r)) tap_reg_st_[i].line_drop_compensation_r = p->line_drop_compensation_r;
            if (!std::isnan(p->line_drop_compensation_x)) tap_reg_st_[i].line_drop_compensation_x = p->line_drop_compensation_x;
        }
    }
}

void Model::restore(Saved const& s) {
    for (auto it = s.branch.rbegin(); it != s.branch.rend(); ++it) branch_st_[it->first] = it->second;
    for (auto it = s.trafo.rbegin(); it != s.trafo.rend(); ++it) trafo_st_[it->first] = it->second;
    for (auto it = s.source.rbegin(); it != s.source.rend(); ++it) source_st_[it->first] = it->second;
    for (auto it = s.shunt.rbegin(); it != s.shunt.rend(); ++it) shunt_st_[it->first] = it->second;
    for (auto it = s.lg.rbegin(); it != s.lg.rend(); ++it) lg_st_[it->first] = it->second;
    for (auto it = s.tap_reg.rbegin(); it != s.tap_reg.rend(); ++it) tap_reg_st_[it->first] = it->second;
    for (auto it = s.reg.rbegin(); it != s.reg.rend(); ++it) reg_st_[it->first] = it->second;
    for (auto it = s.t3w.rbegin(); it != s.t3w.rend(); ++it) t3w_st_[it->first] = it->second;
    if (s.topo) topo_valid_ = false;
    if (s.param) param_valid_[0] = param_valid_[1] = false;
}

// Permanent update: the reference resolves every id before it changes anything (get_all_sequence_idx_map,
// main_model_impl.hpp:191-197), so an unknown id leaves the model untouched.  Here the components are applied one by one with
// their previous state recorded, and a failure rolls everything back before it is reported.
void Model::update_permanent(UpdateData const& update) {
    Saved saved;
    bool const topo_was_valid = topo_valid_;
    bool const param_was_valid[2] = {param_valid_[0], param_valid_[1]};
    try {
        apply_scenario(update, 0, &saved);
    } catch (...) {
        restore(saved);
        // the state is the old one again and nothing was rebuilt in between: the caches are as valid as they were
        topo_valid_ = topo_was_valid;
        param_valid_[0] = param_was_valid[0];
        param_valid_[1] = param_was_valid[1];
        throw;
    }
    ++state_version_;
}

void Model::batch_pf_input(UpdateData const& update, bool symmetric, Idx group, double* s_injection, double* source_u_ref) {
    prepare_topology();
    if (group < 0 || group >= static_cast<Idx>(topo_.math.size())) throw InvalidArgument("math group out of range");
    std::vector<std::vector<double>> sinj(topo_.math.size()), uref(topo_.math.size());
    for (Idx s = 0; s != update.n_scenarios; ++s) {
        Saved saved;
        apply_scenario(update, s, &saved);
        bool const structural = saved.topo || saved.param;
        if (!structural) {
            if (symmetric) {
                gather_pf_input<1>(sinj, uref);
            } else {
                gather_pf_input<3>(sinj, uref);
            }
        }
        restore(saved);
        if (structural) throw InvalidArgument("batch changes topology or parameters: no common math model");
    }
    std::memcpy(s_injection, sinj[group].data(), sinj[group].size() * sizeof(double));
    std::memcpy(source_u_ref, uref[group].data(), uref[group].size() * sizeof(double));
}

// ---- output conversion (host, v1) ------------------------------------------------------------------------------------
// so[0..5] = u, bus_injection(unused), branch, source, shunt, load_gen per group, scenario-major
template <int B>
void Model::write_output(Idx n_scn, Idx first, OutputData const& out, std::vector<std::vector<double>> const (&so)[6],
                         std::vector<std::vector<int8_t>> const& reg_out,
                         std::vector<std::vector<int8_t>> const* lg_status) const {
    constexpr int c2 = 2 * B;
    constexpr double base_power = B == 1 ? kBasePower3p : kBasePower1p;
    constexpr double u_scale = B == 1 ? 1.0 : 1.0 / kSqrt3;
    auto cabs = [](double re, double im) { return std::sqrt(re * re + im * im); };
    Idx const nn = static_cast<Idx>(node_.size());
    auto group_n = [&](Idx g, int what) -> Idx {
        auto const& m = topo_.math[g];
        switch (what) {
        case 0: return m.n_bus;
        case 2: return m.n_branch();
        case 3: return m.n_source();
        case 4: return m.n_shunt();
        default: return m.n_load_gen();
        }
    };
    for (Idx s = 0; s != n_scn; ++s) {
        Idx const os = first + s;
        auto appliance = [&](Coupling c, int what, double base_i, double direction, ID id, bool status) {
            ApplianceOutput<B> o{};
            o.id = id;
            if (c.group == -1) return o;
            o.energized = status ? 1 : 0;
            double const* v = &so[what][c.group][(s * group_n(c.group, what) + c.pos) * 2 * c2];
            for (int p = 0; p != B; ++p) {
                double const sr = v[2 * p], sim = v[2 * p + 1], ir = v[c2 + 2 * p], ii = v[c2 + 2 * p + 1];
                o.p[p] = base_power * sr * direction;
                o.q[p] = base_power * sim * direction;
                o.s[p] = base_power * cabs(sr, sim);
                o.i[p] = base_i * cabs(ir, ii);
                o.pf[p] = o.s[p] < kNumTol ? 0.0 : o.p[p] / o.s[p];
            }
            return o;
        };
        if (out.node != nullptr) {
            std::vector<cplx> inj(nn * B, cplx{});
            auto add = [&](Idx node, Coupling c, int what) {
                if (c.group == -1) return;
                double const* v = &so[what][c.group][(s * group_n(c.group, what) + c.pos) * 2 * c2];
                for (int p = 0; p != B; ++p) inj[node * B + p] += cplx{v[2 * p], v[2 * p + 1]};
            };
            for (size_t i = 0; i != source_in_.size(); ++i) add(node_idx_.at(source_in_[i].node), topo_.source[i], 3);
            Idx const o_sg = 0, o_ag = n_sym_gen_, o_sl = n_sym_gen_ + n_asym_gen_, o_al = o_sl + n_sym_load_;
            for (Idx i = o_sl; i != o_sl + n_sym_load_; ++i) add(lg_[i].node, topo_.load_gen[i], 5);
            for (Idx i = o_sg; i != o_sg + n_sym_gen_; ++i) add(lg_[i].node, topo_.load_gen[i], 5);
            for (Idx i = o_al; i != o_al + n_asym_load_; ++i) add(lg_[i].node, topo_.load_gen[i], 5);
            for (Idx i = o_ag; i != o_ag + n_asym_gen_; ++i) add(lg_[i].node, topo_.load_gen[i], 5);
            auto* dst = static_cast<NodeOutput<B>*>(out.node) + os * nn;
            for (Idx i = 0; i != nn; ++i) {
                NodeOutput<B> o{};
                o.id = node_[i].id;
                Coupling const c = topo_.node[i];
                if (c.group != -1) {
                    o.energized = 1;
                    double const* u = &so[0][c.group][(s * topo_.math[c.group].n_bus + c.pos) * c2];
                    for (int p = 0; p != B; ++p) {
                        o.u_pu[p] = cabs(u[2 * p], u[2 * p + 1]);
                        o.u[p] = u_scale * node_[i].u_rated * o.u_pu[p];
                        o.u_angle[p] = std::atan2(u[2 * p + 1], u[2 * p]);
                        o.p[p] = base_power * inj[i * B + p].real();
                        o.q[p] = base_power * inj[i * B + p].imag();
                    }
                }
                dst[i] = o;
            }
        }
        auto branch = [&](Idx seq, ID id, double base_i_from, double base_i_to, double sn, double i_n) {
            BranchOutput<B> o{};
            o.id = id;
            Coupling const c = topo_.branch[seq];
            if (c.group == -1) return o;
            o.energized = (branch_st_[seq].from_status || branch_st_[seq].to_status) ? 1 : 0;
            double const* v = &so[2][c.group][(s * topo_.math[c.group].n_branch() + c.pos) * 4 * c2];
            double sum_sf = 0, sum_st = 0, max_if = 0, max_it = 0;
            for (int p = 0; p != B; ++p) {
                double const* sf = v + 2 * p;
                double const* st = v + c2 + 2 * p;
                double const* i_f = v + 2 * c2 + 2 * p;
                double const* i_t = v + 3 * c2 + 2 * p;
                o.p_from[p] = base_power * sf[0];
                o.q_from[p] = base_power * sf[1];
                o.i_from[p] = base_i_from * cabs(i_f[0], i_f[1]);
                o.s_from[p] = base_power * cabs(sf[0], sf[1]);
                o.p_to[p] = base_power * st[0];
                o.q_to[p] = base_power * st[1];
                o.i_to[p] = base_i_to * cabs(i_t[0], i_t[1]);
                o.s_to[p] = base_power * cabs(st[0], st[1]);
                sum_sf = p == 0 ? o.s_from[p] : sum_sf + o.s_from[p];
                sum_st = p == 0 ? o.s_to[p] : sum_st + o.s_to[p];
                max_if = p == 0 ? o.i_from[p] : std::max(max_if, o.i_from[p]);
                max_it = p == 0 ? o.i_to[p] : std::max(max_it, o.i_to[p]);
            }
            o.loading = sn > 0.0 ? std::max(sum_sf, sum_st) / sn : std::max(max_if, max_it) / i_n;
            return o;
        };
        if (out.line != nullptr) {
            auto* dst = static_cast<BranchOutput<B>*>(out.line) + os * n_line();
            for (Idx i = 0; i != n_line(); ++i) dst[i] = branch(i, line_in_[i].id, line_c_[i].base_i, line_c_[i].base_i, -1.0, line_in_[i].i_n);
        }
        auto branch_range = [&](void* base, Idx first_b, Idx count) {
            if (base == nullptr) return;
            auto* dst = static_cast<BranchOutput<B>*>(base) + os * count;
            for (Idx i = 0; i != count; ++i) {
                BranchInfo const info = branch_info(first_b + i);
                dst[i] = branch(first_b + i, info.id, info.base_i_from, info.base_i_to, info.rating > 0.0 ? info.rating : -1.0,
                                info.rating > 0.0 ? 0.0 : -info.rating);
            }
        };
        branch_range(out.asym_line, off_aline(), n_aline());
        branch_range(out.link, off_link(), n_link());
        branch_range(out.generic_branch, off_gb(), n_gb());
        if (out.three_winding_transformer != nullptr) { // Branch3::get_output (branch3.hpp:93-122)
            auto* dst = static_cast<Branch3Output<B>*>(out.three_winding_transformer) + os * n_t3w();
            for (Idx i = 0; i != n_t3w(); ++i) {
                Branch3Output<B> o{};
                auto const& c = t3w_c_[i];
                o.id = c.in.id;
                Coupling3 const& cp = topo_.branch3[i];
                if (cp.group != -1) {
                    o.energized = (t3w_st_[i].status[0] || t3w_st_[i].status[1] || t3w_st_[i].status[2]) ? 1 : 0;
                    double* const fields[3][4] = {{o.p_1, o.q_1, o.i_1, o.s_1}, {o.p_2, o.q_2, o.i_2, o.s_2}, {o.p_3, o.q_3, o.i_3, o.s_3}};
                    double const sn[3] = {c.in.sn_1, c.in.sn_2, c.in.sn_3};
                    double loading[3];
                    for (int k = 0; k != 3; ++k) {
                        double const* v = &so[2][cp.group][(s * topo_.math[cp.group].n_branch() + cp.pos[k]) * 4 * c2];
                        double sum_s = 0.0;
                        for (int p = 0; p != B; ++p) {
                            double const* sf = v + 2 * p;
                            double const* i_f = v + 2 * c2 + 2 * p;
                            fields[k][0][p] = base_power * sf[0];
                            fields[k][1][p] = base_power * sf[1];
                            fields[k][2][p] = c.base_i[k] * cabs(i_f[0], i_f[1]);
                            fields[k][3][p] = base_power * cabs(sf[0], sf[1]);
                            sum_s = p == 0 ? fields[k][3][p] : sum_s + fields[k][3][p];
                        }
                        loading[k] = sum_s / sn[k];
                    }
                    o.loading_1 = loading[0];
                    o.loading_2 = loading[1];
                    o.loading_3 = loading[2];
                    o.loading = std::max({loading[0], loading[1], loading[2]});
                }
                dst[i] = o;
            }
        }
        branch_range(out.transformer, off_trafo(), n_trafo());
        if (out.shunt != nullptr) {
            Idx const n = static_cast<Idx>(shunt_in_.size());
            auto* dst = static_cast<ApplianceOutput<B>*>(out.shunt) + os * n;
            for (Idx i = 0; i != n; ++i) {
                double const u = node_[node_idx_.at(shunt_in_[i].node)].u_rated;
                dst[i] = appliance(topo_.shunt[i], 4, kBasePower3p / u / kSqrt3, -1.0, shunt_in_[i].id, shunt_st_[i].status);
            }
        }
        if (out.source != nullptr) {
            Idx const n = static_cast<Idx>(source_in_.size());
            auto* dst = static_cast<ApplianceOutput<B>*>(out.source) + os * n;
            for (Idx i = 0; i != n; ++i) {
                double const u = node_[node_idx_.at(source_in_[i].node)].u_rated;
                dst[i] = appliance(topo_.source[i], 3, kBasePower3p / u / kSqrt3, 1.0, source_in_[i].id, source_st_[i].status);
            }
        }
        auto lg_out = [&](void* base, Idx begin, Idx count) {
            if (base == nullptr) return;
            auto* dst = static_cast<ApplianceOutput<B>*>(base) + os * count;
            for (Idx k = 0; k != count; ++k) {
                Idx const i = begin + k;
                Coupling const c = topo_.load_gen[i];
                bool status = lg_st_[i].status;
                if (lg_status != nullptr && c.group != -1) status = (*lg_status)[c.group][s * topo_.math[c.group].n_load_gen() + c.pos] != 0;
                dst[k] = appliance(c, 5, lg_[i].base_i, lg_[i].direction, lg_[i].id, status);
            }
        };
        lg_out(out.sym_gen, 0, n_sym_gen_);
        lg_out(out.asym_gen, n_sym_gen_, n_asym_gen_);
        lg_out(out.sym_load, n_sym_gen_ + n_asym_gen_, n_sym_load_);
        lg_out(out.asym_load, n_sym_gen_ + n_asym_gen_ + n_sym_load_, n_asym_load_);
        if (out.transformer_tap_regulator != nullptr) { // main_core/output.hpp:381-396: null output unless the optimizer ran
            Idx const n = static_cast<Idx>(tap_reg_in_.size());
            auto* dst = static_cast<TransformerTapRegulatorOutput*>(out.transformer_tap_regulator) + os * n;
            for (Idx i = 0; i != n; ++i) {
                IntS const tap = i < static_cast<Idx>(tap_positions_out_.size()) ? tap_positions_out_[i] : kNaIntS;
                std::memset(&dst[i], 0, sizeof(dst[i])); // padding bytes too: outputs compare byte for byte between runs
                dst[i].id = tap_reg_in_[i].id;
                dst[i].energized = tap != kNaIntS ? 1 : 0;
                dst[i].tap_pos = tap;
            }
        }
        if (out.voltage_regulator != nullptr) { // main_core/output.hpp:407-421, VoltageRegulator::get_output
            Idx const n = static_cast<Idx>(reg_in_.size());
            auto* dst = static_cast<VoltageRegulatorOutput*>(out.voltage_regulator) + os * n;
            for (Idx i = 0; i != n; ++i) {
                Coupling const c = topo_.voltage_regulator[i];
                VoltageRegulatorOutput o{reg_in_[i].id, 0, 0};
                if (c.group != -1) {
                    int8_t const* v = &reg_out[c.group][(s * topo_.math[c.group].n_voltage_regulator() + c.pos) * 2];
                    o.energized = (reg_st_[i].status && v[1] != 0) ? 1 : 0;
                    o.limit_violated = v[0];
                }
                dst[i] = o;
            }
        }
    }
}

template <int B>
void Model::solve_block(ModelOptions const& opt, Idx n_scn, std::vector<std::vector<double>> const& sinj,
                        std::vector<std::vector<double>> const& uref, RegulatorInput const* reg, BlockSolution& sol) {
    constexpr int si = B == 1 ? 0 : 1;
    constexpr int c2 = 2 * B;
    auto& so = sol.so;
    for (auto& v : so) v.resize(topo_.math.size());
    sol.reg_out.assign(topo_.math.size(), {});
    sol.status.assign(n_scn, 0);
    sol.n_iter.assign(n_scn, 0);
    sol.max_dev.assign(n_scn, 0.0);
    for (size_t g = 0; g != topo_.math.size(); ++g) {
        auto const& m = topo_.math[g];
        Engine& e = *engines_[g].engine[si];
        so[0][g].resize(n_scn * m.n_bus * c2);
        so[2][g].resize(n_scn * m.n_branch() * 4 * c2);
        so[3][g].resize(n_scn * m.n_source() * 2 * c2);
        so[4][g].resize(n_scn * m.n_shunt() * 2 * c2);
        so[5][g].resize(n_scn * m.n_load_gen() * 2 * c2);
        std::vector<int32_t> st(n_scn), it(n_scn);
        std::vector<double> dev(n_scn);
        SolverOutputView view{so[0][g].data(), nullptr, so[2][g].data(), so[3][g].data(), so[4][g].data(), so[5][g].data(),
                              st.data(), it.data(), dev.data()};
        PfInputView in_view{n_scn, uref[g].data(), false, sinj[g].data()};
        if (e.has_regulators()) {
            if (reg == nullptr) throw InvalidArgument("internal: regulator input missing");
            sol.reg_out[g].resize(n_scn * m.n_voltage_regulator() * 2);
            view.voltage_regulator = sol.reg_out[g].data();
            in_view.voltage_regulator = reg->param[g].data();
            in_view.load_gen_status = reg->lg_status[g].data();
        }
        auto t0 = Clock::now();
        e.stage(in_view);
        timing[1] += ms_since(t0);
        timing[2] += e.solve_staged({opt.method, opt.err_tol, static_cast<int32_t>(opt.max_iter)});
        t0 = Clock::now();
        e.fetch(view);
        timing[4] += ms_since(t0);
        for (Idx s = 0; s != n_scn; ++s) {
            if (sol.status[s] == 0 && st[s] != 0) sol.max_dev[s] = dev[s];
            if (sol.status[s] == 0) sol.status[s] = st[s];
            sol.n_iter[s] = std::max(sol.n_iter[s], it[s]);
        }
    }
}

template <int B>
int64_t Model::run_block(ModelOptions const& opt, Idx n_scn, std::vector<std::vector<double>> const& sinj,
                         std::vector<std::vector<double>> const& uref, OutputData const& out, Idx first, int32_t* n_iter,
                         int32_t* status, RegulatorInput const* reg) {
    BlockSolution sol;
    solve_block<B>(opt, n_scn, sinj, uref, reg, sol);
    auto t0 = Clock::now();
    write_output<B>(n_scn, first, out, sol.so, sol.reg_out, reg != nullptr ? &reg->lg_status : nullptr);
    timing[3] += ms_since(t0);
    int64_t failed = 0;
    for (Idx s = 0; s != n_scn; ++s) {
        if (n_iter != nullptr) n_iter[first + s] = sol.n_iter[s];
        if (status != nullptr) status[first + s] = sol.status[s];
        if (sol.status[s] != 0) {
            ++failed;
            batch_message += "Error in batch #" + std::to_string(first + s) + ": " +
                             scenario_failure_text(sol.status[s], opt.max_iter, sol.max_dev[s], opt.err_tol) + "\n";
        }
    }
    return failed;
}

// bridges of the graph of fully connected branches (iterative Tarjan; parallel branches are told apart by their edge id)
Model::BridgeInfo Model::bridge_analysis() const {
    Idx const n = static_cast<Idx>(node_.size());
    Idx const nb = n_branch_comp();
    auto ends = [&](Idx b) -> std::pair<Idx, Idx> {
        BranchInfo const info = branch_info(b);
        return {info.from, info.to};
    };
    std::vector<Idx> ptr(n + 1, 0);
    std::vector<std::pair<Idx, Idx>> e(nb);
    for (Idx b = 0; b != nb; ++b) {
        e[b] = ends(b);
        if (!branch_st_[b].from_status || !branch_st_[b].to_status || e[b].first == e[b].second) continue;
        ++ptr[e[b].first + 1];
        ++ptr[e[b].second + 1];
    }
    for (Idx i = 0; i != n; ++i) ptr[i + 1] += ptr[i];
    std::vector<Idx> adj_node(ptr[n]), adj_edge(ptr[n]), cur(ptr.begin(), ptr.end() - 1);
    for (Idx b = 0; b != nb; ++b) {
        if (!branch_st_[b].from_status || !branch_st_[b].to_status || e[b].first == e[b].second) continue;
        adj_node[cur[e[b].first]] = e[b].second;
        adj_edge[cur[e[b].first]++] = b;
        adj_node[cur[e[b].second]] = e[b].first;
        adj_edge[cur[e[b].second]++] = b;
    }
    BridgeInfo info;
    info.bridge.assign(nb, 0);
    info.child.assign(nb, -1);
    info.size.assign(n, 1);
    info.n_source.assign(n, 0);
    info.root.assign(n, -1);
    info.order.assign(n, -1);
    for (size_t i = 0; i != source_in_.size(); ++i)
        if (source_st_[i].status) ++info.n_source[node_idx_.at(source_in_[i].node)];
    info.self_source = info.n_source;
    std::vector<char>& bridge = info.bridge;
    std::vector<Idx>& disc = info.disc;
    disc.assign(n, -1);
    std::vector<Idx> low(n, 0), parent_edge(n, -1), it(ptr.begin(), ptr.end() - 1), stack;
    Idx timer = 0;
    for (Idx root = 0; root != n; ++root) {
        if (disc[root] != -1) continue;
        info.order[timer] = root;
        disc[root] = low[root] = timer++;
        info.root[root] = root;
        stack.push_back(root);
        while (!stack.empty()) {
            Idx const v = stack.back();
            if (it[v] != ptr[v + 1]) {
                Idx const w = adj_node[it[v]], eid = adj_edge[it[v]];
                ++it[v];
                if (eid == parent_edge[v]) continue;
                if (disc[w] == -1) {
                    info.order[timer] = w;
                    disc[w] = low[w] = timer++;
                    info.root[w] = root;
                    parent_edge[w] = eid;
                    stack.push_back(w);
                } else {
                    low[v] = std::min(low[v], disc[w]);
                }
            } else {
                stack.pop_back();
                if (!stack.empty()) {
                    Idx const p = stack.back();
                    low[p] = std::min(low[p], low[v]);
                    info.size[p] += info.size[v];
                    info.n_source[p] += info.n_source[v];
                    if (low[v] > disc[p]) {
                        bridge[parent_edge[v]] = 1;
                        info.child[parent_edge[v]] = v;
                    }
                }
            }
        }
    }
    info.adj_ptr = std::move(ptr);
    info.adj_node = std::move(adj_node);
    info.adj_edge = std::move(adj_edge);
    return info;
}

std::vector<Idx> Model::closing_branches(UpdateData const& u) {
    std::vector<Idx> found;
    if (u.line.data == nullptr && u.transformer.data == nullptr) return found;
    prepare_topology();
    if (topo_.math.size() != 1) return found;
    auto consider = [&](Idx bi, IntS from, IntS to) {
        bool const closes = (from != kNaIntS && from != 0 && !branch_st_[bi].from_status) || (to != kNaIntS && to != 0 && !branch_st_[bi].to_status);
        if (!closes || std::find(found.begin(), found.end(), bi) != found.end()) return;
        BranchInfo const info = branch_info(bi);
        // both ends in the base grid's math model: the copy's buses are the base grid's
        if (info.from == info.to || topo_.node[info.from].group != 0 || topo_.node[info.to].group != 0) return;
        found.push_back(bi);
    };
    for (Idx s = 0; s != u.n_scenarios; ++s) {
        {
            auto [b, e] = scenario_span<BranchUpdate>(u.line, s);
            for (auto p = b; p != e; ++p) {
                Idx bi = -1;
                if (p->id == kNaID) {
                    if (e - b == n_line()) bi = p - b;
                } else if (auto it = line_idx_.find(p->id); it != line_idx_.end()) {
                    bi = it->second;
                }
                if (bi >= 0) consider(bi, p->from_status, p->to_status);
            }
        }
        {
            auto [b, e] = scenario_span<TransformerUpdate>(u.transformer, s);
            for (auto p = b; p != e; ++p) {
                Idx ti = -1;
                if (p->id == kNaID) {
                    if (e - b == n_trafo()) ti = p - b;
                } else if (auto it = trafo_idx_.find(p->id); it != trafo_idx_.end()) {
                    ti = it->second;
                }
                if (ti >= 0) consider(off_trafo() + ti, p->from_status, p->to_status);
            }
        }
        if (static_cast<int>(found.size()) > kMaxOutageSlots) return {};
    }
    std::sort(found.begin(), found.end());
    return found;
}

Model* Model::outage_host(UpdateData const& update) {
    if (std::getenv("PGMB_NO_UNION_GRID") != nullptr) return this;
    std::vector<Idx> const closing = closing_branches(update);
    if (closing.empty()) return this;
    Model* const host = union_model(closing);
    host->device_ = device_;
    return host;
}

void Model::outage_plan_summary(UpdateData const& update, bool symmetric, int64_t* out) {
    Model* const host = outage_host(update);
    host->prepare_topology();
    OutagePlan plan;
    bool const planned = symmetric ? host->plan_outage_batch<1>(update, plan) : host->plan_outage_batch<3>(update, plan);
    Idx const n = update.n_scenarios;
    for (Idx s = 0; s != n; ++s) {
        int64_t* o = out + 4 * s;
        o[0] = 1; // own topology
        o[1] = o[2] = 0;
        o[3] = host != this ? 1 : 0;
    }
    if (!planned) return;
    Idx const n_bus = host->topo_.math[0].n_bus;
    for (Idx s = 0; s != n; ++s) {
        int64_t* o = out + 4 * s;
        o[0] = 0;
        for (int j = 0; j != plan.n_slot; ++j) o[1] += plan.math_branch[s * plan.n_slot + j] >= 0 ? 1 : 0;
        if (plan.dead_off[s] >= 0) {
            uint8_t const* mask = &plan.dead[static_cast<size_t>(plan.dead_off[s]) * n_bus];
            o[2] = std::count(mask, mask + n_bus, uint8_t{1});
        }
    }
    for (Idx const s : plan.exact) {
        out[4 * s] = 1;
        out[4 * s + 1] = out[4 * s + 2] = 0;
    }
}

Model* Model::union_model(std::vector<Idx> const& closing) {
    if (union_model_ != nullptr && union_key_ == closing && union_version_ == state_version_) return union_model_.get();
    union_model_.reset();
    std::unique_ptr<Model> copy = clone();
    std::vector<BranchUpdate> lines;
    std::vector<TransformerUpdate> trafos;
    for (Idx const bi : closing) {
        copy->outage_base_state_.push_back({bi, branch_st_[bi].from_status, branch_st_[bi].to_status});
        if (bi < n_line()) {
            lines.push_back({line_in_[bi].id, 1, 1});
        } else {
            trafos.push_back({trafo_in_[bi - off_trafo()].id, 1, 1, kNaIntS});
        }
    }
    UpdateData close{};
    close.n_scenarios = 1;
    if (!lines.empty()) close.line = {static_cast<int64_t>(lines.size()), nullptr, lines.data()};
    if (!trafos.empty()) close.transformer = {static_cast<int64_t>(trafos.size()), nullptr, trafos.data()};
    copy->update_permanent(close);
    union_model_ = std::move(copy);
    union_key_ = closing;
    union_version_ = state_version_;
    return union_model_.get();
}

template <int B> bool Model::plan_outage_batch(UpdateData const& u, OutagePlan& plan) const {
    // load / generator updates may ride along (contingency x load profile): the device pipeline applies them as in any load batch
    if (u.shunt.data != nullptr || u.source.data != nullptr || u.voltage_regulator.data != nullptr || u.asym_line.data != nullptr ||
        u.generic_branch.data != nullptr || u.link.data != nullptr || u.three_winding_transformer.data != nullptr) {
        return false;
    }
    if (topo_.math.size() != 1 || u.n_scenarios <= 0) return false;
    if (n_t3w() != 0) return false; // the bridge analysis below walks two-way branches only
    MathTopology const& m = topo_.math[0];
    if (std::all_of(m.load_gen_type.begin(), m.load_gen_type.end(), [](int8_t t) { return t == 1; })) return false; // linear method
    constexpr size_t bb2 = static_cast<size_t>(B) * B * 2;
    Idx const n = u.n_scenarios;
    BridgeInfo const info = bridge_analysis();
    std::vector<char> const& bridge = info.bridge;
    std::unordered_map<Idx, int32_t> mask_of_branch; // bridge -> its mask in plan.dead
    plan.dead_off.assign(n, -1);
    plan.dead.clear();
    plan.exact.clear();
    auto lookup = [](ID id, Idx pos, Idx n_in_scenario, Idx n_comp, std::unordered_map<ID, Idx> const& map) -> Idx {
        if (id == kNaID) {
            if (n_in_scenario != n_comp) throw InvalidArgument("update without ids must cover every element of the component");
            return pos;
        }
        auto it = map.find(id);
        if (it == map.end()) throw InvalidArgument("The id cannot be found: " + std::to_string(id) + "\n");
        return it->second;
    };
    struct Change {
        Idx branch;
        bool from, to;
        IntS tap; // transformers: the scenario's tap position
    };
    auto base_tap = [this](Idx bi) -> IntS { return bi >= off_trafo() ? trafo_st_[bi - off_trafo()].tap_pos : IntS{0}; };
    // PGMB_OUTAGE_SLOTS=1: only one switched branch per scenario on the shared pattern (comparison)
    int max_slots = kMaxOutageSlots;
    if (char const* env = std::getenv("PGMB_OUTAGE_SLOTS")) max_slots = std::clamp(std::atoi(env), 1, kMaxOutageSlots);
    // pass 1: what every scenario switches
    std::vector<std::vector<Change>> all(n);
    std::vector<char> exact_flag(n, 0);
    for (Idx s = 0; s != n; ++s) {
        std::vector<Change>& changes = all[s];
        for (BranchSwitch const& b : outage_base_state_) changes.push_back({b.branch, b.from, b.to, base_tap(b.branch)}); // unless the scenario says otherwise
        auto note = [&](Idx bi, IntS from, IntS to, IntS tap = kNaIntS) {
            auto it = std::find_if(changes.begin(), changes.end(), [bi](Change const& c) { return c.branch == bi; });
            if (it == changes.end()) {
                changes.push_back({bi, branch_st_[bi].from_status, branch_st_[bi].to_status, base_tap(bi)});
                it = changes.end() - 1;
            }
            if (from != kNaIntS) it->from = from != 0;
            if (to != kNaIntS) it->to = to != 0;
            if (tap != kNaIntS) it->tap = tap_limit(trafo_c_[bi - off_trafo()], tap); // a tap position is one more set of branch parameters
        };
        try {
            {
                auto [b, e] = scenario_span<BranchUpdate>(u.line, s);
                for (auto p = b; p != e; ++p) note(lookup(p->id, p - b, e - b, n_line(), line_idx_), p->from_status, p->to_status);
            }
            {
                auto [b, e] = scenario_span<TransformerUpdate>(u.transformer, s);
                for (auto p = b; p != e; ++p) {
                    Idx const i = lookup(p->id, p - b, e - b, n_trafo(), trafo_idx_);
                    note(off_trafo() + i, p->from_status, p->to_status, p->tap_pos);
                }
            }
        } catch (InvalidArgument const&) {
            // an update that cannot be applied fails this scenario alone (job_dispatch.hpp:162-206): the per-scenario route
            // records its message and the other scenarios are still calculated
            exact_flag[s] = 1;
            changes.clear();
            continue;
        }
        std::erase_if(changes, [&](Change const& c) {
            return c.from == branch_st_[c.branch].from_status && c.to == branch_st_[c.branch].to_status && c.tap == base_tap(c.branch);
        });
        if (static_cast<int>(changes.size()) > max_slots) exact_flag[s] = 1;
        // only branches that are closed, in the math model and between two different buses in the base state keep the pattern
        for (Change const& c : changes) {
            Coupling const cp = topo_.branch[c.branch];
            bool const base_closed = branch_st_[c.branch].from_status && branch_st_[c.branch].to_status && cp.group == 0 &&
                                     m.branch_bus_idx[2 * cp.pos] >= 0 && m.branch_bus_idx[2 * cp.pos + 1] >= 0 &&
                                     m.branch_bus_idx[2 * cp.pos] != m.branch_bus_idx[2 * cp.pos + 1];
            if (!base_closed) exact_flag[s] = 1;
        }
    }
    // pass 2: which buses lose their supply.  One switched branch: the bridge analysis knows; several: a search from a supplied
    // node over the remaining closed branches.  Exactly one supplied part must remain (several = several math models: exact route)
    Idx n_supplied_base = 0, n_source_base = 0, start_node = -1;
    for (Idx i = 0; i != static_cast<Idx>(node_.size()); ++i) {
        if (topo_.node[i].group != 0) continue;
        ++n_supplied_base;
        n_source_base += info.self_source[i];
        if (start_node < 0 && info.self_source[i] != 0) start_node = i;
    }
    auto new_mask = [&]() -> uint8_t* {
        plan.dead.resize(plan.dead.size() + m.n_bus, 0);
        return &plan.dead[plan.dead.size() - m.n_bus];
    };
    {
        // distinct branch sets, searched by the host threads; verdict -2: not on the shared pattern, -1: nothing goes dark,
        // >= 0: mask index (the buses the search did not reach)
        std::map<std::vector<Idx>, Idx> set_index;
        std::vector<std::vector<Idx> const*> sets;
        std::vector<Idx> set_of_scenario(n, -1);
        for (Idx s = 0; s != n; ++s) {
            if (exact_flag[s] != 0 || all[s].size() < 2) continue;
            std::vector<Idx> key; // the branches that stop connecting their ends (a tap change alone keeps the connection)
            for (Change const& c : all[s])
                if (!(c.from && c.to)) key.push_back(c.branch);
            if (key.empty()) continue;
            std::sort(key.begin(), key.end());
            auto const [it, fresh] = set_index.emplace(std::move(key), static_cast<Idx>(sets.size()));
            if (fresh) sets.push_back(&it->first);
            set_of_scenario[s] = it->second;
        }
        Idx const n_set = static_cast<Idx>(sets.size());
        std::vector<int32_t> verdict(n_set, -2);
        std::vector<std::vector<Idx>> unreached(n_set); // bus positions
        auto search = [&](Idx first, Idx step) {
            std::vector<Idx> stamp(node_.size(), -1), stack;
            for (Idx k = first; k < n_set; k += step) {
                std::vector<Idx> const& key = *sets[k];
                stack.assign(1, start_node);
                stamp[start_node] = k;
                Idx reached = 0, sources = 0;
                while (!stack.empty()) {
                    Idx const v = stack.back();
                    stack.pop_back();
                    ++reached;
                    sources += info.self_source[v];
                    for (Idx e = info.adj_ptr[v]; e != info.adj_ptr[v + 1]; ++e) {
                        Idx const w = info.adj_node[e];
                        if (stamp[w] == k || std::find(key.begin(), key.end(), info.adj_edge[e]) != key.end()) continue;
                        stamp[w] = k;
                        stack.push_back(w);
                    }
                }
                if (sources != n_source_base) continue; // a second supplied part: its own math model in the reference
                verdict[k] = -1;
                if (reached == n_supplied_base) continue;
                verdict[k] = 0;
                for (Idx i = 0; i != static_cast<Idx>(node_.size()); ++i)
                    if (topo_.node[i].group == 0 && stamp[i] != k) unreached[k].push_back(topo_.node[i].pos);
            }
        };
        if (start_node >= 0 && n_set != 0) {
            Idx n_thread = std::min<Idx>(std::max<Idx>(1, std::thread::hardware_concurrency()), 32);
            if (char const* env = std::getenv("PGMB_MAX_HOST_THREADS")) n_thread = std::max(1, std::atoi(env));
            n_thread = std::min<Idx>(n_thread, std::max<Idx>(1, n_set * static_cast<Idx>(node_.size()) / 200000));
            if (n_thread <= 1) {
                search(0, 1);
            } else {
                std::vector<std::thread> pool;
                for (Idx t = 0; t != n_thread; ++t) pool.emplace_back(search, t, n_thread);
                for (std::thread& t : pool) t.join();
            }
        }
        for (Idx k = 0; k != n_set; ++k) {
            if (verdict[k] != 0) continue;
            verdict[k] = static_cast<int32_t>(plan.dead.size() / m.n_bus);
            uint8_t* mask = new_mask();
            for (Idx const pos : unreached[k]) mask[pos] = 1;
        }
        for (Idx s = 0; s != n; ++s) {
            if (set_of_scenario[s] < 0) continue;
            int32_t const v = verdict[set_of_scenario[s]];
            if (v == -2) {
                exact_flag[s] = 1;
            } else {
                plan.dead_off[s] = v;
            }
        }
    }
    for (Idx s = 0; s != n; ++s) {
        std::vector<Change> const& changes = all[s];
        if (exact_flag[s] != 0 || changes.size() != 1) continue;
        Change const& c = changes[0];
        if ((c.from && c.to) || !bridge[c.branch]) continue;
        // a bridge cuts a subtree of the DFS off: fine when exactly one side keeps a source (the other side goes dark)
        Idx const v = info.child[c.branch];
        Idx const inside = info.n_source[v], outside = info.n_source[info.root[v]] - inside;
        if ((inside == 0) == (outside == 0)) {
            exact_flag[s] = 1; // two supplied islands (two math models), or an island that was dark already
            continue;
        }
        auto it = mask_of_branch.find(c.branch);
        if (it == mask_of_branch.end()) {
            int32_t const k = static_cast<int32_t>(plan.dead.size() / m.n_bus);
            uint8_t* mask = new_mask();
            Idx const r = info.root[v];
            auto mark = [&](Idx t0, Idx t1) {
                for (Idx t = t0; t != t1; ++t) {
                    Coupling const nc = topo_.node[info.order[t]];
                    if (nc.group == 0) mask[nc.pos] = 1;
                }
            };
            if (inside == 0) {
                mark(info.disc[v], info.disc[v] + info.size[v]);
            } else {
                mark(info.disc[r], info.disc[v]);
                mark(info.disc[v] + info.size[v], info.disc[r] + info.size[r]);
            }
            it = mask_of_branch.emplace(c.branch, k).first;
        }
        plan.dead_off[s] = it->second;
    }
    // pass 3: slots per scenario = the most branches any shared-pattern scenario switches; parameters of the switched branches
    size_t K = 1;
    for (Idx s = 0; s != n; ++s)
        if (exact_flag[s] == 0) K = std::max(K, all[s].size());
    plan.n_slot = static_cast<int>(K);
    plan.math_branch.assign(n * K, -1);
    plan.bparam.assign(n * K * 4 * bb2, 0.0);
    plan.comp.assign(n * K, -1);
    plan.energized.assign(n * K, 0);
    for (Idx s = 0; s != n; ++s) {
        if (exact_flag[s] != 0) {
            plan.dead_off[s] = -1;
            plan.exact.push_back(s);
            continue;
        }
        uint8_t const* const mask = plan.dead_off[s] >= 0 ? &plan.dead[static_cast<size_t>(plan.dead_off[s]) * m.n_bus] : nullptr;
        for (size_t j = 0; j != all[s].size(); ++j) {
            Change const& c = all[s][j];
            Coupling const cp = topo_.branch[c.branch];
            size_t const slot = s * K + j;
            plan.math_branch[slot] = cp.pos;
            plan.comp[slot] = static_cast<int32_t>(c.branch);
            plan.energized[slot] = (c.from || c.to) ? 1 : 0;
            double* const bp = &plan.bparam[slot * 4 * bb2];
            BranchState const st{c.from, c.to};
            if (c.branch < n_line()) {
                line_param<B>(line_c_[c.branch], st, bp);
            } else {
                Idx const i = c.branch - off_trafo();
                transformer_param<B>(trafo_c_[i], st, c.tap, bp);
            }
            if (mask != nullptr) { // still connected to a supplied bus?  otherwise the branch itself goes dark
                bool const from_live = c.from && mask[m.branch_bus_idx[2 * cp.pos]] == 0;
                bool const to_live = c.to && mask[m.branch_bus_idx[2 * cp.pos + 1]] == 0;
                if (!from_live && !to_live) {
                    std::fill_n(bp, 4 * bb2, 0.0);
                    plan.energized[slot] = 0;
                }
            }
        }
    }
    return true;
}

template <int B>
int64_t Model::calculate_impl(ModelOptions const& opt, UpdateData const* update, OutputData const& out, int32_t* n_iter,
                              int32_t* status) {
    auto const t_all = Clock::now();
    for (double& t : timing) t = 0.0;
    batch_message.clear();
    device_ = opt.device;
    int64_t failed = 0;
    auto t0 = Clock::now();
    bool const has_reg = !reg_in_.empty();
    if (out.transformer_tap_regulator != nullptr && !tap_reg_in_.empty()) {
        // TransformerTapRegulator::get_null_output for every scenario; the optimizer route overwrites what it regulates
        Idx const n_scn = update == nullptr ? 1 : update->n_scenarios, n_reg = static_cast<Idx>(tap_reg_in_.size());
        auto* dst = static_cast<TransformerTapRegulatorOutput*>(out.transformer_tap_regulator);
        for (Idx s = 0; s != n_scn; ++s)
            for (Idx i = 0; i != n_reg; ++i) {
                std::memset(&dst[s * n_reg + i], 0, sizeof(*dst));
                dst[s * n_reg + i].id = tap_reg_in_[i].id;
                dst[s * n_reg + i].tap_pos = kNaIntS;
            }
    }
    if (update == nullptr && opt.tap_strategy != 0) {
        timing[0] += ms_since(t0);
        failed = run_tap_optimizer<B>(opt, out, 0, n_iter, status);
    } else if (update == nullptr) {
        check_regulators<B>(opt);
        prepare_engines<B>();
        std::vector<std::vector<double>> sinj(topo_.math.size()), uref(topo_.math.size());
        RegulatorInput reg;
        gather_pf_input<B>(sinj, uref, &reg);
        timing[0] += ms_since(t0);
        failed = run_block<B>(opt, 1, sinj, uref, out, 0, n_iter, status, &reg);
    } else {
        Idx const n = update->n_scenarios;
        // does any scenario touch something other than loads / generators / source references?
        // (regulator updates change the parameters every scenario of one engine call shares: scenario by scenario as well)
        bool const structural = update->line.data != nullptr || update->transformer.data != nullptr ||
                                update->asym_line.data != nullptr || update->generic_branch.data != nullptr ||
                                update->link.data != nullptr || update->three_winding_transformer.data != nullptr ||
                                update->shunt.data != nullptr || (has_reg && update->voltage_regulator.data != nullptr) ||
                                opt.tap_strategy != 0; // the tap search of a scenario is its own sequence of power flows
        bool source_param_change = false;
        if (update->source.data != nullptr) {
            for (Idx s = 0; s != n && !source_param_change; ++s) {
                auto [b, e] = scenario_span<SourceUpdate>(update->source, s);
                for (auto p = b; p != e; ++p)
                    if (p->status != kNaIntS || !std::isnan(p->sk) || !std::isnan(p->rx_ratio) || !std::isnan(p->z01_ratio))
                        source_param_change = true;
            }
        }
        bool device_done = false;
        bool scenario_errors = false; // some scenario of a load / source-reference batch cannot be applied
        // branch-switching batches (N-1): scenarios that keep the grid connected run together on the base pattern; `todo` is
        // what remains for the exact per-scenario route below
        std::vector<Idx> todo;
        bool todo_is_subset = false;
        std::vector<int32_t> status_local;
        // a device pass solved the batch; `exact` are the scenarios the scenario-by-scenario route below recomputes
        auto adopt_device_result = [&](int64_t r, std::vector<Idx> exact) {
            failed = r;
            for (Idx const s : exact) // placeholders of the scenarios that are recomputed below
                if (status[s] != 0) --failed;
            todo = std::move(exact);
            todo_is_subset = true;
            if (!todo.empty()) { // drop the placeholder messages
                std::string kept;
                size_t pos = 0;
                while (pos < batch_message.size()) {
                    size_t const end = batch_message.find('\n', pos);
                    std::string const line = batch_message.substr(pos, end == std::string::npos ? std::string::npos : end - pos + 1);
                    bool drop = false;
                    for (Idx const s : todo)
                        if (line.rfind("Error in batch #" + std::to_string(s) + ":", 0) == 0) drop = true;
                    if (!drop) kept += line;
                    if (end == std::string::npos) break;
                    pos = end + 1;
                }
                batch_message = std::move(kept);
            }
        };
        if (structural && !source_param_change && !has_reg && opt.tap_strategy == 0 && (opt.method == 1 || opt.method == -128) && n > 0 &&
            (update->line.data != nullptr || update->transformer.data != nullptr) && std::getenv("PGMB_N1_EXACT") == nullptr) {
            // scenarios that close branches run on the union grid's pattern (model.hpp: union_model)
            Model* const host = outage_host(*update);
            host->template prepare_engines<B>();
            OutagePlan plan;
            bool const planned = host->template plan_outage_batch<B>(*update, plan);
            if (std::getenv("PGMB_DEBUG_N1") != nullptr) {
                std::fprintf(stderr, "[pgmb n-1] planned=%d scenarios=%lld exact=%zu plan_ms=%.2f\n", planned ? 1 : 0,
                             static_cast<long long>(n), plan.exact.size(), ms_since(t0));
            }
            if (planned && 2 * plan.exact.size() <= static_cast<size_t>(n) && host->device_path_eligible(*update)) {
                if (status == nullptr) {
                    status_local.assign(n, 0);
                    status = status_local.data();
                }
                UpdateData none{}; // the branch updates are in the plan; what remains are the load / generator updates
                none.n_scenarios = n;
                none.sym_gen = update->sym_gen;
                none.asym_gen = update->asym_gen;
                none.sym_load = update->sym_load;
                none.asym_load = update->asym_load;
                host->outage_plan_ = &plan;
                timing[0] += ms_since(t0);
                int64_t r = -1;
                try {
                    r = host->run_batch_device(opt, B, none, out, n_iter, status);
                } catch (...) {
                    host->outage_plan_ = nullptr;
                    throw;
                }
                host->outage_plan_ = nullptr;
                if (host != this) {
                    batch_message = std::move(host->batch_message);
                    host->batch_message.clear();
                    for (int k = 1; k != 8; ++k) timing[k] = host->timing[k];
                }
                if (r < 0) batch_message.clear(); // messages of parts that ran before the attempt was given up
                t0 = Clock::now();
                if (r >= 0) adopt_device_result(r, std::move(plan.exact));
            }
        }
        // automatic tap changer on a load-profile batch with one regulated transformer: the scenarios search in lockstep, one
        // batched power flow per step (model_tap.cpp); PGMB_TAP_EXACT=1: every scenario searches on its own (comparison)
        if (opt.tap_strategy != 0 && n > 0 && std::getenv("PGMB_TAP_EXACT") == nullptr) {
            if (status == nullptr) {
                status_local.assign(n, 0);
                status = status_local.data();
            }
            std::vector<Idx> exact;
            timing[0] += ms_since(t0);
            int64_t const r = run_tap_lockstep<B>(opt, *update, out, n_iter, status, exact);
            t0 = Clock::now();
            if (r >= 0) {
                adopt_device_result(r, std::move(exact));
            } else {
                batch_message.clear();
            }
        }
        if (!structural && !source_param_change) prepare_engines<B>(); // the eligibility test looks at the math topology
        if (!structural && !source_param_change && has_reg) check_regulators<B>(opt);
        if (!structural && !source_param_change && device_path_eligible(*update)) {
            timing[0] += ms_since(t0);
            int64_t const r = run_batch_device(opt, B, *update, out, n_iter, status);
            if (r >= 0) {
                failed = r;
                device_done = true;
            } else {
                batch_message.clear(); // messages of parts that ran before the attempt was given up
            }
            t0 = Clock::now();
        }
        if (device_done || (todo_is_subset && todo.empty())) {
            // results written by the device path
        } else if (!structural && !source_param_change) {
            // fast path: one engine call for the whole batch.  A scenario whose update cannot be applied (unknown id, ...) fails
            // alone in the reference (job_dispatch.hpp:162-206): such a batch takes the scenario-by-scenario route below, which
            // records the message of every failing scenario and calculates the others.
            prepare_engines<B>();
            std::vector<std::vector<double>> sinj(topo_.math.size()), uref(topo_.math.size());
            RegulatorInput reg;
            for (Idx s = 0; s != n && !scenario_errors; ++s) {
                Saved saved;
                try {
                    apply_scenario(*update, s, &saved);
                    gather_pf_input<B>(sinj, uref, &reg);
                } catch (CudaError const&) {
                    restore(saved);
                    throw;
                } catch (std::exception const&) {
                    scenario_errors = true;
                }
                restore(saved);
            }
            timing[0] += ms_since(t0);
            if (n != 0 && !scenario_errors) failed = run_block<B>(opt, n, sinj, uref, out, 0, n_iter, status, &reg);
        }
        if (device_done || (todo_is_subset && todo.empty())) {
        } else if (!structural && !source_param_change && !scenario_errors) {
        } else {
            // general path: scenario by scenario (topology / parameters may change), still on the GPU.  Like the reference's
            // job dispatch (job_dispatch.hpp:88-160) the scenarios are spread over host threads, thread t taking scenarios
            // t, t + n_threads, ...; every thread works on its own copy of the model, whose engines own their CUDA streams,
            // so the symbolic stage of one scenario overlaps the kernels of the others.
            if (!todo_is_subset) {
                todo.resize(n);
                for (Idx s = 0; s != n; ++s) todo[s] = s;
            }
            Idx const n_todo = static_cast<Idx>(todo.size());
            Idx n_threads = opt.threading > 0 ? opt.threading : static_cast<Idx>(std::thread::hardware_concurrency());
            Idx max_threads = 32;
            if (char const* env = std::getenv("PGMB_MAX_HOST_THREADS")) max_threads = std::max(1, std::atoi(env));
            n_threads = std::max<Idx>(1, std::min<Idx>({n_threads, n_todo, max_threads}));
            // several GPUs (opt.n_devices / PGMB_DEVICES): the host threads are dealt round robin over the devices
            Idx n_dev_general = opt.n_devices;
            if (n_dev_general <= 1) {
                char const* env = std::getenv("PGMB_DEVICES");
                n_dev_general = env != nullptr ? std::atoi(env) : 1;
            }
            {
                int available = 0;
                if (cudaGetDeviceCount(&available) != cudaSuccess) available = 0;
                n_dev_general = std::max<Idx>(1, std::min<Idx>(n_dev_general, available - opt.device));
            }
            std::vector<std::string> messages(n);
            std::vector<int64_t> failed_per_thread(n_threads, 0);
            std::vector<std::exception_ptr> fatal(n_threads);
            // Scenarios with the SAME structural update (byte-identical rows of everything but loads / generators) share one
            // topology, one set of symbolic structures and engines: such a group is applied once and its load updates run as one
            // batch through the device pipeline -- "grouped by symbolic pattern" for batches that revisit a few switching states.
            // PGMB_GROUP_SCENARIOS=0: every scenario on its own (comparison).
            std::vector<std::vector<Idx>> groups;
            UpdateData structural_part = *update, load_part{};
            structural_part.sym_gen = structural_part.asym_gen = structural_part.sym_load = structural_part.asym_load = ComponentBuffer{};
            load_part.n_scenarios = n;
            load_part.sym_gen = update->sym_gen;
            load_part.asym_gen = update->asym_gen;
            load_part.sym_load = update->sym_load;
            load_part.asym_load = update->asym_load;
            {
                bool dense_loads = true;
                for (ComponentBuffer const* b : {&update->sym_gen, &update->asym_gen, &update->sym_load, &update->asym_load})
                    if (b->data != nullptr && (b->indptr != nullptr || b->n < 0)) dense_loads = false;
                char const* env = std::getenv("PGMB_GROUP_SCENARIOS");
                if (dense_loads && opt.tap_strategy == 0 && n_todo >= 2 && !(env != nullptr && std::atoi(env) == 0)) {
                    std::pair<ComponentBuffer const*, size_t> const parts[] = {
                        {&update->line, sizeof(BranchUpdate)}, {&update->transformer, sizeof(TransformerUpdate)},
                        {&update->shunt, sizeof(ShuntUpdate)}, {&update->source, sizeof(SourceUpdate)},
                        {&update->voltage_regulator, sizeof(VoltageRegulatorUpdate)}, {&update->asym_line, sizeof(BranchUpdate)},
                        {&update->generic_branch, sizeof(BranchUpdate)}, {&update->link, sizeof(BranchUpdate)},
                        {&update->three_winding_transformer, sizeof(ThreeWindingTransformerUpdate)},
                        {&update->transformer_tap_regulator, sizeof(TransformerTapRegulatorUpdate)}};
                    std::unordered_map<std::string, size_t> group_of;
                    for (Idx const sc : todo) {
                        std::string key;
                        for (auto const& [b, row] : parts) {
                            if (b->data == nullptr) continue;
                            auto const* base = static_cast<char const*>(b->data);
                            Idx const r0 = b->indptr != nullptr ? b->indptr[sc] : sc * b->n, r1 = b->indptr != nullptr ? b->indptr[sc + 1] : (sc + 1) * b->n;
                            int64_t const count = r1 - r0;
                            key.append(reinterpret_cast<char const*>(&count), sizeof(count));
                            key.append(base + r0 * static_cast<Idx>(row), static_cast<size_t>(count) * row);
                        }
                        auto const [it, fresh] = group_of.emplace(std::move(key), groups.size());
                        if (fresh) groups.emplace_back();
                        groups[it->second].push_back(sc);
                    }
                    if (groups.size() == static_cast<size_t>(n_todo)) groups.clear(); // nothing shared
                }
            }
            if (!groups.empty()) n_threads = std::max<Idx>(1, std::min<Idx>(n_threads, static_cast<Idx>(groups.size())));
            if (std::getenv("PGMB_DEBUG_N1") != nullptr && !groups.empty()) {
                std::fprintf(stderr, "[pgmb groups] %zu scenarios in %zu groups\n", todo.size(), groups.size());
            }
            // one group: its structure once, its loads as one device batch; false = take the scenarios one by one
            auto run_group = [&](Model& model, std::vector<Idx> const& g, Idx t) -> bool {
                if (g.size() < 2) return false;
                Saved saved;
                bool handled = false;
                try {
                    model.apply_scenario(structural_part, g[0], &saved);
                    model.template check_regulators<B>(opt);
                    model.template prepare_engines<B>();
                    Idx const gn = static_cast<Idx>(g.size());
                    UpdateData compact{};
                    compact.n_scenarios = gn;
                    std::vector<unsigned char> rows[4];
                    ComponentBuffer const* src[4] = {&load_part.sym_gen, &load_part.asym_gen, &load_part.sym_load, &load_part.asym_load};
                    ComponentBuffer* dst[4] = {&compact.sym_gen, &compact.asym_gen, &compact.sym_load, &compact.asym_load};
                    size_t const urow[4] = {sizeof(SymLoadGenUpdate), sizeof(AsymLoadGenUpdate), sizeof(SymLoadGenUpdate), sizeof(AsymLoadGenUpdate)};
                    for (int b = 0; b != 4; ++b) {
                        if (src[b]->data == nullptr) continue;
                        size_t const bytes = static_cast<size_t>(src[b]->n) * urow[b];
                        rows[b].resize(bytes * g.size());
                        for (size_t k = 0; k != g.size(); ++k)
                            std::memcpy(rows[b].data() + k * bytes, static_cast<unsigned char const*>(src[b]->data) + static_cast<size_t>(g[k]) * bytes, bytes);
                        *dst[b] = ComponentBuffer{src[b]->n, nullptr, rows[b].data()};
                    }
                    if (model.device_path_eligible(compact)) {
                        std::vector<int32_t> it(g.size(), 0), st(g.size(), 0);
                        ModelOptions go = opt;
                        go.n_devices = 1;
                        go.flags = 0;
                        go.device = model.device_;
                        model.batch_message.clear();
                        // the pass writes every scenario's output rows straight to its place in the caller's batch
                        model.out_scatter_ = g.data();
                        int64_t r = -1;
                        try {
                            r = model.run_batch_device_one(go, B, compact, out, it.data(), st.data(), 0);
                        } catch (...) {
                            model.out_scatter_ = nullptr;
                            throw;
                        }
                        model.out_scatter_ = nullptr;
                        if (r >= 0) {
                            for (size_t i = 0; i != g.size(); ++i) {
                                if (n_iter != nullptr) n_iter[g[i]] = it[i];
                                if (status != nullptr) status[g[i]] = st[i];
                            }
                            // the pipeline numbers its failures inside the group: back to the caller's scenario numbers
                            std::string const& all = model.batch_message;
                            std::string const tag = "Error in batch #";
                            size_t pos = all.find(tag);
                            while (pos != std::string::npos) {
                                size_t const next = all.find(tag, pos + tag.size());
                                std::string const entry = all.substr(pos, next == std::string::npos ? std::string::npos : next - pos);
                                size_t const colon = entry.find(':');
                                Idx const local = std::strtoll(entry.c_str() + tag.size(), nullptr, 10);
                                if (colon != std::string::npos && local >= 0 && local < gn) messages[g[local]] = tag + std::to_string(g[local]) + entry.substr(colon);
                                pos = next;
                            }
                            failed_per_thread[t] += r;
                            handled = true;
                        }
                        model.batch_message.clear();
                    }
                } catch (CudaError const&) {
                    model.restore(saved);
                    throw;
                } catch (std::exception const&) {
                    handled = false; // the scenario-by-scenario route records each scenario's own message
                    model.batch_message.clear();
                }
                model.restore(saved);
                return handled;
            };
            auto worker = [&](Model& model, Idx t) {
                try {
                    std::vector<Idx> mine; // scenarios this thread takes one by one
                    if (groups.empty()) {
                        for (Idx k = t; k < n_todo; k += n_threads) mine.push_back(todo[k]);
                    } else {
                        for (size_t gi = static_cast<size_t>(t); gi < groups.size(); gi += static_cast<size_t>(n_threads)) {
                            if (!run_group(model, groups[gi], t)) mine.insert(mine.end(), groups[gi].begin(), groups[gi].end());
                        }
                    }
                    for (Idx const s : mine) {
                        Saved saved;
                        try {
                            model.apply_scenario(*update, s, &saved);
                            if (opt.tap_strategy != 0) {
                                failed_per_thread[t] += model.template run_tap_optimizer<B>(opt, out, s, n_iter, status);
                            } else {
                                model.template check_regulators<B>(opt);
                                model.template prepare_engines<B>();
                                std::vector<std::vector<double>> sinj(model.topo_.math.size()), uref(model.topo_.math.size());
                                RegulatorInput reg;
                                model.template gather_pf_input<B>(sinj, uref, &reg);
                                failed_per_thread[t] += model.template run_block<B>(opt, 1, sinj, uref, out, s, n_iter, status, &reg);
                            }
                            messages[s] = std::move(model.batch_message);
                            model.batch_message.clear();
                        } catch (CudaError const&) {
                            model.restore(saved);
                            throw;
                        } catch (std::exception const& ex) {
                            ++failed_per_thread[t];
                            if (status != nullptr) status[s] = 3;
                            messages[s] = "Error in batch #" + std::to_string(s) + ": " + ex.what() + "\n";
                            model.batch_message.clear();
                        }
                        model.restore(saved);
                    }
                } catch (...) {
                    fatal[t] = std::current_exception();
                }
            };
            if (n_threads == 1) {
                worker(*this, 0);
            } else {
                std::vector<std::unique_ptr<Model>> copies;
                for (Idx t = 0; t != n_threads; ++t) {
                    copies.push_back(std::make_unique<Model>(*this));
                    copies.back()->dev_.reset();
                    copies.back()->device_ = opt.device + static_cast<int>(t % n_dev_general);
                    copies.back()->batch_message.clear();
                }
                std::vector<std::thread> pool;
                for (Idx t = 0; t != n_threads; ++t) pool.emplace_back(worker, std::ref(*copies[t]), t);
                for (auto& th : pool) th.join();
            }
            for (auto const& ex : fatal)
                if (ex) std::rethrow_exception(ex);
            for (Idx t = 0; t != n_threads; ++t) failed += failed_per_thread[t];
            for (Idx const s : todo) batch_message += messages[s];
        }
    }
    timing[5] = ms_since(t_all);
    return failed;
}

// used by the tap optimizer (model_tap.cpp)
#define PGMB_INSTANTIATE(B)                                                                                                          \
    template void Model::check_regulators<B>(ModelOptions const&) const;                                                             \
    template void Model::prepare_engines<B>();                                                                                       \
    template void Model::gather_pf_input<B>(std::vector<std::vector<double>>&, std::vector<std::vector<double>>&, RegulatorInput*)   \
        const;                                                                                                                       \
    template void Model::solve_block<B>(ModelOptions const&, Idx, std::vector<std::vector<double>> const&,                           \
                                        std::vector<std::vector<double>> const&, RegulatorInput const*, BlockSolution&);             \
    template void Model::write_output<B>(Idx, Idx, OutputData const&, std::vector<std::vector<double>> const (&)[6],                 \
                                         std::vector<std::vector<int8_t>> const&, std::vector<std::vector<int8_t>> const*) const;
PGMB_INSTANTIATE(1)
PGMB_INSTANTIATE(3)
#undef PGMB_INSTANTIATE

int64_t Model::calculate(ModelOptions const& opt, UpdateData const* update, OutputData const& out, int32_t* n_iter,
                         int32_t* status) {
    ModelOptions o = opt;
    if (tap_reg_in_.empty()) o.tap_strategy = 0; // nothing to regulate: the optimizer is the plain power flow
    return o.symmetric ? calculate_impl<1>(o, update, out, n_iter, status) : calculate_impl<3>(o, update, out, n_iter, status);
}

// ---- copy / indexer (PGM_copy_model, PGM_get_indexer) -------------------------------------------------------------
std::unique_ptr<Model> Model::clone() const {
    auto copy = std::make_unique<Model>(*this);
    copy->dev_.reset();
    copy->outage_plan_ = nullptr;
    copy->union_model_.reset();
    copy->outage_base_state_.clear();
    copy->batch_message.clear();
    return copy;
}

Idx Model::component_count(std::string const& c) const {
    if (c == "node") return static_cast<Idx>(node_.size());
    if (c == "line") return n_line();
    if (c == "asym_line") return n_aline();
    if (c == "generic_branch") return n_gb();
    if (c == "link") return n_link();
    if (c == "three_winding_transformer") return n_t3w();
    if (c == "transformer_tap_regulator") return static_cast<Idx>(tap_reg_in_.size());
    if (c == "transformer") return n_trafo();
    if (c == "shunt") return static_cast<Idx>(shunt_in_.size());
    if (c == "source") return static_cast<Idx>(source_in_.size());
    if (c == "sym_gen") return n_sym_gen_;
    if (c == "asym_gen") return n_asym_gen_;
    if (c == "sym_load") return n_sym_load_;
    if (c == "asym_load") return n_asym_load_;
    if (c == "voltage_regulator") return static_cast<Idx>(reg_in_.size());
    return -1;
}

void Model::get_indexer(std::string const& c, ID const* ids, Idx size, Idx* indexer) const {
    // load_gen components share one sequence (sym_gen, asym_gen, sym_load, asym_load): position = index - offset of the type
    std::unordered_map<ID, Idx> const* map = nullptr;
    Idx offset = 0, count = 0;
    if (c == "node") map = &node_idx_;
    else if (c == "line") map = &line_idx_;
    else if (c == "asym_line") map = &aline_idx_;
    else if (c == "generic_branch") map = &gb_idx_;
    else if (c == "link") map = &link_idx_;
    else if (c == "three_winding_transformer") map = &t3w_idx_;
    else if (c == "transformer_tap_regulator") map = &tap_reg_idx_;
    else if (c == "transformer") map = &trafo_idx_;
    else if (c == "shunt") map = &shunt_idx_;
    else if (c == "source") map = &source_idx_;
    else if (c == "voltage_regulator") map = &reg_idx_;
    else if (c == "sym_gen") { map = &lg_idx_; offset = 0; count = n_sym_gen_; }
    else if (c == "asym_gen") { map = &lg_idx_; offset = n_sym_gen_; count = n_asym_gen_; }
    else if (c == "sym_load") { map = &lg_idx_; offset = n_sym_gen_ + n_asym_gen_; count = n_sym_load_; }
    else if (c == "asym_load") { map = &lg_idx_; offset = n_sym_gen_ + n_asym_gen_ + n_sym_load_; count = n_asym_load_; }
    else return; // the reference matches the name against its component list and does nothing when none matches
    for (Idx i = 0; i != size; ++i) {
        auto const it = map->find(ids[i]);
        if (it == map->end()) {
            if (all_ids_.count(ids[i]) != 0) throw InvalidArgument("Wrong type for object with id " + std::to_string(ids[i]) + "\n");
            throw InvalidArgument("The id cannot be found: " + std::to_string(ids[i]) + "\n");
        }
        Idx const pos = it->second - offset;
        if (map == &lg_idx_ && (pos < 0 || pos >= count)) throw InvalidArgument("Wrong type for object with id " + std::to_string(ids[i]) + "\n");
        indexer[i] = pos;
    }
}

// ---- introspection -----------------------------------------------------------------------------------------------
Idx Model::n_math_groups() {
    prepare_topology();
    return static_cast<Idx>(topo_.math.size());
}

std::vector<int64_t> const& Model::get_index(Idx group, std::string const& name) {
    prepare_topology();
    std::string const key = std::to_string(group) + "." + name;
    auto it = index_cache_.find(key);
    if (it != index_cache_.end() && name != "tap_rank") return it->second;
    std::vector<int64_t> v;
    auto coupling = [](std::vector<Coupling> const& c) {
        std::vector<int64_t> o;
        for (auto const& x : c) {
            o.push_back(x.group);
            o.push_back(x.pos);
        }
        return o;
    };
    if (name == "coup.node") {
        v = coupling(topo_.node);
        v.resize(2 * node_.size()); // user nodes only (the internal nodes of three-way branches follow them)
    }
    else if (name == "coup.branch") v = coupling(topo_.branch);
    else if (name == "coup.shunt") v = coupling(topo_.shunt);
    else if (name == "coup.load_gen") v = coupling(topo_.load_gen);
    else if (name == "coup.source") v = coupling(topo_.source);
    else if (name == "coup.voltage_regulator") v = coupling(topo_.voltage_regulator);
    else if (name == "coup.branch3") {
        for (auto const& c : topo_.branch3) {
            v.push_back(c.group);
            v.insert(v.end(), c.pos.begin(), c.pos.end());
        }
    }
    else if (name == "tap_rank") {
        // ranking of the regulated transformers by the automatic tap changer (model_tap.cpp): kind (0 transformer, 1 three-winding
        // transformer), index within the kind, rank group -- in the order the optimizer visits them
        v = tap_rank_table();
        return index_cache_[key] = v; // depends on the regulators' state: recomputed on every call
    }
    else if (name == "branch_is_bridge" || name == "bridge_cut_size") {
        // host logic of the shared-pattern N-1 route (plan_outage_batch): per branch component (lines then transformers)
        // whether it is a bridge of the closed-branch graph, and how many nodes its DFS subtree holds
        BridgeInfo const info = bridge_analysis();
        for (size_t b = 0; b != info.bridge.size(); ++b) {
            v.push_back(name == "branch_is_bridge" ? info.bridge[b] : (info.bridge[b] ? info.size[info.child[b]] : 0));
        }
    }
    else {
        if (group < 0 || group >= static_cast<Idx>(topo_.math.size())) throw InvalidArgument("math group out of range");
        auto const& m = topo_.math[group];
        if (name == "slack_bus") v = {m.slack_bus};
        else if (name == "is_radial") v = {m.is_radial ? 1 : 0};
        else if (name == "branch_bus_idx") v = m.branch_bus_idx;
        else if (name == "fill_in") v = m.fill_in;
        else if (name == "sources_per_bus") v = m.sources_per_bus;
        else if (name == "shunts_per_bus") v = m.shunts_per_bus;
        else if (name == "load_gens_per_bus") v = m.load_gens_per_bus;
        else if (name == "load_gen_type") v.assign(m.load_gen_type.begin(), m.load_gen_type.end());
        else if (name == "voltage_regulators_per_load_gen") { // indptr form of the reference
            v.assign(m.n_load_gen() + 1, 0);
            for (size_t i = 0; i != m.load_gen_regulator.size(); ++i) v[i + 1] = v[i] + (m.load_gen_regulator[i] >= 0 ? 1 : 0);
        }
        else {
            LuPattern const p{m};
            if (name == "row_indptr") v = p.row_indptr;
            else if (name == "col_indices") v = p.col_indices;
            else if (name == "bus_entry") v = p.bus_entry;
            else if (name == "row_indptr_lu") v = p.row_indptr_lu;
            else if (name == "col_indices_lu") v = p.col_indices_lu;
            else if (name == "diag_lu") v = p.diag_lu;
            else if (name == "map_lu_y_bus") v = p.map_lu_y_bus;
            else if (name == "lu_transpose_entry") v = p.lu_transpose_entry;
            else if (name == "path_program") { // symbolic.hpp PathProgram words (empty: not a radial grid)
                EliminationSchedule const sch{p};
                RowProgram const rows{p, sch, m};
                PathProgram const pp{p, sch, m, rows};
                if (pp.valid) v.assign(pp.words.begin(), pp.words.end());
            }
            else throw InvalidArgument("unknown index array: " + name);
        }
    }
    return index_cache_.emplace(key, std::move(v)).first->second;
}

std::vector<double> const& Model::get_real(Idx group, bool symmetric, std::string const& name) {
    prepare_topology();
    if (group < 0 || group >= static_cast<Idx>(topo_.math.size())) throw InvalidArgument("math group out of range");
    std::string const key = std::to_string(group) + (symmetric ? ".s." : ".a.") + name;
    real_cache_.erase(key);
    std::vector<double> v;
    if (name == "phase_shift") {
        v = topo_.math[group].phase_shift;
    } else {
        std::vector<double> bp, sp, srcp;
        std::vector<std::vector<double>> sinj(topo_.math.size()), uref(topo_.math.size());
        if (symmetric) {
            param_arrays<1>(group, bp, sp, srcp);
            gather_pf_input<1>(sinj, uref);
        } else {
            param_arrays<3>(group, bp, sp, srcp);
            gather_pf_input<3>(sinj, uref);
        }
        if (name == "branch_param") v = bp;
        else if (name == "shunt_param") v = sp;
        else if (name == "source_param") v = srcp;
        else if (name == "s_injection") v = sinj[group];
        else if (name == "source_u_ref") v = uref[group];
        else throw InvalidArgument("unknown real array: " + name);
    }
    return real_cache_.emplace(key, std::move(v)).first->second;
}

} // namespace pgmb
