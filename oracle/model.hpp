// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the model-level glue around the solvers:
//   main_core/topology.hpp        construct_topology / construct_components_connections :231-255
//   main_core/y_bus.hpp           get_math_param :186-212
//   main_core/calculation_input_preparation.hpp  prepare_power_flow_input :163-188
//   main_core/output.hpp          output_result :60-187 ; main_core/topological_node_output.hpp :72-117
//   calculation_preparation.hpp   SolversCacheStatus / prepare_solvers :96-114, 240-279
//   main_model_impl.hpp           calculate_ :288-315, update_component :139-160, restore_components :254-261
//   job_dispatch.hpp              batch_calculation :37-68, single_thread_job :88-138, job_dispatch :142-172
//   job_adapter.hpp               setup_impl / winddown_impl :127-139
//   main_model_impl.hpp           calculate_with_optimizer :318-348 (automatic tap changer, tap_optimizer.hpp);
//   main_core/input.hpp           transformer tap regulators :168-214 ; main_core/output.hpp :381-396
// Component storage order (all_components.hpp:36-39): Node, Line, Transformer, Shunt, Source, SymGenerator,
// AsymGenerator, SymLoad, AsymLoad; Branch = Line..Transformer; GenericLoadGen = SymGen, AsymGen, SymLoad, AsymLoad.
#pragma once

#include "components.hpp"
#include "pf_solvers.hpp"
#include "tap_optimizer.hpp"
#include "topology.hpp"

#include <set>
#include <thread>
#include <unordered_map>

namespace pgm_oracle {

struct ModelInput {
    Idx n_node;
    NodeInput const* node;
    Idx n_line;
    LineInput const* line;
    Idx n_transformer;
    TransformerInput const* transformer;
    Idx n_shunt;
    ShuntInput const* shunt;
    Idx n_source;
    SourceInput const* source;
    Idx n_sym_gen;
    SymLoadGenInput const* sym_gen;
    Idx n_asym_gen;
    AsymLoadGenInput const* asym_gen;
    Idx n_sym_load;
    SymLoadGenInput const* sym_load;
    Idx n_asym_load;
    AsymLoadGenInput const* asym_load;
    Idx n_voltage_regulator;
    VoltageRegulatorInput const* voltage_regulator;
    Idx n_asym_line;
    AsymLineInput const* asym_line;
    Idx n_generic_branch;
    GenericBranchInput const* generic_branch;
    Idx n_link;
    LinkInput const* link;
    Idx n_three_winding_transformer;
    ThreeWindingTransformerInput const* three_winding_transformer;
    Idx n_transformer_tap_regulator;
    TransformerTapRegulatorInput const* transformer_tap_regulator;
};

// one buffer of a batch update dataset: uniform (indptr == nullptr, n_per_scenario elements each) or sparse
template <class T> struct UpdateBuffer {
    Idx n_per_scenario;
    Idx const* indptr;
    T const* data;
    std::pair<T const*, T const*> scenario(Idx s) const {
        if (data == nullptr) return {nullptr, nullptr};
        if (indptr != nullptr) return {data + indptr[s], data + indptr[s + 1]};
        return {data + s * n_per_scenario, data + (s + 1) * n_per_scenario};
    }
};
struct BatchUpdate {
    Idx n_scenarios;
    UpdateBuffer<BranchUpdate> line;
    UpdateBuffer<TransformerUpdate> transformer;
    UpdateBuffer<ShuntUpdate> shunt;
    UpdateBuffer<SourceUpdate> source;
    UpdateBuffer<SymLoadGenUpdate> sym_gen;
    UpdateBuffer<AsymLoadGenUpdate> asym_gen;
    UpdateBuffer<SymLoadGenUpdate> sym_load;
    UpdateBuffer<AsymLoadGenUpdate> asym_load;
    UpdateBuffer<VoltageRegulatorUpdate> voltage_regulator;
    UpdateBuffer<BranchUpdate> asym_line;
    UpdateBuffer<BranchUpdate> generic_branch;
    UpdateBuffer<BranchUpdate> link;
    UpdateBuffer<ThreeWindingTransformerUpdate> three_winding_transformer;
    UpdateBuffer<TransformerTapRegulatorUpdate> transformer_tap_regulator;
};
// output buffers, each [n_scenarios][n_component] or nullptr when the caller does not want that component
template <int B> struct BatchOutput {
    NodeOutput<B>* node;
    BranchOutput<B>* line;
    BranchOutput<B>* transformer;
    ApplianceOutput<B>* shunt;
    ApplianceOutput<B>* source;
    ApplianceOutput<B>* sym_gen;
    ApplianceOutput<B>* asym_gen;
    ApplianceOutput<B>* sym_load;
    ApplianceOutput<B>* asym_load;
    VoltageRegulatorOutput* voltage_regulator;
    BranchOutput<B>* asym_line;
    BranchOutput<B>* generic_branch;
    BranchOutput<B>* link;
    Branch3Output<B>* three_winding_transformer;
    TransformerTapRegulatorOutput* transformer_tap_regulator;
};

struct CalcOptions {
    CalculationMethod method{CalculationMethod::newton_raphson};
    double err_tol{1e-8};
    Idx max_iter{20};
    Idx threading{-1};
    bool reuse_ic_factorization{false};
    int tap_strategy{0}; // PGM_TapChangingStrategy: 0 disabled, 1 any_valid_tap, 2 min_voltage_tap, 3 max_voltage_tap, 4 fast_any_tap
};

class Model {
  public:
    Model(double system_frequency, ModelInput const& in) : system_frequency_{system_frequency} {
        for (Idx i = 0; i != in.n_node; ++i) {
            add_id(in.node[i].id);
            node_idx_[in.node[i].id] = i;
            nodes_.push_back(in.node[i]);
        }
        auto u_rated = [this](ID node_id) {
            auto it = node_idx_.find(node_id);
            if (it == node_idx_.end()) throw PgmError{"The id cannot be found: " + std::to_string(node_id) + "\n"};
            return nodes_[it->second].u_rated;
        };
        for (Idx i = 0; i != in.n_line; ++i) {
            add_id(in.line[i].id);
            lines_.emplace_back(in.line[i], system_frequency_, u_rated(in.line[i].from_node), u_rated(in.line[i].to_node));
        }
        for (Idx i = 0; i != in.n_asym_line; ++i) {
            add_id(in.asym_line[i].id);
            asym_lines_.emplace_back(in.asym_line[i], system_frequency_, u_rated(in.asym_line[i].from_node),
                                     u_rated(in.asym_line[i].to_node));
        }
        for (Idx i = 0; i != in.n_link; ++i) {
            add_id(in.link[i].id);
            links_.emplace_back(in.link[i], u_rated(in.link[i].from_node), u_rated(in.link[i].to_node));
        }
        for (Idx i = 0; i != in.n_generic_branch; ++i) {
            add_id(in.generic_branch[i].id);
            generic_branches_.emplace_back(in.generic_branch[i], u_rated(in.generic_branch[i].from_node),
                                           u_rated(in.generic_branch[i].to_node));
        }
        for (Idx i = 0; i != in.n_transformer; ++i) {
            add_id(in.transformer[i].id);
            transformers_.emplace_back(in.transformer[i], u_rated(in.transformer[i].from_node),
                                       u_rated(in.transformer[i].to_node));
        }
        for (Idx i = 0; i != in.n_three_winding_transformer; ++i) {
            auto const& t = in.three_winding_transformer[i];
            add_id(t.id);
            t3w_.emplace_back(t, u_rated(t.node_1), u_rated(t.node_2), u_rated(t.node_3));
        }
        for (Idx i = 0; i != in.n_shunt; ++i) {
            add_id(in.shunt[i].id);
            shunts_.emplace_back(in.shunt[i], u_rated(in.shunt[i].node));
        }
        for (Idx i = 0; i != in.n_source; ++i) {
            add_id(in.source[i].id);
            sources_.emplace_back(in.source[i], u_rated(in.source[i].node));
        }
        auto add_sym = [&](SymLoadGenInput const* p, Idx n, double dir) {
            for (Idx i = 0; i != n; ++i) {
                add_id(p[i].id);
                load_gens_.emplace_back(p[i].id, p[i].node, p[i].status, p[i].type, u_rated(p[i].node), 1, dir,
                                        &p[i].p_specified, &p[i].q_specified);
            }
        };
        auto add_asym = [&](AsymLoadGenInput const* p, Idx n, double dir) {
            for (Idx i = 0; i != n; ++i) {
                add_id(p[i].id);
                load_gens_.emplace_back(p[i].id, p[i].node, p[i].status, p[i].type, u_rated(p[i].node), 3, dir,
                                        p[i].p_specified, p[i].q_specified);
            }
        };
        n_sym_gen_ = in.n_sym_gen;
        n_asym_gen_ = in.n_asym_gen;
        n_sym_load_ = in.n_sym_load;
        n_asym_load_ = in.n_asym_load;
        add_sym(in.sym_gen, in.n_sym_gen, 1.0);
        add_asym(in.asym_gen, in.n_asym_gen, 1.0);
        add_sym(in.sym_load, in.n_sym_load, -1.0);
        add_asym(in.asym_load, in.n_asym_load, -1.0);
        // id -> sequence index within each updatable type
        for (size_t i = 0; i != lines_.size(); ++i) line_idx_[lines_[i].id] = static_cast<Idx>(i);
        for (size_t i = 0; i != transformers_.size(); ++i) transformer_idx_[transformers_[i].id] = static_cast<Idx>(i);
        for (size_t i = 0; i != asym_lines_.size(); ++i) asym_line_idx_[asym_lines_[i].id] = static_cast<Idx>(i);
        for (size_t i = 0; i != generic_branches_.size(); ++i) generic_branch_idx_[generic_branches_[i].id] = static_cast<Idx>(i);
        for (size_t i = 0; i != links_.size(); ++i) link_idx_[links_[i].id] = static_cast<Idx>(i);
        for (size_t i = 0; i != t3w_.size(); ++i) t3w_idx_[t3w_[i].in.id] = static_cast<Idx>(i);
        for (size_t i = 0; i != shunts_.size(); ++i) shunt_idx_[shunts_[i].id] = static_cast<Idx>(i);
        for (size_t i = 0; i != sources_.size(); ++i) source_idx_[sources_[i].id] = static_cast<Idx>(i);
        for (size_t i = 0; i != load_gens_.size(); ++i) load_gen_idx_[load_gens_[i].id] = static_cast<Idx>(i);
        // transformer tap regulators (main_core/input.hpp:168-214): the regulated object must be a transformer or a three-winding
        // transformer, the control side one of its terminals; at most one regulator per object
        {
            std::set<ID> tap_regulated;
            for (Idx i = 0; i != in.n_transformer_tap_regulator; ++i) {
                auto const& r = in.transformer_tap_regulator[i];
                add_id(r.id);
                if (all_ids_.find(r.regulated_object) == all_ids_.end()) {
                    throw PgmError{"The id cannot be found: " + std::to_string(r.regulated_object) + "\n"};
                }
                auto bad_side = [&] {
                    return PgmError{"transformer_tap_regulator item retrieval is not implemented for ControlSide #" +
                                    std::to_string(static_cast<int>(r.control_side)) + "!\n"};
                };
                if (auto it = transformer_idx_.find(r.regulated_object); it != transformer_idx_.end()) {
                    if (r.control_side != 0 && r.control_side != 1) throw bad_side();
                    auto const& t = transformers_[it->second];
                    tap_regulators_.emplace_back(r, false, u_rated(r.control_side == 0 ? t.from_node : t.to_node));
                } else if (auto it3 = t3w_idx_.find(r.regulated_object); it3 != t3w_idx_.end()) {
                    if (r.control_side < 0 || r.control_side > 2) throw bad_side();
                    tap_regulators_.emplace_back(r, true, t3w_[it3->second].u_rated[r.control_side]);
                } else {
                    throw PgmError{"transformer_tap_regulator regulator is not supported for object with ID " +
                                   std::to_string(r.regulated_object) + "\n"};
                }
                tap_regulated.insert(r.regulated_object);
                tap_regulator_idx_[r.id] = i;
            }
            if (static_cast<Idx>(tap_regulated.size()) != in.n_transformer_tap_regulator) {
                throw PgmError{"There are objects regulated by more than one regulator. Maximum one regulator is allowed."};
            }
        }
        // voltage regulators (main_core/input.hpp:216-241): the regulated object must be a load / generator, at most one
        // regulator per object
        std::set<ID> regulated;
        for (Idx i = 0; i != in.n_voltage_regulator; ++i) {
            add_id(in.voltage_regulator[i].id);
            auto it = load_gen_idx_.find(in.voltage_regulator[i].regulated_object);
            if (it == load_gen_idx_.end()) {
                throw PgmError{"voltage_regulator has invalid regulated object " + std::to_string(in.voltage_regulator[i].regulated_object)};
            }
            if (!regulated.insert(in.voltage_regulator[i].regulated_object).second) {
                throw PgmError{"There are objects regulated by more than one regulator. Maximum one regulator is allowed."};
            }
            regulators_.emplace_back(in.voltage_regulator[i]);
            regulator_idx_[in.voltage_regulator[i].id] = i;
        }
    }

    Idx n_node() const { return static_cast<Idx>(nodes_.size()); }
    Idx n_branch() const {
        return static_cast<Idx>(lines_.size() + asym_lines_.size() + links_.size() + generic_branches_.size() + transformers_.size());
    }

    // ---- single calculation ----
    // calculation_preparation.hpp:163-225 check_state_validity + main_model_impl.hpp:362-366, 400-420
    template <int B> void check_regulators(CalcOptions const& opt) const {
        if (!regulators_.empty() && !tap_regulators_.empty()) { // calculation_preparation.hpp:219-222
            throw PgmError{"The combination of voltage regulators and transformer tap regulators is not supported in the same model."};
        }
        if (regulators_.empty()) return;
        if (opt.method != CalculationMethod::newton_raphson && opt.method != CalculationMethod::default_method) {
            throw PgmError{"The calculation method is invalid for this calculation!"};
        }
        std::unordered_map<ID, std::pair<ID, double>> node_ref; // node -> (regulator id, u_ref)
        for (auto const& r : regulators_) {
            if (!r.status) continue;
            auto const& lg = load_gens_[load_gen_idx_.at(r.regulated_object)];
            auto it = node_ref.find(lg.node);
            if (it != node_ref.end()) {
                if (it->second.second != r.u_ref) {
                    throw PgmError{"Conflicting u_ref values detected for voltage regulators " + std::to_string(it->second.first) + ", " +
                                   std::to_string(r.id) + "."}; // exception.hpp:168-172
                }
            } else {
                node_ref[lg.node] = {r.id, r.u_ref};
            }
        }
        for (auto const& r : regulators_) {
            if (!r.status) continue;
            if (load_gens_[load_gen_idx_.at(r.regulated_object)].type != LoadGenType::const_pq) {
                throw PgmError{"Unsupported load_gen type for voltage regulators " + std::to_string(r.id) + "."}; // exception.hpp:174-178
            }
        }
        for (auto const& src : sources_) {
            if (src.status && node_ref.count(src.node) != 0) {
                throw PgmError{"Nodes with a source and a voltage regulated load/generator are not supported when both are enabled. Found at node with id " +
                               std::to_string(src.node)}; // exception.hpp:180-186
            }
        }
        if constexpr (B == 3) {
            for (auto const& r : regulators_)
                if (!is_nan(r.q_min) || !is_nan(r.q_max)) {
                    throw PgmError{"Voltage Regulator with Qmin/Qmax limits for asymmetric calculations is an experimental feature"};
                }
        }
    }

    template <int B> std::vector<SolverOutput<B>> solve(CalcOptions const& opt, bool cache_run = false) {
        check_regulators<B>(opt);
        prepare_solvers<B>();
        auto const pf_input = prepare_power_flow_input<B>();
        auto& ys = y_bus<B>();
        auto& solvers = math_solvers<B>();
        std::vector<SolverOutput<B>> so;
        for (size_t g = 0; g != ys.size(); ++g) {
            so.push_back(solvers[g].run_power_flow(pf_input[g], opt.err_tol, opt.max_iter, opt.method, ys[g],
                                                   opt.reuse_ic_factorization, cache_run));
            last_num_iter_ = std::max(last_num_iter_, so.back().num_iter);
        }
        return so;
    }
    template <int B> void calculate(CalcOptions const& opt, BatchOutput<B> const& out, Idx scenario, bool cache_run = false) {
        last_num_iter_ = 0;
        if (opt.tap_strategy != 0 && !cache_run) {
            calculate_with_optimizer<B>(opt, out, scenario);
            return;
        }
        auto const so = solve<B>(opt, cache_run);
        output_result<B>(so, out, scenario);
    }

    // ---- automatic tap changer (main_model_impl.hpp:318-348, tap_optimizer.hpp) ----
    // build_transformer_graph (tap_position_optimizer.hpp:143-305): Transformer, ThreeWindingTransformer, Line, Link edges
    tap::RankedGroups rank_transformers() const {
        std::vector<tap::GraphEdge> edges;
        auto node = [this](ID id) { return node_idx_.at(id); };
        auto both_ways = [&](Idx a, Idx b) {
            edges.push_back({a, b, 0, {-1, -1}, na_IntID});
            edges.push_back({b, a, 0, {-1, -1}, na_IntID});
        };
        auto regulator_of = [this](ID object) -> TransformerTapRegulator const* {
            for (auto const& r : tap_regulators_)
                if (r.status && r.regulated_object == object) return &r;
            return nullptr;
        };
        for (size_t i = 0; i != transformers_.size(); ++i) {
            auto const& t = transformers_[i];
            if (!t.from_status || !t.to_status) continue;
            if (auto const* r = regulator_of(t.id)) {
                Idx const control = node(r->control_side == 0 ? t.from_node : t.to_node);
                Idx const other = node(r->control_side == 0 ? t.to_node : t.from_node);
                edges.push_back({other, control, 1, {0, static_cast<Idx>(i)}, t.id});
            } else {
                both_ways(node(t.from_node), node(t.to_node));
            }
        }
        for (size_t i = 0; i != t3w_.size(); ++i) {
            auto const& t = t3w_[i];
            Idx const nodes[3] = {node(t.in.node_1), node(t.in.node_2), node(t.in.node_3)};
            auto const* r = regulator_of(t.in.id);
            constexpr int combinations[3][2] = {{0, 1}, {0, 2}, {1, 2}};
            for (auto const& c : combinations) {
                int const first = c[0], second = c[1];
                if (!t.status[first] || !t.status[second]) continue;
                bool const tap_at_first = t.in.tap_side == first;
                if (r != nullptr && (tap_at_first || t.in.tap_side == second)) {
                    bool const tap_at_control = r->control_side == t.in.tap_side;
                    Idx const tap_side_node = tap_at_first ? nodes[first] : nodes[second];
                    Idx const non_tap_side_node = tap_at_first ? nodes[second] : nodes[first];
                    edges.push_back({tap_at_control ? non_tap_side_node : tap_side_node, tap_at_control ? tap_side_node : non_tap_side_node,
                                     1, {1, static_cast<Idx>(i)}, t.in.id});
                } else {
                    both_ways(nodes[first], nodes[second]);
                }
            }
        }
        for (auto const& l : lines_)
            if (l.from_status && l.to_status) both_ways(node(l.from_node), node(l.to_node));
        for (auto const& l : links_)
            if (l.from_status && l.to_status) both_ways(node(l.from_node), node(l.to_node));
        std::vector<char> is_source(nodes_.size(), 0);
        for (auto const& src : sources_) is_source[node(src.node)] = src.status ? 1 : 0;
        return tap::rank_transformers(n_node(), std::move(edges), is_source);
    }

    template <int B> struct TapHost {
        Model& m;
        CalcOptions opt;
        std::vector<SolverOutput<B>> so;
        IntS tap_pos(tap::Regulated const& r) const {
            return r.index.group == 0 ? m.transformers_[r.index.pos].tap_pos : m.t3w_[r.index.pos].tap_pos;
        }
        void set_taps(std::vector<std::pair<tap::Regulated const*, IntS>> const& updates) {
            for (auto const& [r, pos] : updates) {
                bool const changed = r->index.group == 0 ? m.transformers_[r->index.pos].set_tap(pos) : m.t3w_[r->index.pos].set_tap(pos);
                if (changed) m.param_valid_[0] = m.param_valid_[1] = false;
            }
        }
        void calculate(CalculationMethod method) {
            CalcOptions o = opt;
            o.method = method;
            so = m.template solve<B>(o);
        }
        // u_pu / i_pu of the controlled node (:628-700), NodeState <=> TransformerTapRegulatorCalcParam (:702-731)
        tap::Comparison compare(tap::Regulated const& r) const {
            auto const& reg = m.tap_regulators_[r.regulator];
            Idx2D bus{-1, -1};
            CVec<B> i_pu{};
            if (r.index.group == 0) {
                auto const& t = m.transformers_[r.index.pos];
                bus = m.coup_.node[m.node_idx_.at(reg.control_side == 0 ? t.from_node : t.to_node)];
                Idx2D const br = m.coup_.branch[m.branch_seq_transformer(r.index.pos)];
                if (br.group != -1) i_pu = reg.control_side == 0 ? so[br.group].branch[br.pos].i_f : so[br.group].branch[br.pos].i_t;
            } else {
                auto const& t = m.t3w_[r.index.pos];
                ID const node_id = reg.control_side == 0 ? t.in.node_1 : reg.control_side == 1 ? t.in.node_2 : t.in.node_3;
                bus = m.coup_.node[m.node_idx_.at(node_id)];
                auto const& b3 = m.coup_.branch3[r.index.pos];
                if (b3.first != -1) i_pu = so[b3.first].branch[b3.second[reg.control_side]].i_f;
            }
            if (bus.group == -1) return {false, 0};
            auto const param = reg.template calc_param<B>();
            CVec<B> const& u = so[bus.group].u[bus.pos];
            double v = 0.0;
            for (int p = 0; p != B; ++p) v += cabs(u(p) + param.z_compensation * i_pu(p));
            v /= B; // mean_val(cabs(u_compensated))
            double const lower = param.u_set - 0.5 * param.u_band, upper = param.u_set + 0.5 * param.u_band;
            return {true, v < lower ? -1 : v > upper ? 1 : 0};
        }
    };

    template <int B> void calculate_with_optimizer(CalcOptions const& opt, BatchOutput<B> const& out, Idx scenario) {
        prepare_topology();
        tap::RankedGroups const ranked = rank_transformers();
        std::vector<std::vector<tap::Regulated>> order;
        for (auto const& group : ranked) {
            order.emplace_back();
            for (Idx2D const& idx : group) {
                ID const object = idx.group == 0 ? transformers_[idx.pos].id : t3w_[idx.pos].in.id;
                Idx regulator = -1;
                for (size_t k = 0; k != tap_regulators_.size() && regulator < 0; ++k)
                    if (tap_regulators_[k].regulated_object == object) regulator = static_cast<Idx>(k);
                IntS const tap_side = idx.group == 0 ? static_cast<IntS>(transformers_[idx.pos].tap_side) : t3w_[idx.pos].in.tap_side;
                IntS const tap_min = idx.group == 0 ? transformers_[idx.pos].tap_min : t3w_[idx.pos].in.tap_min;
                IntS const tap_max = idx.group == 0 ? transformers_[idx.pos].tap_max : t3w_[idx.pos].in.tap_max;
                order.back().push_back({idx, regulator, tap_min, tap_max, tap_regulators_[regulator].control_side == tap_side});
            }
        }
        TapHost<B> host{*this, opt, {}};
        // cache_states / update_state(cache) (:966-984, :1379-1395): the tap positions come back whatever happens
        std::vector<std::pair<tap::Regulated const*, IntS>> cache;
        tap::TapPositionOptimizer<TapHost<B>> optimizer{host, order, static_cast<tap::Strategy>(opt.tap_strategy)};
        for (auto const& group : optimizer.order())
            for (auto const& r : group) cache.emplace_back(&r, host.tap_pos(r));
        try {
            optimizer.optimize(opt.method);
        } catch (...) {
            host.set_taps(cache);
            throw;
        }
        tap_positions_found_.assign(tap_regulators_.size(), na_IntS);
        for (auto const& group : optimizer.order())
            for (auto const& r : group) tap_positions_found_[r.regulator] = host.tap_pos(r);
        host.set_taps(cache);
        output_result<B>(host.so, out, scenario);
        tap_positions_found_.clear();
    }
    Idx last_num_iter() const { return last_num_iter_; }

    // ---- batch calculation (job_dispatch.hpp:37-68) ----
    // returns per-scenario error messages (empty string = ok); n_iter[s] filled when not null
    template <int B>
    std::vector<std::string> batch_calculate(CalcOptions const& opt, BatchUpdate const& upd, BatchOutput<B> const& out,
                                             Idx* n_iter) {
        Idx const n = upd.n_scenarios;
        std::vector<std::string> messages(n);
        if (n == 0) return messages;
        // cache run: one calculation to warm topology/solver caches; errors ignored like the reference
        try {
            BatchOutput<B> none{};
            CalcOptions cache_opt = opt;
            cache_opt.err_tol = std::numeric_limits<double>::max();
            cache_opt.max_iter = 1;
            calculate<B>(cache_opt, none, 0, true);
        } catch (SparseMatrixError const&) {
        } catch (IterationDiverge const&) {
        }
        auto job = [&](Idx start, Idx stride) {
            Model local{*this}; // deep copy per thread (job_adapter.hpp:34-41)
            for (Idx s = start; s < n; s += stride) {
                Saved saved;
                try {
                    local.apply_update(upd, s, saved);
                    local.template calculate<B>(opt, out, s);
                    if (n_iter != nullptr) n_iter[s] = local.last_num_iter();
                } catch (std::exception const& e) {
                    messages[s] = e.what();
                }
                local.restore(saved);
            }
        };
        Idx const hw = static_cast<Idx>(std::thread::hardware_concurrency());
        Idx n_thread = opt.threading == 0 ? hw : opt.threading;
        if (opt.threading < 0 || n_thread < 2) {
            job(0, 1);
        } else {
            n_thread = std::min(n_thread, n);
            std::vector<std::thread> threads;
            for (Idx t = 0; t != n_thread; ++t) threads.emplace_back(job, t, n_thread);
            for (auto& t : threads) t.join();
        }
        return messages;
    }

    // permanent update with scenario 0 of the given dataset
    void update_permanent(BatchUpdate const& upd) {
        Saved saved;
        apply_update(upd, 0, saved);
    }

    // math-level accessors for tests
    void ensure_topology() { prepare_topology(); }
    std::vector<std::shared_ptr<MathTopology const>> const& math_topology() const { return math_topo_; }
    ComponentToMathCoupling const& coupling() const { return coup_; }
    template <int B> std::vector<YBus<B>>& prepared_y_bus() {
        prepare_solvers<B>();
        return y_bus<B>();
    }
    template <int B> std::vector<PowerFlowInput<B>> power_flow_input() {
        prepare_solvers<B>();
        return prepare_power_flow_input<B>();
    }

  private:
    double system_frequency_;
    std::vector<NodeInput> nodes_;
    std::vector<Line> lines_;
    std::vector<AsymLine> asym_lines_;
    std::vector<GenericBranch> generic_branches_;
    std::vector<Link> links_;
    std::vector<Transformer> transformers_;
    std::vector<ThreeWindingTransformer> t3w_;
    std::unordered_map<ID, Idx> asym_line_idx_, generic_branch_idx_, link_idx_, t3w_idx_;
    std::vector<Shunt> shunts_;
    std::vector<Source> sources_;
    std::vector<LoadGen> load_gens_; // sym_gen, asym_gen, sym_load, asym_load
    std::vector<VoltageRegulator> regulators_;
    std::unordered_map<ID, Idx> regulator_idx_;
    std::vector<TransformerTapRegulator> tap_regulators_;
    std::unordered_map<ID, Idx> tap_regulator_idx_;
    std::vector<IntS> tap_positions_found_; // per tap regulator, filled by the optimizer for output_result (na: not regulated)
    Idx n_sym_gen_{}, n_asym_gen_{}, n_sym_load_{}, n_asym_load_{};
    std::unordered_map<ID, Idx> all_ids_, node_idx_, line_idx_, transformer_idx_, shunt_idx_, source_idx_, load_gen_idx_;

    // caches
    bool topo_valid_{false};
    bool param_valid_[2]{false, false}; // [sym, asym]
    std::vector<std::shared_ptr<MathTopology const>> math_topo_;
    ComponentToMathCoupling coup_;
    ComponentTopology comp_topo_;
    std::vector<YBus<1>> y_bus_sym_;
    std::vector<YBus<3>> y_bus_asym_;
    std::vector<MathSolver<1>> solver_sym_;
    std::vector<MathSolver<3>> solver_asym_;
    Idx last_num_iter_{};

    void add_id(ID id) {
        if (!all_ids_.emplace(id, 0).second) throw PgmError{"Conflicting id detected: " + std::to_string(id) + "\n"};
    }
    template <int B> std::vector<YBus<B>>& y_bus() {
        if constexpr (B == 1) {
            return y_bus_sym_;
        } else {
            return y_bus_asym_;
        }
    }
    template <int B> std::vector<MathSolver<B>>& math_solvers() {
        if constexpr (B == 1) {
            return solver_sym_;
        } else {
            return solver_asym_;
        }
    }
    Idx branch_seq_line(Idx i) const { return i; }
    // branch sequence = component order of the reference (all_components.hpp:36-39): line, asym_line, (link), generic_branch,
    // transformer
    Idx branch_seq_asym_line(Idx i) const { return static_cast<Idx>(lines_.size()) + i; }
    Idx branch_seq_link(Idx i) const { return static_cast<Idx>(lines_.size() + asym_lines_.size()) + i; }
    Idx branch_seq_generic_branch(Idx i) const { return static_cast<Idx>(lines_.size() + asym_lines_.size() + links_.size()) + i; }
    Idx branch_seq_transformer(Idx i) const {
        return static_cast<Idx>(lines_.size() + asym_lines_.size() + links_.size() + generic_branches_.size()) + i;
    }

    void prepare_topology() {
        if (topo_valid_) return;
        comp_topo_ = ComponentTopology{};
        ComponentConnections conn;
        comp_topo_.n_node = n_node();
        auto add_branch = [&](BranchBase const& b, double shift) {
            comp_topo_.branch_node_idx.push_back({node_idx_.at(b.from_node), node_idx_.at(b.to_node)});
            conn.branch_connected.push_back({static_cast<IntS>(b.from_status), static_cast<IntS>(b.to_status)});
            conn.branch_phase_shift.push_back(shift);
        };
        for (auto const& l : lines_) add_branch(l, l.phase_shift());
        for (auto const& l : asym_lines_) add_branch(l, l.phase_shift());
        for (auto const& l : links_) add_branch(l, l.phase_shift());
        for (auto const& g : generic_branches_) add_branch(g, g.phase_shift());
        for (auto const& t : transformers_) add_branch(t, t.phase_shift());
        for (auto const& t : t3w_) { // main_core/topology.hpp: branch3_node_idx / branch3_connected / branch3_phase_shift
            comp_topo_.branch3_node_idx.push_back({node_idx_.at(t.in.node_1), node_idx_.at(t.in.node_2), node_idx_.at(t.in.node_3)});
            conn.branch3_connected.push_back({static_cast<IntS>(t.status[0]), static_cast<IntS>(t.status[1]), static_cast<IntS>(t.status[2])});
            conn.branch3_phase_shift.push_back(t.phase_shift());
        }
        for (auto const& s : shunts_) comp_topo_.shunt_node_idx.push_back(node_idx_.at(s.node));
        for (auto const& s : sources_) {
            comp_topo_.source_node_idx.push_back(node_idx_.at(s.node));
            conn.source_connected.push_back(static_cast<IntS>(s.status));
        }
        for (auto const& lg : load_gens_) {
            comp_topo_.load_gen_node_idx.push_back(node_idx_.at(lg.node));
            comp_topo_.load_gen_type.push_back(lg.type);
        }
        for (auto const& r : regulators_) comp_topo_.regulated_load_gen_idx.push_back(load_gen_idx_.at(r.regulated_object));
        Topology topology{comp_topo_, conn};
        auto [math, coup] = topology.build_topology();
        math_topo_.clear();
        for (auto& m : math) math_topo_.push_back(std::make_shared<MathTopology const>(std::move(m)));
        coup_ = std::move(coup);
        y_bus_sym_.clear();
        y_bus_asym_.clear();
        solver_sym_.clear();
        solver_asym_.clear();
        param_valid_[0] = param_valid_[1] = false;
        topo_valid_ = true;
    }

    template <int B> std::vector<MathParam<B>> get_math_param() const {
        std::vector<MathParam<B>> param(math_topo_.size());
        for (size_t g = 0; g != math_topo_.size(); ++g) {
            param[g].branch_param.resize(math_topo_[g]->n_branch());
            param[g].shunt_param.resize(math_topo_[g]->n_shunt());
            param[g].source_param.resize(math_topo_[g]->n_source());
        }
        for (size_t i = 0; i != lines_.size(); ++i) {
            Idx2D const m = coup_.branch[branch_seq_line(static_cast<Idx>(i))];
            if (m.group != -1) param[m.group].branch_param[m.pos] = lines_[i].calc_param<B>();
        }
        for (size_t i = 0; i != asym_lines_.size(); ++i) {
            Idx2D const m = coup_.branch[branch_seq_asym_line(static_cast<Idx>(i))];
            if (m.group != -1) param[m.group].branch_param[m.pos] = asym_lines_[i].calc_param<B>();
        }
        for (size_t i = 0; i != generic_branches_.size(); ++i) {
            Idx2D const m = coup_.branch[branch_seq_generic_branch(static_cast<Idx>(i))];
            if (m.group != -1) param[m.group].branch_param[m.pos] = generic_branches_[i].calc_param<B>();
        }
        for (size_t i = 0; i != links_.size(); ++i) {
            Idx2D const m = coup_.branch[branch_seq_link(static_cast<Idx>(i))];
            if (m.group != -1) param[m.group].branch_param[m.pos] = links_[i].calc_param<B>();
        }
        for (size_t i = 0; i != t3w_.size(); ++i) { // main_core/y_bus.hpp:196-204: three branches per Branch3
            auto const& [group, pos] = coup_.branch3[i];
            if (group == -1) continue;
            auto const p3 = t3w_[i].calc_param<B>();
            for (int k = 0; k != 3; ++k) param[group].branch_param[pos[k]] = p3[k];
        }
        for (size_t i = 0; i != transformers_.size(); ++i) {
            Idx2D const m = coup_.branch[branch_seq_transformer(static_cast<Idx>(i))];
            if (m.group != -1) param[m.group].branch_param[m.pos] = transformers_[i].calc_param<B>();
        }
        for (size_t i = 0; i != shunts_.size(); ++i) {
            Idx2D const m = coup_.shunt[i];
            if (m.group != -1) param[m.group].shunt_param[m.pos] = shunts_[i].calc_param<B>();
        }
        for (size_t i = 0; i != sources_.size(); ++i) {
            Idx2D const m = coup_.source[i];
            if (m.group != -1) param[m.group].source_param[m.pos] = sources_[i].math_param();
        }
        return param;
    }

    template <int B> void prepare_solvers() {
        prepare_topology();
        auto& ys = y_bus<B>();
        auto& solvers = math_solvers<B>();
        constexpr int sym_idx = B == 1 ? 0 : 1;
        if (ys.empty() && !math_topo_.empty()) {
            auto params = get_math_param<B>();
            for (size_t g = 0; g != math_topo_.size(); ++g) {
                std::shared_ptr<YBusStructure const> shared;
                if constexpr (B == 1) {
                    if (!y_bus_asym_.empty()) shared = y_bus_asym_[g].shared_structure();
                } else {
                    if (!y_bus_sym_.empty()) shared = y_bus_sym_[g].shared_structure();
                }
                ys.emplace_back(math_topo_[g], std::move(params[g]), shared);
            }
            param_valid_[sym_idx] = true;
        }
        if (solvers.empty()) {
            for (size_t g = 0; g != math_topo_.size(); ++g) solvers.emplace_back(*math_topo_[g]);
        }
        if (!param_valid_[sym_idx]) {
            auto params = get_math_param<B>();
            for (size_t g = 0; g != math_topo_.size(); ++g) {
                ys[g].update_admittance(std::move(params[g]));
                solvers[g].parameters_changed();
            }
            param_valid_[sym_idx] = true;
        }
    }

    template <int B> std::vector<PowerFlowInput<B>> prepare_power_flow_input() const {
        std::vector<PowerFlowInput<B>> in(math_topo_.size());
        for (size_t g = 0; g != math_topo_.size(); ++g) {
            in[g].s_injection.resize(math_topo_[g]->n_load_gen());
            in[g].source.resize(math_topo_[g]->n_source());
        }
        for (size_t i = 0; i != sources_.size(); ++i) {
            Idx2D const m = coup_.source[i];
            if (m.group != -1) in[m.group].source[m.pos] = sources_[i].calc_param();
        }
        for (size_t i = 0; i != load_gens_.size(); ++i) {
            Idx2D const m = coup_.load_gen[i];
            if (m.group != -1) in[m.group].s_injection[m.pos] = load_gens_[i].calc_param<B>();
        }
        if (!regulators_.empty()) {
            for (size_t g = 0; g != math_topo_.size(); ++g) {
                in[g].voltage_regulator.resize(math_topo_[g]->n_voltage_regulator());
                in[g].load_gen_status.resize(math_topo_[g]->n_load_gen());
            }
            for (size_t i = 0; i != regulators_.size(); ++i) {
                Idx2D const m = coup_.voltage_regulator[i];
                if (m.group != -1) in[m.group].voltage_regulator[m.pos] = regulators_[i].calc_param();
            }
            for (size_t i = 0; i != load_gens_.size(); ++i) {
                Idx2D const m = coup_.load_gen[i];
                if (m.group != -1) in[m.group].load_gen_status[m.pos] = static_cast<IntS>(load_gens_[i].status);
            }
        }
        return in;
    }

    template <int B>
    BranchOutput<B> branch_output(BranchBase const& b, BranchSolverOutput<B> const& so, double loading_sn,
                                  double loading_in) const {
        BranchOutput<B> o{};
        o.id = b.id;
        o.energized = (b.from_status || b.to_status) ? 1 : 0;
        double sum_sf = 0.0, sum_st = 0.0, max_if = 0.0, max_it = 0.0;
        for (int p = 0; p < B; ++p) {
            o.p_from[p] = base_power<B> * so.s_f.v[p].real();
            o.q_from[p] = base_power<B> * so.s_f.v[p].imag();
            o.i_from[p] = b.base_i_from * cabs(so.i_f.v[p]);
            o.s_from[p] = base_power<B> * cabs(so.s_f.v[p]);
            o.p_to[p] = base_power<B> * so.s_t.v[p].real();
            o.q_to[p] = base_power<B> * so.s_t.v[p].imag();
            o.i_to[p] = b.base_i_to * cabs(so.i_t.v[p]);
            o.s_to[p] = base_power<B> * cabs(so.s_t.v[p]);
            sum_sf = (p == 0) ? o.s_from[p] : sum_sf + o.s_from[p];
            sum_st = (p == 0) ? o.s_to[p] : sum_st + o.s_to[p];
            max_if = (p == 0) ? o.i_from[p] : std::max(max_if, o.i_from[p]);
            max_it = (p == 0) ? o.i_to[p] : std::max(max_it, o.i_to[p]);
        }
        double const max_s = std::max(sum_sf, sum_st);
        double const max_i = std::max(max_if, max_it);
        o.loading = loading_sn > 0.0 ? max_s / loading_sn : max_i / loading_in;
        return o;
    }

    template <int B>
    void output_result(std::vector<SolverOutput<B>> const& so, BatchOutput<B> const& out, Idx scenario) const {
        Idx const nn = n_node();
        if (out.node != nullptr) {
            // node injection = sum of source / load_gen injections at the user node, order Source, SymLoad, SymGen,
            // AsymLoad, AsymGen (topological_node_output.hpp:98-109)
            std::vector<CVec<B>> inj(nn);
            for (size_t i = 0; i != sources_.size(); ++i) {
                Idx2D const m = coup_.source[i];
                if (m.group != -1) inj[node_idx_.at(sources_[i].node)] += so[m.group].source[m.pos].s;
            }
            auto add_lg = [&](Idx begin, Idx count) {
                for (Idx i = begin; i != begin + count; ++i) {
                    Idx2D const m = coup_.load_gen[i];
                    if (m.group != -1) inj[node_idx_.at(load_gens_[i].node)] += so[m.group].load_gen[m.pos].s;
                }
            };
            Idx const o_sym_gen = 0, o_asym_gen = n_sym_gen_, o_sym_load = n_sym_gen_ + n_asym_gen_,
                      o_asym_load = n_sym_gen_ + n_asym_gen_ + n_sym_load_;
            add_lg(o_sym_load, n_sym_load_);
            add_lg(o_sym_gen, n_sym_gen_);
            add_lg(o_asym_load, n_asym_load_);
            add_lg(o_asym_gen, n_asym_gen_);
            for (Idx i = 0; i != nn; ++i) {
                NodeOutput<B> o{};
                o.id = nodes_[i].id;
                Idx2D const m = coup_.node[i];
                if (m.group != -1) {
                    o.energized = 1;
                    for (int p = 0; p < B; ++p) {
                        cplx const u = so[m.group].u[m.pos].v[p];
                        o.u_pu[p] = cabs(u);
                        o.u[p] = u_scale<B> * nodes_[i].u_rated * o.u_pu[p];
                        o.u_angle[p] = std::arg(u);
                        o.p[p] = base_power<B> * inj[i].v[p].real();
                        o.q[p] = base_power<B> * inj[i].v[p].imag();
                    }
                }
                out.node[scenario * nn + i] = o;
            }
        }
        auto null_branch = [](BranchBase const& b) {
            BranchOutput<B> o{};
            o.id = b.id;
            o.energized = 0;
            return o;
        };
        if (out.line != nullptr) {
            Idx const n = static_cast<Idx>(lines_.size());
            for (Idx i = 0; i != n; ++i) {
                Idx2D const m = coup_.branch[branch_seq_line(i)];
                out.line[scenario * n + i] = m.group == -1
                                                 ? null_branch(lines_[i])
                                                 : branch_output<B>(lines_[i], so[m.group].branch[m.pos], -1.0, lines_[i].i_n);
            }
        }
        if (out.asym_line != nullptr) {
            Idx const n = static_cast<Idx>(asym_lines_.size());
            for (Idx i = 0; i != n; ++i) {
                Idx2D const m = coup_.branch[branch_seq_asym_line(i)];
                out.asym_line[scenario * n + i] =
                    m.group == -1 ? null_branch(asym_lines_[i])
                                  : branch_output<B>(asym_lines_[i], so[m.group].branch[m.pos], -1.0, asym_lines_[i].i_n);
            }
        }
        if (out.link != nullptr) { // Link::loading = 0 (link.hpp:30)
            Idx const n = static_cast<Idx>(links_.size());
            for (Idx i = 0; i != n; ++i) {
                Idx2D const m = coup_.branch[branch_seq_link(i)];
                out.link[scenario * n + i] =
                    m.group == -1 ? null_branch(links_[i])
                                  : branch_output<B>(links_[i], so[m.group].branch[m.pos], std::numeric_limits<double>::infinity(), 0.0);
            }
        }
        if (out.three_winding_transformer != nullptr) { // Branch3::get_output (branch3.hpp:93-122), main_core/output.hpp
            Idx const n = static_cast<Idx>(t3w_.size());
            for (Idx i = 0; i != n; ++i) {
                auto const& t = t3w_[i];
                auto const& [group, pos] = coup_.branch3[i];
                Branch3Output<B> o{};
                o.id = t.in.id;
                if (group != -1) {
                    o.energized = t.energized() ? 1 : 0;
                    double sum_s[3];
                    auto side = [&](int k, double* p, double* q, double* cur, double* sv) {
                        auto const& b = so[group].branch[pos[k]];
                        sum_s[k] = 0.0;
                        for (int ph = 0; ph < B; ++ph) {
                            p[ph] = base_power<B> * b.s_f.v[ph].real();
                            q[ph] = base_power<B> * b.s_f.v[ph].imag();
                            cur[ph] = t.base_i[k] * cabs(b.i_f.v[ph]);
                            sv[ph] = base_power<B> * cabs(b.s_f.v[ph]);
                            sum_s[k] = ph == 0 ? sv[ph] : sum_s[k] + sv[ph];
                        }
                    };
                    side(0, o.p_1, o.q_1, o.i_1, o.s_1);
                    side(1, o.p_2, o.q_2, o.i_2, o.s_2);
                    side(2, o.p_3, o.q_3, o.i_3, o.s_3);
                    o.loading_1 = t.loading_side(0, sum_s[0]);
                    o.loading_2 = t.loading_side(1, sum_s[1]);
                    o.loading_3 = t.loading_side(2, sum_s[2]);
                    o.loading = std::max({o.loading_1, o.loading_2, o.loading_3});
                }
                out.three_winding_transformer[scenario * n + i] = o;
            }
        }
        if (out.generic_branch != nullptr) {
            Idx const n = static_cast<Idx>(generic_branches_.size());
            for (Idx i = 0; i != n; ++i) {
                Idx2D const m = coup_.branch[branch_seq_generic_branch(i)];
                out.generic_branch[scenario * n + i] =
                    m.group == -1 ? null_branch(generic_branches_[i])
                                  : branch_output<B>(generic_branches_[i], so[m.group].branch[m.pos], generic_branches_[i].loading_sn(), 0.0);
            }
        }
        if (out.transformer != nullptr) {
            Idx const n = static_cast<Idx>(transformers_.size());
            for (Idx i = 0; i != n; ++i) {
                Idx2D const m = coup_.branch[branch_seq_transformer(i)];
                out.transformer[scenario * n + i] =
                    m.group == -1 ? null_branch(transformers_[i])
                                  : branch_output<B>(transformers_[i], so[m.group].branch[m.pos], transformers_[i].sn, 0.0);
            }
        }
        if (out.shunt != nullptr) {
            Idx const n = static_cast<Idx>(shunts_.size());
            for (Idx i = 0; i != n; ++i) {
                Idx2D const m = coup_.shunt[i];
                out.shunt[scenario * n + i] = m.group == -1 ? shunts_[i].get_null_output<B>()
                                                            : shunts_[i].get_output<B>(so[m.group].shunt[m.pos], -1.0);
            }
        }
        if (out.source != nullptr) {
            Idx const n = static_cast<Idx>(sources_.size());
            for (Idx i = 0; i != n; ++i) {
                Idx2D const m = coup_.source[i];
                out.source[scenario * n + i] = m.group == -1 ? sources_[i].get_null_output<B>()
                                                             : sources_[i].get_output<B>(so[m.group].source[m.pos], 1.0);
            }
        }
        auto lg_out = [&](ApplianceOutput<B>* dst, Idx begin, Idx count) {
            if (dst == nullptr) return;
            for (Idx k = 0; k != count; ++k) {
                auto const& lg = load_gens_[begin + k];
                Idx2D const m = coup_.load_gen[begin + k];
                dst[scenario * count + k] =
                    m.group == -1 ? lg.get_null_output<B>() : lg.get_output<B>(so[m.group].load_gen[m.pos], lg.direction);
            }
        };
        lg_out(out.sym_gen, 0, n_sym_gen_);
        lg_out(out.asym_gen, n_sym_gen_, n_asym_gen_);
        lg_out(out.sym_load, n_sym_gen_ + n_asym_gen_, n_sym_load_);
        lg_out(out.asym_load, n_sym_gen_ + n_asym_gen_ + n_sym_load_, n_asym_load_);
        if (out.voltage_regulator != nullptr) { // main_core/output.hpp:407-421
            Idx const n = static_cast<Idx>(regulators_.size());
            for (Idx i = 0; i != n; ++i) {
                Idx2D const m = coup_.voltage_regulator[i];
                out.voltage_regulator[scenario * n + i] =
                    m.group == -1 ? regulators_[i].get_null_output() : regulators_[i].get_output(so[m.group].voltage_regulator[m.pos]);
            }
        }
        if (out.transformer_tap_regulator != nullptr) { // main_core/output.hpp:381-396
            Idx const n = static_cast<Idx>(tap_regulators_.size());
            for (Idx i = 0; i != n; ++i) {
                IntS const tap = i < static_cast<Idx>(tap_positions_found_.size()) ? tap_positions_found_[i] : na_IntS;
                out.transformer_tap_regulator[scenario * n + i] =
                    tap == na_IntS ? tap_regulators_[i].get_null_output() : tap_regulators_[i].get_output(tap);
            }
        }
    }

    // ---- update / restore ----
    struct Saved {
        std::vector<std::pair<Idx, Line>> lines;
        std::vector<std::pair<Idx, Transformer>> transformers;
        std::vector<std::pair<Idx, AsymLine>> asym_lines;
        std::vector<std::pair<Idx, GenericBranch>> generic_branches;
        std::vector<std::pair<Idx, Link>> links;
        std::vector<std::pair<Idx, ThreeWindingTransformer>> t3w;
        std::vector<std::pair<Idx, Shunt>> shunts;
        std::vector<std::pair<Idx, Source>> sources;
        std::vector<std::pair<Idx, LoadGen>> load_gens;
        std::vector<std::pair<Idx, VoltageRegulator>> regulators;
        std::vector<std::pair<Idx, TransformerTapRegulator>> tap_regulators;
        bool topo{false}, param{false};
    };
    void mark(bool topo, bool param, Saved& saved) {
        if (topo) {
            topo_valid_ = false;
            saved.topo = true;
        }
        if (param || topo) {
            param_valid_[0] = param_valid_[1] = false;
            saved.param = true;
        }
    }
    template <class T, class Map>
    static Idx find_seq(T const& upd, Idx pos_in_scenario, Idx n_in_scenario, Idx n_component, Map const& map, Idx offset) {
        if (upd.id == na_IntID) {
            if (n_in_scenario != n_component) throw PgmError{"update without ids must cover every element"};
            return offset + pos_in_scenario;
        }
        auto it = map.find(upd.id);
        if (it == map.end()) throw PgmError{"The id cannot be found: " + std::to_string(upd.id) + "\n"};
        return it->second;
    }
    void apply_update(BatchUpdate const& upd, Idx s, Saved& saved) {
        {
            auto [b, e] = upd.line.scenario(s);
            for (auto p = b; p != e; ++p) {
                Idx const i = find_seq(*p, p - b, e - b, static_cast<Idx>(lines_.size()), line_idx_, 0);
                saved.lines.emplace_back(i, lines_[i]);
                bool const changed = lines_[i].set_status(p->from_status, p->to_status);
                mark(changed, changed, saved);
            }
        }
        {
            auto [b, e] = upd.asym_line.scenario(s);
            for (auto p = b; p != e; ++p) {
                Idx const i = find_seq(*p, p - b, e - b, static_cast<Idx>(asym_lines_.size()), asym_line_idx_, 0);
                saved.asym_lines.emplace_back(i, asym_lines_[i]);
                bool const changed = asym_lines_[i].set_status(p->from_status, p->to_status);
                mark(changed, changed, saved);
            }
        }
        {
            auto [b, e] = upd.generic_branch.scenario(s);
            for (auto p = b; p != e; ++p) {
                Idx const i = find_seq(*p, p - b, e - b, static_cast<Idx>(generic_branches_.size()), generic_branch_idx_, 0);
                saved.generic_branches.emplace_back(i, generic_branches_[i]);
                bool const changed = generic_branches_[i].set_status(p->from_status, p->to_status);
                mark(changed, changed, saved);
            }
        }
        {
            auto [b, e] = upd.link.scenario(s);
            for (auto p = b; p != e; ++p) {
                Idx const i = find_seq(*p, p - b, e - b, static_cast<Idx>(links_.size()), link_idx_, 0);
                saved.links.emplace_back(i, links_[i]);
                bool const changed = links_[i].set_status(p->from_status, p->to_status);
                mark(changed, changed, saved);
            }
        }
        {
            auto [b, e] = upd.three_winding_transformer.scenario(s);
            for (auto p = b; p != e; ++p) { // ThreeWindingTransformer::update (three_winding_transformer.hpp:158-163)
                Idx const i = find_seq(*p, p - b, e - b, static_cast<Idx>(t3w_.size()), t3w_idx_, 0);
                saved.t3w.emplace_back(i, t3w_[i]);
                bool const topo = t3w_[i].set_status(p->status_1, p->status_2, p->status_3);
                bool const param = t3w_[i].set_tap(p->tap_pos) || topo;
                mark(topo, param, saved);
            }
        }
        {
            auto [b, e] = upd.transformer.scenario(s);
            for (auto p = b; p != e; ++p) {
                Idx const i = find_seq(*p, p - b, e - b, static_cast<Idx>(transformers_.size()), transformer_idx_, 0);
                saved.transformers.emplace_back(i, transformers_[i]);
                bool const topo = transformers_[i].set_status(p->from_status, p->to_status);
                bool const param = transformers_[i].set_tap(p->tap_pos) || topo;
                mark(topo, param, saved);
            }
        }
        {
            auto [b, e] = upd.shunt.scenario(s);
            for (auto p = b; p != e; ++p) {
                Idx const i = find_seq(*p, p - b, e - b, static_cast<Idx>(shunts_.size()), shunt_idx_, 0);
                saved.shunts.emplace_back(i, shunts_[i]);
                bool changed = shunts_[i].set_status(p->status);
                changed = shunts_[i].update_params(p->g1, p->b1, p->g0, p->b0) || changed;
                mark(false, changed, saved);
            }
        }
        {
            auto [b, e] = upd.source.scenario(s);
            for (auto p = b; p != e; ++p) {
                Idx const i = find_seq(*p, p - b, e - b, static_cast<Idx>(sources_.size()), source_idx_, 0);
                saved.sources.emplace_back(i, sources_[i]);
                auto& src = sources_[i];
                bool const topo = src.set_status(p->status);
                if (!is_nan(p->u_ref)) src.u_ref = p->u_ref;
                if (!is_nan(p->u_ref_angle)) src.u_ref_angle = p->u_ref_angle;
                bool param = false;
                if (!is_nan(p->sk)) {
                    src.sk = p->sk;
                    param = true;
                }
                if (!is_nan(p->rx_ratio)) {
                    src.rx_ratio = p->rx_ratio;
                    param = true;
                }
                if (!is_nan(p->z01_ratio)) {
                    src.z01_ratio = p->z01_ratio;
                    param = true;
                }
                mark(topo, param || topo, saved);
            }
        }
        auto upd_lg = [&](auto const& buffer, Idx offset, Idx count) {
            auto [b, e] = buffer.scenario(s);
            for (auto p = b; p != e; ++p) {
                Idx const i = find_seq(*p, p - b, e - b, count, load_gen_idx_, offset);
                if (i < offset || i >= offset + count) throw PgmError{"The id cannot be found: " + std::to_string(p->id) + "\n"};
                saved.load_gens.emplace_back(i, load_gens_[i]);
                load_gens_[i].set_status(p->status);
                if constexpr (std::is_same_v<std::remove_cvref_t<decltype(*p)>, SymLoadGenUpdate>) {
                    load_gens_[i].set_power(&p->p_specified, &p->q_specified);
                } else {
                    load_gens_[i].set_power(p->p_specified, p->q_specified);
                }
            }
        };
        upd_lg(upd.sym_gen, 0, n_sym_gen_);
        upd_lg(upd.asym_gen, n_sym_gen_, n_asym_gen_);
        upd_lg(upd.sym_load, n_sym_gen_ + n_asym_gen_, n_sym_load_);
        upd_lg(upd.asym_load, n_sym_gen_ + n_asym_gen_ + n_sym_load_, n_asym_load_);
        {
            auto [b, e] = upd.voltage_regulator.scenario(s);
            for (auto p = b; p != e; ++p) {
                Idx const i = find_seq(*p, p - b, e - b, static_cast<Idx>(regulators_.size()), regulator_idx_, 0);
                saved.regulators.emplace_back(i, regulators_[i]);
                regulators_[i].update(*p); // UpdateChange{false, false} (voltage_regulator.hpp:35-41)
            }
        }
        {
            auto [b, e] = upd.transformer_tap_regulator.scenario(s);
            for (auto p = b; p != e; ++p) {
                Idx const i = find_seq(*p, p - b, e - b, static_cast<Idx>(tap_regulators_.size()), tap_regulator_idx_, 0);
                saved.tap_regulators.emplace_back(i, tap_regulators_[i]);
                tap_regulators_[i].update(*p); // UpdateChange{false, false} (transformer_tap_regulator.hpp:41-50)
            }
        }
    }
    void restore(Saved const& saved) {
        // restore in reverse order so repeated updates of one component end at the original value
        for (auto it = saved.lines.rbegin(); it != saved.lines.rend(); ++it) lines_[it->first] = it->second;
        for (auto it = saved.asym_lines.rbegin(); it != saved.asym_lines.rend(); ++it) asym_lines_[it->first] = it->second;
        for (auto it = saved.generic_branches.rbegin(); it != saved.generic_branches.rend(); ++it)
            generic_branches_[it->first] = it->second;
        for (auto it = saved.transformers.rbegin(); it != saved.transformers.rend(); ++it)
            transformers_[it->first] = it->second;
        for (auto it = saved.links.rbegin(); it != saved.links.rend(); ++it) links_[it->first] = it->second;
        for (auto it = saved.t3w.rbegin(); it != saved.t3w.rend(); ++it) t3w_[it->first] = it->second;
        for (auto it = saved.shunts.rbegin(); it != saved.shunts.rend(); ++it) shunts_[it->first] = it->second;
        for (auto it = saved.sources.rbegin(); it != saved.sources.rend(); ++it) sources_[it->first] = it->second;
        for (auto it = saved.load_gens.rbegin(); it != saved.load_gens.rend(); ++it) load_gens_[it->first] = it->second;
        for (auto it = saved.regulators.rbegin(); it != saved.regulators.rend(); ++it) regulators_[it->first] = it->second;
        for (auto it = saved.tap_regulators.rbegin(); it != saved.tap_regulators.rend(); ++it) tap_regulators_[it->first] = it->second;
        if (saved.topo) topo_valid_ = false;
        if (saved.param) param_valid_[0] = param_valid_[1] = false;
    }
};

} // namespace pgm_oracle
