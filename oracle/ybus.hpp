// ORACLE (test infrastructure, NOT product code).
// CPU restatement of
//   power_grid_model/calculation_parameters.hpp (MathModelTopology :160-213, MathModelParam :240-255,
//     PowerFlowInput :270-277, SolverOutput :338-350, SourceCalcParam::y_ref :215-227)
//   power_grid_model/math_solver/y_bus.hpp (counting_sort_element :55-88, YBusStructure :122-293,
//     update_admittance_entries :400-431, calculate_injection :482-496, calculate_branch_flow :501-525,
//     calculate_shunt_flow :531-546)
#pragma once

#include "tensor.hpp"

#include <array>
#include <memory>

namespace pgm_oracle {

using BranchIdx = std::array<Idx, 2>;

enum class LoadGenType : IntS { const_pq = 0, const_y = 1, const_i = 2 };
enum class YBusElementType : IntS { bff = 0, bft = 1, btf = 2, btt = 3, shunt = 4, fill_in_ft = 5, fill_in_tf = 6 };

// grouped index vectors are kept in their sparse (indptr) form
struct MathTopology {
    Idx slack_bus{};
    bool is_radial{};
    std::vector<double> phase_shift;
    std::vector<BranchIdx> branch_bus_idx;
    std::vector<BranchIdx> fill_in;
    IdxVector sources_per_bus;   // indptr, size n_bus + 1
    IdxVector shunts_per_bus;    // indptr
    IdxVector load_gens_per_bus; // indptr
    std::vector<LoadGenType> load_gen_type;
    IdxVector voltage_regulators_per_load_gen; // indptr over the load_gens (empty = no regulators), topology.hpp:594-600

    Idx n_bus() const { return static_cast<Idx>(phase_shift.size()); }
    Idx n_branch() const { return static_cast<Idx>(branch_bus_idx.size()); }
    Idx n_source() const { return sources_per_bus.empty() ? 0 : sources_per_bus.back(); }
    Idx n_shunt() const { return shunts_per_bus.empty() ? 0 : shunts_per_bus.back(); }
    Idx n_load_gen() const { return load_gens_per_bus.empty() ? 0 : load_gens_per_bus.back(); }
    Idx n_voltage_regulator() const { return voltage_regulators_per_load_gen.empty() ? 0 : voltage_regulators_per_load_gen.back(); }
};

struct SourceCalcParam {
    cplx y1;
    cplx y0;
    template <int B> CMat<B> y_ref() const {
        if constexpr (B == 1) {
            return cmat_diag<1>(y1);
        } else {
            return cmat_sm<3>((2.0 * y1 + y0) / 3.0, (y0 - y1) / 3.0);
        }
    }
};

template <int B> struct BranchCalcParam {
    CMat<B> value[4]; // yff, yft, ytf, ytt
};

template <int B> struct MathParam {
    std::vector<BranchCalcParam<B>> branch_param;
    std::vector<CMat<B>> shunt_param;
    std::vector<SourceCalcParam> source_param;
};

enum class BusType : IntS { pq = 0, pv = 1, slack = 2 };
enum class LimitViolation : IntS { none = 0, lower = 1, upper = 2 };
struct VoltageRegulatorCalcParam { // calculation_parameters.hpp:228-236
    IntS status{};
    cplx u_ref{};
    double q_min{};
    double q_max{};
    ID generator_id{};
};
struct VoltageRegulatorSolverOutput { // calculation_parameters.hpp:94-100
    LimitViolation limit_violated{};
    ID generator_id{};
    IntS generator_status{};
};
template <int B> struct PowerFlowInput {
    std::vector<cplx> source;         // u_ref of each source
    std::vector<CVec<B>> s_injection; // specified power of each load_gen
    std::vector<VoltageRegulatorCalcParam> voltage_regulator; // math order (grouped by load_gen)
    std::vector<IntS> load_gen_status;                        // only filled when the grid has voltage regulators
};

template <int B> struct BranchSolverOutput {
    CVec<B> s_f, s_t, i_f, i_t;
};
template <int B> struct ApplianceSolverOutput {
    CVec<B> s, i;
};
template <int B> struct SolverOutput {
    std::vector<CVec<B>> u;
    std::vector<CVec<B>> bus_injection;
    std::vector<BranchSolverOutput<B>> branch;
    std::vector<ApplianceSolverOutput<B>> source;
    std::vector<ApplianceSolverOutput<B>> shunt;
    std::vector<ApplianceSolverOutput<B>> load_gen;
    std::vector<LimitViolation> bus_q_limit_violated; // BusSolverOutput (calculation_parameters.hpp:46-49)
    std::vector<VoltageRegulatorSolverOutput> voltage_regulator;
    Idx num_iter{}; // the reference only logs this (iterative_pf_solver.hpp:87)
};

struct YBusElement {
    YBusElementType element_type{};
    Idx idx{};
};

struct YBusStructure {
    IdxVector row_indptr, col_indices;
    std::vector<YBusElement> y_bus_element;
    IdxVector y_bus_entry_indptr;
    IdxVector bus_entry;
    IdxVector row_indptr_lu, col_indices_lu, diag_lu, map_lu_y_bus, lu_transpose_entry;

    struct ElementMap {
        Idx row, col;
        YBusElement element;
    };

    static void counting_sort_element(std::vector<ElementMap>& vec, Idx n_bus) {
        std::vector<ElementMap> temp(vec.size());
        IdxVector counter(n_bus, 0);
        for (auto const& e : vec) ++counter[e.col];
        for (size_t i = 1; i < counter.size(); ++i) counter[i] += counter[i - 1];
        for (auto it = vec.rbegin(); it != vec.rend(); ++it) temp[--counter[it->col]] = *it;
        vec.swap(temp);
        std::fill(counter.begin(), counter.end(), 0);
        for (auto const& e : vec) ++counter[e.row];
        for (size_t i = 1; i < counter.size(); ++i) counter[i] += counter[i - 1];
        for (auto it = vec.rbegin(); it != vec.rend(); ++it) temp[--counter[it->row]] = *it;
        vec.swap(temp);
    }

    explicit YBusStructure(MathTopology const& topo) {
        Idx const n_bus = topo.n_bus();
        Idx const n_branch = topo.n_branch();
        Idx const n_fill_in = static_cast<Idx>(topo.fill_in.size());
        std::vector<ElementMap> vec;
        vec.reserve(4 * n_branch + n_bus + 2 * n_fill_in);
        auto append = [&vec](Idx b1, Idx b2, YBusElementType t, Idx idx) {
            if (b1 == -1 || b2 == -1) return;
            vec.push_back({b1, b2, {t, idx}});
        };
        std::vector<std::array<Idx, 2>> off_diag_map(n_branch + n_fill_in, {0, 0});
        for (Idx branch = 0; branch != n_branch; ++branch) {
            for (int i = 0; i != 4; ++i) {
                append(topo.branch_bus_idx[branch][i / 2], topo.branch_bus_idx[branch][i % 2],
                       static_cast<YBusElementType>(i), branch);
            }
        }
        for (Idx bus = 0; bus != n_bus; ++bus) {
            for (Idx shunt = topo.shunts_per_bus[bus]; shunt != topo.shunts_per_bus[bus + 1]; ++shunt) {
                append(bus, bus, YBusElementType::shunt, shunt);
            }
        }
        for (Idx f = 0; f != n_fill_in; ++f) {
            append(topo.fill_in[f][0], topo.fill_in[f][1], YBusElementType::fill_in_ft, f);
            append(topo.fill_in[f][1], topo.fill_in[f][0], YBusElementType::fill_in_tf, f);
        }
        counting_sort_element(vec, n_bus);

        Idx nnz_counter = 0, row_start = 0, nnz_counter_lu = 0, row_start_lu = 0, fill_in_counter = 0;
        row_indptr.assign(n_bus + 1, 0);
        row_indptr_lu.assign(n_bus + 1, 0);
        bus_entry.assign(n_bus, 0);
        diag_lu.assign(n_bus, 0);
        y_bus_entry_indptr.push_back(0);
        auto is_fill = [](ElementMap const& m) {
            return m.element.element_type == YBusElementType::fill_in_ft ||
                   m.element.element_type == YBusElementType::fill_in_tf;
        };
        for (auto const& m : vec) {
            if (!is_fill(m)) y_bus_element.push_back(m.element);
        }
        for (size_t it = 0; it < vec.size();) {
            Idx const row = vec[it].row;
            Idx const col = vec[it].col;
            col_indices_lu.push_back(col);
            if (row > row_start_lu) row_indptr_lu[++row_start_lu] = nnz_counter_lu;
            if (!is_fill(vec[it])) {
                col_indices.push_back(col);
                map_lu_y_bus.push_back(nnz_counter);
                if (row > row_start) row_indptr[++row_start] = nnz_counter;
                if (row == col) {
                    bus_entry[row] = nnz_counter;
                    diag_lu[row] = nnz_counter_lu;
                } else {
                    off_diag_map[vec[it].element.idx][static_cast<Idx>(vec[it].element.element_type) - 1] =
                        nnz_counter_lu;
                }
                for (; it < vec.size() && vec[it].row == row && vec[it].col == col; ++it) {
                }
                y_bus_entry_indptr.push_back(static_cast<Idx>(it) - fill_in_counter);
                ++nnz_counter;
                ++nnz_counter_lu;
            } else {
                off_diag_map[vec[it].element.idx + n_branch][static_cast<Idx>(vec[it].element.element_type) - 5] =
                    nnz_counter_lu;
                map_lu_y_bus.push_back(-1);
                ++fill_in_counter;
                ++nnz_counter_lu;
                ++it;
            }
        }
        row_indptr[++row_start] = nnz_counter;
        row_indptr_lu[++row_start_lu] = nnz_counter_lu;
        if (topo.n_branch() == 0 && topo.n_shunt() == 0) {
            nnz_counter = 1;
            nnz_counter_lu = 1;
            row_indptr = {0, 1};
            col_indices = {0};
            bus_entry = {0};
            lu_transpose_entry = {0};
            y_bus_entry_indptr = {0, 0};
            row_indptr_lu = {0, 1};
            col_indices_lu = {0};
            diag_lu = {0};
            map_lu_y_bus = {0};
        }
        lu_transpose_entry.resize(nnz_counter_lu);
        for (Idx i = 0; i != nnz_counter_lu; ++i) lu_transpose_entry[i] = i;
        for (auto const& [e1, e2] : off_diag_map) {
            lu_transpose_entry[e1] = e2;
            lu_transpose_entry[e2] = e1;
        }
    }
};

template <int B> class YBus {
  public:
    YBus(std::shared_ptr<MathTopology const> topo, MathParam<B> param,
         std::shared_ptr<YBusStructure const> structure = {})
        : topo_{std::move(topo)},
          sp_{structure ? std::move(structure) : std::make_shared<YBusStructure const>(*topo_)} {
        update_admittance(std::move(param));
    }
    YBus(MathTopology const& topo, MathParam<B> param)
        : YBus{std::make_shared<MathTopology const>(topo), std::move(param)} {}

    Idx size() const { return static_cast<Idx>(sp_->bus_entry.size()); }
    Idx nnz() const { return sp_->row_indptr.back(); }
    Idx nnz_lu() const { return sp_->row_indptr_lu.back(); }
    YBusStructure const& structure() const { return *sp_; }
    std::shared_ptr<YBusStructure const> shared_structure() const { return sp_; }
    MathTopology const& topo() const { return *topo_; }
    MathParam<B> const& param() const { return param_; }
    std::vector<CMat<B>> const& admittance() const { return admittance_; }

    void update_admittance(MathParam<B> param) {
        param_ = std::move(param);
        admittance_.assign(nnz(), CMat<B>{});
        for (Idx entry = 0; entry != nnz(); ++entry) {
            CMat<B> y{};
            for (Idx e = sp_->y_bus_entry_indptr[entry]; e != sp_->y_bus_entry_indptr[entry + 1]; ++e) {
                auto const& c = sp_->y_bus_element[e];
                if (c.element_type == YBusElementType::shunt) {
                    y += param_.shunt_param[c.idx];
                } else {
                    y += param_.branch_param[c.idx].value[static_cast<int>(c.element_type)];
                }
            }
            admittance_[entry] = y;
        }
    }

    CVec<B> calculate_injection(std::vector<CVec<B>> const& u, Idx bus) const {
        CVec<B> i_inj{};
        for (Idx k = sp_->row_indptr[bus]; k != sp_->row_indptr[bus + 1]; ++k) {
            i_inj += dot(admittance_[k], u[sp_->col_indices[k]]);
        }
        return conj(i_inj) * u[bus];
    }
    std::vector<BranchSolverOutput<B>> calculate_branch_flow(std::vector<CVec<B>> const& u) const {
        std::vector<BranchSolverOutput<B>> out(topo_->n_branch());
        for (Idx b = 0; b != topo_->n_branch(); ++b) {
            auto const [f, t] = topo_->branch_bus_idx[b];
            CVec<B> const uf = f != -1 ? u[f] : CVec<B>{};
            CVec<B> const ut = t != -1 ? u[t] : CVec<B>{};
            auto const& p = param_.branch_param[b];
            out[b].i_f = dot(p.value[0], uf) + dot(p.value[1], ut);
            out[b].i_t = dot(p.value[2], uf) + dot(p.value[3], ut);
            out[b].s_f = uf * conj(out[b].i_f);
            out[b].s_t = ut * conj(out[b].i_t);
        }
        return out;
    }
    std::vector<ApplianceSolverOutput<B>> calculate_shunt_flow(std::vector<CVec<B>> const& u) const {
        std::vector<ApplianceSolverOutput<B>> out(topo_->n_shunt());
        for (Idx bus = 0; bus != topo_->n_bus(); ++bus) {
            for (Idx sh = topo_->shunts_per_bus[bus]; sh != topo_->shunts_per_bus[bus + 1]; ++sh) {
                CVec<B> const yu = dot(param_.shunt_param[sh], u[bus]);
                for (int p = 0; p < B; ++p) out[sh].i.v[p] = -yu.v[p];
                out[sh].s = u[bus] * conj(out[sh].i);
            }
        }
        return out;
    }

  private:
    std::shared_ptr<MathTopology const> topo_;
    std::shared_ptr<YBusStructure const> sp_;
    MathParam<B> param_;
    std::vector<CMat<B>> admittance_;
};

} // namespace pgm_oracle
