// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the reference power-flow solvers:
//   math_solver/iterative_pf_solver.hpp        run_power_flow :33-90
//   math_solver/newton_raphson_pf_solver.hpp   initialize_derived_solver :255-303, calculate_hnml :462-471,
//       prepare_matrix_and_rhs_from_network_perspective :473-521, add_loads :764-822, add_sources :824-852,
//       iterate_unknown :325-349, set_linear_block :856-863, add_linear_initial_guess_* :706-762
//   math_solver/iterative_current_pf_solver.hpp initialize :95-123, add_loads/add_sources :172-205,
//       make_flat_start :207-225, iterate_unknown :148-160
//   math_solver/linear_pf_solver.hpp           run_power_flow :67-92
//   math_solver/common_solver_functions.hpp    add_sources :33-42, add_linear_loads :44-51, copy_y_bus :68-76,
//       calculate_multiple_source_result :83-139, calculate_source_result :143-160,
//       calculate_load_gen_result :383-409, calculate_pf_result :411-446
//   math_solver/math_solver.hpp                method switch :43-64
// Voltage regulators (PV buses, newton_raphson_pf_solver.hpp:400-453, 549-704: bus types and Q limits, PV rows of the Jacobian,
// limit check from iteration 2 with the PV -> PQ switch, Q allocation in the result step) are restated below and pinned by the
// reference's six pv-node validation cases (tests/test_oracle_validation.py).
#pragma once

#include "sparse_lu.hpp"
#include "ybus.hpp"

namespace pgm_oracle {

enum class CalculationMethod : IntS {
    default_method = -128,
    linear = 0,
    newton_raphson = 1,
    iterative_current = 3,
    linear_current = 4,
};

// ---- result extraction (common_solver_functions.hpp) -------------------------------------------------------
template <int B>
inline void calculate_multiple_source_result(Idx s_begin, Idx s_end, YBus<B> const& y_bus,
                                             PowerFlowInput<B> const& input, CVec<B> const& i_inj_t,
                                             SolverOutput<B>& output, Idx bus) {
    auto const& y_ref = y_bus.param().source_param;
    if constexpr (B == 1) {
        cplx y_ref_t{};
        for (Idx s = s_begin; s != s_end; ++s) y_ref_t += y_ref[s].y1;
        cplx const z_ref_t = 1.0 / y_ref_t;
        cplx i_ref_t{};
        for (Idx s = s_begin; s != s_end; ++s) i_ref_t += input.source[s] * y_ref[s].y1;
        for (Idx s = s_begin; s != s_end; ++s) {
            cplx const ratio = y_ref[s].y1 * z_ref_t;
            cplx const lhs = ratio * (input.source[s] * y_ref_t - i_ref_t);
            output.source[s].i.v[0] = lhs + (ratio * i_inj_t.v[0]);
            output.source[s].s.v[0] = output.u[bus].v[0] * std::conj(output.source[s].i.v[0]);
        }
    } else {
        cplx y_ref_t_012[3]{};
        for (Idx s = s_begin; s != s_end; ++s) {
            y_ref_t_012[0] += y_ref[s].y0;
            y_ref_t_012[1] += y_ref[s].y1;
            y_ref_t_012[2] += y_ref[s].y1;
        }
        cplx i_ref_1_t{};
        for (Idx s = s_begin; s != s_end; ++s) i_ref_1_t += input.source[s] * y_ref[s].y1;
        CVec<3> const i_inj_t_012 = dot(get_sym_matrix_inv(), i_inj_t);
        for (Idx s = s_begin; s != s_end; ++s) {
            cplx const ratio1 = y_ref[s].y1 / y_ref_t_012[1];
            cplx const lhs1 = ratio1 * (input.source[s] * y_ref_t_012[1] - i_ref_1_t);
            CVec<3> i_012;
            i_012.v[0] = (y_ref[s].y0 / y_ref_t_012[0]) * i_inj_t_012.v[0];
            i_012.v[1] = lhs1 + (ratio1 * i_inj_t_012.v[1]);
            i_012.v[2] = (y_ref[s].y1 / y_ref_t_012[2]) * i_inj_t_012.v[2];
            output.source[s].i = dot(get_sym_matrix(), i_012);
            output.source[s].s = output.u[bus] * conj(output.source[s].i);
        }
    }
}

// ---- reactive power of the regulated generators (common_solver_functions.hpp:162-381) ----------------------------------
template <int B> struct RegulatorQState {
    Idx regulator_idx;
    Idx load_gen_idx;
    bool has_available_capacity{false};
    double q_allocated[B]{};
};
template <int B> inline double total_q(double const* q) {
    if constexpr (B == 1) {
        return q[0];
    } else {
        return q[0] + q[1] + q[2];
    }
}
// distribute_q :223-238: keep the per-phase proportions of the base distribution, or split equally when it is ~0
template <int B> inline void distribute_q(double q_scalar, double const* base, double* out) {
    if constexpr (B == 1) {
        (void)base;
        out[0] = q_scalar;
    } else {
        double const base_total = total_q<3>(base);
        if (std::abs(base_total) > numerical_tolerance) {
            double const scale = q_scalar / base_total;
            for (int p = 0; p < 3; ++p) out[p] = base[p] * scale;
        } else {
            for (int p = 0; p < 3; ++p) out[p] = q_scalar / 3.0;
        }
    }
}

template <int B>
inline void calculate_voltage_regulator_result(Idx bus, Idx lg_begin, Idx lg_end, MathTopology const& topo,
                                               PowerFlowInput<B> const& input, SolverOutput<B>& output) {
    if (lg_begin == lg_end) return;
    auto const& vr = topo.voltage_regulators_per_load_gen;
    // 1. regulator outputs + the set of regulating generators (:170-216)
    std::vector<RegulatorQState<B>> state;
    CVec<B> s_load_gen_bus{};
    int num_regulating = 0;
    for (Idx lg = lg_begin; lg != lg_end; ++lg) {
        if (vr[lg] == vr[lg + 1]) {
            s_load_gen_bus += output.load_gen[lg].s;
            continue;
        }
        Idx const regulator = vr[lg]; // std::map::insert keeps the first regulator of the load_gen (:434-438)
        auto& out_reg = output.voltage_regulator[regulator];
        out_reg.generator_id = input.voltage_regulator[regulator].generator_id;
        out_reg.generator_status = input.load_gen_status[lg];
        out_reg.limit_violated = LimitViolation::none;
        bool const is_regulating = input.load_gen_status[lg] != 0 && input.voltage_regulator[regulator].status != 0;
        if (!is_regulating) {
            s_load_gen_bus += output.load_gen[lg].s;
            continue;
        }
        ++num_regulating;
        RegulatorQState<B> st{};
        st.regulator_idx = regulator;
        st.load_gen_idx = lg;
        state.push_back(st);
    }
    if (num_regulating == 0) return;
    // 2. distribution under the regulator limits (:254-270)
    double q_remaining[B];
    for (int p = 0; p < B; ++p) q_remaining[p] = (output.bus_injection[bus].v[p] - s_load_gen_bus.v[p]).imag();
    LimitViolation const bus_limit_violated =
        output.bus_q_limit_violated.empty() ? LimitViolation::none : output.bus_q_limit_violated[bus];
    if (bus_limit_violated == LimitViolation::none) {
        // allocate_q_iterative_distribution (:321-381)
        for (auto& st : state) st.has_available_capacity = true;
        int n_active = num_regulating;
        while (std::abs(total_q<B>(q_remaining)) > numerical_tolerance && n_active > 0) {
            double q_per_regulator[B], q_unallocated[B] = {};
            for (int p = 0; p < B; ++p) q_per_regulator[p] = q_remaining[p] / n_active;
            for (auto& st : state) {
                if (!st.has_available_capacity) continue;
                auto const& in_reg = input.voltage_regulator[st.regulator_idx];
                double q_prev[B], q_next[B];
                for (int p = 0; p < B; ++p) {
                    q_prev[p] = st.q_allocated[p];
                    q_next[p] = q_prev[p] + q_per_regulator[p];
                }
                double const q_next_scalar = total_q<B>(q_next);
                bool const hit_upper = !std::isnan(in_reg.q_max) && q_next_scalar > in_reg.q_max + numerical_tolerance;
                bool const hit_lower =
                    !hit_upper && !std::isnan(in_reg.q_min) && q_next_scalar < in_reg.q_min - numerical_tolerance;
                if (hit_upper || hit_lower) {
                    distribute_q<B>(hit_upper ? in_reg.q_max : in_reg.q_min, q_next, st.q_allocated);
                    st.has_available_capacity = false;
                    for (int p = 0; p < B; ++p) q_unallocated[p] += q_per_regulator[p] - (st.q_allocated[p] - q_prev[p]);
                    n_active -= 1;
                } else {
                    for (int p = 0; p < B; ++p) st.q_allocated[p] = q_next[p];
                }
                output.voltage_regulator[st.regulator_idx].limit_violated = LimitViolation::none;
            }
            double diff[B];
            for (int p = 0; p < B; ++p) diff[p] = q_remaining[p] - q_unallocated[p];
            if (std::abs(total_q<B>(diff)) < numerical_tolerance) {
                throw PgmError{"Unallocated Q remains after distribution on"};
            }
            for (int p = 0; p < B; ++p) q_remaining[p] = q_unallocated[p];
        }
    } else {
        // allocate_q_bus_limit_violated (:293-318)
        for (auto& st : state) {
            auto const& in_reg = input.voltage_regulator[st.regulator_idx];
            output.voltage_regulator[st.regulator_idx].limit_violated = bus_limit_violated;
            double const limit_value = bus_limit_violated == LimitViolation::upper ? in_reg.q_max : in_reg.q_min;
            double const q_limit_scalar = std::isnan(limit_value) ? 0.0 : limit_value;
            double base_q[B];
            for (int p = 0; p < B; ++p) base_q[p] = output.load_gen[st.load_gen_idx].s.v[p].imag();
            distribute_q<B>(q_limit_scalar, base_q, st.q_allocated);
            st.has_available_capacity = false;
        }
    }
    // 3. apply to the generators (:272-289)
    for (auto const& st : state) {
        auto& lg_out = output.load_gen[st.load_gen_idx];
        for (int p = 0; p < B; ++p) lg_out.s.v[p] = cplx{lg_out.s.v[p].real(), st.q_allocated[p]};
        lg_out.i = conj(lg_out.s / output.u[bus]);
    }
}

template <int B, class LoadGenFunc>
inline void calculate_pf_result(YBus<B> const& y_bus, PowerFlowInput<B> const& input, SolverOutput<B>& output,
                                LoadGenFunc load_gen_func) {
    auto const& topo = y_bus.topo();
    output.branch = y_bus.calculate_branch_flow(output.u);
    output.shunt = y_bus.calculate_shunt_flow(output.u);
    output.source.assign(topo.n_source(), {});
    output.load_gen.assign(topo.n_load_gen(), {});
    output.bus_injection.resize(topo.n_bus());
    for (Idx bus = 0; bus != topo.n_bus(); ++bus) output.bus_injection[bus] = y_bus.calculate_injection(output.u, bus);
    output.voltage_regulator.assign(topo.n_voltage_regulator(), {});

    for (Idx bus = 0; bus != topo.n_bus(); ++bus) {
        Idx const lg_begin = topo.load_gens_per_bus[bus], lg_end = topo.load_gens_per_bus[bus + 1];
        Idx const s_begin = topo.sources_per_bus[bus], s_end = topo.sources_per_bus[bus + 1];
        for (Idx lg = lg_begin; lg != lg_end; ++lg) {
            switch (load_gen_func(lg)) {
            case LoadGenType::const_pq:
                output.load_gen[lg].s = input.s_injection[lg];
                break;
            case LoadGenType::const_y:
                output.load_gen[lg].s = input.s_injection[lg] * abs2(output.u[bus]);
                break;
            case LoadGenType::const_i:
                output.load_gen[lg].s = input.s_injection[lg] * cabs(output.u[bus]);
                break;
            default:
                throw PgmError{"unknown load_gen type"};
            }
            output.load_gen[lg].i = conj(output.load_gen[lg].s / output.u[bus]);
        }
        if (!topo.voltage_regulators_per_load_gen.empty()) {
            calculate_voltage_regulator_result<B>(bus, lg_begin, lg_end, topo, input, output);
        }
        if (s_begin == s_end) continue;
        CVec<B> i_load_gen_bus{};
        for (Idx lg = lg_begin; lg != lg_end; ++lg) i_load_gen_bus += output.load_gen[lg].i;
        CVec<B> const i_inj_t = conj(output.bus_injection[bus] / output.u[bus]) - i_load_gen_bus;
        if (s_end - s_begin == 1) {
            output.source[s_begin].i = i_inj_t;
            output.source[s_begin].s = output.u[bus] * conj(output.source[s_begin].i);
        } else {
            calculate_multiple_source_result<B>(s_begin, s_end, y_bus, input, i_inj_t, output, bus);
        }
    }
}

// ---- Newton-Raphson -------------------------------------------------------------------------------------------
// Jacobian block: (2B x 2B) real, column-major, sub-blocks H (0,0) N (0,1) M (1,0) L (1,1)  (block_matrix.hpp:19-93)
template <int B> class NewtonRaphsonPFSolver {
  public:
    static constexpr int N = 2 * B;
    static constexpr int NN = N * N;
    using Solver = SparseLU<double, N>;

    explicit NewtonRaphsonPFSolver(YBus<B> const& y_bus)
        : n_bus_{y_bus.size()},
          data_jac_(y_bus.nnz_lu() * NN),
          x_(n_bus_ * N),
          del_x_pq_(n_bus_ * N),
          solver_{y_bus.structure().row_indptr_lu, y_bus.structure().col_indices_lu, y_bus.structure().diag_lu},
          perm_(n_bus_),
          bus_control_(n_bus_) {}

    SolverOutput<B> run_power_flow(YBus<B> const& y_bus, PowerFlowInput<B> const& input, double err_tol, Idx max_iter,
                                   bool cache_run = false) {
        SolverOutput<B> output;
        output.u.resize(n_bus_);
        double max_dev = std::numeric_limits<double>::infinity();
        initialize(y_bus, input, output);
        Idx num_iter = 0;
        while (max_dev > err_tol || num_iter == 0) {
            if (num_iter++ == max_iter) {
                throw IterationDiverge{max_iter, max_dev, err_tol};
            }
            // prepare_matrix_and_rhs (:306-319)
            build_jacobian_and_rhs(y_bus, input, output.u, false);
            if (limit_check_countdown_ > 0) --limit_check_countdown_;
            bool const buses_switched = (limit_check_countdown_ == 0) && enforce_q_limits(y_bus, input);
            if (buses_switched) build_jacobian_and_rhs(y_bus, input, output.u, true);
            apply_pv_constraints(y_bus);
            solver_.prefactorize_and_solve(data_jac_, perm_, del_x_pq_, del_x_pq_);
            max_dev = iterate_unknown(output.u, err_tol, cache_run);
        }
        output.num_iter = num_iter;
        // finalize_result (:351-360)
        output.bus_q_limit_violated.resize(n_bus_);
        for (Idx i = 0; i != n_bus_; ++i) output.bus_q_limit_violated[i] = bus_control_[i].limit_violated;
        auto const& lgt = y_bus.topo().load_gen_type;
        calculate_pf_result<B>(y_bus, input, output, [&lgt](Idx i) { return lgt[i]; });
        return output;
    }

    // exposed for the parity tests of the GPU kernels
    std::vector<double> const& jacobian() const { return data_jac_; }

  private:
    Idx n_bus_;
    std::vector<double> data_jac_;
    std::vector<double> x_;        // per bus: theta[B], v[B]
    std::vector<double> del_x_pq_; // per bus: p[B], q[B]
    Solver solver_;
    typename Solver::PermArray perm_;
    // PV buses / Q limits (:374-397)
    struct BusControlState {
        BusType type{BusType::pq};
        cplx u_ref{};
        bool has_q_limits{false};
        bool recalc_after_limit_violation{false};
        LimitViolation limit_violated{LimitViolation::none};
        double bus_q_min{0.0};
        double bus_q_max{0.0};
    };
    std::vector<BusControlState> bus_control_;
    Idx limit_check_countdown_{-1}; // -1 no check, 0 check now, > 0 iterations to wait
    std::vector<std::array<double, B>> clamped_q_; // per load_gen, NaN = not clamped

    static bool regulates(YBus<B> const& y_bus, PowerFlowInput<B> const& input, Idx lg, Idx reg) {
        (void)y_bus;
        return input.voltage_regulator[reg].status != 0 && input.load_gen_status[lg] != 0;
    }

    // set_bus_types_and_q_limits (:400-444)
    bool set_bus_types_and_q_limits(YBus<B> const& y_bus, PowerFlowInput<B> const& input) {
        auto const& topo = y_bus.topo();
        bool has_usable_q_limits = false;
        if (topo.voltage_regulators_per_load_gen.empty()) {
            for (Idx bus = 0; bus != n_bus_; ++bus)
                if (topo.sources_per_bus[bus] != topo.sources_per_bus[bus + 1]) bus_control_[bus].type = BusType::slack;
            return false;
        }
        for (Idx bus = 0; bus != n_bus_; ++bus) {
            auto& bc = bus_control_[bus];
            if (topo.sources_per_bus[bus] != topo.sources_per_bus[bus + 1]) {
                bc.type = BusType::slack;
                continue;
            }
            for (Idx lg = topo.load_gens_per_bus[bus]; lg != topo.load_gens_per_bus[bus + 1]; ++lg) {
                if (input.load_gen_status.empty() || input.load_gen_status[lg] == 0) continue;
                for (Idx reg = topo.voltage_regulators_per_load_gen[lg]; reg != topo.voltage_regulators_per_load_gen[lg + 1]; ++reg) {
                    auto const& regulator = input.voltage_regulator[reg];
                    if (regulator.status != 0) {
                        bc.type = BusType::pv;
                        bc.u_ref = regulator.u_ref;
                        bc.bus_q_min += regulator.q_min;
                        bc.bus_q_max += regulator.q_max;
                    }
                }
            }
            if (bc.type == BusType::pv) {
                bc.has_q_limits = !std::isnan(bc.bus_q_min) || !std::isnan(bc.bus_q_max);
                if (bc.has_q_limits) has_usable_q_limits = true;
            }
        }
        return has_usable_q_limits;
    }

    // apply_pv_constraints (:549-587): M row zero, L row zero except diag = V, q mismatch = 0
    void apply_pv_constraints(YBus<B> const& y_bus) {
        auto const& s = y_bus.structure();
        for (Idx row = 0; row != n_bus_; ++row) {
            if (bus_control_[row].type != BusType::pv) continue;
            for (Idx k = s.row_indptr_lu[row]; k != s.row_indptr_lu[row + 1]; ++k) {
                for (int r = 0; r < B; ++r)
                    for (int c = 0; c < B; ++c) {
                        el(k, 1, 0, r, c) = 0.0;
                        el(k, 1, 1, r, c) = 0.0;
                    }
                if (s.col_indices_lu[k] == row)
                    for (int p = 0; p < B; ++p) el(k, 1, 1, p, p) = v(row, p);
            }
            for (int p = 0; p < B; ++p) dq(row, p) = 0.0;
        }
    }

    // enforce_q_limits (:605-704)
    bool enforce_q_limits(YBus<B> const& y_bus, PowerFlowInput<B> const& input) {
        auto const& topo = y_bus.topo();
        bool switched = false;
        for (Idx bus = 0; bus != n_bus_; ++bus) {
            auto& bc = bus_control_[bus];
            if (bc.type != BusType::pv || !bc.has_q_limits) continue;
            double specified_regulating_q[B] = {};
            for (Idx lg = topo.load_gens_per_bus[bus]; lg != topo.load_gens_per_bus[bus + 1]; ++lg) {
                if (input.load_gen_status[lg] == 0) continue;
                for (Idx reg = topo.voltage_regulators_per_load_gen[lg]; reg != topo.voltage_regulators_per_load_gen[lg + 1]; ++reg) {
                    if (input.voltage_regulator[reg].status != 0) {
                        if (topo.load_gen_type[lg] != LoadGenType::const_pq) {
                            throw PgmError{"Unsupported load_gen type for voltage regulators " + std::to_string(reg) + "."}; // newton_raphson_pf_solver.hpp:655
                        }
                        for (int p = 0; p < B; ++p) specified_regulating_q[p] += input.s_injection[lg].v[p].imag();
                    }
                }
            }
            double q_total = 0.0;
            for (int p = 0; p < B; ++p) {
                double const q_required = specified_regulating_q[p] - dq(bus, p);
                q_total = p == 0 ? q_required : q_total + q_required;
            }
            LimitViolation limit = LimitViolation::none;
            if (!std::isnan(bc.bus_q_max) && q_total > bc.bus_q_max + numerical_tolerance) {
                limit = LimitViolation::upper;
            } else if (!std::isnan(bc.bus_q_min) && q_total < bc.bus_q_min - numerical_tolerance) {
                limit = LimitViolation::lower;
            }
            if (limit == LimitViolation::none) continue;
            for (Idx lg = topo.load_gens_per_bus[bus]; lg != topo.load_gens_per_bus[bus + 1]; ++lg) {
                for (Idx reg = topo.voltage_regulators_per_load_gen[lg]; reg != topo.voltage_regulators_per_load_gen[lg + 1]; ++reg) {
                    if (input.load_gen_status[lg] == 0 || input.voltage_regulator[reg].status == 0) continue;
                    double const q_limit_scalar =
                        limit == LimitViolation::upper ? input.voltage_regulator[reg].q_max : input.voltage_regulator[reg].q_min;
                    if constexpr (B == 1) {
                        clamped_q_[lg][0] = q_limit_scalar;
                    } else {
                        double base_q[3];
                        for (int p = 0; p < 3; ++p) base_q[p] = input.s_injection[lg].v[p].imag();
                        double const base_q_total = base_q[0] + base_q[1] + base_q[2];
                        if (std::abs(base_q_total) > numerical_tolerance) {
                            double const scale = q_limit_scalar / base_q_total;
                            for (int p = 0; p < 3; ++p) clamped_q_[lg][p] = base_q[p] * scale;
                        } else {
                            for (int p = 0; p < 3; ++p) clamped_q_[lg][p] = q_limit_scalar / 3.0;
                        }
                    }
                }
            }
            bc.limit_violated = limit;
            bc.recalc_after_limit_violation = true;
            bc.type = BusType::pq;
            switched = true;
        }
        return switched;
    }

    double* blk(Idx k) { return &data_jac_[k * NN]; }
    // element (r, c) of sub-block (br, bc) of block k
    double& el(Idx k, int br, int bc, int r, int c) { return data_jac_[k * NN + (bc * B + c) * N + (br * B + r)]; }
    double& theta(Idx i, int p) { return x_[i * N + p]; }
    double& v(Idx i, int p) { return x_[i * N + B + p]; }
    double& dp(Idx i, int p) { return del_x_pq_[i * N + p]; }
    double& dq(Idx i, int p) { return del_x_pq_[i * N + B + p]; }

    void initialize(YBus<B> const& y_bus, PowerFlowInput<B> const& input, SolverOutput<B>& output) {
        auto const& s = y_bus.structure();
        auto const& topo = y_bus.topo();
        std::fill(data_jac_.begin(), data_jac_.end(), 0.0);
        std::fill(del_x_pq_.begin(), del_x_pq_.end(), 0.0);
        bus_control_.assign(n_bus_, BusControlState{});
        std::array<double, B> nan_q;
        nan_q.fill(std::numeric_limits<double>::quiet_NaN());
        clamped_q_.assign(topo.n_load_gen(), nan_q);
        bool const has_usable_limits = set_bus_types_and_q_limits(y_bus, input);
        limit_check_countdown_ = has_usable_limits ? 2 : -1; // limit_check_at_iteration / no_limit_check (:169-171)
        bool const has_regulators = !topo.voltage_regulators_per_load_gen.empty();
        auto const& ydata = y_bus.admittance();
        for (Idx k = 0; k != y_bus.nnz_lu(); ++k) {
            Idx const ky = s.map_lu_y_bus[k];
            if (ky == -1) continue;
            // set_linear_block: [[G, -B], [B, G]]
            for (int r = 0; r < B; ++r)
                for (int c = 0; c < B; ++c) {
                    double const g = ydata[ky].m[r][c].real();
                    double const b = ydata[ky].m[r][c].imag();
                    el(k, 0, 1, r, c) = -b;
                    el(k, 0, 0, r, c) = g;
                    el(k, 1, 1, r, c) = g;
                    el(k, 1, 0, r, c) = b;
                }
        }
        for (Idx bus = 0; bus != n_bus_; ++bus) {
            Idx const d = s.diag_lu[bus];
            for (Idx lg = topo.load_gens_per_bus[bus]; lg != topo.load_gens_per_bus[bus + 1]; ++lg) {
                // y_load = -conj(s); the specified Q of a regulated load_gen is ignored (:706-742)
                bool is_regulated = false;
                if (has_regulators)
                    for (Idx reg = topo.voltage_regulators_per_load_gen[lg]; reg != topo.voltage_regulators_per_load_gen[lg + 1]; ++reg)
                        if (regulates(y_bus, input, lg, reg)) {
                            is_regulated = true;
                            break;
                        }
                for (int p = 0; p < B; ++p) {
                    cplx const s_in = is_regulated ? cplx{input.s_injection[lg].v[p].real(), 0.0} : input.s_injection[lg].v[p];
                    cplx const y_load = -std::conj(s_in);
                    el(d, 0, 1, p, p) += -y_load.imag();
                    el(d, 0, 0, p, p) += y_load.real();
                    el(d, 1, 1, p, p) += y_load.real();
                    el(d, 1, 0, p, p) += y_load.imag();
                }
            }
            for (Idx src = topo.sources_per_bus[bus]; src != topo.sources_per_bus[bus + 1]; ++src) {
                CMat<B> const y_source = y_bus.param().source_param[src].template y_ref<B>();
                CVec<B> const u_source = cvec_rotated<B>(input.source[src]);
                for (int r = 0; r < B; ++r)
                    for (int c = 0; c < B; ++c) {
                        el(d, 0, 1, r, c) -= y_source.m[r][c].imag();
                        el(d, 0, 0, r, c) += y_source.m[r][c].real();
                        el(d, 1, 1, r, c) += y_source.m[r][c].real();
                        el(d, 1, 0, r, c) += y_source.m[r][c].imag();
                    }
                CVec<B> const i_rhs = dot(y_source, u_source);
                for (int p = 0; p < B; ++p) {
                    dp(bus, p) += i_rhs.v[p].real();
                    dq(bus, p) += i_rhs.v[p].imag();
                }
            }
        }
        solver_.prefactorize_and_solve(data_jac_, perm_, del_x_pq_, del_x_pq_);
        for (Idx i = 0; i != n_bus_; ++i) {
            for (int p = 0; p < B; ++p) {
                output.u[i].v[p] = cplx{dp(i, p), dq(i, p)};
                // set_reference_voltage_for_pv_buses (:446-452)
                if (bus_control_[i].type == BusType::pv) output.u[i].v[p] = bus_control_[i].u_ref * phase_shift(output.u[i].v[p]);
                v(i, p) = cabs(output.u[i].v[p]);
                theta(i, p) = std::arg(output.u[i].v[p]);
            }
        }
    }

    // block = hnml(yij, ui, uj) written into block k; returns nothing
    void set_hnml(double* block, CMat<B> const& yij, CVec<B> const& ui, CVec<B> const& uj) {
        for (int r = 0; r < B; ++r)
            for (int c = 0; c < B; ++c) {
                cplx const pf = (ui.v[r] * std::conj(uj.v[c])) * std::conj(yij.m[r][c]);
                double const h = pf.imag();
                double const n = pf.real();
                block[(0 * B + c) * N + (0 * B + r)] = h;
                block[(1 * B + c) * N + (0 * B + r)] = n;
                block[(0 * B + c) * N + (1 * B + r)] = -n;
                block[(1 * B + c) * N + (1 * B + r)] = h;
            }
    }
    static double bh(double const* b, int r, int c) { return b[(0 * B + c) * N + r]; }
    static double bn(double const* b, int r, int c) { return b[(1 * B + c) * N + r]; }

    void build_jacobian_and_rhs(YBus<B> const& y_bus, PowerFlowInput<B> const& input, std::vector<CVec<B>> const& u,
                                bool partial_rebuild) {
        auto const& s = y_bus.structure();
        auto const& topo = y_bus.topo();
        auto const& ydata = y_bus.admittance();
        for (Idx row = 0; row != n_bus_; ++row) {
            // only the rows of buses that switched from PV to PQ are rebuilt (:477-482)
            if (partial_rebuild && !bus_control_[row].recalc_after_limit_violation) continue;
            for (int p = 0; p < B; ++p) {
                dp(row, p) = 0.0;
                dq(row, p) = 0.0;
            }
            for (Idx k = s.row_indptr_lu[row]; k != s.row_indptr_lu[row + 1]; ++k) {
                Idx const ky = s.map_lu_y_bus[k];
                if (ky == -1) {
                    std::fill(blk(k), blk(k) + NN, 0.0);
                    continue;
                }
                Idx const j = s.col_indices_lu[k];
                set_hnml(blk(k), ydata[ky], u[row], u[j]);
                for (int r = 0; r < B; ++r) {
                    double sn = bn(blk(k), r, 0);
                    double sh = bh(blk(k), r, 0);
                    for (int c = 1; c < B; ++c) {
                        sn += bn(blk(k), r, c);
                        sh += bh(blk(k), r, c);
                    }
                    dp(row, r) -= sn;
                    dq(row, r) -= sh;
                }
            }
            Idx const k = s.diag_lu[row];
            for (int p = 0; p < B; ++p) {
                el(k, 0, 0, p, p) += dq(row, p);
                el(k, 0, 1, p, p) += -dp(row, p);
                el(k, 1, 0, p, p) += -dp(row, p);
                el(k, 1, 1, p, p) += -dq(row, p);
            }
        }
        for (Idx bus = 0; bus != n_bus_; ++bus) {
            if (partial_rebuild) {
                if (!bus_control_[bus].recalc_after_limit_violation) continue;
                bus_control_[bus].recalc_after_limit_violation = false;
            }
            Idx const d = s.diag_lu[bus];
            for (Idx lg = topo.load_gens_per_bus[bus]; lg != topo.load_gens_per_bus[bus + 1]; ++lg) {
                if (!std::isnan(clamped_q_[lg][0])) { // the load_gen hit a Q limit: use the clamped value (:768-776)
                    for (int p = 0; p < B; ++p) {
                        dp(bus, p) += input.s_injection[lg].v[p].real();
                        dq(bus, p) += clamped_q_[lg][p];
                    }
                    continue;
                }
                for (int p = 0; p < B; ++p) {
                    double const ps = input.s_injection[lg].v[p].real();
                    double const qs = input.s_injection[lg].v[p].imag();
                    double const vv = v(bus, p);
                    switch (topo.load_gen_type[lg]) {
                    case LoadGenType::const_pq:
                        dp(bus, p) += ps;
                        dq(bus, p) += qs;
                        break;
                    case LoadGenType::const_y:
                        dp(bus, p) += ps * vv * vv;
                        dq(bus, p) += qs * vv * vv;
                        el(d, 0, 1, p, p) += -ps * 2.0 * vv * vv;
                        el(d, 1, 1, p, p) += -qs * 2.0 * vv * vv;
                        break;
                    case LoadGenType::const_i:
                        dp(bus, p) += ps * vv;
                        dq(bus, p) += qs * vv;
                        el(d, 0, 1, p, p) += -ps * vv;
                        el(d, 1, 1, p, p) += -qs * vv;
                        break;
                    default:
                        throw PgmError{"unknown load_gen type"};
                    }
                }
            }
            for (Idx src = topo.sources_per_bus[bus]; src != topo.sources_per_bus[bus + 1]; ++src) {
                CMat<B> const y_ref = y_bus.param().source_param[src].template y_ref<B>();
                CVec<B> const u_ref = cvec_rotated<B>(input.source[src]);
                double mm[NN], ms[NN];
                set_hnml(mm, y_ref, u[bus], u[bus]);
                set_hnml(ms, -y_ref, u[bus], u_ref);
                double p_cal[B], q_cal[B];
                for (int r = 0; r < B; ++r) {
                    double sn = bn(mm, r, 0) + bn(ms, r, 0);
                    double sh = bh(mm, r, 0) + bh(ms, r, 0);
                    for (int c = 1; c < B; ++c) {
                        sn += bn(mm, r, c) + bn(ms, r, c);
                        sh += bh(mm, r, c) + bh(ms, r, c);
                    }
                    p_cal[r] = sn;
                    q_cal[r] = sh;
                }
                for (int p = 0; p < B; ++p) {
                    mm[(0 * B + p) * N + (0 * B + p)] += -q_cal[p]; // h
                    mm[(1 * B + p) * N + (0 * B + p)] += p_cal[p];  // n
                    mm[(0 * B + p) * N + (1 * B + p)] += p_cal[p];  // m
                    mm[(1 * B + p) * N + (1 * B + p)] += q_cal[p];  // l
                    dp(bus, p) -= p_cal[p];
                    dq(bus, p) -= q_cal[p];
                }
                for (int i = 0; i < NN; ++i) blk(d)[i] += mm[i];
            }
        }
    }

    double iterate_unknown(std::vector<CVec<B>>& u, double err_tol, bool cache_run) {
        double max_dev = 0.0;
        for (Idx i = 0; i != n_bus_; ++i) {
            double dev_bus = 0.0;
            for (int p = 0; p < B; ++p) {
                theta(i, p) += dp(i, p);           // del theta
                v(i, p) += v(i, p) * dq(i, p);     // del v / v
                cplx const u_tmp = v(i, p) * std::exp(cplx{0.0, 1.0} * theta(i, p));
                double const dev = cabs(u_tmp - u[i].v[p]);
                dev_bus = (p == 0) ? dev : std::max(dev_bus, dev);
                u[i].v[p] = u_tmp;
            }
            max_dev = std::max(dev_bus, max_dev);
        }
        if (max_dev <= err_tol && limit_check_countdown_ > 0 && !cache_run) {
            // converged before the limit check happened: force the check in one more iteration (:343-347)
            limit_check_countdown_ = 0;
            return std::numeric_limits<double>::infinity();
        }
        return max_dev;
    }
};

// ---- complex-domain solvers --------------------------------------------------------------------------------
template <int B> inline std::vector<cplx> flatten_cmats(std::vector<CMat<B>> const& m) { // to column-major blocks
    std::vector<cplx> out(m.size() * B * B);
    for (size_t k = 0; k != m.size(); ++k)
        for (int c = 0; c < B; ++c)
            for (int r = 0; r < B; ++r) out[k * B * B + c * B + r] = m[k].m[r][c];
    return out;
}

template <int B> inline std::vector<cplx> copy_y_bus(YBus<B> const& y_bus) {
    auto const& s = y_bus.structure();
    std::vector<cplx> mat(y_bus.nnz_lu() * B * B, cplx{});
    for (Idx k = 0; k != y_bus.nnz_lu(); ++k) {
        Idx const ky = s.map_lu_y_bus[k];
        if (ky == -1) continue;
        for (int c = 0; c < B; ++c)
            for (int r = 0; r < B; ++r) mat[k * B * B + c * B + r] = y_bus.admittance()[ky].m[r][c];
    }
    return mat;
}

template <int B> class IterativeCurrentPFSolver {
  public:
    using Solver = SparseLU<cplx, B>;
    explicit IterativeCurrentPFSolver(YBus<B> const& y_bus)
        : n_bus_{y_bus.size()},
          rhs_u_(n_bus_ * B),
          solver_{y_bus.structure().row_indptr_lu, y_bus.structure().col_indices_lu, y_bus.structure().diag_lu} {}

    // `reuse_factorization = false` reproduces the reference snapshot, where run_power_flow works on a copy of the
    // solver so the factorisation is redone every call (iterative_pf_solver.hpp:36).
    SolverOutput<B> run_power_flow(YBus<B> const& y_bus, PowerFlowInput<B> const& input, double err_tol, Idx max_iter,
                                   bool reuse_factorization = false) {
        auto const& topo = y_bus.topo();
        auto const& s = y_bus.structure();
        SolverOutput<B> output;
        output.u.resize(n_bus_);
        double max_dev = std::numeric_limits<double>::infinity();
        // make_flat_start
        {
            cplx sum_u_ref = 0.0;
            for (Idx bus = 0; bus != n_bus_; ++bus)
                for (Idx src = topo.sources_per_bus[bus]; src != topo.sources_per_bus[bus + 1]; ++src)
                    sum_u_ref += input.source[src] * std::exp(cplx{0.0, 1.0} * -topo.phase_shift[bus]);
            cplx const u_ref = sum_u_ref / static_cast<double>(input.source.size());
            for (Idx i = 0; i != n_bus_; ++i)
                output.u[i] = cvec_rotated<B>(u_ref * std::exp(cplx{0.0, 1.0} * topo.phase_shift[i]));
        }
        if (!factorized_ || !reuse_factorization) {
            mat_data_ = copy_y_bus<B>(y_bus);
            for (Idx bus = 0; bus != n_bus_; ++bus) {
                Idx const d = s.diag_lu[bus];
                for (Idx src = topo.sources_per_bus[bus]; src != topo.sources_per_bus[bus + 1]; ++src) {
                    CMat<B> const y = y_bus.param().source_param[src].template y_ref<B>();
                    for (int c = 0; c < B; ++c)
                        for (int r = 0; r < B; ++r) mat_data_[d * B * B + c * B + r] += y.m[r][c];
                }
            }
            solver_.prefactorize(mat_data_, perm_);
            factorized_ = true;
        }
        Idx num_iter = 0;
        while (max_dev > err_tol || num_iter == 0) {
            if (num_iter++ == max_iter) {
                throw IterationDiverge{max_iter, max_dev, err_tol};
            }
            // prepare rhs
            std::fill(rhs_u_.begin(), rhs_u_.end(), cplx{});
            for (Idx bus = 0; bus != n_bus_; ++bus) {
                CVec<B> rhs{};
                for (Idx lg = topo.load_gens_per_bus[bus]; lg != topo.load_gens_per_bus[bus + 1]; ++lg) {
                    auto const& sinj = input.s_injection[lg];
                    switch (topo.load_gen_type[lg]) {
                    case LoadGenType::const_pq:
                        rhs += conj(sinj / output.u[bus]);
                        break;
                    case LoadGenType::const_y:
                        rhs += conj(sinj) * output.u[bus];
                        break;
                    case LoadGenType::const_i:
                        rhs += conj(sinj * cabs(output.u[bus]) / output.u[bus]);
                        break;
                    default:
                        throw PgmError{"unknown load_gen type"};
                    }
                }
                for (Idx src = topo.sources_per_bus[bus]; src != topo.sources_per_bus[bus + 1]; ++src) {
                    rhs += dot(y_bus.param().source_param[src].template y_ref<B>(), cvec_rotated<B>(input.source[src]));
                }
                for (int p = 0; p < B; ++p) rhs_u_[bus * B + p] = rhs.v[p];
            }
            solver_.solve_with_prefactorized_matrix(mat_data_, perm_, rhs_u_, rhs_u_);
            max_dev = 0.0;
            for (Idx bus = 0; bus != n_bus_; ++bus) {
                double dev_bus = 0.0;
                for (int p = 0; p < B; ++p) {
                    double const dev = cabs(rhs_u_[bus * B + p] - output.u[bus].v[p]);
                    dev_bus = (p == 0) ? dev : std::max(dev_bus, dev);
                    output.u[bus].v[p] = rhs_u_[bus * B + p];
                }
                max_dev = std::max(dev_bus, max_dev);
            }
        }
        output.num_iter = num_iter;
        auto const& lgt = topo.load_gen_type;
        calculate_pf_result<B>(y_bus, input, output, [&lgt](Idx i) { return lgt[i]; });
        return output;
    }
    void parameters_changed() { factorized_ = false; }

  private:
    Idx n_bus_;
    std::vector<cplx> rhs_u_;
    std::vector<cplx> mat_data_;
    Solver solver_;
    typename Solver::PermArray perm_;
    bool factorized_{false};
};

template <int B> class LinearPFSolver {
  public:
    using Solver = SparseLU<cplx, B>;
    explicit LinearPFSolver(YBus<B> const& y_bus)
        : n_bus_{y_bus.size()},
          solver_{y_bus.structure().row_indptr_lu, y_bus.structure().col_indices_lu, y_bus.structure().diag_lu} {}

    SolverOutput<B> run_power_flow(YBus<B> const& y_bus, PowerFlowInput<B> const& input) {
        auto const& topo = y_bus.topo();
        auto const& s = y_bus.structure();
        SolverOutput<B> output;
        output.u.assign(n_bus_, CVec<B>{});
        std::vector<cplx> mat = copy_y_bus<B>(y_bus);
        std::vector<cplx> rhs(n_bus_ * B, cplx{});
        for (Idx bus = 0; bus != n_bus_; ++bus) {
            Idx const d = s.diag_lu[bus];
            for (Idx lg = topo.load_gens_per_bus[bus]; lg != topo.load_gens_per_bus[bus + 1]; ++lg) {
                for (int p = 0; p < B; ++p) mat[d * B * B + p * B + p] += -std::conj(input.s_injection[lg].v[p]);
            }
            for (Idx src = topo.sources_per_bus[bus]; src != topo.sources_per_bus[bus + 1]; ++src) {
                CMat<B> const y = y_bus.param().source_param[src].template y_ref<B>();
                for (int c = 0; c < B; ++c)
                    for (int r = 0; r < B; ++r) mat[d * B * B + c * B + r] += y.m[r][c];
                CVec<B> const yu = dot(y, cvec_rotated<B>(input.source[src]));
                for (int p = 0; p < B; ++p) rhs[bus * B + p] += yu.v[p];
            }
        }
        typename Solver::PermArray perm;
        solver_.prefactorize_and_solve(mat, perm, rhs, rhs);
        for (Idx bus = 0; bus != n_bus_; ++bus)
            for (int p = 0; p < B; ++p) output.u[bus].v[p] = rhs[bus * B + p];
        output.num_iter = 1;
        calculate_pf_result<B>(y_bus, input, output, [](Idx) { return LoadGenType::const_y; });
        return output;
    }

  private:
    Idx n_bus_;
    Solver solver_;
};

// math_solver.hpp:43-64, 124-156
template <int B> class MathSolver {
  public:
    explicit MathSolver(MathTopology const& topo)
        : all_const_y_{std::all_of(topo.load_gen_type.begin(), topo.load_gen_type.end(),
                                   [](LoadGenType x) { return x == LoadGenType::const_y; })} {}

    SolverOutput<B> run_power_flow(PowerFlowInput<B> const& input, double err_tol, Idx max_iter,
                                   CalculationMethod method, YBus<B> const& y_bus, bool reuse_ic_factorization = false,
                                   bool cache_run = false) {
        method = all_const_y_ ? CalculationMethod::linear : method;
        switch (method) {
        case CalculationMethod::default_method:
        case CalculationMethod::newton_raphson:
            if (!nr_) nr_.emplace(y_bus);
            return nr_->run_power_flow(y_bus, input, err_tol, max_iter, cache_run);
        case CalculationMethod::linear:
            if (!lin_) lin_.emplace(y_bus);
            return lin_->run_power_flow(y_bus, input);
        case CalculationMethod::linear_current:
            if (!ic_) ic_.emplace(y_bus);
            return ic_->run_power_flow(y_bus, input, std::numeric_limits<double>::infinity(), 1, reuse_ic_factorization);
        case CalculationMethod::iterative_current:
            if (!ic_) ic_.emplace(y_bus);
            return ic_->run_power_flow(y_bus, input, err_tol, max_iter, reuse_ic_factorization);
        default:
            throw PgmError{"The calculation method is invalid for this calculation!"};
        }
    }
    void parameters_changed() {
        if (ic_) ic_->parameters_changed();
    }

  private:
    bool all_const_y_;
    std::optional<NewtonRaphsonPFSolver<B>> nr_;
    std::optional<LinearPFSolver<B>> lin_;
    std::optional<IterativeCurrentPFSolver<B>> ic_;
};

} // namespace pgm_oracle
