// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the PF-relevant components (physical input -> per-unit parameters -> physical output):
//   auxiliary/input.hpp, update.hpp, output.hpp          struct layouts
//   component/branch.hpp      calc_param_y_sym :197-227, calc_param_y_asym :228-241, get_output :94-111
//   component/line.hpp        ctor :26-36
//   component/transformer.hpp ctor :33-83, transformer_params :181-241, sym/asym_calc_param :243-345
//   component/transformer_utils.hpp tap_adjust_impedance :33-47
//   component/source.hpp      math_param :40-48, calc_param :64
//   component/shunt.hpp       calc_param :37-51, update_params :86-98
//   component/load_gen.hpp    set_power :86-95, sym/asym_calc_param :124-139
//   component/appliance.hpp   get_output :67-93 ; component/node.hpp get_output :37-46
#pragma once

#include "tensor.hpp"
#include "ybus.hpp"

namespace pgm_oracle {

enum class WindingType : IntS { wye = 0, wye_n = 1, delta = 2, zigzag = 3, zigzag_n = 4 };
enum class BranchSide : IntS { from = 0, to = 1 };

// ---- C-API struct layouts (natural alignment) ----
struct NodeInput {
    ID id;
    double u_rated;
};
struct LineInput {
    ID id, from_node, to_node;
    IntS from_status, to_status;
    double r1, x1, c1, tan1, r0, x0, c0, tan0, i_n;
};
struct AsymLineInput { // auxiliary/input.hpp:99-142
    ID id, from_node, to_node;
    IntS from_status, to_status;
    double r_aa, r_ba, r_bb, r_ca, r_cb, r_cc, r_na, r_nb, r_nc, r_nn;
    double x_aa, x_ba, x_bb, x_ca, x_cb, x_cc, x_na, x_nb, x_nc, x_nn;
    double c_aa, c_ba, c_bb, c_ca, c_cb, c_cc, c0, c1, i_n;
};
struct GenericBranchInput { // auxiliary/input.hpp:144-165
    ID id, from_node, to_node;
    IntS from_status, to_status;
    double r1, x1, g1, b1, k, theta, sn;
};
static_assert(sizeof(AsymLineInput) == 248 && sizeof(GenericBranchInput) == 72);
struct LinkInput { // auxiliary/input.hpp: LinkInput = BranchInput
    ID id, from_node, to_node;
    IntS from_status, to_status;
};
struct ThreeWindingTransformerInput { // auxiliary/input.hpp: Branch3Input + ThreeWindingTransformerInput
    ID id, node_1, node_2, node_3;
    IntS status_1, status_2, status_3;
    double u1, u2, u3, sn_1, sn_2, sn_3, uk_12, uk_13, uk_23, pk_12, pk_13, pk_23, i0, p0;
    IntS winding_1, winding_2, winding_3, clock_12, clock_13, tap_side, tap_pos, tap_min, tap_max, tap_nom;
    double tap_size, uk_12_min, uk_12_max, uk_13_min, uk_13_max, uk_23_min, uk_23_max, pk_12_min, pk_12_max, pk_13_min, pk_13_max,
        pk_23_min, pk_23_max, r_grounding_1, x_grounding_1, r_grounding_2, x_grounding_2, r_grounding_3, x_grounding_3;
};
static_assert(sizeof(LinkInput) == 16 && sizeof(ThreeWindingTransformerInput) == 304);
struct ThreeWindingTransformerUpdate { // auxiliary/update.hpp: Branch3Update + tap_pos
    ID id;
    IntS status_1, status_2, status_3, tap_pos;
};
static_assert(sizeof(ThreeWindingTransformerUpdate) == 8);
struct TransformerInput {
    ID id, from_node, to_node;
    IntS from_status, to_status;
    double u1, u2, sn, uk, pk, i0, p0, i0_zero_sequence, p0_zero_sequence;
    IntS winding_from, winding_to, clock, tap_side, tap_pos, tap_min, tap_max, tap_nom;
    double tap_size, uk_min, uk_max, pk_min, pk_max, r_grounding_from, x_grounding_from, r_grounding_to,
        x_grounding_to;
};
struct SourceInput {
    ID id, node;
    IntS status;
    double u_ref, u_ref_angle, sk, rx_ratio, z01_ratio;
};
struct ShuntInput {
    ID id, node;
    IntS status;
    double g1, b1, g0, b0;
};
struct SymLoadGenInput {
    ID id, node;
    IntS status, type;
    double p_specified, q_specified;
};
struct AsymLoadGenInput {
    ID id, node;
    IntS status, type;
    double p_specified[3], q_specified[3];
};
struct BranchUpdate {
    ID id;
    IntS from_status, to_status;
};
struct TransformerUpdate {
    ID id;
    IntS from_status, to_status, tap_pos;
};
struct SourceUpdate {
    ID id;
    IntS status;
    double u_ref, u_ref_angle, sk, rx_ratio, z01_ratio;
};
struct ShuntUpdate {
    ID id;
    IntS status;
    double g1, b1, g0, b0;
};
struct SymLoadGenUpdate {
    ID id;
    IntS status;
    double p_specified, q_specified;
};
struct AsymLoadGenUpdate {
    ID id;
    IntS status;
    double p_specified[3], q_specified[3];
};
template <int B> struct NodeOutput {
    ID id;
    IntS energized;
    double u_pu[B], u[B], u_angle[B], p[B], q[B];
};
template <int B> struct BranchOutput {
    ID id;
    IntS energized;
    double loading;
    double p_from[B], q_from[B], i_from[B], s_from[B], p_to[B], q_to[B], i_to[B], s_to[B];
};
template <int B> struct ApplianceOutput {
    ID id;
    IntS energized;
    double p[B], q[B], i[B], s[B], pf[B];
};
// voltage regulator (auxiliary/input.hpp:492-498, update.hpp:213-219, output.hpp:239-243)
struct VoltageRegulatorInput {
    ID id, regulated_object;
    IntS status;
    double u_ref, q_min, q_max;
};
struct VoltageRegulatorUpdate {
    ID id;
    IntS status;
    double u_ref, q_min, q_max;
};
struct VoltageRegulatorOutput {
    ID id;
    IntS energized;
    IntS limit_violated;
};
static_assert(sizeof(VoltageRegulatorInput) == 40 && sizeof(VoltageRegulatorUpdate) == 32 && sizeof(VoltageRegulatorOutput) == 8);
static_assert(sizeof(LineInput) == 88 && sizeof(TransformerInput) == 168 && sizeof(SourceInput) == 56);
static_assert(sizeof(SymLoadGenUpdate) == 24 && sizeof(AsymLoadGenUpdate) == 56);
template <int B> struct Branch3Output { // auxiliary/output.hpp: Branch3Output<sym>
    ID id;
    IntS energized;
    double loading_1, loading_2, loading_3, loading;
    double p_1[B], q_1[B], i_1[B], s_1[B], p_2[B], q_2[B], i_2[B], s_2[B], p_3[B], q_3[B], i_3[B], s_3[B];
};
static_assert(sizeof(Branch3Output<1>) == 136 && sizeof(Branch3Output<3>) == 328);
static_assert(sizeof(NodeOutput<1>) == 48 && sizeof(NodeOutput<3>) == 128);
static_assert(sizeof(BranchOutput<1>) == 80 && sizeof(BranchOutput<3>) == 208);
static_assert(sizeof(ApplianceOutput<1>) == 48 && sizeof(ApplianceOutput<3>) == 128);

// ---- Branch ----
struct BranchBase {
    ID id{}, from_node{}, to_node{};
    bool from_status{}, to_status{};
    double base_i_from{}, base_i_to{};

    bool branch_status() const { return from_status && to_status; }
    bool set_status(IntS f, IntS t) {
        bool changed = false;
        if (f != na_IntS) {
            changed = changed || (from_status != static_cast<bool>(f));
            from_status = static_cast<bool>(f);
        }
        if (t != na_IntS) {
            changed = changed || (to_status != static_cast<bool>(t));
            to_status = static_cast<bool>(t);
        }
        return changed;
    }
    // branch.hpp:197-227
    BranchCalcParam<1> calc_param_y_sym(cplx const& y_series, cplx const& y_shunt, cplx const& tap_ratio) const {
        double const tap = cabs(tap_ratio);
        BranchCalcParam<1> param{};
        auto& yff = param.value[0].m[0][0];
        auto& yft = param.value[1].m[0][0];
        auto& ytf = param.value[2].m[0][0];
        auto& ytt = param.value[3].m[0][0];
        if (!branch_status()) {
            if (from_status || to_status) {
                cplx branch_shunt;
                if (cabs(y_shunt) < numerical_tolerance) {
                    branch_shunt = 0.0;
                } else {
                    branch_shunt = 0.5 * y_shunt + 1.0 / (1.0 / y_series + 2.0 / y_shunt);
                }
                yff = from_status ? (1.0 / tap / tap) * branch_shunt : 0.0;
                ytt = to_status ? branch_shunt : 0.0;
            }
        } else {
            ytt = y_series + 0.5 * y_shunt;
            yff = (1.0 / tap / tap) * ytt;
            yft = (-1.0 / std::conj(tap_ratio)) * y_series;
            ytf = (-1.0 / tap_ratio) * y_series;
        }
        return param;
    }
    BranchCalcParam<3> calc_param_y_asym(cplx const& y1_series, cplx const& y1_shunt, cplx const& y0_series,
                                         cplx const& y0_shunt, cplx const& tap_ratio) const {
        auto const p1 = calc_param_y_sym(y1_series, y1_shunt, tap_ratio);
        auto const p0 = calc_param_y_sym(y0_series, y0_shunt, tap_ratio);
        BranchCalcParam<3> param{};
        for (int i = 0; i < 4; ++i) {
            cplx const v1 = p1.value[i].m[0][0];
            cplx const v0 = p0.value[i].m[0][0];
            param.value[i] = cmat_sm<3>((2.0 * v1 + v0) / 3.0, (v0 - v1) / 3.0);
        }
        return param;
    }
};

struct Line : BranchBase {
    double i_n{};
    cplx y1_series, y1_shunt, y0_series, y0_shunt;
    Line(LineInput const& in, double system_frequency, double u1, double u2) {
        id = in.id;
        from_node = in.from_node;
        to_node = in.to_node;
        from_status = in.from_status != 0;
        to_status = in.to_status != 0;
        i_n = in.i_n;
        double const base_i = base_power_3p / u1 / sqrt3;
        base_i_from = base_i_to = base_i;
        if (cabs(u1 - u2) > numerical_tolerance) {
            throw PgmError{"Conflicting voltage for line " + std::to_string(id)};
        }
        double const base_y = base_i / (u1 / sqrt3);
        cplx const j{0.0, 1.0};
        y1_series = 1.0 / (in.r1 + j * in.x1) / base_y;
        y1_shunt = 2.0 * pi * system_frequency * in.c1 / base_y * (in.tan1 + j);
        y0_series = 1.0 / (in.r0 + j * in.x0) / base_y;
        y0_shunt = 2.0 * pi * system_frequency * in.c0 / base_y * (in.tan0 + j);
    }
    double loading(double /*max_s*/, double max_i) const { return max_i / i_n; }
    double phase_shift() const { return 0.0; }
    template <int B> BranchCalcParam<B> calc_param() const {
        if (!(from_status || to_status)) return BranchCalcParam<B>{};
        if constexpr (B == 1) {
            return calc_param_y_sym(y1_series, y1_shunt, 1.0);
        } else {
            return calc_param_y_asym(y1_series, y1_shunt, y0_series, y0_shunt, 1.0);
        }
    }
};

// component/generic_branch.hpp:37-93: pi model behind a complex ratio k e^{j theta}; no asymmetric parameters
struct GenericBranch : BranchBase {
    double sn{}, theta{};
    cplx y1_series, y1_shunt, ratio;
    GenericBranch(GenericBranchInput const& in, double u1, double u2) {
        id = in.id;
        from_node = in.from_node;
        to_node = in.to_node;
        from_status = in.from_status != 0;
        to_status = in.to_status != 0;
        sn = in.sn;
        double const k = is_nan(in.k) ? 1.0 : in.k;
        theta = is_nan(in.theta) ? 0.0 : std::fmod(in.theta, 2 * pi);
        base_i_from = base_power_3p / u1 / sqrt3;
        base_i_to = base_power_3p / u2 / sqrt3;
        double const base_y = base_i_to / (u2 / sqrt3);
        cplx const j{0.0, 1.0};
        y1_series = 1.0 / (in.r1 + j * in.x1) / base_y;
        y1_shunt = (in.g1 + j * in.b1) / base_y;
        ratio = k * std::exp(j * theta);
    }
    double phase_shift() const { return theta; }
    double loading_sn() const { return is_nan(sn) ? std::numeric_limits<double>::infinity() : sn; } // NaN sn: loading 0
    template <int B> BranchCalcParam<B> calc_param() const {
        if constexpr (B == 1) {
            if (!(from_status || to_status)) return BranchCalcParam<1>{};
            return calc_param_y_sym(y1_series, y1_shunt, ratio);
        } else {
            throw PgmError{"Function not yet implemented: generic_branch in an asymmetric calculation"};
        }
    }
};

// component/link.hpp:16-39: a branch with the fixed series admittance y_link (common/common.hpp:100-101), no shunt, ratio 1.
// At this snapshot of the reference links are ordinary branches of the math model (main_core/topology.hpp never fills
// link_node_idx, so supernodes::reduce_topology takes its dont_reduce_topology branch).
constexpr double g_link = 1e6 / (base_power_3p / 10e3 / 10e3);
struct Link : BranchBase {
    Link(LinkInput const& in, double u1, double u2) {
        id = in.id;
        from_node = in.from_node;
        to_node = in.to_node;
        from_status = in.from_status != 0;
        to_status = in.to_status != 0;
        base_i_from = base_power_3p / u1 / sqrt3;
        base_i_to = base_power_3p / u2 / sqrt3;
    }
    double phase_shift() const { return 0.0; }
    template <int B> BranchCalcParam<B> calc_param() const {
        if (!(from_status || to_status)) return BranchCalcParam<B>{};
        cplx const y_link{g_link, g_link};
        if constexpr (B == 1) {
            return calc_param_y_sym(y_link, 0.0, 1.0);
        } else {
            return calc_param_y_asym(y_link, 0.0, y_link, 0.0, 1.0);
        }
    }
};

// component/asym_line.hpp:26-127 (+ component/line_utils.hpp kron_reduction, common/matrix_utils.hpp averages)
struct AsymLine : BranchBase {
    double i_n{};
    CMat<3> y_series, y_shunt;
    static CMat<3> sym3(cplx s1, cplx s2, cplx s3, cplx m12, cplx m13, cplx m23) {
        CMat<3> r;
        r.m[0][0] = s1, r.m[1][1] = s2, r.m[2][2] = s3;
        r.m[0][1] = r.m[1][0] = m12;
        r.m[0][2] = r.m[2][0] = m13;
        r.m[1][2] = r.m[2][1] = m23;
        return r;
    }
    // fixed-size 3 x 3 inverse by cofactors: inv(i, j) = cofactor(j, i) / det, det expanded along the first column
    static CMat<3> inv3(CMat<3> const& a) {
        auto cof = [&a](int i, int j) {
            int const i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return a.m[i1][j1] * a.m[i2][j2] - a.m[i1][j2] * a.m[i2][j1];
        };
        cplx const c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
        cplx const det = c0 * a.m[0][0] + c1 * a.m[1][0] + c2 * a.m[2][0];
        cplx const invdet = 1.0 / det;
        CMat<3> r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[i][j] = cof(j, i) * invdet;
        return r;
    }
    AsymLine(AsymLineInput const& in, double system_frequency, double u1, double u2) {
        id = in.id;
        from_node = in.from_node;
        to_node = in.to_node;
        from_status = in.from_status != 0;
        to_status = in.to_status != 0;
        i_n = in.i_n;
        double const base_i = base_power_3p / u1 / sqrt3;
        base_i_from = base_i_to = base_i;
        if (cabs(u1 - u2) > numerical_tolerance) throw PgmError{"Conflicting voltage for line " + std::to_string(id)};
        cplx const j{0.0, 1.0};
        CMat<3> z;
        if (is_nan(in.r_na) && is_nan(in.x_na)) {
            CMat<3> const r = sym3(in.r_aa, in.r_bb, in.r_cc, in.r_ba, in.r_ca, in.r_cb);
            CMat<3> const x = sym3(in.x_aa, in.x_bb, in.x_cc, in.x_ba, in.x_ca, in.x_cb);
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) z.m[a][b] = r.m[a][b] + j * x.m[a][b];
        } else {
            auto zz = [&j](double r, double x) { return cplx{r} + j * cplx{x}; };
            CMat<3> const z_pp = sym3(zz(in.r_aa, in.x_aa), zz(in.r_bb, in.x_bb), zz(in.r_cc, in.x_cc), zz(in.r_ba, in.x_ba),
                                      zz(in.r_ca, in.x_ca), zz(in.r_cb, in.x_cb));
            cplx const z_pn[3] = {zz(in.r_na, in.x_na), zz(in.r_nb, in.x_nb), zz(in.r_nc, in.x_nc)};
            cplx const z_nn_inv = 1.0 / zz(in.r_nn, in.x_nn);
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) z.m[a][b] = z_pp.m[a][b] - (z_pn[a] * z_pn[b]) * z_nn_inv;
        }
        CMat<3> cm;
        if (!is_nan(in.c0) && !is_nan(in.c1)) {
            cplx const sdiag = (2.0 * in.c1 + in.c0) / 3.0, moff = (in.c0 - in.c1) / 3.0;
            cm = sym3(sdiag, sdiag, sdiag, moff, moff, moff);
        } else {
            cm = sym3(in.c_aa, in.c_bb, in.c_cc, in.c_ba, in.c_ca, in.c_cb);
        }
        double const base_y = base_i / (u1 / sqrt3);
        double const inv_base_y = 1 / base_y;
        CMat<3> const zi = inv3(z);
        cplx const w = 2.0 * j * pi * system_frequency;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                y_series.m[a][b] = inv_base_y * zi.m[a][b];
                y_shunt.m[a][b] = inv_base_y * (w * cm.m[a][b]);
            }
    }
    double phase_shift() const { return 0.0; }
    template <int B> BranchCalcParam<B> calc_param() const {
        if constexpr (B == 1) {
            if (!(from_status || to_status)) return BranchCalcParam<1>{};
            auto avg_diag = [](CMat<3> const& a) { return (a.m[0][0] + a.m[1][1] + a.m[2][2]) / 3.0; };
            // matrix_utils.hpp:16-18 as written: (1,2) counted twice, (0,2) not at all
            auto avg_off = [](CMat<3> const& a) { return (a.m[0][1] + a.m[1][2] + a.m[1][0] + a.m[1][2] + a.m[2][0] + a.m[2][1]) / 6.0; };
            return calc_param_y_sym(avg_diag(y_series) - avg_off(y_series), avg_diag(y_shunt) - avg_off(y_shunt), 1.0);
        } else {
            BranchCalcParam<3> param{};
            auto scaled = [](CMat<3> const& a, double f) {
                CMat<3> r;
                for (int i = 0; i < 3; ++i)
                    for (int k = 0; k < 3; ++k) r.m[i][k] = f * a.m[i][k];
                return r;
            };
            auto added = [](CMat<3> const& a, CMat<3> const& b) {
                CMat<3> r;
                for (int i = 0; i < 3; ++i)
                    for (int k = 0; k < 3; ++k) r.m[i][k] = a.m[i][k] + b.m[i][k];
                return r;
            };
            if (!branch_status()) {
                if (from_status || to_status) {
                    CMat<3> branch_shunt{};
                    bool all_above = true;
                    for (int i = 0; i < 3; ++i)
                        for (int k = 0; k < 3; ++k) all_above = all_above && cabs(y_shunt.m[i][k]) >= numerical_tolerance;
                    if (all_above) branch_shunt = added(scaled(y_shunt, 0.5), inv3(added(inv3(y_series), scaled(inv3(y_shunt), 2.0))));
                    if (from_status) param.value[0] = branch_shunt;
                    if (to_status) param.value[3] = branch_shunt;
                }
            } else {
                param.value[3] = added(y_series, scaled(y_shunt, 0.5));
                param.value[0] = param.value[3];
                for (int i = 0; i < 3; ++i)
                    for (int k = 0; k < 3; ++k) param.value[1].m[i][k] = -y_series.m[i][k];
                param.value[2] = param.value[1];
            }
            return param;
        }
    }
};

inline double tap_adjust_impedance(double tap_pos, double tap_min, double tap_max, double tap_nom, double xk,
                                   double xk_min, double xk_max) {
    if (tap_pos <= std::max(tap_nom, tap_max) && tap_pos >= std::min(tap_nom, tap_max)) {
        if (tap_max == tap_nom) return xk;
        double const inc = (xk_max - xk) / (tap_max - tap_nom);
        return xk + (tap_pos - tap_nom) * inc;
    }
    if (tap_min == tap_nom) return xk;
    double const inc = (xk_min - xk) / (tap_min - tap_nom);
    return xk + (tap_pos - tap_nom) * inc;
}

struct Transformer : BranchBase {
    double u1, u2, sn, tap_size, uk, pk, i0, p0, i0_zero_sequence, p0_zero_sequence;
    WindingType winding_from, winding_to;
    IntS clock;
    BranchSide tap_side;
    IntS tap_pos, tap_min, tap_max, tap_nom, tap_direction;
    double uk_min, uk_max, pk_min, pk_max;
    double nominal_ratio;
    cplx z_grounding_from, z_grounding_to;

    static cplx calculate_z_pu(double r, double x, double u) {
        r = is_nan(r) ? 0 : r;
        x = is_nan(x) ? 0 : x;
        double const base_z = u * u / base_power_3p;
        return {r / base_z, x / base_z};
    }
    IntS tap_limit(IntS new_tap) const {
        new_tap = std::min(new_tap, std::max(tap_max, tap_min));
        new_tap = std::max(new_tap, std::min(tap_max, tap_min));
        return new_tap;
    }
    Transformer(TransformerInput const& in, double u1_rated, double u2_rated) {
        id = in.id;
        from_node = in.from_node;
        to_node = in.to_node;
        from_status = in.from_status != 0;
        to_status = in.to_status != 0;
        u1 = in.u1;
        u2 = in.u2;
        sn = in.sn;
        tap_size = in.tap_size;
        uk = in.uk;
        pk = in.pk;
        i0 = in.i0;
        p0 = in.p0;
        i0_zero_sequence = is_nan(in.i0_zero_sequence) ? i0 : in.i0_zero_sequence;
        p0_zero_sequence =
            is_nan(in.p0_zero_sequence) ? p0 + pk * (i0_zero_sequence * i0_zero_sequence - i0 * i0) : in.p0_zero_sequence;
        winding_from = static_cast<WindingType>(in.winding_from);
        winding_to = static_cast<WindingType>(in.winding_to);
        clock = in.clock;
        tap_side = static_cast<BranchSide>(in.tap_side);
        tap_min = in.tap_min;
        tap_max = in.tap_max;
        tap_nom = in.tap_nom == na_IntS ? IntS{0} : in.tap_nom;
        tap_direction = tap_max > tap_min ? IntS{1} : IntS{-1};
        uk_min = is_nan(in.uk_min) ? uk : in.uk_min;
        uk_max = is_nan(in.uk_max) ? uk : in.uk_max;
        pk_min = is_nan(in.pk_min) ? pk : in.pk_min;
        pk_max = is_nan(in.pk_max) ? pk : in.pk_max;
        base_i_from = base_power_3p / u1_rated / sqrt3;
        base_i_to = base_power_3p / u2_rated / sqrt3;
        nominal_ratio = u1_rated / u2_rated;
        z_grounding_from = calculate_z_pu(in.r_grounding_from, in.x_grounding_from, u1_rated);
        z_grounding_to = calculate_z_pu(in.r_grounding_to, in.x_grounding_to, u2_rated);
        if (in.tap_pos == na_IntS) {
            tap_pos = in.tap_nom == na_IntS ? IntS{0} : in.tap_nom;
        } else {
            tap_pos = in.tap_pos;
        }
        bool const clock_is_even = (clock % 2) == 0;
        bool const is_from_wye = winding_from == WindingType::wye || winding_from == WindingType::wye_n;
        bool const is_to_wye = winding_to == WindingType::wye || winding_to == WindingType::wye_n;
        if (clock_is_even != (is_from_wye == is_to_wye)) {
            throw PgmError{"Invalid clock for transformer " + std::to_string(id)};
        }
        clock = static_cast<IntS>((clock % 12 + 12) % 12);
        tap_pos = tap_limit(tap_pos);
    }
    bool set_tap(IntS new_tap) {
        if (new_tap == na_IntS || new_tap == tap_pos) return false;
        tap_pos = tap_limit(new_tap);
        return true;
    }
    double loading(double max_s, double /*max_i*/) const { return max_s / sn; }
    double phase_shift() const { return clock * deg_30; }

    struct Params {
        cplx y_series, y_shunt, y0_shunt;
        double k;
    };
    Params transformer_params() const {
        double const base_y_to = base_i_to * base_i_to / base_power_1p;
        double ru1 = u1, ru2 = u2;
        if (tap_side == BranchSide::from) {
            ru1 += tap_direction * (tap_pos - tap_nom) * tap_size;
        } else {
            ru2 += tap_direction * (tap_pos - tap_nom) * tap_size;
        }
        double const k = (ru1 / ru2) / nominal_ratio;
        double const uk_t = tap_adjust_impedance(tap_pos, tap_min, tap_max, tap_nom, uk, uk_min, uk_max);
        double const pk_t = tap_adjust_impedance(tap_pos, tap_min, tap_max, tap_nom, pk, pk_min, pk_max);
        cplx z_series{};
        double const uk_sign = (uk_t >= 0) ? 1.0 : -1.0;
        double const z_series_abs = cabs(uk_t) * ru2 * ru2 / sn;
        z_series.real(pk_t * ru2 * ru2 / sn / sn);
        double const zi2 = z_series_abs * z_series_abs - z_series.real() * z_series.real();
        z_series.imag(uk_sign * (zi2 > 0.0 ? std::sqrt(zi2) : 0.0));
        cplx const y_series = (1.0 / z_series) / base_y_to;
        auto shunt = [&](double i0_, double p0_) {
            cplx y;
            double const y_abs = i0_ * sn / ru2 / ru2;
            y.real(p0_ / ru2 / ru2);
            double const yi2 = y_abs * y_abs - y.real() * y.real();
            y.imag(yi2 > 0.0 ? -std::sqrt(yi2) : 0.0);
            return y / base_y_to;
        };
        return {y_series, shunt(i0, p0), shunt(i0_zero_sequence, p0_zero_sequence), k};
    }
    template <int B> BranchCalcParam<B> calc_param() const {
        if (!(from_status || to_status)) return BranchCalcParam<B>{};
        cplx const j{0.0, 1.0};
        auto const [y_series, y_shunt, y0_shunt, k] = transformer_params();
        if constexpr (B == 1) {
            return calc_param_y_sym(y_series, y_shunt, k * std::exp(j * (clock * deg_30)));
        } else {
            auto const param1 = calc_param_y_sym(y_series, y_shunt, k * std::exp(j * (clock * deg_30)));
            auto const param2 = calc_param_y_sym(y_series, y_shunt, k * std::exp(j * (-clock * deg_30)));
            BranchCalcParam<1> param0{};
            auto& p0ff = param0.value[0].m[0][0];
            auto& p0tt = param0.value[3].m[0][0];
            using enum WindingType;
            if (winding_from == wye_n && winding_to == wye_n) {
                double phase_shift_0 = 0.0;
                if (clock == 2 || clock == 6 || clock == 10) phase_shift_0 = 6.0 * deg_30;
                cplx const z0_series = 1.0 / y_series + 3.0 * (z_grounding_to + z_grounding_from / k / k);
                cplx const y0_series = 1.0 / z0_series;
                param0 = calc_param_y_sym(y0_series, y0_shunt, k * std::exp(j * phase_shift_0));
            } else if (winding_from == wye_n && from_status) {
                cplx y0 = y0_shunt;
                if (winding_to == delta) y0 += y_series;
                if (y0 != cplx{0.0, 0.0}) {
                    cplx const z0 = 1.0 / y0 + 3.0 * z_grounding_from / k / k;
                    y0 = 1.0 / z0;
                    p0ff = y0 / k / k;
                }
            } else if (winding_to == wye_n && to_status) {
                cplx y0 = y0_shunt;
                if (winding_from == delta) y0 += y_series;
                if (y0 != cplx{0.0, 0.0}) {
                    cplx const z0 = 1.0 / y0 + 3.0 * z_grounding_to;
                    y0 = 1.0 / z0;
                    p0tt = y0;
                }
            }
            if (winding_from == zigzag_n && from_status) {
                cplx const z0_series = (1.0 / y_series) * 0.1 + 3.0 * z_grounding_from / k / k;
                p0ff = (1.0 / z0_series) / k / k;
            }
            if (winding_to == zigzag_n && to_status) {
                cplx const z0_series = (1.0 / y_series) * 0.1 + 3.0 * z_grounding_to;
                p0tt = 1.0 / z0_series;
            }
            double const low_susceptance = -transformer_low_susceptance_ratio * sn / base_power_3p / uk;
            auto zero_seq_available = [](WindingType this_side, WindingType other_side) {
                switch (this_side) {
                case wye_n:
                    return other_side == wye_n || other_side == delta;
                case zigzag_n:
                    return true;
                default:
                    return false;
                }
            };
            if (!zero_seq_available(winding_from, winding_to) && from_status) p0ff += cplx{0.0, low_susceptance};
            if (!zero_seq_available(winding_to, winding_from) && to_status) p0tt += cplx{0.0, low_susceptance};
            CMat<3> const sm = get_sym_matrix();
            CMat<3> const smi = get_sym_matrix_inv();
            BranchCalcParam<3> param;
            for (int i = 0; i != 4; ++i) {
                CMat<3> y012{};
                y012.m[0][0] = param0.value[i].m[0][0];
                y012.m[1][1] = param1.value[i].m[0][0];
                y012.m[2][2] = param2.value[i].m[0][0];
                param.value[i] = dot(dot(sm, y012), smi);
            }
            return param;
        }
    }
};

// component/three_winding_transformer.hpp:28-435 (+ component/branch3.hpp): three two-winding transformers T1, T2, T3 between the
// three nodes and an internal node; uk / pk converted delta -> wye relative to side 1 (:226-280); T1 is a YNyn0 transformer that
// carries i0 / p0, T2 / T3 take the reversed clocks (:301-408)
struct ThreeWindingTransformer {
    ThreeWindingTransformerInput in;
    double u_rated[3];
    bool status[3];
    IntS tap_pos, tap_nom, tap_direction, clock_12, clock_13;
    double base_i[3];

    ThreeWindingTransformer(ThreeWindingTransformerInput const& input, double u1_rated, double u2_rated, double u3_rated)
        : in{input}, u_rated{u1_rated, u2_rated, u3_rated}, status{input.status_1 != 0, input.status_2 != 0, input.status_3 != 0} {
        auto nz = [](double& v, double fallback) { v = is_nan(v) ? fallback : v; };
        tap_nom = in.tap_nom == na_IntS ? IntS{0} : in.tap_nom;
        tap_direction = in.tap_max > in.tap_min ? IntS{1} : IntS{-1};
        nz(in.uk_12_min, in.uk_12), nz(in.uk_12_max, in.uk_12), nz(in.uk_13_min, in.uk_13), nz(in.uk_13_max, in.uk_13);
        nz(in.uk_23_min, in.uk_23), nz(in.uk_23_max, in.uk_23), nz(in.pk_12_min, in.pk_12), nz(in.pk_12_max, in.pk_12);
        nz(in.pk_13_min, in.pk_13), nz(in.pk_13_max, in.pk_13), nz(in.pk_23_min, in.pk_23), nz(in.pk_23_max, in.pk_23);
        for (int k = 0; k != 3; ++k) base_i[k] = base_power_3p / u_rated[k] / sqrt3;
        tap_pos = in.tap_pos == na_IntS ? tap_nom : in.tap_pos;
        auto valid_clock = [](IntS clock, IntS w_a, IntS w_b) { // transformer_utils.hpp is_valid_clock
            auto wye = [](IntS w) { return w == 0 || w == 1; };
            bool const clock_is_even = (clock % 2) == 0;
            return clock_is_even == (wye(w_a) == wye(w_b));
        };
        if (!valid_clock(in.clock_12, in.winding_1, in.winding_2)) throw PgmError{"Invalid clock for transformer " + std::to_string(in.id)};
        if (!valid_clock(in.clock_13, in.winding_1, in.winding_3)) throw PgmError{"Invalid clock for transformer " + std::to_string(in.id)};
        clock_12 = static_cast<IntS>((in.clock_12 % 12 + 12) % 12);
        clock_13 = static_cast<IntS>((in.clock_13 % 12 + 12) % 12);
        tap_pos = tap_limit(tap_pos);
    }
    IntS tap_limit(IntS new_tap) const {
        new_tap = std::min(new_tap, std::max(in.tap_max, in.tap_min));
        new_tap = std::max(new_tap, std::min(in.tap_max, in.tap_min));
        return new_tap;
    }
    bool energized() const { return status[0] || status[1] || status[2]; }
    std::array<double, 3> phase_shift() const { return {0.0, -clock_12 * deg_30, -clock_13 * deg_30}; }
    double loading_side(int k, double s) const { return s / (k == 0 ? in.sn_1 : (k == 1 ? in.sn_2 : in.sn_3)); }
    bool set_status(IntS s1, IntS s2, IntS s3) {
        IntS const v[3] = {s1, s2, s3};
        bool changed = false;
        for (int k = 0; k != 3; ++k) {
            if (v[k] == na_IntS) continue;
            changed = changed || (status[k] != static_cast<bool>(v[k]));
            status[k] = static_cast<bool>(v[k]);
        }
        return changed;
    }
    bool set_tap(IntS new_tap) {
        if (new_tap == na_IntS || new_tap == tap_pos) return false;
        tap_pos = tap_limit(new_tap);
        return true;
    }
    std::array<Transformer, 3> two_winding() const {
        double u1 = in.u1, u2 = in.u2, u3 = in.u3;
        double const du = tap_direction * (tap_pos - tap_nom) * in.tap_size;
        if (in.tap_side == 0) {
            u1 += du;
        } else if (in.tap_side == 1) {
            u2 += du;
        } else {
            u3 += du;
        }
        auto adj = [&](double x, double x_min, double x_max) { return tap_adjust_impedance(tap_pos, in.tap_min, in.tap_max, tap_nom, x, x_min, x_max); };
        double const sn_1 = in.sn_1, sn_2 = in.sn_2, sn_3 = in.sn_3;
        // calculate_uk (:226-252)
        double const uk_12 = adj(in.uk_12, in.uk_12_min, in.uk_12_max) * sn_1 / std::min(sn_1, sn_2);
        double const uk_13 = adj(in.uk_13, in.uk_13_min, in.uk_13_max) * sn_1 / std::min(sn_1, sn_3);
        double const uk_23 = adj(in.uk_23, in.uk_23_min, in.uk_23_max) * sn_1 / std::min(sn_2, sn_3);
        double const uk_t1 = 0.5 * (uk_12 + uk_13 - uk_23);
        double const uk_t2 = 0.5 * (uk_12 + uk_23 - uk_13) * (sn_2 / sn_1);
        double const uk_t3 = 0.5 * (uk_13 + uk_23 - uk_12) * (sn_3 / sn_1);
        // calculate_pk (:254-280)
        double const pk_12 = adj(in.pk_12, in.pk_12_min, in.pk_12_max) * (sn_1 / std::min(sn_1, sn_2)) * (sn_1 / std::min(sn_1, sn_2));
        double const pk_13 = adj(in.pk_13, in.pk_13_min, in.pk_13_max) * (sn_1 / std::min(sn_1, sn_3)) * (sn_1 / std::min(sn_1, sn_3));
        double const pk_23 = adj(in.pk_23, in.pk_23_min, in.pk_23_max) * (sn_1 / std::min(sn_2, sn_3)) * (sn_1 / std::min(sn_2, sn_3));
        double const pk_t1 = 0.5 * (pk_12 + pk_13 - pk_23);
        double const pk_t2 = 0.5 * (pk_12 + pk_23 - pk_13) * (sn_2 / sn_1) * (sn_2 / sn_1);
        double const pk_t3 = 0.5 * (pk_13 + pk_23 - pk_12) * (sn_3 / sn_1) * (sn_3 / sn_1);
        auto make = [&](bool st, double ua, double sn, double uk, double pk, double i0, double p0, IntS w_from, IntS w_to, IntS clock, double rg,
                        double xg, double u_rated_from) {
            TransformerInput t{};
            t.id = 2;
            t.from_node = 0;
            t.to_node = 1;
            t.from_status = st ? 1 : 0;
            t.to_status = 1;
            t.u1 = ua;
            t.u2 = u1;
            t.sn = sn;
            t.uk = uk;
            t.pk = pk;
            t.i0 = i0;
            t.p0 = p0;
            t.i0_zero_sequence = nan;
            t.p0_zero_sequence = nan;
            t.winding_from = w_from;
            t.winding_to = w_to;
            t.clock = clock;
            t.tap_side = 0;
            t.tap_pos = t.tap_min = t.tap_max = t.tap_nom = 0;
            t.tap_size = 0.0;
            t.uk_min = t.uk_max = t.pk_min = t.pk_max = nan;
            t.r_grounding_from = rg;
            t.x_grounding_from = xg;
            t.r_grounding_to = 0.0;
            t.x_grounding_to = 0.0;
            return Transformer{t, u_rated_from, u_rated[0]};
        };
        return {make(status[0], u1, sn_1, uk_t1, pk_t1, in.i0, in.p0, 1, 1, 0, in.r_grounding_1, in.x_grounding_1, u_rated[0]),
                make(status[1], u2, sn_2, uk_t2, pk_t2, 0.0, 0.0, in.winding_2, in.winding_1, static_cast<IntS>(12 - clock_12), in.r_grounding_2,
                     in.x_grounding_2, u_rated[1]),
                make(status[2], u3, sn_3, uk_t3, pk_t3, 0.0, 0.0, in.winding_3, in.winding_1, static_cast<IntS>(12 - clock_13), in.r_grounding_3,
                     in.x_grounding_3, u_rated[2])};
    }
    template <int B> std::array<BranchCalcParam<B>, 3> calc_param() const {
        if (!energized()) return {};
        auto const t = two_winding();
        return {t[0].template calc_param<B>(), t[1].template calc_param<B>(), t[2].template calc_param<B>()};
    }
};

// ---- Appliances ----
struct ApplianceBase {
    ID id{}, node{};
    bool status{};
    double base_i{};
    bool set_status(IntS s) {
        if (s == na_IntS) return false;
        if (static_cast<bool>(s) == status) return false;
        status = static_cast<bool>(s);
        return true;
    }
    template <int B>
    ApplianceOutput<B> get_output(ApplianceSolverOutput<B> const& so, double direction) const {
        ApplianceOutput<B> out{};
        out.id = id;
        out.energized = status ? 1 : 0;
        for (int p = 0; p < B; ++p) {
            out.p[p] = base_power<B> * so.s.v[p].real() * direction;
            out.q[p] = base_power<B> * so.s.v[p].imag() * direction;
            out.s[p] = base_power<B> * cabs(so.s.v[p]);
            out.i[p] = base_i * cabs(so.i.v[p]);
            out.pf[p] = out.s[p] < numerical_tolerance ? 0.0 : out.p[p] / out.s[p];
        }
        return out;
    }
    template <int B> ApplianceOutput<B> get_null_output() const {
        ApplianceOutput<B> out{};
        out.id = id;
        out.energized = 0;
        return out;
    }
};

struct Source : ApplianceBase {
    double u_ref, u_ref_angle, sk, rx_ratio, z01_ratio;
    Source(SourceInput const& in, double u) {
        id = in.id;
        node = in.node;
        status = in.status != 0;
        base_i = base_power_3p / u / sqrt3;
        u_ref = in.u_ref;
        u_ref_angle = is_nan(in.u_ref_angle) ? 0.0 : in.u_ref_angle;
        sk = is_nan(in.sk) ? default_source_sk : in.sk;
        rx_ratio = is_nan(in.rx_ratio) ? default_source_rx_ratio : in.rx_ratio;
        z01_ratio = is_nan(in.z01_ratio) ? default_source_z01_ratio : in.z01_ratio;
    }
    SourceCalcParam math_param() const {
        double const z_abs = base_power_3p / sk;
        double const x1 = z_abs / std::sqrt(rx_ratio * rx_ratio + 1.0);
        double const r1 = x1 * rx_ratio;
        cplx const y1_ref = 1.0 / cplx{r1, x1};
        cplx const y0_ref = y1_ref / z01_ratio;
        return {y1_ref, y0_ref};
    }
    cplx calc_param() const { return u_ref * std::exp(cplx{0.0, 1.0} * u_ref_angle); }
};

struct Shunt : ApplianceBase {
    double base_y{nan}, g1{nan}, b1{nan}, g0{nan}, b0{nan};
    cplx y1{nan}, y0{nan};
    Shunt(ShuntInput const& in, double u) {
        id = in.id;
        node = in.node;
        status = in.status != 0;
        base_i = base_power_3p / u / sqrt3;
        base_y = base_i / (u / sqrt3);
        update_params(in.g1, in.b1, in.g0, in.b0);
    }
    static bool update_param(double value, double& target) {
        if (is_nan(value) || value == target) return false;
        target = value;
        return true;
    }
    bool update_params(double ng1, double nb1, double ng0, double nb0) {
        bool changed = update_param(ng1, g1);
        changed = update_param(nb1, b1) || changed;
        changed = update_param(ng0, g0) || changed;
        changed = update_param(nb0, b0) || changed;
        if (changed) {
            y1 = (g1 + cplx{0.0, 1.0} * b1) / base_y;
            y0 = (g0 + cplx{0.0, 1.0} * b0) / base_y;
        }
        return changed;
    }
    template <int B> CMat<B> calc_param() const {
        if (!status) return CMat<B>{};
        if constexpr (B == 1) {
            return cmat_diag<1>(y1);
        } else {
            return cmat_sm<3>((2.0 * y1 + y0) / 3.0, (y0 - y1) / 3.0);
        }
    }
};

// LoadGen<loadgen symmetry LB, direction>
struct LoadGen : ApplianceBase {
    int lb{1};          // 1 = sym_load/sym_gen, 3 = asym_load/asym_gen
    double direction{}; // +1 generator, -1 load
    LoadGenType type{};
    cplx s_specified[3]{{nan, nan}, {nan, nan}, {nan, nan}};

    LoadGen(ID id_, ID node_, IntS status_, IntS type_, double u, int lb_, double direction_, double const* p,
            double const* q) {
        id = id_;
        node = node_;
        status = status_ != 0;
        base_i = base_power_3p / u / sqrt3;
        lb = lb_;
        direction = direction_;
        type = static_cast<LoadGenType>(type_);
        set_power(p, q);
    }
    void set_power(double const* new_p, double const* new_q) {
        double const scalar = direction / (lb == 1 ? base_power_3p : base_power_1p);
        for (int i = 0; i < lb; ++i) {
            double ps = s_specified[i].real();
            double qs = s_specified[i].imag();
            if (!is_nan(new_p[i])) ps = scalar * new_p[i];
            if (!is_nan(new_q[i])) qs = scalar * new_q[i];
            s_specified[i] = cplx{ps, qs};
        }
    }
    template <int B> CVec<B> calc_param() const {
        if (!status) return CVec<B>{};
        if constexpr (B == 1) {
            if (lb == 1) {
                if (is_nan(s_specified[0].real()) || is_nan(s_specified[0].imag())) return {cplx{nan, nan}};
                return {s_specified[0]};
            }
            // mean_val over the three phases (Eigen mean = sum / 3)
            return {(s_specified[0] + s_specified[1] + s_specified[2]) / 3.0};
        } else {
            if (lb == 1) {
                if (is_nan(s_specified[0].real()) || is_nan(s_specified[0].imag())) return vec_piecewise<cplx, 3>(cplx{nan, nan});
                return vec_piecewise<cplx, 3>(s_specified[0]);
            }
            CVec<3> r;
            for (int i = 0; i < 3; ++i) r.v[i] = s_specified[i];
            return r;
        }
    }
};

// component/voltage_regulator.hpp:22-101, component/regulator.hpp
// component/transformer_tap_regulator.hpp:23-99 + regulator.hpp:15-60; auxiliary/{input,update,output}.hpp for the structs
struct TransformerTapRegulatorInput {
    ID id, regulated_object;
    IntS status, control_side; // ControlSide: from = side_1 = 0, to = side_2 = 1, side_3 = 2
    double u_set, u_band, line_drop_compensation_r, line_drop_compensation_x;
};
struct TransformerTapRegulatorUpdate {
    ID id;
    IntS status;
    double u_set, u_band, line_drop_compensation_r, line_drop_compensation_x;
};
struct TransformerTapRegulatorOutput {
    ID id;
    IntS energized;
    IntS tap_pos;
};
static_assert(sizeof(TransformerTapRegulatorInput) == 48 && sizeof(TransformerTapRegulatorUpdate) == 40 &&
              sizeof(TransformerTapRegulatorOutput) == 8);
struct TransformerTapRegulatorCalcParam { // calculation_parameters.hpp
    double u_set, u_band;
    cplx z_compensation;
    IntS status;
};
struct TransformerTapRegulator {
    ID id{}, regulated_object{};
    bool regulates_branch3{}; // regulated_object_type: branch (transformer) or branch3 (three-winding transformer)
    bool status{};
    IntS control_side{};
    double u_rated{}; // of the node at the control side
    double u_set{}, u_band{}, line_drop_compensation_r{}, line_drop_compensation_x{};
    TransformerTapRegulator(TransformerTapRegulatorInput const& in, bool branch3, double u_rated_control)
        : id{in.id}, regulated_object{in.regulated_object}, regulates_branch3{branch3}, status{in.status != 0},
          control_side{in.control_side}, u_rated{u_rated_control}, u_set{in.u_set}, u_band{in.u_band},
          line_drop_compensation_r{in.line_drop_compensation_r}, line_drop_compensation_x{in.line_drop_compensation_x} {}
    void update(TransformerTapRegulatorUpdate const& u) { // :41-50; Regulator::set_status takes the value as it is
        status = static_cast<bool>(u.status);
        if (!is_nan(u.u_set)) u_set = u.u_set;
        if (!is_nan(u.u_band)) u_band = u.u_band;
        if (!is_nan(u.line_drop_compensation_r)) line_drop_compensation_r = u.line_drop_compensation_r;
        if (!is_nan(u.line_drop_compensation_x)) line_drop_compensation_x = u.line_drop_compensation_x;
    }
    template <int B> TransformerTapRegulatorCalcParam calc_param() const { // :76-87
        double const z_base = u_rated * u_rated / base_power<B>;
        cplx const z{is_nan(line_drop_compensation_r) ? 0.0 : line_drop_compensation_r,
                     is_nan(line_drop_compensation_x) ? 0.0 : line_drop_compensation_x};
        return {u_set / u_rated, u_band / u_rated, z / z_base, static_cast<IntS>(status)};
    }
    TransformerTapRegulatorOutput get_null_output() const { return {id, 0, na_IntS}; }
    TransformerTapRegulatorOutput get_output(IntS tap_pos) const { return {id, 1, tap_pos}; } // a regulator is always energized
};

struct VoltageRegulator {
    ID id{};
    ID regulated_object{};
    bool status{};
    double u_ref{}, q_min{}, q_max{};
    explicit VoltageRegulator(VoltageRegulatorInput const& in)
        : id{in.id}, regulated_object{in.regulated_object}, status{in.status != 0}, u_ref{in.u_ref}, q_min{in.q_min}, q_max{in.q_max} {}
    void update(VoltageRegulatorUpdate const& u) {
        if (u.status != na_IntS) status = u.status != 0;
        if (!is_nan(u.u_ref)) u_ref = u.u_ref;
        if (!is_nan(u.q_min)) q_min = u.q_min;
        if (!is_nan(u.q_max)) q_max = u.q_max;
    }
    VoltageRegulatorCalcParam calc_param() const {
        return {static_cast<IntS>(status), cplx{u_ref, 0.0}, q_min / base_power_3p, q_max / base_power_3p, regulated_object};
    }
    VoltageRegulatorOutput get_output(VoltageRegulatorSolverOutput const& so) const {
        return {id, static_cast<IntS>(status && so.generator_status != 0), static_cast<IntS>(so.limit_violated)};
    }
    VoltageRegulatorOutput get_null_output() const { return {id, 0, 0}; }
};

} // namespace pgm_oracle
