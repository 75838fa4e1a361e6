// Physical component data -> per-unit calculation parameters and back to physical outputs (host).
// Formulas follow the reference operation by operation so that parameters are bit-identical:
//   Branch::calc_param_y_sym / calc_param_y_asym   component/branch.hpp:197-245
//   Line                                           component/line.hpp:26-58
//   Transformer::transformer_params / *_calc_param component/transformer.hpp:181-345, transformer_utils.hpp:33-47
//   Source::math_param / calc_param                component/source.hpp:40-48, 64
//   Shunt::calc_param / update_params              component/shunt.hpp:37-51, 86-98
//   LoadGen::set_power / calc_param                component/load_gen.hpp:86-95, 124-139
// Struct layouts are the reference's dataset structs (auxiliary/input.hpp, update.hpp, output.hpp).
// Design: plain records + free functions (no class hierarchy, no virtual dispatch); a parameter is a small complex
// tensor written straight into the flat arrays the engine uploads.
#pragma once

#include <cmath>
#include <complex>
#include <cstdint>
#include <limits>
#include <numbers>

namespace pgmb {

using ID = int32_t;
using IntS = int8_t;
using cplx = std::complex<double>;

constexpr double kNaN = std::numeric_limits<double>::quiet_NaN();
constexpr IntS kNaIntS = std::numeric_limits<IntS>::min();
constexpr ID kNaID = std::numeric_limits<ID>::min();
constexpr double kSqrt3 = std::numbers::sqrt3;
constexpr double kPi = std::numbers::pi;
constexpr double kDeg30 = (1.0 / 6.0) * kPi;
constexpr double kBasePower3p = 1e6;
constexpr double kBasePower1p = kBasePower3p / 3.0;
constexpr double kNumTol = 1e-8;
constexpr cplx kA2{-0.5, -kSqrt3 / 2.0};
constexpr cplx kA{-0.5, kSqrt3 / 2.0};

// ---- dataset structs ------------------------------------------------------------------------------------------
struct NodeInput {
    ID id;
    double u_rated;
};
struct LineInput {
    ID id, from_node, to_node;
    IntS from_status, to_status;
    double r1, x1, c1, tan1, r0, x0, c0, tan0, i_n;
};
struct TransformerInput {
    ID id, from_node, to_node;
    IntS from_status, to_status;
    double u1, u2, sn, uk, pk, i0, p0, i0_zero_sequence, p0_zero_sequence;
    IntS winding_from, winding_to, clock, tap_side, tap_pos, tap_min, tap_max, tap_nom;
    double tap_size, uk_min, uk_max, pk_min, pk_max, r_grounding_from, x_grounding_from, r_grounding_to, x_grounding_to;
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
// asym_line / generic_branch (auxiliary/input.hpp:99-165); both are updated with BranchUpdate and report BranchOutput
struct AsymLineInput {
    ID id, from_node, to_node;
    IntS from_status, to_status;
    double r_aa, r_ba, r_bb, r_ca, r_cb, r_cc, r_na, r_nb, r_nc, r_nn;
    double x_aa, x_ba, x_bb, x_ca, x_cb, x_cc, x_na, x_nb, x_nc, x_nn;
    double c_aa, c_ba, c_bb, c_ca, c_cb, c_cc, c0, c1, i_n;
};
struct GenericBranchInput {
    ID id, from_node, to_node;
    IntS from_status, to_status;
    double r1, x1, g1, b1, k, theta, sn;
};
static_assert(sizeof(AsymLineInput) == 248 && sizeof(GenericBranchInput) == 72);
// voltage regulator (auxiliary/input.hpp:492-498, update.hpp:213-219, output.hpp:239-243)
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
struct ThreeWindingTransformerUpdate { // auxiliary/update.hpp: Branch3Update + tap_pos
    ID id;
    IntS status_1, status_2, status_3, tap_pos;
};
static_assert(sizeof(LinkInput) == 16 && sizeof(ThreeWindingTransformerInput) == 304 && sizeof(ThreeWindingTransformerUpdate) == 8);
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
// component/transformer_tap_regulator.hpp:23-99 (+ regulator.hpp): the regulator of the automatic tap changer.  control_side is a
// ControlSide (from = side_1 = 0, to = side_2 = 1, side_3 = 2)
struct TransformerTapRegulatorInput {
    ID id, regulated_object;
    IntS status, control_side;
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
struct TapRegulatorState { // what TransformerTapRegulator::update changes (transformer_tap_regulator.hpp:41-50)
    bool status;
    double u_set, u_band, line_drop_compensation_r, line_drop_compensation_x;
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
template <int B> struct Branch3Output { // auxiliary/output.hpp: Branch3Output<sym>
    ID id;
    IntS energized;
    double loading_1, loading_2, loading_3, loading;
    double p_1[B], q_1[B], i_1[B], s_1[B], p_2[B], q_2[B], i_2[B], s_2[B], p_3[B], q_3[B], i_3[B], s_3[B];
};
static_assert(sizeof(Branch3Output<1>) == 136 && sizeof(Branch3Output<3>) == 328);
template <int B> struct ApplianceOutput {
    ID id;
    IntS energized;
    double p[B], q[B], i[B], s[B], pf[B];
};
static_assert(sizeof(NodeInput) == 16 && sizeof(LineInput) == 88 && sizeof(TransformerInput) == 168);
static_assert(sizeof(SourceInput) == 56 && sizeof(ShuntInput) == 48 && sizeof(SymLoadGenInput) == 32);
static_assert(sizeof(AsymLoadGenInput) == 64 && sizeof(SymLoadGenUpdate) == 24 && sizeof(AsymLoadGenUpdate) == 56);
static_assert(sizeof(NodeOutput<1>) == 48 && sizeof(NodeOutput<3>) == 128 && sizeof(BranchOutput<1>) == 80);
static_assert(sizeof(BranchOutput<3>) == 208 && sizeof(ApplianceOutput<1>) == 48 && sizeof(ApplianceOutput<3>) == 128);

// ---- mutable state of the components (what updates may change) --------------------------------------------------------
struct BranchState {
    bool from_status, to_status;
};
struct TransformerState {
    IntS tap_pos;
};
struct SourceState {
    bool status;
    double u_ref, u_ref_angle, sk, rx_ratio, z01_ratio;
};
struct ShuntState {
    bool status;
    double g1, b1, g0, b0;
    cplx y1, y0;
};
struct LoadGenState {
    bool status;
    cplx s[3]; // per-unit specified power in injection direction; [0] only for symmetric loads
};
struct RegulatorState { // component/voltage_regulator.hpp:22-101
    bool status;
    double u_ref, q_min, q_max; // q in VAr (3-phase), divided by the base power in calc_param
};

// write a B x B complex tensor (row-major, interleaved) --------------------------------------------------------------
template <int B> inline void put_scalar_tensor(double* out, cplx const& s, cplx const& m) {
    for (int r = 0; r != B; ++r)
        for (int c = 0; c != B; ++c) {
            cplx const v = (r == c) ? s : m;
            out[2 * (r * B + c)] = v.real();
            out[2 * (r * B + c) + 1] = v.imag();
        }
}

// ---- branches -----------------------------------------------------------------------------------------------------
struct SymBranchParam {
    cplx yff{}, yft{}, ytf{}, ytt{};
};

// pi-model seen from the "to" side with complex tap ratio (branch.hpp:197-227)
inline SymBranchParam branch_pi_model(bool from_status, bool to_status, cplx const& y_series, cplx const& y_shunt,
                                      cplx const& tap_ratio) {
    double const tap = std::sqrt(std::norm(tap_ratio));
    SymBranchParam p;
    if (!(from_status && to_status)) {
        if (from_status || to_status) {
            cplx branch_shunt;
            if (std::sqrt(std::norm(y_shunt)) < kNumTol) {
                branch_shunt = 0.0;
            } else {
                branch_shunt = 0.5 * y_shunt + 1.0 / (1.0 / y_series + 2.0 / y_shunt);
            }
            p.yff = from_status ? (1.0 / tap / tap) * branch_shunt : 0.0;
            p.ytt = to_status ? branch_shunt : 0.0;
        }
    } else {
        p.ytt = y_series + 0.5 * y_shunt;
        p.yff = (1.0 / tap / tap) * p.ytt;
        p.yft = (-1.0 / std::conj(tap_ratio)) * y_series;
        p.ytf = (-1.0 / tap_ratio) * y_series;
    }
    return p;
}

struct LineConst { // derived once at construction (line.hpp:26-36)
    double base_i;
    cplx y1_series, y1_shunt, y0_series, y0_shunt;
};
inline LineConst line_constants(LineInput const& in, double system_frequency, double u_rated) {
    LineConst c;
    c.base_i = kBasePower3p / u_rated / kSqrt3;
    double const base_y = c.base_i / (u_rated / kSqrt3);
    cplx const j{0.0, 1.0};
    c.y1_series = 1.0 / (in.r1 + j * in.x1) / base_y;
    c.y1_shunt = 2.0 * kPi * system_frequency * in.c1 / base_y * (in.tan1 + j);
    c.y0_series = 1.0 / (in.r0 + j * in.x0) / base_y;
    c.y0_shunt = 2.0 * kPi * system_frequency * in.c0 / base_y * (in.tan0 + j);
    return c;
}
// out: [4][B][B] complex
template <int B> inline void line_param(LineConst const& c, BranchState const& st, double* out) {
    SymBranchParam const p1 = branch_pi_model(st.from_status, st.to_status, c.y1_series, c.y1_shunt, 1.0);
    cplx const v1[4] = {p1.yff, p1.yft, p1.ytf, p1.ytt};
    if constexpr (B == 1) {
        for (int k = 0; k != 4; ++k) put_scalar_tensor<1>(out + 2 * k, v1[k], 0.0);
    } else {
        SymBranchParam const p0 = branch_pi_model(st.from_status, st.to_status, c.y0_series, c.y0_shunt, 1.0);
        cplx const v0[4] = {p0.yff, p0.yft, p0.ytf, p0.ytt};
        for (int k = 0; k != 4; ++k) put_scalar_tensor<3>(out + 18 * k, (2.0 * v1[k] + v0[k]) / 3.0, (v0[k] - v1[k]) / 3.0);
    }
}

// ---- generic branch (component/generic_branch.hpp:37-93): pi model with a complex ratio k e^{j theta}; symmetric only ----
struct GenericBranchConst {
    double base_i_from, base_i_to, sn, theta;
    cplx y1_series, y1_shunt, ratio;
};
inline GenericBranchConst generic_branch_constants(GenericBranchInput const& in, double u1_rated, double u2_rated) {
    GenericBranchConst c;
    c.sn = in.sn;
    double const k = std::isnan(in.k) ? 1.0 : in.k;
    c.theta = std::isnan(in.theta) ? 0.0 : std::fmod(in.theta, 2 * kPi);
    c.base_i_from = kBasePower3p / u1_rated / kSqrt3;
    c.base_i_to = kBasePower3p / u2_rated / kSqrt3;
    double const base_y = c.base_i_to / (u2_rated / kSqrt3);
    cplx const j{0.0, 1.0};
    c.y1_series = 1.0 / (in.r1 + j * in.x1) / base_y;
    c.y1_shunt = (in.g1 + j * in.b1) / base_y;
    c.ratio = k * std::exp(j * c.theta);
    return c;
}
inline void generic_branch_param(GenericBranchConst const& c, BranchState const& st, double* out) { // [4] complex, B = 1
    SymBranchParam const p = branch_pi_model(st.from_status, st.to_status, c.y1_series, c.y1_shunt, c.ratio);
    cplx const v[4] = {p.yff, p.yft, p.ytf, p.ytt};
    for (int k = 0; k != 4; ++k) put_scalar_tensor<1>(out + 2 * k, v[k], 0.0);
}

// ---- asymmetric line (component/asym_line.hpp:26-127, component/line_utils.hpp, common/matrix_utils.hpp) ---------------------
struct Mat3 {
    cplx m[3][3]{};
};
inline Mat3 mat3_sym(cplx s1, cplx s2, cplx s3, cplx m12, cplx m13, cplx m23) {
    Mat3 r;
    r.m[0][0] = s1, r.m[1][1] = s2, r.m[2][2] = s3;
    r.m[0][1] = r.m[1][0] = m12;
    r.m[0][2] = r.m[2][0] = m13;
    r.m[1][2] = r.m[2][1] = m23;
    return r;
}
template <class F> inline Mat3 mat3_map(Mat3 const& a, F f) {
    Mat3 r;
    for (int i = 0; i != 3; ++i)
        for (int j = 0; j != 3; ++j) r.m[i][j] = f(a.m[i][j]);
    return r;
}
inline Mat3 mat3_add(Mat3 const& a, Mat3 const& b) {
    Mat3 r;
    for (int i = 0; i != 3; ++i)
        for (int j = 0; j != 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
    return r;
}
// inverse by cofactors (what a fixed-size 3 x 3 inverse does): inv(i, j) = cofactor(j, i) / det
inline Mat3 mat3_inv(Mat3 const& a) {
    auto cof = [&a](int i, int j) {
        int const i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return a.m[i1][j1] * a.m[i2][j2] - a.m[i1][j2] * a.m[i2][j1];
    };
    cplx const c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    cplx const det = c0 * a.m[0][0] + c1 * a.m[1][0] + c2 * a.m[2][0];
    cplx const invdet = 1.0 / det;
    Mat3 r;
    for (int i = 0; i != 3; ++i)
        for (int j = 0; j != 3; ++j) r.m[i][j] = cof(j, i) * invdet;
    return r;
}
struct AsymLineConst {
    double base_i, i_n;
    Mat3 y_series, y_shunt;
};
inline AsymLineConst asym_line_constants(AsymLineInput const& in, double system_frequency, double u_rated) {
    AsymLineConst c;
    c.i_n = in.i_n;
    c.base_i = kBasePower3p / u_rated / kSqrt3;
    cplx const j{0.0, 1.0};
    Mat3 z;
    if (std::isnan(in.r_na) && std::isnan(in.x_na)) {
        Mat3 const r = mat3_sym(in.r_aa, in.r_bb, in.r_cc, in.r_ba, in.r_ca, in.r_cb);
        Mat3 const x = mat3_sym(in.x_aa, in.x_bb, in.x_cc, in.x_ba, in.x_ca, in.x_cb);
        for (int a = 0; a != 3; ++a)
            for (int b = 0; b != 3; ++b) z.m[a][b] = r.m[a][b] + j * x.m[a][b];
    } else { // neutral conductor given: Kron reduction of the 4 x 4 matrix
        auto zz = [&j](double r, double x) { return cplx{r} + j * cplx{x}; };
        Mat3 const z_pp = mat3_sym(zz(in.r_aa, in.x_aa), zz(in.r_bb, in.x_bb), zz(in.r_cc, in.x_cc), zz(in.r_ba, in.x_ba),
                                   zz(in.r_ca, in.x_ca), zz(in.r_cb, in.x_cb));
        cplx const z_pn[3] = {zz(in.r_na, in.x_na), zz(in.r_nb, in.x_nb), zz(in.r_nc, in.x_nc)};
        cplx const z_nn_inv = 1.0 / zz(in.r_nn, in.x_nn);
        for (int a = 0; a != 3; ++a)
            for (int b = 0; b != 3; ++b) z.m[a][b] = z_pp.m[a][b] - (z_pn[a] * z_pn[b]) * z_nn_inv;
    }
    Mat3 cm;
    if (!std::isnan(in.c0) && !std::isnan(in.c1)) {
        cplx const sdiag = (2.0 * in.c1 + in.c0) / 3.0, moff = (in.c0 - in.c1) / 3.0;
        cm = mat3_sym(sdiag, sdiag, sdiag, moff, moff, moff);
    } else {
        cm = mat3_sym(in.c_aa, in.c_bb, in.c_cc, in.c_ba, in.c_ca, in.c_cb);
    }
    double const base_y = c.base_i / (u_rated / kSqrt3);
    double const inv_base_y = 1 / base_y;
    c.y_series = mat3_map(mat3_inv(z), [inv_base_y](cplx v) { return inv_base_y * v; });
    cplx const w = 2.0 * j * kPi * system_frequency;
    c.y_shunt = mat3_map(cm, [inv_base_y, w](cplx v) { return inv_base_y * (w * v); });
    return c;
}
template <int B> inline void asym_line_param(AsymLineConst const& c, BranchState const& st, double* out) { // [4][B][B] complex
    if constexpr (B == 1) {
        auto avg_diag = [](Mat3 const& a) { return (a.m[0][0] + a.m[1][1] + a.m[2][2]) / 3.0; };
        // the reference's off-diagonal average as it is written (matrix_utils.hpp:16-18: (1,2) twice, (0,2) absent)
        auto avg_off = [](Mat3 const& a) { return (a.m[0][1] + a.m[1][2] + a.m[1][0] + a.m[1][2] + a.m[2][0] + a.m[2][1]) / 6.0; };
        cplx const y1_series = avg_diag(c.y_series) - avg_off(c.y_series);
        cplx const y1_shunt = avg_diag(c.y_shunt) - avg_off(c.y_shunt);
        SymBranchParam const p = branch_pi_model(st.from_status, st.to_status, y1_series, y1_shunt, 1.0);
        cplx const v[4] = {p.yff, p.yft, p.ytf, p.ytt};
        for (int k = 0; k != 4; ++k) put_scalar_tensor<1>(out + 2 * k, v[k], 0.0);
    } else {
        Mat3 yff, yft, ytf, ytt;
        if (!(st.from_status && st.to_status)) {
            if (st.from_status || st.to_status) {
                Mat3 branch_shunt;
                bool all_above = true;
                for (int a = 0; a != 3; ++a)
                    for (int b = 0; b != 3; ++b) all_above = all_above && std::sqrt(std::norm(c.y_shunt.m[a][b])) >= kNumTol;
                if (all_above) {
                    Mat3 const inner = mat3_add(mat3_inv(c.y_series), mat3_map(mat3_inv(c.y_shunt), [](cplx v) { return 2.0 * v; }));
                    branch_shunt = mat3_add(mat3_map(c.y_shunt, [](cplx v) { return 0.5 * v; }), mat3_inv(inner));
                }
                if (st.from_status) yff = branch_shunt;
                if (st.to_status) ytt = branch_shunt;
            }
        } else {
            ytt = mat3_add(c.y_series, mat3_map(c.y_shunt, [](cplx v) { return 0.5 * v; }));
            yff = ytt;
            yft = mat3_map(c.y_series, [](cplx v) { return -v; });
            ytf = yft;
        }
        Mat3 const* blocks[4] = {&yff, &yft, &ytf, &ytt};
        for (int k = 0; k != 4; ++k)
            for (int a = 0; a != 3; ++a)
                for (int b = 0; b != 3; ++b) {
                    out[18 * k + 2 * (a * 3 + b)] = blocks[k]->m[a][b].real();
                    out[18 * k + 2 * (a * 3 + b) + 1] = blocks[k]->m[a][b].imag();
                }
    }
}

struct TransformerConst { // transformer.hpp:33-83
    double u1, u2, sn, tap_size, uk, pk, i0, p0, i0_zero, p0_zero;
    IntS winding_from, winding_to, clock, tap_side, tap_min, tap_max, tap_nom, tap_direction;
    double uk_min, uk_max, pk_min, pk_max;
    double base_i_from, base_i_to, nominal_ratio;
    cplx z_grounding_from, z_grounding_to;
    IntS initial_tap_pos;
    bool clock_valid;
};
inline IntS tap_limit(TransformerConst const& c, IntS tap) {
    tap = std::min(tap, std::max(c.tap_max, c.tap_min));
    tap = std::max(tap, std::min(c.tap_max, c.tap_min));
    return tap;
}
inline TransformerConst transformer_constants(TransformerInput const& in, double u1_rated, double u2_rated) {
    auto nan_or = [](double v, double fallback) { return std::isnan(v) ? fallback : v; };
    auto z_pu = [](double r, double x, double u) {
        r = std::isnan(r) ? 0 : r;
        x = std::isnan(x) ? 0 : x;
        double const base_z = u * u / kBasePower3p;
        return cplx{r / base_z, x / base_z};
    };
    TransformerConst c{};
    c.u1 = in.u1;
    c.u2 = in.u2;
    c.sn = in.sn;
    c.tap_size = in.tap_size;
    c.uk = in.uk;
    c.pk = in.pk;
    c.i0 = in.i0;
    c.p0 = in.p0;
    c.i0_zero = nan_or(in.i0_zero_sequence, c.i0);
    c.p0_zero = std::isnan(in.p0_zero_sequence) ? c.p0 + c.pk * (c.i0_zero * c.i0_zero - c.i0 * c.i0) : in.p0_zero_sequence;
    c.winding_from = in.winding_from;
    c.winding_to = in.winding_to;
    c.tap_side = in.tap_side;
    c.tap_min = in.tap_min;
    c.tap_max = in.tap_max;
    c.tap_nom = in.tap_nom == kNaIntS ? IntS{0} : in.tap_nom;
    c.tap_direction = c.tap_max > c.tap_min ? IntS{1} : IntS{-1};
    c.uk_min = nan_or(in.uk_min, c.uk);
    c.uk_max = nan_or(in.uk_max, c.uk);
    c.pk_min = nan_or(in.pk_min, c.pk);
    c.pk_max = nan_or(in.pk_max, c.pk);
    c.base_i_from = kBasePower3p / u1_rated / kSqrt3;
    c.base_i_to = kBasePower3p / u2_rated / kSqrt3;
    c.nominal_ratio = u1_rated / u2_rated;
    c.z_grounding_from = z_pu(in.r_grounding_from, in.x_grounding_from, u1_rated);
    c.z_grounding_to = z_pu(in.r_grounding_to, in.x_grounding_to, u2_rated);
    IntS const tap0 = in.tap_pos == kNaIntS ? c.tap_nom : in.tap_pos;
    bool const even = (in.clock % 2) == 0;
    bool const from_wye = in.winding_from == 0 || in.winding_from == 1;
    bool const to_wye = in.winding_to == 0 || in.winding_to == 1;
    c.clock_valid = (even == (from_wye == to_wye));
    c.clock = static_cast<IntS>((in.clock % 12 + 12) % 12);
    c.initial_tap_pos = tap_limit(c, tap0);
    return c;
}
inline double tap_adjust_impedance(double tap_pos, double tap_min, double tap_max, double tap_nom, double xk, double xk_min,
                                   double xk_max) {
    if (tap_pos <= std::max(tap_nom, tap_max) && tap_pos >= std::min(tap_nom, tap_max)) {
        if (tap_max == tap_nom) return xk;
        double const step = (xk_max - xk) / (tap_max - tap_nom);
        return xk + (tap_pos - tap_nom) * step;
    }
    if (tap_min == tap_nom) return xk;
    double const step = (xk_min - xk) / (tap_min - tap_nom);
    return xk + (tap_pos - tap_nom) * step;
}
template <int B>
inline void transformer_param(TransformerConst const& c, BranchState const& st, IntS tap_pos, double* out) {
    cplx const j{0.0, 1.0};
    // transformer_params (transformer.hpp:181-241)
    double const base_y_to = c.base_i_to * c.base_i_to / kBasePower1p;
    double u1 = c.u1, u2 = c.u2;
    if (c.tap_side == 0) {
        u1 += c.tap_direction * (tap_pos - c.tap_nom) * c.tap_size;
    } else {
        u2 += c.tap_direction * (tap_pos - c.tap_nom) * c.tap_size;
    }
    double const k = (u1 / u2) / c.nominal_ratio;
    double const uk = tap_adjust_impedance(tap_pos, c.tap_min, c.tap_max, c.tap_nom, c.uk, c.uk_min, c.uk_max);
    double const pk = tap_adjust_impedance(tap_pos, c.tap_min, c.tap_max, c.tap_nom, c.pk, c.pk_min, c.pk_max);
    cplx z_series{};
    double const uk_sign = (uk >= 0) ? 1.0 : -1.0;
    double const z_abs = std::abs(uk) * u2 * u2 / c.sn;
    z_series.real(pk * u2 * u2 / c.sn / c.sn);
    double const zi2 = z_abs * z_abs - z_series.real() * z_series.real();
    z_series.imag(uk_sign * (zi2 > 0.0 ? std::sqrt(zi2) : 0.0));
    cplx const y_series = (1.0 / z_series) / base_y_to;
    auto magnetising = [&](double i0, double p0) {
        cplx y;
        double const y_abs = i0 * c.sn / u2 / u2;
        y.real(p0 / u2 / u2);
        double const yi2 = y_abs * y_abs - y.real() * y.real();
        y.imag(yi2 > 0.0 ? -std::sqrt(yi2) : 0.0);
        return y / base_y_to;
    };
    cplx const y_shunt = magnetising(c.i0, c.p0);
    cplx const y0_shunt = magnetising(c.i0_zero, c.p0_zero);

    SymBranchParam const p1 = branch_pi_model(st.from_status, st.to_status, y_series, y_shunt, k * std::exp(j * (c.clock * kDeg30)));
    if constexpr (B == 1) {
        cplx const v[4] = {p1.yff, p1.yft, p1.ytf, p1.ytt};
        for (int i = 0; i != 4; ++i) put_scalar_tensor<1>(out + 2 * i, v[i], 0.0);
    } else {
        SymBranchParam const p2 = branch_pi_model(st.from_status, st.to_status, y_series, y_shunt, k * std::exp(j * (-c.clock * kDeg30)));
        SymBranchParam p0;
        constexpr IntS wye_n = 1, delta = 2, zigzag_n = 4;
        if (c.winding_from == wye_n && c.winding_to == wye_n) {
            double shift0 = 0.0;
            if (c.clock == 2 || c.clock == 6 || c.clock == 10) shift0 = 6.0 * kDeg30;
            cplx const z0_series = 1.0 / y_series + 3.0 * (c.z_grounding_to + c.z_grounding_from / k / k);
            cplx const y0_series = 1.0 / z0_series;
            p0 = branch_pi_model(st.from_status, st.to_status, y0_series, y0_shunt, k * std::exp(j * shift0));
        } else if (c.winding_from == wye_n && st.from_status) {
            cplx y0 = y0_shunt;
            if (c.winding_to == delta) y0 += y_series;
            if (y0 != cplx{0.0, 0.0}) {
                cplx const z0 = 1.0 / y0 + 3.0 * c.z_grounding_from / k / k;
                y0 = 1.0 / z0;
                p0.yff = y0 / k / k;
            }
        } else if (c.winding_to == wye_n && st.to_status) {
            cplx y0 = y0_shunt;
            if (c.winding_from == delta) y0 += y_series;
            if (y0 != cplx{0.0, 0.0}) {
                cplx const z0 = 1.0 / y0 + 3.0 * c.z_grounding_to;
                y0 = 1.0 / z0;
                p0.ytt = y0;
            }
        }
        if (c.winding_from == zigzag_n && st.from_status) {
            cplx const z0_series = (1.0 / y_series) * 0.1 + 3.0 * c.z_grounding_from / k / k;
            p0.yff = (1.0 / z0_series) / k / k;
        }
        if (c.winding_to == zigzag_n && st.to_status) {
            cplx const z0_series = (1.0 / y_series) * 0.1 + 3.0 * c.z_grounding_to;
            p0.ytt = 1.0 / z0_series;
        }
        double const low_susceptance = -1e-8 * c.sn / kBasePower3p / c.uk;
        auto zero_seq_available = [](IntS self, IntS other) {
            if (self == wye_n) return other == wye_n || other == delta;
            return self == zigzag_n;
        };
        if (!zero_seq_available(c.winding_from, c.winding_to) && st.from_status) p0.yff += cplx{0.0, low_susceptance};
        if (!zero_seq_available(c.winding_to, c.winding_from) && st.to_status) p0.ytt += cplx{0.0, low_susceptance};
        // y_abc = A diag(y0, y1, y2) A^-1, evaluated as (A * D) * A^-1 with sequential dot products
        cplx const A[3][3] = {{1.0, 1.0, 1.0}, {1.0, kA2, kA}, {1.0, kA, kA2}};
        cplx Ai[3][3] = {{1.0, 1.0, 1.0}, {1.0, kA, kA2}, {1.0, kA2, kA}};
        for (auto& row : Ai)
            for (auto& x : row) x = x / 3.0;
        cplx const s0[4] = {p0.yff, p0.yft, p0.ytf, p0.ytt};
        cplx const s1[4] = {p1.yff, p1.yft, p1.ytf, p1.ytt};
        cplx const s2[4] = {p2.yff, p2.yft, p2.ytf, p2.ytt};
        for (int i = 0; i != 4; ++i) {
            cplx D[3][3] = {};
            D[0][0] = s0[i];
            D[1][1] = s1[i];
            D[2][2] = s2[i];
            cplx AD[3][3];
            for (int r = 0; r != 3; ++r)
                for (int q = 0; q != 3; ++q) {
                    cplx s = A[r][0] * D[0][q];
                    for (int m = 1; m != 3; ++m) s += A[r][m] * D[m][q];
                    AD[r][q] = s;
                }
            for (int r = 0; r != 3; ++r)
                for (int q = 0; q != 3; ++q) {
                    cplx s = AD[r][0] * Ai[0][q];
                    for (int m = 1; m != 3; ++m) s += AD[r][m] * Ai[m][q];
                    out[18 * i + 2 * (r * 3 + q)] = s.real();
                    out[18 * i + 2 * (r * 3 + q) + 1] = s.imag();
                }
        }
    }
}

// ---- source / shunt / load_gen ------------------------------------------------------------------------------------
// ---- link (component/link.hpp:16-39): fixed series admittance y_link (common/common.hpp:100-101), no shunt, ratio 1 ------------
constexpr double kGLink = 1e6 / (kBasePower3p / 10e3 / 10e3);
inline LineConst link_constants(double u_rated_from) {
    return LineConst{kBasePower3p / u_rated_from / kSqrt3, cplx{kGLink, kGLink}, cplx{0.0, 0.0}, cplx{kGLink, kGLink}, cplx{0.0, 0.0}};
}

// ---- three-winding transformer (component/three_winding_transformer.hpp:28-435, component/branch3.hpp) -------------------------
// Three two-winding transformers T1, T2, T3 between the three nodes and an internal node: uk / pk converted delta -> wye relative to
// side 1 (:226-280); T1 is a YNyn0 transformer carrying i0 / p0, T2 / T3 take the reversed clocks (:301-408).
struct ThreeWindingConst {
    ThreeWindingTransformerInput in; // optional values resolved
    double u_rated[3];
    double base_i[3];
    IntS tap_nom, tap_direction, clock_12, clock_13, initial_tap_pos;
    bool clock_valid;
};
struct ThreeWindingState {
    bool status[3];
    IntS tap_pos;
};
inline IntS tap_limit(ThreeWindingConst const& c, IntS tap) {
    tap = std::min(tap, std::max(c.in.tap_max, c.in.tap_min));
    tap = std::max(tap, std::min(c.in.tap_max, c.in.tap_min));
    return tap;
}
inline ThreeWindingConst three_winding_constants(ThreeWindingTransformerInput const& input, double u1_rated, double u2_rated, double u3_rated) {
    ThreeWindingConst c{input, {u1_rated, u2_rated, u3_rated}, {}, 0, 0, 0, 0, 0, true};
    auto nz = [](double& v, double fallback) { v = std::isnan(v) ? fallback : v; };
    auto& in = c.in;
    c.tap_nom = in.tap_nom == kNaIntS ? IntS{0} : in.tap_nom;
    c.tap_direction = in.tap_max > in.tap_min ? IntS{1} : IntS{-1};
    nz(in.uk_12_min, in.uk_12), nz(in.uk_12_max, in.uk_12), nz(in.uk_13_min, in.uk_13), nz(in.uk_13_max, in.uk_13);
    nz(in.uk_23_min, in.uk_23), nz(in.uk_23_max, in.uk_23), nz(in.pk_12_min, in.pk_12), nz(in.pk_12_max, in.pk_12);
    nz(in.pk_13_min, in.pk_13), nz(in.pk_13_max, in.pk_13), nz(in.pk_23_min, in.pk_23), nz(in.pk_23_max, in.pk_23);
    for (int k = 0; k != 3; ++k) c.base_i[k] = kBasePower3p / c.u_rated[k] / kSqrt3;
    auto valid = [](IntS clock, IntS wa, IntS wb) {
        auto wye = [](IntS w) { return w == 0 || w == 1; };
        return ((clock % 2) == 0) == (wye(wa) == wye(wb));
    };
    c.clock_valid = valid(in.clock_12, in.winding_1, in.winding_2) && valid(in.clock_13, in.winding_1, in.winding_3);
    c.clock_12 = static_cast<IntS>((in.clock_12 % 12 + 12) % 12);
    c.clock_13 = static_cast<IntS>((in.clock_13 % 12 + 12) % 12);
    c.initial_tap_pos = tap_limit(c, in.tap_pos == kNaIntS ? c.tap_nom : in.tap_pos);
    return c;
}
// parameters of the three math branches (side k -> internal node): out = [3][4][B][B] complex
template <int B> inline void three_winding_param(ThreeWindingConst const& c, ThreeWindingState const& st, double* out) {
    constexpr int bb2 = B * B * 2;
    std::fill_n(out, 3 * 4 * bb2, 0.0);
    if (!(st.status[0] || st.status[1] || st.status[2])) return; // Branch3::calc_param: not energized
    auto const& in = c.in;
    double u1 = in.u1, u2 = in.u2, u3 = in.u3;
    double const du = c.tap_direction * (st.tap_pos - c.tap_nom) * in.tap_size;
    if (in.tap_side == 0) {
        u1 += du;
    } else if (in.tap_side == 1) {
        u2 += du;
    } else {
        u3 += du;
    }
    auto adj = [&](double x, double x_min, double x_max) { return tap_adjust_impedance(st.tap_pos, in.tap_min, in.tap_max, c.tap_nom, x, x_min, x_max); };
    double const sn_1 = in.sn_1, sn_2 = in.sn_2, sn_3 = in.sn_3;
    double const uk_12 = adj(in.uk_12, in.uk_12_min, in.uk_12_max) * sn_1 / std::min(sn_1, sn_2);
    double const uk_13 = adj(in.uk_13, in.uk_13_min, in.uk_13_max) * sn_1 / std::min(sn_1, sn_3);
    double const uk_23 = adj(in.uk_23, in.uk_23_min, in.uk_23_max) * sn_1 / std::min(sn_2, sn_3);
    double const uk_t[3] = {0.5 * (uk_12 + uk_13 - uk_23), 0.5 * (uk_12 + uk_23 - uk_13) * (sn_2 / sn_1), 0.5 * (uk_13 + uk_23 - uk_12) * (sn_3 / sn_1)};
    double const pk_12 = adj(in.pk_12, in.pk_12_min, in.pk_12_max) * (sn_1 / std::min(sn_1, sn_2)) * (sn_1 / std::min(sn_1, sn_2));
    double const pk_13 = adj(in.pk_13, in.pk_13_min, in.pk_13_max) * (sn_1 / std::min(sn_1, sn_3)) * (sn_1 / std::min(sn_1, sn_3));
    double const pk_23 = adj(in.pk_23, in.pk_23_min, in.pk_23_max) * (sn_1 / std::min(sn_2, sn_3)) * (sn_1 / std::min(sn_2, sn_3));
    double const pk_t[3] = {0.5 * (pk_12 + pk_13 - pk_23), 0.5 * (pk_12 + pk_23 - pk_13) * (sn_2 / sn_1) * (sn_2 / sn_1),
                            0.5 * (pk_13 + pk_23 - pk_12) * (sn_3 / sn_1) * (sn_3 / sn_1)};
    double const u_side[3] = {u1, u2, u3};
    double const sn[3] = {sn_1, sn_2, sn_3};
    IntS const w_from[3] = {1, in.winding_2, in.winding_3};
    IntS const w_to[3] = {1, in.winding_1, in.winding_1};
    IntS const clock[3] = {0, static_cast<IntS>(12 - c.clock_12), static_cast<IntS>(12 - c.clock_13)};
    double const rg[3] = {in.r_grounding_1, in.r_grounding_2, in.r_grounding_3};
    double const xg[3] = {in.x_grounding_1, in.x_grounding_2, in.x_grounding_3};
    for (int k = 0; k != 3; ++k) {
        TransformerInput t{};
        t.id = 2;
        t.from_node = 0;
        t.to_node = 1;
        t.from_status = st.status[k] ? 1 : 0;
        t.to_status = 1;
        t.u1 = u_side[k];
        t.u2 = u1;
        t.sn = sn[k];
        t.uk = uk_t[k];
        t.pk = pk_t[k];
        t.i0 = k == 0 ? in.i0 : 0.0;
        t.p0 = k == 0 ? in.p0 : 0.0;
        t.i0_zero_sequence = kNaN;
        t.p0_zero_sequence = kNaN;
        t.winding_from = w_from[k];
        t.winding_to = w_to[k];
        t.clock = clock[k];
        t.tap_side = 0;
        t.tap_pos = t.tap_min = t.tap_max = t.tap_nom = 0;
        t.tap_size = 0.0;
        t.uk_min = t.uk_max = t.pk_min = t.pk_max = kNaN;
        t.r_grounding_from = rg[k];
        t.x_grounding_from = xg[k];
        t.r_grounding_to = 0.0;
        t.x_grounding_to = 0.0;
        TransformerConst const tc = transformer_constants(t, c.u_rated[k], c.u_rated[0]);
        transformer_param<B>(tc, BranchState{st.status[k], true}, 0, out + k * 4 * bb2);
    }
}

inline void source_param(SourceState const& s, double* out4) { // y1, y0 (source.hpp:40-48)
    double const z_abs = kBasePower3p / s.sk;
    double const x1 = z_abs / std::sqrt(s.rx_ratio * s.rx_ratio + 1.0);
    double const r1 = x1 * s.rx_ratio;
    cplx const y1 = 1.0 / cplx{r1, x1};
    cplx const y0 = y1 / s.z01_ratio;
    out4[0] = y1.real();
    out4[1] = y1.imag();
    out4[2] = y0.real();
    out4[3] = y0.imag();
}
inline cplx source_u_ref(SourceState const& s) { return s.u_ref * std::exp(cplx{0.0, 1.0} * s.u_ref_angle); }

inline bool shunt_set(ShuntState& st, double base_y, double g1, double b1, double g0, double b0) { // shunt.hpp:86-107
    auto upd = [](double v, double& target) {
        if (std::isnan(v) || v == target) return false;
        target = v;
        return true;
    };
    bool changed = upd(g1, st.g1);
    changed = upd(b1, st.b1) || changed;
    changed = upd(g0, st.g0) || changed;
    changed = upd(b0, st.b0) || changed;
    if (changed) {
        st.y1 = (st.g1 + cplx{0.0, 1.0} * st.b1) / base_y;
        st.y0 = (st.g0 + cplx{0.0, 1.0} * st.b0) / base_y;
    }
    return changed;
}
template <int B> inline void shunt_param(ShuntState const& st, double* out) {
    if (!st.status) {
        put_scalar_tensor<B>(out, 0.0, 0.0);
    } else if constexpr (B == 1) {
        put_scalar_tensor<1>(out, st.y1, 0.0);
    } else {
        put_scalar_tensor<3>(out, (2.0 * st.y1 + st.y0) / 3.0, (st.y0 - st.y1) / 3.0);
    }
}

// per-unit injection of a load/generator for a B-phase calculation; lb = phases of the component (1 or 3)
template <int B> inline void load_gen_injection(LoadGenState const& st, int lb, cplx* out) {
    if (!st.status) {
        for (int p = 0; p != B; ++p) out[p] = 0.0;
        return;
    }
    if constexpr (B == 1) {
        if (lb == 1) {
            bool const bad = std::isnan(st.s[0].real()) || std::isnan(st.s[0].imag());
            out[0] = bad ? cplx{kNaN, kNaN} : st.s[0];
        } else {
            out[0] = (st.s[0] + st.s[1] + st.s[2]) / 3.0;
        }
    } else {
        if (lb == 1) {
            bool const bad = std::isnan(st.s[0].real()) || std::isnan(st.s[0].imag());
            for (int p = 0; p != 3; ++p) out[p] = bad ? cplx{kNaN, kNaN} : st.s[0];
        } else {
            for (int p = 0; p != 3; ++p) out[p] = st.s[p];
        }
    }
}

} // namespace pgmb
