// Device helpers of the generic (2B x 2B real block) Newton-Raphson row tasks, shared by nr_block.cu (asymmetric, B = 3) and
// nr_sym_v2.cu (wide rows of meshed symmetric grids, B = 1).  Blocks are column-major like the reference's Eigen blocks;
// sub-blocks H (0,0) N (0,1) M (1,0) L (1,1).
//   newton_raphson_pf_solver.hpp:255-303 (linear start), 462-547, 764-852 (Jacobian + mismatch), 325-349 (update)
//   sparse_lu_solver.hpp:86-165 (full-pivot block LU, first maximum in column-major order), 171-200 (triangular solves),
//                        346-495 (prefactorize), 769-827 (solve_once)
#pragma once

#include "kernels.cuh"

#include <cfloat>
#include <cmath>
#include <cuda_runtime.h>

namespace pgmb {
namespace blk {

constexpr int kStatusOk = 0, kStatusDiverged = 1, kStatusSingular = 2;
enum class Mode { linear_init, newton };

__device__ __forceinline__ bool not_normal(double x) { return !(fabs(x) >= DBL_MIN) || isinf(x); }

// SymPerm: the symmetric kernels keep the 2 x 2 pivot permutation of a bus in one byte (bit 0 row swap, bit 1 column swap)
// at perm[bus]; the generic layout keeps p[N] then q[N] at perm[bus * 2N + i].
template <int T, int B, bool SymPerm = false> struct TileB {
    static constexpr int N = 2 * B, NN = N * N;
    double* jac;
    double* xvec;
    double* pol;
    double* u;
    uint8_t* perm;
    double const* sinj;
    double const* usrc;
    // scratch of the cooperative wide-row elimination (may be null when the structure has no wide rows)
    double* wide_terms; // [max_upd][NN][T]
    double* wide_rhs;   // [max_lower][N][T]
    double* wide_sum;   // [max_entries][N][T]
    // PV buses (only read by the REG instantiations)
    uint8_t const* lg_status; // [n_load_gen][T]
    uint8_t* qviol;           // [n_bus][T]
    // branch-outage overlay of this lane's scenario (DevOverlay), null = none
    int32_t const* ovr_entry{nullptr}; // [ovr_n]
    double const* ovr_y{nullptr};      // [ovr_n][B*B][2]
    int ovr_n{0};                      // 4 per switched-branch slot of the batch
    uint8_t const* dead{nullptr};      // [n_bus] buses without supply in this scenario: identity rows, voltage 0
    __device__ __forceinline__ void get_q(int bus, uint8_t* q) const {
        if constexpr (SymPerm) {
            static_assert(!SymPerm || B == 1, "byte-packed permutation is a 2 x 2 layout");
            uint8_t const sw = (perm[(size_t)bus * T] >> 1) & 1;
            q[0] = sw;
            q[1] = 1 - sw;
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) q[i] = perm[(size_t)(bus * 2 * N + N + i) * T];
        }
    }
    __device__ __forceinline__ void get_p(int bus, uint8_t* p) const {
        if constexpr (SymPerm) {
            uint8_t const sw = perm[(size_t)bus * T] & 1;
            p[0] = sw;
            p[1] = 1 - sw;
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) p[i] = perm[(size_t)(bus * 2 * N + i) * T];
        }
    }
    __device__ __forceinline__ void set_perm(int bus, uint8_t const* p, uint8_t const* q) const {
        if constexpr (SymPerm) {
            perm[(size_t)bus * T] = static_cast<uint8_t>((p[0] != 0 ? 1 : 0) | ((q[0] != 0 ? 1 : 0) << 1));
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                perm[(size_t)(bus * 2 * N + i) * T] = p[i];
                perm[(size_t)(bus * 2 * N + N + i) * T] = q[i];
            }
        }
    }
    __device__ __forceinline__ void load_blk(int k, double* a) const {
        double const* p = jac + (size_t)k * NN * T;
#pragma unroll
        for (int i = 0; i < NN; ++i) a[i] = p[(size_t)i * T];
    }
    __device__ __forceinline__ void store_blk(int k, double const* a) const {
        double* p = jac + (size_t)k * NN * T;
#pragma unroll
        for (int i = 0; i < NN; ++i) p[(size_t)i * T] = a[i];
    }
    // voltages: component 2p = re, 2p + 1 = im of phase p
    __device__ __forceinline__ void load_u(int bus, double* ur, double* ui) const {
#pragma unroll
        for (int p = 0; p < B; ++p) {
            ur[p] = u[(size_t)(bus * N + 2 * p) * T];
            ui[p] = u[(size_t)(bus * N + 2 * p + 1) * T];
        }
    }
};

// element (r, c) of sub-block (br, bc), column-major N x N
template <int B> __device__ __forceinline__ int el(int br, int bc, int r, int c) { return (bc * B + c) * (2 * B) + (br * B + r); }

// block = hnml(y, ui, uj): power_flow[r][c] = (ui[r] * conj(uj[c])) * conj(y[r][c]); H = L = imag, N = -M = real
template <int B>
__device__ __forceinline__ void hnml(double* blk, double const* y /*[B*B][2] row-major*/, double sign, double const* uir,
                                     double const* uii, double const* ujr, double const* uji) {
#pragma unroll
    for (int r = 0; r < B; ++r)
#pragma unroll
        for (int c = 0; c < B; ++c) {
            double const yr = sign * y[2 * (r * B + c)], yi = sign * y[2 * (r * B + c) + 1];
            double const cr = ujr[c], ci = -uji[c];
            double const ar = uir[r] * cr - uii[r] * ci;
            double const ai = uir[r] * ci + uii[r] * cr;
            double const dr = yr, di = -yi;
            double const n = ar * dr - ai * di;
            double const h = ar * di + ai * dr;
            blk[el<B>(0, 0, r, c)] = h;
            blk[el<B>(0, 1, r, c)] = n;
            blk[el<B>(1, 0, r, c)] = -n;
            blk[el<B>(1, 1, r, c)] = h;
        }
}
template <int B> __device__ __forceinline__ void linear_block(double* blk, double const* y) {
#pragma unroll
    for (int r = 0; r < B; ++r)
#pragma unroll
        for (int c = 0; c < B; ++c) {
            double const g = y[2 * (r * B + c)], b = y[2 * (r * B + c) + 1];
            blk[el<B>(0, 1, r, c)] = -b;
            blk[el<B>(0, 0, r, c)] = g;
            blk[el<B>(1, 1, r, c)] = g;
            blk[el<B>(1, 0, r, c)] = b;
        }
}

// DenseLUFactor::factorize_block_in_place for an N x N real block; p / q are the accumulated permutation index vectors
template <int N> __device__ bool factorize_block(double* m, uint8_t* p, uint8_t* q) {
    int rt[N], ct[N];
    double max_pivot = 0.0;
    bool stopped = false;
    for (int pivot = 0; pivot < N; ++pivot) {
        int rb = pivot, cb = pivot;
        double best = m[pivot * N + pivot] * m[pivot * N + pivot];
        for (int c = pivot; c < N; ++c)
            for (int r = pivot; r < N; ++r) {
                double const sc = m[c * N + r] * m[c * N + r];
                if (sc > best) {
                    best = sc;
                    rb = r;
                    cb = c;
                }
            }
        if (best == 0.0) {
            for (int k = pivot; k < N; ++k) {
                rt[k] = k;
                ct[k] = k;
            }
            stopped = true;
            break;
        }
        max_pivot = fmax(max_pivot, sqrt(best));
        rt[pivot] = rb;
        ct[pivot] = cb;
        if (rb != pivot)
            for (int c = 0; c < N; ++c) {
                double const x = m[c * N + pivot];
                m[c * N + pivot] = m[c * N + rb];
                m[c * N + rb] = x;
            }
        if (cb != pivot)
            for (int r = 0; r < N; ++r) {
                double const x = m[pivot * N + r];
                m[pivot * N + r] = m[cb * N + r];
                m[cb * N + r] = x;
            }
        if (pivot < N - 1) {
            for (int r = pivot + 1; r < N; ++r) m[pivot * N + r] /= m[pivot * N + pivot];
            for (int c = pivot + 1; c < N; ++c)
                for (int r = pivot + 1; r < N; ++r) m[c * N + r] -= m[pivot * N + r] * m[c * N + pivot];
        }
    }
    (void)stopped;
    for (int i = 0; i < N; ++i) {
        p[i] = (uint8_t)i;
        q[i] = (uint8_t)i;
    }
    for (int pivot = N - 1; pivot >= 0; --pivot) {
        uint8_t const x = p[pivot];
        p[pivot] = p[rt[pivot]];
        p[rt[pivot]] = x;
    }
    for (int pivot = 0; pivot < N; ++pivot) {
        uint8_t const x = q[pivot];
        q[pivot] = q[ct[pivot]];
        q[ct[pivot]] = x;
    }
    double const threshold = DBL_EPSILON * max_pivot;
    bool singular = false;
    for (int pivot = 0; pivot < N; ++pivot) {
        double const d = m[pivot * N + pivot];
        singular = singular || fabs(d) < threshold || not_normal(d);
    }
    return singular;
}

// ---- PV buses / reactive-power limits of voltage regulators -------------------------------------------------------------
//   newton_raphson_pf_solver.hpp:400-444 (bus types and limits), 446-452 (reference voltage after the linear start),
//   549-587 (PV rows of the Jacobian), 605-704 (limit check, PV -> PQ switch with clamped Q), 706-742 (linear start ignores
//   the Q of a regulating generator).  Every decision is local to the bus; the only per-scenario state is the iteration
//   number (the check runs from iteration 2 on) and DevBatch.qviol.
constexpr double kNumTol = 1e-8;
struct BusControl {
    bool regulated;  // a regulating (status on, generator on) regulator sits on this non-slack bus
    bool has_limits;
    double u_ref, q_min, q_max;
};
template <int T, int B, bool SP>
__device__ __forceinline__ bool lg_regulating(DevStructure const& s, TileB<T, B, SP> const& t, int lg, int& reg) {
    reg = __ldg(s.lg_reg + lg);
    return reg >= 0 && __ldg(s.reg_param + 4 * reg) != 0.0 && t.lg_status[(size_t)lg * T] != 0;
}
template <int T, int B, bool SP>
__device__ BusControl bus_control(DevStructure const& s, TileB<T, B, SP> const& t, int row) {
    BusControl c{false, false, 0.0, 0.0, 0.0};
    if (__ldg(s.src_ptr + row) != __ldg(s.src_ptr + row + 1)) return c; // slack bus
    for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) {
        int reg;
        if (lg_regulating<T, B, SP>(s, t, lg, reg)) {
            c.regulated = true;
            c.u_ref = __ldg(s.reg_param + 4 * reg + 1);
            c.q_min += __ldg(s.reg_param + 4 * reg + 2);
            c.q_max += __ldg(s.reg_param + 4 * reg + 3);
        }
    }
    c.has_limits = c.regulated && (!isnan(c.q_min) || !isnan(c.q_max));
    return c;
}
// Q of a regulating generator whose bus ran into limit `viol` (1 lower, 2 upper): the regulator's limit, spread over the
// phases in proportion to the specified Q (equally when that is ~0)
template <int T, int B, bool SP>
__device__ __forceinline__ void clamped_q(DevStructure const& s, TileB<T, B, SP> const& t, int lg, int reg, int viol, double* out) {
    constexpr int N = 2 * B;
    double const lim = __ldg(s.reg_param + 4 * reg + (viol == 2 ? 3 : 2));
    if constexpr (B == 1) {
        out[0] = lim;
    } else {
        double base[B];
#pragma unroll
        for (int p = 0; p < B; ++p) base[p] = t.sinj[(size_t)(lg * N + 2 * p + 1) * T];
        double const total = base[0] + base[1] + base[2];
        if (fabs(total) > kNumTol) {
            double const scale = lim / total;
#pragma unroll
            for (int p = 0; p < B; ++p) out[p] = base[p] * scale;
        } else {
#pragma unroll
            for (int p = 0; p < B; ++p) out[p] = lim / 3.0;
        }
    }
}
// enforce_q_limits for one PV bus; acc = mismatch of the freshly built row.  Returns the violated limit (0 = none).
template <int T, int B, bool SP>
__device__ int check_q_limit(DevStructure const& s, TileB<T, B, SP> const& t, int row, BusControl const& c, double const* acc) {
    constexpr int N = 2 * B;
    double spec[B];
#pragma unroll
    for (int p = 0; p < B; ++p) spec[p] = 0.0;
    for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) {
        int reg;
        if (!lg_regulating<T, B, SP>(s, t, lg, reg)) continue;
#pragma unroll
        for (int p = 0; p < B; ++p) spec[p] += t.sinj[(size_t)(lg * N + 2 * p + 1) * T];
    }
    double q_total = spec[0] - acc[B];
#pragma unroll
    for (int p = 1; p < B; ++p) q_total += spec[p] - acc[B + p];
    if (!isnan(c.q_max) && q_total > c.q_max + kNumTol) return 2;
    if (!isnan(c.q_min) && q_total < c.q_min - kNumTol) return 1;
    return 0;
}
// apply_pv_constraints on the diagonal block and the mismatch of a PV row (the other blocks: zero_pv_rows)
template <int T, int B, bool SP>
__device__ __forceinline__ void pv_diag(TileB<T, B, SP> const& t, int row, double* d, double* acc) {
    constexpr int N = 2 * B;
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int r = B; r < N; ++r) d[c * N + r] = 0.0;
#pragma unroll
    for (int p = 0; p < B; ++p) {
        d[el<B>(1, 1, p, p)] = t.pol[(size_t)(row * N + B + p) * T];
        acc[B + p] = 0.0;
    }
}
template <int T, int B, bool SP> __device__ __forceinline__ void zero_pv_rows(TileB<T, B, SP> const& t, int k) {
    constexpr int N = 2 * B, NN = N * N;
    double* p = t.jac + (size_t)k * NN * T;
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int r = B; r < N; ++r) p[(size_t)(c * N + r) * T] = 0.0;
}

// block of LU entry k of row `row` (zero for a fill-in) and, for the Newton step, its contribution to the row sums
// (contrib[r] = sum_c N[r][c], contrib[B + r] = sum_c H[r][c])
template <int T, int B, Mode mode, bool SP>
__device__ __forceinline__ void build_entry(DevStructure const& s, TileB<T, B, SP> const& t, double const* uir, double const* uii,
                                            int k, double* blk, double* contrib, int row = -1) {
    constexpr int N = 2 * B, NN = N * N, BB2 = B * B * 2;
    int ky = __ldg(s.map_y + k);
    if (t.dead != nullptr && ky >= 0 && (t.dead[row] != 0 || t.dead[__ldg(s.col_idx + k)] != 0)) ky = -1; // no coupling
#pragma unroll
    for (int i = 0; i < NN; ++i) blk[i] = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) contrib[i] = 0.0;
    if (ky >= 0) {
        double y[BB2];
        double const* ysrc = s.ydata + (size_t)ky * BB2;
        if (t.ovr_entry != nullptr) { // this scenario replaces the entries of its switched branch
            for (int j = 0; j < t.ovr_n; ++j)
                if (t.ovr_entry[j] == ky) ysrc = t.ovr_y + j * BB2;
        }
#pragma unroll
        for (int i = 0; i < BB2; ++i) y[i] = __ldg(ysrc + i);
        if constexpr (mode == Mode::newton) {
            int const j = __ldg(s.col_idx + k);
            double ujr[B], uji[B];
            t.load_u(j, ujr, uji);
            hnml<B>(blk, y, 1.0, uir, uii, ujr, uji);
#pragma unroll
            for (int r = 0; r < B; ++r) {
                double sn = blk[el<B>(0, 1, r, 0)], sh = blk[el<B>(0, 0, r, 0)];
#pragma unroll
                for (int c = 1; c < B; ++c) {
                    sn += blk[el<B>(0, 1, r, c)];
                    sh += blk[el<B>(0, 0, r, c)];
                }
                contrib[r] = sn;
                contrib[B + r] = sh;
            }
        } else {
            linear_block<B>(blk, y);
        }
    }
}

// diagonal corrections from the row sums, loads and sources of the bus (acc = row sums on entry, mismatch / rhs on exit)
// REG: the grid has voltage regulators; viol = limit the bus ran into (its regulating generators then inject the clamped Q)
template <int T, int B, Mode mode, bool SP, bool REG = false>
__device__ __forceinline__ void finish_diag(DevStructure const& s, TileB<T, B, SP> const& t, int row, double const* uir,
                                            double const* uii, double* acc, double* d, int viol = 0) {
    constexpr int N = 2 * B, NN = N * N, BB2 = B * B * 2;
    if (t.dead != nullptr && t.dead[row] != 0) { // bus without supply: identity row, zero right-hand side -> u stays 0
#pragma unroll
        for (int i = 0; i < NN; ++i) d[i] = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            d[i * N + i] = 1.0;
            acc[i] = 0.0;
        }
        return;
    }
    if constexpr (mode == Mode::newton) {
#pragma unroll
        for (int p = 0; p < B; ++p) {
            d[el<B>(0, 0, p, p)] += acc[B + p];
            d[el<B>(0, 1, p, p)] += -acc[p];
            d[el<B>(1, 0, p, p)] += -acc[p];
            d[el<B>(1, 1, p, p)] += -acc[B + p];
        }
    }
    // loads
    for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) {
        int const type = __ldg(s.lg_type + lg);
        [[maybe_unused]] double qclamp[B];
        [[maybe_unused]] bool regulating = false;
        if constexpr (REG) {
            int reg;
            regulating = lg_regulating<T, B, SP>(s, t, lg, reg);
            if (mode == Mode::newton && regulating && viol != 0) clamped_q<T, B, SP>(s, t, lg, reg, viol, qclamp);
        }
#pragma unroll
        for (int p = 0; p < B; ++p) {
            double const ps = t.sinj[(size_t)(lg * N + 2 * p) * T];
            double qs = t.sinj[(size_t)(lg * N + 2 * p + 1) * T];
            if constexpr (REG) {
                // linear start: the specified Q of a regulating generator is ignored; Newton: clamped once the bus hit a limit
                if (regulating && mode == Mode::linear_init) qs = 0.0;
                if (regulating && mode == Mode::newton && viol != 0) qs = qclamp[p];
            }
            if constexpr (mode == Mode::newton) {
                double const v = t.pol[(size_t)(row * N + B + p) * T];
                if (type == 0) {
                    acc[p] += ps;
                    acc[B + p] += qs;
                } else if (type == 1) {
                    acc[p] += ps * v * v;
                    acc[B + p] += qs * v * v;
                    d[el<B>(0, 1, p, p)] += -ps * 2.0 * v * v;
                    d[el<B>(1, 1, p, p)] += -qs * 2.0 * v * v;
                } else {
                    acc[p] += ps * v;
                    acc[B + p] += qs * v;
                    d[el<B>(0, 1, p, p)] += -ps * v;
                    d[el<B>(1, 1, p, p)] += -qs * v;
                }
            } else {
                double const ylr = -ps, yli = qs; // y_load = -conj(s)
                d[el<B>(0, 1, p, p)] += -yli;
                d[el<B>(0, 0, p, p)] += ylr;
                d[el<B>(1, 1, p, p)] += ylr;
                d[el<B>(1, 0, p, p)] += yli;
            }
        }
    }
    // sources
    for (int sr = __ldg(s.src_ptr + row), sre = __ldg(s.src_ptr + row + 1); sr < sre; ++sr) {
        double y[BB2];
#pragma unroll
        for (int i = 0; i < BB2; ++i) y[i] = __ldg(s.src_yref + (size_t)sr * BB2 + i);
        double const u0r = t.usrc[(size_t)(sr * 2) * T], u0i = t.usrc[(size_t)(sr * 2 + 1) * T];
        double usr[B], usi[B];
        usr[0] = u0r;
        usi[0] = u0i;
        if constexpr (B == 3) { // ComplexValue<asym>{u} = (u, u a^2, u a)   (three_phase_tensor.hpp:47-53)
            double const a2r = -0.5, a2i = -0.8660254037844386 /* -sqrt3/2 */, ar = -0.5, ai = 0.8660254037844386;
            usr[1] = u0r * a2r - u0i * a2i;
            usi[1] = u0r * a2i + u0i * a2r;
            usr[2] = u0r * ar - u0i * ai;
            usi[2] = u0r * ai + u0i * ar;
        }
        if constexpr (mode == Mode::newton) {
            double mm[NN], ms[NN];
            hnml<B>(mm, y, 1.0, uir, uii, uir, uii);
            hnml<B>(ms, y, -1.0, uir, uii, usr, usi);
            double p_cal[B], q_cal[B];
#pragma unroll
            for (int r = 0; r < B; ++r) {
                double sn = mm[el<B>(0, 1, r, 0)] + ms[el<B>(0, 1, r, 0)];
                double sh = mm[el<B>(0, 0, r, 0)] + ms[el<B>(0, 0, r, 0)];
#pragma unroll
                for (int c = 1; c < B; ++c) {
                    sn += mm[el<B>(0, 1, r, c)] + ms[el<B>(0, 1, r, c)];
                    sh += mm[el<B>(0, 0, r, c)] + ms[el<B>(0, 0, r, c)];
                }
                p_cal[r] = sn;
                q_cal[r] = sh;
            }
#pragma unroll
            for (int p = 0; p < B; ++p) {
                mm[el<B>(0, 0, p, p)] += -q_cal[p];
                mm[el<B>(0, 1, p, p)] += p_cal[p];
                mm[el<B>(1, 0, p, p)] += p_cal[p];
                mm[el<B>(1, 1, p, p)] += q_cal[p];
                acc[p] -= p_cal[p];
                acc[B + p] -= q_cal[p];
            }
#pragma unroll
            for (int i = 0; i < NN; ++i) d[i] += mm[i];
        } else {
#pragma unroll
            for (int r = 0; r < B; ++r) {
#pragma unroll
                for (int c = 0; c < B; ++c) {
                    double const yr = y[2 * (r * B + c)], yi = y[2 * (r * B + c) + 1];
                    d[el<B>(0, 1, r, c)] -= yi;
                    d[el<B>(0, 0, r, c)] += yr;
                    d[el<B>(1, 1, r, c)] += yr;
                    d[el<B>(1, 0, r, c)] += yi;
                }
                // rhs += dot(y_source, u_source): sequential sum over the phases
                double sr_ = y[2 * (r * B)] * usr[0] - y[2 * (r * B) + 1] * usi[0];
                double si_ = y[2 * (r * B)] * usi[0] + y[2 * (r * B) + 1] * usr[0];
#pragma unroll
                for (int c = 1; c < B; ++c) {
                    sr_ += y[2 * (r * B + c)] * usr[c] - y[2 * (r * B + c) + 1] * usi[c];
                    si_ += y[2 * (r * B + c)] * usi[c] + y[2 * (r * B + c) + 1] * usr[c];
                }
                acc[r] += sr_;
                acc[B + r] += si_;
            }
        }
    }
}

template <int T, int B, Mode mode, bool SP = false, bool REG = false>
__device__ bool up_row(DevStructure const& s, TileB<T, B, SP> const& t, int row, bool check_now = false) {
    constexpr int N = 2 * B, NN = N * N;
    int const rb = __ldg(s.row_ptr + row), re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    double uir[B], uii[B];
    t.load_u(row, uir, uii);
    double acc[N]; // NR: -P[B], -Q[B] then mismatch ; linear: rhs real[B], imag[B]
    double d[NN];
    [[maybe_unused]] BusControl ctl{};
    [[maybe_unused]] int viol = 0;
    if constexpr (REG && mode == Mode::newton) {
        ctl = bus_control<T, B, SP>(s, t, row);
        viol = t.qviol[(size_t)row * T];
    }
    constexpr int n_pass = (REG && mode == Mode::newton) ? 2 : 1;
#pragma unroll 1
    for (int pass = 0; pass < n_pass; ++pass) { // a second pass only for a PV bus that has just run into a Q limit
#pragma unroll
        for (int i = 0; i < N; ++i) acc[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NN; ++i) d[i] = 0.0;
        // 1. build the row
        for (int k = rb; k < re; ++k) {
            double blk[NN], contrib[N];
            build_entry<T, B, mode, SP>(s, t, uir, uii, k, blk, contrib, row);
            if (__ldg(s.map_y + k) >= 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] -= contrib[i];
            }
            if (k == dg) {
#pragma unroll
                for (int i = 0; i < NN; ++i) d[i] = blk[i];
            } else {
                t.store_blk(k, blk);
            }
        }
        finish_diag<T, B, mode, SP, REG>(s, t, row, uir, uii, acc, d, viol);
        if constexpr (REG && mode == Mode::newton) {
            if (pass == 0 && check_now && ctl.has_limits && viol == 0) {
                viol = check_q_limit<T, B, SP>(s, t, row, ctl, acc);
                if (viol != 0) {
                    t.qviol[(size_t)row * T] = (uint8_t)viol;
                    continue; // the bus is PQ from now on: rebuild its row with the clamped generators
                }
            }
        }
        break;
    }
    if constexpr (REG && mode == Mode::newton) {
        if (ctl.regulated && viol == 0) { // PV row
            pv_diag<T, B, SP>(t, row, d, acc);
            for (int k = rb; k < re; ++k)
                if (k != dg) zero_pv_rows<T, B, SP>(t, k);
        }
    }

    // 2. eliminate against finished rows (L block in local memory only)
    for (int e = rb; e < dg; ++e) {
        int const c = __ldg(s.col_idx + e);
        int const dc = __ldg(s.diag + c);
        double a[NN], piv[NN], l[NN];
        t.load_blk(e, a);
        t.load_blk(dc, piv);
        uint8_t qc[N];
        t.get_q(c, qc);
        // l = (a Q_c): column i of l = column q[i] of a ; then right / upper triangular solve
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int r = 0; r < N; ++r) l[i * N + r] = a[qc[i] * N + r];
        for (int idx = 0; idx < N; ++idx) {
            for (int prev = 0; prev < idx; ++prev) {
                double const uv = piv[idx * N + prev];
#pragma unroll
                for (int r = 0; r < N; ++r) l[idx * N + r] -= uv * l[prev * N + r];
            }
            double const dd = piv[idx * N + idx];
#pragma unroll
            for (int r = 0; r < N; ++r) l[idx * N + r] /= dd;
        }
        for (int q = __ldg(s.upd_ptr + e), qe = __ldg(s.upd_ptr + e + 1); q < qe; ++q) {
            int const ui = __ldg(s.upd_u + q), ai = __ldg(s.upd_a + q);
            double ub[NN];
            t.load_blk(ui, ub);
            double* tgt = d;
            double tb[NN];
            if (ai != dg) {
                t.load_blk(ai, tb);
                tgt = tb;
            }
            for (int cc = 0; cc < N; ++cc)
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double sum = l[0 * N + r] * ub[cc * N + 0];
#pragma unroll
                    for (int k = 1; k < N; ++k) sum += l[k * N + r] * ub[cc * N + k];
                    tgt[cc * N + r] -= sum;
                }
            if (ai != dg) t.store_blk(ai, tb);
        }
        double yc[N];
#pragma unroll
        for (int i = 0; i < N; ++i) yc[i] = t.xvec[(size_t)(c * N + i) * T];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            double sum = l[0 * N + r] * yc[0];
#pragma unroll
            for (int k = 1; k < N; ++k) sum += l[k * N + r] * yc[k];
            acc[r] -= sum;
        }
    }

    // 3. factorise the diagonal block
    uint8_t p[N], q[N];
    bool const singular = factorize_block<N>(d, p, q);
    t.store_blk(dg, d);
    t.set_perm(row, p, q);
    // 4. U blocks: L_pp^-1 (P A)   (P A: row p[i] of the result = row i of A)
    for (int e = dg + 1; e < re; ++e) {
        double a[NN], ub[NN];
        t.load_blk(e, a);
#pragma unroll
        for (int cc = 0; cc < N; ++cc)
#pragma unroll
            for (int i = 0; i < N; ++i) ub[cc * N + p[i]] = a[cc * N + i];
        for (int idx = 0; idx < N; ++idx)
            for (int prev = 0; prev < idx; ++prev) {
                double const lv = d[prev * N + idx];
#pragma unroll
                for (int cc = 0; cc < N; ++cc) ub[cc * N + idx] -= lv * ub[cc * N + prev];
            }
        t.store_blk(e, ub);
    }
    // 5. forward substitution inside the block: x = L_pp^-1 (P t)
    double xr[N];
#pragma unroll
    for (int i = 0; i < N; ++i) xr[p[i]] = acc[i];
    for (int idx = 0; idx < N; ++idx)
        for (int prev = 0; prev < idx; ++prev) xr[idx] -= d[prev * N + idx] * xr[prev];
#pragma unroll
    for (int i = 0; i < N; ++i) t.xvec[(size_t)(row * N + i) * T] = xr[i];
    return singular;
}

template <int T, int B, Mode mode, bool SP = false, bool REG = false>
__device__ double down_row(DevStructure const& s, TileB<T, B, SP> const& t, int row) {
    constexpr int N = 2 * B, NN = N * N;
    int const re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    double y[N];
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = t.xvec[(size_t)(row * N + i) * T];
    for (int e = re - 1; e > dg; --e) {
        int const j = __ldg(s.col_idx + e);
        double ub[NN], xj[N];
        uint8_t qj[N];
        t.load_blk(e, ub);
#pragma unroll
        for (int i = 0; i < N; ++i) xj[i] = t.xvec[(size_t)(j * N + i) * T]; // final solution of row j (column permutation applied)
        t.get_q(j, qj);
        // reference: x_row -= (U Q_j) x'_j with x_j[q[i]] = x'_j[i]  =>  sum over i of U[:, q[i]] * x_j[q[i]], in that order
#pragma unroll
        for (int r = 0; r < N; ++r) {
            double sum = ub[qj[0] * N + r] * xj[qj[0]];
#pragma unroll
            for (int i = 1; i < N; ++i) sum += ub[qj[i] * N + r] * xj[qj[i]];
            y[r] -= sum;
        }
    }
    double d[NN];
    t.load_blk(dg, d);
    for (int step = 0; step < N; ++step) { // left / upper solve, backward traversal
        int const idx = N - 1 - step;
        for (int ps = 0; ps < step; ++ps) {
            int const prev = N - 1 - ps;
            y[idx] -= d[prev * N + idx] * y[prev];
        }
        y[idx] /= d[idx * N + idx];
    }
    double x[N];
    uint8_t qr[N];
    t.get_q(row, qr);
#pragma unroll
    for (int i = 0; i < N; ++i) x[qr[i]] = y[i];
#pragma unroll
    for (int i = 0; i < N; ++i) t.xvec[(size_t)(row * N + i) * T] = x[i];
    double dev = 0.0;
    [[maybe_unused]] BusControl ctl{};
    if constexpr (REG && mode == Mode::linear_init) ctl = bus_control<T, B, SP>(s, t, row);
#pragma unroll
    for (int p = 0; p < B; ++p) {
        double* const pth = t.pol + (size_t)(row * N + p) * T;
        double* const pv = t.pol + (size_t)(row * N + B + p) * T;
        double* const pur = t.u + (size_t)(row * N + 2 * p) * T;
        double* const pui = t.u + (size_t)(row * N + 2 * p + 1) * T;
        if constexpr (mode == Mode::newton) {
            double theta = *pth, v = *pv;
            theta += x[p];
            v += v * x[B + p];
            double sn, cs;
            sincos(theta, &sn, &cs);
            double const nr = v * cs, ni = v * sn;
            double const dr = nr - *pur, di = ni - *pui;
            *pth = theta;
            *pv = v;
            *pur = nr;
            *pui = ni;
            double const dp = sqrt(dr * dr + di * di);
            dev = p == 0 ? dp : fmax(dev, dp);
        } else {
            double xr = x[p], xi = x[B + p]; // linear start: real part in the P rows, imaginary part in the Q rows
            if constexpr (REG) { // a PV bus starts at its reference magnitude: u = u_ref * u / |u|
                if (ctl.regulated) {
                    double const ax = sqrt(xr * xr + xi * xi);
                    double sr = 1.0, si = 0.0;
                    if (ax > 0.0) {
                        sr = xr / ax;
                        si = xi / ax;
                    }
                    xr = ctl.u_ref * sr - 0.0 * si;
                    xi = ctl.u_ref * si + 0.0 * sr;
                }
            }
            *pur = xr;
            *pui = xi;
            *pv = sqrt(xr * xr + xi * xi);
            *pth = atan2(xi, xr);
        }
    }
    return dev;
}

// ---- cooperative elimination of a wide row (symbolic.hpp: WideRowPlan) ------------------------------------------------------
// Called by ALL threads of the block (barriers inside); `active` masks the lanes of finished scenarios.  Phases:
//   1  every entry of the row is built in parallel (blocks to their LU slots, row-sum contributions to scratch)
//   2  slot 0 adds the contributions in entry order, applies loads / sources -> raw diagonal block + mismatch
//   3  per sub-level, children in parallel: subtract the terms the child's own block receives (ascending child order),
//      L = (A Q_c) U_c^-1, update terms L * U(c, j) and L * x_c to scratch
//   4  diagonal / upper entries in parallel subtract their terms in ascending child order; slot 0 subtracts the L * x_c terms
//   5  slot 0 factorises the diagonal block and forward-substitutes; 6  U blocks in parallel
template <int T, int B, Mode mode, bool SP, bool REG = false>
__device__ void wide_up_row(DevStructure const& s, TileB<T, B, SP> const& t, int w, int slot, int n_slot, bool active,
                            bool& singular, bool check_now = false) {
    constexpr int N = 2 * B, NN = N * N;
    int32_t const* const tab = s.wide_table + 8 * w;
    int const row = __ldg(tab), n_sub = __ldg(tab + 1);
    int32_t const* const sub_ptr = s.wide_data + __ldg(tab + 2);
    int32_t const* const order = s.wide_data + __ldg(tab + 3);
    int32_t const* const in_ptr = s.wide_data + __ldg(tab + 4);
    int32_t const* const in_idx = s.wide_data + __ldg(tab + 5);
    int const rb = __ldg(s.row_ptr + row), re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    int const n_entries = re - rb, n_lower = dg - rb;
    int const upd_base = __ldg(s.upd_ptr + rb);
    auto sub_terms = [&](int entry_pos, double* blk) { // blk -= incoming terms, ascending child order
        for (int q = __ldg(in_ptr + entry_pos), qe = __ldg(in_ptr + entry_pos + 1); q < qe; ++q) {
            double const* term = t.wide_terms + (size_t)__ldg(in_idx + q) * NN * T;
#pragma unroll
            for (int i = 0; i < NN; ++i) blk[i] -= term[(size_t)i * T];
        }
    };
    // 1
    if (active) {
        double uir[B], uii[B];
        t.load_u(row, uir, uii);
        for (int idx = slot; idx < n_entries; idx += n_slot) {
            double blk[NN], contrib[N];
            build_entry<T, B, mode, SP>(s, t, uir, uii, rb + idx, blk, contrib, row);
            t.store_blk(rb + idx, blk);
#pragma unroll
            for (int i = 0; i < N; ++i) t.wide_sum[(size_t)(idx * N + i) * T] = contrib[i];
        }
    }
    __syncthreads();
    // 2
    if (active && slot == 0) {
        double uir[B], uii[B];
        t.load_u(row, uir, uii);
        double acc[N], d[NN];
        [[maybe_unused]] BusControl ctl{};
        [[maybe_unused]] int viol = 0;
        if constexpr (REG && mode == Mode::newton) {
            ctl = bus_control<T, B, SP>(s, t, row);
            viol = t.qviol[(size_t)row * T];
        }
        constexpr int n_pass = (REG && mode == Mode::newton) ? 2 : 1;
#pragma unroll 1
        for (int pass = 0; pass < n_pass; ++pass) {
#pragma unroll
            for (int i = 0; i < N; ++i) acc[i] = 0.0;
            for (int idx = 0; idx < n_entries; ++idx) {
                if (__ldg(s.map_y + rb + idx) < 0) continue;
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] -= t.wide_sum[(size_t)(idx * N + i) * T];
            }
            t.load_blk(dg, d);
            finish_diag<T, B, mode, SP, REG>(s, t, row, uir, uii, acc, d, viol);
            if constexpr (REG && mode == Mode::newton) {
                if (pass == 0 && check_now && ctl.has_limits && viol == 0) {
                    viol = check_q_limit<T, B, SP>(s, t, row, ctl, acc);
                    if (viol != 0) {
                        t.qviol[(size_t)row * T] = (uint8_t)viol;
                        continue;
                    }
                }
            }
            break;
        }
        if constexpr (REG && mode == Mode::newton) {
            if (ctl.regulated && viol == 0) { // PV row: the built blocks of the row lose their Q rows before phase 3 reads them
                pv_diag<T, B, SP>(t, row, d, acc);
                for (int k = rb; k < re; ++k)
                    if (k != dg) zero_pv_rows<T, B, SP>(t, k);
            }
        }
        t.store_blk(dg, d);
#pragma unroll
        for (int i = 0; i < N; ++i) t.xvec[(size_t)(row * N + i) * T] = acc[i];
    }
    if constexpr (REG && mode == Mode::newton) __syncthreads(); // phase 2 may have rewritten blocks that phase 3 loads
    // 3 (without regulators the diagonal block is the only thing phase 2 writes, so it needs no barrier of its own)
    for (int sl = 0; sl < n_sub; ++sl) {
        if (active) {
            for (int oi = __ldg(sub_ptr + sl) + slot; oi < __ldg(sub_ptr + sl + 1); oi += n_slot) {
                int const pos = __ldg(order + oi);
                int const e = rb + pos;
                int const c = __ldg(s.col_idx + e);
                double a[NN], piv[NN], l[NN];
                t.load_blk(e, a);
                sub_terms(pos, a);
                t.load_blk(__ldg(s.diag + c), piv);
                uint8_t qc[N];
                t.get_q(c, qc);
#pragma unroll
                for (int i = 0; i < N; ++i)
#pragma unroll
                    for (int r = 0; r < N; ++r) l[i * N + r] = a[qc[i] * N + r];
                for (int idx = 0; idx < N; ++idx) {
                    for (int prev = 0; prev < idx; ++prev) {
                        double const uv = piv[idx * N + prev];
#pragma unroll
                        for (int r = 0; r < N; ++r) l[idx * N + r] -= uv * l[prev * N + r];
                    }
                    double const dd = piv[idx * N + idx];
#pragma unroll
                    for (int r = 0; r < N; ++r) l[idx * N + r] /= dd;
                }
                for (int q = __ldg(s.upd_ptr + e), qe = __ldg(s.upd_ptr + e + 1); q < qe; ++q) {
                    double ub[NN];
                    t.load_blk(__ldg(s.upd_u + q), ub);
                    double* term = t.wide_terms + (size_t)(q - upd_base) * NN * T;
                    for (int cc = 0; cc < N; ++cc)
#pragma unroll
                        for (int r = 0; r < N; ++r) {
                            double sum = l[0 * N + r] * ub[cc * N + 0];
#pragma unroll
                            for (int k = 1; k < N; ++k) sum += l[k * N + r] * ub[cc * N + k];
                            term[(size_t)(cc * N + r) * T] = sum;
                        }
                }
                double yc[N];
#pragma unroll
                for (int i = 0; i < N; ++i) yc[i] = t.xvec[(size_t)(c * N + i) * T];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double sum = l[0 * N + r] * yc[0];
#pragma unroll
                    for (int k = 1; k < N; ++k) sum += l[k * N + r] * yc[k];
                    t.wide_rhs[(size_t)(pos * N + r) * T] = sum;
                }
            }
        }
        __syncthreads();
    }
    // 4
    if (active) {
        for (int pos = n_lower + slot; pos < n_entries; pos += n_slot) {
            double blk[NN];
            t.load_blk(rb + pos, blk);
            sub_terms(pos, blk);
            t.store_blk(rb + pos, blk);
        }
        if (slot == n_slot - 1) { // the L * x_c terms in ascending child order (another slot than the diagonal block's)
            double acc[N];
#pragma unroll
            for (int i = 0; i < N; ++i) acc[i] = t.xvec[(size_t)(row * N + i) * T];
            for (int pos = 0; pos < n_lower; ++pos) {
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] -= t.wide_rhs[(size_t)(pos * N + i) * T];
            }
#pragma unroll
            for (int i = 0; i < N; ++i) t.xvec[(size_t)(row * N + i) * T] = acc[i];
        }
    }
    __syncthreads();
    // 5
    if (active && slot == 0) {
        double d[NN], acc[N], xr[N];
        uint8_t p[N], q[N];
        t.load_blk(dg, d);
        singular |= factorize_block<N>(d, p, q);
        t.store_blk(dg, d);
        t.set_perm(row, p, q);
#pragma unroll
        for (int i = 0; i < N; ++i) acc[i] = t.xvec[(size_t)(row * N + i) * T];
#pragma unroll
        for (int i = 0; i < N; ++i) xr[p[i]] = acc[i];
        for (int idx = 0; idx < N; ++idx)
            for (int prev = 0; prev < idx; ++prev) xr[idx] -= d[prev * N + idx] * xr[prev];
#pragma unroll
        for (int i = 0; i < N; ++i) t.xvec[(size_t)(row * N + i) * T] = xr[i];
    }
    __syncthreads();
    // 6
    if (active) {
        for (int e = dg + 1 + slot; e < re; e += n_slot) {
            double a[NN], ub[NN], d[NN];
            uint8_t p[N];
            t.load_blk(e, a);
            t.load_blk(dg, d);
            t.get_p(row, p);
#pragma unroll
            for (int cc = 0; cc < N; ++cc)
#pragma unroll
                for (int i = 0; i < N; ++i) ub[cc * N + p[i]] = a[cc * N + i];
            for (int idx = 0; idx < N; ++idx)
                for (int prev = 0; prev < idx; ++prev) {
                    double const lv = d[prev * N + idx];
#pragma unroll
                    for (int cc = 0; cc < N; ++cc) ub[cc * N + idx] -= lv * ub[cc * N + prev];
                }
            t.store_blk(e, ub);
        }
    }
    __syncthreads();
}

} // namespace blk
} // namespace pgmb
