// Shared device helpers of the symmetric NR kernels: 2x2 block arithmetic restating the reference
// (sparse_lu_solver.hpp:86-165, 171-200; newton_raphson_pf_solver.hpp:462-471) and the generic row tasks.
#pragma once

#include "kernels.cuh"

#include <cfloat>
#include <cmath>

namespace pgmb {
namespace nrsym {

constexpr int kStatusOk = 0, kStatusDiverged = 1, kStatusSingular = 2;

enum class Mode { linear_init, newton };

struct Blk { // 2x2 block, named by (row, col)
    double a00, a10, a01, a11;
};

__device__ __forceinline__ bool not_normal(double x) { return !(fabs(x) >= DBL_MIN) || isinf(x); }

// power_flow_ij = (ui * conj(uj)) * conj(y);  H = L = imag, N = -M = real      (newton_raphson_pf_solver.hpp:462-471)
__device__ __forceinline__ void hnml(double yr, double yi, double uir, double uii, double ujr, double uji, double& h,
                                     double& n) {
    double const cr = ujr, ci = -uji;
    double const ar = uir * cr - uii * ci;
    double const ai = uir * ci + uii * cr;
    double const dr = yr, di = -yi;
    n = ar * dr - ai * di;
    h = ar * di + ai * dr;
}

// full-pivot LU of a 2x2 block in place (DenseLUFactor::factorize_block_in_place, sparse_lu_solver.hpp:86-165):
// pivot = first maximum of |a|^2 in column-major order; pr / pcq = row / column transposition of step 0.
__device__ __forceinline__ bool factor_diag(Blk& d, int& pr, int& pcq) {
    double const s00 = d.a00 * d.a00, s10 = d.a10 * d.a10, s01 = d.a01 * d.a01, s11 = d.a11 * d.a11;
    double best = s00;
    pr = 0;
    pcq = 0;
    if (s10 > best) {
        best = s10;
        pr = 1;
        pcq = 0;
    }
    if (s01 > best) {
        best = s01;
        pr = 0;
        pcq = 1;
    }
    if (s11 > best) {
        best = s11;
        pr = 1;
        pcq = 1;
    }
    bool singular = (best == 0.0);
    if (pr) {
        double x = d.a00;
        d.a00 = d.a10;
        d.a10 = x;
        x = d.a01;
        d.a01 = d.a11;
        d.a11 = x;
    }
    if (pcq) {
        double x = d.a00;
        d.a00 = d.a01;
        d.a01 = x;
        x = d.a10;
        d.a10 = d.a11;
        d.a11 = x;
    }
    d.a10 /= d.a00;
    d.a11 -= d.a10 * d.a01;
    // reference: max(sqrt(best), sqrt(|d11|^2)); sqrt is monotone and correctly rounded, so one square root of the larger
    // argument gives the same bits
    double const max_pivot = sqrt(fmax(best, d.a11 * d.a11));
    double const threshold = DBL_EPSILON * max_pivot;
    singular = singular || fabs(d.a00) < threshold || not_normal(d.a00) || fabs(d.a11) < threshold || not_normal(d.a11);
    return singular;
}

// iterate_unknown (newton_raphson_pf_solver.hpp:325-349) / start values from the linear solve (:291-302) for one bus.
// pol -> theta, pol + T -> V ; u -> re, u + T -> im.  Returns |U_new - U_old| (newton) or 0.
template <int T, Mode mode>
__device__ __forceinline__ double polar_update(double* pol, double* u, double x0, double x1, double theta, double v,
                                               double old_re, double old_im) {
    if constexpr (mode == Mode::newton) {
        theta += x0;
        v += v * x1;
        double sn, cs;
        sincos(theta, &sn, &cs);
        double const nr = v * cs, ni = v * sn;
        double const dr = nr - old_re, di = ni - old_im;
        pol[0] = theta;
        pol[T] = v;
        u[0] = nr;
        u[T] = ni;
        return sqrt(dr * dr + di * di);
    } else {
        u[0] = x0;
        u[T] = x1;
        pol[T] = sqrt(x0 * x0 + x1 * x1);
        pol[0] = atan2(x1, x0);
        return 0.0;
    }
}

// Branch-outage overlay of a lane's scenario (kernels.cuh: DevOverlay): replaced Y-bus entry values and the buses that lost
// their supply.  All null for ordinary batches.
// OVL is a compile-time switch: the ordinary instantiations carry none of this code.
template <bool OVL, class TileT>
__device__ __forceinline__ void load_y(DevStructure const& s, TileT const& t, int ky, double& yr, double& yi) {
    double const* p = s.ydata + 2 * ky;
    if (OVL && t.ovr_entry != nullptr) {
        for (int j = 0; j < t.ovr_n; ++j)
            if (t.ovr_entry[j] == ky) p = t.ovr_y + 2 * j;
    }
    yr = __ldg(p);
    yi = __ldg(p + 1);
}
template <bool OVL, class TileT> __device__ __forceinline__ bool is_dead(TileT const& t, int bus) {
    return OVL && t.dead != nullptr && t.dead[bus] != 0;
}


// ---- PV buses (voltage regulators; newton_raphson_pf_solver.hpp:400-452, 549-587, 605-742) -----------------------------------
// Shared by the symmetric kernels' REG instantiations.  The same decisions, in the same order, as the generic block kernel takes
// for B = 1 (block_common.cuh: bus_control, check_q_limit, pv_diag, zero_pv_rows): results are bit-identical to it.
// TileT needs lg_status (status of every load_gen) and sinj.
constexpr double kQTol = 1e-8;
struct PvControl {
    bool regulated, has_limits;
    double u_ref, q_min, q_max;
};
template <int T, class TileT> __device__ __forceinline__ bool lg_regulating(DevStructure const& s, TileT const& t, int lg, int& reg) {
    reg = __ldg(s.lg_reg + lg);
    return reg >= 0 && __ldg(s.reg_param + 4 * reg) != 0.0 && t.lg_status[(size_t)lg * T] != 0;
}
template <int T, class TileT> __device__ __forceinline__ PvControl pv_control(DevStructure const& s, TileT const& t, int lg0, int n_lg, int n_src) {
    PvControl c{false, false, 0.0, 0.0, 0.0};
    if (n_src != 0) return c; // slack bus
    for (int lg = lg0; lg < lg0 + n_lg; ++lg) {
        int reg;
        if (lg_regulating<T>(s, t, lg, reg)) {
            c.regulated = true;
            c.u_ref = __ldg(s.reg_param + 4 * reg + 1);
            c.q_min += __ldg(s.reg_param + 4 * reg + 2);
            c.q_max += __ldg(s.reg_param + 4 * reg + 3);
        }
    }
    c.has_limits = c.regulated && (!isnan(c.q_min) || !isnan(c.q_max));
    return c;
}
template <int T, class TileT> __device__ __forceinline__ PvControl pv_control_of_row(DevStructure const& s, TileT const& t, int row) {
    int const lg0 = __ldg(s.lg_ptr + row), lg1 = __ldg(s.lg_ptr + row + 1);
    return pv_control<T>(s, t, lg0, lg1 - lg0, __ldg(s.src_ptr + row + 1) - __ldg(s.src_ptr + row));
}
// specified Q of a load_gen as the row build sees it: ignored for a regulating generator in the linear start, the regulator's
// limit once its bus ran into one (viol: 1 lower, 2 upper)
template <int T, Mode mode, class TileT>
__device__ __forceinline__ double regulated_q(DevStructure const& s, TileT const& t, int lg, int viol, double qs) {
    int reg;
    if (!lg_regulating<T>(s, t, lg, reg)) return qs;
    if (mode == Mode::linear_init) return 0.0;
    return viol != 0 ? __ldg(s.reg_param + 4 * reg + (viol == 2 ? 3 : 2)) : qs;
}
// enforce_q_limits for one PV bus (:605-704); acc1 = Q mismatch of the freshly built row.  Returns the violated limit (0 = none).
template <int T, class TileT>
__device__ __forceinline__ int check_q_limit(DevStructure const& s, TileT const& t, int lg0, int n_lg, PvControl const& c, double acc1) {
    double spec = 0.0;
    for (int lg = lg0; lg < lg0 + n_lg; ++lg) {
        int reg;
        if (lg_regulating<T>(s, t, lg, reg)) spec += t.sinj[(size_t)(lg * 2 + 1) * T];
    }
    double const q_total = spec - acc1;
    if (!isnan(c.q_max) && q_total > c.q_max + kQTol) return 2;
    if (!isnan(c.q_min) && q_total < c.q_min - kQTol) return 1;
    return 0;
}
// a PV bus starts at its reference magnitude after the linear start: u = u_ref * u / |u| (:446-452)
__device__ __forceinline__ void pv_start_voltage(PvControl const& c, double& y0, double& y1) {
    if (!c.regulated) return;
    double const ax = sqrt(y0 * y0 + y1 * y1);
    double sr = 1.0, si = 0.0;
    if (ax > 0.0) {
        sr = y0 / ax;
        si = y1 / ax;
    }
    y0 = c.u_ref * sr - 0.0 * si;
    y1 = c.u_ref * si + 0.0 * sr;
}

template <int T> struct Tile {
    double* jac;
    double* xvec;
    double* pol;
    double* u;
    uint8_t* perm;
    double const* sinj;
    double const* usrc;
    int32_t const* ovr_entry{nullptr}; // [ovr_n]
    double const* ovr_y{nullptr};      // [ovr_n][2]
    uint8_t const* dead{nullptr};      // [n_bus]
    uint8_t const* lg_status{nullptr}; // REG instantiations: status of every load_gen
    uint8_t* qviol{nullptr};           // REG instantiations: Q limit each bus ran into (0 none, 1 lower, 2 upper)
    int ovr_n{0};                      // replaced Y-bus entries of the lane's scenario (4 per switched-branch slot)

    __device__ __forceinline__ Blk load_blk(int k) const {
        double const* p = jac + (size_t)k * 4 * T;
        return {p[0], p[T], p[2 * T], p[3 * T]};
    }
    __device__ __forceinline__ void store_blk(int k, Blk const& b) const {
        double* p = jac + (size_t)k * 4 * T;
        p[0] = b.a00;
        p[T] = b.a10;
        p[2 * T] = b.a01;
        p[3 * T] = b.a11;
    }
};

// ---- up-sweep row task -------------------------------------------------------------------------------------------
template <int T, Mode mode, bool OVL = false, bool REG = false>
__device__ __forceinline__ bool up_row(DevStructure const& s, Tile<T> const& t, int row, [[maybe_unused]] bool check_now = false) {
    int const rb = __ldg(s.row_ptr + row), re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    double const uir = t.u[(size_t)(row * 2) * T], uii = t.u[(size_t)(row * 2 + 1) * T];
    double acc0 = 0.0, acc1 = 0.0; // NR: -P, -Q then mismatch ; linear: rhs (re, im)
    Blk d{0.0, 0.0, 0.0, 0.0};

    // 1. build the row
    for (int k = rb; k < re; ++k) {
        int ky = __ldg(s.map_y + k);
        if (OVL && t.dead != nullptr && ky >= 0 && (t.dead[row] != 0 || t.dead[__ldg(s.col_idx + k)] != 0)) ky = -1; // no coupling
        Blk b{0.0, 0.0, 0.0, 0.0};
        if (ky >= 0) {
            double yr, yi;
            load_y<OVL>(s, t, ky, yr, yi);
            if constexpr (mode == Mode::newton) {
                int const j = __ldg(s.col_idx + k);
                double ujr = uir, uji = uii;
                if (j != row) {
                    ujr = t.u[(size_t)(j * 2) * T];
                    uji = t.u[(size_t)(j * 2 + 1) * T];
                }
                double h, n;
                hnml(yr, yi, uir, uii, ujr, uji, h, n);
                b = {h, -n, n, h};
                acc0 -= n;
                acc1 -= h;
            } else {
                b = {yr, yi, -yi, yr}; // [[G, -B], [B, G]]
            }
        }
        if (k == dg) {
            d = b;
        } else {
            t.store_blk(k, b);
        }
    }
    if constexpr (mode == Mode::newton) {
        // diagonal correction: H += -Q, N -= -P, M -= -P, L -= -Q   (newton_raphson_pf_solver.hpp:509-519)
        d.a00 += acc1;
        d.a01 += -acc0;
        d.a10 += -acc0;
        d.a11 += -acc1;
    }
    // loads
    double const v = t.pol[(size_t)(row * 2 + 1) * T];
    // REG (PV buses): loads and sources depend on the Q limit the bus may run into right now, the row sums above do not: a bus
    // that hits its limit repeats what follows with its regulating generators clamped
    [[maybe_unused]] PvControl ctl{false, false, 0.0, 0.0, 0.0};
    [[maybe_unused]] int viol = 0;
    [[maybe_unused]] Blk const d_rows = d;
    [[maybe_unused]] double const rows0 = acc0, rows1 = acc1;
    [[maybe_unused]] int const lg_first = __ldg(s.lg_ptr + row), lg_count = __ldg(s.lg_ptr + row + 1) - lg_first;
    if constexpr (REG && mode == Mode::newton) {
        ctl = pv_control_of_row<T>(s, t, row);
        viol = t.qviol[(size_t)row * T];
    }
    constexpr int n_pass = (REG && mode == Mode::newton) ? 2 : 1;
#pragma unroll 1
    for (int pass = 0; pass < n_pass; ++pass) {
    if constexpr (REG && mode == Mode::newton) {
        d = d_rows;
        acc0 = rows0;
        acc1 = rows1;
    }
    for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) {
        double const ps = t.sinj[(size_t)(lg * 2) * T];
        double qs = t.sinj[(size_t)(lg * 2 + 1) * T];
        if constexpr (REG) qs = regulated_q<T, mode>(s, t, lg, viol, qs);
        if constexpr (mode == Mode::newton) {
            int const type = __ldg(s.lg_type + lg);
            if (type == 0) {
                acc0 += ps;
                acc1 += qs;
            } else if (type == 1) {
                acc0 += ps * v * v;
                acc1 += qs * v * v;
                d.a01 += -ps * 2.0 * v * v;
                d.a11 += -qs * 2.0 * v * v;
            } else {
                acc0 += ps * v;
                acc1 += qs * v;
                d.a01 += -ps * v;
                d.a11 += -qs * v;
            }
        } else {
            // y_load = -conj(s) = (-ps, qs)   (newton_raphson_pf_solver.hpp:736-740)
            double const ylr = -ps, yli = qs;
            d.a01 += -yli;
            d.a00 += ylr;
            d.a11 += ylr;
            d.a10 += yli;
        }
    }
    // sources
    for (int sr = __ldg(s.src_ptr + row), sre = __ldg(s.src_ptr + row + 1); sr < sre; ++sr) {
        double const yr = __ldg(s.src_yref + 2 * sr), yi = __ldg(s.src_yref + 2 * sr + 1);
        double const usr = t.usrc[(size_t)(sr * 2) * T], usi = t.usrc[(size_t)(sr * 2 + 1) * T];
        if constexpr (mode == Mode::newton) {
            double hmm, nmm, hms, nms;
            hnml(yr, yi, uir, uii, uir, uii, hmm, nmm);
            hnml(-yr, -yi, uir, uii, usr, usi, hms, nms);
            double const p_cal = nmm + nms;
            double const q_cal = hmm + hms;
            Blk mm{hmm, -nmm, nmm, hmm};
            mm.a00 += -q_cal;
            mm.a01 += p_cal;
            mm.a10 += p_cal;
            mm.a11 += q_cal;
            acc0 -= p_cal;
            acc1 -= q_cal;
            d.a00 += mm.a00;
            d.a01 += mm.a01;
            d.a10 += mm.a10;
            d.a11 += mm.a11;
        } else {
            d.a01 -= yi;
            d.a00 += yr;
            d.a11 += yr;
            d.a10 += yi;
            acc0 += yr * usr - yi * usi;
            acc1 += yr * usi + yi * usr;
        }
    }
    if constexpr (REG && mode == Mode::newton) {
        if (pass == 0 && check_now && ctl.has_limits && viol == 0) {
            viol = check_q_limit<T>(s, t, lg_first, lg_count, ctl, acc1);
            if (viol != 0) {
                t.qviol[(size_t)row * T] = (uint8_t)viol;
                continue;
            }
        }
    }
    break;
    }
    if constexpr (REG && mode == Mode::newton) {
        if (ctl.regulated && viol == 0) { // PV row (:549-587): the Q row of every block of the row goes, |V| is held
            d.a10 = 0.0;
            d.a11 = v;
            acc1 = 0.0;
            for (int k = rb; k < re; ++k) {
                if (k == dg) continue;
                double* p = t.jac + (size_t)k * 4 * T;
                p[T] = 0.0;
                p[3 * T] = 0.0;
            }
        }
    }

    if (is_dead<OVL>(t, row)) { // bus without supply (branch-outage overlay): identity row, zero right-hand side -> u stays 0
        d = {1.0, 0.0, 0.0, 1.0};
        acc0 = 0.0;
        acc1 = 0.0;
    }

    // 2. eliminate against finished rows; L block lives in registers only
    for (int e = rb; e < dg; ++e) {
        int const c = __ldg(s.col_idx + e);
        int const dc = __ldg(s.diag + c);
        Blk a = t.load_blk(e);
        Blk const piv = t.load_blk(dc);
        uint8_t const pc = t.perm[(size_t)c * T];
        if (pc & 2) { // A * Q_c: swap columns
            double x = a.a00;
            a.a00 = a.a01;
            a.a01 = x;
            x = a.a10;
            a.a10 = a.a11;
            a.a11 = x;
        }
        // L = (A Q) U^-1  (right / upper triangular solve)
        Blk l;
        l.a00 = a.a00 / piv.a00;
        l.a10 = a.a10 / piv.a00;
        l.a01 = (a.a01 - piv.a01 * l.a00) / piv.a11;
        l.a11 = (a.a11 - piv.a01 * l.a10) / piv.a11;
        for (int q = __ldg(s.upd_ptr + e), qe = __ldg(s.upd_ptr + e + 1); q < qe; ++q) {
            int const ui = __ldg(s.upd_u + q), ai = __ldg(s.upd_a + q);
            Blk const ub = t.load_blk(ui);
            double const s00 = l.a00 * ub.a00 + l.a01 * ub.a10;
            double const s10 = l.a10 * ub.a00 + l.a11 * ub.a10;
            double const s01 = l.a00 * ub.a01 + l.a01 * ub.a11;
            double const s11 = l.a10 * ub.a01 + l.a11 * ub.a11;
            if (ai == dg) {
                d.a00 -= s00;
                d.a10 -= s10;
                d.a01 -= s01;
                d.a11 -= s11;
            } else {
                Blk tb = t.load_blk(ai);
                tb.a00 -= s00;
                tb.a10 -= s10;
                tb.a01 -= s01;
                tb.a11 -= s11;
                t.store_blk(ai, tb);
            }
        }
        // forward substitution with the un-permuted L block
        double const y0 = t.xvec[(size_t)(c * 2) * T], y1 = t.xvec[(size_t)(c * 2 + 1) * T];
        acc0 -= l.a00 * y0 + l.a01 * y1;
        acc1 -= l.a10 * y0 + l.a11 * y1;
    }

    // 3. full-pivot LU of the diagonal block
    int pr, pcq;
    bool const singular = factor_diag(d, pr, pcq);
    t.store_blk(dg, d);
    t.perm[(size_t)row * T] = static_cast<uint8_t>(pr | (pcq << 1));

    // 4. U blocks: L_pp^-1 (P A)
    for (int e = dg + 1; e < re; ++e) {
        Blk a = t.load_blk(e);
        if (pr) {
            double x = a.a00;
            a.a00 = a.a10;
            a.a10 = x;
            x = a.a01;
            a.a01 = a.a11;
            a.a11 = x;
        }
        a.a10 -= d.a10 * a.a00;
        a.a11 -= d.a10 * a.a01;
        t.store_blk(e, a);
    }
    // 5. forward substitution inside the block
    if (pr) {
        double const x = acc0;
        acc0 = acc1;
        acc1 = x;
    }
    acc1 -= d.a10 * acc0;
    t.xvec[(size_t)(row * 2) * T] = acc0;
    t.xvec[(size_t)(row * 2 + 1) * T] = acc1;
    return singular;
}

// ---- down-sweep row task: returns |dU| of the bus (newton) -----------------------------------------------------------
template <int T, Mode mode, bool REG = false> __device__ __forceinline__ double down_row(DevStructure const& s, Tile<T> const& t, int row) {
    int const re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    double y0 = t.xvec[(size_t)(row * 2) * T], y1 = t.xvec[(size_t)(row * 2 + 1) * T];
    for (int e = re - 1; e > dg; --e) {
        int const j = __ldg(s.col_idx + e);
        Blk const ub = t.load_blk(e);
        double const x0 = t.xvec[(size_t)(j * 2) * T], x1 = t.xvec[(size_t)(j * 2 + 1) * T];
        y0 -= ub.a00 * x0 + ub.a01 * x1;
        y1 -= ub.a10 * x0 + ub.a11 * x1;
    }
    Blk const d = t.load_blk(dg);
    y1 /= d.a11;
    y0 -= d.a01 * y1;
    y0 /= d.a00;
    if (t.perm[(size_t)row * T] & 2) {
        double const x = y0;
        y0 = y1;
        y1 = x;
    }
    t.xvec[(size_t)(row * 2) * T] = y0;
    t.xvec[(size_t)(row * 2 + 1) * T] = y1;
    double th = 0.0, v = 0.0, our = 0.0, oui = 0.0;
    if constexpr (mode == Mode::newton) {
        th = t.pol[(size_t)(row * 2) * T];
        v = t.pol[(size_t)(row * 2 + 1) * T];
        our = t.u[(size_t)(row * 2) * T];
        oui = t.u[(size_t)(row * 2 + 1) * T];
    }
    if constexpr (REG && mode == Mode::linear_init) pv_start_voltage(pv_control_of_row<T>(s, t, row), y0, y1);
    return polar_update<T, mode>(t.pol + (size_t)(row * 2) * T, t.u + (size_t)(row * 2) * T, y0, y1, th, v, our, oui);
}

} // namespace nrsym
} // namespace pgmb
