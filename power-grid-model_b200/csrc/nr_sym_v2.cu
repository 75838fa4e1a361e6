// Symmetric batched Newton-Raphson, version 2: same arithmetic as nr_sym.cu (see there for the reference citations), with
//   * the per-row index data ("row programs", symbolic.hpp) staged once per thread block into shared memory by one TMA bulk
//     copy (cp.async.bulk + mbarrier) instead of dependent global index loads in every row task,
//   * "tree rows" (all rows of a radial grid) processed entirely in registers: the row's Jacobian blocks are never written
//     to the scratch matrix, only the factorised diagonal block, the U block towards the parent and the forward-substituted
//     right-hand side leave the thread; every global load of a row task is issued before its first store.
// Rows that are not tree rows (the cyclic core of a meshed grid) run the generic row task of nr_sym_common.cuh.
#include "block_common.cuh"
#include "nr_sym_common.cuh"

#include <cuda_runtime.h>

#include <cstdio>
#include <stdexcept>

namespace pgmb {

using namespace nrsym;

namespace {

// ---- TMA bulk copy global -> shared ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(void const* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void stage_program(int32_t* dst, int32_t const* src, uint32_t bytes, uint64_t* mbar) {
    uint32_t const bar = smem_u32(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        uint32_t const chunk = 32768;
        for (uint32_t off = 0; off < bytes; off += chunk) {
            uint32_t const n = min(chunk, bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(reinterpret_cast<unsigned char*>(dst) + off)),
                         "l"(reinterpret_cast<unsigned char const*>(src) + off), "r"(n), "r"(bar)
                         : "memory");
        }
    }
    // every thread waits for phase 0 of the barrier
    uint32_t done = 0;
    while (done == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar)
                     : "memory");
    }
}

// ---- tree-row tasks ---------------------------------------------------------------------------------------------------
template <int T> struct TileR { // restrict-qualified view: the arrays never alias each other
    double* __restrict__ jac;
    double* __restrict__ xvec;
    double* __restrict__ pol;
    double* __restrict__ u;
    uint8_t* __restrict__ perm;
    double const* __restrict__ sinj;
    double const* __restrict__ usrc;
    int32_t const* ovr_entry{nullptr}; // branch-outage overlay of the lane's scenario (nr_sym_common.cuh: load_y / is_dead)
    double const* ovr_y{nullptr};
    uint8_t const* dead{nullptr};
    uint8_t const* lg_status{nullptr}; // REG instantiations (PV buses, nr_sym_common.cuh)
    uint8_t* qviol{nullptr};
    int ovr_n{0};
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

template <int T, Mode mode, bool OVL, bool REG = false>
__device__ __forceinline__ bool up_tree_row(DevStructure const& s, TileR<T> const& t, int32_t const* __restrict__ rec,
                                            [[maybe_unused]] bool check_now = false) {
    int const row = rec[0], k_d = rec[1], ky_d = rec[2];
    int const n_lower = rec[3] & 0xfff, n_upper = (rec[3] >> 12) & 0xfff;
    int const lg0 = rec[4] & 0xffffff, n_lg = (rec[4] >> 24) & 0x7f;
    int const sr0 = rec[5] & 0xffffff, n_src = (rec[5] >> 24) & 0x7f;
    int32_t const* __restrict__ lower = rec + 6;
    int32_t const* __restrict__ upper = lower + 4 * n_lower;

    double const uir = t.u[(size_t)(row * 2) * T], uii = t.u[(size_t)(row * 2 + 1) * T];
    double acc0 = 0.0, acc1 = 0.0;
    Blk d{0.0, 0.0, 0.0, 0.0}, ub{0.0, 0.0, 0.0, 0.0};

    // pass 1: row sums in entry order (lower entries, diagonal, upper entry); only the diagonal and upper blocks are kept
    if constexpr (mode == Mode::newton) {
#pragma unroll 2
        for (int e = 0; e < n_lower; ++e) {
            int const c = lower[4 * e], ky = lower[4 * e + 1];
            if (ky >= 0 && !(OVL && t.dead != nullptr && (t.dead[row] != 0 || t.dead[c] != 0))) {
                double h, n, yr, yi;
                load_y<OVL>(s, t, ky, yr, yi);
                hnml(yr, yi, uir, uii, t.u[(size_t)(c * 2) * T],
                     t.u[(size_t)(c * 2 + 1) * T], h, n);
                acc0 -= n;
                acc1 -= h;
            }
        }
    }
    {
        double yr, yi;
        load_y<OVL>(s, t, ky_d, yr, yi);
        if constexpr (mode == Mode::newton) {
            double h, n;
            hnml(yr, yi, uir, uii, uir, uii, h, n);
            d = {h, -n, n, h};
            acc0 -= n;
            acc1 -= h;
        } else {
            d = {yr, yi, -yi, yr};
        }
    }
    int k_u = -1;
    if (n_upper != 0) {
        k_u = upper[0];
        int const j = upper[1], ky = upper[2];
        if (ky >= 0 && !(OVL && t.dead != nullptr && (t.dead[row] != 0 || t.dead[j] != 0))) {
            double yr, yi;
            load_y<OVL>(s, t, ky, yr, yi);
            if constexpr (mode == Mode::newton) {
                double h, n;
                hnml(yr, yi, uir, uii, t.u[(size_t)(j * 2) * T], t.u[(size_t)(j * 2 + 1) * T], h, n);
                ub = {h, -n, n, h};
                acc0 -= n;
                acc1 -= h;
            } else {
                ub = {yr, yi, -yi, yr};
            }
        }
    }
    if constexpr (mode == Mode::newton) {
        d.a00 += acc1;
        d.a01 += -acc0;
        d.a10 += -acc0;
        d.a11 += -acc1;
    }
    // loads and sources on the diagonal (same statements as the generic row task)
    double const v = t.pol[(size_t)(row * 2 + 1) * T];
    // REG (PV buses): see the generic row task (nr_sym_common.cuh up_row)
    [[maybe_unused]] PvControl ctl{false, false, 0.0, 0.0, 0.0};
    [[maybe_unused]] int viol = 0;
    [[maybe_unused]] Blk const d_rows = d;
    [[maybe_unused]] double const rows0 = acc0, rows1 = acc1;
    if constexpr (REG && mode == Mode::newton) {
        ctl = pv_control<T>(s, t, lg0, n_lg, n_src);
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
    for (int lg = lg0; lg < lg0 + n_lg; ++lg) {
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
            double const ylr = -ps, yli = qs;
            d.a01 += -yli;
            d.a00 += ylr;
            d.a11 += ylr;
            d.a10 += yli;
        }
    }
    for (int sr = sr0; sr < sr0 + n_src; ++sr) {
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
            viol = check_q_limit<T>(s, t, lg0, n_lg, ctl, acc1);
            if (viol != 0) {
                t.qviol[(size_t)row * T] = (uint8_t)viol;
                continue;
            }
        }
    }
    break;
    }
    [[maybe_unused]] bool pv_row = false;
    if constexpr (REG && mode == Mode::newton) {
        pv_row = ctl.regulated && viol == 0;
        if (pv_row) { // PV row (:549-587): the Q row of every block of the row goes, |V| is held
            d.a10 = 0.0;
            d.a11 = v;
            acc1 = 0.0;
            ub.a10 = 0.0;
            ub.a11 = 0.0;
        }
    }

    if (is_dead<OVL>(t, row)) { // bus without supply: identity row, zero right-hand side (its entries were skipped above)
        d = {1.0, 0.0, 0.0, 1.0};
        acc0 = 0.0;
        acc1 = 0.0;
    }
    // pass 2: eliminate against the children; the lower block is rebuilt in registers (never stored)
#pragma unroll 2
    for (int e = 0; e < n_lower; ++e) {
        int const c = lower[4 * e], ky = lower[4 * e + 1], kd_c = lower[4 * e + 2], k_uc = lower[4 * e + 3];
        Blk a{0.0, 0.0, 0.0, 0.0};
        if (ky >= 0 && !(OVL && t.dead != nullptr && (t.dead[row] != 0 || t.dead[c] != 0))) {
            double yr, yi;
            load_y<OVL>(s, t, ky, yr, yi);
            if constexpr (mode == Mode::newton) {
                double h, n;
                hnml(yr, yi, uir, uii, t.u[(size_t)(c * 2) * T], t.u[(size_t)(c * 2 + 1) * T], h, n);
                a = {h, -n, n, h};
            } else {
                a = {yr, yi, -yi, yr};
            }
        }
        if constexpr (REG && mode == Mode::newton) {
            if (pv_row) {
                a.a10 = 0.0;
                a.a11 = 0.0;
            }
        }
        Blk const piv = t.load_blk(kd_c);
        Blk const uc = t.load_blk(k_uc);
        uint8_t const pc = t.perm[(size_t)c * T];
        double const y0 = t.xvec[(size_t)(c * 2) * T], y1 = t.xvec[(size_t)(c * 2 + 1) * T];
        if (pc & 2) {
            double x = a.a00;
            a.a00 = a.a01;
            a.a01 = x;
            x = a.a10;
            a.a10 = a.a11;
            a.a11 = x;
        }
        Blk l;
        l.a00 = a.a00 / piv.a00;
        l.a10 = a.a10 / piv.a00;
        l.a01 = (a.a01 - piv.a01 * l.a00) / piv.a11;
        l.a11 = (a.a11 - piv.a01 * l.a10) / piv.a11;
        d.a00 -= l.a00 * uc.a00 + l.a01 * uc.a10;
        d.a10 -= l.a10 * uc.a00 + l.a11 * uc.a10;
        d.a01 -= l.a00 * uc.a01 + l.a01 * uc.a11;
        d.a11 -= l.a10 * uc.a01 + l.a11 * uc.a11;
        acc0 -= l.a00 * y0 + l.a01 * y1;
        acc1 -= l.a10 * y0 + l.a11 * y1;
    }

    int pr, pcq;
    bool const singular = factor_diag(d, pr, pcq);
    t.store_blk(k_d, d);
    t.perm[(size_t)row * T] = static_cast<uint8_t>(pr | (pcq << 1));
    if (k_u >= 0) {
        if (pr) {
            double x = ub.a00;
            ub.a00 = ub.a10;
            ub.a10 = x;
            x = ub.a01;
            ub.a01 = ub.a11;
            ub.a11 = x;
        }
        ub.a10 -= d.a10 * ub.a00;
        ub.a11 -= d.a10 * ub.a01;
        t.store_blk(k_u, ub);
    }
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

template <int T, Mode mode, bool REG = false>
__device__ __forceinline__ double down_tree_row(TileR<T> const& t, int32_t const* __restrict__ rec, [[maybe_unused]] DevStructure const* s = nullptr) {
    int const row = rec[0], k_d = rec[1];
    int const n_lower = rec[3] & 0xfff, n_upper = (rec[3] >> 12) & 0xfff;
    double y0 = t.xvec[(size_t)(row * 2) * T], y1 = t.xvec[(size_t)(row * 2 + 1) * T];
    Blk const d = t.load_blk(k_d);
    uint8_t const pm = t.perm[(size_t)row * T];
    double th = 0.0, v = 0.0, our = 0.0, oui = 0.0;
    if constexpr (mode == Mode::newton) {
        th = t.pol[(size_t)(row * 2) * T];
        v = t.pol[(size_t)(row * 2 + 1) * T];
        our = t.u[(size_t)(row * 2) * T];
        oui = t.u[(size_t)(row * 2 + 1) * T];
    }
    if (n_upper != 0) {
        int32_t const* __restrict__ upper = rec + 6 + 4 * n_lower;
        int const k_u = upper[0], j = upper[1];
        Blk const ub = t.load_blk(k_u);
        double const x0 = t.xvec[(size_t)(j * 2) * T], x1 = t.xvec[(size_t)(j * 2 + 1) * T];
        y0 -= ub.a00 * x0 + ub.a01 * x1;
        y1 -= ub.a10 * x0 + ub.a11 * x1;
    }
    y1 /= d.a11;
    y0 -= d.a01 * y1;
    y0 /= d.a00;
    if (pm & 2) {
        double const x = y0;
        y0 = y1;
        y1 = x;
    }
    t.xvec[(size_t)(row * 2) * T] = y0;
    t.xvec[(size_t)(row * 2 + 1) * T] = y1;
    if constexpr (REG && mode == Mode::linear_init) pv_start_voltage(pv_control_of_row<T>(*s, t, row), y0, y1);
    return polar_update<T, mode>(t.pol + (size_t)(row * 2) * T, t.u + (size_t)(row * 2) * T, y0, y1, th, v, our, oui);
}

template <int T, Mode mode, bool OVL, bool REG>
__device__ __forceinline__ void sweeps_v2(DevStructure const& s, Tile<T> const& tg, TileR<T> const& t,
                                          blk::TileB<T, 1, true> const& tw, int32_t const* prog, int slot, int n_slot,
                                          bool active, bool& singular, double& dev, unsigned long long* phase, bool check_now) {
    constexpr blk::Mode wmode = mode == Mode::newton ? blk::Mode::newton : blk::Mode::linear_init;
    int32_t const* level_ptr = prog;
    int32_t const* task_off = prog + s.n_level + 1;
    long long t0 = clock64();
    for (int lv = 0; lv < s.n_level; ++lv) {
        int const b = level_ptr[lv], e = level_ptr[lv + 1];
        if (active) {
            for (int i = b + slot; i < e; i += n_slot) {
                int32_t const* rec = prog + task_off[i];
                if (s.n_wide != 0 && __ldg(s.row_is_wide + rec[0])) continue; // eliminated below by the whole block
                if (rec[3] >> 24) {
                    singular |= up_tree_row<T, mode, OVL, REG>(s, t, rec, check_now);
                } else {
                    singular |= up_row<T, mode, OVL, REG>(s, tg, rec[0], check_now);
                }
            }
        }
        __syncthreads();
        if (s.n_wide != 0)
            for (int w = __ldg(s.wide_level_ptr + lv); w < __ldg(s.wide_level_ptr + lv + 1); ++w)
                blk::wide_up_row<T, 1, wmode, true, REG>(s, tw, w, slot, n_slot, active, singular, check_now);
        if (phase != nullptr && threadIdx.x == 0) {
            long long const t1 = clock64();
            phase[lv == 0 ? 0 : 1] += (unsigned long long)(t1 - t0);
#ifdef V2_LEVEL_DEBUG
            if (blockIdx.x == 0 && mode == Mode::newton)
                printf("up level %d rows %d wide %d kcycles %lld\n", lv, e - b,
                       s.n_wide != 0 ? __ldg(s.wide_level_ptr + lv + 1) - __ldg(s.wide_level_ptr + lv) : 0, (t1 - t0) / 1000);
            t0 = clock64();
#else
            t0 = t1;
#endif
        }
    }
    for (int lv = s.n_level - 1; lv >= 0; --lv) {
        int const b = level_ptr[lv], e = level_ptr[lv + 1];
        if (active) {
            for (int i = b + slot; i < e; i += n_slot) {
                int32_t const* rec = prog + task_off[i];
                if (rec[3] >> 24) {
                    dev = fmax(dev, down_tree_row<T, mode, REG>(t, rec, &s));
                } else {
                    dev = fmax(dev, down_row<T, mode, REG>(s, tg, rec[0]));
                }
            }
        }
        __syncthreads();
        if (phase != nullptr && threadIdx.x == 0) {
            long long const t1 = clock64();
            phase[lv == 0 ? 3 : 2] += (unsigned long long)(t1 - t0);
            t0 = t1;
        }
    }
}

} // namespace

// OVL: batches with a branch-outage overlay (kernels.cuh: DevOverlay); capped at 128 registers like the plain kernel uses
// REG: the grid has voltage regulators (PV buses; nr_sym_common.cuh)
template <int T, bool OVL, bool REG> __global__ void __launch_bounds__(512, 1) nr_sym_v2_kernel(DevStructure s, DevBatch b, SolveOptions opt, int prog_in_smem) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long sh_dev[T];
    __shared__ int sh_singular[T];
    __shared__ int sh_has_limits[T]; // REG: the scenario has a PV bus with a usable Q limit (limit check from iteration 2 on)
    __shared__ __align__(8) uint64_t sh_mbar;
    int const lane = threadIdx.x % T;
    int const slot = threadIdx.x / T;
    int const n_slot = blockDim.x / T;
    int const tile = blockIdx.x;
    int64_t const scn = (int64_t)tile * T + lane;
    bool const valid = scn < b.n_scn;

    int32_t const* prog = s.prog;
    if (prog_in_smem) {
        stage_program(reinterpret_cast<int32_t*>(smem_raw), s.prog, (uint32_t)s.prog_words * 4u, &sh_mbar);
        prog = reinterpret_cast<int32_t const*>(smem_raw);
    }

    Tile<T> tg;
    tg.jac = b.jac + (size_t)tile * s.nnz_lu * 4 * T + lane;
    tg.xvec = b.xvec + (size_t)tile * s.n_bus * 2 * T + lane;
    tg.pol = b.pol + (size_t)tile * s.n_bus * 2 * T + lane;
    tg.u = b.u + (size_t)tile * s.n_bus * 2 * T + lane;
    tg.perm = b.perm + (size_t)tile * s.n_bus * T + lane;
    tg.sinj = b.sinj + (size_t)tile * s.n_load_gen * 2 * T + lane;
    tg.usrc = b.usrc + (size_t)tile * s.n_source * 2 * T + lane;
    if (OVL && b.ovl.entry != nullptr && valid) { // branch-outage overlay of this lane's scenario
        tg.ovr_n = 4 * b.ovl.n_branch;
        tg.ovr_entry = b.ovl.entry + scn * tg.ovr_n;
        tg.ovr_y = b.ovl.y + scn * tg.ovr_n * 2;
        if (b.ovl.dead_off != nullptr && b.ovl.dead_off[scn] >= 0) tg.dead = b.ovl.dead + (size_t)b.ovl.dead_off[scn] * s.n_bus;
    }
    if (REG) {
        tg.lg_status = b.lg_status + (size_t)tile * s.n_load_gen * T + lane;
        tg.qviol = b.qviol + (size_t)tile * s.n_bus * T + lane;
    }
    TileR<T> const t{tg.jac, tg.xvec, tg.pol, tg.u, tg.perm, tg.sinj, tg.usrc, tg.ovr_entry, tg.ovr_y, tg.dead, tg.lg_status, tg.qviol, tg.ovr_n};
    blk::TileB<T, 1, true> tw;
    tw.ovr_entry = tg.ovr_entry;
    tw.ovr_n = tg.ovr_n;
    tw.ovr_y = tg.ovr_y;
    tw.dead = tg.dead;
    tw.jac = tg.jac;
    tw.xvec = tg.xvec;
    tw.pol = tg.pol;
    tw.u = tg.u;
    tw.perm = tg.perm;
    tw.sinj = tg.sinj;
    tw.usrc = tg.usrc;
    tw.lg_status = tg.lg_status;
    tw.qviol = tg.qviol;
    tw.wide_terms = b.wide_terms ? b.wide_terms + (size_t)tile * s.wide_max_upd * 4 * T + lane : nullptr;
    tw.wide_rhs = b.wide_rhs ? b.wide_rhs + (size_t)tile * s.wide_max_lower * 2 * T + lane : nullptr;
    tw.wide_sum = b.wide_sum ? b.wide_sum + (size_t)tile * s.wide_max_entries * 2 * T + lane : nullptr;

    if (threadIdx.x < T) {
        sh_dev[threadIdx.x] = 0ull;
        sh_singular[threadIdx.x] = 0;
        sh_has_limits[threadIdx.x] = 0;
    }
    __syncthreads();
    if constexpr (REG) { // set_bus_types_and_q_limits (newton_raphson_pf_solver.hpp:400-444); no limit has been hit yet
        if (valid) {
            for (int row = slot; row < s.n_bus; row += n_slot) {
                tg.qviol[(size_t)row * T] = 0;
                if (__ldg(s.lg_ptr + row) != __ldg(s.lg_ptr + row + 1) && pv_control_of_row<T>(s, tg, row).has_limits) sh_has_limits[lane] = 1;
            }
        }
        __syncthreads();
    }

    bool done = !valid;
    int status = kStatusOk;
    int num_iter = 0;
    double max_dev = INFINITY;
    unsigned long long* const phase = b.phase_cycles ? b.phase_cycles + tile * 16 : nullptr;
    {
        bool singular = false;
        double dev = 0.0;
        sweeps_v2<T, Mode::linear_init, OVL, REG>(s, tg, t, tw, prog, slot, n_slot, !done, singular, dev, phase, false);
        if (singular) sh_singular[lane] = 1;
        __syncthreads();
        if (!done && sh_singular[lane]) {
            status = kStatusSingular;
            done = true;
        }
    }
    while (true) {
        if (!done) {
            if (num_iter == opt.max_iter) {
                status = kStatusDiverged;
                done = true;
            } else {
                ++num_iter;
            }
        }
        if (!__syncthreads_or(!done)) break;
        bool singular = false;
        double dev = 0.0;
        sweeps_v2<T, Mode::newton, OVL, REG>(s, tg, t, tw, prog, slot, n_slot, !done, singular, dev, phase ? phase + 4 : nullptr, REG && num_iter >= 2);
        if (!done) {
            if (singular) sh_singular[lane] = 1;
            atomicMax(&sh_dev[lane], (unsigned long long)__double_as_longlong(dev));
        }
        __syncthreads();
        if (!done) {
            if (sh_singular[lane]) {
                status = kStatusSingular;
                done = true;
            } else {
                max_dev = __longlong_as_double((long long)sh_dev[lane]);
                if (!(max_dev > opt.err_tol)) {
                    if (REG && sh_has_limits[lane] && num_iter < 2) {
                        max_dev = INFINITY; // converged before the limit check: one more iteration (:343-347)
                    } else {
                        done = true;
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < T) sh_dev[threadIdx.x] = 0ull;
    }
    if (slot == 0 && valid) {
        b.status[scn] = status;
        b.n_iter[scn] = num_iter;
        b.max_dev[scn] = max_dev;
    }
}

template <int T>
static void launch_v2_t(DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot, cudaStream_t st) {
    size_t const prog_bytes = (size_t)s.prog_words * 4;
    int dev = 0, max_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    bool const in_smem = prog_bytes + 1024 <= (size_t)max_optin;
    size_t const dyn = in_smem ? prog_bytes : 0;
    if (s.lg_reg != nullptr) { // voltage regulators (the engine sends overlay batches of such grids to the block kernel)
        if (b.ovl.entry != nullptr || b.qviol == nullptr || b.lg_status == nullptr) {
            throw std::logic_error("nr_sym_v2: a grid with voltage regulators needs the qviol / lg_status buffers and no overlay");
        }
        cudaFuncSetAttribute(nr_sym_v2_kernel<T, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - 1024);
        nr_sym_v2_kernel<T, false, true><<<b.n_tile, T * n_slot, dyn, st>>>(s, b, opt, in_smem ? 1 : 0);
    } else if (b.ovl.entry != nullptr) {
        cudaFuncSetAttribute(nr_sym_v2_kernel<T, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - 1024);
        nr_sym_v2_kernel<T, true, false><<<b.n_tile, T * n_slot, dyn, st>>>(s, b, opt, in_smem ? 1 : 0);
    } else {
        cudaFuncSetAttribute(nr_sym_v2_kernel<T, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - 1024);
        nr_sym_v2_kernel<T, false, false><<<b.n_tile, T * n_slot, dyn, st>>>(s, b, opt, in_smem ? 1 : 0);
    }
}

void launch_nr_sym_v2(int tile_width, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot,
                      cudaStream_t st) {
    count_kernel_launch();
    switch (tile_width) {
    case 4: launch_v2_t<4>(s, b, opt, n_slot, st); break;
    case 8: launch_v2_t<8>(s, b, opt, n_slot, st); break;
    case 16: launch_v2_t<16>(s, b, opt, n_slot, st); break;
    default: launch_v2_t<32>(s, b, opt, n_slot, st); break;
    }
}

} // namespace pgmb
