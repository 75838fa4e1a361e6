// Symmetric batched Newton-Raphson, version 3 ("path kernel") for radial grids: same arithmetic, statement for statement,
// as nr_sym.cu / nr_sym_v2.cu (see nr_sym.cu for the reference citations), reorganised around the elimination tree:
//   build phase   every row's own blocks (diagonal, block towards the parent, blocks towards the children, mismatch) depend
//                 only on the voltages, so they are built for all rows at once with all threads busy; leaves (47 % of the
//                 rows of a feeder grid) are factorised in the same pass, and leaf children are eliminated from their parents
//                 (or their update term is precomputed where the reference subtracts it after another child's term);
//   chain phase   the elimination tree is cut into paths (symbolic.hpp: PathProgram); one thread walks a path bottom-up with
//                 the carried child's factor / U block / permutation / right-hand side in registers, the next row's operands
//                 prefetched while the current row's dependent divide chain runs.  Block barriers only between stages.
//   down sweep    the same paths top-down with the parent's solution in registers; leaves in one parallel pass at the end.
// The dependent work of a 40-node feeder is then 40 x (divide chain + 2x2 pivoted factor) instead of 40 x (barrier + L2 round
// trips + row build).  The path program is staged in shared memory by one TMA bulk copy.
#include "nr_sym_common.cuh"

#include <cuda_runtime.h>

#include <cstdlib>
#include <stdexcept>

#ifndef V3_CHAIN_AHEAD
#define V3_CHAIN_AHEAD 2
#endif
#ifndef V3_RING_DOWN
#define V3_RING_DOWN 0 // 1: the down chains read their first operands through the ring too (measured slower)
#endif
#ifndef V3_THREADS
#define V3_THREADS 512
#endif

namespace pgmb {

using namespace nrsym;

namespace {

constexpr int kChainAhead = V3_CHAIN_AHEAD; // rows whose operands are pulled into L1 ahead of the chain step that uses them

__device__ __forceinline__ uint32_t smem_u32(void const* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Operand ring of the chain walks: the operands a chain step needs first (10 doubles per thread) are copied global -> shared
// with cp.async one row ahead, into one of two buffers; no register holds them meanwhile and the step reads them with LDS.
// Layout [buffer][operand][thread]: conflict-free.
constexpr int kRingOperands = 10;
struct Ring {
    double* base; // this thread's element of operand 0, buffer 0
    int stride;   // doubles between operands (= threads of the block)
    __device__ __forceinline__ double* at(int buf, int k) const { return base + (size_t)(buf * kRingOperands + k) * stride; }
};
__device__ __forceinline__ void cp_async8(double* dst, double const* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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
    uint32_t done = 0;
    while (done == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar)
                     : "memory");
    }
}

template <int T> struct TileP {
    double* __restrict__ jac;
    double* __restrict__ xvec;
    double* __restrict__ pol;
    double* __restrict__ u;
    uint8_t* __restrict__ perm;
    double* __restrict__ side; // [n_bus][2][T] right-hand-side part of the precomputed leaf update term (kind 2)
    double const* __restrict__ sinj;
    double const* __restrict__ usrc;
    // voltage regulators (REG instantiations only): status of every load_gen, Q limit the bus ran into (0 none, 1 lower, 2 upper)
    uint8_t const* __restrict__ lg_status;
    uint8_t* __restrict__ qviol;
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

// L block of one child: l = (a Q_c) U_c^-1   (sparse_lu_solver.hpp:404-419), identical statements to nr_sym_v2.cu pass 2
__device__ __forceinline__ Blk lower_block(Blk a, Blk const& piv, int q_swap) {
    if (q_swap) {
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
    return l;
}

// diagonal factor + U block towards the parent + forward substitution inside the block; stores everything the later
// phases read.  Returns singular.
template <int T>
__device__ __forceinline__ bool finish_row(TileP<T> const& t, int row, int k_d, int k_u, Blk& d, Blk& ub, double& acc0,
                                           double& acc1, int& pcq) {
    int pr;
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

// ---- build phase -----------------------------------------------------------------------------------------------------
// own blocks of one row: diagonal d, block towards the parent ub, mismatch / right-hand side acc, blocks towards the
// children (stored to their LU slots).  rec = 8-word head; lower = per-child words (null for a leaf).
template <int T, Mode mode, bool REG = false>
__device__ __forceinline__ void build_row(DevStructure const& s, TileP<T> const& t, int32_t const* __restrict__ rec, int n_lower,
                                          int32_t const* __restrict__ lower, Blk& d, Blk& ub, double& acc0, double& acc1,
                                          [[maybe_unused]] bool check_now = false) {
    int const row = rec[0], k_d = rec[1], ky_d = rec[2], k_u = rec[3], j = rec[4], ky_u = rec[5];
    int const lg0 = rec[6] & 0xffffff, n_lg = (rec[6] >> 24) & 0x7f;
    int const sr0 = rec[7] & 0xffffff, n_src = (rec[7] >> 24) & 0x7f;
    double const uir = t.u[(size_t)(row * 2) * T], uii = t.u[(size_t)(row * 2 + 1) * T];
    acc0 = 0.0;
    acc1 = 0.0;
    d = {0.0, 0.0, 0.0, 0.0};
    ub = {0.0, 0.0, 0.0, 0.0};
    // row sums in entry order: lower entries, diagonal, upper entry
    for (int e = 0; e < n_lower; ++e) {
        int const c = lower[4 * e], ky = lower[4 * e + 1];
        Blk a{0.0, 0.0, 0.0, 0.0};
        if (ky >= 0) {
            double const yr = __ldg(s.ydata + 2 * ky), yi = __ldg(s.ydata + 2 * ky + 1);
            if constexpr (mode == Mode::newton) {
                double h, n;
                hnml(yr, yi, uir, uii, t.u[(size_t)(c * 2) * T], t.u[(size_t)(c * 2 + 1) * T], h, n);
                a = {h, -n, n, h};
                acc0 -= n;
                acc1 -= h;
            } else {
                a = {yr, yi, -yi, yr};
            }
        }
        t.store_blk(k_d - n_lower + e, a);
    }
    {
        double const yr = __ldg(s.ydata + 2 * ky_d), yi = __ldg(s.ydata + 2 * ky_d + 1);
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
    if (k_u >= 0 && ky_u >= 0) {
        double const yr = __ldg(s.ydata + 2 * ky_u), yi = __ldg(s.ydata + 2 * ky_u + 1);
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
    if constexpr (mode == Mode::newton) {
        d.a00 += acc1;
        d.a01 += -acc0;
        d.a10 += -acc0;
        d.a11 += -acc1;
    }
    double const v = t.pol[(size_t)(row * 2 + 1) * T];
    // REG: loads and sources depend on the Q limit the bus may run into right now; the part above does not, so a bus that
    // hits its limit repeats only what follows (the block kernel rebuilds the whole row: the same sums in the same order)
    [[maybe_unused]] PvControl ctl{false, false, 0.0, 0.0, 0.0};
    [[maybe_unused]] int viol = 0;
    [[maybe_unused]] Blk d_rows = d;
    [[maybe_unused]] double rows0 = acc0, rows1 = acc1;
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
                continue; // the bus is PQ from now on: its injections again, with the clamped generators
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
            ub.a10 = 0.0;
            ub.a11 = 0.0;
            for (int e = 0; e < n_lower; ++e) {
                double* p = t.jac + (size_t)(k_d - n_lower + e) * 4 * T;
                p[T] = 0.0;
                p[3 * T] = 0.0;
            }
        }
    }
}

template <int T, Mode mode, bool REG = false>
__device__ __forceinline__ bool build_leaf(DevStructure const& s, TileP<T> const& t, int32_t const* __restrict__ rec, bool check_now = false) {
    Blk d, ub;
    double acc0, acc1;
    build_row<T, mode, REG>(s, t, rec, 0, nullptr, d, ub, acc0, acc1, check_now);
    int pcq;
    return finish_row<T>(t, rec[0], rec[1], rec[3], d, ub, acc0, acc1, pcq);
}

// non-leaf row: build, then eliminate / precompute the leaf children (they were finished by build_leaf before the barrier)
template <int T, Mode mode, bool REG = false>
__device__ __forceinline__ void build_inner(DevStructure const& s, TileP<T> const& t, int32_t const* __restrict__ rec, bool check_now = false) {
    int const row = rec[0], k_d = rec[1], k_u = rec[3];
    int const n_lower = rec[8] & 0xfff;
    int32_t const* __restrict__ lower = rec + 9;
    Blk d, ub;
    double acc0, acc1;
    build_row<T, mode, REG>(s, t, rec, n_lower, lower, d, ub, acc0, acc1, check_now);
    for (int e = 0; e < n_lower; ++e) {
        int const kind = (lower[4 * e + 3] >> 28) & 3;
        if (kind != 0 && kind != 2) continue;
        int const c = lower[4 * e], kd_c = lower[4 * e + 2], k_uc = lower[4 * e + 3] & 0x0fffffff;
        int const k_e = k_d - n_lower + e;
        Blk const a = t.load_blk(k_e); // written above by this thread
        Blk const piv = t.load_blk(kd_c);
        Blk const uc = t.load_blk(k_uc);
        uint8_t const pc = t.perm[(size_t)c * T];
        double const y0 = t.xvec[(size_t)(c * 2) * T], y1 = t.xvec[(size_t)(c * 2 + 1) * T];
        Blk const l = lower_block(a, piv, pc & 2);
        Blk sterm;
        sterm.a00 = l.a00 * uc.a00 + l.a01 * uc.a10;
        sterm.a10 = l.a10 * uc.a00 + l.a11 * uc.a10;
        sterm.a01 = l.a00 * uc.a01 + l.a01 * uc.a11;
        sterm.a11 = l.a10 * uc.a01 + l.a11 * uc.a11;
        double const s0 = l.a00 * y0 + l.a01 * y1;
        double const s1 = l.a10 * y0 + l.a11 * y1;
        if (kind == 0) {
            d.a00 -= sterm.a00;
            d.a10 -= sterm.a10;
            d.a01 -= sterm.a01;
            d.a11 -= sterm.a11;
            acc0 -= s0;
            acc1 -= s1;
        } else {
            t.store_blk(k_e, sterm);
            t.side[(size_t)(row * 2) * T] = s0;
            t.side[(size_t)(row * 2 + 1) * T] = s1;
        }
    }
    t.store_blk(k_d, d);
    if (k_u >= 0) t.store_blk(k_u, ub);
    t.xvec[(size_t)(row * 2) * T] = acc0;
    t.xvec[(size_t)(row * 2 + 1) * T] = acc1;
}

// ---- chain phase ---------------------------------------------------------------------------------------------------------
// One chain step needs the block towards the carried child at once (first divide of the chain) and the prebuilt diagonal
// right after it: those two are fetched one row ahead into registers; all operands of the row after next are pulled into L1
// (prefetch.global.L1).  Everything else of the row (mismatch, block towards the parent, precomputed leaf term) is requested
// when the step starts and arrives behind the divide chain.  Chain records (8 words) are read as two 16-byte vectors.
template <int T> __device__ __forceinline__ void prefetch_blk(double const* jac, int k) {
    double const* p = jac + (size_t)k * 4 * T;
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + i * T));
}
// RING: the diagonal, the block towards the carried child and the mismatch come through the operand ring instead
template <int T, bool RING> __device__ __forceinline__ void prefetch_row(TileP<T> const& t, int4 const c0, int4 const c1) {
    if constexpr (!RING) prefetch_blk<T>(t.jac, c0.y);
    if (c0.z >= 0) prefetch_blk<T>(t.jac, c0.z);
    if constexpr (!RING) {
        if (c0.w >= 0) prefetch_blk<T>(t.jac, c0.w);
    }
    if (c1.x >= 0) {
        prefetch_blk<T>(t.jac, c1.x);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.side + (size_t)(c0.x * 2) * T));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.side + (size_t)(c0.x * 2 + 1) * T));
    }
    if constexpr (!RING) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.xvec + (size_t)(c0.x * 2) * T));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.xvec + (size_t)(c0.x * 2 + 1) * T));
    }
}

__device__ __forceinline__ void eliminate(Blk& d, double& acc0, double& acc1, Blk const& a, Blk const& piv, Blk const& uc,
                                          double y0, double y1, int q) {
    Blk const l = lower_block(a, piv, q);
    d.a00 -= l.a00 * uc.a00 + l.a01 * uc.a10;
    d.a10 -= l.a10 * uc.a00 + l.a11 * uc.a10;
    d.a01 -= l.a00 * uc.a01 + l.a01 * uc.a11;
    d.a11 -= l.a10 * uc.a01 + l.a11 * uc.a11;
    acc0 -= l.a00 * y0 + l.a01 * y1;
    acc1 -= l.a10 * y0 + l.a11 * y1;
}

template <int T, bool RING>
__device__ __forceinline__ bool up_path(TileP<T> const& t, int32_t const* __restrict__ prog, int4 const* __restrict__ chain,
                                        int first_rec, int n_rows, Ring const& ring) {
    bool singular = false;
    Blk c_piv{1.0, 0.0, 0.0, 1.0}, c_uc{0.0, 0.0, 0.0, 0.0};
    double c_y0 = 0.0, c_y1 = 0.0;
    int c_q = 0;
    int4 n0 = chain[2 * first_rec], n1 = chain[2 * first_rec + 1];
    // RING: block towards the carried child, prebuilt diagonal and mismatch of the next row travel through the ring
    auto stage = [&](int buf, int4 const c) {
        if (c.w >= 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async8(ring.at(buf, i), t.jac + (size_t)c.w * 4 * T + i * T);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async8(ring.at(buf, 4 + i), t.jac + (size_t)c.y * 4 * T + i * T);
        cp_async8(ring.at(buf, 8), t.xvec + (size_t)(c.x * 2) * T);
        cp_async8(ring.at(buf, 9), t.xvec + (size_t)(c.x * 2 + 1) * T);
        cp_async_commit();
    };
    Blk a_next{0.0, 0.0, 0.0, 0.0}, d_next{0.0, 0.0, 0.0, 0.0};
    int buf = 0;
    if constexpr (RING) {
        stage(0, n0);
    } else {
        if (n0.w >= 0) a_next = t.load_blk(n0.w);
        d_next = t.load_blk(n0.y);
    }
    for (int k = 1; k < kChainAhead && k < n_rows; ++k) prefetch_row<T, RING>(t, chain[2 * (first_rec + k)], chain[2 * (first_rec + k) + 1]);
    for (int i = 0; i < n_rows; ++i) {
        int4 const c0 = n0, c1 = n1;
        int const row = c0.x, k_d = c0.y, k_u = c0.z, k_s = c1.x, pattern = c1.y;
        Blk a_carry, d;
        double acc0, acc1;
        if constexpr (RING) {
            cp_async_wait_all();
            a_carry = c0.w >= 0 ? Blk{*ring.at(buf, 0), *ring.at(buf, 1), *ring.at(buf, 2), *ring.at(buf, 3)} : Blk{0.0, 0.0, 0.0, 0.0};
            d = Blk{*ring.at(buf, 4), *ring.at(buf, 5), *ring.at(buf, 6), *ring.at(buf, 7)};
            acc0 = *ring.at(buf, 8);
            acc1 = *ring.at(buf, 9);
        } else {
            a_carry = a_next;
            d = d_next;
            acc0 = t.xvec[(size_t)(row * 2) * T];
            acc1 = t.xvec[(size_t)(row * 2 + 1) * T];
        }
        // operands of this row that are needed behind the first divide chain
        Blk ub = k_u >= 0 ? t.load_blk(k_u) : Blk{0.0, 0.0, 0.0, 0.0};
        Blk sterm{0.0, 0.0, 0.0, 0.0};
        double s0 = 0.0, s1 = 0.0;
        if (k_s >= 0) {
            sterm = t.load_blk(k_s);
            s0 = t.side[(size_t)(row * 2) * T];
            s1 = t.side[(size_t)(row * 2 + 1) * T];
        }
        if (i + 1 < n_rows) {
            n0 = chain[2 * (first_rec + i + 1)];
            n1 = chain[2 * (first_rec + i + 1) + 1];
            if constexpr (RING) {
                stage(buf ^ 1, n0);
            } else {
                a_next = n0.w >= 0 ? t.load_blk(n0.w) : Blk{0.0, 0.0, 0.0, 0.0};
                d_next = t.load_blk(n0.y);
            }
            if (i + kChainAhead < n_rows) prefetch_row<T, RING>(t, chain[2 * (first_rec + i + kChainAhead)], chain[2 * (first_rec + i + kChainAhead) + 1]);
        }
        buf ^= 1;
        if (pattern >= 2) { // carry child, then (pattern 3) the precomputed leaf term
            eliminate(d, acc0, acc1, a_carry, c_piv, c_uc, c_y0, c_y1, c_q);
            if (pattern == 3) {
                d.a00 -= sterm.a00;
                d.a10 -= sterm.a10;
                d.a01 -= sterm.a01;
                d.a11 -= sterm.a11;
                acc0 -= s0;
                acc1 -= s1;
            }
        } else if (pattern == 0) { // generic: children in entry order from the row record
            int32_t const* __restrict__ cur = prog + c1.z;
            int const n_lower = cur[8] & 0xfff;
            int32_t const* __restrict__ lower = cur + 9;
            for (int e = 0; e < n_lower; ++e) {
                int const w = lower[4 * e + 3];
                int const kind = (w >> 28) & 3;
                if (kind == 0) continue;
                if (kind == 2) {
                    d.a00 -= sterm.a00;
                    d.a10 -= sterm.a10;
                    d.a01 -= sterm.a01;
                    d.a11 -= sterm.a11;
                    acc0 -= s0;
                    acc1 -= s1;
                } else if (kind == 1) {
                    eliminate(d, acc0, acc1, a_carry, c_piv, c_uc, c_y0, c_y1, c_q);
                } else {
                    int const c = lower[4 * e];
                    Blk const a = t.load_blk(k_d - n_lower + e);
                    Blk const piv = t.load_blk(lower[4 * e + 2]);
                    Blk const uc = t.load_blk(w & 0x0fffffff);
                    int const q = t.perm[(size_t)c * T] & 2;
                    double const y0 = t.xvec[(size_t)(c * 2) * T], y1 = t.xvec[(size_t)(c * 2 + 1) * T];
                    eliminate(d, acc0, acc1, a, piv, uc, y0, y1, q);
                }
            }
        }
        int pcq;
        singular |= finish_row<T>(t, row, k_d, k_u, d, ub, acc0, acc1, pcq);
        c_piv = d;
        c_uc = ub;
        c_y0 = acc0;
        c_y1 = acc1;
        c_q = pcq;
    }
    return singular;
}

struct DownOperands {
    Blk d, ub;
    double y0, y1, th, v, our, oui;
    int pm;
};
template <int T, Mode mode>
__device__ __forceinline__ DownOperands fetch_down_k(TileP<T> const& t, int row, int k_d, int k_u) {
    DownOperands o;
    o.y0 = t.xvec[(size_t)(row * 2) * T];
    o.y1 = t.xvec[(size_t)(row * 2 + 1) * T];
    o.d = t.load_blk(k_d);
    o.pm = t.perm[(size_t)row * T];
    o.ub = k_u >= 0 ? t.load_blk(k_u) : Blk{0.0, 0.0, 0.0, 0.0};
    o.th = o.v = o.our = o.oui = 0.0;
    if constexpr (mode == Mode::newton) {
        o.th = t.pol[(size_t)(row * 2) * T];
        o.v = t.pol[(size_t)(row * 2 + 1) * T];
        o.our = t.u[(size_t)(row * 2) * T];
        o.oui = t.u[(size_t)(row * 2 + 1) * T];
    }
    return o;
}
template <int T, Mode mode>
__device__ __forceinline__ DownOperands fetch_down(TileP<T> const& t, int32_t const* __restrict__ rec) {
    return fetch_down_k<T, mode>(t, rec[0], rec[1], rec[3]);
}
template <int T, Mode mode, bool RING = false> __device__ __forceinline__ void prefetch_down(TileP<T> const& t, int4 const c0) {
    if constexpr (!RING) {
        prefetch_blk<T>(t.jac, c0.y);
        if (c0.z >= 0) prefetch_blk<T>(t.jac, c0.z);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.xvec + (size_t)(c0.x * 2) * T));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.xvec + (size_t)(c0.x * 2 + 1) * T));
    }
    asm volatile("prefetch.global.L1 [%0];" ::"l"(t.perm + (size_t)c0.x * T));
    if constexpr (mode == Mode::newton) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.pol + (size_t)(c0.x * 2) * T));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.pol + (size_t)(c0.x * 2 + 1) * T));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.u + (size_t)(c0.x * 2) * T));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.u + (size_t)(c0.x * 2 + 1) * T));
    }
}
// backward substitution of one row given the parent's solution; returns the voltage change
template <int T, Mode mode, bool REG = false>
__device__ __forceinline__ double down_step(TileP<T> const& t, int row, bool has_parent, DownOperands const& o, double& x0,
                                            double& x1, [[maybe_unused]] DevStructure const* s = nullptr) {
    double y0 = o.y0, y1 = o.y1;
    if (has_parent) {
        y0 -= o.ub.a00 * x0 + o.ub.a01 * x1;
        y1 -= o.ub.a10 * x0 + o.ub.a11 * x1;
    }
    y1 /= o.d.a11;
    y0 -= o.d.a01 * y1;
    y0 /= o.d.a00;
    if (o.pm & 2) {
        double const x = y0;
        y0 = y1;
        y1 = x;
    }
    t.xvec[(size_t)(row * 2) * T] = y0;
    t.xvec[(size_t)(row * 2 + 1) * T] = y1;
    x0 = y0;
    x1 = y1;
    if constexpr (REG && mode == Mode::linear_init) pv_start_voltage(pv_control_of_row<T>(*s, t, row), y0, y1);
    return polar_update<T, mode>(t.pol + (size_t)(row * 2) * T, t.u + (size_t)(row * 2) * T, y0, y1, o.th, o.v, o.our, o.oui);
}

template <int T, Mode mode, bool RING, bool REG = false>
__device__ __forceinline__ double down_path(TileP<T> const& t, int4 const* __restrict__ chain, int first_rec, int n_rows,
                                            Ring const& ring, [[maybe_unused]] DevStructure const* s = nullptr) {
    double dev = 0.0;
    int4 c0 = chain[2 * (first_rec + n_rows - 1)];
    int const j_top = chain[2 * (first_rec + n_rows - 1) + 1].w;
    // RING: U block, factorised diagonal and right-hand side of the next row (what the step needs first) travel through the
    // ring; permutation and old voltage (needed behind the divides) are plain loads when the step starts
    auto stage = [&](int buf, int4 const c) {
        if (c.z >= 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cp_async8(ring.at(buf, i), t.jac + (size_t)c.z * 4 * T + i * T);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async8(ring.at(buf, 4 + i), t.jac + (size_t)c.y * 4 * T + i * T);
        cp_async8(ring.at(buf, 8), t.xvec + (size_t)(c.x * 2) * T);
        cp_async8(ring.at(buf, 9), t.xvec + (size_t)(c.x * 2 + 1) * T);
        cp_async_commit();
    };
    DownOperands next;
    int buf = 0;
    auto fetch_late = [&](int row) { // needed behind the divides: one row ahead into registers
        next.pm = t.perm[(size_t)row * T];
        next.th = next.v = next.our = next.oui = 0.0;
        if constexpr (mode == Mode::newton) {
            next.th = t.pol[(size_t)(row * 2) * T];
            next.v = t.pol[(size_t)(row * 2 + 1) * T];
            next.our = t.u[(size_t)(row * 2) * T];
            next.oui = t.u[(size_t)(row * 2 + 1) * T];
        }
    };
    if constexpr (RING) {
        stage(0, c0);
        fetch_late(c0.x);
    } else {
        next = fetch_down_k<T, mode>(t, c0.x, c0.y, c0.z);
    }
    double x0 = 0.0, x1 = 0.0;
    bool has_parent = c0.z >= 0;
    if (has_parent) {
        x0 = t.xvec[(size_t)(j_top * 2) * T];
        x1 = t.xvec[(size_t)(j_top * 2 + 1) * T];
    }
    for (int k = 2; k <= kChainAhead && k <= n_rows; ++k) prefetch_down<T, mode, RING>(t, chain[2 * (first_rec + n_rows - k)]);
    for (int i = n_rows - 1; i >= 0; --i) {
        DownOperands o;
        int const row = c0.x;
        if constexpr (RING) {
            cp_async_wait_all();
            o.ub = c0.z >= 0 ? Blk{*ring.at(buf, 0), *ring.at(buf, 1), *ring.at(buf, 2), *ring.at(buf, 3)} : Blk{0.0, 0.0, 0.0, 0.0};
            o.d = Blk{*ring.at(buf, 4), *ring.at(buf, 5), *ring.at(buf, 6), *ring.at(buf, 7)};
            o.y0 = *ring.at(buf, 8);
            o.y1 = *ring.at(buf, 9);
            o.pm = next.pm;
            o.th = next.th;
            o.v = next.v;
            o.our = next.our;
            o.oui = next.oui;
        } else {
            o = next;
        }
        if (i > 0) {
            c0 = chain[2 * (first_rec + i - 1)];
            if constexpr (RING) {
                stage(buf ^ 1, c0);
                fetch_late(c0.x);
            } else {
                next = fetch_down_k<T, mode>(t, c0.x, c0.y, c0.z);
            }
            if (i >= kChainAhead) prefetch_down<T, mode, RING>(t, chain[2 * (first_rec + i - kChainAhead)]);
        }
        buf ^= 1;
        dev = fmax(dev, down_step<T, mode, REG>(t, row, has_parent, o, x0, x1, s));
        has_parent = true;
    }
    return dev;
}

template <int T, Mode mode, bool RING, bool REG>
__device__ __forceinline__ void sweeps_v3(DevStructure const& s, TileP<T> const& t, int32_t const* __restrict__ prog, int slot,
                                          int n_slot, bool active, bool& singular, double& dev, unsigned long long* phase,
                                          Ring const& ring, bool check_now) {
    int32_t const* __restrict__ const recs = s.path_prog; // leaf / row records: global memory
    int const n_leaf = prog[0], n_rec = prog[1], n_stage = prog[2];
    int32_t const* __restrict__ leaf = recs + prog[3];
    int32_t const* __restrict__ rec_off = recs + prog[4];
    int32_t const* __restrict__ stage_ptr = prog + prog[5];
    int32_t const* __restrict__ path = prog + prog[6];
    int4 const* __restrict__ chain = reinterpret_cast<int4 const*>(prog + prog[8]);
    long long t0 = clock64();
    auto lap = [&](int k) {
        if (phase != nullptr && threadIdx.x == 0) {
            long long const t1 = clock64();
            phase[k] += (unsigned long long)(t1 - t0);
            t0 = t1;
        }
    };
    // the voltages / loads of the row this thread builds next are pulled into L1 while the current row is computed
    auto prefetch_row_inputs = [&](int32_t const* rec, int n_lower, int32_t const* lower) {
        int const row = rec[0];
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.u + (size_t)(row * 2) * T));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.u + (size_t)(row * 2 + 1) * T));
        if constexpr (mode == Mode::newton) asm volatile("prefetch.global.L1 [%0];" ::"l"(t.pol + (size_t)(row * 2 + 1) * T));
        if (rec[4] >= 0) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(t.u + (size_t)(rec[4] * 2) * T));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(t.u + (size_t)(rec[4] * 2 + 1) * T));
        }
        int const lg0 = rec[6] & 0xffffff, n_lg = (rec[6] >> 24) & 0x7f;
        for (int lg = lg0; lg < lg0 + n_lg; ++lg) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(t.sinj + (size_t)(lg * 2) * T));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(t.sinj + (size_t)(lg * 2 + 1) * T));
        }
        for (int e = 0; e < n_lower; ++e) {
            int const c = lower[4 * e];
            asm volatile("prefetch.global.L1 [%0];" ::"l"(t.u + (size_t)(c * 2) * T));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(t.u + (size_t)(c * 2 + 1) * T));
            int const kind = (lower[4 * e + 3] >> 28) & 3;
            if (kind == 0 || kind == 2) { // leaf child: its factor, U block, permutation and rhs are read when the row is built
                prefetch_blk<T>(t.jac, lower[4 * e + 2]);
                prefetch_blk<T>(t.jac, lower[4 * e + 3] & 0x0fffffff);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(t.perm + (size_t)c * T));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(t.xvec + (size_t)(c * 2) * T));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(t.xvec + (size_t)(c * 2 + 1) * T));
            }
        }
    };
    if (active) {
        for (int i = slot; i < n_leaf; i += n_slot) {
            if (i + n_slot < n_leaf) prefetch_row_inputs(leaf + 8 * (i + n_slot), 0, nullptr);
            singular |= build_leaf<T, mode, REG>(s, t, leaf + 8 * i, check_now);
        }
    }
    __syncthreads();
    lap(0);
    if (active) {
        for (int i = slot; i < n_rec; i += n_slot) {
            if (i + n_slot < n_rec) {
                int32_t const* nx = recs + rec_off[i + n_slot];
                prefetch_row_inputs(nx, nx[8] & 0xfff, nx + 9);
            }
            build_inner<T, mode, REG>(s, t, recs + rec_off[i], check_now);
        }
    }
    __syncthreads();
    lap(1);
    for (int st = 1; st < n_stage; ++st) {
        if (active) {
            for (int p = stage_ptr[st - 1] + slot; p < stage_ptr[st]; p += n_slot) {
                singular |= up_path<T, RING>(t, recs, chain, path[2 * p], path[2 * p + 1], ring);
            }
        }
        __syncthreads();
    }
    lap(2);
    for (int st = n_stage - 1; st >= 1; --st) {
        if (active) {
            for (int p = stage_ptr[st - 1] + slot; p < stage_ptr[st]; p += n_slot) {
                dev = fmax(dev, down_path<T, mode, RING && (V3_RING_DOWN != 0), REG>(t, chain, path[2 * p], path[2 * p + 1], ring, &s));
            }
        }
        __syncthreads();
    }
    lap(3);
    if (active && slot < n_leaf) {
        // software pipeline in registers: the operands of the thread's next leaf are loaded while the current one is solved
        // (prefetch.global.L1 only reaches L2 on this part, tools/microbench_prefetch.cu); the leaf after that is prefetched
        auto fetch_leaf = [&](int32_t const* rec, DownOperands& o, double& x0, double& x1) {
            o = fetch_down<T, mode>(t, rec);
            x0 = x1 = 0.0;
            if (rec[3] >= 0) {
                int const j = rec[4];
                x0 = t.xvec[(size_t)(j * 2) * T];
                x1 = t.xvec[(size_t)(j * 2 + 1) * T];
            }
        };
        DownOperands next;
        double nx0, nx1;
        fetch_leaf(leaf + 8 * slot, next, nx0, nx1);
        for (int i = slot; i < n_leaf; i += n_slot) {
            int32_t const* rec = leaf + 8 * i;
            DownOperands const o = next;
            double x0 = nx0, x1 = nx1;
            if (i + n_slot < n_leaf) {
                fetch_leaf(leaf + 8 * (i + n_slot), next, nx0, nx1);
                if (i + 2 * n_slot < n_leaf) {
                    int32_t const* nx = leaf + 8 * (i + 2 * n_slot);
                    prefetch_down<T, mode>(t, make_int4(nx[0], nx[1], nx[3], 0));
                    if (nx[3] >= 0) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.xvec + (size_t)(nx[4] * 2) * T));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(t.xvec + (size_t)(nx[4] * 2 + 1) * T));
                    }
                }
            }
            dev = fmax(dev, down_step<T, mode, REG>(t, rec[0], rec[3] >= 0, o, x0, x1, &s));
        }
    }
    __syncthreads();
    lap(4);
}

} // namespace

// SMEM: the chain part of the path program is staged in shared memory; RING (needs SMEM): operand ring of the chain walks
// behind it
// REG: the grid has voltage regulators (PV buses with reactive-power limits)
template <int T, bool SMEM, bool RING, bool REG> __global__ void __launch_bounds__(V3_THREADS, 1) nr_sym_v3_kernel(DevStructure s, DevBatch b, SolveOptions opt) {
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

    if constexpr (SMEM) stage_program(reinterpret_cast<int32_t*>(smem_raw), s.path_prog, (uint32_t)s.path_prog_smem_words * 4u, &sh_mbar);
    // SMEM: the pointer is derived from the shared array in this scope, so every read of the staged prefix (header, stages,
    // paths, chain records) compiles to LDS; leaf / row records stay in global memory (s.path_prog, read through L1)
    int32_t const* const prog = SMEM ? reinterpret_cast<int32_t const*>(smem_raw) : s.path_prog;
    Ring ring{nullptr, 0};
    if constexpr (RING) {
        ring.base = reinterpret_cast<double*>(smem_raw + (((size_t)s.path_prog_smem_words * 4 + 127) / 128) * 128) + threadIdx.x;
        ring.stride = blockDim.x;
    }
    TileP<T> t;
    t.jac = b.jac + (size_t)tile * s.nnz_lu * 4 * T + lane;
    t.xvec = b.xvec + (size_t)tile * s.n_bus * 2 * T + lane;
    t.pol = b.pol + (size_t)tile * s.n_bus * 2 * T + lane;
    t.u = b.u + (size_t)tile * s.n_bus * 2 * T + lane;
    t.perm = b.perm + (size_t)tile * s.n_bus * T + lane;
    t.side = b.side + (size_t)tile * s.n_bus * 2 * T + lane;
    t.sinj = b.sinj + (size_t)tile * s.n_load_gen * 2 * T + lane;
    t.usrc = b.usrc + (size_t)tile * s.n_source * 2 * T + lane;
    t.lg_status = (REG && b.lg_status != nullptr) ? b.lg_status + (size_t)tile * s.n_load_gen * T + lane : nullptr;
    t.qviol = (REG && b.qviol != nullptr) ? b.qviol + (size_t)tile * s.n_bus * T + lane : nullptr;

    if (threadIdx.x < T) {
        sh_dev[threadIdx.x] = 0ull;
        sh_singular[threadIdx.x] = 0;
        sh_has_limits[threadIdx.x] = 0;
    }
    __syncthreads();
    if constexpr (REG) { // set_bus_types_and_q_limits (newton_raphson_pf_solver.hpp:400-444); no limit has been hit yet
        if (valid) {
            for (int row = slot; row < s.n_bus; row += n_slot) {
                t.qviol[(size_t)row * T] = 0;
                if (__ldg(s.lg_ptr + row) != __ldg(s.lg_ptr + row + 1) && pv_control_of_row<T>(s, t, row).has_limits) sh_has_limits[lane] = 1;
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
        sweeps_v3<T, Mode::linear_init, RING, REG>(s, t, prog, slot, n_slot, !done, singular, dev, phase, ring, false);
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
        sweeps_v3<T, Mode::newton, RING, REG>(s, t, prog, slot, n_slot, !done, singular, dev, phase ? phase + 8 : nullptr, ring, REG && num_iter >= 2);
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

template <int T, bool REG>
static void launch_v3_t(DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot, cudaStream_t st) {
    size_t const prog_bytes = (((size_t)s.path_prog_smem_words * 4 + 127) / 128) * 128;
    size_t const ring_bytes = (size_t)2 * kRingOperands * sizeof(double) * T * n_slot;
    int dev = 0, max_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    static int const want_ring = [] {
        char const* v = std::getenv("PGMB_V3_RING"); // 0: chain operands through registers / L1 prefetch only (comparison)
        return v != nullptr && *v != '\0' ? std::atoi(v) : 1;
    }();
    // The attribute is one value per kernel and device, shared by every host thread that launches it (the scenario-by-scenario
    // route runs engines of different grids at once): always the device limit, never this launch's own size.
    bool const in_smem = prog_bytes + 1024 <= (size_t)max_optin;
    bool const with_ring = in_smem && want_ring != 0 && prog_bytes + ring_bytes + 1024 <= (size_t)max_optin;
    if (with_ring) {
        size_t const dyn = prog_bytes + ring_bytes;
        cudaFuncSetAttribute(nr_sym_v3_kernel<T, true, true, REG>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - 1024);
        nr_sym_v3_kernel<T, true, true, REG><<<b.n_tile, T * n_slot, dyn, st>>>(s, b, opt);
    } else if (in_smem) {
        cudaFuncSetAttribute(nr_sym_v3_kernel<T, true, false, REG>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - 1024);
        nr_sym_v3_kernel<T, true, false, REG><<<b.n_tile, T * n_slot, prog_bytes, st>>>(s, b, opt);
    } else {
        nr_sym_v3_kernel<T, false, false, REG><<<b.n_tile, T * n_slot, 0, st>>>(s, b, opt);
    }
}

void launch_nr_sym_v3(int tile_width, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot,
                      cudaStream_t st) {
    count_kernel_launch();
    if (n_slot * tile_width > V3_THREADS) n_slot = V3_THREADS / tile_width;
    bool const reg = s.lg_reg != nullptr;
    if (reg && (b.qviol == nullptr || b.lg_status == nullptr)) {
        throw std::logic_error("nr_sym_v3: a grid with voltage regulators needs the qviol / lg_status buffers of the batch");
    }
    switch (tile_width) {
    case 4: reg ? launch_v3_t<4, true>(s, b, opt, n_slot, st) : launch_v3_t<4, false>(s, b, opt, n_slot, st); break;
    case 8: reg ? launch_v3_t<8, true>(s, b, opt, n_slot, st) : launch_v3_t<8, false>(s, b, opt, n_slot, st); break;
    case 16: reg ? launch_v3_t<16, true>(s, b, opt, n_slot, st) : launch_v3_t<16, false>(s, b, opt, n_slot, st); break;
    default: reg ? launch_v3_t<32, true>(s, b, opt, n_slot, st) : launch_v3_t<32, false>(s, b, opt, n_slot, st); break;
    }
}

} // namespace pgmb
