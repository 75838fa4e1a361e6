// Asymmetric (three-phase) Newton-Raphson with the row task of a bus SPLIT OVER SIX THREADS: thread r owns block row r of every
// 6 x 6 block of the bus row (and element r of its right-hand side).  nr_block.cu gives one thread the whole row task -- ~5000
// dependent FP64 operations, ~130 k cycles -- and the elimination tree of a meshed grid has 60 levels with ~20 rows each, so
// that kernel walks a critical path of row latencies with most of the SM idle.  Here the L row, the Schur-update row and the
// right-hand-side element are thread-local, the full-pivot search / row swaps / pivot-row broadcasts of the 6 x 6 factorisation
// are warp shuffles, and the U blocks are finished column-wise.  Per matrix element the floating-point operations and their
// order are exactly those of nr_block.cu (and of the reference), so the results are bit-identical.
//   newton_raphson_pf_solver.hpp:255-303, 462-547, 764-852, 325-349 ; sparse_lu_solver.hpp:86-165, 171-200, 346-495, 769-827
// Warp layout: lane = r * 4 + sc, r = 0..7 (rows 6, 7 idle: they mirror rows 0, 1 without storing), sc = scenario within a group
// of four; a tile of T scenarios is T / 4 such groups; one warp = one row task for one group.  Hub rows (WideRowPlan) run the
// cooperative phases of block_common.cuh's wide_up_row with the same row split (wide_up_row6): one warp per entry / child, so
// no thread of this kernel ever holds a whole 6 x 6 block and the register budget allows 24 - 32 warps per SM.
#include "kernels.cuh"
#include "block_common.cuh"

#include <algorithm>
#include <type_traits>

namespace pgmb {
using namespace blk;
namespace {

#ifndef B6_THREADS
#define B6_THREADS 768
#endif
constexpr int kB = 3, kN = 6, kNN = 36, kBB2 = 18;
constexpr unsigned kFull = 0xffffffffu;

template <int T> struct Tile6 {
    double* jac;
    double* xvec;
    double* pol;
    double* u;
    uint8_t* perm;
    double const* sinj;
    double const* usrc;
    int32_t const* ovr_entry;
    double const* ovr_y;
    int ovr_n;
    uint8_t const* dead;
    double* sm;         // shared scratch of my warp + my scenario: element e at sm[e * 4]; 36 block elements, then 6 vector slots
    double* wide_terms; // scratch of the cooperative hub rows (block_common.cuh: wide_up_row), may be null
    double* wide_rhs;
    double* wide_sum;
    int r;    // block row of this thread (0..5; the idle lanes carry 0 / 1)
    int sc;   // scenario within the group of four = lane & 3
    bool real; // lane carries a real block row (not one of the two idle mirrors)
    bool act; // this thread stores results (real row, valid and unfinished scenario)
};

__device__ __forceinline__ double sel6(double const* d, int c) {
    double x = d[0];
#pragma unroll
    for (int i = 1; i < kN; ++i) x = (c == i) ? d[i] : x;
    return x;
}
__device__ __forceinline__ int nib(uint32_t packed, int i) { return (packed >> (4 * i)) & 15; }
// a permutation index read from memory; lanes of finished / padding scenarios may read anything: keep their addresses in range
__device__ __forceinline__ int perm_at(uint8_t const* p) { return min((int)*p, kN - 1); }
__device__ __forceinline__ uint32_t nib_swap(uint32_t packed, int i, int j) {
    uint32_t const a = nib(packed, i), b = nib(packed, j);
    packed &= ~((15u << (4 * i)) | (15u << (4 * j)));
    return packed | (b << (4 * i)) | (a << (4 * j));
}
constexpr uint32_t kIdentityPerm = 0x543210u;

// power-flow term of hnml for one element: h = imag, n = real of (ui * conj(uj)) * conj(y)
__device__ __forceinline__ void pf_term(double yr, double yi, double uir, double uii, double ujr, double uji, double& h, double& n) {
    double const cr = ujr, ci = -uji;
    double const ar = uir * cr - uii * ci;
    double const ai = uir * ci + uii * cr;
    double const dr = yr, di = -yi;
    n = ar * dr - ai * di;
    h = ar * di + ai * dr;
}

// ---- full-pivot LU of the 6 x 6 diagonal block (DenseLUFactor::factorize_block_in_place) through the warp's shared scratch -----
// The block sits in shared memory (element (r, c) of my scenario at sm[(c * 6 + r) * 4]).  Every thread scans the whole trailing
// sub-block for the pivot -- first maximum in column-major order, exactly the reference's maxCoeff -- so the search needs no
// cross-thread reduction and all its addresses are compile-time constants; the row swap is done by the six threads column-wise,
// the column swap row-wise, and thread r eliminates row r.  Far fewer instructions than a shuffle version: no select chains, no
// double-width shuffles, no collective synchronisation inside the search.  On return d[] holds row r of the factor and the
// scratch holds the whole factor (read by finish_row6 for the U blocks and the forward substitution).
__device__ bool factorize6s(double* sm, double* d, int r, bool real_row, uint32_t& p_out, uint32_t& q_out) {
    double* const mine = sm + r * 4; // row r: element (r, c) at mine[c * 24]
    if (real_row) {
#pragma unroll
        for (int c = 0; c < kN; ++c) mine[c * 24] = d[c];
    }
    __syncwarp();
    int rt[kN], ct[kN];
    double max_pivot = 0.0;
    bool stopped = false;
#pragma unroll
    for (int pivot = 0; pivot < kN; ++pivot) {
        int best_at = pivot * kN + pivot; // c * 6 + r of the best element
        double best;
        {
            double const v = sm[(pivot * kN + pivot) * 4];
            best = v * v;
        }
#pragma unroll
        for (int c = pivot; c < kN; ++c)
#pragma unroll
            for (int rr = pivot; rr < kN; ++rr) {
                double const v = sm[(c * kN + rr) * 4];
                double const sq = v * v;
                if (sq > best) {
                    best = sq;
                    best_at = c * kN + rr;
                }
            }
        int rb = best_at % kN, cb = best_at / kN;
        if (stopped || best == 0.0) { // the reference stops here: the remaining transpositions are identities
            stopped = true;
            rb = pivot;
            cb = pivot;
        } else {
            max_pivot = fmax(max_pivot, sqrt(best));
        }
        rt[pivot] = rb;
        ct[pivot] = cb;
        __syncwarp(); // every thread has finished its scan before the block changes
        if (real_row) { // row swap pivot <-> rb: thread r moves column r
            double const a = sm[(r * kN + pivot) * 4], b = sm[(r * kN + rb) * 4];
            sm[(r * kN + pivot) * 4] = b;
            sm[(r * kN + rb) * 4] = a;
        }
        __syncwarp();
        if (real_row) { // column swap pivot <-> cb: thread r moves row r
            double const a = mine[pivot * 24], b = mine[cb * 24];
            mine[pivot * 24] = b;
            mine[cb * 24] = a;
        }
        __syncwarp();
        if (pivot < kN - 1) {
            if (!stopped && real_row && r > pivot) {
                double const m = mine[pivot * 24] / sm[(pivot * kN + pivot) * 4];
                mine[pivot * 24] = m;
#pragma unroll
                for (int c = pivot + 1; c < kN; ++c) mine[c * 24] -= m * sm[(c * kN + pivot) * 4];
            }
            __syncwarp();
        }
    }
    uint32_t p = kIdentityPerm, q = kIdentityPerm;
#pragma unroll
    for (int pivot = kN - 1; pivot >= 0; --pivot) p = nib_swap(p, pivot, rt[pivot]);
#pragma unroll
    for (int pivot = 0; pivot < kN; ++pivot) q = nib_swap(q, pivot, ct[pivot]);
    p_out = p;
    q_out = q;
    double const threshold = DBL_EPSILON * max_pivot;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < kN; ++i) {
        double const dd = sm[(i * kN + i) * 4];
        bad = bad || fabs(dd) < threshold || not_normal(dd);
    }
#pragma unroll
    for (int c = 0; c < kN; ++c) d[c] = mine[c * 24];
    return bad;
}

// ---- L1 prefetch of everything a row task will read ------------------------------------------------------------------------------
// A row task is a chain of dependent global round trips (index -> child -> its factor -> its U blocks -> targets ...), ~1-2 k cycles
// each; measured (PGMB_DEBUG_PHASES): 45 k cycles of a 73 k-cycle row task sit in the elimination step.  All those addresses follow
// from the shared index arrays, so the task starts by pulling the lines into L1; the six row threads share the work (a block is
// 36 elements x T scenarios x 8 B = 18 lines of 128 B for T = 8).
__device__ __forceinline__ void pf_l1(void const* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <int T> __device__ __forceinline__ void prefetch_block6(double const* blk, int r) { // thread r: column r (6 elements)
    double const* p = blk + (size_t)(r * kN) * T;
#pragma unroll
    for (int i = 0; i < kN; i += (T >= 16 ? 1 : 16 / T)) pf_l1(p + (size_t)i * T);
}
template <int T> __device__ __forceinline__ void prefetch_vec6(double const* v, int r) { // thread r: element r
    pf_l1(v + (size_t)r * T);
}

// ---- pieces of the up-sweep row task, each for block row r of the calling thread -----------------------------------------------
// block row of LU entry k of bus row `row` (zero for a fill-in) and my phase's row sums of N / H (Newton step)
template <int T, Mode mode>
__device__ __forceinline__ void build_entry6(DevStructure const& s, Tile6<T> const& t, bool dead_row, double const* uir, double const* uii,
                                             int k, double* bl, double& sn, double& sh) {
    int const r = t.r, p = r % kB;
    bool const top = r < kB;
    int ky = __ldg(s.map_y + k);
    int const j = __ldg(s.col_idx + k);
    if (t.dead != nullptr && ky >= 0 && (dead_row || t.dead[j] != 0)) ky = -1;
#pragma unroll
    for (int i = 0; i < kN; ++i) bl[i] = 0.0;
    sn = 0.0;
    sh = 0.0;
    if (ky >= 0) {
        double const* ysrc = s.ydata + (size_t)ky * kBB2;
        if (t.ovr_entry != nullptr) {
            for (int o = 0; o < t.ovr_n; ++o)
                if (t.ovr_entry[o] == ky) ysrc = t.ovr_y + o * kBB2;
        }
#pragma unroll
        for (int c = 0; c < kB; ++c) {
            double const yr = __ldg(ysrc + 2 * (p * kB + c)), yi = __ldg(ysrc + 2 * (p * kB + c) + 1);
            if constexpr (mode == Mode::newton) {
                double const ujr = t.u[(size_t)(j * kN + 2 * c) * T], uji = t.u[(size_t)(j * kN + 2 * c + 1) * T];
                double h, n;
                pf_term(yr, yi, uir[p], uii[p], ujr, uji, h, n);
                bl[c] = top ? h : -n;
                bl[kB + c] = top ? n : h;
                sn = c == 0 ? n : sn + n;
                sh = c == 0 ? h : sh + h;
            } else {
                bl[c] = top ? yr : yi;
                bl[kB + c] = top ? -yi : yr;
            }
        }
    }
}

// diagonal corrections, loads, sources (finish_diag of block_common.cuh) for row r of the diagonal block d
template <int T, Mode mode>
__device__ __forceinline__ void finish_diag6(DevStructure const& s, Tile6<T> const& t, int row, bool dead_row, double const* uir,
                                             double const* uii, double* d, double& acc_p, double& acc_q) {
    int const r = t.r, p = r % kB;
    bool const top = r < kB;
    if (dead_row) {
#pragma unroll
        for (int c = 0; c < kN; ++c) d[c] = (c == r) ? 1.0 : 0.0;
        acc_p = 0.0;
        acc_q = 0.0;
        return;
    }
    double dpp = sel6(d, p), dp3 = sel6(d, kB + p); // the two entries of my row that the corrections touch: columns p, 3 + p
    if constexpr (mode == Mode::newton) {
        if (top) {
            dpp += acc_q;
            dp3 += -acc_p;
        } else {
            dpp += -acc_p;
            dp3 += -acc_q;
        }
    }
    for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) {
        int const type = __ldg(s.lg_type + lg);
        double const ps = t.sinj[(size_t)(lg * kN + 2 * p) * T], qs = t.sinj[(size_t)(lg * kN + 2 * p + 1) * T];
        if constexpr (mode == Mode::newton) {
            double const v = t.pol[(size_t)(row * kN + kB + p) * T];
            if (type == 0) {
                acc_p += ps;
                acc_q += qs;
            } else if (type == 1) {
                acc_p += ps * v * v;
                acc_q += qs * v * v;
                dp3 += top ? -ps * 2.0 * v * v : -qs * 2.0 * v * v;
            } else {
                acc_p += ps * v;
                acc_q += qs * v;
                dp3 += top ? -ps * v : -qs * v;
            }
        } else {
            double const ylr = -ps, yli = qs;
            if (top) {
                dp3 += -yli;
                dpp += ylr;
            } else {
                dp3 += ylr;
                dpp += yli;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < kN; ++c) d[c] = (c == p) ? dpp : ((c == kB + p) ? dp3 : d[c]);
    for (int sr = __ldg(s.src_ptr + row), sre = __ldg(s.src_ptr + row + 1); sr < sre; ++sr) {
        double const* y = s.src_yref + (size_t)sr * kBB2;
        double const u0r = t.usrc[(size_t)(sr * 2) * T], u0i = t.usrc[(size_t)(sr * 2 + 1) * T];
        double usr[kB], usi[kB];
        double const a2r = -0.5, a2i = -0.8660254037844386, ar = -0.5, ai = 0.8660254037844386;
        usr[0] = u0r;
        usi[0] = u0i;
        usr[1] = u0r * a2r - u0i * a2i;
        usi[1] = u0r * a2i + u0i * a2r;
        usr[2] = u0r * ar - u0i * ai;
        usi[2] = u0r * ai + u0i * ar;
        if constexpr (mode == Mode::newton) {
            double mm[kN];
            double p_cal = 0.0, q_cal = 0.0;
#pragma unroll
            for (int c = 0; c < kB; ++c) {
                double const yr = __ldg(y + 2 * (p * kB + c)), yi = __ldg(y + 2 * (p * kB + c) + 1);
                double hmm, nmm, hms, nms;
                pf_term(1.0 * yr, 1.0 * yi, uir[p], uii[p], uir[c], uii[c], hmm, nmm);
                pf_term(-1.0 * yr, -1.0 * yi, uir[p], uii[p], usr[c], usi[c], hms, nms);
                mm[c] = top ? hmm : -nmm;
                mm[kB + c] = top ? nmm : hmm;
                p_cal = c == 0 ? nmm + nms : p_cal + (nmm + nms);
                q_cal = c == 0 ? hmm + hms : q_cal + (hmm + hms);
            }
#pragma unroll
            for (int c = 0; c < kN; ++c) {
                double add = 0.0;
                bool hit = false;
                if (c == p) {
                    add = top ? -q_cal : p_cal;
                    hit = true;
                }
                if (c == kB + p) {
                    add = top ? p_cal : q_cal;
                    hit = true;
                }
                if (hit) mm[c] += add;
            }
            acc_p -= p_cal;
            acc_q -= q_cal;
#pragma unroll
            for (int c = 0; c < kN; ++c) d[c] += mm[c];
        } else {
            double sr_ = 0.0, si_ = 0.0;
#pragma unroll
            for (int c = 0; c < kB; ++c) {
                double const yr = __ldg(y + 2 * (p * kB + c)), yi = __ldg(y + 2 * (p * kB + c) + 1);
                if (top) {
                    d[kB + c] -= yi;
                    d[c] += yr;
                } else {
                    d[kB + c] += yr;
                    d[c] += yi;
                }
                sr_ = c == 0 ? yr * usr[0] - yi * usi[0] : sr_ + (yr * usr[c] - yi * usi[c]);
                si_ = c == 0 ? yr * usi[0] + yi * usr[0] : si_ + (yr * usi[c] + yi * usr[c]);
            }
            acc_p += sr_;
            acc_q += si_;
        }
    }
}

// my row of L = (A Q_c) U_c^-1 for the lower entry whose (already column-permuted) row is in l[]; pp = factorised diagonal of c
template <int T> __device__ __forceinline__ void lower_row6(double const* pp, double* l) {
#pragma unroll
    for (int idx = 0; idx < kN; ++idx) {
#pragma unroll
        for (int prev = 0; prev < idx; ++prev) l[idx] -= pp[(size_t)(idx * kN + prev) * T] * l[prev];
        l[idx] /= pp[(size_t)(idx * kN + idx) * T];
    }
}
// my row of L * U for one update: out[cc] = sum_k l[k] * U[k][cc] in the reference's summation order
template <int T> __device__ __forceinline__ void schur_row6(double const* l, double const* ubp, double* out) {
#pragma unroll
    for (int cc = 0; cc < kN; ++cc) {
        double sum = l[0] * ubp[(size_t)(cc * kN) * T];
#pragma unroll
        for (int k = 1; k < kN; ++k) sum += l[k] * ubp[(size_t)(cc * kN + k) * T];
        out[cc] = sum;
    }
}

// after factorize6: store my row of the factor + the permutations, then column r of every U block and the forward substitution
// inside the block (x = L_pp^-1 (P t)); a_r = my element of the right-hand side
template <int T>
__device__ __forceinline__ void finish_row6(Tile6<T> const& t, int row, int dg, int e_begin, int e_end, int e_step, bool diag_part,
                                            double const* d, uint32_t pk, uint32_t qk, double a_r) {
    int const r = t.r;
    uint32_t pinv = 0;
#pragma unroll
    for (int i = 0; i < kN; ++i) pinv |= (uint32_t)i << (4 * min(nib(pk, i), kN - 1));
    double lo[kNN]; // unit-lower factors lo[prev * 6 + idx], prev < idx
    if (diag_part) { // factorised by this warp a moment ago: the factor is in the warp's scratch in logical order
        if (t.act) {
            double* dp = t.jac + ((size_t)dg * kNN + r) * T;
#pragma unroll
            for (int c = 0; c < kN; ++c) dp[(size_t)(c * kN) * T] = d[c];
            t.perm[(size_t)(row * 2 * kN + r) * T] = (uint8_t)nib(pk, r);
            t.perm[(size_t)(row * 2 * kN + kN + r) * T] = (uint8_t)nib(qk, r);
        }
#pragma unroll
        for (int idx = 0; idx < kN; ++idx)
#pragma unroll
            for (int prev = 0; prev < idx; ++prev) lo[prev * kN + idx] = t.sm[(prev * kN + idx) * 4];
    } else { // factorised by another warp before the last block barrier
        double const* dp = t.jac + (size_t)dg * kNN * T;
#pragma unroll
        for (int idx = 0; idx < kN; ++idx)
#pragma unroll
            for (int prev = 0; prev < idx; ++prev) lo[prev * kN + idx] = dp[(size_t)(prev * kN + idx) * T];
    }
    for (int e = e_begin; e < e_end; e += e_step) { // U blocks, column r of each: L_pp^-1 (P A), row permutation in the load addresses
        double* ap = t.jac + ((size_t)e * kNN + r * kN) * T;
        double col[kN];
#pragma unroll
        for (int jj = 0; jj < kN; ++jj) col[jj] = ap[(size_t)nib(pinv, jj) * T];
#pragma unroll
        for (int idx = 0; idx < kN; ++idx)
#pragma unroll
            for (int prev = 0; prev < idx; ++prev) col[idx] -= lo[prev * kN + idx] * col[prev];
        if (t.act) {
#pragma unroll
            for (int jj = 0; jj < kN; ++jj) ap[(size_t)jj * T] = col[jj];
        }
    }
    if (diag_part) { // x = L_pp^-1 (P t): thread i drops its element at position p[i]; every thread solves, stores its own
        double* const vs = t.sm + kNN * 4;
        if (t.real) vs[nib(pk, r) * 4] = a_r;
        __syncwarp();
        double xr[kN];
#pragma unroll
        for (int jj = 0; jj < kN; ++jj) xr[jj] = vs[jj * 4];
#pragma unroll
        for (int idx = 0; idx < kN; ++idx)
#pragma unroll
            for (int prev = 0; prev < idx; ++prev) xr[idx] -= lo[prev * kN + idx] * xr[prev];
        if (t.act) t.xvec[(size_t)(row * kN + r) * T] = sel6(xr, r);
        __syncwarp(); // the scratch is reused by the next row task of this warp
    }
}

// operands of the up-sweep task of `row`: the children's factors, permutations, right-hand sides and U blocks, the voltages of the
// neighbours, the loads of the bus
template <int T, Mode mode> __device__ __forceinline__ void prefetch_up6(DevStructure const& s, Tile6<T> const& t, int row) {
    int const r = t.r;
    int const rb = __ldg(s.row_ptr + row), re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    for (int e = rb; e < dg; ++e) {
        int const c = __ldg(s.col_idx + e);
        prefetch_block6<T>(t.jac + (size_t)__ldg(s.diag + c) * kNN * T, r);
        prefetch_vec6<T>(t.xvec + (size_t)(c * kN) * T, r);
        if (r == 0) pf_l1(t.perm + (size_t)(c * 2 * kN + kN) * T);
        for (int q = __ldg(s.upd_ptr + e), qe = __ldg(s.upd_ptr + e + 1); q < qe; ++q)
            prefetch_block6<T>(t.jac + (size_t)__ldg(s.upd_u + q) * kNN * T, r);
    }
    if constexpr (mode == Mode::newton) {
        for (int k = rb; k < re; ++k) prefetch_vec6<T>(t.u + (size_t)(__ldg(s.col_idx + k) * kN) * T, r);
        prefetch_vec6<T>(t.pol + (size_t)(row * kN) * T, r);
    }
    for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) prefetch_vec6<T>(t.sinj + (size_t)(lg * kN) * T, r);
}

// ---- up-sweep row task (build + eliminate + factorise + U blocks + forward substitution) ------------------------------------
template <int T, Mode mode>
__device__ bool up_row6(DevStructure const& s, Tile6<T> const& t, int row, bool prefetched, int next_row, unsigned long long* steps = nullptr) {
    int const r = t.r;
    long long c0 = steps != nullptr ? clock64() : 0;
    auto step = [&](int k) { // PGMB_DEBUG_PHASES: cycles of one designated warp per step of the row task (slots 5, 6, 7)
        if (steps != nullptr) {
            long long const c1 = clock64();
            steps[k] += (unsigned long long)(c1 - c0);
            c0 = c1;
        }
    };
    bool const top = r < kB;
    int const rb = __ldg(s.row_ptr + row), re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    bool const dead_row = t.dead != nullptr && t.dead[row] != 0;
    double uir[kB], uii[kB];
#pragma unroll
    for (int c = 0; c < kB; ++c) {
        uir[c] = t.u[(size_t)(row * kN + 2 * c) * T];
        uii[c] = t.u[(size_t)(row * kN + 2 * c + 1) * T];
    }
    double acc_p = 0.0, acc_q = 0.0; // NR: -P / -Q of phase p then mismatch ; linear: rhs real / imag
    double d[kN];
#pragma unroll
    for (int i = 0; i < kN; ++i) d[i] = 0.0;

    // 0. the operands of this task (first task of a level) and of the warp's next task in the level go into L1 now
    if (!prefetched) prefetch_up6<T, mode>(s, t, row);
    if (next_row >= 0) prefetch_up6<T, mode>(s, t, next_row);

    // 1. build my block row of every entry
    for (int k = rb; k < re; ++k) {
        double bl[kN], sn, sh;
        build_entry6<T, mode>(s, t, dead_row, uir, uii, k, bl, sn, sh);
        if constexpr (mode == Mode::newton) {
            acc_p -= sn;
            acc_q -= sh;
        }
        if (k == dg) {
#pragma unroll
            for (int i = 0; i < kN; ++i) d[i] = bl[i];
        } else if (t.act) {
            double* bp = t.jac + ((size_t)k * kNN + r) * T;
#pragma unroll
            for (int c = 0; c < kN; ++c) bp[(size_t)(c * kN) * T] = bl[c];
        }
    }
    finish_diag6<T, mode>(s, t, row, dead_row, uir, uii, d, acc_p, acc_q);
    double a_r = top ? acc_p : acc_q; // my element of the right-hand side
    step(5);

    // 2. eliminate against finished rows: my row of L, of every update and of the rhs are thread-local
    for (int e = rb; e < dg; ++e) {
        int const c = __ldg(s.col_idx + e);
        double const* ap = t.jac + (size_t)e * kNN * T;
        double const* pp = t.jac + (size_t)__ldg(s.diag + c) * kNN * T;
        uint8_t const* qp = t.perm + (size_t)(c * 2 * kN + kN) * T;
        double l[kN];
#pragma unroll
        for (int i = 0; i < kN; ++i) l[i] = ap[(size_t)(perm_at(qp + (size_t)i * T) * kN + r) * T];
        lower_row6<T>(pp, l);
        for (int q = __ldg(s.upd_ptr + e), qe = __ldg(s.upd_ptr + e + 1); q < qe; ++q) {
            int const ui = __ldg(s.upd_u + q), ai = __ldg(s.upd_a + q);
            double sum[kN];
            schur_row6<T>(l, t.jac + (size_t)ui * kNN * T, sum);
            if (ai == dg) {
#pragma unroll
                for (int cc = 0; cc < kN; ++cc) d[cc] -= sum[cc];
            } else if (t.act) {
                double* tp = t.jac + (size_t)ai * kNN * T;
#pragma unroll
                for (int cc = 0; cc < kN; ++cc) tp[(size_t)(cc * kN + r) * T] -= sum[cc];
            }
        }
        double sum = l[0] * t.xvec[(size_t)(c * kN) * T];
#pragma unroll
        for (int k = 1; k < kN; ++k) sum += l[k] * t.xvec[(size_t)(c * kN + k) * T];
        a_r -= sum;
    }

    step(6);
    // 3. factorise the diagonal block across the six row threads; 4. U blocks; 5. forward substitution
    uint32_t pk, qk;
    bool const singular = factorize6s(t.sm, d, r, t.real, pk, qk);
    finish_row6<T>(t, row, dg, dg + 1, re, 1, true, d, pk, qk, a_r);
    step(7);
    return singular;
}

// ---- cooperative elimination of a hub row (symbolic.hpp: WideRowPlan; block_common.cuh: wide_up_row), row-split -----------------
// Called by ALL threads of the block (barriers inside).  slot6 / n_slot6: this warp's position among the warps that serve the same
// group of four scenarios.  Phases as in wide_up_row: 1 entries built in parallel (+ row-sum contributions to scratch), 2 one warp
// adds the contributions in entry order and applies loads / sources, 3 per sub-level the children in parallel (incoming terms
// subtracted in ascending child order, L row, update terms and rhs terms to scratch), 4 diagonal / upper entries subtract their
// terms in ascending child order and one warp subtracts the rhs terms, 5 one warp factorises the diagonal block and
// forward-substitutes, 6 U blocks in parallel.  Per matrix element the operations and their order are those of wide_up_row.
template <int T, Mode mode>
__device__ void wide_up_row6(DevStructure const& s, Tile6<T> const& t, int w, int slot6, int n_slot6, bool warp_active, bool& singular) {
    int const r = t.r, sc = t.sc, p = r % kB;
    bool const top = r < kB;
    int32_t const* const tab = s.wide_table + 8 * w;
    int const row = __ldg(tab), n_sub = __ldg(tab + 1);
    int32_t const* const sub_ptr = s.wide_data + __ldg(tab + 2);
    int32_t const* const order = s.wide_data + __ldg(tab + 3);
    int32_t const* const in_ptr = s.wide_data + __ldg(tab + 4);
    int32_t const* const in_idx = s.wide_data + __ldg(tab + 5);
    int const rb = __ldg(s.row_ptr + row), re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    int const n_entries = re - rb, n_lower = dg - rb;
    int const upd_base = __ldg(s.upd_ptr + rb);
    bool const dead_row = t.dead != nullptr && t.dead[row] != 0;
    double uir[kB], uii[kB];
#pragma unroll
    for (int c = 0; c < kB; ++c) {
        uir[c] = t.u[(size_t)(row * kN + 2 * c) * T];
        uii[c] = t.u[(size_t)(row * kN + 2 * c + 1) * T];
    }
    auto load_row = [&](int k, double* a) {
        double const* ap = t.jac + (size_t)k * kNN * T;
#pragma unroll
        for (int c = 0; c < kN; ++c) a[c] = ap[(size_t)(c * kN + r) * T];
    };
    auto store_row = [&](int k, double const* a) {
        if (!t.act) return;
        double* ap = t.jac + (size_t)k * kNN * T;
#pragma unroll
        for (int c = 0; c < kN; ++c) ap[(size_t)(c * kN + r) * T] = a[c];
    };
    auto sub_terms = [&](int entry_pos, double* a) { // a -= incoming terms, ascending child order
        for (int q = __ldg(in_ptr + entry_pos), qe = __ldg(in_ptr + entry_pos + 1); q < qe; ++q) {
            double const* term = t.wide_terms + (size_t)__ldg(in_idx + q) * kNN * T;
#pragma unroll
            for (int c = 0; c < kN; ++c) a[c] -= term[(size_t)(c * kN + r) * T];
        }
    };
    // 1
    if (warp_active) {
        for (int idx = slot6; idx < n_entries; idx += n_slot6) {
            double bl[kN], sn, sh;
            build_entry6<T, mode>(s, t, dead_row, uir, uii, rb + idx, bl, sn, sh);
            store_row(rb + idx, bl);
            if (t.act) t.wide_sum[(size_t)(idx * kN + r) * T] = top ? sn : sh;
        }
    }
    __syncthreads();
    // 2
    if (warp_active && slot6 == 0) {
        double acc_p = 0.0, acc_q = 0.0;
        if constexpr (mode == Mode::newton) {
            for (int idx = 0; idx < n_entries; ++idx) {
                if (__ldg(s.map_y + rb + idx) < 0) continue;
                acc_p -= t.wide_sum[(size_t)(idx * kN + p) * T];
                acc_q -= t.wide_sum[(size_t)(idx * kN + kB + p) * T];
            }
        }
        double d[kN];
        load_row(dg, d);
        finish_diag6<T, mode>(s, t, row, dead_row, uir, uii, d, acc_p, acc_q);
        store_row(dg, d);
        if (t.act) t.xvec[(size_t)(row * kN + r) * T] = top ? acc_p : acc_q;
    }
    // 3 (phase 2 writes the diagonal block and the rhs of the row only, which phase 3 does not read: no barrier)
    for (int sl = 0; sl < n_sub; ++sl) {
        if (warp_active) {
            for (int oi = __ldg(sub_ptr + sl) + slot6; oi < __ldg(sub_ptr + sl + 1); oi += n_slot6) {
                int const pos = __ldg(order + oi);
                int const e = rb + pos;
                int const c = __ldg(s.col_idx + e);
                double a[kN], l[kN];
                load_row(e, a);
                sub_terms(pos, a);
                uint8_t const* qp = t.perm + (size_t)(c * 2 * kN + kN) * T;
#pragma unroll
                for (int i = 0; i < kN; ++i) l[i] = sel6(a, perm_at(qp + (size_t)i * T));
                lower_row6<T>(t.jac + (size_t)__ldg(s.diag + c) * kNN * T, l);
                for (int q = __ldg(s.upd_ptr + e), qe = __ldg(s.upd_ptr + e + 1); q < qe; ++q) {
                    double sum[kN];
                    schur_row6<T>(l, t.jac + (size_t)__ldg(s.upd_u + q) * kNN * T, sum);
                    if (t.act) {
                        double* term = t.wide_terms + (size_t)(q - upd_base) * kNN * T;
#pragma unroll
                        for (int cc = 0; cc < kN; ++cc) term[(size_t)(cc * kN + r) * T] = sum[cc];
                    }
                }
                double sum = l[0] * t.xvec[(size_t)(c * kN) * T];
#pragma unroll
                for (int k = 1; k < kN; ++k) sum += l[k] * t.xvec[(size_t)(c * kN + k) * T];
                if (t.act) t.wide_rhs[(size_t)(pos * kN + r) * T] = sum;
            }
        }
        __syncthreads();
    }
    // 4
    if (warp_active) {
        for (int pos = n_lower + slot6; pos < n_entries; pos += n_slot6) {
            double a[kN];
            load_row(rb + pos, a);
            sub_terms(pos, a);
            store_row(rb + pos, a);
        }
        if (slot6 == n_slot6 - 1) { // the L * x_c terms in ascending child order (another warp than the diagonal block's when possible)
            double a_r = t.xvec[(size_t)(row * kN + r) * T];
            for (int pos = 0; pos < n_lower; ++pos) a_r -= t.wide_rhs[(size_t)(pos * kN + r) * T];
            if (t.act) t.xvec[(size_t)(row * kN + r) * T] = a_r;
        }
    }
    __syncthreads();
    // 5 (the U blocks of the row are finished in phase 6 by all warps)
    if (warp_active && slot6 == 0) {
        double d[kN];
        load_row(dg, d);
        double const a_r = t.xvec[(size_t)(row * kN + r) * T];
        uint32_t pk, qk;
        singular |= factorize6s(t.sm, d, r, t.real, pk, qk);
        finish_row6<T>(t, row, dg, 0, 0, 1, true, d, pk, qk, a_r);
    }
    __syncthreads();
    // 6
    if (warp_active && dg + 1 + slot6 < re) {
        uint32_t pk = 0;
#pragma unroll
        for (int i = 0; i < kN; ++i) pk |= (uint32_t)perm_at(t.perm + (size_t)(row * 2 * kN + i) * T) << (4 * i);
        double const dummy[kN] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        finish_row6<T>(t, row, dg, dg + 1 + slot6, re, n_slot6, false, dummy, pk, 0u, 0.0);
    }
    __syncthreads();
}

// ---- down-sweep row task ----------------------------------------------------------------------------------------------------
template <int T, Mode mode> __device__ __forceinline__ void prefetch_down6(DevStructure const& s, Tile6<T> const& t, int row) {
    int const r = t.r;
    int const re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    for (int e = re - 1; e > dg; --e) {
        int const j = __ldg(s.col_idx + e);
        prefetch_block6<T>(t.jac + (size_t)e * kNN * T, r);
        prefetch_vec6<T>(t.xvec + (size_t)(j * kN) * T, r);
        if (r == 0) pf_l1(t.perm + (size_t)(j * 2 * kN + kN) * T);
    }
    prefetch_block6<T>(t.jac + (size_t)dg * kNN * T, r);
    prefetch_vec6<T>(t.xvec + (size_t)(row * kN) * T, r);
    if (r == 0) pf_l1(t.perm + (size_t)(row * 2 * kN + kN) * T);
    if constexpr (mode == Mode::newton) {
        prefetch_vec6<T>(t.pol + (size_t)(row * kN) * T, r);
        prefetch_vec6<T>(t.u + (size_t)(row * kN) * T, r);
    }
}
template <int T, Mode mode> __device__ double down_row6(DevStructure const& s, Tile6<T> const& t, int row, bool prefetched, int next_row) {
    int const r = t.r;
    int const re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    if (!prefetched) prefetch_down6<T, mode>(s, t, row);
    if (next_row >= 0) prefetch_down6<T, mode>(s, t, next_row);
    double y_r = t.xvec[(size_t)(row * kN + r) * T];
    for (int e = re - 1; e > dg; --e) {
        int const j = __ldg(s.col_idx + e);
        uint8_t const* qp = t.perm + (size_t)(j * 2 * kN + kN) * T;
        double const* up = t.jac + (size_t)e * kNN * T;
        int const q0 = perm_at(qp);
        double sum = up[(size_t)(q0 * kN + r) * T] * t.xvec[(size_t)(j * kN + q0) * T];
#pragma unroll
        for (int i = 1; i < kN; ++i) {
            int const qi = perm_at(qp + (size_t)i * T);
            sum += up[(size_t)(qi * kN + r) * T] * t.xvec[(size_t)(j * kN + qi) * T];
        }
        y_r -= sum;
    }
    // gather the six elements through the warp's scratch; every thread solves the block, the solution goes back through the
    // scratch at its permuted position: x[q[i]] = y[i]
    double* const vs = t.sm + kNN * 4;
    if (t.real) vs[r * 4] = y_r;
    __syncwarp();
    double y[kN];
#pragma unroll
    for (int i = 0; i < kN; ++i) y[i] = vs[i * 4];
    double const* dp = t.jac + (size_t)dg * kNN * T;
#pragma unroll
    for (int step = 0; step < kN; ++step) {
        int const idx = kN - 1 - step;
#pragma unroll
        for (int ps = 0; ps < step; ++ps) {
            int const prev = kN - 1 - ps;
            y[idx] -= dp[(size_t)(prev * kN + idx) * T] * y[prev];
        }
        y[idx] /= dp[(size_t)(idx * kN + idx) * T];
    }
    uint8_t const* qr = t.perm + (size_t)(row * 2 * kN + kN) * T;
    __syncwarp();
    if (t.real) vs[perm_at(qr + (size_t)r * T) * 4] = sel6(y, r); // thread i places y[i] at position q[i]
    __syncwarp();
    double const x_r = vs[r * 4];
    if (t.act) t.xvec[(size_t)(row * kN + r) * T] = x_r;
    double dev = 0.0;
    if (r < kB && t.act) { // thread p updates phase p
        int const p = r;
        double const xa = x_r, xb = vs[(kB + p) * 4];
        double* const pth = t.pol + (size_t)(row * kN + p) * T;
        double* const pv = t.pol + (size_t)(row * kN + kB + p) * T;
        double* const pur = t.u + (size_t)(row * kN + 2 * p) * T;
        double* const pui = t.u + (size_t)(row * kN + 2 * p + 1) * T;
        if constexpr (mode == Mode::newton) {
            double theta = *pth, v = *pv;
            theta += xa;
            v += v * xb;
            double sn, cs;
            sincos(theta, &sn, &cs);
            double const nr = v * cs, ni = v * sn;
            double const dr = nr - *pur, di = ni - *pui;
            *pth = theta;
            *pv = v;
            *pur = nr;
            *pui = ni;
            dev = sqrt(dr * dr + di * di);
        } else {
            *pur = xa;
            *pui = xb;
            *pv = sqrt(xa * xa + xb * xb);
            *pth = atan2(xb, xa);
        }
    }
    __syncwarp(); // the scratch is reused by the next row task of this warp
    return dev;
}

template <int T, Mode mode>
__device__ void sweeps6(DevStructure const& s, Tile6<T>& t6, int slot6, int n_slot6, bool& singular6, double& dev,
                        unsigned long long* phase) {
    long long t0 = clock64();
    auto lap = [&](int k) { // PGMB_DEBUG_PHASES: up level 0 | up wide rows | up other levels | down levels >= 1 | down level 0
        if (phase != nullptr && threadIdx.x == 0) {
            long long const t1 = clock64();
            phase[k] += (unsigned long long)(t1 - t0);
            t0 = t1;
        }
    };
    bool const warp_active = __any_sync(kFull, t6.act); // a warp whose four scenarios are all finished skips its row tasks
    for (int lv = 0; lv < s.n_level; ++lv) {
        int const b = __ldg(s.level_ptr + lv), e = __ldg(s.level_ptr + lv + 1);
        if (warp_active) {
            // (prefetching the warp's NEXT row of the level during the current task was measured: no gain, 37.5 -> 38.9 ms)
            for (int i = b + slot6; i < e; i += n_slot6) {
                int const row = __ldg(s.level_rows + i);
                if (s.n_wide != 0 && __ldg(s.row_is_wide + row)) continue; // eliminated below by the whole block
                singular6 |= up_row6<T, mode>(s, t6, row, false, -1, (phase != nullptr && threadIdx.x == 0 && lv != 0) ? phase : nullptr);
            }
        }
        __syncthreads();
        lap(lv == 0 ? 0 : 2);
        if (s.n_wide != 0)
            for (int w = __ldg(s.wide_level_ptr + lv); w < __ldg(s.wide_level_ptr + lv + 1); ++w)
                wide_up_row6<T, mode>(s, t6, w, slot6, n_slot6, warp_active, singular6);
        lap(1);
    }
    for (int lv = s.n_level - 1; lv >= 0; --lv) {
        int const b = __ldg(s.level_ptr + lv), e = __ldg(s.level_ptr + lv + 1);
        if (warp_active)
            for (int i = b + slot6; i < e; i += n_slot6) dev = fmax(dev, down_row6<T, mode>(s, t6, __ldg(s.level_rows + i), false, -1));
        __syncthreads();
        lap(lv == 0 ? 4 : 3);
    }
}

template <int T, int MAXT> __global__ void __launch_bounds__(MAXT, 1) nr_block6_kernel(DevStructure s, DevBatch b, SolveOptions opt) {
    constexpr int TS = T / 4; // groups of four scenarios per tile
    __shared__ unsigned long long sh_dev[T];
    __shared__ int sh_singular[T], sh_done[T], sh_status[T], sh_iter[T];
    __shared__ double sh_max_dev[T];
    __shared__ double sh_scratch[MAXT / 32][(kNN + kN) * 4]; // per warp: one 6 x 6 block + one 6-vector for each of its 4 scenarios
    int const tile = blockIdx.x;
    int const warp = threadIdx.x / 32, wl = threadIdx.x % 32;
    int const n_warp = blockDim.x / 32;
    int const st = warp % TS, slot6 = warp / TS, n_slot6 = n_warp / TS;
    int const r8 = wl >> 2, sc = wl & 3;
    int const ln6 = st * 4 + sc; // my scenario lane within the tile
    Tile6<T> t6;
    {
        t6.jac = b.jac + (size_t)tile * s.nnz_lu * kNN * T + ln6;
        t6.xvec = b.xvec + (size_t)tile * s.n_bus * kN * T + ln6;
        t6.pol = b.pol + (size_t)tile * s.n_bus * kN * T + ln6;
        t6.u = b.u + (size_t)tile * s.n_bus * kN * T + ln6;
        t6.perm = b.perm + (size_t)tile * s.n_bus * 2 * kN * T + ln6;
        t6.sinj = b.sinj + (size_t)tile * s.n_load_gen * kN * T + ln6;
        t6.usrc = b.usrc + (size_t)tile * s.n_source * 2 * T + ln6;
        int64_t const scn = (int64_t)tile * T + ln6;
        bool const valid = scn < b.n_scn;
        t6.ovr_n = 4 * b.ovl.n_branch;
        t6.ovr_entry = (b.ovl.entry != nullptr && valid) ? b.ovl.entry + scn * t6.ovr_n : nullptr;
        t6.ovr_y = (b.ovl.entry != nullptr && valid) ? b.ovl.y + scn * t6.ovr_n * kBB2 : nullptr;
        t6.dead = (b.ovl.dead_off != nullptr && valid && b.ovl.dead_off[scn] >= 0) ? b.ovl.dead + (size_t)b.ovl.dead_off[scn] * s.n_bus : nullptr;
        t6.wide_terms = b.wide_terms ? b.wide_terms + (size_t)tile * s.wide_max_upd * kNN * T + ln6 : nullptr;
        t6.wide_rhs = b.wide_rhs ? b.wide_rhs + (size_t)tile * s.wide_max_lower * kN * T + ln6 : nullptr;
        t6.wide_sum = b.wide_sum ? b.wide_sum + (size_t)tile * s.wide_max_entries * kN * T + ln6 : nullptr;
    }
    t6.sm = &sh_scratch[warp][sc];
    t6.r = r8 < kN ? r8 : r8 - kN;
    t6.sc = sc;
    t6.real = r8 < kN;
    if (threadIdx.x < T) {
        int64_t const scn = (int64_t)tile * T + threadIdx.x;
        sh_dev[threadIdx.x] = 0ull;
        sh_singular[threadIdx.x] = 0;
        sh_done[threadIdx.x] = scn < b.n_scn ? 0 : 1;
        sh_status[threadIdx.x] = kStatusOk;
        sh_iter[threadIdx.x] = 0;
        sh_max_dev[threadIdx.x] = INFINITY;
    }
    __syncthreads();
    auto run = [&](auto mode_tag) {
        constexpr Mode mode = decltype(mode_tag)::value;
        bool const done6 = sh_done[ln6] != 0;
        t6.act = !done6 && r8 < kN;
        bool singular6 = false;
        double dev = 0.0;
        unsigned long long* const phase = b.phase_cycles ? b.phase_cycles + tile * 16 + (mode == Mode::newton ? 8 : 0) : nullptr;
        sweeps6<T, mode>(s, t6, slot6, n_slot6, singular6, dev, phase);
        if (!done6 && singular6) sh_singular[ln6] = 1;
        if (t6.act && mode == Mode::newton) atomicMax(&sh_dev[ln6], (unsigned long long)__double_as_longlong(dev));
        __syncthreads();
    };
    using LinTag = std::integral_constant<Mode, Mode::linear_init>;
    using NrTag = std::integral_constant<Mode, Mode::newton>;
    run(LinTag{});
    if (threadIdx.x < T && !sh_done[threadIdx.x] && sh_singular[threadIdx.x]) {
        sh_status[threadIdx.x] = kStatusSingular;
        sh_done[threadIdx.x] = 1;
    }
    __syncthreads();
    while (true) {
        if (threadIdx.x < T && !sh_done[threadIdx.x]) {
            if (sh_iter[threadIdx.x] == opt.max_iter) {
                sh_status[threadIdx.x] = kStatusDiverged;
                sh_done[threadIdx.x] = 1;
            } else {
                ++sh_iter[threadIdx.x];
            }
        }
        __syncthreads();
        bool any = false;
#pragma unroll
        for (int i = 0; i < T; ++i) any |= sh_done[i] == 0;
        if (!any) break;
        run(NrTag{});
        if (threadIdx.x < T && !sh_done[threadIdx.x]) {
            if (sh_singular[threadIdx.x]) {
                sh_status[threadIdx.x] = kStatusSingular;
                sh_done[threadIdx.x] = 1;
            } else {
                double const md = __longlong_as_double((long long)sh_dev[threadIdx.x]);
                sh_max_dev[threadIdx.x] = md;
                if (!(md > opt.err_tol)) sh_done[threadIdx.x] = 1;
            }
        }
        if (threadIdx.x < T) sh_dev[threadIdx.x] = 0ull;
        __syncthreads();
    }
    if (threadIdx.x < T) {
        int64_t const scn = (int64_t)tile * T + threadIdx.x;
        if (scn < b.n_scn) {
            b.status[scn] = sh_status[threadIdx.x];
            b.n_iter[scn] = sh_iter[threadIdx.x];
            b.max_dev[scn] = sh_max_dev[threadIdx.x];
        }
    }
}

template <int T> void launch_b6(DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int threads, cudaStream_t st) {
    int const unit = 32 * (T / 4); // every warp runs one row task for four scenarios
    threads = std::max(unit, threads / unit * unit);
    if (threads > 768) {
        nr_block6_kernel<T, 1024><<<b.n_tile, std::min(threads, 1024 / unit * unit), 0, st>>>(s, b, opt);
    } else if (threads > 512) {
        nr_block6_kernel<T, 768><<<b.n_tile, threads, 0, st>>>(s, b, opt);
    } else {
        nr_block6_kernel<T, 512><<<b.n_tile, threads, 0, st>>>(s, b, opt);
    }
}

} // namespace

void launch_nr_block6(int tw, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int threads, cudaStream_t st) {
    count_kernel_launch();
    switch (tw) {
    case 4: launch_b6<4>(s, b, opt, threads, st); break;
    case 8: launch_b6<8>(s, b, opt, threads, st); break;
    case 16: launch_b6<16>(s, b, opt, threads, st); break;
    default: launch_b6<32>(s, b, opt, threads, st); break;
    }
}

} // namespace pgmb
