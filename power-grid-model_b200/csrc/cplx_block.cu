// Complex-domain solvers with B x B complex blocks (asymmetric calculation, B = 3), batched like nr_block.cu:
//   linear_block_kernel      LinearPFSolver<asymmetric_t>::run_power_flow (math_solver/linear_pf_solver.hpp:67-113): per scenario
//                            Y + diag(-conj(S_load)) + Y_source on the diagonal blocks, rhs = Y_source * U_ref, one block
//                            factorisation with full pivoting inside the blocks + solve (sparse_lu_solver.hpp:346-495, 769-827)
//   ic_factor_block_kernel   IterativeCurrentPFSolver::initialize_derived_solver (iterative_current_pf_solver.hpp:95-123):
//                            Y + Y_source factorised once per parameter set, L and U blocks and the block permutations kept
//   ic_iterate_block_kernel  the iteration (:126-160, 172-225) with the shared factor
// Storage: a block is column-major like the reference's Eigen blocks; element (r, c) of LU entry k lives at
// ((k * B*B + c * B + r) * 2 + {re, im}) * T + lane.  L blocks are kept with their rows in the order of the row's own
// variables (the reference permutes them by P of the row when that row becomes the pivot; applying P to the accumulated
// right-hand side afterwards performs the same operations on the same numbers).  U blocks are kept without the later
// column permutation Q_j of their column's pivot; the backward substitution walks the columns in the permuted order.
#include "result_common.cuh"

#include <cfloat>
#include <cmath>
#include <cuda_runtime.h>

namespace pgmb {
using namespace res;
namespace {

constexpr int kStatusOk = 0, kStatusDiverged = 1, kStatusSingular = 2;

__device__ __forceinline__ bool not_normal_d(double x) { return !(fabs(x) >= DBL_MIN) || isinf(x); }
__device__ __forceinline__ bool not_normal_c(C v) { // is_normal(complex) (three_phase_tensor.hpp:380-392)
    if (v.r == 0.0) return not_normal_d(v.i);
    if (v.i == 0.0) return not_normal_d(v.r);
    return not_normal_d(v.r) || not_normal_d(v.i);
}

// matrix / vector views; TM = lane stride of the matrix (T, or 1 for the shared iterative-current factor)
template <int TM, int B> struct CMatView {
    double* m;
    uint8_t* perm; // [bus][2B] : p[B] then q[B], lane stride TM
    __device__ __forceinline__ void load(int k, C* a) const {
        double const* p = m + (size_t)k * B * B * 2 * TM;
#pragma unroll
        for (int i = 0; i < B * B; ++i) a[i] = C{p[(size_t)(2 * i) * TM], p[(size_t)(2 * i + 1) * TM]};
    }
    __device__ __forceinline__ void store(int k, C const* a) const {
        double* p = m + (size_t)k * B * B * 2 * TM;
#pragma unroll
        for (int i = 0; i < B * B; ++i) {
            p[(size_t)(2 * i) * TM] = a[i].r;
            p[(size_t)(2 * i + 1) * TM] = a[i].i;
        }
    }
    __device__ __forceinline__ uint8_t p_of(int bus, int i) const { return perm[(size_t)(bus * 2 * B + i) * TM]; }
    __device__ __forceinline__ uint8_t q_of(int bus, int i) const { return perm[(size_t)(bus * 2 * B + B + i) * TM]; }
};
template <int T, int B> struct CVecView {
    double* x; // [bus][B][2][T]
    __device__ __forceinline__ C get(int bus, int p) const {
        return C{x[(size_t)((bus * B + p) * 2) * T], x[(size_t)((bus * B + p) * 2 + 1) * T]};
    }
    __device__ __forceinline__ void set(int bus, int p, C v) const {
        x[(size_t)((bus * B + p) * 2) * T] = v.r;
        x[(size_t)((bus * B + p) * 2 + 1) * T] = v.i;
    }
};

// DenseLUFactor::factorize_block_in_place for a complex B x B block (sparse_lu_solver.hpp:86-165)
template <int B> __device__ bool factorize_cblock(C* m, uint8_t* p, uint8_t* q) {
    int rt[B], ct[B];
    double max_pivot = 0.0;
    for (int pivot = 0; pivot < B; ++pivot) {
        int rb = pivot, cb = pivot;
        C const first = m[pivot * B + pivot];
        double best = first.r * first.r + first.i * first.i;
        for (int c = pivot; c < B; ++c)
            for (int r = pivot; r < B; ++r) {
                C const v = m[c * B + r];
                double const sc = v.r * v.r + v.i * v.i;
                if (sc > best) {
                    best = sc;
                    rb = r;
                    cb = c;
                }
            }
        if (best == 0.0) {
            for (int k = pivot; k < B; ++k) {
                rt[k] = k;
                ct[k] = k;
            }
            break;
        }
        max_pivot = fmax(max_pivot, sqrt(best));
        rt[pivot] = rb;
        ct[pivot] = cb;
        if (rb != pivot)
            for (int c = 0; c < B; ++c) {
                C const x = m[c * B + pivot];
                m[c * B + pivot] = m[c * B + rb];
                m[c * B + rb] = x;
            }
        if (cb != pivot)
            for (int r = 0; r < B; ++r) {
                C const x = m[pivot * B + r];
                m[pivot * B + r] = m[cb * B + r];
                m[cb * B + r] = x;
            }
        if (pivot < B - 1) {
            for (int r = pivot + 1; r < B; ++r) m[pivot * B + r] = cdiv(m[pivot * B + r], m[pivot * B + pivot]);
            for (int c = pivot + 1; c < B; ++c)
                for (int r = pivot + 1; r < B; ++r) m[c * B + r] = csub(m[c * B + r], cmul(m[pivot * B + r], m[c * B + pivot]));
        }
    }
    for (int i = 0; i < B; ++i) {
        p[i] = (uint8_t)i;
        q[i] = (uint8_t)i;
    }
    for (int pivot = B - 1; pivot >= 0; --pivot) {
        uint8_t const x = p[pivot];
        p[pivot] = p[rt[pivot]];
        p[rt[pivot]] = x;
    }
    for (int pivot = 0; pivot < B; ++pivot) {
        uint8_t const x = q[pivot];
        q[pivot] = q[ct[pivot]];
        q[ct[pivot]] = x;
    }
    double const threshold = DBL_EPSILON * max_pivot;
    bool singular = false;
    for (int pivot = 0; pivot < B; ++pivot) {
        C const d = m[pivot * B + pivot];
        singular = singular || sqrt(d.r * d.r + d.i * d.i) < threshold || not_normal_c(d);
    }
    return singular;
}

// xr -= dot(blk, xc)   (sparse_lu_solver.hpp solve_once: sum over k in ascending order, then one subtraction)
template <int B> __device__ __forceinline__ void sub_dot(C* xr, C const* blk, C const* xc) {
#pragma unroll
    for (int r = 0; r < B; ++r) {
        C sum = cmul(blk[0 * B + r], xc[0]);
#pragma unroll
        for (int k = 1; k < B; ++k) sum = cadd(sum, cmul(blk[k * B + r], xc[k]));
        xr[r] = csub(xr[r], sum);
    }
}

// (u, u a^2, u a): ComplexValue<asymmetric_t>{u}   (three_phase_tensor.hpp:47-53)
template <int B> __device__ __forceinline__ void rotated(C u, C* out) {
    out[0] = u;
    if constexpr (B == 3) {
        C const a2{-0.5, -0.8660254037844386}, a{-0.5, 0.8660254037844386};
        out[1] = cmul(u, a2);
        out[2] = cmul(u, a);
    }
}

// factorise row `row`.  with_loads: linear PF matrix (loads as admittance) + fused forward substitution of the rhs;
// otherwise the shared iterative-current factor whose L blocks are kept.
template <int T, int TM, int B, bool with_loads>
__device__ bool factor_row(DevStructure const& s, CMatView<TM, B> const& mv, CVecView<T, B> const& xv, double const* sinj,
                           double const* usrc, int row) {
    constexpr int BB = B * B;
    int const rb = __ldg(s.row_ptr + row), re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    C d[BB];
    for (int k = rb; k < re; ++k) {
        int const ky = __ldg(s.map_y + k);
        C blk[BB];
#pragma unroll
        for (int r = 0; r < B; ++r)
#pragma unroll
            for (int c = 0; c < B; ++c) blk[c * B + r] = ky >= 0 ? ldc(s.ydata, (int64_t)ky * BB + r * B + c) : C{0.0, 0.0};
        if (k == dg) {
#pragma unroll
            for (int i = 0; i < BB; ++i) d[i] = blk[i];
        } else {
            mv.store(k, blk);
        }
    }
    C rhs[B];
#pragma unroll
    for (int p = 0; p < B; ++p) rhs[p] = C{0.0, 0.0};
    if constexpr (with_loads) {
        for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) {
#pragma unroll
            for (int p = 0; p < B; ++p) {
                double const ps = sinj[(size_t)((lg * B + p) * 2) * T], qs = sinj[(size_t)((lg * B + p) * 2 + 1) * T];
                d[p * B + p] = cadd(d[p * B + p], C{-ps, qs}); // -conj(s)
            }
        }
    }
    for (int sr = __ldg(s.src_ptr + row), sre = __ldg(s.src_ptr + row + 1); sr < sre; ++sr) {
        C y[BB]; // row-major like src_yref
#pragma unroll
        for (int i = 0; i < BB; ++i) y[i] = ldc(s.src_yref, (int64_t)sr * BB + i);
#pragma unroll
        for (int c = 0; c < B; ++c)
#pragma unroll
            for (int r = 0; r < B; ++r) d[c * B + r] = cadd(d[c * B + r], y[r * B + c]);
        if constexpr (with_loads) {
            C us[B];
            rotated<B>(C{usrc[(size_t)(sr * 2) * T], usrc[(size_t)(sr * 2 + 1) * T]}, us);
#pragma unroll
            for (int r = 0; r < B; ++r) {
                C sum = cmul(y[r * B], us[0]);
#pragma unroll
                for (int k = 1; k < B; ++k) sum = cadd(sum, cmul(y[r * B + k], us[k]));
                rhs[r] = cadd(rhs[r], sum);
            }
        }
    }
    // eliminate against finished rows
    for (int e = rb; e < dg; ++e) {
        int const c = __ldg(s.col_idx + e);
        C a[BB], piv[BB], l[BB];
        mv.load(e, a);
        mv.load(__ldg(s.diag + c), piv);
        uint8_t qc[B];
#pragma unroll
        for (int i = 0; i < B; ++i) qc[i] = mv.q_of(c, i);
#pragma unroll
        for (int i = 0; i < B; ++i)
#pragma unroll
            for (int r = 0; r < B; ++r) l[i * B + r] = a[qc[i] * B + r];
        for (int idx = 0; idx < B; ++idx) { // right / upper triangular solve
            for (int prev = 0; prev < idx; ++prev) {
                C const uv = piv[idx * B + prev];
#pragma unroll
                for (int r = 0; r < B; ++r) l[idx * B + r] = csub(l[idx * B + r], cmul(uv, l[prev * B + r]));
            }
            C const dd = piv[idx * B + idx];
#pragma unroll
            for (int r = 0; r < B; ++r) l[idx * B + r] = cdiv(l[idx * B + r], dd);
        }
        if constexpr (!with_loads) mv.store(e, l);
        for (int q = __ldg(s.upd_ptr + e), qe = __ldg(s.upd_ptr + e + 1); q < qe; ++q) {
            int const ui = __ldg(s.upd_u + q), ai = __ldg(s.upd_a + q);
            C ub[BB], tb[BB];
            mv.load(ui, ub);
            C* tgt = d;
            if (ai != dg) {
                mv.load(ai, tb);
                tgt = tb;
            }
            for (int cc = 0; cc < B; ++cc)
#pragma unroll
                for (int r = 0; r < B; ++r) {
                    C sum = cmul(l[0 * B + r], ub[cc * B + 0]);
#pragma unroll
                    for (int k = 1; k < B; ++k) sum = cadd(sum, cmul(l[k * B + r], ub[cc * B + k]));
                    tgt[cc * B + r] = csub(tgt[cc * B + r], sum);
                }
            if (ai != dg) mv.store(ai, tb);
        }
        if constexpr (with_loads) {
            C xc[B];
#pragma unroll
            for (int i = 0; i < B; ++i) xc[i] = xv.get(c, i);
            sub_dot<B>(rhs, l, xc);
        }
    }
    uint8_t p[B], q[B];
    bool const singular = factorize_cblock<B>(d, p, q);
    mv.store(dg, d);
#pragma unroll
    for (int i = 0; i < B; ++i) {
        mv.perm[(size_t)(row * 2 * B + i) * TM] = p[i];
        mv.perm[(size_t)(row * 2 * B + B + i) * TM] = q[i];
    }
    for (int e = dg + 1; e < re; ++e) { // U blocks: L_pp^-1 (P A)
        C a[BB], ub[BB];
        mv.load(e, a);
#pragma unroll
        for (int cc = 0; cc < B; ++cc)
#pragma unroll
            for (int i = 0; i < B; ++i) ub[cc * B + p[i]] = a[cc * B + i];
        for (int idx = 0; idx < B; ++idx)
            for (int prev = 0; prev < idx; ++prev) {
                C const lv = d[prev * B + idx];
#pragma unroll
                for (int cc = 0; cc < B; ++cc) ub[cc * B + idx] = csub(ub[cc * B + idx], cmul(lv, ub[cc * B + prev]));
            }
        mv.store(e, ub);
    }
    if constexpr (with_loads) { // forward substitution inside the block: x = L_pp^-1 (P rhs)
        C xr[B];
#pragma unroll
        for (int i = 0; i < B; ++i) xr[p[i]] = rhs[i];
        for (int idx = 0; idx < B; ++idx)
            for (int prev = 0; prev < idx; ++prev) xr[idx] = csub(xr[idx], cmul(d[prev * B + idx], xr[prev]));
#pragma unroll
        for (int i = 0; i < B; ++i) xv.set(row, i, xr[i]);
    }
    return singular;
}

// forward substitution of one row with stored L blocks (iterative current): x_row = L_pp^-1 P (rhs - sum L x_c)
template <int T, int TM, int B>
__device__ void forward_row(DevStructure const& s, CMatView<TM, B> const& mv, CVecView<T, B> const& xv, C* acc, int row) {
    constexpr int BB = B * B;
    int const rb = __ldg(s.row_ptr + row), dg = __ldg(s.diag + row);
    for (int e = rb; e < dg; ++e) {
        int const c = __ldg(s.col_idx + e);
        C l[BB], xc[B];
        mv.load(e, l);
#pragma unroll
        for (int i = 0; i < B; ++i) xc[i] = xv.get(c, i);
        sub_dot<B>(acc, l, xc);
    }
    C d[BB], xr[B];
    mv.load(dg, d);
#pragma unroll
    for (int i = 0; i < B; ++i) xr[mv.p_of(row, i)] = acc[i];
    for (int idx = 0; idx < B; ++idx)
        for (int prev = 0; prev < idx; ++prev) xr[idx] = csub(xr[idx], cmul(d[prev * B + idx], xr[prev]));
#pragma unroll
    for (int i = 0; i < B; ++i) xv.set(row, i, xr[i]);
}

// backward substitution of one row; out = final solution of the row (column permutation applied)
template <int T, int TM, int B>
__device__ void backward_row(DevStructure const& s, CMatView<TM, B> const& mv, CVecView<T, B> const& xv, int row, C* out) {
    constexpr int BB = B * B;
    int const re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    C y[B];
#pragma unroll
    for (int i = 0; i < B; ++i) y[i] = xv.get(row, i);
    for (int e = re - 1; e > dg; --e) {
        int const j = __ldg(s.col_idx + e);
        C ub[BB], xj[B];
        uint8_t qj[B];
        mv.load(e, ub);
#pragma unroll
        for (int i = 0; i < B; ++i) {
            xj[i] = xv.get(j, i);
            qj[i] = mv.q_of(j, i);
        }
#pragma unroll
        for (int r = 0; r < B; ++r) {
            C sum = cmul(ub[qj[0] * B + r], xj[qj[0]]);
#pragma unroll
            for (int i = 1; i < B; ++i) sum = cadd(sum, cmul(ub[qj[i] * B + r], xj[qj[i]]));
            y[r] = csub(y[r], sum);
        }
    }
    C d[BB];
    mv.load(dg, d);
    for (int step = 0; step < B; ++step) {
        int const idx = B - 1 - step;
        for (int ps = 0; ps < step; ++ps) {
            int const prev = B - 1 - ps;
            y[idx] = csub(y[idx], cmul(d[prev * B + idx], y[prev]));
        }
        y[idx] = cdiv(y[idx], d[idx * B + idx]);
    }
#pragma unroll
    for (int i = 0; i < B; ++i) out[mv.q_of(row, i)] = y[i];
#pragma unroll
    for (int i = 0; i < B; ++i) xv.set(row, i, out[i]);
}

// ---- linear ----------------------------------------------------------------------------------------------------------
template <int T, int B> __global__ void linear_block_kernel(DevStructure s, DevBatch b) {
    constexpr int N = 2 * B;
    __shared__ int sh_singular[T];
    int const lane = threadIdx.x % T, slot = threadIdx.x / T, n_slot = blockDim.x / T, tile = blockIdx.x;
    int64_t const scn = (int64_t)tile * T + lane;
    bool const valid = scn < b.n_scn;
    CMatView<T, B> const mv{b.jac + (size_t)tile * s.nnz_lu * N * N * T + lane, b.perm + (size_t)tile * s.n_bus * 2 * N * T + lane};
    CVecView<T, B> const xv{b.xvec + (size_t)tile * s.n_bus * N * T + lane};
    CVecView<T, B> const uv{b.u + (size_t)tile * s.n_bus * N * T + lane};
    double const* const sinj = b.sinj + (size_t)tile * s.n_load_gen * N * T + lane;
    double const* const usrc = b.usrc + (size_t)tile * s.n_source * 2 * T + lane;
    if (threadIdx.x < T) sh_singular[threadIdx.x] = 0;
    __syncthreads();
    bool singular = false;
    for (int lv = 0; lv < s.n_level; ++lv) {
        if (valid)
            for (int i = __ldg(s.level_ptr + lv) + slot; i < __ldg(s.level_ptr + lv + 1); i += n_slot)
                singular |= factor_row<T, T, B, true>(s, mv, xv, sinj, usrc, __ldg(s.level_rows + i));
        __syncthreads();
    }
    for (int lv = s.n_level - 1; lv >= 0; --lv) {
        if (valid)
            for (int i = __ldg(s.level_ptr + lv) + slot; i < __ldg(s.level_ptr + lv + 1); i += n_slot) {
                int const row = __ldg(s.level_rows + i);
                C out[B];
                backward_row<T, T, B>(s, mv, xv, row, out);
#pragma unroll
                for (int p = 0; p < B; ++p) uv.set(row, p, out[p]);
            }
        __syncthreads();
    }
    if (singular) sh_singular[lane] = 1;
    __syncthreads();
    if (slot == 0 && valid) {
        b.status[scn] = sh_singular[lane] ? kStatusSingular : kStatusOk;
        b.n_iter[scn] = 1;
        b.max_dev[scn] = 0.0;
    }
}

// ---- iterative current -----------------------------------------------------------------------------------------------
template <int B> __global__ void ic_factor_block_kernel(DevStructure s, double* factor, uint8_t* perm, int* flag) {
    CMatView<1, B> const mv{factor, perm};
    CVecView<1, B> const none{nullptr};
    bool singular = false;
    for (int lv = 0; lv < s.n_level; ++lv) {
        for (int i = __ldg(s.level_ptr + lv) + threadIdx.x; i < __ldg(s.level_ptr + lv + 1); i += blockDim.x)
            singular |= factor_row<1, 1, B, false>(s, mv, none, nullptr, nullptr, __ldg(s.level_rows + i));
        __syncthreads();
    }
    if (singular) *flag = 1;
}

template <int T, int B>
__global__ void ic_iterate_block_kernel(DevStructure s, DevBatch b, SolveOptions opt, double* factor, uint8_t* factor_perm,
                                        int const* __restrict__ factor_flag) {
    constexpr int N = 2 * B, BB = B * B;
    __shared__ unsigned long long sh_dev[T];
    int const lane = threadIdx.x % T, slot = threadIdx.x / T, n_slot = blockDim.x / T, tile = blockIdx.x;
    int64_t const scn = (int64_t)tile * T + lane;
    bool const valid = scn < b.n_scn;
    CMatView<1, B> const mv{factor, factor_perm};
    CVecView<T, B> const xv{b.xvec + (size_t)tile * s.n_bus * N * T + lane};
    CVecView<T, B> const uv{b.u + (size_t)tile * s.n_bus * N * T + lane};
    double const* const sinj = b.sinj + (size_t)tile * s.n_load_gen * N * T + lane;
    double const* const usrc = b.usrc + (size_t)tile * s.n_source * 2 * T + lane;
    if (threadIdx.x < T) sh_dev[threadIdx.x] = 0ull;
    bool const factor_singular = *factor_flag != 0;

    if (valid) { // flat start (make_flat_start :207-225)
        C sum{0.0, 0.0};
        for (int sr = 0; sr < s.n_source; ++sr) {
            double sn, cs;
            sincos(-__ldg(s.phase_shift + __ldg(s.src_bus + sr)), &sn, &cs);
            sum = cadd(sum, cmul(C{usrc[(size_t)(sr * 2) * T], usrc[(size_t)(sr * 2 + 1) * T]}, C{cs, sn}));
        }
        C const u_ref{sum.r / (double)s.n_source, sum.i / (double)s.n_source};
        for (int i = slot; i < s.n_bus; i += n_slot) {
            double sn, cs;
            sincos(__ldg(s.phase_shift + i), &sn, &cs);
            C rot[B];
            rotated<B>(cmul(u_ref, C{cs, sn}), rot);
#pragma unroll
            for (int p = 0; p < B; ++p) uv.set(i, p, rot[p]);
        }
    }
    __syncthreads();

    bool done = !valid;
    int status = kStatusOk, num_iter = 0;
    double max_dev = INFINITY;
    if (!done && factor_singular) {
        status = kStatusSingular;
        done = true;
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
        double dev = 0.0;
        for (int lv = 0; lv < s.n_level; ++lv) {
            if (!done) {
                for (int i = __ldg(s.level_ptr + lv) + slot; i < __ldg(s.level_ptr + lv + 1); i += n_slot) {
                    int const row = __ldg(s.level_rows + i);
                    C ui[B], rhs[B];
#pragma unroll
                    for (int p = 0; p < B; ++p) {
                        ui[p] = uv.get(row, p);
                        rhs[p] = C{0.0, 0.0};
                    }
                    for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) {
                        int const type = __ldg(s.lg_type + lg);
#pragma unroll
                        for (int p = 0; p < B; ++p) {
                            C const sv{sinj[(size_t)((lg * B + p) * 2) * T], sinj[(size_t)((lg * B + p) * 2 + 1) * T]};
                            if (type == 0) {
                                rhs[p] = cadd(rhs[p], conj(cdiv(sv, ui[p])));
                            } else if (type == 1) {
                                rhs[p] = cadd(rhs[p], cmul(conj(sv), ui[p]));
                            } else {
                                rhs[p] = cadd(rhs[p], conj(cdiv(cscale(sv, sqrt(ui[p].r * ui[p].r + ui[p].i * ui[p].i)), ui[p])));
                            }
                        }
                    }
                    for (int sr = __ldg(s.src_ptr + row), sre = __ldg(s.src_ptr + row + 1); sr < sre; ++sr) {
                        C us[B];
                        rotated<B>(C{usrc[(size_t)(sr * 2) * T], usrc[(size_t)(sr * 2 + 1) * T]}, us);
#pragma unroll
                        for (int r = 0; r < B; ++r) {
                            C sum = cmul(ldc(s.src_yref, (int64_t)sr * BB + r * B), us[0]);
#pragma unroll
                            for (int k = 1; k < B; ++k) sum = cadd(sum, cmul(ldc(s.src_yref, (int64_t)sr * BB + r * B + k), us[k]));
                            rhs[r] = cadd(rhs[r], sum);
                        }
                    }
                    forward_row<T, 1, B>(s, mv, xv, rhs, row);
                }
            }
            __syncthreads();
        }
        for (int lv = s.n_level - 1; lv >= 0; --lv) {
            if (!done) {
                for (int i = __ldg(s.level_ptr + lv) + slot; i < __ldg(s.level_ptr + lv + 1); i += n_slot) {
                    int const row = __ldg(s.level_rows + i);
                    C out[B];
                    backward_row<T, 1, B>(s, mv, xv, row, out);
                    double dev_bus = 0.0;
#pragma unroll
                    for (int p = 0; p < B; ++p) {
                        C const uo = uv.get(row, p);
                        double const dr = out[p].r - uo.r, di = out[p].i - uo.i;
                        double const dp = sqrt(dr * dr + di * di);
                        dev_bus = p == 0 ? dp : fmax(dev_bus, dp);
                        uv.set(row, p, out[p]);
                    }
                    dev = fmax(dev_bus, dev);
                }
            }
            __syncthreads();
        }
        if (!done) atomicMax(&sh_dev[lane], (unsigned long long)__double_as_longlong(dev));
        __syncthreads();
        if (!done) {
            max_dev = __longlong_as_double((long long)sh_dev[lane]);
            if (!(max_dev > opt.err_tol)) done = true;
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

} // namespace

void launch_linear_asym(int tw, DevStructure const& s, DevBatch const& b, int n_slot, cudaStream_t st) {
    count_kernel_launch();
    switch (tw) {
    case 4: linear_block_kernel<4, 3><<<b.n_tile, 4 * n_slot, 0, st>>>(s, b); break;
    case 8: linear_block_kernel<8, 3><<<b.n_tile, 8 * n_slot, 0, st>>>(s, b); break;
    case 16: linear_block_kernel<16, 3><<<b.n_tile, 16 * n_slot, 0, st>>>(s, b); break;
    default: linear_block_kernel<32, 3><<<b.n_tile, 32 * n_slot, 0, st>>>(s, b); break;
    }
}
void launch_ic_factor_asym(DevStructure const& s, double* factor, uint8_t* perm, int* flag, cudaStream_t st) {
    count_kernel_launch();
    cudaMemsetAsync(flag, 0, sizeof(int), st);
    ic_factor_block_kernel<3><<<1, 128, 0, st>>>(s, factor, perm, flag);
}
void launch_ic_iterate_asym(int tw, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, double* factor,
                            uint8_t* perm, int const* flag, int n_slot, cudaStream_t st) {
    count_kernel_launch();
    switch (tw) {
    case 4: ic_iterate_block_kernel<4, 3><<<b.n_tile, 4 * n_slot, 0, st>>>(s, b, opt, factor, perm, flag); break;
    case 8: ic_iterate_block_kernel<8, 3><<<b.n_tile, 8 * n_slot, 0, st>>>(s, b, opt, factor, perm, flag); break;
    case 16: ic_iterate_block_kernel<16, 3><<<b.n_tile, 16 * n_slot, 0, st>>>(s, b, opt, factor, perm, flag); break;
    default: ic_iterate_block_kernel<32, 3><<<b.n_tile, 32 * n_slot, 0, st>>>(s, b, opt, factor, perm, flag); break;
    }
}

} // namespace pgmb
