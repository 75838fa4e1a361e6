// Symmetric complex-domain solvers on the scalar (1x1 block) sparse LU, batched like nr_sym.cu (one thread block = one tile of
// T scenarios, lane = scenario, slots work on the rows of one dependency level):
//   linear_sym_kernel     LinearPFSolver::run_power_flow (math_solver/linear_pf_solver.hpp:67-92): per scenario
//                         Y + diag(-conj(S_load)) + diag(Y_source), rhs = Y_source * U_ref, one factorisation + solve
//                         (prepare_linear_matrix_and_rhs, common_solver_functions.hpp:33-66; scalar LU without pivoting,
//                         sparse_lu_solver.hpp:377-388, 464-467, 777-826)
//   ic_factor_kernel      IterativeCurrentPFSolver::initialize_derived_solver (iterative_current_pf_solver.hpp:95-123):
//                         Y + diag(Y_source) factorised ONCE per parameter set and shared by every scenario and iteration
//   ic_iterate_sym_kernel the iteration (:126-160, 172-225): flat start, injected currents by load type, two triangular
//                         sweeps with the shared factor, max |U_new - U_old| per scenario; linear_current = one iteration
// Complex products / quotients follow std::complex (libgcc __muldc3 / __divdc3 operation order), see result_common.cuh.
#include "result_common.cuh"

#include <cfloat>
#include <cmath>
#include <cuda_runtime.h>

namespace pgmb {
using namespace res;
namespace {

constexpr int kStatusOk = 0, kStatusDiverged = 1, kStatusSingular = 2;

__device__ __forceinline__ bool not_normal_d(double x) { return !(fabs(x) >= DBL_MIN) || isinf(x); }
// is_normal(complex) (three_phase_tensor.hpp:380-392)
__device__ __forceinline__ bool not_normal_c(C v) {
    if (v.r == 0.0) return not_normal_d(v.i);
    if (v.i == 0.0) return not_normal_d(v.r);
    return not_normal_d(v.r) || not_normal_d(v.i);
}

template <int T> struct CTile { // complex entries: [item][2][T]
    double* m;       // matrix entries of this tile (or the shared factor with T = 1)
    double* x;       // rhs / solution
    double* u;       // voltages
    double const* sinj;
    double const* usrc;
    __device__ __forceinline__ C ldm(int k) const { return {m[(size_t)(k * 2) * T], m[(size_t)(k * 2 + 1) * T]}; }
    __device__ __forceinline__ void stm(int k, C v) const {
        m[(size_t)(k * 2) * T] = v.r;
        m[(size_t)(k * 2 + 1) * T] = v.i;
    }
    __device__ __forceinline__ C ldx(int i) const { return {x[(size_t)(i * 2) * T], x[(size_t)(i * 2 + 1) * T]}; }
    __device__ __forceinline__ void stx(int i, C v) const {
        x[(size_t)(i * 2) * T] = v.r;
        x[(size_t)(i * 2 + 1) * T] = v.i;
    }
    __device__ __forceinline__ C ldu(int i) const { return {u[(size_t)(i * 2) * T], u[(size_t)(i * 2 + 1) * T]}; }
    __device__ __forceinline__ void stu(int i, C v) const {
        u[(size_t)(i * 2) * T] = v.r;
        u[(size_t)(i * 2 + 1) * T] = v.i;
    }
};

// factorise row `row` (IKJ, scalar entries).  with_loads: linear PF matrix (loads as admittance) and fused forward
// substitution of the right-hand side; otherwise the shared iterative-current factor whose L entries are kept.
template <int T, bool with_loads>
__device__ __forceinline__ bool factor_row(DevStructure const& s, CTile<T> const& t, int row) {
    int const rb = __ldg(s.row_ptr + row), re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    C d{0.0, 0.0};
    for (int k = rb; k < re; ++k) {
        int const ky = __ldg(s.map_y + k);
        C const v = ky >= 0 ? ldc(s.ydata, ky) : C{0.0, 0.0};
        if (k == dg) {
            d = v;
        } else {
            t.stm(k, v);
        }
    }
    C rhs{0.0, 0.0};
    if constexpr (with_loads) {
        for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) {
            double const ps = t.sinj[(size_t)(lg * 2) * T], qs = t.sinj[(size_t)(lg * 2 + 1) * T];
            d = cadd(d, C{-ps, qs}); // -conj(s)
        }
    }
    for (int sr = __ldg(s.src_ptr + row), sre = __ldg(s.src_ptr + row + 1); sr < sre; ++sr) {
        C const y = ldc(s.src_yref, sr);
        d = cadd(d, y);
        if constexpr (with_loads) {
            C const us{t.usrc[(size_t)(sr * 2) * T], t.usrc[(size_t)(sr * 2 + 1) * T]};
            rhs = cadd(rhs, cmul(y, us));
        }
    }
    for (int e = rb; e < dg; ++e) {
        int const c = __ldg(s.col_idx + e);
        C const piv = t.ldm(__ldg(s.diag + c));
        C const l = cdiv(t.ldm(e), piv);
        if constexpr (!with_loads) t.stm(e, l);
        for (int q = __ldg(s.upd_ptr + e), qe = __ldg(s.upd_ptr + e + 1); q < qe; ++q) {
            int const ui = __ldg(s.upd_u + q), ai = __ldg(s.upd_a + q);
            C const lu = cmul(l, t.ldm(ui));
            if (ai == dg) {
                d = csub(d, lu);
            } else {
                t.stm(ai, csub(t.ldm(ai), lu));
            }
        }
        if constexpr (with_loads) rhs = csub(rhs, cmul(l, t.ldx(c)));
    }
    t.stm(dg, d);
    if constexpr (with_loads) t.stx(row, rhs);
    return not_normal_c(d);
}

template <int T> __device__ __forceinline__ C backward_row(DevStructure const& s, CTile<T> const& t, int row) {
    int const re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    C x = t.ldx(row);
    for (int e = re - 1; e > dg; --e) x = csub(x, cmul(t.ldm(e), t.ldx(__ldg(s.col_idx + e))));
    x = cdiv(x, t.ldm(dg));
    t.stx(row, x);
    return x;
}

// ---- linear ----------------------------------------------------------------------------------------------------------
template <int T> __global__ void __launch_bounds__(512, 2) linear_sym_kernel(DevStructure s, DevBatch b) {
    __shared__ int sh_singular[T];
    int const lane = threadIdx.x % T, slot = threadIdx.x / T, n_slot = blockDim.x / T, tile = blockIdx.x;
    int64_t const scn = (int64_t)tile * T + lane;
    bool const valid = scn < b.n_scn;
    CTile<T> const t{b.jac + (size_t)tile * s.nnz_lu * 4 * T + lane, b.xvec + (size_t)tile * s.n_bus * 2 * T + lane,
                     b.u + (size_t)tile * s.n_bus * 2 * T + lane, b.sinj + (size_t)tile * s.n_load_gen * 2 * T + lane,
                     b.usrc + (size_t)tile * s.n_source * 2 * T + lane};
    if (threadIdx.x < T) sh_singular[threadIdx.x] = 0;
    __syncthreads();
    bool singular = false;
    for (int lv = 0; lv < s.n_level; ++lv) {
        if (valid)
            for (int i = __ldg(s.level_ptr + lv) + slot; i < __ldg(s.level_ptr + lv + 1); i += n_slot)
                singular |= factor_row<T, true>(s, t, __ldg(s.level_rows + i));
        __syncthreads();
    }
    for (int lv = s.n_level - 1; lv >= 0; --lv) {
        if (valid)
            for (int i = __ldg(s.level_ptr + lv) + slot; i < __ldg(s.level_ptr + lv + 1); i += n_slot) {
                int const row = __ldg(s.level_rows + i);
                t.stu(row, backward_row<T>(s, t, row));
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
// shared factor: one "scenario" (T = 1 layout) in factor[nnz_lu][2]; flag[0] = 1 when a pivot is not normal
__global__ void ic_factor_kernel(DevStructure s, double* factor, int* flag) {
    CTile<1> const t{factor, nullptr, nullptr, nullptr, nullptr};
    bool singular = false;
    for (int lv = 0; lv < s.n_level; ++lv) {
        for (int i = __ldg(s.level_ptr + lv) + threadIdx.x; i < __ldg(s.level_ptr + lv + 1); i += blockDim.x)
            singular |= factor_row<1, false>(s, t, __ldg(s.level_rows + i));
        __syncthreads();
    }
    if (singular) *flag = 1;
}

template <int T>
__global__ void __launch_bounds__(512, 2) ic_iterate_sym_kernel(DevStructure s, DevBatch b, SolveOptions opt, double const* __restrict__ factor,
                                      int const* __restrict__ factor_flag) {
    __shared__ unsigned long long sh_dev[T];
    int const lane = threadIdx.x % T, slot = threadIdx.x / T, n_slot = blockDim.x / T, tile = blockIdx.x;
    int64_t const scn = (int64_t)tile * T + lane;
    bool const valid = scn < b.n_scn;
    double* const x = b.xvec + (size_t)tile * s.n_bus * 2 * T + lane;
    double* const u = b.u + (size_t)tile * s.n_bus * 2 * T + lane;
    double const* const sinj = b.sinj + (size_t)tile * s.n_load_gen * 2 * T + lane;
    double const* const usrc = b.usrc + (size_t)tile * s.n_source * 2 * T + lane;
    auto ldx = [&](int i) { return C{x[(size_t)(i * 2) * T], x[(size_t)(i * 2 + 1) * T]}; };
    auto stx = [&](int i, C v) {
        x[(size_t)(i * 2) * T] = v.r;
        x[(size_t)(i * 2 + 1) * T] = v.i;
    };
    auto ldu = [&](int i) { return C{u[(size_t)(i * 2) * T], u[(size_t)(i * 2 + 1) * T]}; };
    auto ldf = [&](int k) { return C{__ldg(factor + 2 * k), __ldg(factor + 2 * k + 1)}; };
    if (threadIdx.x < T) sh_dev[threadIdx.x] = 0ull;
    bool const factor_singular = *factor_flag != 0;

    // flat start (make_flat_start :207-225): mean source voltage, de-rotated by the bus phase shift, re-rotated per bus
    if (valid) {
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
            C const v = cmul(u_ref, C{cs, sn});
            u[(size_t)(i * 2) * T] = v.r;
            u[(size_t)(i * 2 + 1) * T] = v.i;
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
        // up-sweep: injected current of the bus, then forward substitution with the shared L
        for (int lv = 0; lv < s.n_level; ++lv) {
            if (!done) {
                int const lv_end = __ldg(s.level_ptr + lv + 1);
                for (int i = __ldg(s.level_ptr + lv) + slot; i < lv_end; i += n_slot) {
                    int const row = __ldg(s.level_rows + i);
                    if (i + n_slot < lv_end) { // pull the voltage and the loads of this thread's next bus into L1
                        int const nx = __ldg(s.level_rows + i + n_slot);
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(u + (size_t)(nx * 2) * T));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(u + (size_t)(nx * 2 + 1) * T));
                        for (int lg = __ldg(s.lg_ptr + nx), lge = __ldg(s.lg_ptr + nx + 1); lg < lge; ++lg) {
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(sinj + (size_t)(lg * 2) * T));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(sinj + (size_t)(lg * 2 + 1) * T));
                        }
                    }
                    C const ui = ldu(row);
                    C rhs{0.0, 0.0};
                    for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) {
                        C const sv{sinj[(size_t)(lg * 2) * T], sinj[(size_t)(lg * 2 + 1) * T]};
                        int const type = __ldg(s.lg_type + lg);
                        if (type == 0) {
                            rhs = cadd(rhs, conj(cdiv(sv, ui)));
                        } else if (type == 1) {
                            rhs = cadd(rhs, cmul(conj(sv), ui));
                        } else {
                            rhs = cadd(rhs, conj(cdiv(cscale(sv, sqrt(ui.r * ui.r + ui.i * ui.i)), ui)));
                        }
                    }
                    for (int sr = __ldg(s.src_ptr + row), sre = __ldg(s.src_ptr + row + 1); sr < sre; ++sr) {
                        rhs = cadd(rhs, cmul(ldc(s.src_yref, sr), C{usrc[(size_t)(sr * 2) * T], usrc[(size_t)(sr * 2 + 1) * T]}));
                    }
                    for (int e = __ldg(s.row_ptr + row), dg = __ldg(s.diag + row); e < dg; ++e)
                        rhs = csub(rhs, cmul(ldf(e), ldx(__ldg(s.col_idx + e))));
                    stx(row, rhs);
                }
            }
            __syncthreads();
        }
        for (int lv = s.n_level - 1; lv >= 0; --lv) {
            if (!done) {
                for (int i = __ldg(s.level_ptr + lv) + slot; i < __ldg(s.level_ptr + lv + 1); i += n_slot) {
                    int const row = __ldg(s.level_rows + i);
                    int const re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
                    C xr = ldx(row);
                    for (int e = re - 1; e > dg; --e) xr = csub(xr, cmul(ldf(e), ldx(__ldg(s.col_idx + e))));
                    xr = cdiv(xr, ldf(dg));
                    stx(row, xr);
                    C const uo = ldu(row);
                    double const dr = xr.r - uo.r, di = xr.i - uo.i;
                    dev = fmax(dev, sqrt(dr * dr + di * di));
                    u[(size_t)(row * 2) * T] = xr.r;
                    u[(size_t)(row * 2 + 1) * T] = xr.i;
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

void launch_linear_sym(int tw, DevStructure const& s, DevBatch const& b, int n_slot, cudaStream_t st) {
    count_kernel_launch();
    switch (tw) {
    case 4: linear_sym_kernel<4><<<b.n_tile, 4 * n_slot, 0, st>>>(s, b); break;
    case 8: linear_sym_kernel<8><<<b.n_tile, 8 * n_slot, 0, st>>>(s, b); break;
    case 16: linear_sym_kernel<16><<<b.n_tile, 16 * n_slot, 0, st>>>(s, b); break;
    default: linear_sym_kernel<32><<<b.n_tile, 32 * n_slot, 0, st>>>(s, b); break;
    }
}
void launch_ic_factor(DevStructure const& s, double* factor, int* flag, cudaStream_t st) {
    count_kernel_launch();
    cudaMemsetAsync(flag, 0, sizeof(int), st);
    ic_factor_kernel<<<1, 256, 0, st>>>(s, factor, flag);
}
void launch_ic_iterate_sym(int tw, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, double const* factor,
                           int const* flag, int n_slot, cudaStream_t st) {
    count_kernel_launch();
    switch (tw) {
    case 4: ic_iterate_sym_kernel<4><<<b.n_tile, 4 * n_slot, 0, st>>>(s, b, opt, factor, flag); break;
    case 8: ic_iterate_sym_kernel<8><<<b.n_tile, 8 * n_slot, 0, st>>>(s, b, opt, factor, flag); break;
    case 16: ic_iterate_sym_kernel<16><<<b.n_tile, 16 * n_slot, 0, st>>>(s, b, opt, factor, flag); break;
    default: ic_iterate_sym_kernel<32><<<b.n_tile, 32 * n_slot, 0, st>>>(s, b, opt, factor, flag); break;
    }
}

} // namespace pgmb
