// Batched block-sparse LU solve WITH pivot perturbation and iterative refinement:
//   sparse_lu_solver.hpp:30-33 (constants), 39-48 (perturb_pivot_if_needed), 86-165 (dense full-pivot block LU),
//   346-495 (prefactorize, right-looking, in place), 514-538 (solve_with_refinement), 580-622 (iterate_and_backward_error,
//   calculate_residual), 624-649 (initialize_pivot_perturbation: block-off-diagonal infinity norm), 769-827 (solve_once).
// No power-flow solver of the reference switches the perturbation on (use_pivot_perturbation = false at every PF call site;
// only state estimation passes true), so the Newton-Raphson / iterative-current kernels of this library carry none of it.  This
// kernel is the same solver behind the same flag for a batch of systems that share one pattern: every system of the batch has
// its own matrix values and right-hand side; one thread factorises and solves one system, walking the pattern exactly like the
// reference (pivot by pivot, L blocks permuted / solved when their pivot is reached, Schur updates through find_entry).
// Memory: system-major ([system][entry][block]) -- this entry point serves the known-answer tests and callers that need the
// reference's perturbation semantics, not the throughput path.
#include "engine.hpp"
#include "result_common.cuh"

#include <cfloat>

namespace pgmb {
namespace {

using res::C;

constexpr double kEpsPerturbation = 1e-13;      // epsilon_perturbation
constexpr double kCapBackErrDenominator = 1e-4; // cap_back_error_denominator
constexpr int kMaxRefinement = 5;               // max_iterative_refinement

// scalar traits: double / complex with the operation order of std::complex (products: naive formula; division: __divdc3)
struct Real {
    using S = double;
    static constexpr int W = 1; // doubles per scalar
    __device__ static S ld(double const* p) { return p[0]; }
    __device__ static void st(double* p, S v) { p[0] = v; }
    __device__ static S zero() { return 0.0; }
    __device__ static S mul(S a, S b) { return a * b; }
    __device__ static S add(S a, S b) { return a + b; }
    __device__ static S sub(S a, S b) { return a - b; }
    __device__ static S div(S a, S b) { return a / b; }
    __device__ static double abs2(S a) { return a * a; }
    __device__ static double cabs(S a) { return fabs(a); }
    __device__ static bool is_normal(S a) { return fabs(a) >= DBL_MIN && !isinf(a); }
    __device__ static S scale_to(S v, double abs_v, double target) { // (value / abs_value) * threshold, 1 * threshold for 0
        S const sc = abs_v == 0.0 ? 1.0 : v / abs_v;
        return sc * target;
    }
};
struct Cplx {
    using S = C;
    static constexpr int W = 2;
    __device__ static S ld(double const* p) { return {p[0], p[1]}; }
    __device__ static void st(double* p, S v) {
        p[0] = v.r;
        p[1] = v.i;
    }
    __device__ static S zero() { return {0.0, 0.0}; }
    __device__ static S mul(S a, S b) { return res::cmul(a, b); }
    __device__ static S add(S a, S b) { return res::cadd(a, b); }
    __device__ static S sub(S a, S b) { return res::csub(a, b); }
    __device__ static S div(S a, S b) { return res::cdiv(a, b); }
    __device__ static double abs2(S a) { return a.r * a.r + a.i * a.i; }
    __device__ static double cabs(S a) { return hypot(a.r, a.i); }
    // is_normal of a complex value (three_phase_tensor.hpp:380-392): both parts normal, or one normal and the other zero
    __device__ static bool is_normal(S a) {
        bool const nr = fabs(a.r) >= DBL_MIN && !isinf(a.r), ni = fabs(a.i) >= DBL_MIN && !isinf(a.i);
        if (a.i == 0.0) return nr;
        if (a.r == 0.0) return ni;
        return nr && ni;
    }
    __device__ static S scale_to(S v, double abs_v, double target) {
        if (abs_v == 0.0) return {target, 0.0};
        S const sc{v.r / abs_v, v.i / abs_v};
        return {sc.r * target, sc.i * target};
    }
};

struct LuDev {
    int64_t n, nnz;
    int64_t const* indptr;
    int64_t const* indices;
    int64_t const* diag;
};

template <class Tr, int N> struct Sys {
    using S = typename Tr::S;
    static constexpr int NN = N * N, W = Tr::W;
    LuDev s;
    double* lu;          // [nnz][NN][W]   factorised in place
    double const* orig;  // original matrix (refinement) -- the caller's input
    int8_t* perm;        // [n][2][N]
    int64_t* colpos;     // [n] scratch: col_position_idx
    __device__ S get(int64_t k, int r, int c) const { return Tr::ld(lu + (k * NN + c * N + r) * W); }
    __device__ void put(int64_t k, int r, int c, S v) const { Tr::st(lu + (k * NN + c * N + r) * W, v); }
    __device__ S orig_at(int64_t k, int r, int c) const { return Tr::ld(orig + (k * NN + c * N + r) * W); }

    __device__ int64_t find_entry(int64_t col, int64_t begin, int64_t end) const {
        for (int64_t k = begin; k < end; ++k)
            if (s.indices[k] == col) return k;
        return -1;
    }

    // DenseLUFactor::factorize_block_in_place on block k; returns false when the block is singular
    __device__ bool factorize_block(int64_t k, int8_t* p, int8_t* q, double threshold, bool use_pp, bool& perturbed) const {
        int8_t rt[N], ct[N];
        double max_pivot = 0.0;
        for (int pivot = 0; pivot < N; ++pivot) {
            int rb = pivot, cb = pivot;
            double best = Tr::abs2(get(k, pivot, pivot));
            for (int c = pivot; c < N; ++c)
                for (int r = pivot; r < N; ++r) {
                    double const sc = Tr::abs2(get(k, r, c));
                    if (sc > best) {
                        best = sc;
                        rb = r;
                        cb = c;
                    }
                }
            if (best == 0.0 && !use_pp) {
                for (int rest = pivot; rest < N; ++rest) {
                    rt[rest] = (int8_t)rest;
                    ct[rest] = (int8_t)rest;
                }
                break;
            }
            double abs_pivot = sqrt(best);
            if (abs_pivot < threshold) { // perturb_pivot_if_needed: keeps the phase of the pivot
                put(k, rb, cb, Tr::scale_to(get(k, rb, cb), abs_pivot, threshold));
                perturbed = true;
                abs_pivot = threshold;
            }
            max_pivot = fmax(max_pivot, abs_pivot);
            rt[pivot] = (int8_t)rb;
            ct[pivot] = (int8_t)cb;
            if (rb != pivot)
                for (int c = 0; c < N; ++c) {
                    S const x = get(k, pivot, c);
                    put(k, pivot, c, get(k, rb, c));
                    put(k, rb, c, x);
                }
            if (cb != pivot)
                for (int r = 0; r < N; ++r) {
                    S const x = get(k, r, pivot);
                    put(k, r, pivot, get(k, r, cb));
                    put(k, r, cb, x);
                }
            if (pivot < N - 1) {
                S const pv = get(k, pivot, pivot);
                for (int r = pivot + 1; r < N; ++r) put(k, r, pivot, Tr::div(get(k, r, pivot), pv));
                for (int c = pivot + 1; c < N; ++c)
                    for (int r = pivot + 1; r < N; ++r)
                        put(k, r, c, Tr::sub(get(k, r, c), Tr::mul(get(k, r, pivot), get(k, pivot, c))));
            }
        }
        for (int i = 0; i < N; ++i) {
            p[i] = (int8_t)i;
            q[i] = (int8_t)i;
        }
        for (int pivot = N - 1; pivot >= 0; --pivot) {
            int8_t const x = p[pivot];
            p[pivot] = p[rt[pivot]];
            p[rt[pivot]] = x;
        }
        for (int pivot = 0; pivot < N; ++pivot) {
            int8_t const x = q[pivot];
            q[pivot] = q[ct[pivot]];
            q[ct[pivot]] = x;
        }
        double const pivot_threshold = perturbed ? 0.0 : DBL_EPSILON * max_pivot;
        for (int pivot = 0; pivot < N; ++pivot) {
            S const d = get(k, pivot, pivot);
            if (Tr::cabs(d) < pivot_threshold || !Tr::is_normal(d)) return false;
        }
        return true;
    }
    __device__ void perm_rows(int64_t k, int8_t const* p) const { // row p[i] of the result = row i
        S tmp[NN];
        for (int c = 0; c < N; ++c)
            for (int r = 0; r < N; ++r) tmp[c * N + r] = get(k, r, c);
        for (int c = 0; c < N; ++c)
            for (int i = 0; i < N; ++i) put(k, p[i], c, tmp[c * N + i]);
    }
    __device__ void perm_cols(int64_t k, int8_t const* q) const { // column i of the result = column q[i]
        S tmp[NN];
        for (int c = 0; c < N; ++c)
            for (int r = 0; r < N; ++r) tmp[c * N + r] = get(k, r, c);
        for (int i = 0; i < N; ++i)
            for (int r = 0; r < N; ++r) put(k, r, i, tmp[q[i] * N + r]);
    }

    // SparseLUSolver::prefactorize; returns false when singular
    __device__ bool prefactorize(bool use_pp, double threshold, bool& perturbed) const {
        for (int64_t i = 0; i < s.n; ++i) colpos[i] = s.indptr[i];
        for (int64_t piv = 0; piv < s.n; ++piv) {
            int64_t const pk = s.diag[piv];
            int8_t* const p = perm + piv * 2 * N;
            int8_t* const q = p + N;
            if constexpr (N > 1) {
                if (!factorize_block(pk, p, q, threshold, use_pp, perturbed)) return false;
                for (int64_t l = s.indptr[piv]; l < pk; ++l) { // finished L blocks of this row, U blocks of this column
                    perm_rows(l, p);
                    int64_t const u_row = s.indices[l];
                    perm_cols(colpos[u_row], q);
                    ++colpos[u_row];
                }
                for (int64_t u = pk + 1; u < s.indptr[piv + 1]; ++u) { // U = L_pp^-1 P A
                    perm_rows(u, p);
                    for (int idx = 0; idx < N; ++idx)
                        for (int prev = 0; prev < idx; ++prev)
                            for (int c = 0; c < N; ++c) put(u, idx, c, Tr::sub(get(u, idx, c), Tr::mul(get(pk, idx, prev), get(u, prev, c))));
                }
            } else {
                S v = get(pk, 0, 0);
                if (use_pp) {
                    double const a = Tr::cabs(v);
                    if (a < threshold) {
                        v = Tr::scale_to(v, a, threshold);
                        put(pk, 0, 0, v);
                        perturbed = true;
                    }
                }
                if (!Tr::is_normal(v)) return false;
                p[0] = 0;
                q[0] = 0;
            }
            for (int64_t lr = pk + 1; lr < s.indptr[piv + 1]; ++lr) { // L blocks below the pivot + Schur complement
                int64_t const l_row = s.indices[lr];
                int64_t const l = colpos[l_row];
                if constexpr (N > 1) {
                    perm_cols(l, q);
                    for (int idx = 0; idx < N; ++idx) { // right / upper solve
                        for (int prev = 0; prev < idx; ++prev)
                            for (int r = 0; r < N; ++r) put(l, r, idx, Tr::sub(get(l, r, idx), Tr::mul(get(pk, prev, idx), get(l, r, prev))));
                        for (int r = 0; r < N; ++r) put(l, r, idx, Tr::div(get(l, r, idx), get(pk, idx, idx)));
                    }
                } else {
                    put(l, 0, 0, Tr::div(get(l, 0, 0), get(pk, 0, 0)));
                }
                int64_t a = l;
                for (int64_t u = pk + 1; u < s.indptr[piv + 1]; ++u) {
                    a = find_entry(s.indices[u], a + 1, s.indptr[l_row + 1]);
                    if (a < 0) return false; // pattern not closed under fill-in
                    for (int c = 0; c < N; ++c)
                        for (int r = 0; r < N; ++r) {
                            S sum = Tr::mul(get(l, r, 0), get(u, 0, c));
                            for (int k = 1; k < N; ++k) sum = Tr::add(sum, Tr::mul(get(l, r, k), get(u, k, c)));
                            put(a, r, c, Tr::sub(get(a, r, c), sum));
                        }
                }
                ++colpos[l_row];
            }
            ++colpos[piv];
        }
        return true;
    }

    // solve_once: x = (LU)^-1 rhs; rhs may alias x
    __device__ void solve_once(double const* rhs, double* x) const {
        auto X = [&](int64_t row, int i) { return Tr::ld(x + (row * N + i) * W); };
        auto setX = [&](int64_t row, int i, S v) { Tr::st(x + (row * N + i) * W, v); };
        auto sub_dot = [&](int64_t row, int64_t k, int64_t col) {
            for (int r = 0; r < N; ++r) {
                S sum = Tr::mul(get(k, r, 0), X(col, 0));
                for (int j = 1; j < N; ++j) sum = Tr::add(sum, Tr::mul(get(k, r, j), X(col, j)));
                setX(row, r, Tr::sub(X(row, r), sum));
            }
        };
        for (int64_t row = 0; row < s.n; ++row) {
            int8_t const* const p = perm + row * 2 * N;
            S tmp[N];
            for (int i = 0; i < N; ++i) tmp[i] = Tr::ld(rhs + (row * N + i) * W);
            for (int i = 0; i < N; ++i) setX(row, N > 1 ? p[i] : 0, tmp[i]);
            for (int64_t l = s.indptr[row]; l < s.diag[row]; ++l) sub_dot(row, l, s.indices[l]);
            if constexpr (N > 1) {
                int64_t const dk = s.diag[row];
                for (int idx = 0; idx < N; ++idx)
                    for (int prev = 0; prev < idx; ++prev) setX(row, idx, Tr::sub(X(row, idx), Tr::mul(get(dk, idx, prev), X(row, prev))));
            }
        }
        for (int64_t row = s.n - 1; row >= 0; --row) {
            int64_t const dk = s.diag[row];
            for (int64_t u = s.indptr[row + 1] - 1; u > dk; --u) sub_dot(row, u, s.indices[u]);
            for (int step = 0; step < N; ++step) {
                int const idx = N - 1 - step;
                for (int ps = 0; ps < step; ++ps) {
                    int const prev = N - 1 - ps;
                    setX(row, idx, Tr::sub(X(row, idx), Tr::mul(get(dk, idx, prev), X(row, prev))));
                }
                setX(row, idx, Tr::div(X(row, idx), get(dk, idx, idx)));
            }
        }
        if constexpr (N > 1) {
            for (int64_t row = 0; row < s.n; ++row) {
                int8_t const* const q = perm + row * 2 * N + N;
                S tmp[N];
                for (int i = 0; i < N; ++i) tmp[i] = X(row, i);
                for (int i = 0; i < N; ++i) setX(row, q[i], tmp[i]);
            }
        }
    }
};

// one thread = one system of the batch
template <class Tr, int N>
__global__ void sparse_lu_batch_kernel(LuDev s, int64_t n_batch, double const* data, double const* rhs, int use_pp, double* lu,
                                       int8_t* perm, int64_t* colpos, double* x, double* residual, double* dx, double* denom,
                                       int32_t* status, int32_t* perturbed_out, int32_t* n_solves_out) {
    using S = typename Tr::S;
    constexpr int NN = N * N, W = Tr::W;
    int64_t const b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batch) return;
    int64_t const vec = s.n * N * W;
    Sys<Tr, N> sys{s, lu + b * s.nnz * NN * W, data + b * s.nnz * NN * W, perm + b * s.n * 2 * N, colpos + b * s.n};
    for (int64_t i = 0; i < s.nnz * NN * W; ++i) sys.lu[i] = sys.orig[i];
    double const* const my_rhs = rhs + b * vec;
    double* const my_x = x + b * vec;
    // initialize_pivot_perturbation: max over rows of the sum of the off-diagonal blocks' infinity norms
    double matrix_norm = 0.0;
    if (use_pp != 0) {
        for (int64_t row = 0; row < s.n; ++row) {
            double row_norm = 0.0;
            for (int64_t k = s.indptr[row]; k < s.indptr[row + 1]; ++k) {
                if (s.indices[k] == row) continue;
                double block_norm = 0.0;
                for (int r = 0; r < N; ++r) {
                    double sum = 0.0;
                    for (int c = 0; c < N; ++c) sum += Tr::cabs(sys.orig_at(k, r, c));
                    block_norm = r == 0 ? sum : fmax(block_norm, sum);
                }
                row_norm += block_norm;
            }
            matrix_norm = fmax(matrix_norm, row_norm);
        }
    }
    double const threshold = kEpsPerturbation * matrix_norm;
    bool perturbed = false;
    int n_solves = 0;
    bool ok = sys.prefactorize(use_pp != 0, threshold, perturbed);
    if (ok) {
        if (!perturbed) {
            sys.solve_once(my_rhs, my_x);
            n_solves = 1;
        } else { // solve_with_refinement
            double* const res_v = residual + b * vec;
            double* const dx_v = dx + b * vec;
            double* const den = denom + b * s.n * N;
            for (int64_t i = 0; i < vec; ++i) {
                my_x[i] = 0.0;
                res_v[i] = my_rhs[i];
                dx_v[i] = 0.0;
            }
            double backward_error = DBL_MAX;
            int num_iter = 0;
            while (backward_error > kEpsPerturbation) {
                if (num_iter++ == kMaxRefinement + 1) {
                    ok = false;
                    break;
                }
                sys.solve_once(res_v, dx_v);
                ++n_solves;
                // iterate_and_backward_error
                double max_den = 0.0;
                for (int64_t row = 0; row < s.n; ++row) {
                    for (int r = 0; r < N; ++r) den[row * N + r] = Tr::cabs(Tr::ld(my_rhs + (row * N + r) * W));
                    for (int64_t k = s.indptr[row]; k < s.indptr[row + 1]; ++k) {
                        int64_t const col = s.indices[k];
                        for (int r = 0; r < N; ++r) {
                            double sum = Tr::cabs(sys.orig_at(k, r, 0)) * Tr::cabs(Tr::ld(my_x + (col * N) * W));
                            for (int j = 1; j < N; ++j) sum += Tr::cabs(sys.orig_at(k, r, j)) * Tr::cabs(Tr::ld(my_x + (col * N + j) * W));
                            den[row * N + r] += sum;
                        }
                    }
                    for (int r = 0; r < N; ++r) max_den = fmax(max_den, den[row * N + r]);
                }
                double const min_den = kCapBackErrDenominator * max_den;
                double max_berr = 0.0;
                for (int64_t i = 0; i < s.n * N; ++i) {
                    double const d = fmax(den[i], min_den);
                    max_berr = fmax(max_berr, Tr::cabs(Tr::ld(res_v + i * W)) / d);
                    Tr::st(my_x + i * W, Tr::add(Tr::ld(my_x + i * W), Tr::ld(dx_v + i * W)));
                }
                backward_error = max_berr;
                // calculate_residual
                for (int64_t row = 0; row < s.n; ++row) {
                    for (int r = 0; r < N; ++r) Tr::st(res_v + (row * N + r) * W, Tr::ld(my_rhs + (row * N + r) * W));
                    for (int64_t k = s.indptr[row]; k < s.indptr[row + 1]; ++k) {
                        int64_t const col = s.indices[k];
                        for (int r = 0; r < N; ++r) {
                            S sum = Tr::mul(sys.orig_at(k, r, 0), Tr::ld(my_x + (col * N) * W));
                            for (int j = 1; j < N; ++j) sum = Tr::add(sum, Tr::mul(sys.orig_at(k, r, j), Tr::ld(my_x + (col * N + j) * W)));
                            Tr::st(res_v + (row * N + r) * W, Tr::sub(Tr::ld(res_v + (row * N + r) * W), sum));
                        }
                    }
                }
            }
        }
    }
    status[b] = ok ? 0 : 2; // PGMB_SCN_SINGULAR: SparseMatrixError
    if (perturbed_out != nullptr) perturbed_out[b] = perturbed ? 1 : 0;
    if (n_solves_out != nullptr) n_solves_out[b] = n_solves;
}

} // namespace

// host side: pattern on the device + the batch call
SparseLuBatch::SparseLuBatch(int64_t n, int64_t const* indptr, int64_t const* indices, int64_t const* diag, int block_size,
                             bool is_complex, int device)
    : n_{n}, nnz_{indptr[n]}, block_{block_size}, complex_{is_complex}, device_{device} {
    bool const supported = is_complex ? (block_size == 1 || block_size == 3) : (block_size == 1 || block_size == 2 || block_size == 3 || block_size == 6);
    if (!supported) throw InvalidArgument("sparse LU: block size / scalar type combination is not instantiated");
    for (int64_t r = 0; r != n; ++r) {
        if (diag[r] < indptr[r] || diag[r] >= indptr[r + 1] || indices[diag[r]] != r) throw InvalidArgument("sparse LU: diag_lu does not point at the diagonal entries");
    }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) throw CudaError("no CUDA device available: pgm_b200 has no CPU fallback");
    if (device < 0 || device >= n_dev) throw InvalidArgument("device index out of range");
    PGMB_CUDA(cudaSetDevice(device_));
    d_indptr_.upload(std::vector<int64_t>(indptr, indptr + n + 1), nullptr);
    d_indices_.upload(std::vector<int64_t>(indices, indices + nnz_), nullptr);
    d_diag_.upload(std::vector<int64_t>(diag, diag + n), nullptr);
    PGMB_CUDA(cudaDeviceSynchronize());
}

void SparseLuBatch::solve(int64_t n_batch, double const* data, double const* rhs, bool use_pivot_perturbation, double* x,
                          int32_t* status, int32_t* perturbed, int32_t* n_solves, double* lu_out, int8_t* perm_out) {
    if (n_batch <= 0) return;
    PGMB_CUDA(cudaSetDevice(device_));
    size_t const W = complex_ ? 2 : 1, N = block_, NN = N * N;
    size_t const mat = static_cast<size_t>(nnz_) * NN * W, vec = static_cast<size_t>(n_) * N * W;
    DevBuf<double> d_data, d_rhs, d_lu, d_x, d_res, d_dx, d_den;
    DevBuf<int8_t> d_perm;
    DevBuf<int64_t> d_colpos;
    DevBuf<int32_t> d_status, d_pert, d_solves;
    d_data.ensure(n_batch * mat);
    d_lu.ensure(n_batch * mat);
    d_rhs.ensure(n_batch * vec);
    d_x.ensure(n_batch * vec);
    d_res.ensure(n_batch * vec);
    d_dx.ensure(n_batch * vec);
    d_den.ensure(static_cast<size_t>(n_batch) * n_ * N);
    d_perm.ensure(static_cast<size_t>(n_batch) * n_ * 2 * N);
    d_colpos.ensure(static_cast<size_t>(n_batch) * n_);
    d_status.ensure(n_batch);
    d_pert.ensure(n_batch);
    d_solves.ensure(n_batch);
    PGMB_CUDA(cudaMemcpy(d_data.get(), data, n_batch * mat * sizeof(double), cudaMemcpyHostToDevice));
    PGMB_CUDA(cudaMemcpy(d_rhs.get(), rhs, n_batch * vec * sizeof(double), cudaMemcpyHostToDevice));
    LuDev const s{n_, nnz_, d_indptr_.get(), d_indices_.get(), d_diag_.get()};
    int const threads = 64;
    int const blocks = static_cast<int>((n_batch + threads - 1) / threads);
    auto launch = [&](auto kernel) {
        kernel<<<blocks, threads>>>(s, n_batch, d_data.get(), d_rhs.get(), use_pivot_perturbation ? 1 : 0, d_lu.get(), d_perm.get(),
                                    d_colpos.get(), d_x.get(), d_res.get(), d_dx.get(), d_den.get(), d_status.get(), d_pert.get(),
                                    d_solves.get());
    };
    count_kernel_launch();
    if (!complex_) {
        switch (block_) {
        case 1: launch(sparse_lu_batch_kernel<Real, 1>); break;
        case 2: launch(sparse_lu_batch_kernel<Real, 2>); break;
        case 3: launch(sparse_lu_batch_kernel<Real, 3>); break;
        default: launch(sparse_lu_batch_kernel<Real, 6>); break;
        }
    } else {
        if (block_ == 1) {
            launch(sparse_lu_batch_kernel<Cplx, 1>);
        } else {
            launch(sparse_lu_batch_kernel<Cplx, 3>);
        }
    }
    PGMB_CUDA(cudaGetLastError());
    PGMB_CUDA(cudaDeviceSynchronize());
    auto back = [&](void* host, void const* dev, size_t bytes) {
        if (host != nullptr && bytes != 0) PGMB_CUDA(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost));
    };
    back(x, d_x.get(), n_batch * vec * sizeof(double));
    back(status, d_status.get(), n_batch * sizeof(int32_t));
    back(perturbed, d_pert.get(), n_batch * sizeof(int32_t));
    back(n_solves, d_solves.get(), n_batch * sizeof(int32_t));
    back(lu_out, d_lu.get(), n_batch * mat * sizeof(double));
    back(perm_out, d_perm.get(), static_cast<size_t>(n_batch) * n_ * 2 * N);
}

} // namespace pgmb
