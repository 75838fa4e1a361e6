// Symmetric (positive-sequence) batched Newton-Raphson power flow for sm_100a.
//
// One thread block = one tile of T scenarios; threads are (slot, lane): lane = scenario inside the tile, slot = which
// matrix row of the current dependency level the thread works on.  Per NR iteration there are exactly two sweeps over
// the dependency levels of the (shared) LU pattern:
//
//   up-sweep   (row task)  build row k of the Jacobian + mismatch  (newton_raphson_pf_solver.hpp:473-547, 764-852)
//                          -> eliminate it against the finished rows c < k   (sparse_lu_solver.hpp:437-487, as IKJ)
//                          -> full-pivot LU of the 2x2 diagonal block        (sparse_lu_solver.hpp:86-165)
//                          -> U blocks right of the diagonal                  (sparse_lu_solver.hpp:420-429)
//                          -> forward substitution of the row                 (sparse_lu_solver.hpp:777-799)
//   down-sweep (row task)  backward substitution + column permutation        (sparse_lu_solver.hpp:802-826)
//                          -> polar update, |dU| per scenario                 (newton_raphson_pf_solver.hpp:325-349)
//
// The arithmetic on every matrix entry is the reference's, in the reference's order (file compiled with -fmad=false so
// nvcc does not contract a*b+c; the CPU reference build has no FMA contraction either).  L blocks are consumed in
// registers and never stored; U blocks are stored un-permuted and the column permutation Q is applied to the solution
// instead (identical products).
#include "nr_sym_common.cuh"

namespace pgmb {

using namespace nrsym;

namespace {

template <int T, Mode mode>
__device__ __forceinline__ void sweeps(DevStructure const& s, Tile<T> const& t, int slot, int n_slot, bool active,
                                       bool& singular, double& dev, unsigned long long* phase) {
    long long t0 = clock64();
    for (int lv = 0; lv < s.n_level; ++lv) {
        int const b = __ldg(s.level_ptr + lv), e = __ldg(s.level_ptr + lv + 1);
        if (active) {
            for (int i = b + slot; i < e; i += n_slot) singular |= up_row<T, mode>(s, t, __ldg(s.level_rows + i));
        }
        __syncthreads();
        if (phase != nullptr && threadIdx.x == 0) {
            long long const t1 = clock64();
            phase[lv == 0 ? 0 : 1] += (unsigned long long)(t1 - t0);
            t0 = t1;
        }
    }
    for (int lv = s.n_level - 1; lv >= 0; --lv) {
        int const b = __ldg(s.level_ptr + lv), e = __ldg(s.level_ptr + lv + 1);
        if (active) {
            for (int i = b + slot; i < e; i += n_slot) dev = fmax(dev, down_row<T, mode>(s, t, __ldg(s.level_rows + i)));
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

// grid = n_tile blocks, block = T * n_slot threads
template <int T> __global__ void nr_sym_kernel(DevStructure s, DevBatch b, SolveOptions opt) {
    __shared__ unsigned long long sh_dev[T];
    __shared__ int sh_singular[T];
    int const lane = threadIdx.x % T;
    int const slot = threadIdx.x / T;
    int const n_slot = blockDim.x / T;
    int const tile = blockIdx.x;
    int64_t const scn = (int64_t)tile * T + lane;
    bool const valid = scn < b.n_scn;

    Tile<T> t;
    t.jac = b.jac + (size_t)tile * s.nnz_lu * 4 * T + lane;
    t.xvec = b.xvec + (size_t)tile * s.n_bus * 2 * T + lane;
    t.pol = b.pol + (size_t)tile * s.n_bus * 2 * T + lane;
    t.u = b.u + (size_t)tile * s.n_bus * 2 * T + lane;
    t.perm = b.perm + (size_t)tile * s.n_bus * T + lane;
    t.sinj = b.sinj + (size_t)tile * s.n_load_gen * 2 * T + lane;
    t.usrc = b.usrc + (size_t)tile * s.n_source * 2 * T + lane;

    if (threadIdx.x < T) {
        sh_dev[threadIdx.x] = 0ull;
        sh_singular[threadIdx.x] = 0;
    }
    __syncthreads();

    bool done = !valid;
    int status = kStatusOk;
    int num_iter = 0;
    double max_dev = INFINITY;

    // initial voltages from the real-domain linear solve (newton_raphson_pf_solver.hpp:255-303)
    {
        bool singular = false;
        double dev = 0.0;
        sweeps<T, Mode::linear_init>(s, t, slot, n_slot, !done, singular, dev, b.phase_cycles ? b.phase_cycles + tile * 16 : nullptr);
        if (singular) sh_singular[lane] = 1;
        __syncthreads();
        if (!done && sh_singular[lane]) {
            status = kStatusSingular;
            done = true;
        }
    }
    // iteration driver (iterative_pf_solver.hpp:55-74): the check happens when the next iteration would start
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
        sweeps<T, Mode::newton>(s, t, slot, n_slot, !done, singular, dev, b.phase_cycles ? b.phase_cycles + tile * 16 + 4 : nullptr);
        if (!done) {
            if (singular) sh_singular[lane] = 1;
            atomicMax(&sh_dev[lane], (unsigned long long)__double_as_longlong(dev)); // dev >= 0: order-preserving
        }
        __syncthreads();
        if (!done) {
            if (sh_singular[lane]) {
                status = kStatusSingular;
                done = true;
            } else {
                max_dev = __longlong_as_double((long long)sh_dev[lane]);
                if (!(max_dev > opt.err_tol)) done = true;
            }
        }
        __syncthreads();
        if (threadIdx.x < T) sh_dev[threadIdx.x] = 0ull;
        // (the next __syncthreads_or orders this reset before the next atomicMax)
    }
    if (slot == 0 && valid) {
        b.status[scn] = status;
        b.n_iter[scn] = num_iter;
        b.max_dev[scn] = max_dev;
    }
}

// ---- layout conversion kernels ---------------------------------------------------------------------------------
// [scn][item][comp] (host layout) -> [tile][item][comp][T]
template <int T>
__global__ void to_tile_kernel(double const* __restrict__ src, double* __restrict__ dst, int64_t n_scn, int n_item,
                               int n_comp, int shared_src) {
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // over tile-layout elements
    int64_t const per_tile = (int64_t)n_item * n_comp * T;
    int64_t const n_tile = (n_scn + T - 1) / T;
    if (idx >= per_tile * n_tile) return;
    int64_t const tile = idx / per_tile;
    int64_t const r = idx % per_tile;
    int const lane = r % T;
    int64_t const ic = r / T; // item * n_comp + comp
    int64_t const scn = tile * T + lane;
    double v = 0.0;
    if (scn < n_scn) v = src[(shared_src ? 0 : scn * (int64_t)n_item * n_comp) + ic];
    dst[idx] = v;
}

// [tile][item][comp][T] -> [scn][item][comp]

// ---- host launchers ----------------------------------------------------------------------------------------------
template <int T>
static void launch_nr_sym_t(DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot, cudaStream_t st) {
    nr_sym_kernel<T><<<b.n_tile, T * n_slot, 0, st>>>(s, b, opt);
}

void launch_nr_sym(int tile_width, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot,
                   cudaStream_t st) {
    count_kernel_launch();
    switch (tile_width) {
    case 4: launch_nr_sym_t<4>(s, b, opt, n_slot, st); break;
    case 8: launch_nr_sym_t<8>(s, b, opt, n_slot, st); break;
    case 16: launch_nr_sym_t<16>(s, b, opt, n_slot, st); break;
    default: launch_nr_sym_t<32>(s, b, opt, n_slot, st); break;
    }
}

void launch_to_tile(int tile_width, double const* src, double* dst, int64_t n_scn, int n_item, int n_comp, int shared_src,
                    cudaStream_t st) {
    count_kernel_launch();
    int64_t const n_tile = (n_scn + tile_width - 1) / tile_width;
    int64_t const total = n_tile * n_item * n_comp * tile_width;
    if (total == 0) return;
    int const block = 256;
    unsigned const grid = (unsigned)((total + block - 1) / block);
    switch (tile_width) {
    case 4: to_tile_kernel<4><<<grid, block, 0, st>>>(src, dst, n_scn, n_item, n_comp, shared_src); break;
    case 8: to_tile_kernel<8><<<grid, block, 0, st>>>(src, dst, n_scn, n_item, n_comp, shared_src); break;
    case 16: to_tile_kernel<16><<<grid, block, 0, st>>>(src, dst, n_scn, n_item, n_comp, shared_src); break;
    default: to_tile_kernel<32><<<grid, block, 0, st>>>(src, dst, n_scn, n_item, n_comp, shared_src); break;
    }
}


} // namespace pgmb
