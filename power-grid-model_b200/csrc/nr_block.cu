// Batched Newton-Raphson power flow with generic (2B x 2B) Jacobian blocks -- the asymmetric three-phase calculation (B = 3,
// 6 x 6 real blocks); also instantiated for B = 1 as a cross-check of the specialised symmetric kernels.
// Same organisation as nr_sym.cu: one thread block = one tile of T scenarios, lane = scenario, slots = rows of a dependency
// level, up-sweep = build + IKJ elimination + full-pivot block LU + forward substitution, down-sweep = backward substitution
// + polar update.  Blocks are column-major like the reference's Eigen blocks; sub-blocks H (0,0) N (0,1) M (1,0) L (1,1).
//   newton_raphson_pf_solver.hpp:255-303 (linear start), 462-547, 764-852 (Jacobian + mismatch), 325-349 (update)
//   sparse_lu_solver.hpp:86-165 (full-pivot block LU, first maximum in column-major order), 171-200 (triangular solves),
//                        346-495 (prefactorize), 769-827 (solve_once)
// A thread keeps its working blocks in (lane-interleaved, L1-cached) local memory: the pivot permutations index them
// dynamically.  U blocks are stored with the row permutation / L^-1 applied but WITHOUT the later column permutation Q_j;
// the backward substitution sums the products in the permuted order instead, so every rounding equals the reference's.
#include "kernels.cuh"

#include <stdexcept>
#include "block_common.cuh"

namespace pgmb {
using namespace blk;
namespace {

template <int T, int B, Mode mode, bool REG>
__device__ void sweeps(DevStructure const& s, TileB<T, B> const& t, int slot, int n_slot, bool active, bool& singular,
                       double& dev, unsigned long long* phase, bool check_now) {
    long long t0 = clock64();
    auto lap = [&](int k) { // PGMB_DEBUG_PHASES: up level 0 | up wide rows | up other levels | down levels >= 1 | down level 0
        if (phase != nullptr && threadIdx.x == 0) {
            long long const t1 = clock64();
            phase[k] += (unsigned long long)(t1 - t0);
            t0 = t1;
        }
    };
    for (int lv = 0; lv < s.n_level; ++lv) {
        int const b = __ldg(s.level_ptr + lv), e = __ldg(s.level_ptr + lv + 1);
        if (active)
            for (int i = b + slot; i < e; i += n_slot) {
                int const row = __ldg(s.level_rows + i);
                if (s.n_wide != 0 && __ldg(s.row_is_wide + row)) continue; // eliminated below by the whole block
                singular |= up_row<T, B, mode, false, REG>(s, t, row, check_now);
            }
        __syncthreads();
        lap(lv == 0 ? 0 : 2);
        if (s.n_wide != 0)
            for (int w = __ldg(s.wide_level_ptr + lv); w < __ldg(s.wide_level_ptr + lv + 1); ++w)
                wide_up_row<T, B, mode, false, REG>(s, t, w, slot, n_slot, active, singular, check_now);
        lap(1);
    }
    for (int lv = s.n_level - 1; lv >= 0; --lv) {
        int const b = __ldg(s.level_ptr + lv), e = __ldg(s.level_ptr + lv + 1);
        if (active)
            for (int i = b + slot; i < e; i += n_slot) dev = fmax(dev, down_row<T, B, mode, false, REG>(s, t, __ldg(s.level_rows + i)));
        __syncthreads();
        lap(lv == 0 ? 4 : 3);
    }
}

// REG: grids with voltage regulators (PV buses, Q limits); the plain instantiation carries none of that code
template <int T, int B, bool REG> __global__ void nr_block_kernel(DevStructure s, DevBatch b, SolveOptions opt) {
    constexpr int N = 2 * B;
    __shared__ unsigned long long sh_dev[T];
    __shared__ int sh_singular[T];
    __shared__ int sh_has_limits[T]; // REG: the scenario has a PV bus with a usable Q limit (limit check at iteration 2)
    int const lane = threadIdx.x % T, slot = threadIdx.x / T, n_slot = blockDim.x / T, tile = blockIdx.x;
    int64_t const scn = (int64_t)tile * T + lane;
    bool const valid = scn < b.n_scn;
    TileB<T, B> t;
    t.jac = b.jac + (size_t)tile * s.nnz_lu * N * N * T + lane;
    t.xvec = b.xvec + (size_t)tile * s.n_bus * N * T + lane;
    t.pol = b.pol + (size_t)tile * s.n_bus * N * T + lane;
    t.u = b.u + (size_t)tile * s.n_bus * N * T + lane;
    t.perm = b.perm + (size_t)tile * s.n_bus * 2 * N * T + lane;
    t.sinj = b.sinj + (size_t)tile * s.n_load_gen * N * T + lane;
    t.usrc = b.usrc + (size_t)tile * s.n_source * 2 * T + lane;
    t.wide_terms = b.wide_terms ? b.wide_terms + (size_t)tile * s.wide_max_upd * N * N * T + lane : nullptr;
    t.wide_rhs = b.wide_rhs ? b.wide_rhs + (size_t)tile * s.wide_max_lower * N * T + lane : nullptr;
    t.wide_sum = b.wide_sum ? b.wide_sum + (size_t)tile * s.wide_max_entries * N * T + lane : nullptr;
    t.lg_status = b.lg_status ? b.lg_status + (size_t)tile * s.n_load_gen * T + lane : nullptr;
    t.qviol = b.qviol ? b.qviol + (size_t)tile * s.n_bus * T + lane : nullptr;
    t.ovr_n = 4 * b.ovl.n_branch;
    t.ovr_entry = (b.ovl.entry != nullptr && valid) ? b.ovl.entry + scn * t.ovr_n : nullptr;
    t.ovr_y = (b.ovl.entry != nullptr && valid) ? b.ovl.y + scn * t.ovr_n * (2 * B * B) : nullptr;
    t.dead = (b.ovl.dead_off != nullptr && valid && b.ovl.dead_off[scn] >= 0) ? b.ovl.dead + (size_t)b.ovl.dead_off[scn] * s.n_bus : nullptr;
    if (threadIdx.x < T) {
        sh_dev[threadIdx.x] = 0ull;
        sh_singular[threadIdx.x] = 0;
        sh_has_limits[threadIdx.x] = 0;
    }
    __syncthreads();
    if constexpr (REG) { // set_bus_types_and_q_limits (newton_raphson_pf_solver.hpp:400-444); no limit has been hit yet
        if (valid)
            for (int row = slot; row < s.n_bus; row += n_slot) {
                t.qviol[(size_t)row * T] = 0;
                if (__ldg(s.lg_ptr + row) != __ldg(s.lg_ptr + row + 1) && bus_control<T, B, false>(s, t, row).has_limits)
                    sh_has_limits[lane] = 1;
            }
        __syncthreads();
    }
    unsigned long long* const phase = b.phase_cycles ? b.phase_cycles + tile * 16 : nullptr;
    bool done = !valid;
    int status = kStatusOk, num_iter = 0;
    double max_dev = INFINITY;
    {
        bool singular = false;
        double dev = 0.0;
        sweeps<T, B, Mode::linear_init, REG>(s, t, slot, n_slot, !done, singular, dev, phase, false);
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
        sweeps<T, B, Mode::newton, REG>(s, t, slot, n_slot, !done, singular, dev, phase ? phase + 8 : nullptr, REG && num_iter >= 2);
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

} // namespace

template <int B, bool REG>
static void launch_nr_block_b(int tw, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot,
                              cudaStream_t st) {
    switch (tw) {
    case 4: nr_block_kernel<4, B, REG><<<b.n_tile, 4 * n_slot, 0, st>>>(s, b, opt); break;
    case 8: nr_block_kernel<8, B, REG><<<b.n_tile, 8 * n_slot, 0, st>>>(s, b, opt); break;
    case 16: nr_block_kernel<16, B, REG><<<b.n_tile, 16 * n_slot, 0, st>>>(s, b, opt); break;
    default: nr_block_kernel<32, B, REG><<<b.n_tile, 32 * n_slot, 0, st>>>(s, b, opt); break;
    }
}

void launch_nr_block(int phases, int tw, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot,
                     cudaStream_t st) {
    count_kernel_launch();
    bool const reg = s.lg_reg != nullptr;
    if (reg && (b.qviol == nullptr || b.lg_status == nullptr)) {
        throw std::logic_error("nr_block: a grid with voltage regulators needs the qviol / lg_status buffers of the batch");
    }
    if (phases == 1) {
        reg ? launch_nr_block_b<1, true>(tw, s, b, opt, n_slot, st) : launch_nr_block_b<1, false>(tw, s, b, opt, n_slot, st);
    } else {
        reg ? launch_nr_block_b<3, true>(tw, s, b, opt, n_slot, st) : launch_nr_block_b<3, false>(tw, s, b, opt, n_slot, st);
    }
}

} // namespace pgmb
