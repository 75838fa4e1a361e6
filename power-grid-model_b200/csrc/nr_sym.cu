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
#include "kernels.cuh"

#include <cfloat>
#include <cmath>

namespace pgmb {

namespace {

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

template <int T> struct Tile {
    double* jac;
    double* xvec;
    double* pol;
    double* u;
    uint8_t* perm;
    double const* sinj;
    double const* usrc;

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
template <int T, Mode mode>
__device__ __forceinline__ bool up_row(DevStructure const& s, Tile<T> const& t, int row) {
    int const rb = __ldg(s.row_ptr + row), re = __ldg(s.row_ptr + row + 1), dg = __ldg(s.diag + row);
    double const uir = t.u[(size_t)(row * 2) * T], uii = t.u[(size_t)(row * 2 + 1) * T];
    double acc0 = 0.0, acc1 = 0.0; // NR: -P, -Q then mismatch ; linear: rhs (re, im)
    Blk d{0.0, 0.0, 0.0, 0.0};

    // 1. build the row
    for (int k = rb; k < re; ++k) {
        int const ky = __ldg(s.map_y + k);
        Blk b{0.0, 0.0, 0.0, 0.0};
        if (ky >= 0) {
            double const yr = __ldg(s.ydata + 2 * ky), yi = __ldg(s.ydata + 2 * ky + 1);
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
    for (int lg = __ldg(s.lg_ptr + row), lge = __ldg(s.lg_ptr + row + 1); lg < lge; ++lg) {
        double const ps = t.sinj[(size_t)(lg * 2) * T], qs = t.sinj[(size_t)(lg * 2 + 1) * T];
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
    double const s00 = d.a00 * d.a00, s10 = d.a10 * d.a10, s01 = d.a01 * d.a01, s11 = d.a11 * d.a11;
    double best = s00;
    int pr = 0, pcq = 0;
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
    double max_pivot = sqrt(best);
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
    max_pivot = fmax(max_pivot, sqrt(d.a11 * d.a11)); // sqrt(|x|^2) like the reference, not |x|
    double const threshold = DBL_EPSILON * max_pivot;
    singular = singular || fabs(d.a00) < threshold || not_normal(d.a00) || fabs(d.a11) < threshold || not_normal(d.a11);
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
template <int T, Mode mode> __device__ __forceinline__ double down_row(DevStructure const& s, Tile<T> const& t, int row) {
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
    double* const pth = t.pol + (size_t)(row * 2) * T;
    double* const pv = t.pol + (size_t)(row * 2 + 1) * T;
    double* const pur = t.u + (size_t)(row * 2) * T;
    double* const pui = t.u + (size_t)(row * 2 + 1) * T;
    if constexpr (mode == Mode::newton) {
        double theta = *pth, v = *pv;
        theta += y0;
        v += v * y1;
        double sn, cs;
        sincos(theta, &sn, &cs);
        double const nr = v * cs, ni = v * sn;
        double const dr = nr - *pur, di = ni - *pui;
        *pth = theta;
        *pv = v;
        *pur = nr;
        *pui = ni;
        return sqrt(dr * dr + di * di);
    } else {
        *pur = y0;
        *pui = y1;
        *pv = sqrt(y0 * y0 + y1 * y1);
        *pth = atan2(y1, y0);
        return 0.0;
    }
}

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
        sweeps<T, Mode::linear_init>(s, t, slot, n_slot, !done, singular, dev, b.phase_cycles ? b.phase_cycles + tile * 8 : nullptr);
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
        sweeps<T, Mode::newton>(s, t, slot, n_slot, !done, singular, dev, b.phase_cycles ? b.phase_cycles + tile * 8 + 4 : nullptr);
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
template <int T>
__global__ void from_tile_kernel(double const* __restrict__ src, double* __restrict__ dst, int64_t n_scn, int n_item,
                                 int n_comp) {
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // over host-layout elements
    int64_t const per_scn = (int64_t)n_item * n_comp;
    if (idx >= per_scn * n_scn) return;
    int64_t const scn = idx / per_scn;
    int64_t const ic = idx % per_scn;
    int64_t const tile = scn / T;
    int const lane = scn % T;
    dst[idx] = src[(tile * per_scn + ic) * T + lane];
}

// ---- host launchers ----------------------------------------------------------------------------------------------
template <int T>
static void launch_nr_sym_t(DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot, cudaStream_t st) {
    nr_sym_kernel<T><<<b.n_tile, T * n_slot, 0, st>>>(s, b, opt);
}

void launch_nr_sym(int tile_width, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot,
                   cudaStream_t st) {
    switch (tile_width) {
    case 4: launch_nr_sym_t<4>(s, b, opt, n_slot, st); break;
    case 8: launch_nr_sym_t<8>(s, b, opt, n_slot, st); break;
    case 16: launch_nr_sym_t<16>(s, b, opt, n_slot, st); break;
    default: launch_nr_sym_t<32>(s, b, opt, n_slot, st); break;
    }
}

void launch_to_tile(int tile_width, double const* src, double* dst, int64_t n_scn, int n_item, int n_comp, int shared_src,
                    cudaStream_t st) {
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

void launch_from_tile(int tile_width, double const* src, double* dst, int64_t n_scn, int n_item, int n_comp,
                      cudaStream_t st) {
    int64_t const total = n_scn * n_item * n_comp;
    if (total == 0) return;
    int const block = 256;
    unsigned const grid = (unsigned)((total + block - 1) / block);
    switch (tile_width) {
    case 4: from_tile_kernel<4><<<grid, block, 0, st>>>(src, dst, n_scn, n_item, n_comp); break;
    case 8: from_tile_kernel<8><<<grid, block, 0, st>>>(src, dst, n_scn, n_item, n_comp); break;
    case 16: from_tile_kernel<16><<<grid, block, 0, st>>>(src, dst, n_scn, n_item, n_comp); break;
    default: from_tile_kernel<32><<<grid, block, 0, st>>>(src, dst, n_scn, n_item, n_comp); break;
    }
}

} // namespace pgmb
