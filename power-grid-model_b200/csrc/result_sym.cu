// Result extraction for the symmetric calculation at the math-solver seam (SolverOutput<symmetric_t>):
//   YBus::calculate_branch_flow / calculate_shunt_flow / calculate_injection   (math_solver/y_bus.hpp:482-546)
//   calculate_load_gen_result / calculate_source_result / calculate_multiple_source_result
//                                                       (math_solver/common_solver_functions.hpp:83-101, 143-160, 383-409)
// One thread per (scenario, element); output is scenario-major like the caller's host buffers, so consecutive threads
// write consecutive elements.  Bus voltages are read from the tile layout written by the solver kernel.
#include "kernels.cuh"

#include <cuComplex.h>

namespace pgmb {
namespace {

struct C {
    double r, i;
};
__device__ __forceinline__ C cmul(C a, C b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
__device__ __forceinline__ C cadd(C a, C b) { return {a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ C csub(C a, C b) { return {a.r - b.r, a.i - b.i}; }
__device__ __forceinline__ C conj(C a) { return {a.r, -a.i}; }
__device__ __forceinline__ C cscale(C a, double s) { return {a.r * s, a.i * s}; }
// complex division as libgcc's __divdc3 performs it for finite, well-scaled operands (Smith's algorithm)
__device__ __forceinline__ C cdiv(C x, C y) {
    double ratio, denom;
    C out;
    if (fabs(y.r) < fabs(y.i)) {
        ratio = y.r / y.i;
        denom = (y.r * ratio) + y.i;
        out.r = ((x.r * ratio) + x.i) / denom;
        out.i = ((x.i * ratio) - x.r) / denom;
    } else {
        ratio = y.i / y.r;
        denom = (y.i * ratio) + y.r;
        out.r = ((x.i * ratio) + x.r) / denom;
        out.i = (x.i - (x.r * ratio)) / denom;
    }
    return out;
}
__device__ __forceinline__ C ldc(double const* p, int64_t k) { return {__ldg(p + 2 * k), __ldg(p + 2 * k + 1)}; }

template <int T> struct UView {
    double const* u;
    int n_bus;
    __device__ __forceinline__ C get(int64_t scn, int bus) const {
        int64_t const tile = scn / T;
        int const lane = scn % T;
        double const* p = u + ((tile * n_bus + bus) * 2) * T + lane;
        return {p[0], p[T]};
    }
};

template <int T> __device__ __forceinline__ C bus_injection(DevStructure const& s, UView<T> const& uv, int64_t scn, int bus) {
    C i_inj{0.0, 0.0};
    for (int k = __ldg(s.y_row_ptr + bus), ke = __ldg(s.y_row_ptr + bus + 1); k < ke; ++k) {
        i_inj = cadd(i_inj, cmul(ldc(s.ydata, k), uv.get(scn, __ldg(s.y_col_idx + k))));
    }
    return cmul(conj(i_inj), uv.get(scn, bus));
}

template <int T>
__device__ __forceinline__ C load_gen_s(DevStructure const& s, double const* sinj, int64_t scn, int lg, C u, int type) {
    int64_t const tile = scn / T;
    int const lane = scn % T;
    double const* p = sinj + ((tile * s.n_load_gen + lg) * 2) * T + lane;
    C const sv{p[0], p[T]};
    if (type == 0) return sv;
    if (type == 1) return cscale(sv, u.r * u.r + u.i * u.i);
    return cscale(sv, sqrt(u.r * u.r + u.i * u.i));
}

template <int T>
__global__ void math_result_sym_kernel(DevStructure s, DevBatch b, int force_const_y, double* out_u, double* out_inj,
                                       double* out_branch, double* out_source, double* out_shunt, double* out_lg) {
    UView<T> const uv{b.u, s.n_bus};
    int64_t const per_scn = (int64_t)s.n_bus * 2 + s.n_branch + s.n_shunt + s.n_load_gen + s.n_source;
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_scn * b.n_scn) return;
    int64_t const scn = idx / per_scn;
    int64_t r = idx % per_scn;
    if (r < s.n_bus) { // u
        if (out_u != nullptr) {
            C const u = uv.get(scn, (int)r);
            out_u[(scn * s.n_bus + r) * 2] = u.r;
            out_u[(scn * s.n_bus + r) * 2 + 1] = u.i;
        }
        return;
    }
    r -= s.n_bus;
    if (r < s.n_bus) { // bus injection
        if (out_inj != nullptr) {
            C const v = bus_injection<T>(s, uv, scn, (int)r);
            out_inj[(scn * s.n_bus + r) * 2] = v.r;
            out_inj[(scn * s.n_bus + r) * 2 + 1] = v.i;
        }
        return;
    }
    r -= s.n_bus;
    if (r < s.n_branch) {
        if (out_branch != nullptr) {
            int const f = __ldg(s.branch_bus + 2 * r), t = __ldg(s.branch_bus + 2 * r + 1);
            C const uf = f >= 0 ? uv.get(scn, f) : C{0.0, 0.0};
            C const ut = t >= 0 ? uv.get(scn, t) : C{0.0, 0.0};
            C const i_f = cadd(cmul(ldc(s.branch_param, r * 4 + 0), uf), cmul(ldc(s.branch_param, r * 4 + 1), ut));
            C const i_t = cadd(cmul(ldc(s.branch_param, r * 4 + 2), uf), cmul(ldc(s.branch_param, r * 4 + 3), ut));
            C const s_f = cmul(uf, conj(i_f));
            C const s_t = cmul(ut, conj(i_t));
            double* o = out_branch + (scn * s.n_branch + r) * 8;
            o[0] = s_f.r, o[1] = s_f.i, o[2] = s_t.r, o[3] = s_t.i, o[4] = i_f.r, o[5] = i_f.i, o[6] = i_t.r, o[7] = i_t.i;
        }
        return;
    }
    r -= s.n_branch;
    if (r < s.n_shunt) {
        if (out_shunt != nullptr) {
            C const u = uv.get(scn, __ldg(s.shunt_bus + r));
            C const yu = cmul(ldc(s.shunt_param, r), u);
            C const i{-yu.r, -yu.i};
            C const sv = cmul(u, conj(i));
            double* o = out_shunt + (scn * s.n_shunt + r) * 4;
            o[0] = sv.r, o[1] = sv.i, o[2] = i.r, o[3] = i.i;
        }
        return;
    }
    r -= s.n_shunt;
    if (r < s.n_load_gen) {
        if (out_lg != nullptr) {
            C const u = uv.get(scn, __ldg(s.lg_bus + r));
            C const sv = load_gen_s<T>(s, b.sinj, scn, (int)r, u, force_const_y ? 1 : __ldg(s.lg_type + r));
            C const i = conj(cdiv(sv, u));
            double* o = out_lg + (scn * s.n_load_gen + r) * 4;
            o[0] = sv.r, o[1] = sv.i, o[2] = i.r, o[3] = i.i;
        }
        return;
    }
    r -= s.n_load_gen;
    if (out_source == nullptr) return;
    // source r
    int const bus = __ldg(s.src_bus + r);
    C const u = uv.get(scn, bus);
    C i_lg{0.0, 0.0};
    for (int lg = __ldg(s.lg_ptr + bus), lge = __ldg(s.lg_ptr + bus + 1); lg < lge; ++lg) {
        C const sv = load_gen_s<T>(s, b.sinj, scn, lg, u, force_const_y ? 1 : __ldg(s.lg_type + lg));
        i_lg = cadd(i_lg, conj(cdiv(sv, u)));
    }
    C const i_inj_t = csub(conj(cdiv(bus_injection<T>(s, uv, scn, bus), u)), i_lg);
    int const sb = __ldg(s.src_ptr + bus), se = __ldg(s.src_ptr + bus + 1);
    C i_src;
    if (se - sb == 1) {
        i_src = i_inj_t;
    } else {
        int64_t const tile = scn / T;
        int const lane = scn % T;
        C y_ref_t{0.0, 0.0}, i_ref_t{0.0, 0.0};
        for (int k = sb; k < se; ++k) y_ref_t = cadd(y_ref_t, ldc(s.src_y1y0, 2 * k));
        C const z_ref_t = cdiv(C{1.0, 0.0}, y_ref_t);
        for (int k = sb; k < se; ++k) {
            double const* p = b.usrc + ((tile * s.n_source + k) * 2) * T + lane;
            i_ref_t = cadd(i_ref_t, cmul(C{p[0], p[T]}, ldc(s.src_y1y0, 2 * k)));
        }
        double const* p = b.usrc + ((tile * s.n_source + r) * 2) * T + lane;
        C const ratio = cmul(ldc(s.src_y1y0, 2 * r), z_ref_t);
        C const lhs = cmul(ratio, csub(cmul(C{p[0], p[T]}, y_ref_t), i_ref_t));
        i_src = cadd(lhs, cmul(ratio, i_inj_t));
    }
    C const sv = cmul(u, conj(i_src));
    double* o = out_source + (scn * s.n_source + r) * 4;
    o[0] = sv.r, o[1] = sv.i, o[2] = i_src.r, o[3] = i_src.i;
}

} // namespace

void launch_math_result_sym(int tile_width, DevStructure const& s, DevBatch const& b, int force_const_y, double* out_u,
                            double* out_inj, double* out_branch, double* out_source, double* out_shunt, double* out_lg,
                            cudaStream_t st) {
    int64_t const per_scn = (int64_t)s.n_bus * 2 + s.n_branch + s.n_shunt + s.n_load_gen + s.n_source;
    int64_t const total = per_scn * b.n_scn;
    if (total == 0) return;
    int const block = 256;
    unsigned const grid = (unsigned)((total + block - 1) / block);
#define PGMB_LAUNCH(TW) \
    math_result_sym_kernel<TW><<<grid, block, 0, st>>>(s, b, force_const_y, out_u, out_inj, out_branch, out_source, out_shunt, out_lg)
    switch (tile_width) {
    case 4: PGMB_LAUNCH(4); break;
    case 8: PGMB_LAUNCH(8); break;
    case 16: PGMB_LAUNCH(16); break;
    default: PGMB_LAUNCH(32); break;
    }
#undef PGMB_LAUNCH
}

} // namespace pgmb
