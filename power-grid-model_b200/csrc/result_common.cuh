// Device helpers shared by the result-extraction and output kernels: complex arithmetic as the reference's CPU build
// performs it (std::complex operator* / libgcc __divdc3 for well-scaled operands), tile-layout voltage access, and the
// per-element result formulas of YBus::calculate_injection (math_solver/y_bus.hpp:482-496) and
// calculate_load_gen_result / calculate_source_result (math_solver/common_solver_functions.hpp:83-101, 143-160, 383-409).
#pragma once

#include "kernels.cuh"

namespace pgmb {
namespace res {

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

// value block of Y-bus entry k for scenario scn: the shared admittance or the scenario's branch-outage replacement
__device__ __forceinline__ double const* y_entry(DevStructure const& s, DevOverlay const& o, int64_t scn, int k, int bb2) {
    if (o.entry != nullptr) {
        int const n = 4 * o.n_branch;
        for (int j = 0; j < n; ++j)
            if (__ldg(o.entry + scn * n + j) == k) return o.y + (scn * n + j) * bb2;
    }
    return s.ydata + (size_t)k * bb2;
}
__device__ __forceinline__ bool bus_is_dead(DevOverlay const& o, int64_t scn, int bus, int n_bus) {
    if (o.dead_off == nullptr) return false;
    int32_t const off = __ldg(o.dead_off + scn);
    return off >= 0 && __ldg(o.dead + (size_t)off * n_bus + bus) != 0;
}
// parameters of math branch r for scenario scn
__device__ __forceinline__ double const* branch_param_of(DevStructure const& s, DevOverlay const& o, int64_t scn, int64_t r, int bb2) {
    if (o.branch != nullptr) {
        for (int j = 0; j < o.n_branch; ++j)
            if (__ldg(o.branch + scn * o.n_branch + j) == r) return o.bparam + (scn * o.n_branch + j) * 4 * bb2;
    }
    return s.branch_param + (size_t)r * 4 * bb2;
}
// `energized` flag of branch component `comp` in scenario scn when the scenario switches it, else `base`
__device__ __forceinline__ int branch_energized_of(DevOverlay const& o, int64_t scn, int comp, int base) {
    if (o.comp != nullptr) {
        for (int j = 0; j < o.n_branch; ++j)
            if (__ldg(o.comp + scn * o.n_branch + j) == comp) return __ldg(o.energized + scn * o.n_branch + j);
    }
    return base;
}

template <int T>
__device__ __forceinline__ C bus_injection(DevStructure const& s, UView<T> const& uv, int64_t scn, int bus, DevOverlay const& ovl) {
    C i_inj{0.0, 0.0};
    for (int k = __ldg(s.y_row_ptr + bus), ke = __ldg(s.y_row_ptr + bus + 1); k < ke; ++k) {
        i_inj = cadd(i_inj, cmul(ldc(y_entry(s, ovl, scn, k, 2), 0), uv.get(scn, __ldg(s.y_col_idx + k))));
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

// current and power of source r (single source: all the bus injection that is not load current; several sources: split by
// admittance share, calculate_multiple_source_result, common_solver_functions.hpp:83-101)
template <int T>
__device__ __forceinline__ void source_result(DevStructure const& s, DevBatch const& b, UView<T> const& uv, int64_t scn, int r,
                                              int force_const_y, C& sv, C& i_src_out) {
    int const bus = __ldg(s.src_bus + r);
    C const u = uv.get(scn, bus);
    C i_lg{0.0, 0.0};
    for (int lg = __ldg(s.lg_ptr + bus), lge = __ldg(s.lg_ptr + bus + 1); lg < lge; ++lg) {
        C const sv = load_gen_s<T>(s, b.sinj, scn, lg, u, force_const_y ? 1 : __ldg(s.lg_type + lg));
        i_lg = cadd(i_lg, conj(cdiv(sv, u)));
    }
    C const i_inj_t = csub(conj(cdiv(bus_injection<T>(s, uv, scn, bus, b.ovl), u)), i_lg);
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
    sv = cmul(u, conj(i_src));
    i_src_out = i_src;
}

} // namespace res
} // namespace pgmb
