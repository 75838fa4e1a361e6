// Device helpers of the asymmetric (three-phase) result extraction, shared by result_asym.cu (math-solver seam) and
// output_asym.cu (output structs): phase vectors, 3 x 3 tensor products in the reference's summation order, and the per-element
// formulas of YBus::calculate_injection (math_solver/y_bus.hpp:482-496), calculate_load_gen_result / calculate_source_result /
// calculate_multiple_source_result (math_solver/common_solver_functions.hpp:83-139, 143-160, 383-409).
#pragma once

#include "result_common.cuh"

namespace pgmb {
namespace res3 {
using namespace res;

struct V3 {
    C v[3];
};
__device__ __forceinline__ V3 vconj(V3 a) {
    for (int p = 0; p < 3; ++p) a.v[p] = conj(a.v[p]);
    return a;
}
__device__ __forceinline__ V3 vmul(V3 a, V3 const& b) {
    for (int p = 0; p < 3; ++p) a.v[p] = cmul(a.v[p], b.v[p]);
    return a;
}
__device__ __forceinline__ V3 vdiv(V3 a, V3 const& b) {
    for (int p = 0; p < 3; ++p) a.v[p] = cdiv(a.v[p], b.v[p]);
    return a;
}
__device__ __forceinline__ V3 vadd(V3 a, V3 const& b) {
    for (int p = 0; p < 3; ++p) a.v[p] = cadd(a.v[p], b.v[p]);
    return a;
}
__device__ __forceinline__ V3 vsub(V3 a, V3 const& b) {
    for (int p = 0; p < 3; ++p) a.v[p] = csub(a.v[p], b.v[p]);
    return a;
}
// dot(tensor, vector) with the tensor at y[(r * 3 + c) * 2]
__device__ __forceinline__ V3 mat_vec(double const* y, V3 const& u) {
    V3 r;
    for (int i = 0; i < 3; ++i) {
        C sum = cmul(C{__ldg(y + (i * 3) * 2), __ldg(y + (i * 3) * 2 + 1)}, u.v[0]);
        for (int k = 1; k < 3; ++k) sum = cadd(sum, cmul(C{__ldg(y + (i * 3 + k) * 2), __ldg(y + (i * 3 + k) * 2 + 1)}, u.v[k]));
        r.v[i] = sum;
    }
    return r;
}
__device__ __forceinline__ V3 mat_vec_c(C const (&m)[9], V3 const& u) {
    V3 r;
    for (int i = 0; i < 3; ++i) {
        C sum = cmul(m[i * 3], u.v[0]);
        for (int k = 1; k < 3; ++k) sum = cadd(sum, cmul(m[i * 3 + k], u.v[k]));
        r.v[i] = sum;
    }
    return r;
}

template <int T> struct UView3 {
    double const* u;
    int n_bus;
    __device__ __forceinline__ V3 get(int64_t scn, int bus) const {
        int64_t const tile = scn / T;
        int const lane = scn % T;
        double const* p = u + ((tile * n_bus + bus) * 6) * T + lane;
        V3 r;
        for (int ph = 0; ph < 3; ++ph) r.v[ph] = C{p[(2 * ph) * T], p[(2 * ph + 1) * T]};
        return r;
    }
};

template <int T>
__device__ V3 bus_injection3(DevStructure const& s, UView3<T> const& uv, int64_t scn, int bus, DevOverlay const& ovl) {
    V3 i_inj{};
    for (int k = __ldg(s.y_row_ptr + bus), ke = __ldg(s.y_row_ptr + bus + 1); k < ke; ++k) {
        i_inj = vadd(i_inj, mat_vec(y_entry(s, ovl, scn, k, 18), uv.get(scn, __ldg(s.y_col_idx + k))));
    }
    return vmul(vconj(i_inj), uv.get(scn, bus));
}

template <int T> __device__ V3 load_gen_s3(DevStructure const& s, double const* sinj, int64_t scn, int lg, V3 const& u, int type) {
    int64_t const tile = scn / T;
    int const lane = scn % T;
    double const* p = sinj + ((tile * s.n_load_gen + lg) * 6) * T + lane;
    V3 sv;
    for (int ph = 0; ph < 3; ++ph) {
        C const x{p[(2 * ph) * T], p[(2 * ph + 1) * T]};
        C const uu = u.v[ph];
        if (type == 0) {
            sv.v[ph] = x;
        } else if (type == 1) {
            sv.v[ph] = cscale(x, uu.r * uu.r + uu.i * uu.i);
        } else {
            sv.v[ph] = cscale(x, sqrt(uu.r * uu.r + uu.i * uu.i));
        }
    }
    return sv;
}

__device__ __forceinline__ void store3(double* o, V3 const& v) {
    for (int p = 0; p < 3; ++p) {
        o[2 * p] = v.v[p].r;
        o[2 * p + 1] = v.v[p].i;
    }
}

// current of source r and the voltage of its bus
template <int T>
__device__ void source_result3(DevStructure const& s, DevBatch const& b, UView3<T> const& uv, int64_t scn, int r,
                               int force_const_y, V3& u_out, V3& i_src_out) {
    int const bus = __ldg(s.src_bus + r);
    V3 const u = uv.get(scn, bus);
    V3 i_lg{};
    for (int lg = __ldg(s.lg_ptr + bus), lge = __ldg(s.lg_ptr + bus + 1); lg < lge; ++lg) {
        V3 const sv = load_gen_s3<T>(s, b.sinj, scn, lg, u, force_const_y ? 1 : __ldg(s.lg_type + lg));
        i_lg = vadd(i_lg, vconj(vdiv(sv, u)));
    }
    V3 const i_inj_t = vsub(vconj(vdiv(bus_injection3<T>(s, uv, scn, bus, b.ovl), u)), i_lg);
    int const sb = __ldg(s.src_ptr + bus), se = __ldg(s.src_ptr + bus + 1);
    V3 i_src;
    if (se - sb == 1) {
        i_src = i_inj_t;
    } else {
        int64_t const tile = scn / T;
        int const lane = scn % T;
        C y_t0{0.0, 0.0}, y_t1{0.0, 0.0}, y_t2{0.0, 0.0}, i_ref_1_t{0.0, 0.0};
        for (int k = sb; k < se; ++k) {
            y_t0 = cadd(y_t0, ldc(s.src_y1y0, 2 * k + 1));
            y_t1 = cadd(y_t1, ldc(s.src_y1y0, 2 * k));
            y_t2 = cadd(y_t2, ldc(s.src_y1y0, 2 * k));
        }
        for (int k = sb; k < se; ++k) {
            double const* p = b.usrc + ((tile * s.n_source + k) * 2) * T + lane;
            i_ref_1_t = cadd(i_ref_1_t, cmul(C{p[0], p[T]}, ldc(s.src_y1y0, 2 * k)));
        }
        double const h = 0.8660254037844386; // sqrt3 / 2
        C const a{-0.5, h}, a2{-0.5, -h}, one{1.0, 0.0};
        C const sym[9] = {one, one, one, one, a2, a, one, a, a2};
        C const third[3] = {C{1.0 / 3.0, 0.0 / 3.0}, C{-0.5 / 3.0, h / 3.0}, C{-0.5 / 3.0, -h / 3.0}}; // 1/3, a/3, a2/3
        C const sym_inv[9] = {third[0], third[0], third[0], third[0], third[1], third[2], third[0], third[2], third[1]};
        V3 const i_inj_t_012 = mat_vec_c(sym_inv, i_inj_t);
        double const* p = b.usrc + ((tile * s.n_source + r) * 2) * T + lane;
        C const y1 = ldc(s.src_y1y0, 2 * r), y0 = ldc(s.src_y1y0, 2 * r + 1);
        C const ratio1 = cdiv(y1, y_t1);
        C const lhs1 = cmul(ratio1, csub(cmul(C{p[0], p[T]}, y_t1), i_ref_1_t));
        V3 i_012;
        i_012.v[0] = cmul(cdiv(y0, y_t0), i_inj_t_012.v[0]);
        i_012.v[1] = cadd(lhs1, cmul(ratio1, i_inj_t_012.v[1]));
        i_012.v[2] = cmul(cdiv(y1, y_t2), i_inj_t_012.v[2]);
        i_src = mat_vec_c(sym, i_012);
    }
    u_out = u;
    i_src_out = i_src;
}

} // namespace res3
} // namespace pgmb
