// Result extraction for the asymmetric (three-phase) calculation at the math-solver seam (SolverOutput<asymmetric_t>):
//   YBus::calculate_branch_flow / calculate_shunt_flow / calculate_injection   (math_solver/y_bus.hpp:482-546)
//   calculate_load_gen_result / calculate_source_result / calculate_multiple_source_result
//                                                       (math_solver/common_solver_functions.hpp:83-139, 143-160, 383-409)
// One thread per (scenario, element); phase vectors are per-thread, tensors are read row-major [3][3] complex.
#include "result_common.cuh"

namespace pgmb {
using namespace res;
namespace {

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

template <int T> __device__ V3 bus_injection3(DevStructure const& s, UView3<T> const& uv, int64_t scn, int bus) {
    V3 i_inj{};
    for (int k = __ldg(s.y_row_ptr + bus), ke = __ldg(s.y_row_ptr + bus + 1); k < ke; ++k) {
        i_inj = vadd(i_inj, mat_vec(s.ydata + (size_t)k * 18, uv.get(scn, __ldg(s.y_col_idx + k))));
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

template <int T>
__global__ void math_result_asym_kernel(DevStructure s, DevBatch b, int force_const_y, double* out_u, double* out_inj,
                                        double* out_branch, double* out_source, double* out_shunt, double* out_lg) {
    UView3<T> const uv{b.u, s.n_bus};
    int64_t const per_scn = (int64_t)s.n_bus * 2 + s.n_branch + s.n_shunt + s.n_load_gen + s.n_source;
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_scn * b.n_scn) return;
    int64_t const scn = idx / per_scn;
    int64_t r = idx % per_scn;
    if (r < s.n_bus) {
        if (out_u != nullptr) store3(out_u + (scn * s.n_bus + r) * 6, uv.get(scn, (int)r));
        return;
    }
    r -= s.n_bus;
    if (r < s.n_bus) {
        if (out_inj != nullptr) store3(out_inj + (scn * s.n_bus + r) * 6, bus_injection3<T>(s, uv, scn, (int)r));
        return;
    }
    r -= s.n_bus;
    if (r < s.n_branch) {
        if (out_branch != nullptr) {
            int const f = __ldg(s.branch_bus + 2 * r), t = __ldg(s.branch_bus + 2 * r + 1);
            V3 const uf = f >= 0 ? uv.get(scn, f) : V3{};
            V3 const ut = t >= 0 ? uv.get(scn, t) : V3{};
            double const* bp = s.branch_param + (size_t)r * 4 * 18;
            V3 const i_f = vadd(mat_vec(bp, uf), mat_vec(bp + 18, ut));
            V3 const i_t = vadd(mat_vec(bp + 36, uf), mat_vec(bp + 54, ut));
            double* o = out_branch + (scn * s.n_branch + r) * 24;
            store3(o, vmul(uf, vconj(i_f)));
            store3(o + 6, vmul(ut, vconj(i_t)));
            store3(o + 12, i_f);
            store3(o + 18, i_t);
        }
        return;
    }
    r -= s.n_branch;
    if (r < s.n_shunt) {
        if (out_shunt != nullptr) {
            V3 const u = uv.get(scn, __ldg(s.shunt_bus + r));
            V3 i = mat_vec(s.shunt_param + (size_t)r * 18, u);
            for (int p = 0; p < 3; ++p) i.v[p] = C{-i.v[p].r, -i.v[p].i};
            double* o = out_shunt + (scn * s.n_shunt + r) * 12;
            store3(o, vmul(u, vconj(i)));
            store3(o + 6, i);
        }
        return;
    }
    r -= s.n_shunt;
    if (r < s.n_load_gen) {
        if (out_lg != nullptr) {
            V3 const u = uv.get(scn, __ldg(s.lg_bus + r));
            V3 const sv = load_gen_s3<T>(s, b.sinj, scn, (int)r, u, force_const_y ? 1 : __ldg(s.lg_type + r));
            double* o = out_lg + (scn * s.n_load_gen + r) * 12;
            store3(o, sv);
            store3(o + 6, vconj(vdiv(sv, u)));
        }
        return;
    }
    r -= s.n_load_gen;
    if (out_source == nullptr) return;
    int const bus = __ldg(s.src_bus + r);
    V3 const u = uv.get(scn, bus);
    V3 i_lg{};
    for (int lg = __ldg(s.lg_ptr + bus), lge = __ldg(s.lg_ptr + bus + 1); lg < lge; ++lg) {
        V3 const sv = load_gen_s3<T>(s, b.sinj, scn, lg, u, force_const_y ? 1 : __ldg(s.lg_type + lg));
        i_lg = vadd(i_lg, vconj(vdiv(sv, u)));
    }
    V3 const i_inj_t = vsub(vconj(vdiv(bus_injection3<T>(s, uv, scn, bus), u)), i_lg);
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
    double* o = out_source + (scn * s.n_source + r) * 12;
    store3(o, vmul(u, vconj(i_src)));
    store3(o + 6, i_src);
}

} // namespace

void launch_math_result_asym(int tile_width, DevStructure const& s, DevBatch const& b, int force_const_y, double* out_u,
                             double* out_inj, double* out_branch, double* out_source, double* out_shunt, double* out_lg,
                             cudaStream_t st) {
    count_kernel_launch();
    int64_t const per_scn = (int64_t)s.n_bus * 2 + s.n_branch + s.n_shunt + s.n_load_gen + s.n_source;
    int64_t const total = per_scn * b.n_scn;
    if (total == 0) return;
    int const block = 128;
    unsigned const grid = (unsigned)((total + block - 1) / block);
#define PGMB_LAUNCH(TW) \
    math_result_asym_kernel<TW><<<grid, block, 0, st>>>(s, b, force_const_y, out_u, out_inj, out_branch, out_source, out_shunt, out_lg)
    switch (tile_width) {
    case 4: PGMB_LAUNCH(4); break;
    case 8: PGMB_LAUNCH(8); break;
    case 16: PGMB_LAUNCH(16); break;
    default: PGMB_LAUNCH(32); break;
    }
#undef PGMB_LAUNCH
}

} // namespace pgmb
