// Result extraction for the asymmetric (three-phase) calculation at the math-solver seam (SolverOutput<asymmetric_t>):
//   YBus::calculate_branch_flow / calculate_shunt_flow / calculate_injection   (math_solver/y_bus.hpp:482-546)
//   calculate_load_gen_result / calculate_source_result / calculate_multiple_source_result
//                                                       (math_solver/common_solver_functions.hpp:83-139, 143-160, 383-409)
// One thread per (scenario, element); phase vectors are per-thread, tensors are read row-major [3][3] complex.
#include "result_asym_common.cuh"

namespace pgmb {
using namespace res3;
namespace {

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
        if (out_inj != nullptr) store3(out_inj + (scn * s.n_bus + r) * 6, bus_injection3<T>(s, uv, scn, (int)r, b.ovl));
        return;
    }
    r -= s.n_bus;
    if (r < s.n_branch) {
        if (out_branch != nullptr) {
            int const f = __ldg(s.branch_bus + 2 * r), t = __ldg(s.branch_bus + 2 * r + 1);
            V3 const uf = f >= 0 ? uv.get(scn, f) : V3{};
            V3 const ut = t >= 0 ? uv.get(scn, t) : V3{};
            double const* bp = branch_param_of(s, b.ovl, scn, r, 18);
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
    V3 u, i_src;
    source_result3<T>(s, b, uv, scn, (int)r, force_const_y, u, i_src);
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
