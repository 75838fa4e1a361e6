// Result extraction for the symmetric calculation at the math-solver seam (SolverOutput<symmetric_t>):
//   YBus::calculate_branch_flow / calculate_shunt_flow / calculate_injection   (math_solver/y_bus.hpp:482-546)
//   calculate_load_gen_result / calculate_source_result / calculate_multiple_source_result
//                                                       (math_solver/common_solver_functions.hpp:83-101, 143-160, 383-409)
// One thread per (scenario, element); output is scenario-major like the caller's host buffers, so consecutive threads
// write consecutive elements.  Bus voltages are read from the tile layout written by the solver kernel.
#include "result_common.cuh"

namespace pgmb {
using namespace res;
namespace {

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
            C const v = bus_injection<T>(s, uv, scn, (int)r, b.ovl);
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
            double const* bp = branch_param_of(s, b.ovl, scn, r, 2);
            C const i_f = cadd(cmul(ldc(bp, 0), uf), cmul(ldc(bp, 1), ut));
            C const i_t = cadd(cmul(ldc(bp, 2), uf), cmul(ldc(bp, 3), ut));
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
    C sv, i_src;
    source_result<T>(s, b, uv, scn, (int)r, force_const_y, sv, i_src);
    double* o = out_source + (scn * s.n_source + r) * 4;
    o[0] = sv.r, o[1] = sv.i, o[2] = i_src.r, o[3] = i_src.i;
}

} // namespace

void launch_math_result_sym(int tile_width, DevStructure const& s, DevBatch const& b, int force_const_y, double* out_u,
                            double* out_inj, double* out_branch, double* out_source, double* out_shunt, double* out_lg,
                            cudaStream_t st) {
    count_kernel_launch();
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
