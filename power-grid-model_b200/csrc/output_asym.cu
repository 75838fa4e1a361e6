// Device-side model level for the asymmetric (three-phase) calculation, the counterpart of output_sym.cu:
//   apply_load_update   LoadGen::update / set_power (component/load_gen.hpp:86-104) + calc_param (:124-139): raw LoadGenUpdate
//                       rows -> per-phase per-unit injections in the solver's tile layout (a symmetric load feeds every phase
//                       with its per-unit value, an asymmetric one phase by phase)
//   pack_*              Node / Branch / Appliance::get_output for asymmetric results (component/node.hpp:37-46,
//                       branch.hpp:94-111 incl. the loading rule, appliance.hpp:67-93) fused with the result extraction
// Output rows: NodeOutput<asym> / ApplianceOutput<asym> 128 B, BranchOutput<asym> 208 B, written as doubles; the first word
// packs id + energized.  One thread per (scenario, element).
#include "result_asym_common.cuh"

#include <cuda_runtime.h>

namespace pgmb {
using namespace res3;
namespace {

struct SymLoadGenUpdateRow { // LoadGenUpdate<symmetric_t>
    int32_t id;
    int8_t status;
    double p_specified, q_specified;
};
struct AsymLoadGenUpdateRow { // LoadGenUpdate<asymmetric_t>
    int32_t id;
    int8_t status;
    double p_specified[3], q_specified[3];
};
constexpr int8_t kNaIntS = -128;
constexpr double kBasePower1p = 1e6 / 3.0; // base_power<asymmetric_t>
constexpr double kInvSqrt3 = 1.0 / 1.7320508075688772;

template <int T>
__global__ void apply_load_update_asym_kernel(DevStructure s, DevBatch b, DevModelTables m, DevUpdateBuffers ub) {
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // (tile, lg, lane)
    int64_t const total = (int64_t)b.n_tile * s.n_load_gen * T;
    if (idx >= total) return;
    int const lane = idx % T;
    int64_t const r = idx / T;
    int const lg = r % s.n_load_gen;
    int64_t const tile = r / s.n_load_gen;
    int64_t const scn = tile * T + lane;
    double out_r[3] = {0.0, 0.0, 0.0}, out_i[3] = {0.0, 0.0, 0.0};
    uint8_t status = 0;
    if (scn < b.n_scn) {
        int const phases = m.lg_phases[lg];
        double sr[3], si[3];
        for (int p = 0; p < 3; ++p) {
            sr[p] = m.lg_base_s[(lg * 3 + p) * 2];
            si[p] = m.lg_base_s[(lg * 3 + p) * 2 + 1];
        }
        status = m.lg_base_status[lg];
        int const buf = m.lg_upd_buf[lg];
        if (buf >= 0) {
            double const scale = m.lg_scale[lg];
            int64_t const row = (scn * ub.n_per_scenario[buf] + m.lg_upd_pos[lg]);
            if (phases == 1) {
                SymLoadGenUpdateRow const u = static_cast<SymLoadGenUpdateRow const*>(ub.data[buf])[row];
                if (u.id != m.lg_upd_id[lg]) *ub.id_mismatch = 1;
                if (u.status != kNaIntS) status = u.status != 0;
                if (!isnan(u.p_specified)) sr[0] = scale * u.p_specified;
                if (!isnan(u.q_specified)) si[0] = scale * u.q_specified;
            } else {
                AsymLoadGenUpdateRow const* u = static_cast<AsymLoadGenUpdateRow const*>(ub.data[buf]) + row;
                int8_t const st = u->status;
                if (u->id != m.lg_upd_id[lg]) *ub.id_mismatch = 1;
                if (st != kNaIntS) status = st != 0;
                for (int p = 0; p < 3; ++p) {
                    double const pp = u->p_specified[p], qq = u->q_specified[p];
                    if (!isnan(pp)) sr[p] = scale * pp;
                    if (!isnan(qq)) si[p] = scale * qq;
                }
            }
        }
        if (status) {
            if (phases == 1) {
                bool const bad = isnan(sr[0]) || isnan(si[0]);
                for (int p = 0; p < 3; ++p) {
                    out_r[p] = bad ? NAN : sr[0];
                    out_i[p] = bad ? NAN : si[0];
                }
            } else {
                for (int p = 0; p < 3; ++p) {
                    out_r[p] = sr[p];
                    out_i[p] = si[p];
                }
            }
        }
    }
    double* o = b.sinj + ((tile * s.n_load_gen + lg) * 6) * T + lane;
    for (int p = 0; p < 3; ++p) {
        o[(size_t)(2 * p) * T] = out_r[p];
        o[(size_t)(2 * p + 1) * T] = out_i[p];
    }
    b.lg_status[(tile * s.n_load_gen + lg) * T + lane] = status;
}

__device__ __forceinline__ double head_word(int32_t id, int energized) {
    unsigned long long const w = (unsigned long long)(unsigned int)id | ((unsigned long long)(energized & 0xff) << 32);
    return __longlong_as_double((long long)w);
}
__device__ __forceinline__ double cabs_(C a) { return sqrt(a.r * a.r + a.i * a.i); }

// source results [scn][n_source][12] = s[3] (re, im), i[3] (re, im)
template <int T>
__global__ void source_result_asym_kernel(DevStructure s, DevBatch b, int force_const_y, double* __restrict__ out) {
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b.n_scn * s.n_source) return;
    int64_t const scn = idx / s.n_source;
    int const r = idx % s.n_source;
    UView3<T> const uv{b.u, s.n_bus};
    V3 u, i_src;
    source_result3<T>(s, b, uv, scn, r, force_const_y, u, i_src);
    store3(out + idx * 12, vmul(u, vconj(i_src)));
    store3(out + idx * 12 + 6, i_src);
}

// NodeOutput<asymmetric_t>: 16 doubles = head, u_pu[3], u[3], u_angle[3], p[3], q[3]
template <int T>
__device__ __forceinline__ void pack_node_asym_row(DevStructure s, DevBatch b, DevModelTables m, int force_const_y,
                                      double const* __restrict__ src_res, int64_t idx, double* __restrict__ o) {
    int64_t const scn = idx / m.n_node;
    int const node = idx % m.n_node;
    int const bus = __ldg(m.node_bus + node);
    int32_t const id = __ldg(m.node_id + node);
    if (bus < 0 || bus_is_dead(b.ovl, scn, bus, s.n_bus)) {
        o[0] = head_word(id, 0);
        for (int k = 1; k < 16; ++k) o[k] = 0.0;
        return;
    }
    UView3<T> const uv{b.u, s.n_bus};
    V3 const u = uv.get(scn, bus);
    V3 inj{};
    for (int k = __ldg(m.node_app_ptr + node), ke = __ldg(m.node_app_ptr + node + 1); k < ke; ++k) {
        int const code = __ldg(m.node_app + k);
        int const a = code & 0x0fffffff;
        if ((code >> 28) == 0) {
            double const* p = src_res + (scn * s.n_source + a) * 12;
            for (int ph = 0; ph < 3; ++ph) inj.v[ph] = cadd(inj.v[ph], C{p[2 * ph], p[2 * ph + 1]});
        } else {
            inj = vadd(inj, load_gen_s3<T>(s, b.sinj, scn, a, u, force_const_y ? 1 : __ldg(s.lg_type + a)));
        }
    }
    o[0] = head_word(id, 1);
    double const u_rated = __ldg(m.node_u_rated + node);
    for (int ph = 0; ph < 3; ++ph) {
        double const u_pu = cabs_(u.v[ph]);
        o[1 + ph] = u_pu;
        o[4 + ph] = kInvSqrt3 * u_rated * u_pu; // u_scale<asym> = 1 / sqrt3
        o[7 + ph] = atan2(u.v[ph].i, u.v[ph].r);
        o[10 + ph] = kBasePower1p * inj.v[ph].r;
        o[13 + ph] = kBasePower1p * inj.v[ph].i;
    }
}

// rows are staged in shared memory and leave the block as one contiguous, fully coalesced run of 16-byte stores: the rows of
// consecutive threads are adjacent in the caller's layout, so a block owns 128 * 16 consecutive doubles of the output
template <int T>
__global__ void __launch_bounds__(128) pack_node_asym_kernel(DevStructure s, DevBatch b, DevModelTables m, int force_const_y,
                                      double const* __restrict__ src_res, double* __restrict__ out) {
    __shared__ __align__(16) double rows[128 * 16];
    int64_t const total = b.n_scn * m.n_node;
    int64_t const row0 = (int64_t)blockIdx.x * 128;
    int64_t const idx = row0 + threadIdx.x;
    if (idx < total) pack_node_asym_row<T>(s, b, m, force_const_y, src_res, idx, rows + threadIdx.x * 16);
    __syncthreads();
    int const n2 = (int)(min((int64_t)128, total - row0) * 16 / 2);
    double2* dst = reinterpret_cast<double2*>(out + row0 * 16);
    double2 const* src = reinterpret_cast<double2 const*>(rows);
    for (int i = threadIdx.x; i < n2; i += 128) dst[i] = src[i];
}

// BranchOutput<asymmetric_t>: 26 doubles = head, loading, p_from[3], q_from[3], i_from[3], s_from[3], p_to .. s_to
template <int T>
__device__ __forceinline__ void pack_branch_asym_row(DevStructure s, DevBatch b, DevModelTables m, int first, int count, int64_t idx, double* __restrict__ o) {
    int64_t const scn = idx / count;
    int const comp = first + (int)(idx % count);
    int const mb = __ldg(m.branch_math + comp);
    int32_t const id = __ldg(m.branch_id + comp);
    bool all_dead = false;
    if (mb >= 0 && b.ovl.dead_off != nullptr) { // a branch whose connected sides all sit on buses that lost their supply
        int const bf = __ldg(s.branch_bus + 2 * mb), bt = __ldg(s.branch_bus + 2 * mb + 1);
        all_dead = (bf < 0 || bus_is_dead(b.ovl, scn, bf, s.n_bus)) && (bt < 0 || bus_is_dead(b.ovl, scn, bt, s.n_bus));
    }
    if (mb < 0 || all_dead) {
        o[0] = head_word(id, 0);
        for (int k = 1; k < 26; ++k) o[k] = 0.0;
        return;
    }
    UView3<T> const uv{b.u, s.n_bus};
    int const f = __ldg(s.branch_bus + 2 * mb), t = __ldg(s.branch_bus + 2 * mb + 1);
    V3 const uf = f >= 0 ? uv.get(scn, f) : V3{};
    V3 const ut = t >= 0 ? uv.get(scn, t) : V3{};
    double const* bp = branch_param_of(s, b.ovl, scn, mb, 18);
    V3 const i_f = vadd(mat_vec(bp, uf), mat_vec(bp + 18, ut));
    V3 const i_t = vadd(mat_vec(bp + 36, uf), mat_vec(bp + 54, ut));
    V3 const s_f = vmul(uf, vconj(i_f));
    V3 const s_t = vmul(ut, vconj(i_t));
    double const base_f = __ldg(m.branch_base_i + 2 * comp), base_t = __ldg(m.branch_base_i + 2 * comp + 1);
    double sum_sf = 0.0, sum_st = 0.0, max_if = 0.0, max_it = 0.0;
    for (int ph = 0; ph < 3; ++ph) {
        double const i_from = base_f * cabs_(i_f.v[ph]);
        double const i_to = base_t * cabs_(i_t.v[ph]);
        double const s_from = kBasePower1p * cabs_(s_f.v[ph]);
        double const s_to = kBasePower1p * cabs_(s_t.v[ph]);
        o[2 + ph] = kBasePower1p * s_f.v[ph].r;
        o[5 + ph] = kBasePower1p * s_f.v[ph].i;
        o[8 + ph] = i_from;
        o[11 + ph] = s_from;
        o[14 + ph] = kBasePower1p * s_t.v[ph].r;
        o[17 + ph] = kBasePower1p * s_t.v[ph].i;
        o[20 + ph] = i_to;
        o[23 + ph] = s_to;
        sum_sf = ph == 0 ? s_from : sum_sf + s_from;
        sum_st = ph == 0 ? s_to : sum_st + s_to;
        max_if = ph == 0 ? i_from : fmax(max_if, i_from);
        max_it = ph == 0 ? i_to : fmax(max_it, i_to);
    }
    double const rating = __ldg(m.branch_rating + comp);
    int const energized = branch_energized_of(b.ovl, scn, comp, __ldg(m.branch_energized + comp));
    o[0] = head_word(id, energized);
    o[1] = rating > 0.0 ? fmax(sum_sf, sum_st) / rating : fmax(max_if, max_it) / (-rating);
}

// rows are staged in shared memory and leave the block as one contiguous, fully coalesced run of 16-byte stores: the rows of
// consecutive threads are adjacent in the caller's layout, so a block owns 128 * 26 consecutive doubles of the output
template <int T>
__global__ void __launch_bounds__(128) pack_branch_asym_kernel(DevStructure s, DevBatch b, DevModelTables m, int first, int count,
                                        double* __restrict__ out) {
    __shared__ __align__(16) double rows[128 * 26];
    int64_t const total = b.n_scn * count;
    int64_t const row0 = (int64_t)blockIdx.x * 128;
    int64_t const idx = row0 + threadIdx.x;
    if (idx < total) pack_branch_asym_row<T>(s, b, m, first, count, idx, rows + threadIdx.x * 26);
    __syncthreads();
    int const n2 = (int)(min((int64_t)128, total - row0) * 26 / 2);
    double2* dst = reinterpret_cast<double2*>(out + row0 * 26);
    double2 const* src = reinterpret_cast<double2 const*>(rows);
    for (int i = threadIdx.x; i < n2; i += 128) dst[i] = src[i];
}

// ApplianceOutput<asymmetric_t>: 16 doubles = head, p[3], q[3], i[3], s[3], pf[3]
template <int T>
__device__ __forceinline__ void pack_appliance_asym_row(DevStructure s, DevBatch b, DevModelTables m, int force_const_y, int first,
                                           int count, double const* __restrict__ src_res, int64_t idx, double* __restrict__ o) {
    int64_t const scn = idx / count;
    int const comp = first + (int)(idx % count);
    int const a = __ldg(m.app_math + comp);
    int const kind = __ldg(m.app_kind + comp);
    int32_t const id = __ldg(m.app_id + comp);
    bool dead = false;
    if (a >= 0 && b.ovl.dead_off != nullptr) {
        int const bus = kind == 0 ? __ldg(s.shunt_bus + a) : (kind == 1 ? __ldg(s.src_bus + a) : __ldg(s.lg_bus + a));
        dead = bus_is_dead(b.ovl, scn, bus, s.n_bus);
    }
    if (a < 0 || dead) {
        o[0] = head_word(id, 0);
        for (int k = 1; k < 16; ++k) o[k] = 0.0;
        return;
    }
    UView3<T> const uv{b.u, s.n_bus};
    V3 sv, iv;
    int energized;
    if (kind == 0) {
        V3 const u = uv.get(scn, __ldg(s.shunt_bus + a));
        iv = mat_vec(s.shunt_param + (size_t)a * 18, u);
        for (int ph = 0; ph < 3; ++ph) iv.v[ph] = C{-iv.v[ph].r, -iv.v[ph].i};
        sv = vmul(u, vconj(iv));
        energized = __ldg(m.app_status + comp);
    } else if (kind == 1) {
        double const* p = src_res + (scn * s.n_source + a) * 12;
        for (int ph = 0; ph < 3; ++ph) {
            sv.v[ph] = C{p[2 * ph], p[2 * ph + 1]};
            iv.v[ph] = C{p[6 + 2 * ph], p[6 + 2 * ph + 1]};
        }
        energized = __ldg(m.app_status + comp);
    } else {
        V3 const u = uv.get(scn, __ldg(s.lg_bus + a));
        sv = load_gen_s3<T>(s, b.sinj, scn, a, u, force_const_y ? 1 : __ldg(s.lg_type + a));
        iv = vconj(vdiv(sv, u));
        energized = b.lg_status[((scn / T) * s.n_load_gen + a) * T + (scn % T)];
    }
    double const dir = __ldg(m.app_dir + comp);
    double const base_i = __ldg(m.app_base_i + comp);
    o[0] = head_word(id, energized);
    for (int ph = 0; ph < 3; ++ph) {
        double const pw = kBasePower1p * sv.v[ph].r * dir;
        double const sa = kBasePower1p * cabs_(sv.v[ph]);
        o[1 + ph] = pw;
        o[4 + ph] = kBasePower1p * sv.v[ph].i * dir;
        o[7 + ph] = base_i * cabs_(iv.v[ph]);
        o[10 + ph] = sa;
        o[13 + ph] = sa < 1e-8 ? 0.0 : pw / sa;
    }
}

// rows are staged in shared memory and leave the block as one contiguous, fully coalesced run of 16-byte stores: the rows of
// consecutive threads are adjacent in the caller's layout, so a block owns 128 * 16 consecutive doubles of the output
template <int T>
__global__ void __launch_bounds__(128) pack_appliance_asym_kernel(DevStructure s, DevBatch b, DevModelTables m, int force_const_y, int first,
                                           int count, double const* __restrict__ src_res, double* __restrict__ out) {
    __shared__ __align__(16) double rows[128 * 16];
    int64_t const total = b.n_scn * count;
    int64_t const row0 = (int64_t)blockIdx.x * 128;
    int64_t const idx = row0 + threadIdx.x;
    if (idx < total) pack_appliance_asym_row<T>(s, b, m, force_const_y, first, count, src_res, idx, rows + threadIdx.x * 16);
    __syncthreads();
    int const n2 = (int)(min((int64_t)128, total - row0) * 16 / 2);
    double2* dst = reinterpret_cast<double2*>(out + row0 * 16);
    double2 const* src = reinterpret_cast<double2 const*>(rows);
    for (int i = threadIdx.x; i < n2; i += 128) dst[i] = src[i];
}

inline unsigned grid_for(int64_t total, int block) { return (unsigned)((total + block - 1) / block); }

} // namespace

#define PGMB_DISPATCH_T(TW, KERNEL, GRID, BLOCK, ST, ...)                    \
    switch (TW) {                                                             \
    case 4: KERNEL<4><<<GRID, BLOCK, 0, ST>>>(__VA_ARGS__); break;            \
    case 8: KERNEL<8><<<GRID, BLOCK, 0, ST>>>(__VA_ARGS__); break;            \
    case 16: KERNEL<16><<<GRID, BLOCK, 0, ST>>>(__VA_ARGS__); break;          \
    default: KERNEL<32><<<GRID, BLOCK, 0, ST>>>(__VA_ARGS__); break;          \
    }

void launch_apply_load_update_asym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m,
                                   DevUpdateBuffers const& ub, cudaStream_t st) {
    count_kernel_launch();
    int64_t const total = (int64_t)b.n_tile * s.n_load_gen * tw;
    if (total == 0) return;
    PGMB_DISPATCH_T(tw, apply_load_update_asym_kernel, grid_for(total, 256), 256, st, s, b, m, ub)
}
void launch_source_result_asym(int tw, DevStructure const& s, DevBatch const& b, int force_const_y, double* out, cudaStream_t st) {
    count_kernel_launch();
    int64_t const total = b.n_scn * s.n_source;
    if (total == 0) return;
    PGMB_DISPATCH_T(tw, source_result_asym_kernel, grid_for(total, 128), 128, st, s, b, force_const_y, out)
}
void launch_pack_node_asym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int force_const_y,
                           double const* src_res, void* out, cudaStream_t st) {
    count_kernel_launch();
    int64_t const total = b.n_scn * m.n_node;
    if (total == 0) return;
    PGMB_DISPATCH_T(tw, pack_node_asym_kernel, grid_for(total, 128), 128, st, s, b, m, force_const_y, src_res, static_cast<double*>(out))
}
void launch_pack_branch_asym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int first, int count,
                             void* out, cudaStream_t st) {
    count_kernel_launch();
    int64_t const total = b.n_scn * count;
    if (total == 0) return;
    PGMB_DISPATCH_T(tw, pack_branch_asym_kernel, grid_for(total, 128), 128, st, s, b, m, first, count, static_cast<double*>(out))
}
void launch_pack_appliance_asym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int force_const_y,
                                int first, int count, double const* src_res, void* out, cudaStream_t st) {
    count_kernel_launch();
    int64_t const total = b.n_scn * count;
    if (total == 0) return;
    PGMB_DISPATCH_T(tw, pack_appliance_asym_kernel, grid_for(total, 128), 128, st, s, b, m, force_const_y, first, count, src_res,
                    static_cast<double*>(out))
}

} // namespace pgmb
