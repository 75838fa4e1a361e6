// Device-side model level for the symmetric calculation: apply load-profile updates and write the caller's output structs.
//   apply_load_update   LoadGen::update / set_power (component/load_gen.hpp:86-104) + calc_param (:124-139) for every
//                       scenario: raw LoadGenUpdate rows (24 B / 56 B) -> per-unit injections in the solver's tile layout
//   pack_*              Node / Branch / Appliance::get_output (component/node.hpp:37-46, branch.hpp:94-111,
//                       appliance.hpp:67-93) fused with the result extraction they consume (y_bus.hpp:482-546,
//                       common_solver_functions.hpp:383-446) and the node injection sum (topological_node_output.hpp:72-117)
// One thread per (scenario, element): consecutive threads write consecutive output structs (scenario-major, exactly the
// caller's buffer layout), so the 48 / 80 byte rows coalesce; voltages are gathered from the solver's tile layout.
#include "result_common.cuh"

#include <cuda_runtime.h>

namespace pgmb {
using namespace res;
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
constexpr double kBasePower = 1e6; // base_power<symmetric_t>

template <int T>
__global__ void apply_load_update_sym_kernel(DevStructure s, DevBatch b, DevModelTables m, DevUpdateBuffers ub) {
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // (tile, lg, lane)
    int64_t const total = (int64_t)b.n_tile * s.n_load_gen * T;
    if (idx >= total) return;
    int const lane = idx % T;
    int64_t const r = idx / T;
    int const lg = r % s.n_load_gen;
    int64_t const tile = r / s.n_load_gen;
    int64_t const scn = tile * T + lane;
    double re = 0.0, im = 0.0;
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
                re = bad ? NAN : sr[0];
                im = bad ? NAN : si[0];
            } else { // mean over the phases: (s0 + s1 + s2) / 3
                re = (sr[0] + sr[1] + sr[2]) / 3.0;
                im = (si[0] + si[1] + si[2]) / 3.0;
            }
        }
    }
    double* o = b.sinj + ((tile * s.n_load_gen + lg) * 2) * T + lane;
    o[0] = re;
    o[T] = im;
    b.lg_status[(tile * s.n_load_gen + lg) * T + lane] = status;
}

__device__ __forceinline__ double head_word(int32_t id, int energized) {
    unsigned long long const w = (unsigned long long)(unsigned int)id | ((unsigned long long)(energized & 0xff) << 32);
    return __longlong_as_double((long long)w);
}
__device__ __forceinline__ double cabs_(C a) { return sqrt(a.r * a.r + a.i * a.i); }

template <int T> __device__ __forceinline__ C load_gen_power(DevStructure const& s, DevBatch const& b, int64_t scn, int lg, C u, int force_const_y) {
    return load_gen_s<T>(s, b.sinj, scn, lg, u, force_const_y ? 1 : __ldg(s.lg_type + lg));
}

// source results [scn][n_source][4] = s.re, s.im, i.re, i.im
template <int T>
__global__ void source_result_sym_kernel(DevStructure s, DevBatch b, int force_const_y, double* __restrict__ out) {
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b.n_scn * s.n_source) return;
    int64_t const scn = idx / s.n_source;
    int const r = idx % s.n_source;
    UView<T> const uv{b.u, s.n_bus};
    C sv, i_src;
    source_result<T>(s, b, uv, scn, r, force_const_y, sv, i_src);
    double* o = out + idx * 4;
    o[0] = sv.r, o[1] = sv.i, o[2] = i_src.r, o[3] = i_src.i;
}

// NodeOutput<symmetric_t>: 48 bytes = head, u_pu, u, u_angle, p, q
template <int T>
__device__ __forceinline__ void pack_node_sym_row(DevStructure s, DevBatch b, DevModelTables m, int force_const_y,
                                     double const* __restrict__ src_res, int64_t idx, double* __restrict__ o) {
    int64_t const scn = idx / m.n_node;
    int const node = idx % m.n_node;
    int const bus = __ldg(m.node_bus + node);
    int32_t const id = __ldg(m.node_id + node);
    if (bus < 0 || bus_is_dead(b.ovl, scn, bus, s.n_bus)) {
        o[0] = head_word(id, 0);
        o[1] = o[2] = o[3] = o[4] = o[5] = 0.0;
        return;
    }
    UView<T> const uv{b.u, s.n_bus};
    C const u = uv.get(scn, bus);
    C inj{0.0, 0.0};
    for (int k = __ldg(m.node_app_ptr + node), ke = __ldg(m.node_app_ptr + node + 1); k < ke; ++k) {
        int const code = __ldg(m.node_app + k);
        int const a = code & 0x0fffffff;
        if ((code >> 28) == 0) {
            inj = cadd(inj, C{src_res[(scn * s.n_source + a) * 4], src_res[(scn * s.n_source + a) * 4 + 1]});
        } else {
            inj = cadd(inj, load_gen_power<T>(s, b, scn, a, u, force_const_y));
        }
    }
    double const u_pu = cabs_(u);
    o[0] = head_word(id, 1);
    o[1] = u_pu;
    o[2] = 1.0 * __ldg(m.node_u_rated + node) * u_pu; // u_scale<sym> = 1
    o[3] = atan2(u.i, u.r);
    o[4] = kBasePower * inj.r;
    o[5] = kBasePower * inj.i;
}

// rows are staged in shared memory and leave the block as one contiguous, fully coalesced run of 16-byte stores: the rows of
// consecutive threads are adjacent in the caller's layout, so a block owns 256 * 6 consecutive doubles of the output
template <int T>
__global__ void __launch_bounds__(256) pack_node_sym_kernel(DevStructure s, DevBatch b, DevModelTables m, int force_const_y,
                                     double const* __restrict__ src_res, double* __restrict__ out) {
    __shared__ __align__(16) double rows[256 * 6];
    int64_t const total = b.n_scn * m.n_node;
    int64_t const row0 = (int64_t)blockIdx.x * 256;
    int64_t const idx = row0 + threadIdx.x;
    if (idx < total) pack_node_sym_row<T>(s, b, m, force_const_y, src_res, idx, rows + threadIdx.x * 6);
    __syncthreads();
    int const n2 = (int)(min((int64_t)256, total - row0) * 6 / 2);
    double2* dst = reinterpret_cast<double2*>(out + row0 * 6);
    double2 const* src = reinterpret_cast<double2 const*>(rows);
    for (int i = threadIdx.x; i < n2; i += 256) dst[i] = src[i];
}

// BranchOutput<symmetric_t>: 80 bytes = head, loading, p_from, q_from, i_from, s_from, p_to, q_to, i_to, s_to
template <int T>
__device__ __forceinline__ void pack_branch_sym_row(DevStructure s, DevBatch b, DevModelTables m, int first, int count, int64_t idx, double* __restrict__ o) {
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
        for (int k = 1; k < 10; ++k) o[k] = 0.0;
        return;
    }
    UView<T> const uv{b.u, s.n_bus};
    int const f = __ldg(s.branch_bus + 2 * mb), t = __ldg(s.branch_bus + 2 * mb + 1);
    C const uf = f >= 0 ? uv.get(scn, f) : C{0.0, 0.0};
    C const ut = t >= 0 ? uv.get(scn, t) : C{0.0, 0.0};
    double const* bp = branch_param_of(s, b.ovl, scn, mb, 2);
    C const i_f = cadd(cmul(ldc(bp, 0), uf), cmul(ldc(bp, 1), ut));
    C const i_t = cadd(cmul(ldc(bp, 2), uf), cmul(ldc(bp, 3), ut));
    C const s_f = cmul(uf, conj(i_f));
    C const s_t = cmul(ut, conj(i_t));
    double const i_from = __ldg(m.branch_base_i + 2 * comp) * cabs_(i_f);
    double const i_to = __ldg(m.branch_base_i + 2 * comp + 1) * cabs_(i_t);
    double const s_from = kBasePower * cabs_(s_f);
    double const s_to = kBasePower * cabs_(s_t);
    double const rating = __ldg(m.branch_rating + comp);
    int const energized = branch_energized_of(b.ovl, scn, comp, __ldg(m.branch_energized + comp));
    o[0] = head_word(id, energized);
    o[1] = rating > 0.0 ? fmax(s_from, s_to) / rating : fmax(i_from, i_to) / (-rating);
    o[2] = kBasePower * s_f.r;
    o[3] = kBasePower * s_f.i;
    o[4] = i_from;
    o[5] = s_from;
    o[6] = kBasePower * s_t.r;
    o[7] = kBasePower * s_t.i;
    o[8] = i_to;
    o[9] = s_to;
}

// rows are staged in shared memory and leave the block as one contiguous, fully coalesced run of 16-byte stores: the rows of
// consecutive threads are adjacent in the caller's layout, so a block owns 256 * 10 consecutive doubles of the output
template <int T>
__global__ void __launch_bounds__(256) pack_branch_sym_kernel(DevStructure s, DevBatch b, DevModelTables m, int first, int count,
                                       double* __restrict__ out) {
    __shared__ __align__(16) double rows[256 * 10];
    int64_t const total = b.n_scn * count;
    int64_t const row0 = (int64_t)blockIdx.x * 256;
    int64_t const idx = row0 + threadIdx.x;
    if (idx < total) pack_branch_sym_row<T>(s, b, m, first, count, idx, rows + threadIdx.x * 10);
    __syncthreads();
    int const n2 = (int)(min((int64_t)256, total - row0) * 10 / 2);
    double2* dst = reinterpret_cast<double2*>(out + row0 * 10);
    double2 const* src = reinterpret_cast<double2 const*>(rows);
    for (int i = threadIdx.x; i < n2; i += 256) dst[i] = src[i];
}

// ApplianceOutput<symmetric_t>: 48 bytes = head, p, q, i, s, pf
template <int T>
__device__ __forceinline__ void pack_appliance_sym_row(DevStructure s, DevBatch b, DevModelTables m, int force_const_y, int first,
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
        o[1] = o[2] = o[3] = o[4] = o[5] = 0.0;
        return;
    }
    UView<T> const uv{b.u, s.n_bus};
    C sv, iv;
    int energized;
    if (kind == 0) {
        C const u = uv.get(scn, __ldg(s.shunt_bus + a));
        C const yu = cmul(ldc(s.shunt_param, a), u);
        iv = C{-yu.r, -yu.i};
        sv = cmul(u, conj(iv));
        energized = __ldg(m.app_status + comp);
    } else if (kind == 1) {
        double const* p = src_res + (scn * s.n_source + a) * 4;
        sv = C{p[0], p[1]};
        iv = C{p[2], p[3]};
        energized = __ldg(m.app_status + comp);
    } else {
        C const u = uv.get(scn, __ldg(s.lg_bus + a));
        sv = load_gen_power<T>(s, b, scn, a, u, force_const_y);
        iv = conj(cdiv(sv, u));
        energized = b.lg_status[((scn / T) * s.n_load_gen + a) * T + (scn % T)];
    }
    double const dir = __ldg(m.app_dir + comp);
    double const pw = kBasePower * sv.r * dir;
    double const sa = kBasePower * cabs_(sv);
    o[0] = head_word(id, energized);
    o[1] = pw;
    o[2] = kBasePower * sv.i * dir;
    o[3] = __ldg(m.app_base_i + comp) * cabs_(iv);
    o[4] = sa;
    o[5] = sa < 1e-8 ? 0.0 : pw / sa;
}

// rows are staged in shared memory and leave the block as one contiguous, fully coalesced run of 16-byte stores: the rows of
// consecutive threads are adjacent in the caller's layout, so a block owns 256 * 6 consecutive doubles of the output
template <int T>
__global__ void __launch_bounds__(256) pack_appliance_sym_kernel(DevStructure s, DevBatch b, DevModelTables m, int force_const_y, int first,
                                          int count, double const* __restrict__ src_res, double* __restrict__ out) {
    __shared__ __align__(16) double rows[256 * 6];
    int64_t const total = b.n_scn * count;
    int64_t const row0 = (int64_t)blockIdx.x * 256;
    int64_t const idx = row0 + threadIdx.x;
    if (idx < total) pack_appliance_sym_row<T>(s, b, m, force_const_y, first, count, src_res, idx, rows + threadIdx.x * 6);
    __syncthreads();
    int const n2 = (int)(min((int64_t)256, total - row0) * 6 / 2);
    double2* dst = reinterpret_cast<double2*>(out + row0 * 6);
    double2 const* src = reinterpret_cast<double2 const*>(rows);
    for (int i = threadIdx.x; i < n2; i += 256) dst[i] = src[i];
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

void launch_apply_load_update_sym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m,
                                  DevUpdateBuffers const& ub, cudaStream_t st) {
    count_kernel_launch();
    int64_t const total = (int64_t)b.n_tile * s.n_load_gen * tw;
    if (total == 0) return;
    PGMB_DISPATCH_T(tw, apply_load_update_sym_kernel, grid_for(total, 256), 256, st, s, b, m, ub);
}
void launch_source_result_sym(int tw, DevStructure const& s, DevBatch const& b, int force_const_y, double* out, cudaStream_t st) {
    count_kernel_launch();
    int64_t const total = b.n_scn * s.n_source;
    if (total == 0) return;
    PGMB_DISPATCH_T(tw, source_result_sym_kernel, grid_for(total, 128), 128, st, s, b, force_const_y, out);
}
void launch_pack_node_sym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int force_const_y,
                          double const* src_res, void* out, cudaStream_t st) {
    count_kernel_launch();
    int64_t const total = b.n_scn * m.n_node;
    if (total == 0) return;
    PGMB_DISPATCH_T(tw, pack_node_sym_kernel, grid_for(total, 256), 256, st, s, b, m, force_const_y, src_res, static_cast<double*>(out));
}
void launch_pack_branch_sym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int first, int count,
                            void* out, cudaStream_t st) {
    count_kernel_launch();
    int64_t const total = b.n_scn * count;
    if (total == 0) return;
    PGMB_DISPATCH_T(tw, pack_branch_sym_kernel, grid_for(total, 256), 256, st, s, b, m, first, count, static_cast<double*>(out));
}
void launch_pack_appliance_sym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int force_const_y,
                               int first, int count, double const* src_res, void* out, cudaStream_t st) {
    count_kernel_launch();
    int64_t const total = b.n_scn * count;
    if (total == 0) return;
    PGMB_DISPATCH_T(tw, pack_appliance_sym_kernel, grid_for(total, 256), 256, st, s, b, m, force_const_y, first, count, src_res,
                    static_cast<double*>(out));
}

} // namespace pgmb
