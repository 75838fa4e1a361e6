// Reactive power of the generators under a voltage regulator and the regulator outputs, after the Newton-Raphson result step.
//   calculate_voltage_regulator_result / distribute_q / allocate_q_bus_limit_violated / allocate_q_iterative_distribution
//                                                        (math_solver/common_solver_functions.hpp:162-381)
//   finalize_result: bus_q_limit_violated                (math_solver/newton_raphson_pf_solver.hpp:351-360)
// One thread per (scenario, regulated bus).  It reads the scenario-major result arrays the math_result kernels have just
// written (u, bus injection, load_gen s / i), re-assigns the Q of the regulating generators of its bus and rewrites their
// s and i; every decision is local to the bus.  A bus carries few regulated generators (kMaxRegPerBus, checked on the host).
#include "result_asym_common.cuh"

namespace pgmb {
using namespace res;
namespace {

constexpr int kMaxRegPerBus = 16;
constexpr double kNumTol = 1e-8;

template <int B> __device__ __forceinline__ double total_q(double const* q) {
    if constexpr (B == 1) {
        return q[0];
    } else {
        return q[0] + q[1] + q[2];
    }
}
template <int B> __device__ __forceinline__ void distribute_q(double q_scalar, double const* base, double* out) {
    if constexpr (B == 1) {
        out[0] = q_scalar;
    } else {
        double const base_total = total_q<3>(base);
        if (fabs(base_total) > kNumTol) {
            double const scale = q_scalar / base_total;
            for (int p = 0; p < 3; ++p) out[p] = base[p] * scale;
        } else {
            for (int p = 0; p < 3; ++p) out[p] = q_scalar / 3.0;
        }
    }
}

// The allocation itself for one bus: st_* describe the regulating generators (in load_gen order), q_remaining what the bus
// injects beyond its other load_gens, base_q the generators' own Q (bus-limit case).  Returns false when Q stays unallocated.
template <int B>
__device__ bool allocate_bus_q(DevStructure const& s, int bus_limit, int n_regulating, int const* st_reg, double (*st_q)[B],
                               double (*base_q)[B], double* q_remaining, int8_t* reg_flag /*[n_regulating]*/) {
    bool ok = true;
    if (bus_limit == 0) {
        bool st_cap[kMaxRegPerBus];
        for (int k = 0; k < n_regulating; ++k) {
            st_cap[k] = true;
            reg_flag[k] = 0;
        }
        int n_active = n_regulating;
        while (fabs(total_q<B>(q_remaining)) > kNumTol && n_active > 0) {
            double q_per[B], q_unallocated[B];
            for (int p = 0; p < B; ++p) {
                q_per[p] = q_remaining[p] / n_active;
                q_unallocated[p] = 0.0;
            }
            for (int k = 0; k < n_regulating; ++k) {
                if (!st_cap[k]) continue;
                double const q_min = __ldg(s.reg_param + 4 * st_reg[k] + 2), q_max = __ldg(s.reg_param + 4 * st_reg[k] + 3);
                double q_prev[B], q_next[B];
                for (int p = 0; p < B; ++p) {
                    q_prev[p] = st_q[k][p];
                    q_next[p] = q_prev[p] + q_per[p];
                }
                double const q_next_scalar = total_q<B>(q_next);
                bool const hit_upper = !isnan(q_max) && q_next_scalar > q_max + kNumTol;
                bool const hit_lower = !hit_upper && !isnan(q_min) && q_next_scalar < q_min - kNumTol;
                if (hit_upper || hit_lower) {
                    distribute_q<B>(hit_upper ? q_max : q_min, q_next, st_q[k]);
                    st_cap[k] = false;
                    for (int p = 0; p < B; ++p) q_unallocated[p] += q_per[p] - (st_q[k][p] - q_prev[p]);
                    n_active -= 1;
                } else {
                    for (int p = 0; p < B; ++p) st_q[k][p] = q_next[p];
                }
            }
            double diff[B];
            for (int p = 0; p < B; ++p) diff[p] = q_remaining[p] - q_unallocated[p];
            if (fabs(total_q<B>(diff)) < kNumTol) { // "Unallocated Q remains after distribution": the scenario fails
                ok = false;
                break;
            }
            for (int p = 0; p < B; ++p) q_remaining[p] = q_unallocated[p];
        }
    } else {
        for (int k = 0; k < n_regulating; ++k) {
            reg_flag[k] = (int8_t)bus_limit;
            double const limit_value = __ldg(s.reg_param + 4 * st_reg[k] + (bus_limit == 2 ? 3 : 2));
            double const q_limit_scalar = isnan(limit_value) ? 0.0 : limit_value;
            distribute_q<B>(q_limit_scalar, base_q[k], st_q[k]);
        }
    }
    return ok;
}

template <int B>
__global__ void regulator_result_kernel(DevStructure s, DevBatch b, int T, int32_t const* reg_bus, int n_reg_bus,
                                        double const* out_u, double const* out_inj, double* out_lg, int8_t* out_reg) {
    constexpr int c2 = 2 * B;
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_reg_bus * b.n_scn) return;
    int64_t const scn = idx / n_reg_bus;
    int const bus = __ldg(reg_bus + idx % n_reg_bus);
    int64_t const tile = scn / T;
    int const lane = (int)(scn % T);
    uint8_t const* lg_status = b.lg_status + (size_t)tile * s.n_load_gen * T + lane;
    int const bus_limit = b.qviol[((size_t)tile * s.n_bus + bus) * T + lane];
    double* const lg_base = out_lg + scn * s.n_load_gen * 2 * c2;

    // 1. regulator outputs and the set of regulating generators
    int st_lg[kMaxRegPerBus], st_reg[kMaxRegPerBus];
    double st_q[kMaxRegPerBus][B], base_q[kMaxRegPerBus][B];
    int n_regulating = 0;
    C s_other[B];
    for (int p = 0; p < B; ++p) s_other[p] = C{0.0, 0.0};
    for (int lg = __ldg(s.lg_ptr + bus), lge = __ldg(s.lg_ptr + bus + 1); lg < lge; ++lg) {
        double const* so = lg_base + (size_t)lg * 2 * c2;
        int const reg = __ldg(s.lg_reg + lg);
        bool regulating = false;
        if (reg >= 0) {
            int8_t* o = out_reg + (scn * s.n_regulator + reg) * 2;
            o[0] = 0;
            o[1] = (int8_t)lg_status[(size_t)lg * T];
            regulating = lg_status[(size_t)lg * T] != 0 && __ldg(s.reg_param + 4 * reg) != 0.0;
        }
        if (!regulating) {
            for (int p = 0; p < B; ++p) s_other[p] = cadd(s_other[p], C{so[2 * p], so[2 * p + 1]});
            continue;
        }
        if (n_regulating < kMaxRegPerBus) {
            st_lg[n_regulating] = lg;
            st_reg[n_regulating] = reg;
            for (int p = 0; p < B; ++p) {
                st_q[n_regulating][p] = 0.0;
                base_q[n_regulating][p] = so[2 * p + 1];
            }
            ++n_regulating;
        }
    }
    if (n_regulating == 0) return;
    // 2. distribution under the regulator limits
    double q_remaining[B];
    for (int p = 0; p < B; ++p) q_remaining[p] = out_inj[(scn * s.n_bus + bus) * c2 + 2 * p + 1] - s_other[p].i;
    int8_t flag[kMaxRegPerBus];
    if (!allocate_bus_q<B>(s, bus_limit, n_regulating, st_reg, st_q, base_q, q_remaining, flag)) b.status[scn] = 4;
    // 3. the generators take the allocated Q
    for (int k = 0; k < n_regulating; ++k) {
        out_reg[(scn * s.n_regulator + st_reg[k]) * 2] = flag[k];
        double* so = lg_base + (size_t)st_lg[k] * 2 * c2;
        for (int p = 0; p < B; ++p) {
            C const sv{so[2 * p], st_q[k][p]};
            C const u{out_u[(scn * s.n_bus + bus) * c2 + 2 * p], out_u[(scn * s.n_bus + bus) * c2 + 2 * p + 1]};
            C const i = conj(cdiv(sv, u));
            so[2 * p + 1] = sv.i;
            so[c2 + 2 * p] = i.r;
            so[c2 + 2 * p + 1] = i.i;
        }
    }
}

// The same step on the tile layout, for the device pipeline of the model level: the allocated Q REPLACES the specified Q of the
// regulating generators in the batch's injection buffer (they are const_pq, so their result power is exactly that value); the
// output kernels that run afterwards (load_gen / node / source results) then report it without knowing about regulators.
// out_reg [n_scn][n_regulator][2] = limit_violated, generator_status.
template <int T, int B>
__global__ void regulator_apply_kernel(DevStructure s, DevBatch b, int32_t const* reg_bus, int n_reg_bus, int8_t* out_reg) {
    constexpr int c2 = 2 * B;
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_reg_bus * b.n_scn) return;
    int64_t const scn = idx / n_reg_bus;
    int const bus = __ldg(reg_bus + idx % n_reg_bus);
    int64_t const tile = scn / T;
    int const lane = (int)(scn % T);
    uint8_t const* lg_status = b.lg_status + (size_t)tile * s.n_load_gen * T + lane;
    double* const sinj = b.sinj + (size_t)tile * s.n_load_gen * c2 * T + lane;
    int const bus_limit = b.qviol[((size_t)tile * s.n_bus + bus) * T + lane];
    int st_lg[kMaxRegPerBus], st_reg[kMaxRegPerBus];
    double st_q[kMaxRegPerBus][B], base_q[kMaxRegPerBus][B];
    int n_regulating = 0;
    C s_other[B];
    for (int p = 0; p < B; ++p) s_other[p] = C{0.0, 0.0};
    C u[B], inj[B];
    if constexpr (B == 1) {
        res::UView<T> const uv{b.u, s.n_bus};
        u[0] = uv.get(scn, bus);
        inj[0] = res::bus_injection<T>(s, uv, scn, bus, b.ovl);
    } else {
        res3::UView3<T> const uv{b.u, s.n_bus};
        res3::V3 const u3 = uv.get(scn, bus), i3 = res3::bus_injection3<T>(s, uv, scn, bus, b.ovl);
        for (int p = 0; p < B; ++p) {
            u[p] = u3.v[p];
            inj[p] = i3.v[p];
        }
    }
    for (int lg = __ldg(s.lg_ptr + bus), lge = __ldg(s.lg_ptr + bus + 1); lg < lge; ++lg) {
        int const reg = __ldg(s.lg_reg + lg);
        bool regulating = false;
        if (reg >= 0) {
            int8_t* o = out_reg + (scn * s.n_regulator + reg) * 2;
            o[0] = 0;
            o[1] = (int8_t)lg_status[(size_t)lg * T];
            regulating = lg_status[(size_t)lg * T] != 0 && __ldg(s.reg_param + 4 * reg) != 0.0;
        }
        int const type = __ldg(s.lg_type + lg);
        for (int p = 0; p < B; ++p) {
            C const x{sinj[(size_t)(lg * c2 + 2 * p) * T], sinj[(size_t)(lg * c2 + 2 * p + 1) * T]};
            C sv = x; // calculate_load_gen_result (common_solver_functions.hpp:143-160)
            if (type == 1) sv = cscale(x, u[p].r * u[p].r + u[p].i * u[p].i);
            if (type == 2) sv = cscale(x, sqrt(u[p].r * u[p].r + u[p].i * u[p].i));
            if (!regulating) {
                s_other[p] = cadd(s_other[p], sv);
            } else if (n_regulating < kMaxRegPerBus) {
                st_q[n_regulating][p] = 0.0;
                base_q[n_regulating][p] = sv.i;
            }
        }
        if (regulating && n_regulating < kMaxRegPerBus) {
            st_lg[n_regulating] = lg;
            st_reg[n_regulating] = reg;
            ++n_regulating;
        }
    }
    if (n_regulating == 0) return;
    double q_remaining[B];
    for (int p = 0; p < B; ++p) q_remaining[p] = inj[p].i - s_other[p].i;
    int8_t flag[kMaxRegPerBus];
    if (!allocate_bus_q<B>(s, bus_limit, n_regulating, st_reg, st_q, base_q, q_remaining, flag)) b.status[scn] = 4;
    for (int k = 0; k < n_regulating; ++k) {
        out_reg[(scn * s.n_regulator + st_reg[k]) * 2] = flag[k];
        for (int p = 0; p < B; ++p) sinj[(size_t)(st_lg[k] * c2 + 2 * p + 1) * T] = st_q[k][p];
    }
}

// VoltageRegulatorOutput (8 bytes: id, energized, limit_violated) of every regulator component and scenario
__global__ void pack_regulator_kernel(int64_t n_scn, int n_comp, int n_regulator, int32_t const* id, int32_t const* math,
                                      uint8_t const* status, int8_t const* reg_out, int32_t* out) {
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_scn * n_comp) return;
    int64_t const scn = idx / n_comp;
    int const c = (int)(idx % n_comp);
    int const m = __ldg(math + c);
    int energized = 0, violated = 0;
    if (m >= 0) {
        int8_t const* v = reg_out + (scn * n_regulator + m) * 2;
        energized = (__ldg(status + c) != 0 && v[1] != 0) ? 1 : 0;
        violated = v[0];
    }
    out[2 * idx] = __ldg(id + c);
    out[2 * idx + 1] = (energized & 0xff) | ((violated & 0xff) << 8);
}

// load_gen status [n_scn][n_item] -> tile layout [tile][n_item][T]
__global__ void status_to_tile_kernel(uint8_t const* src, uint8_t* dst, int64_t n_scn, int n_item, int T) {
    int64_t const idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t const n_tile = (n_scn + T - 1) / T;
    if (idx >= n_tile * n_item * T) return;
    int const lane = (int)(idx % T);
    int64_t const item = (idx / T) % n_item, tile = idx / ((int64_t)T * n_item);
    int64_t const scn = tile * T + lane;
    dst[idx] = scn < n_scn ? src[scn * n_item + item] : 0;
}

} // namespace

void launch_status_to_tile(int tile_width, uint8_t const* src, uint8_t* dst, int64_t n_scn, int n_item, cudaStream_t st) {
    int64_t const n_tile = (n_scn + tile_width - 1) / tile_width;
    int64_t const total = n_tile * n_item * tile_width;
    if (total == 0) return;
    count_kernel_launch();
    status_to_tile_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, dst, n_scn, n_item, tile_width);
}

void launch_regulator_apply(int phases, int tile_width, DevStructure const& s, DevBatch const& b, int32_t const* reg_bus,
                            int n_reg_bus, int8_t* out_reg, cudaStream_t st) {
    int64_t const total = (int64_t)n_reg_bus * b.n_scn;
    if (total == 0) return;
    count_kernel_launch();
    unsigned const grid = (unsigned)((total + 127) / 128);
#define PGMB_LAUNCH(TW)                                                                               \
    if (phases == 1) {                                                                                \
        regulator_apply_kernel<TW, 1><<<grid, 128, 0, st>>>(s, b, reg_bus, n_reg_bus, out_reg);       \
    } else {                                                                                          \
        regulator_apply_kernel<TW, 3><<<grid, 128, 0, st>>>(s, b, reg_bus, n_reg_bus, out_reg);       \
    }
    switch (tile_width) {
    case 4: PGMB_LAUNCH(4); break;
    case 8: PGMB_LAUNCH(8); break;
    case 16: PGMB_LAUNCH(16); break;
    default: PGMB_LAUNCH(32); break;
    }
#undef PGMB_LAUNCH
}

void launch_pack_regulator(int64_t n_scn, int n_comp, int n_regulator, int32_t const* id, int32_t const* math, uint8_t const* status,
                           int8_t const* reg_out, void* out, cudaStream_t st) {
    int64_t const total = n_scn * n_comp;
    if (total == 0) return;
    count_kernel_launch();
    pack_regulator_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(n_scn, n_comp, n_regulator, id, math, status, reg_out,
                                                                          static_cast<int32_t*>(out));
}

void launch_regulator_result(int phases, int tile_width, DevStructure const& s, DevBatch const& b, int32_t const* reg_bus,
                             int n_reg_bus, double const* out_u, double const* out_inj, double* out_lg, int8_t* out_reg,
                             cudaStream_t st) {
    int64_t const total = (int64_t)n_reg_bus * b.n_scn;
    if (total == 0) return;
    count_kernel_launch();
    int const block = 128;
    unsigned const grid = (unsigned)((total + block - 1) / block);
    if (phases == 1) {
        regulator_result_kernel<1><<<grid, block, 0, st>>>(s, b, tile_width, reg_bus, n_reg_bus, out_u, out_inj, out_lg, out_reg);
    } else {
        regulator_result_kernel<3><<<grid, block, 0, st>>>(s, b, tile_width, reg_bus, n_reg_bus, out_u, out_inj, out_lg, out_reg);
    }
}

} // namespace pgmb
