#include "engine.hpp"

#include <atomic>
#include <map>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace pgmb {

namespace {
template <class To, class From> std::vector<To> narrow_vec(std::vector<From> const& v) {
    std::vector<To> out(v.size());
    for (size_t i = 0; i != v.size(); ++i) out[i] = static_cast<To>(v[i]);
    return out;
}
std::atomic<uint64_t> g_kernel_launches{0};

int env_int(char const* name, int fallback) {
    char const* v = std::getenv(name);
    return (v != nullptr && *v != '\0') ? std::atoi(v) : fallback;
}
} // namespace

namespace {
struct PoolState {
    std::mutex mutex;
    std::map<std::pair<int, size_t>, std::vector<void*>> free_blocks; // (device, size class) -> blocks
    size_t free_bytes{0};
};
PoolState& pool() {
    static PoolState* p = new PoolState; // never destroyed: buffers may be released during static destruction
    return *p;
}
constexpr size_t kPoolKeepBytes = size_t{8} << 30;
} // namespace

void* DevPool::alloc(size_t bytes) {
    size_t const cls = size_class(bytes);
    int dev = 0;
    PGMB_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(pool().mutex);
        auto it = pool().free_blocks.find({dev, cls});
        if (it != pool().free_blocks.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            pool().free_bytes -= cls;
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t err = cudaMalloc(&p, cls);
    if (err != cudaSuccess) {
        cudaGetLastError();
        trim();
        err = cudaMalloc(&p, cls);
    }
    if (err != cudaSuccess) throw CudaError(std::string("cudaMalloc of ") + std::to_string(cls) + " bytes failed: " + cudaGetErrorString(err));
    return p;
}

void DevPool::release(void* p, size_t bytes) {
    size_t const cls = size_class(bytes);
    cudaPointerAttributes attr{};
    int dev = 0;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) dev = attr.device;
    bool too_much = false;
    {
        std::lock_guard<std::mutex> lock(pool().mutex);
        pool().free_blocks[{dev, cls}].push_back(p);
        pool().free_bytes += cls;
        too_much = pool().free_bytes > kPoolKeepBytes;
    }
    if (too_much) trim();
}

void DevPool::trim() {
    std::map<std::pair<int, size_t>, std::vector<void*>> blocks;
    {
        std::lock_guard<std::mutex> lock(pool().mutex);
        blocks.swap(pool().free_blocks);
        pool().free_bytes = 0;
    }
    for (auto& [key, list] : blocks)
        for (void* p : list) cudaFree(p);
}

void count_kernel_launch() { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); }
uint64_t kernel_launch_count() { return g_kernel_launches.load(std::memory_order_relaxed); }

Engine::Engine(MathTopology topo, bool symmetric, int device)
    : topo_{std::move(topo)}, symmetric_{symmetric}, B_{symmetric ? 1 : 3}, device_{device}, pattern_{topo_}, schedule_{pattern_}, program_{pattern_, schedule_, topo_},
      path_program_{pattern_, schedule_, topo_, program_},
      wide_plan_{pattern_, schedule_} {
    if (device < 0) return; // symbolic-only engine (structure introspection on hosts without a GPU); it cannot run
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        throw CudaError("no CUDA device available: pgm_b200 has no CPU fallback");
    }
    if (device >= n_dev) throw InvalidArgument("device index out of range");
    if (pattern_.nnz_lu >= (Idx{1} << 30)) throw InvalidArgument("LU pattern too large for 32-bit device indices");
    PGMB_CUDA(cudaSetDevice(device_));
    PGMB_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    PGMB_CUDA(cudaEventCreate(&ev0_));
    PGMB_CUDA(cudaEventCreate(&ev1_));
    upload_structure();
}

Engine::~Engine() {
    if (device_ < 0) return;
    cudaSetDevice(device_);
    if (stream_ != nullptr) cudaStreamSynchronize(stream_); // the buffers go back to the pool: nothing may still use them
    if (ev0_ != nullptr) cudaEventDestroy(ev0_);
    if (ev1_ != nullptr) cudaEventDestroy(ev1_);
    if (stream_ != nullptr) cudaStreamDestroy(stream_);
}

void Engine::upload_structure() {
    Idx const n_bus = topo_.n_bus;
    d_row_ptr_.upload(narrow_vec<int32_t>(pattern_.row_indptr_lu), stream_);
    d_col_idx_.upload(narrow_vec<int32_t>(pattern_.col_indices_lu), stream_);
    d_diag_.upload(narrow_vec<int32_t>(pattern_.diag_lu), stream_);
    d_map_y_.upload(narrow_vec<int32_t>(pattern_.map_lu_y_bus), stream_);
    d_level_ptr_.upload(schedule_.level_ptr, stream_);
    d_level_rows_.upload(schedule_.level_rows, stream_);
    d_upd_ptr_.upload(schedule_.upd_ptr, stream_);
    d_upd_u_.upload(schedule_.upd_u, stream_);
    d_upd_a_.upload(schedule_.upd_a, stream_);
    d_lg_ptr_.upload(narrow_vec<int32_t>(topo_.load_gens_per_bus), stream_);
    d_src_ptr_.upload(narrow_vec<int32_t>(topo_.sources_per_bus), stream_);
    d_prog_.upload(program_.words, stream_);
    if (path_program_.valid) d_path_prog_.upload(path_program_.words, stream_);
    d_wide_level_ptr_.upload(wide_plan_.level_ptr, stream_);
    d_wide_table_.upload(wide_plan_.table, stream_);
    d_wide_data_.upload(wide_plan_.data, stream_);
    d_row_is_wide_.upload(wide_plan_.is_wide, stream_);
    d_lg_type_.upload(topo_.load_gen_type, stream_);
    d_y_row_ptr_.upload(narrow_vec<int32_t>(pattern_.row_indptr), stream_);
    d_y_col_idx_.upload(narrow_vec<int32_t>(pattern_.col_indices), stream_);
    d_branch_bus_.upload(narrow_vec<int32_t>(topo_.branch_bus_idx), stream_);
    auto group_of = [n_bus](std::vector<Idx> const& indptr) {
        std::vector<int32_t> out(indptr.empty() ? 0 : indptr.back());
        for (Idx b = 0; b != n_bus; ++b)
            for (Idx k = indptr[b]; k != indptr[b + 1]; ++k) out[k] = static_cast<int32_t>(b);
        return out;
    };
    d_shunt_bus_.upload(group_of(topo_.shunts_per_bus), stream_);
    d_lg_bus_.upload(group_of(topo_.load_gens_per_bus), stream_);
    d_src_bus_.upload(group_of(topo_.sources_per_bus), stream_);
    d_phase_shift_.upload(topo_.phase_shift, stream_);
    std::vector<int32_t> reg_bus;
    if (has_regulators()) {
        if (static_cast<Idx>(topo_.load_gen_regulator.size()) != topo_.n_load_gen()) {
            throw InvalidArgument("load_gen_regulator must have one entry per load_gen");
        }
        d_lg_reg_.upload(narrow_vec<int32_t>(topo_.load_gen_regulator), stream_);
        for (Idx b = 0; b != n_bus; ++b) {
            int n_reg = 0;
            for (Idx k = topo_.load_gens_per_bus[b]; k != topo_.load_gens_per_bus[b + 1]; ++k) n_reg += topo_.load_gen_regulator[k] >= 0 ? 1 : 0;
            if (n_reg > 16) throw InvalidArgument("more than 16 regulated generators on one bus are not supported");
            if (n_reg != 0) reg_bus.push_back(static_cast<int32_t>(b));
        }
        d_reg_bus_.upload(reg_bus, stream_);
        n_reg_bus_ = static_cast<int>(reg_bus.size());
    }
    PGMB_CUDA(cudaStreamSynchronize(stream_)); // temporaries above must outlive the copies

    ds_.n_bus = static_cast<int32_t>(n_bus);
    ds_.nnz = static_cast<int32_t>(pattern_.nnz);
    ds_.nnz_lu = static_cast<int32_t>(pattern_.nnz_lu);
    ds_.n_level = schedule_.n_level();
    ds_.n_load_gen = static_cast<int32_t>(topo_.n_load_gen());
    ds_.n_source = static_cast<int32_t>(topo_.n_source());
    ds_.n_branch = static_cast<int32_t>(topo_.n_branch());
    ds_.n_shunt = static_cast<int32_t>(topo_.n_shunt());
    ds_.row_ptr = d_row_ptr_.get();
    ds_.col_idx = d_col_idx_.get();
    ds_.diag = d_diag_.get();
    ds_.map_y = d_map_y_.get();
    ds_.level_ptr = d_level_ptr_.get();
    ds_.level_rows = d_level_rows_.get();
    ds_.upd_ptr = d_upd_ptr_.get();
    ds_.upd_u = d_upd_u_.get();
    ds_.upd_a = d_upd_a_.get();
    ds_.lg_ptr = d_lg_ptr_.get();
    ds_.lg_type = d_lg_type_.get();
    ds_.src_ptr = d_src_ptr_.get();
    ds_.y_row_ptr = d_y_row_ptr_.get();
    ds_.y_col_idx = d_y_col_idx_.get();
    ds_.branch_bus = d_branch_bus_.get();
    ds_.shunt_bus = d_shunt_bus_.get();
    ds_.lg_bus = d_lg_bus_.get();
    ds_.src_bus = d_src_bus_.get();
    ds_.phase_shift = d_phase_shift_.get();
    ds_.prog = d_prog_.get();
    ds_.prog_words = static_cast<int32_t>(program_.words.size());
    ds_.wide_level_ptr = d_wide_level_ptr_.get();
    ds_.wide_table = d_wide_table_.get();
    ds_.wide_data = d_wide_data_.get();
    ds_.row_is_wide = d_row_is_wide_.get();
    ds_.n_wide = env_int("PGMB_WIDE", 1) != 0 ? wide_plan_.n_wide() : 0; // PGMB_WIDE=0: every row by one thread (cross-check)
    ds_.wide_max_upd = wide_plan_.max_upd;
    ds_.wide_max_lower = wide_plan_.max_lower;
    ds_.wide_max_entries = wide_plan_.max_entries;
    ds_.path_prog = path_program_.valid ? d_path_prog_.get() : nullptr;
    ds_.path_prog_words = path_program_.valid ? static_cast<int32_t>(path_program_.words.size()) : 0;
    ds_.path_prog_smem_words = path_program_.valid ? path_program_.smem_words : 0;
    ds_.n_regulator = static_cast<int32_t>(topo_.n_voltage_regulator());
    ds_.lg_reg = has_regulators() ? d_lg_reg_.get() : nullptr;
    ds_.reg_param = nullptr;
}

// VoltageRegulatorCalcParam of every regulator (calculation_parameters.hpp:228-236).  A regulating regulator needs a
// const_pq generator (newton_raphson_pf_solver.hpp:621-625; the reference throws when it reaches the limit check).
void Engine::set_regulators(double const* param) {
    if (!has_regulators()) throw InvalidArgument("the math topology has no voltage regulators");
    if (param == nullptr) throw InvalidArgument("voltage regulator parameters are required for a grid with regulators");
    Idx const n_reg = topo_.n_voltage_regulator();
    reg_param_.assign(param, param + 4 * n_reg);
    for (Idx lg = 0; lg != topo_.n_load_gen(); ++lg) {
        Idx const r = topo_.load_gen_regulator[lg];
        if (r >= 0 && reg_param_[4 * r] != 0.0 && topo_.load_gen_type[lg] != 0) {
            throw InvalidArgument("Voltage regulator(s) " + std::to_string(r) + " regulate(s) a load/generator with unsupported type");
        }
    }
    reg_param_set_ = true;
    if (device_ < 0) return;
    PGMB_CUDA(cudaSetDevice(device_));
    d_reg_param_.ensure(reg_param_.size() + 1);
    PGMB_CUDA(cudaMemcpyAsync(d_reg_param_.get(), reg_param_.data(), reg_param_.size() * sizeof(double), cudaMemcpyHostToDevice, stream_));
    PGMB_CUDA(cudaStreamSynchronize(stream_));
    ds_.reg_param = d_reg_param_.get();
}

// YBus::update_admittance_entries (y_bus.hpp:400-431): sum of the contributions of each entry, in element order
void Engine::set_param(double const* branch_param, double const* shunt_param, double const* source_param) {
    if (device_ >= 0) PGMB_CUDA(cudaSetDevice(device_));
    int const bb2 = B_ * B_ * 2;
    branch_param_.assign(branch_param, branch_param + topo_.n_branch() * 4 * bb2);
    shunt_param_.assign(shunt_param, shunt_param + topo_.n_shunt() * bb2);
    source_param_.assign(source_param, source_param + topo_.n_source() * 4);
    admittance_.assign(pattern_.nnz * bb2, 0.0);
    for (Idx entry = 0; entry != pattern_.nnz; ++entry) {
        double* y = &admittance_[entry * bb2];
        for (Idx e = pattern_.y_bus_entry_indptr[entry]; e != pattern_.y_bus_entry_indptr[entry + 1]; ++e) {
            int const kind = pattern_.element_type[e];
            double const* src = kind == 4 ? &shunt_param_[pattern_.element_idx[e] * bb2]
                                          : &branch_param_[(pattern_.element_idx[e] * 4 + kind) * bb2];
            for (int i = 0; i != bb2; ++i) y[i] += src[i];
        }
    }
    // y_ref tensor of each source (SourceCalcParam::y_ref, calculation_parameters.hpp:215-227)
    std::vector<double> yref(topo_.n_source() * bb2, 0.0);
    for (Idx s = 0; s != topo_.n_source(); ++s) {
        std::complex<double> const y1{source_param_[4 * s], source_param_[4 * s + 1]};
        std::complex<double> const y0{source_param_[4 * s + 2], source_param_[4 * s + 3]};
        if (B_ == 1) {
            yref[2 * s] = y1.real();
            yref[2 * s + 1] = y1.imag();
        } else {
            std::complex<double> const ys = (2.0 * y1 + y0) / 3.0;
            std::complex<double> const ym = (y0 - y1) / 3.0;
            for (int r = 0; r != 3; ++r)
                for (int c = 0; c != 3; ++c) {
                    std::complex<double> const v = r == c ? ys : ym;
                    yref[s * bb2 + (r * 3 + c) * 2] = v.real();
                    yref[s * bb2 + (r * 3 + c) * 2 + 1] = v.imag();
                }
        }
    }
    param_set_ = true;
    if (device_ < 0) return;
    d_ydata_.upload(admittance_, stream_);
    d_src_yref_.upload(yref, stream_);
    d_src_y1y0_.upload(source_param_, stream_);
    d_branch_param_.upload(branch_param_, stream_);
    d_shunt_param_.upload(shunt_param_, stream_);
    PGMB_CUDA(cudaStreamSynchronize(stream_));
    ds_.ydata = d_ydata_.get();
    ds_.src_yref = d_src_yref_.get();
    ds_.src_y1y0 = d_src_y1y0_.get();
    ds_.branch_param = d_branch_param_.get();
    ds_.shunt_param = d_shunt_param_.get();
    ic_factor_valid_ = false; // parameters_changed (iterative_current_pf_solver.hpp:162)
    param_set_ = true;
}

void Engine::choose_tiling(int64_t n_scn) {
    static int sm_cache[64] = {};
    if (sm_cache[device_ % 64] == 0) PGMB_CUDA(cudaDeviceGetAttribute(&sm_cache[device_ % 64], cudaDevAttrMultiProcessorCount, device_));
    int const sm = sm_cache[device_ % 64];
    int t = 4;
    for (int cand : {32, 16, 8, 4}) {
        if ((n_scn + cand - 1) / cand >= (3 * sm) / 4) {
            t = cand;
            break;
        }
    }
    // Iterative-current batches (two resident blocks per SM): a tile of width T costs about (9 + T) units whatever the batch
    // (measured: 12 500 scenarios 14.4 ms at T = 32 in 2 waves, 13.1 ms at T = 16 in 3 waves), so take the width with the
    // cheapest number of waves -- it avoids a thin last wave of wide tiles.
    if (symmetric_ && (method_hint_ == 3 || method_hint_ == 4)) {
        double best = 0.0;
        for (int cand : {32, 16, 8}) {
            int64_t const tiles = (n_scn + cand - 1) / cand;
            if (tiles < (3 * sm) / 4 && cand != 8) continue;
            double const cost = static_cast<double>((tiles + 2 * sm - 1) / (2 * sm)) * (9.0 + cand);
            if (best == 0.0 || cost < best) {
                best = cost;
                t = cand;
            }
        }
    }
    tile_width_ = env_int("PGMB_TILE", t);
    if (tile_width_ != 4 && tile_width_ != 8 && tile_width_ != 16 && tile_width_ != 32) tile_width_ = t;
    // generic-block kernel (asymmetric): 255 registers per thread, at most 256 threads per block (measured best)
    n_slot_ = env_int("PGMB_SLOTS", (symmetric_ ? 512 : 256) / tile_width_);
    if (n_slot_ < 1) n_slot_ = 1;
    if (n_slot_ * tile_width_ > 1024) n_slot_ = 1024 / tile_width_;
}

void Engine::allocate_batch(int64_t n) {
    if (device_ < 0) throw CudaError("engine was created without a CUDA device (symbolic only): pgm_b200 has no CPU fallback");
    if (!param_set_) throw InvalidArgument("pgmb_engine_set_param must be called before running");
    PGMB_CUDA(cudaSetDevice(device_));
    choose_tiling(n);
    int const T = tile_width_;
    int64_t const n_tile = (n + T - 1) / T;
    int const N = 2 * B_;
    size_t const nb = static_cast<size_t>(topo_.n_bus);
    d_jac_.ensure(static_cast<size_t>(n_tile) * pattern_.nnz_lu * N * N * T);
    d_xvec_.ensure(n_tile * nb * N * T);
    d_pol_.ensure(n_tile * nb * N * T);
    d_u_.ensure(n_tile * nb * N * T);
    if (symmetric_ && path_program_.valid) d_side_.ensure(n_tile * nb * N * T);
    if (wide_plan_.n_wide() != 0) {
        d_wide_terms_.ensure(static_cast<size_t>(n_tile) * wide_plan_.max_upd * N * N * T + 1);
        d_wide_rhs_.ensure(static_cast<size_t>(n_tile) * wide_plan_.max_lower * N * T + 1);
        d_wide_sum_.ensure(static_cast<size_t>(n_tile) * wide_plan_.max_entries * N * T + 1);
    }
    d_perm_.ensure(n_tile * nb * T * 2 * N);
    d_sinj_.ensure(static_cast<size_t>(n_tile) * topo_.n_load_gen() * 2 * B_ * T + 1);
    d_lg_status_.ensure(static_cast<size_t>(n_tile) * topo_.n_load_gen() * T + 1);
    d_usrc_.ensure(static_cast<size_t>(n_tile) * topo_.n_source() * 2 * T + 1);
    if (has_regulators()) d_qviol_.ensure(n_tile * nb * T + 1);
    d_status_.ensure(n + 1);
    d_n_iter_.ensure(n + 1);
    d_max_dev_.ensure(n + 1);
    db_.n_scn = n;
    db_.n_tile = static_cast<int32_t>(n_tile);
    db_.jac = d_jac_.get();
    db_.xvec = d_xvec_.get();
    db_.pol = d_pol_.get();
    db_.u = d_u_.get();
    db_.side = d_side_.get();
    db_.wide_terms = d_wide_terms_.get();
    db_.wide_rhs = d_wide_rhs_.get();
    db_.wide_sum = d_wide_sum_.get();
    db_.perm = d_perm_.get();
    db_.sinj = d_sinj_.get();
    db_.usrc = d_usrc_.get();
    db_.status = d_status_.get();
    db_.n_iter = d_n_iter_.get();
    db_.max_dev = d_max_dev_.get();
    db_.lg_status = d_lg_status_.get();
    db_.qviol = has_regulators() ? d_qviol_.get() : nullptr;
    db_.ovl = DevOverlay{};
    db_.phase_cycles = nullptr;
    if (env_int("PGMB_DEBUG_PHASES", 0) != 0) {
        d_phase_.ensure(static_cast<size_t>(n_tile) * 16);
        db_.phase_cycles = d_phase_.get();
    }
}

void Engine::stage(PfInputView const& in) {
    int64_t const n = in.n_scenarios;
    method_hint_ = in.method_hint;
    allocate_batch(n);
    int const T = tile_width_;
    size_t const n_sinj = static_cast<size_t>(n) * topo_.n_load_gen() * 2 * B_;
    size_t const n_usrc = static_cast<size_t>(in.source_is_shared ? 1 : n) * topo_.n_source() * 2;
    d_in_sinj_.ensure(n_sinj + 1);
    d_in_usrc_.ensure(n_usrc + 1);
    if (n_sinj != 0) PGMB_CUDA(cudaMemcpyAsync(d_in_sinj_.get(), in.s_injection, n_sinj * sizeof(double), cudaMemcpyHostToDevice, stream_));
    if (n_usrc != 0) PGMB_CUDA(cudaMemcpyAsync(d_in_usrc_.get(), in.source_u_ref, n_usrc * sizeof(double), cudaMemcpyHostToDevice, stream_));
    // host layout per load_gen: [B] complex = (re, im) interleaved; tile layout wants re[B], im[B]: for B = 1 identical
    launch_to_tile(T, d_in_sinj_.get(), d_sinj_.get(), n, static_cast<int>(topo_.n_load_gen()), 2 * B_, 0, stream_);
    launch_to_tile(T, d_in_usrc_.get(), d_usrc_.get(), n, static_cast<int>(topo_.n_source()), 2, in.source_is_shared ? 1 : 0, stream_);
    PGMB_CUDA(cudaMemsetAsync(d_lg_status_.get(), 1, d_lg_status_.size(), stream_));
    if (has_regulators()) {
        set_regulators(in.voltage_regulator);
        if (in.load_gen_status != nullptr && n * topo_.n_load_gen() != 0) {
            size_t const bytes = static_cast<size_t>(n) * topo_.n_load_gen();
            d_in_lg_status_.ensure(bytes);
            PGMB_CUDA(cudaMemcpyAsync(d_in_lg_status_.get(), in.load_gen_status, bytes, cudaMemcpyHostToDevice, stream_));
            launch_status_to_tile(T, d_in_lg_status_.get(), d_lg_status_.get(), n, static_cast<int>(topo_.n_load_gen()), stream_);
        }
    }
    PGMB_CUDA(cudaStreamSynchronize(stream_));
}

void Engine::stage_device(int64_t n, double const* source_u_ref, bool source_is_shared) {
    allocate_batch(n);
    size_t const n_usrc = static_cast<size_t>(source_is_shared ? 1 : n) * topo_.n_source() * 2;
    d_in_usrc_.ensure(n_usrc + 1);
    if (n_usrc != 0) PGMB_CUDA(cudaMemcpyAsync(d_in_usrc_.get(), source_u_ref, n_usrc * sizeof(double), cudaMemcpyHostToDevice, stream_));
    launch_to_tile(tile_width_, d_in_usrc_.get(), d_usrc_.get(), n, static_cast<int>(topo_.n_source()), 2, source_is_shared ? 1 : 0, stream_);
    PGMB_CUDA(cudaStreamSynchronize(stream_)); // source_u_ref may be a temporary of the caller
}

void Engine::launch_regulator_apply(DevBatch const& view, int8_t* out_reg, cudaStream_t st) {
    if (!has_regulators() || view.n_scn == 0) return;
    pgmb::launch_regulator_apply(B_, tile_width_, ds_, view, d_reg_bus_.get(), n_reg_bus_, out_reg, st);
    PGMB_CUDA(cudaGetLastError());
}

void Engine::set_overlay(int64_t n_scn, int64_t const* math_branch, double const* bparam, int32_t const* comp,
                         uint8_t const* energized, int32_t const* dead_off, uint8_t const* dead, size_t dead_bytes, int n_slot) {
    if (device_ < 0) throw CudaError("engine was created without a CUDA device (symbolic only)");
    if (n_scn != db_.n_scn) throw InvalidArgument("overlay size differs from the staged batch");
    if (n_slot < 1) throw InvalidArgument("overlay needs at least one branch slot per scenario");
    PGMB_CUDA(cudaSetDevice(device_));
    size_t const bb2 = static_cast<size_t>(B_) * B_ * 2;
    if (branch_entries_.empty()) {
        branch_entries_.assign(topo_.n_branch(), {-1, -1, -1, -1});
        for (Idx entry = 0; entry != pattern_.nnz; ++entry)
            for (Idx e = pattern_.y_bus_entry_indptr[entry]; e != pattern_.y_bus_entry_indptr[entry + 1]; ++e)
                if (pattern_.element_type[e] < 4) branch_entries_[pattern_.element_idx[e]][pattern_.element_type[e]] = static_cast<int32_t>(entry);
    }
    size_t const K = static_cast<size_t>(n_slot), E = 4 * K;
    std::vector<int32_t> entry(n_scn * E, -1), branch(n_scn * K, -1);
    std::vector<double> y(n_scn * E * bb2, 0.0);
    for (int64_t s = 0; s != n_scn; ++s) {
        int64_t const* const br = math_branch + s * K;
        double const* const np = bparam + s * K * 4 * bb2;
        size_t n_entry = 0;
        for (size_t j = 0; j != K; ++j) {
            if (br[j] < 0) continue;
            if (br[j] >= topo_.n_branch()) throw InvalidArgument("overlay branch out of range");
            branch[s * K + j] = static_cast<int32_t>(br[j]);
            for (int k = 0; k != 4; ++k) {
                int32_t const en = branch_entries_[br[j]][k];
                int32_t* const list = &entry[s * E];
                if (en < 0 || std::find(list, list + n_entry, en) != list + n_entry) continue; // shared with an earlier branch
                list[n_entry] = en;
                // the same sum as Engine::set_param, with the contributions of this scenario's branches replaced
                double* yo = &y[(s * E + n_entry) * bb2];
                ++n_entry;
                for (Idx e = pattern_.y_bus_entry_indptr[en]; e != pattern_.y_bus_entry_indptr[en + 1]; ++e) {
                    int const kind = pattern_.element_type[e];
                    Idx const idx = pattern_.element_idx[e];
                    double const* src = kind == 4 ? &shunt_param_[idx * bb2] : &branch_param_[(idx * 4 + kind) * bb2];
                    if (kind != 4)
                        for (size_t jj = 0; jj != K; ++jj)
                            if (br[jj] == idx) src = np + (jj * 4 + kind) * bb2;
                    for (size_t i = 0; i != bb2; ++i) yo[i] += src[i];
                }
            }
        }
    }
    d_ovl_entry_.upload(entry, stream_);
    d_ovl_branch_.upload(branch, stream_);
    d_ovl_y_.upload(y, stream_);
    d_ovl_bparam_.ensure(n_scn * K * 4 * bb2 + 1);
    d_ovl_comp_.ensure(n_scn * K + 1);
    d_ovl_energized_.ensure(n_scn * K + 1);
    if (n_scn != 0) {
        PGMB_CUDA(cudaMemcpyAsync(d_ovl_bparam_.get(), bparam, n_scn * K * 4 * bb2 * sizeof(double), cudaMemcpyHostToDevice, stream_));
        PGMB_CUDA(cudaMemcpyAsync(d_ovl_comp_.get(), comp, n_scn * K * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
        PGMB_CUDA(cudaMemcpyAsync(d_ovl_energized_.get(), energized, n_scn * K, cudaMemcpyHostToDevice, stream_));
    }
    bool const any_dead = dead_off != nullptr && dead != nullptr && dead_bytes != 0;
    if (any_dead) {
        d_ovl_dead_off_.ensure(n_scn + 1);
        d_ovl_dead_.ensure(dead_bytes);
        PGMB_CUDA(cudaMemcpyAsync(d_ovl_dead_off_.get(), dead_off, n_scn * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
        PGMB_CUDA(cudaMemcpyAsync(d_ovl_dead_.get(), dead, dead_bytes, cudaMemcpyHostToDevice, stream_));
    }
    PGMB_CUDA(cudaStreamSynchronize(stream_));
    db_.ovl = DevOverlay{d_ovl_entry_.get(), d_ovl_y_.get(), d_ovl_branch_.get(), d_ovl_bparam_.get(), d_ovl_comp_.get(), d_ovl_energized_.get(),
                         any_dead ? d_ovl_dead_off_.get() : nullptr, any_dead ? d_ovl_dead_.get() : nullptr, n_slot};
}

void Engine::fetch_status(int32_t* status, int32_t* n_iter) {
    PGMB_CUDA(cudaSetDevice(device_));
    if (db_.n_scn == 0) return;
    PGMB_CUDA(cudaMemcpyAsync(status, d_status_.get(), sizeof(int32_t) * db_.n_scn, cudaMemcpyDeviceToHost, stream_));
    PGMB_CUDA(cudaMemcpyAsync(n_iter, d_n_iter_.get(), sizeof(int32_t) * db_.n_scn, cudaMemcpyDeviceToHost, stream_));
    PGMB_CUDA(cudaStreamSynchronize(stream_));
}

DevBatch Engine::batch_view(int64_t tile_begin, int64_t tile_end) const {
    DevBatch v = db_;
    int64_t const T = tile_width_;
    int64_t const N = 2 * B_;
    size_t const nb = static_cast<size_t>(topo_.n_bus);
    int64_t const scn_begin = tile_begin * T;
    v.n_tile = static_cast<int32_t>(tile_end - tile_begin);
    v.n_scn = std::min<int64_t>(db_.n_scn, tile_end * T) - scn_begin;
    // per-tile strides are those of allocate_batch(); a kernel that uses a smaller stride stays inside its view's range
    v.jac += static_cast<size_t>(tile_begin) * pattern_.nnz_lu * N * N * T;
    v.xvec += tile_begin * nb * N * T;
    v.pol += tile_begin * nb * N * T;
    v.u += tile_begin * nb * N * T;
    if (v.side != nullptr) v.side += tile_begin * nb * N * T;
    if (v.wide_terms != nullptr) {
        v.wide_terms += static_cast<size_t>(tile_begin) * wide_plan_.max_upd * N * N * T;
        v.wide_rhs += static_cast<size_t>(tile_begin) * wide_plan_.max_lower * N * T;
        v.wide_sum += static_cast<size_t>(tile_begin) * wide_plan_.max_entries * N * T;
    }
    v.perm += tile_begin * nb * T * 2 * N;
    v.sinj += static_cast<size_t>(tile_begin) * topo_.n_load_gen() * 2 * B_ * T;
    v.lg_status += static_cast<size_t>(tile_begin) * topo_.n_load_gen() * T;
    if (v.qviol != nullptr) v.qviol += tile_begin * nb * T;
    v.usrc += static_cast<size_t>(tile_begin) * topo_.n_source() * 2 * T;
    v.status += scn_begin;
    v.n_iter += scn_begin;
    v.max_dev += scn_begin;
    if (v.ovl.entry != nullptr) {
        size_t const bb2 = static_cast<size_t>(B_) * B_ * 2;
        int64_t const K = v.ovl.n_branch;
        v.ovl.entry += scn_begin * 4 * K;
        v.ovl.y += scn_begin * 4 * K * bb2;
        v.ovl.branch += scn_begin * K;
        v.ovl.bparam += scn_begin * K * 4 * bb2;
        v.ovl.comp += scn_begin * K;
        v.ovl.energized += scn_begin * K;
        if (v.ovl.dead_off != nullptr) v.ovl.dead_off += scn_begin;
    }
    if (v.phase_cycles != nullptr) v.phase_cycles += tile_begin * 16;
    return v;
}

SolveOptions Engine::prepare_solve(SolveOptions const& opt_in) {
    if (device_ < 0) throw CudaError("engine was created without a CUDA device (symbolic only)");
    PGMB_CUDA(cudaSetDevice(device_));
    SolveOptions opt = opt_in;
    // all loads const_y => the reference forces the linear method (math_solver.hpp:36-37, 48)
    bool const all_const_y = std::all_of(topo_.load_gen_type.begin(), topo_.load_gen_type.end(), [](int8_t t) { return t == 1; });
    if (all_const_y) opt.method = 0;
    if (opt.method == -128) opt.method = 1;
    if (opt.method != 0 && opt.method != 1 && opt.method != 3 && opt.method != 4) {
        throw InvalidArgument("calculation method " + std::to_string(opt.method) + " is not a power-flow method");
    }
    if (has_regulators()) { // calculation_preparation.hpp:163-225: regulators exist only in the Newton-Raphson formulation
        if (opt_in.method != 1 && opt_in.method != -128) throw InvalidArgument("The calculation method is invalid for this calculation!");
        if (!reg_param_set_) throw InvalidArgument("voltage regulator parameters have not been set");
    }
    last_method_ = opt.method;
    if ((opt.method == 3 || opt.method == 4) && !ic_factor_valid_) {
        d_ic_factor_.ensure(static_cast<size_t>(pattern_.nnz_lu) * 2 * B_ * B_);
        d_ic_flag_.ensure(1);
        if (symmetric_) {
            launch_ic_factor(ds_, d_ic_factor_.get(), reinterpret_cast<int*>(d_ic_flag_.get()), stream_);
        } else {
            d_ic_perm_.ensure(static_cast<size_t>(topo_.n_bus) * 2 * B_ + 1);
            launch_ic_factor_asym(ds_, d_ic_factor_.get(), d_ic_perm_.get(), reinterpret_cast<int*>(d_ic_flag_.get()), stream_);
        }
        PGMB_CUDA(cudaGetLastError());
        ic_factor_valid_ = true;
    }
    if (opt.method == 4) { // linear_current = one iteration, no tolerance (math_solver.hpp:151-156)
        opt.err_tol = INFINITY;
        opt.max_iter = 1;
    }
    return opt;
}

void Engine::launch_solve(DevBatch const& b, SolveOptions const& opt, cudaStream_t st) {
    if (b.n_scn == 0) return;
    if (b.qviol != nullptr) {
        PGMB_CUDA(cudaMemsetAsync(b.qviol, 0, static_cast<size_t>(b.n_tile) * topo_.n_bus * tile_width_, st));
    }
    if (b.ovl.entry != nullptr && opt.method != 1) throw InvalidArgument("a branch-outage overlay needs the Newton-Raphson method");
    switch (opt.method) {
    case 1:
        // grids with voltage regulators (PV buses): the generic block kernel carries that logic for B = 1 and B = 3
        // (a branch-outage overlay is read by the block kernel and the symmetric level kernel)
        if (!symmetric_ && !has_regulators() && env_int("PGMB_BLOCK6", 0) != 0) {
            // row-split kernel: six threads per (bus row, scenario); see nr_block6.cu
            launch_nr_block6(tile_width_, ds_, b, opt, env_int("PGMB_BLOCK6_THREADS", 768), st);
        } else if (symmetric_ && has_regulators() && env_int("PGMB_KERNEL", 3) >= 2 && b.ovl.entry == nullptr && env_int("PGMB_REG_PATH", 1) != 0) {
            // symmetric grid with voltage regulators: the PV instantiations of the path kernel (radial grids) and the level
            // kernel (bit-identical to the block kernel with B = 1, which PGMB_REG_PATH=0 selects)
            if (path_program_.valid && env_int("PGMB_KERNEL", 3) == 3) {
                launch_nr_sym_v3(tile_width_, ds_, b, opt, n_slot_, st);
            } else {
                launch_nr_sym_v2(tile_width_, ds_, b, opt, n_slot_, st);
            }
        } else if (!symmetric_ || has_regulators() || env_int("PGMB_KERNEL", 3) == 0) {
            launch_nr_block(B_, tile_width_, ds_, b, opt, n_slot_, st);
        } else if (env_int("PGMB_KERNEL", 3) == 1 && b.ovl.entry == nullptr) {
            launch_nr_sym(tile_width_, ds_, b, opt, n_slot_, st);
        } else if (path_program_.valid && env_int("PGMB_KERNEL", 3) == 3 && b.ovl.entry == nullptr) {
            // radial grid: the path kernel (measured against the level kernel: 1000 scenarios 1.93 vs 2.39 ms, 8000 scenarios
            // 10.5 vs 11.1 ms)
            launch_nr_sym_v3(tile_width_, ds_, b, opt, n_slot_, st);
        } else {
            launch_nr_sym_v2(tile_width_, ds_, b, opt, n_slot_, st);
        }
        break;
    case 0:
        if (symmetric_) {
            launch_linear_sym(tile_width_, ds_, b, n_slot_, st);
        } else {
            launch_linear_asym(tile_width_, ds_, b, n_slot_, st);
        }
        break;
    default:
        if (symmetric_) {
            launch_ic_iterate_sym(tile_width_, ds_, b, opt, d_ic_factor_.get(), reinterpret_cast<int const*>(d_ic_flag_.get()), n_slot_, st);
        } else {
            launch_ic_iterate_asym(tile_width_, ds_, b, opt, d_ic_factor_.get(), d_ic_perm_.get(),
                                   reinterpret_cast<int const*>(d_ic_flag_.get()), n_slot_, st);
        }
        break;
    }
    PGMB_CUDA(cudaGetLastError());
}

float Engine::solve_staged(SolveOptions const& opt_in) {
    if (device_ < 0) throw CudaError("engine was created without a CUDA device (symbolic only)");
    PGMB_CUDA(cudaSetDevice(device_));
    if (db_.n_scn == 0) return 0.0f;
    SolveOptions const opt = prepare_solve(opt_in);
    if (db_.phase_cycles != nullptr) PGMB_CUDA(cudaMemsetAsync(db_.phase_cycles, 0, sizeof(unsigned long long) * db_.n_tile * 16, stream_));
    PGMB_CUDA(cudaEventRecord(ev0_, stream_));
    launch_solve(db_, opt, stream_);
    PGMB_CUDA(cudaGetLastError());
    PGMB_CUDA(cudaEventRecord(ev1_, stream_));
    PGMB_CUDA(cudaEventSynchronize(ev1_));
    float ms = 0.0f;
    PGMB_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
    if (db_.phase_cycles != nullptr) {
        std::vector<unsigned long long> h(static_cast<size_t>(db_.n_tile) * 16);
        PGMB_CUDA(cudaMemcpy(h.data(), db_.phase_cycles, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        double avg[16] = {};
        for (int t = 0; t != db_.n_tile; ++t)
            for (int k = 0; k != 16; ++k) avg[k] += static_cast<double>(h[t * 16 + k]) / db_.n_tile;
        // level kernels (v1 / v2): init up0 up_rest down_rest down0 ... | newton (+4); path kernel (v3): init leaf-build
        // inner-build up-chains down-chains down-leaves ... | newton (+8)
        std::fprintf(stderr, "[pgmb phases, kcycles/tile]");
        for (int k = 0; k != 16; ++k) std::fprintf(stderr, " %s%.0f", k == 8 ? "| " : "", avg[k] / 1e3);
        std::fprintf(stderr, "\n");
    }
    return ms;
}

void Engine::fetch(SolverOutputView const& out) {
    if (device_ < 0) throw CudaError("engine was created without a CUDA device (symbolic only)");
    PGMB_CUDA(cudaSetDevice(device_));
    int64_t const n = db_.n_scn;
    if (n == 0) return;
    int const c2 = 2 * B_;
    auto want = [&](double* host, DevBuf<double>& buf, size_t per_scn) -> double* {
        if (host == nullptr) return nullptr;
        buf.ensure(static_cast<size_t>(n) * per_scn + 1);
        return buf.get();
    };
    bool const reg = has_regulators();
    double dummy = 0.0; // the regulator step reads u, the bus injection and the load_gen results on the device
    double* const du = want(reg ? &dummy : out.u, d_out_u_, topo_.n_bus * c2);
    double* const di = want(reg ? &dummy : out.bus_injection, d_out_inj_, topo_.n_bus * c2);
    double* const dbr = want(out.branch, d_out_branch_, topo_.n_branch() * 4 * c2);
    double* const dsrc = want(out.source, d_out_source_, topo_.n_source() * 2 * c2);
    double* const dsh = want(out.shunt, d_out_shunt_, topo_.n_shunt() * 2 * c2);
    double* const dlg = want(reg ? &dummy : out.load_gen, d_out_lg_, topo_.n_load_gen() * 2 * c2);
    if (reg) d_out_reg_.ensure(static_cast<size_t>(n) * ds_.n_regulator * 2 + 1);
    if (symmetric_) {
        launch_math_result_sym(tile_width_, ds_, db_, last_method_ == 0 ? 1 : 0, du, di, dbr, dsrc, dsh, dlg, stream_);
    } else {
        launch_math_result_asym(tile_width_, ds_, db_, last_method_ == 0 ? 1 : 0, du, di, dbr, dsrc, dsh, dlg, stream_);
    }
    if (reg) {
        // sources sit on slack buses, which never regulate: the source results above do not depend on this step
        launch_regulator_result(B_, tile_width_, ds_, db_, d_reg_bus_.get(), n_reg_bus_, du, di, dlg, d_out_reg_.get(), stream_);
    }
    PGMB_CUDA(cudaGetLastError());
    auto back = [&](void* host, void const* dev, size_t bytes) {
        if (host != nullptr && bytes != 0) PGMB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, stream_));
    };
    back(out.u, du, sizeof(double) * n * topo_.n_bus * c2);
    back(out.bus_injection, di, sizeof(double) * n * topo_.n_bus * c2);
    back(out.branch, dbr, sizeof(double) * n * topo_.n_branch() * 4 * c2);
    back(out.source, dsrc, sizeof(double) * n * topo_.n_source() * 2 * c2);
    back(out.shunt, dsh, sizeof(double) * n * topo_.n_shunt() * 2 * c2);
    back(out.load_gen, dlg, sizeof(double) * n * topo_.n_load_gen() * 2 * c2);
    back(out.status, d_status_.get(), sizeof(int32_t) * n);
    back(out.n_iter, d_n_iter_.get(), sizeof(int32_t) * n);
    back(out.max_dev, d_max_dev_.get(), sizeof(double) * n);
    if (reg) back(out.voltage_regulator, d_out_reg_.get(), static_cast<size_t>(n) * ds_.n_regulator * 2);
    PGMB_CUDA(cudaStreamSynchronize(stream_));
}

int Engine::run(SolveOptions const& opt, PfInputView const& in, SolverOutputView const& out) {
    stage(in);
    solve_staged(opt);
    std::vector<int32_t> status_local;
    SolverOutputView o = out;
    if (o.status == nullptr) {
        status_local.resize(in.n_scenarios);
        o.status = status_local.data();
    }
    fetch(o);
    int failed = 0;
    for (int64_t s = 0; s != in.n_scenarios; ++s) failed += o.status[s] != 0 ? 1 : 0;
    return failed;
}

} // namespace pgmb
