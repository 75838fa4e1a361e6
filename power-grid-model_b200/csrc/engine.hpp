// Engine = one math sub-grid on one GPU: symbolic structure (shared by all scenarios), Y-bus values, and the batched
// solver kernels.  Host-side replacement of YBus<sym> + MathSolver<sym> (math_solver/y_bus.hpp:297-592,
// math_solver/math_solver.hpp:26-183) for whole batches of scenarios.
#pragma once

#include "kernels.cuh"
#include "symbolic.hpp"

#include <cuda_runtime.h>

#include <array>
#include <complex>
#include <memory>
#include <string>
#include <vector>

namespace pgmb {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct InvalidArgument : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define PGMB_CUDA(expr)                                                                                              \
    do {                                                                                                             \
        cudaError_t const err__ = (expr);                                                                            \
        if (err__ != cudaSuccess) {                                                                                  \
            throw ::pgmb::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(err__));                   \
        }                                                                                                            \
    } while (0)

// Device memory pool: cudaMalloc / cudaFree serialise the whole process (and cudaFree synchronises the device), which hurts
// when engines are rebuilt per scenario (topology-changing batches, one engine per host thread).  Released blocks are kept in
// power-of-two size classes and handed out again; the pool is trimmed when it holds too much or an allocation fails.
// A block must only be released when no work that uses it is in flight (~Engine synchronises its stream first).
class DevPool {
  public:
    static void* alloc(size_t bytes);
    static void release(void* p, size_t bytes);
    static void trim();
    static size_t size_class(size_t bytes) {
        size_t c = 256;
        while (c < bytes) c <<= 1;
        return c;
    }
};

// owning device buffer
template <class T> class DevBuf {
  public:
    DevBuf() = default;
    DevBuf(DevBuf const&) = delete;
    DevBuf& operator=(DevBuf const&) = delete;
    DevBuf(DevBuf&& o) noexcept : p_{o.p_}, n_{o.n_} { o.p_ = nullptr; o.n_ = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p_ = o.p_;
            n_ = o.n_;
            o.p_ = nullptr;
            o.n_ = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void ensure(size_t n) { // grow-only
        if (n <= n_) return;
        release();
        if (n == 0) return;
        p_ = static_cast<T*>(DevPool::alloc(n * sizeof(T)));
        n_ = n;
    }
    void upload(std::vector<T> const& h, cudaStream_t st) {
        ensure(h.size());
        if (!h.empty()) PGMB_CUDA(cudaMemcpyAsync(p_, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    T* get() const { return p_; }
    size_t size() const { return n_; }

  private:
    void release() {
        if (p_ != nullptr) DevPool::release(p_, n_ * sizeof(T));
        p_ = nullptr;
        n_ = 0;
    }
    T* p_{nullptr};
    size_t n_{0};
};

struct PfInputView {
    int64_t n_scenarios;
    double const* source_u_ref;
    bool source_is_shared;
    double const* s_injection;
    double const* voltage_regulator{nullptr}; // [n_regulator][4] status, u_ref, q_min, q_max (shared by the scenarios)
    int8_t const* load_gen_status{nullptr};   // [n_scenarios][n_load_gen], null = all on
    int32_t method_hint{-128};                // method the batch will be solved with (affects the tile width), -128 = unknown
};
struct SolverOutputView {
    double *u, *bus_injection, *branch, *source, *shunt, *load_gen;
    int32_t *status, *n_iter;
    double* max_dev;
    int8_t* voltage_regulator{nullptr}; // [n_scenarios][n_regulator][2] limit_violated, generator_status
};

class Engine {
  public:
    Engine(MathTopology topo, bool symmetric, int device);
    ~Engine();
    Engine(Engine const&) = delete;
    Engine& operator=(Engine const&) = delete;

    // branch_param [n_branch][4][B][B], shunt_param [n_shunt][B][B], source_param [n_source][2] (all complex)
    void set_param(double const* branch_param, double const* shunt_param, double const* source_param);

    // voltage regulator parameters of the next solves: [n_regulator][4] status, u_ref, q_min, q_max
    void set_regulators(double const* param);
    bool has_regulators() const { return !topo_.load_gen_regulator.empty(); }
    // calculation method the next staged batch will be solved with (PGMB_METHOD_*; 0 linear, 3 / 4 iterative current): the
    // tile width is chosen when the batch is staged, and the iterative-current kernels run two thread blocks per SM
    void set_method_hint(int method) { method_hint_ = method; }
    // device pipeline: Q allocation of the regulated generators of a solved chunk (rewrites their Q in the injection buffer) and
    // the per-regulator flags [n_scn][n_regulator][2] for the output kernel
    void launch_regulator_apply(DevBatch const& view, int8_t* out_reg, cudaStream_t st);
    void stage(PfInputView const& in);                      // H2D + layout conversion
    // device path of the model level: allocate the batch, upload per-scenario source references; load injections are then
    // produced on the device by the apply_load_update kernels (model_device.cpp)
    void stage_device(int64_t n_scn, double const* source_u_ref, bool source_is_shared);
    // Branch-outage overlay for the staged batch (call after stage / stage_device): scenario s replaces the parameters of math
    // branch `math_branch[s]` (-1 = none) by bparam[s] ([4][B][B] complex) -- the engine derives the replaced Y-bus entries, summing
    // each entry's elements in the assembly order (y_bus.hpp:400-431).  comp / energized: the branch component's index and
    // `energized` flag for the output kernels.  Only the Newton-Raphson block kernel and the result kernels read the overlay.
    // dead_off / dead: per scenario the index of its mask of buses without supply ([n_mask][n_bus] bytes), -1 = none.
    // n_slot: branch slots per scenario (N-k batches): math_branch / comp / energized are [n_scn][n_slot], bparam [n_scn][n_slot][4]...;
    // Y-bus entries shared by several switched branches of a scenario are replaced once, with all of them taken into account.
    void set_overlay(int64_t n_scn, int64_t const* math_branch, double const* bparam, int32_t const* comp, uint8_t const* energized,
                     int32_t const* dead_off = nullptr, uint8_t const* dead = nullptr, size_t dead_bytes = 0, int n_slot = 1);
    bool has_overlay() const { return db_.ovl.entry != nullptr; }
    void fetch_status(int32_t* status, int32_t* n_iter);
    float solve_staged(SolveOptions const& opt);            // kernels only; returns solver-kernel milliseconds
    // pipelined use (model device path): a view of the staged batch restricted to tiles [tile_begin, tile_end), and the solver
    // launch on a caller-chosen stream without synchronisation.  prepare_solve() resolves the method and builds what all
    // chunks share (the iterative-current factor) on the engine's own stream.
    DevBatch batch_view(int64_t tile_begin, int64_t tile_end) const;
    SolveOptions prepare_solve(SolveOptions const& opt);
    void launch_solve(DevBatch const& view, SolveOptions const& resolved, cudaStream_t st);
    void fetch(SolverOutputView const& out);                // result extraction + D2H
    int run(SolveOptions const& opt, PfInputView const& in, SolverOutputView const& out); // returns #failed

    MathTopology const& topology() const { return topo_; }
    LuPattern const& pattern() const { return pattern_; }
    EliminationSchedule const& schedule() const { return schedule_; }
    RowProgram const& program() const { return program_; }
    PathProgram const& path_program() const { return path_program_; }
    WideRowPlan const& wide_plan() const { return wide_plan_; }
    std::vector<double> const& admittance() const { return admittance_; } // [nnz][B][B] complex
    int device() const { return device_; }
    int phases() const { return B_; }
    int tile_width() const { return tile_width_; }
    cudaStream_t stream() const { return stream_; }
    DevStructure const& dev_structure() const { return ds_; }
    DevBatch const& dev_batch() const { return db_; }
    int last_method() const { return last_method_; }

  private:
    MathTopology topo_;
    bool symmetric_;
    int B_;
    int device_;
    LuPattern pattern_;
    EliminationSchedule schedule_;
    RowProgram program_;
    PathProgram path_program_;
    WideRowPlan wide_plan_;
    std::vector<double> admittance_;
    std::vector<double> branch_param_, shunt_param_, source_param_;
    bool param_set_{false};
    cudaStream_t stream_{};
    cudaEvent_t ev0_{}, ev1_{};

    // device structure
    DevBuf<int32_t> d_row_ptr_, d_col_idx_, d_diag_, d_map_y_, d_level_ptr_, d_level_rows_, d_upd_ptr_, d_upd_u_, d_upd_a_,
        d_lg_ptr_, d_src_ptr_, d_prog_, d_path_prog_, d_wide_level_ptr_, d_wide_table_, d_wide_data_, d_y_row_ptr_, d_y_col_idx_, d_branch_bus_, d_shunt_bus_, d_lg_bus_, d_src_bus_;
    DevBuf<int8_t> d_lg_type_;
    DevBuf<double> d_ydata_, d_src_yref_, d_src_y1y0_, d_branch_param_, d_shunt_param_, d_phase_shift_;
    DevStructure ds_{};

    // batch buffers
    int method_hint_{1};
    int tile_width_{8};
    int n_slot_{64};
    DevBuf<double> d_side_, d_wide_terms_, d_wide_rhs_, d_wide_sum_;
    DevBuf<uint8_t> d_row_is_wide_;
    DevBuf<double> d_jac_, d_xvec_, d_pol_, d_u_, d_sinj_, d_usrc_, d_max_dev_, d_in_sinj_, d_in_usrc_;
    DevBuf<double> d_out_u_, d_out_inj_, d_out_branch_, d_out_source_, d_out_shunt_, d_out_lg_;
    DevBuf<uint8_t> d_perm_, d_lg_status_, d_qviol_, d_in_lg_status_;
    DevBuf<int32_t> d_lg_reg_, d_reg_bus_;
    DevBuf<double> d_reg_param_;
    DevBuf<int8_t> d_out_reg_;
    DevBuf<int32_t> d_ovl_entry_, d_ovl_branch_, d_ovl_comp_;
    DevBuf<double> d_ovl_y_, d_ovl_bparam_;
    DevBuf<uint8_t> d_ovl_energized_, d_ovl_dead_;
    DevBuf<int32_t> d_ovl_dead_off_;
    std::vector<std::array<int32_t, 4>> branch_entries_; // Y-bus entries (ff, ft, tf, tt) of each math branch, -1 = none
    std::vector<double> reg_param_;
    int n_reg_bus_{0};
    bool reg_param_set_{false};
    DevBuf<int32_t> d_status_, d_n_iter_;
    DevBuf<unsigned long long> d_phase_;
    DevBuf<double> d_ic_factor_; // shared iterative-current factor [nnz_lu][2]
    DevBuf<int32_t> d_ic_flag_;
    DevBuf<uint8_t> d_ic_perm_;  // block permutations of the shared factor (asymmetric)
    bool ic_factor_valid_{false};
    DevBatch db_{};
    int last_method_{1};

    void upload_structure();
    void choose_tiling(int64_t n_scn);
    void allocate_batch(int64_t n_scn);
};

// Batched sparse LU with the reference's pivot perturbation + iterative refinement (sparse_lu.cu)
class SparseLuBatch {
  public:
    SparseLuBatch(int64_t n, int64_t const* indptr, int64_t const* indices, int64_t const* diag, int block_size, bool is_complex,
                  int device);
    // data [n_batch][nnz][N*N] column-major blocks, rhs / x [n_batch][n][N] (complex: re, im interleaved); optional outputs may be
    // null: perturbed [n_batch] (a pivot was perturbed), n_solves [n_batch] (solve_once calls incl. refinement), lu_out (factors),
    // perm_out [n_batch][n][2][N] (p then q of every pivot block)
    void solve(int64_t n_batch, double const* data, double const* rhs, bool use_pivot_perturbation, double* x, int32_t* status,
               int32_t* perturbed, int32_t* n_solves, double* lu_out, int8_t* perm_out);
    int64_t size() const { return n_; }
    int64_t nnz() const { return nnz_; }

  private:
    int64_t n_, nnz_;
    int block_;
    bool complex_;
    int device_;
    DevBuf<int64_t> d_indptr_, d_indices_, d_diag_;
};

uint64_t kernel_launch_count(); // kernels launched by this library since it was loaded

// kernel launchers (nr_sym.cu, result_sym.cu)
void launch_nr_sym(int tile_width, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot,
                   cudaStream_t st);
void launch_nr_sym_v2(int tile_width, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot,
                      cudaStream_t st);
void launch_nr_sym_v3(int tile_width, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot,
                      cudaStream_t st);
void launch_apply_load_update_sym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m,
                                  DevUpdateBuffers const& ub, cudaStream_t st);
void launch_source_result_sym(int tw, DevStructure const& s, DevBatch const& b, int force_const_y, double* out, cudaStream_t st);
void launch_pack_node_sym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int force_const_y,
                          double const* src_res, void* out, cudaStream_t st);
void launch_pack_branch_sym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int first, int count,
                            void* out, cudaStream_t st);
void launch_pack_appliance_sym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int force_const_y,
                               int first, int count, double const* src_res, void* out, cudaStream_t st);
void launch_apply_load_update_asym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m,
                                   DevUpdateBuffers const& ub, cudaStream_t st);
void launch_source_result_asym(int tw, DevStructure const& s, DevBatch const& b, int force_const_y, double* out, cudaStream_t st);
void launch_pack_node_asym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int force_const_y,
                           double const* src_res, void* out, cudaStream_t st);
void launch_pack_branch_asym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int first, int count,
                             void* out, cudaStream_t st);
void launch_pack_appliance_asym(int tw, DevStructure const& s, DevBatch const& b, DevModelTables const& m, int force_const_y,
                                int first, int count, double const* src_res, void* out, cudaStream_t st);
void launch_nr_block6(int tile_width, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int threads, cudaStream_t st);
void launch_regulator_result(int phases, int tile_width, DevStructure const& s, DevBatch const& b, int32_t const* reg_bus,
                             int n_reg_bus, double const* out_u, double const* out_inj, double* out_lg, int8_t* out_reg,
                             cudaStream_t st);
void launch_regulator_apply(int phases, int tile_width, DevStructure const& s, DevBatch const& b, int32_t const* reg_bus,
                            int n_reg_bus, int8_t* out_reg, cudaStream_t st);
void launch_pack_regulator(int64_t n_scn, int n_comp, int n_regulator, int32_t const* id, int32_t const* math, uint8_t const* status,
                           int8_t const* reg_out, void* out, cudaStream_t st);
void launch_status_to_tile(int tile_width, uint8_t const* src, uint8_t* dst, int64_t n_scn, int n_item, cudaStream_t st);
void launch_nr_block(int phases, int tile_width, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, int n_slot,
                     cudaStream_t st);
void launch_math_result_asym(int tile_width, DevStructure const& s, DevBatch const& b, int force_const_y, double* out_u,
                             double* out_inj, double* out_branch, double* out_source, double* out_shunt, double* out_lg,
                             cudaStream_t st);
void launch_linear_asym(int tw, DevStructure const& s, DevBatch const& b, int n_slot, cudaStream_t st);
void launch_ic_factor_asym(DevStructure const& s, double* factor, uint8_t* perm, int* flag, cudaStream_t st);
void launch_ic_iterate_asym(int tw, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, double* factor,
                            uint8_t* perm, int const* flag, int n_slot, cudaStream_t st);
void launch_linear_sym(int tw, DevStructure const& s, DevBatch const& b, int n_slot, cudaStream_t st);
void launch_ic_factor(DevStructure const& s, double* factor, int* flag, cudaStream_t st);
void launch_ic_iterate_sym(int tw, DevStructure const& s, DevBatch const& b, SolveOptions const& opt, double const* factor,
                           int const* flag, int n_slot, cudaStream_t st);
void launch_to_tile(int tile_width, double const* src, double* dst, int64_t n_scn, int n_item, int n_comp, int shared_src,
                    cudaStream_t st);
void launch_math_result_sym(int tile_width, DevStructure const& s, DevBatch const& b, int force_const_y, double* out_u,
                            double* out_inj, double* out_branch, double* out_source, double* out_shunt, double* out_lg,
                            cudaStream_t st);

} // namespace pgmb
