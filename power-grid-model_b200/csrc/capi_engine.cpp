// C-ABI, engine level (include/pgm_b200.h).  Exceptions never cross the boundary: every entry is wrapped like the
// reference's call_with_catch (power_grid_model_c/src/handle.hpp:70-89) and reports through a thread-local message.
#include "capi_common.hpp"
#include "engine.hpp"

#include <map>

using namespace pgmb;

struct pgmb_engine {
    std::unique_ptr<Engine> engine;
    std::map<std::string, std::vector<int64_t>> index_cache;
};

namespace pgmb {
thread_local std::string g_last_error;

MathTopology topology_from_view(pgmb_math_topology const& t) {
    if (t.n_bus <= 0) throw InvalidArgument("n_bus must be positive");
    MathTopology m;
    m.n_bus = t.n_bus;
    m.phase_shift.assign(t.phase_shift, t.phase_shift + t.n_bus);
    m.branch_bus_idx.assign(t.branch_bus_idx, t.branch_bus_idx + 2 * t.n_branch);
    if (t.n_fill_in > 0) m.fill_in.assign(t.fill_in, t.fill_in + 2 * t.n_fill_in);
    m.sources_per_bus.assign(t.sources_per_bus, t.sources_per_bus + t.n_bus + 1);
    m.shunts_per_bus.assign(t.shunts_per_bus, t.shunts_per_bus + t.n_bus + 1);
    m.load_gens_per_bus.assign(t.load_gens_per_bus, t.load_gens_per_bus + t.n_bus + 1);
    if (m.n_load_gen() > 0) m.load_gen_type.assign(t.load_gen_type, t.load_gen_type + m.n_load_gen());
    if (t.voltage_regulators_per_load_gen != nullptr) {
        m.load_gen_regulator.assign(m.n_load_gen(), -1);
        for (Idx lg = 0; lg != m.n_load_gen(); ++lg) {
            int64_t const b = t.voltage_regulators_per_load_gen[lg], e = t.voltage_regulators_per_load_gen[lg + 1];
            if (e - b > 1) throw InvalidArgument("There are objects regulated by more than one regulator. Maximum one regulator is allowed.");
            if (e - b == 1) m.load_gen_regulator[lg] = b;
        }
    }
    return m;
}
} // namespace pgmb

namespace {
PfInputView view_of(pgmb_pf_input const& in) {
    return {in.n_scenarios, in.source_u_ref, in.source_is_shared != 0, in.s_injection, in.voltage_regulator, in.load_gen_status,
            in.method_hint_valid != 0 ? in.method_hint : -128};
}
SolverOutputView view_of(pgmb_solver_output const& o) {
    return {o.u, o.bus_injection, o.branch, o.source, o.shunt, o.load_gen, o.status, o.n_iter, o.max_dev, o.voltage_regulator};
}
SolveOptions options_of(pgmb_run_options const& o) {
    if (o.max_iter < 0 || o.max_iter > (int64_t{1} << 30)) throw InvalidArgument("max_iter out of range");
    return {o.method, o.err_tol, static_cast<int32_t>(o.max_iter)};
}

} // namespace

extern "C" {

const char* pgmb_last_error(void) { return g_last_error.c_str(); }
const char* pgmb_version(void) { return "pgm_b200 0.1 (reference power-grid-model 1.13 semantics)"; }
int pgmb_host_alloc(uint64_t bytes, void** ptr) {
    return guarded([&] {
        if (ptr == nullptr) throw InvalidArgument("null argument");
        *ptr = nullptr;
        if (bytes == 0) return;
        cudaError_t const e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
        if (e != cudaSuccess) throw CudaError(std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    });
}
int pgmb_host_free(void* ptr) {
    return guarded([&] {
        if (ptr == nullptr) return;
        cudaError_t const e = cudaFreeHost(ptr);
        if (e != cudaSuccess) throw CudaError(std::string("cudaFreeHost: ") + cudaGetErrorString(e));
    });
}
uint64_t pgmb_kernel_launch_count(void) { return pgmb::kernel_launch_count(); }
int pgmb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int pgmb_engine_create(const pgmb_math_topology* topo, int32_t symmetric, int32_t device, pgmb_engine** out) {
    return guarded([&] {
        if (topo == nullptr || out == nullptr) throw InvalidArgument("null argument");
        auto h = std::make_unique<pgmb_engine>();
        h->engine = std::make_unique<Engine>(topology_from_view(*topo), symmetric != 0, device);
        *out = h.release();
    });
}
void pgmb_engine_destroy(pgmb_engine* engine) { delete engine; }

int pgmb_engine_set_param(pgmb_engine* engine, const pgmb_math_param* param) {
    return guarded([&] {
        if (engine == nullptr || param == nullptr) throw InvalidArgument("null argument");
        engine->engine->set_param(param->branch_param, param->shunt_param, param->source_param);
    });
}

int pgmb_engine_get_index(pgmb_engine* engine, const char* name, const int64_t** data, int64_t* size) {
    return guarded([&] {
        if (engine == nullptr || name == nullptr) throw InvalidArgument("null argument");
        auto const& p = engine->engine->pattern();
        auto const& s = engine->engine->schedule();
        std::string const key{name};
        auto it = engine->index_cache.find(key);
        if (it == engine->index_cache.end()) {
            std::vector<int64_t> v;
            auto widen = [](std::vector<int32_t> const& x) { return std::vector<int64_t>(x.begin(), x.end()); };
            if (key == "row_indptr") v = p.row_indptr;
            else if (key == "col_indices") v = p.col_indices;
            else if (key == "bus_entry") v = p.bus_entry;
            else if (key == "row_indptr_lu") v = p.row_indptr_lu;
            else if (key == "col_indices_lu") v = p.col_indices_lu;
            else if (key == "diag_lu") v = p.diag_lu;
            else if (key == "map_lu_y_bus") v = p.map_lu_y_bus;
            else if (key == "lu_transpose_entry") v = p.lu_transpose_entry;
            else if (key == "y_bus_entry_indptr") v = p.y_bus_entry_indptr;
            else if (key == "level_ptr") v = widen(s.level_ptr);
            else if (key == "level_rows") v = widen(s.level_rows);
            else if (key == "row_program") v = widen(engine->engine->program().words);
            else if (key == "path_program") v = widen(engine->engine->path_program().words); // empty: not a radial grid
            else if (key == "wide_level_ptr") v = widen(engine->engine->wide_plan().level_ptr);
            else if (key == "wide_table") v = widen(engine->engine->wide_plan().table);
            else if (key == "wide_data") v = widen(engine->engine->wide_plan().data);
            else if (key == "upd_ptr") v = widen(s.upd_ptr);
            else if (key == "upd_a") v = widen(s.upd_a);
            else throw InvalidArgument("unknown index array: " + key);
            it = engine->index_cache.emplace(key, std::move(v)).first;
        }
        *data = it->second.data();
        *size = static_cast<int64_t>(it->second.size());
    });
}

int pgmb_engine_get_admittance(pgmb_engine* engine, const double** data, int64_t* size) {
    return guarded([&] {
        if (engine == nullptr) throw InvalidArgument("null argument");
        *data = engine->engine->admittance().data();
        *size = static_cast<int64_t>(engine->engine->admittance().size());
    });
}

int pgmb_engine_run(pgmb_engine* engine, const pgmb_run_options* opt, const pgmb_pf_input* input,
                    const pgmb_solver_output* output) {
    int failed = 0;
    int const rc = guarded([&] {
        if (engine == nullptr || opt == nullptr || input == nullptr || output == nullptr) throw InvalidArgument("null argument");
        failed = engine->engine->run(options_of(*opt), view_of(*input), view_of(*output));
    });
    if (rc != PGMB_OK) return rc;
    if (failed != 0) {
        g_last_error = std::to_string(failed) + " scenario(s) failed; see status[]";
        return PGMB_ERR_BATCH;
    }
    return PGMB_OK;
}

int pgmb_engine_stage(pgmb_engine* engine, const pgmb_pf_input* input) {
    return guarded([&] {
        if (engine == nullptr || input == nullptr) throw InvalidArgument("null argument");
        engine->engine->stage(view_of(*input));
    });
}
int pgmb_engine_solve_staged(pgmb_engine* engine, const pgmb_run_options* opt, float* solve_kernel_ms) {
    return guarded([&] {
        if (engine == nullptr || opt == nullptr) throw InvalidArgument("null argument");
        float const ms = engine->engine->solve_staged(options_of(*opt));
        if (solve_kernel_ms != nullptr) *solve_kernel_ms = ms;
    });
}
int pgmb_engine_fetch(pgmb_engine* engine, const pgmb_solver_output* output) {
    return guarded([&] {
        if (engine == nullptr || output == nullptr) throw InvalidArgument("null argument");
        engine->engine->fetch(view_of(*output));
    });
}

int pgmb_sparse_lu_create(int64_t n, const int64_t* row_indptr, const int64_t* col_indices, const int64_t* diag_lu, int32_t block_size,
                          int32_t is_complex, int32_t device, pgmb_sparse_lu** out) {
    return guarded([&] {
        if (row_indptr == nullptr || col_indices == nullptr || diag_lu == nullptr || out == nullptr || n < 0) throw InvalidArgument("null argument");
        *out = reinterpret_cast<pgmb_sparse_lu*>(new SparseLuBatch(n, row_indptr, col_indices, diag_lu, block_size, is_complex != 0, device));
    });
}
void pgmb_sparse_lu_destroy(pgmb_sparse_lu* solver) { delete reinterpret_cast<SparseLuBatch*>(solver); }
int pgmb_sparse_lu_solve(pgmb_sparse_lu* solver, int64_t n_batch, const double* data, const double* rhs, int32_t use_pivot_perturbation,
                         double* x, int32_t* status, int32_t* perturbed, int32_t* n_solves, double* lu_out, int8_t* perm_out) {
    std::vector<int32_t> st_local;
    int const rc = guarded([&] {
        if (solver == nullptr || data == nullptr || rhs == nullptr || x == nullptr) throw InvalidArgument("null argument");
        if (status == nullptr) {
            st_local.assign(static_cast<size_t>(std::max<int64_t>(n_batch, 0)), 0);
            status = st_local.data();
        }
        reinterpret_cast<SparseLuBatch*>(solver)->solve(n_batch, data, rhs, use_pivot_perturbation != 0, x, status, perturbed, n_solves,
                                                        lu_out, perm_out);
    });
    if (rc != PGMB_OK) return rc;
    for (int64_t b = 0; b < n_batch; ++b)
        if (status[b] != 0) {
            g_last_error = "Sparse matrix error, possibly singular matrix!\n";
            return PGMB_ERR_BATCH;
        }
    return PGMB_OK;
}

} // extern "C"
