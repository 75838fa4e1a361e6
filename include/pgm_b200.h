/* pgm_b200 -- C-ABI of the B200-native batch power-flow engine.
 *
 * Two seams, both plain C (pointers + sizes, no C++/torch types):
 *
 *  1. ENGINE level  = the reference's math-solver seam.  One engine == one math sub-grid == what the reference builds
 *     from a `MathModelTopology` (calculation_parameters.hpp:160-213) as `YBus<sym>` + `MathSolver<sym>`
 *     (math_solver/math_solver_dispatch.hpp:26-107, math_solver/math_solver.hpp:43-64).  It replaces
 *     `MathSolverBase<sym>::run_power_flow(input, err_tol, max_iter, cache_run, log, method, y_bus)` called once per
 *     scenario by `MainModelImpl::calculate_` (main_model_impl.hpp:310-312) with ONE call over all scenarios.
 *
 *  2. MODEL level   = the reference's `PGM_create_model` / `PGM_calculate` seam
 *     (power_grid_model_c/include/power_grid_model_c/model.h:43-48, 116-118) for the PF component subset, with the
 *     dataset handles flattened into structs of caller-owned buffers (same struct layouts as the reference's
 *     `PGM_def_input_* / update_* / sym_output_* / asym_output_*` components, see pgm_b200.structs).  It replaces the
 *     CPU thread-pool dispatcher `JobDispatch::batch_calculation` (job_dispatch.hpp:37-68).
 *
 * Complex numbers are interleaved (re, im) doubles.  B = 1 (symmetric) or 3 (asymmetric); tensors are row-major
 * [r][c].  All index types are int64 like the reference's `Idx`; ids are int32 (`ID`), enums int8 (`IntS`).
 * Every function returns 0 on success or a PGMB_ERR_* code; the message is available from pgmb_last_error()
 * (thread-local, like one PGM_Handle per thread).  Nothing here ever falls back to a CPU solver: if no CUDA device
 * is usable, create/run calls fail with PGMB_ERR_CUDA.
 */
#ifndef PGM_B200_H
#define PGM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGMB_API __attribute__((visibility("default")))

enum {
    PGMB_OK = 0,
    PGMB_ERR_INVALID = 1, /* invalid argument / unsupported option */
    PGMB_ERR_CUDA = 2,    /* CUDA runtime error or no device */
    PGMB_ERR_BATCH = 3,   /* at least one scenario failed: see status[] (PGM_batch_error analogue) */
    PGMB_ERR_INTERNAL = 4
};

/* per-scenario status (the reference throws IterationDiverge / SparseMatrixError per scenario,
 * common/exception.hpp:89-110; job_dispatch.hpp:181-224 collects them) */
enum {
    PGMB_SCN_OK = 0,
    PGMB_SCN_DIVERGED = 1,
    PGMB_SCN_SINGULAR = 2,
    PGMB_SCN_ERROR = 3,
    PGMB_SCN_UNALLOCATED_Q = 4 /* "Unallocated Q remains after distribution" (common_solver_functions.hpp:373-376) */
};

/* CalculationMethod, common/enum.hpp:33-41 */
enum {
    PGMB_METHOD_DEFAULT = -128,
    PGMB_METHOD_LINEAR = 0,
    PGMB_METHOD_NEWTON_RAPHSON = 1,
    PGMB_METHOD_ITERATIVE_CURRENT = 3,
    PGMB_METHOD_LINEAR_CURRENT = 4
};

PGMB_API const char* pgmb_last_error(void);
PGMB_API int pgmb_device_count(void);
PGMB_API const char* pgmb_version(void);
/* page-locked host memory for update / output buffers (cudaHostAlloc): with such buffers the model-level batch call overlaps
 * its transfers with the solver; pageable buffers work too but serialise.  No reference counterpart: the reference's
 * buffers are plain host memory owned by the caller (dataset.h:173-206) and stay caller-owned here. */
PGMB_API int pgmb_host_alloc(uint64_t bytes, void** ptr);
PGMB_API int pgmb_host_free(void* ptr);
/* number of CUDA kernels this library has launched since it was loaded (evidence that the GPU path ran) */
PGMB_API uint64_t pgmb_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * ENGINE level
 * ---------------------------------------------------------------------------------------------------------- */

/* MathModelTopology as arrays (calculation_parameters.hpp:160-213); grouped index vectors in sparse (indptr) form */
typedef struct pgmb_math_topology {
    int64_t n_bus;
    const double* phase_shift;        /* [n_bus] */
    int64_t n_branch;
    const int64_t* branch_bus_idx;    /* [n_branch][2], -1 = disconnected side */
    int64_t n_fill_in;
    const int64_t* fill_in;           /* [n_fill_in][2] */
    const int64_t* sources_per_bus;   /* indptr [n_bus + 1] */
    const int64_t* shunts_per_bus;    /* indptr [n_bus + 1] */
    const int64_t* load_gens_per_bus; /* indptr [n_bus + 1] */
    const int8_t* load_gen_type;      /* [n_load_gen] LoadGenType: 0 const_pq, 1 const_y, 2 const_i */
    /* voltage regulators grouped by the load_gen they regulate (calculation_parameters.hpp:176, topology.hpp:594-600):
     * indptr [n_load_gen + 1], or NULL when the grid has none.  At most one regulator per load_gen (main_core/input.hpp:216-241
     * enforces that for every model). */
    const int64_t* voltage_regulators_per_load_gen;
} pgmb_math_topology;

/* MathModelParam<sym> (calculation_parameters.hpp:240-255) */
typedef struct pgmb_math_param {
    const double* branch_param; /* [n_branch][4 (ff,ft,tf,tt)][B][B] complex */
    const double* shunt_param;  /* [n_shunt][B][B] complex */
    const double* source_param; /* [n_source][2 (y1, y0)] complex  (SourceCalcParam) */
} pgmb_math_param;

typedef struct pgmb_run_options {
    int32_t method;   /* PGMB_METHOD_* */
    double err_tol;   /* PGM_Options.err_tol, default 1e-8 */
    int64_t max_iter; /* PGM_Options.max_iter, default 20 */
} pgmb_run_options;

/* PowerFlowInput<sym> for n_scenarios scenarios (calculation_parameters.hpp:270-277) */
typedef struct pgmb_pf_input {
    int64_t n_scenarios;
    const double* source_u_ref; /* [n_scenarios][n_source] complex; or [n_source] when source_is_shared != 0 */
    int32_t source_is_shared;
    const double* s_injection;  /* [n_scenarios][n_load_gen][B] complex, per-unit, injection direction */
    /* grids with voltage regulators only (NULL otherwise): VoltageRegulatorCalcParam (calculation_parameters.hpp:228-236)
     * as [n_regulator][4] = status, u_ref, q_min, q_max (per unit, NaN = no limit), shared by the scenarios of the call, and
     * the status of each load_gen per scenario (PowerFlowInput::load_gen_status), NULL = all on */
    const double* voltage_regulator;
    const int8_t* load_gen_status; /* [n_scenarios][n_load_gen] */
    /* optional: the method the staged batch will be solved with (PGMB_METHOD_*).  The tile width of the device layout is chosen
     * at staging time and iterative-current batches prefer another width than Newton-Raphson ones.  method_hint is only read
     * when method_hint_valid != 0 (zero-initialised structs keep the default). */
    int32_t method_hint;
    int32_t method_hint_valid;
} pgmb_pf_input;

/* SolverOutput<sym> for n_scenarios scenarios (calculation_parameters.hpp:338-350); any pointer may be NULL */
typedef struct pgmb_solver_output {
    double* u;             /* [n_scenarios][n_bus][B] complex */
    double* bus_injection; /* [n_scenarios][n_bus][B] complex */
    double* branch;        /* [n_scenarios][n_branch][4 (s_f, s_t, i_f, i_t)][B] complex */
    double* source;        /* [n_scenarios][n_source][2 (s, i)][B] complex */
    double* shunt;         /* [n_scenarios][n_shunt][2 (s, i)][B] complex */
    double* load_gen;      /* [n_scenarios][n_load_gen][2 (s, i)][B] complex */
    int32_t* status;       /* [n_scenarios] PGMB_SCN_* */
    int32_t* n_iter;       /* [n_scenarios] iterations used (the reference only logs it, iterative_pf_solver.hpp:87) */
    double* max_dev;       /* [n_scenarios] last max |dU| */
    int8_t* voltage_regulator; /* [n_scenarios][n_regulator][2] = limit_violated (0 none, 1 lower, 2 upper), generator_status
                                * (VoltageRegulatorSolverOutput, calculation_parameters.hpp:94-100) */
} pgmb_solver_output;

typedef struct pgmb_engine pgmb_engine;

/* Symbolic stage, once per topology: Y-bus CSR + LU pattern with fill-ins (YBusStructure, y_bus.hpp:122-293) and
 * the elimination schedule (flattened index walk of SparseLUSolver::prefactorize, sparse_lu_solver.hpp:346-495),
 * uploaded to `device`. */
PGMB_API int pgmb_engine_create(const pgmb_math_topology* topo, int32_t symmetric, int32_t device, pgmb_engine** out);
PGMB_API void pgmb_engine_destroy(pgmb_engine* engine);

/* Y-bus assembly (YBus::update_admittance, y_bus.hpp:342-431) + upload. */
PGMB_API int pgmb_engine_set_param(pgmb_engine* engine, const pgmb_math_param* param);

/* Structure introspection for parity tests: name in {row_indptr, col_indices, bus_entry, row_indptr_lu,
 * col_indices_lu, diag_lu, map_lu_y_bus, lu_transpose_entry, y_bus_entry_indptr, level_ptr, level_rows};
 * admittance via pgmb_engine_get_admittance ([nnz][B][B] complex). Pointers stay valid until destroy/set_param. */
PGMB_API int pgmb_engine_get_index(pgmb_engine* engine, const char* name, const int64_t** data, int64_t* size);
PGMB_API int pgmb_engine_get_admittance(pgmb_engine* engine, const double** data, int64_t* size);

/* Batch power flow, host buffers in / host buffers out (copies inside). Returns PGMB_ERR_BATCH when some scenario
 * failed; the other scenarios' results are valid (BatchCalculationError semantics, job_dispatch.hpp:208-224). */
PGMB_API int pgmb_engine_run(pgmb_engine* engine, const pgmb_run_options* opt, const pgmb_pf_input* input,
                             const pgmb_solver_output* output);

/* Device-resident variant used for kernel timing: stage inputs once, run the solver kernels on them (no host
 * transfers), read kernel time measured with CUDA events on the engine's stream. */
PGMB_API int pgmb_engine_stage(pgmb_engine* engine, const pgmb_pf_input* input);
PGMB_API int pgmb_engine_solve_staged(pgmb_engine* engine, const pgmb_run_options* opt, float* solve_kernel_ms);
PGMB_API int pgmb_engine_fetch(pgmb_engine* engine, const pgmb_solver_output* output);

/* ------------------------------------------------------------------------------------------------------------
 * SparseLUSolver with pivot perturbation + iterative refinement (sparse_lu_solver.hpp:277-828; the perturbation path :514-649),
 * batched: every system of the batch shares one block-CSR pattern (row_indptr / col_indices / diag_lu as the reference's
 * constructor takes them, fill-ins included) and has its own values and right-hand side.
 * block_size: 1 (scalar), 2, 3, 6 for real values; 1, 3 for complex values (is_complex != 0; re, im interleaved).
 * Blocks are column-major like the reference's Eigen blocks.  use_pivot_perturbation as in SparseLUSolver::prefactorize(data,
 * perm, use_pivot_perturbation): threshold 1e-13 * (block-off-diagonal infinity norm), the perturbed pivot keeps its phase, and a
 * perturbed factorisation is solved with iterative refinement (backward error <= 1e-13, at most 6 solves, else singular).
 * The power-flow solvers never enable it (as in the reference); state estimation would.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct pgmb_sparse_lu pgmb_sparse_lu;
PGMB_API int pgmb_sparse_lu_create(int64_t n, const int64_t* row_indptr, const int64_t* col_indices, const int64_t* diag_lu,
                                   int32_t block_size, int32_t is_complex, int32_t device, pgmb_sparse_lu** out);
PGMB_API void pgmb_sparse_lu_destroy(pgmb_sparse_lu* solver);
/* data [n_batch][nnz][block_size^2], rhs and x [n_batch][n][block_size]; status [n_batch]: PGMB_SCN_OK or PGMB_SCN_SINGULAR
 * (SparseMatrixError).  Optional (NULL = not wanted): perturbed [n_batch] 1 when a pivot was perturbed, n_solves [n_batch] number
 * of triangular solves (1 without refinement), lu_out (the factors, same layout as data), perm_out [n_batch][n][2][block_size]
 * (block permutations p then q). Returns PGMB_ERR_BATCH when some system is singular. */
PGMB_API int pgmb_sparse_lu_solve(pgmb_sparse_lu* solver, int64_t n_batch, const double* data, const double* rhs,
                                  int32_t use_pivot_perturbation, double* x, int32_t* status, int32_t* perturbed,
                                  int32_t* n_solves, double* lu_out, int8_t* perm_out);

/* ------------------------------------------------------------------------------------------------------------
 * MODEL level (component structs; layouts == the reference's dataset structs, see pgm_b200/structs.py)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct pgmb_component_buffer {
    int64_t n;        /* elements (input) or elements per scenario (uniform batch); -1 => sparse batch, use indptr */
    const int64_t* indptr; /* [n_scenarios + 1] for sparse batches, else NULL */
    const void* data; /* packed row buffer of the component's struct */
} pgmb_component_buffer;

/* order = component storage order of the reference (all_components.hpp:36-39), PF subset */
typedef struct pgmb_input_data {
    pgmb_component_buffer node, line, transformer, shunt, source, sym_gen, asym_gen, sym_load, asym_load;
    pgmb_component_buffer voltage_regulator; /* VoltageRegulatorInput (auxiliary/input.hpp:492-498) */
    /* further branch components; in the branch sequence of the model they sit between `line` and `transformer` like in the
     * reference's component list (all_components.hpp:36-39): line, asym_line, generic_branch, transformer */
    pgmb_component_buffer asym_line;      /* AsymLineInput (auxiliary/input.hpp:99-142) */
    pgmb_component_buffer generic_branch; /* GenericBranchInput (auxiliary/input.hpp:144-165); symmetric calculations only */
    pgmb_component_buffer link;           /* LinkInput = BranchInput; branch sequence: line, asym_line, link, generic_branch, transformer */
    pgmb_component_buffer three_winding_transformer; /* ThreeWindingTransformerInput (Branch3Input + transformer data, 304 bytes) */
    pgmb_component_buffer transformer_tap_regulator; /* TransformerTapRegulatorInput (auxiliary/input.hpp, 48 bytes): automatic tap changer */
} pgmb_input_data;

typedef struct pgmb_update_data {
    int64_t n_scenarios;
    pgmb_component_buffer line, transformer, shunt, source, sym_gen, asym_gen, sym_load, asym_load;
    pgmb_component_buffer voltage_regulator; /* VoltageRegulatorUpdate (auxiliary/update.hpp:213-219) */
    pgmb_component_buffer asym_line, generic_branch; /* BranchUpdate */
    pgmb_component_buffer link;                      /* BranchUpdate */
    pgmb_component_buffer three_winding_transformer; /* ThreeWindingTransformerUpdate: id, status_1, status_2, status_3, tap_pos */
    pgmb_component_buffer transformer_tap_regulator; /* TransformerTapRegulatorUpdate: id, status, u_set, u_band, line_drop_compensation_r/x */
} pgmb_update_data;

/* caller-owned output buffers [n_scenarios][n_component]; NULL = component not requested
 * (only components present in the output dataset are produced, main_model_impl.hpp:450-460) */
typedef struct pgmb_output_data {
    void *node, *line, *transformer, *shunt, *source, *sym_gen, *asym_gen, *sym_load, *asym_load;
    void* voltage_regulator; /* VoltageRegulatorOutput (auxiliary/output.hpp:239-243) */
    void *asym_line, *generic_branch; /* BranchOutput */
    void* link;                       /* BranchOutput (loading 0) */
    void* three_winding_transformer;  /* Branch3Output (auxiliary/output.hpp): loading_1..3, loading, p/q/i/s per side */
    void* transformer_tap_regulator;  /* TransformerTapRegulatorOutput: id, energized, tap_pos (na unless a tap strategy ran) */
} pgmb_output_data;

/* pgmb_options.flags -- measurement of the device-resident pipeline (bench.py `value`): one load-profile batch on one device.
 * RESIDENT_INPUT:  the update rows of this batch were uploaded by the previous calculate call on this model (same buffers and
 *                  sizes) and are still in HBM: no host-to-device copy.
 * RESIDENT_OUTPUT: the output structs are produced in HBM and stay there (no device-to-host copy; the caller's output buffers
 *                  only select the components).  A following call without this flag delivers them. */
#define PGMB_FLAG_RESIDENT_INPUT 1u
#define PGMB_FLAG_RESIDENT_OUTPUT 2u

/* PGM_Options (power_grid_model_c/src/options.hpp:16-27), PF subset */
typedef struct pgmb_options {
    int32_t calculation_method; /* PGMB_METHOD_* */
    int32_t symmetric;          /* 1 symmetric, 0 asymmetric */
    double err_tol;
    int64_t max_iter;
    int32_t n_devices;          /* GPUs ONE batch is spread over inside the call, starting at first_device: contiguous scenario
                                 * blocks per device, one host thread + stream set per device, symbolic structures replicated,
                                 * results into disjoint slices of the caller's buffers, no collective (the reference fans a batch
                                 * out over host threads, job_dispatch.hpp:131-172).  0 / 1 = one device unless the environment
                                 * variable PGMB_DEVICES names more (that is how a PGM_calculate client selects it). */
    int32_t first_device;       /* CUDA device ordinal of the first device */
    int32_t threading;          /* host threads for batches whose scenarios change topology / parameters (each thread owns a
                                 * model copy, job_dispatch.hpp:88-160): -1 or 0 = all cores, n > 0 = n threads, 1 = sequential */
    uint32_t flags;             /* PGMB_FLAG_* */
    int32_t tap_changing_strategy; /* PGM_TapChangingStrategy (basics.h:225-235): 0 disabled, 1 any_valid_tap, 2 min_voltage_tap,
                                    * 3 max_voltage_tap, 4 fast_any_tap.  Not 0: every scenario runs the automatic tap changer
                                    * (optimizer/tap_position_optimizer.hpp) around its power flows, scenario by scenario */
} pgmb_options;

typedef struct pgmb_model pgmb_model;

/* PGM_create_model (model.h:43-48) */
PGMB_API int pgmb_model_create(double system_frequency, const pgmb_input_data* input, pgmb_model** out);
PGMB_API void pgmb_model_destroy(pgmb_model* model);
/* PGM_update_model (model.h:54): permanent update with scenario 0 of `update` */
PGMB_API int pgmb_model_update(pgmb_model* model, const pgmb_update_data* update);
/* PGM_calculate (model.h:116-118): update == NULL => single calculation, else batch.
 * n_iter / status: optional [n_scenarios] arrays. */
PGMB_API int pgmb_model_calculate(pgmb_model* model, const pgmb_options* opt, const pgmb_update_data* update,
                                  const pgmb_output_data* output, int32_t* n_iter, int32_t* status);
/* Math model export for parity tests (same names as pgmb_engine_get_index plus: slack_bus, phase_shift (f64),
 * branch_bus_idx, fill_in, sources_per_bus, shunts_per_bus, load_gens_per_bus, load_gen_type, node_coupling ...) */
PGMB_API int pgmb_model_get_index(pgmb_model* model, int64_t math_group, const char* name, const int64_t** data,
                                  int64_t* size);
PGMB_API int pgmb_model_get_real(pgmb_model* model, int64_t math_group, int32_t symmetric, const char* name,
                                 const double** data, int64_t* size);
PGMB_API int64_t pgmb_model_n_math_groups(pgmb_model* model);
/* PowerFlowInput of every scenario of `update` for one math group (prepare_power_flow_input after applying the scenario's
 * update, main_core/calculation_input_preparation.hpp:163-188): s_injection [n_scenarios][n_load_gen][B] complex,
 * source_u_ref [n_scenarios][n_source] complex; caller-allocated. Only valid for batches that change loads / sources. */
PGMB_API int pgmb_model_batch_pf_input(pgmb_model* model, const pgmb_update_data* update, int32_t symmetric,
                                       int64_t math_group, double* s_injection, double* source_u_ref);
/* Host planning of a branch-switching batch, for tests (no device needed): plan [n_scenarios][4] =
 * route (0: shared symbolic pattern through the branch overlay, 1: own topology), overlay slots in use, buses that lose their
 * supply, 1 when the batch is planned on the union grid (scenarios close branches that are open in the base state).
 * Reference behaviour it stands in for: per-scenario rebuild_topology, main_model_impl.hpp:139-160. */
PGMB_API int pgmb_model_outage_plan(pgmb_model* model, const pgmb_update_data* update, int32_t symmetric, int64_t* plan);
/* timing of the last calculate call, milliseconds: [0] host prepare (tables, source references), [1] host time to
 * enqueue the chunk pipeline (H2D, kernels, D2H of every chunk), [2] solver kernels (CUDA events, summed over the chunks,
 * which overlap), [3] host output conversion (per-scenario route only), [4] wait for the pipeline to drain + status read-back, [5] total wall */
PGMB_API int pgmb_model_last_timing(pgmb_model* model, double* ms6);
/* device time of the last calculate call's pipeline in milliseconds: CUDA events on the engine stream around
 * (H2D ->) apply update -> solve -> result extraction / output structs (-> D2H) of all chunks; with several devices the
 * maximum over the devices; 0 when the batch did not take the device pipeline */
PGMB_API int pgmb_model_device_pipeline_ms(pgmb_model* model, double* ms);

/* ------------------------------------------------------------------------------------------------------------
 * Benchmark input: the reference's fictional grid generator (tests/benchmark_cpp/fictional_grid_generator.hpp)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct pgmb_grid_option {
    int64_t n_node_total_specified, n_mv_feeder, n_node_per_mv_feeder, n_lv_feeder, n_connection_per_lv_feeder;
    int32_t has_mv_ring, has_lv_ring;
} pgmb_grid_option;
typedef struct pgmb_fictional_grid pgmb_fictional_grid;
PGMB_API int pgmb_fictional_grid_create(const pgmb_grid_option* option, uint32_t seed, pgmb_fictional_grid** out);
PGMB_API void pgmb_fictional_grid_destroy(pgmb_fictional_grid* grid);
/* component in {node,line,transformer,shunt,source,sym_load,asym_load}; data points at packed input structs */
PGMB_API int pgmb_fictional_grid_get(pgmb_fictional_grid* grid, const char* component, const void** data, int64_t* n);
/* load-profile batch (generate_batch_input): fills caller buffers [batch_size][n_sym_load] / [batch_size][n_asym_load] */
PGMB_API int pgmb_fictional_grid_batch(pgmb_fictional_grid* grid, int64_t batch_size, uint32_t seed, void* sym_load_update,
                                       void* asym_load_update);

#ifdef __cplusplus
}
#endif
#endif /* PGM_B200_H */
