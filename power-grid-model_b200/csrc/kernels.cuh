// Device-side data layout shared by the kernels and the engine (host).
//
// Scenario-interleaved structure of arrays: scenarios are grouped in tiles of T consecutive scenarios; every per-scenario
// quantity q[item][component] of a tile is stored as q[item][component][T] (scenario fastest), so the T lanes that work
// on the same matrix entry for T scenarios touch T consecutive doubles (T*8 bytes, 2 sectors for T = 8, 8 for T = 32).
// One thread block owns one tile for the whole Newton-Raphson loop: no grid-wide synchronisation is ever needed.
#pragma once

#include <cstdint>

namespace pgmb {

struct DevStructure {
    int32_t n_bus, nnz, nnz_lu, n_level, n_load_gen, n_source, n_branch, n_shunt;
    // LU pattern (with fill-ins), CSR
    int32_t const* row_ptr;   // [n_bus + 1]
    int32_t const* col_idx;   // [nnz_lu]
    int32_t const* diag;      // [n_bus]
    int32_t const* map_y;     // [nnz_lu] -> Y-bus entry, -1 for fill-ins
    // elimination schedule
    int32_t const* level_ptr;  // [n_level + 1]
    int32_t const* level_rows; // [n_bus]
    int32_t const* upd_ptr;    // [nnz_lu + 1]
    int32_t const* upd_u;      // U entry (c, j)
    int32_t const* upd_a;      // target entry (k, j)
    // appliances per bus
    int32_t const* lg_ptr;    // [n_bus + 1]
    int8_t const* lg_type;    // [n_load_gen]
    int32_t const* src_ptr;   // [n_bus + 1]
    // Y-bus CSR (without fill-ins) for result extraction
    int32_t const* y_row_ptr; // [n_bus + 1]
    int32_t const* y_col_idx; // [nnz]
    // branches / shunts for result extraction
    int32_t const* branch_bus; // [n_branch][2]
    int32_t const* shunt_bus;  // [n_shunt]
    int32_t const* lg_bus;     // [n_load_gen]
    int32_t const* src_bus;    // [n_source]
    // values shared by all scenarios (complex, B x B row-major)
    double const* ydata;        // [nnz][B*B][2]
    double const* src_yref;     // [n_source][B*B][2]   y_ref tensor of each source
    double const* src_y1y0;     // [n_source][2][2]     y1, y0
    double const* branch_param; // [n_branch][4][B*B][2]
    double const* shunt_param;  // [n_shunt][B*B][2]
    double const* phase_shift;  // [n_bus]
    // row programs (symbolic.hpp RowProgram): level_ptr | task_off | records
    int32_t const* prog;
    int32_t prog_words;
    // wide rows (symbolic.hpp WideRowPlan): rows eliminated by the whole thread block together
    int32_t const* wide_level_ptr; // [n_level + 1]
    int32_t const* wide_table;     // 8 words per wide row
    int32_t const* wide_data;
    uint8_t const* row_is_wide;    // [n_bus]
    int32_t n_wide, wide_max_upd, wide_max_lower, wide_max_entries;
    // path programs (symbolic.hpp PathProgram) for radial grids; null when the grid has a cyclic core
    int32_t const* path_prog;
    int32_t path_prog_words;
    int32_t path_prog_smem_words; // prefix of the path program that the kernel stages in shared memory
    // voltage regulators (PV buses, newton_raphson_pf_solver.hpp:374-453); lg_reg == nullptr when the grid has none
    int32_t n_regulator;
    int32_t const* lg_reg;    // [n_load_gen] regulator of the load_gen or -1 (at most one regulator per load_gen)
    double const* reg_param;  // [n_regulator][4] status, u_ref, q_min, q_max (per unit; NaN = no limit)
};

// Branch-outage overlay of a batch whose scenarios each switch a few branches of the shared topology (N-1 / N-k studies): the
// symbolic pattern stays the base grid's; a scenario replaces the values of the Y-bus entries its branches contribute to (at most
// four per branch, shared entries once) and the parameters of those branches.  Scenario-major arrays with n_branch slots per
// scenario (the largest number of switched branches in the batch); entry == nullptr when the batch has no overlay.
struct DevOverlay {
    int32_t const* entry;     // [n_scn][4 * n_branch] Y-bus entries with replaced values, -1 = unused slot
    double const* y;          // [n_scn][4 * n_branch][B*B][2] replacement values
    int32_t const* branch;    // [n_scn][n_branch] math branches with replaced parameters, -1 = unused slot
    double const* bparam;     // [n_scn][n_branch][4][B*B][2]
    int32_t const* comp;      // [n_scn][n_branch] component index of each branch (lines then transformers), -1 = none
    uint8_t const* energized; // [n_scn][n_branch] its `energized` flag in this scenario
    // buses that lose their supply in a scenario (the switched branches cut them off): their rows become identity rows, their
    // voltage stays 0 and everything on them is reported as not energized
    int32_t const* dead_off;  // [n_scn] index of the scenario's mask in `dead`, -1 = no bus is lost
    uint8_t const* dead;      // [n_mask][n_bus]
    int32_t n_branch;         // slots per scenario (>= 1 when entry != nullptr)
};

// per-batch device buffers, tile layout (see above); B = phases, N = 2B
struct DevBatch {
    int64_t n_scn;
    int32_t n_tile;
    double* jac;   // [tile][nnz_lu][N*N][T]   NR Jacobian / LU factors (col-major blocks)
    double* xvec;  // [tile][n_bus][N][T]      rhs -> forward-substituted -> solution of the linear system
    double* pol;   // [tile][n_bus][N][T]      theta[B], V[B]
    double* u;     // [tile][n_bus][2B][T]     re[B], im[B]
    uint8_t* perm; // [tile][n_bus][2*N][T]    block permutations p[N], q[N]  (sym: packed into one byte: bit0 p, bit1 q)
    double* sinj;  // [tile][n_load_gen][2B][T] re[B], im[B]
    double* usrc;  // [tile][n_source][2][T]   source reference voltage (re, im)
    int32_t* status; // [n_scn]
    int32_t* n_iter; // [n_scn]
    double* max_dev; // [n_scn]
    double* side;    // [tile][n_bus][2][T]      precomputed leaf update terms of the path kernel (nr_sym_v3.cu)
    double* wide_terms; // [tile][wide_max_upd][N*N][T]   update terms of the wide row in flight
    double* wide_rhs;   // [tile][wide_max_lower][N][T]
    double* wide_sum;   // [tile][wide_max_entries][N][T]
    uint8_t* lg_status; // [tile][n_load_gen][T] per-scenario status of each load_gen (device update path), may be null
    uint8_t* qviol;     // [tile][n_bus][T] reactive-power limit a PV bus ran into: 0 none, 1 lower, 2 upper; null = no regulators
    DevOverlay ovl;     // branch-outage overlay (all null when unused)
    unsigned long long* phase_cycles; // optional [n_tile][8] clock64 totals per phase (PGMB_DEBUG_PHASES), may be null
};

// component-level tables of one math group for the device-side update and output kernels (model level)
struct DevModelTables {
    int32_t n_node, n_branch_comp, n_appliance;
    // nodes
    int32_t const* node_id;
    double const* node_u_rated;
    int32_t const* node_bus;      // math bus or -1 (not connected to a source)
    int32_t const* node_app_ptr;  // CSR over nodes
    int32_t const* node_app;      // (kind << 28) | math index ; kind 0 = source, 1 = load_gen; reference summation order
    // branch components: lines then transformers
    int32_t const* branch_id;
    int32_t const* branch_math;   // math branch or -1
    double const* branch_base_i;  // [n][2] from, to
    double const* branch_rating;  // > 0: sn (loading = max_s / sn) ; < 0: -i_n (loading = max_i / i_n)
    uint8_t const* branch_energized;
    // appliances: shunt, source, sym_gen, asym_gen, sym_load, asym_load (concatenated in this order)
    int32_t const* app_id;
    int32_t const* app_math;      // math index inside its kind or -1
    int8_t const* app_kind;       // 0 shunt, 1 source, 2 load_gen
    double const* app_base_i;
    double const* app_dir;        // injection direction
    uint8_t const* app_status;    // shunt / source status (load_gen status is per scenario: DevBatch.lg_status)
    // load update application: per math load_gen
    int8_t const* lg_phases;      // 1 (sym component) or 3 (asym component)
    int8_t const* lg_upd_buf;     // update buffer 0..3 (sym_gen, asym_gen, sym_load, asym_load) or -1 = never updated
    int32_t const* lg_upd_pos;    // element position inside one scenario of that buffer
    int32_t const* lg_upd_id;     // raw id found at that position in scenario 0 (every scenario must repeat it)
    double const* lg_base_s;      // [n_lg][3][2] per-unit specified power of the permanent state
    uint8_t const* lg_base_status;
    double const* lg_scale;       // direction / base_power of the component
};

struct DevUpdateBuffers {
    void const* data[4];          // device copies of the scenario-major update rows
    int64_t n_per_scenario[4];
    int32_t* id_mismatch;         // set to 1 when a scenario's row carries another id than scenario 0 (dependent batch)
};

// every launcher reports here; pgmb_kernel_launch_count() exposes the total (bench.py: gpu_launches)
void count_kernel_launch();

struct SolveOptions {
    int32_t method;
    double err_tol;
    int32_t max_iter;
};

} // namespace pgmb
