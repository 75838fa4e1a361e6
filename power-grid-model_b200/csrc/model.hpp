// Model level (host): the PF slice of the reference's MainModel behind a flat C interface.
//   construction            main_core/input.hpp:49-86 (components built with the rated voltage of their nodes)
//   topology / parameters   calculation_preparation.hpp:131-156, 240-279 ; main_core/y_bus.hpp:186-212
//   PF input                main_core/calculation_input_preparation.hpp:163-188
//   update / restore        main_model_impl.hpp:139-160, 254-261 ; main_core/update.hpp:58-155
//   batch dispatch          job_dispatch.hpp:37-68 -- replaced: scenarios that only change loads / source references are
//                           solved in ONE engine call; other scenarios fall back to per-scenario engine calls (still GPU)
//   output                  main_core/output.hpp:60-187 ; topological_node_output.hpp:72-117
#pragma once

#include "components.hpp"
#include "engine.hpp"
#include "topology.hpp"

#include <chrono>
#include <map>
#include <memory>
#include <unordered_map>

namespace pgmb {

struct ComponentBuffer {
    int64_t n;
    int64_t const* indptr;
    void const* data;
};
struct InputData {
    ComponentBuffer node, line, transformer, shunt, source, sym_gen, asym_gen, sym_load, asym_load, voltage_regulator, asym_line, generic_branch,
        link, three_winding_transformer, transformer_tap_regulator;
};
struct UpdateData {
    int64_t n_scenarios;
    ComponentBuffer line, transformer, shunt, source, sym_gen, asym_gen, sym_load, asym_load, voltage_regulator, asym_line, generic_branch,
        link, three_winding_transformer, transformer_tap_regulator;
};
struct OutputData {
    void *node, *line, *transformer, *shunt, *source, *sym_gen, *asym_gen, *sym_load, *asym_load, *voltage_regulator, *asym_line, *generic_branch,
        *link, *three_winding_transformer, *transformer_tap_regulator;
};
struct ModelOptions {
    int32_t method;
    bool symmetric;
    double err_tol;
    int64_t max_iter;
    int32_t device;
    int32_t threading{-1}; // structural batches: -1 / 0 = all cores, n > 0 = n host threads
    int32_t n_devices{1};  // GPUs one batch is spread over (contiguous scenario blocks), starting at `device`
    uint32_t flags{0};     // PGMB_FLAG_* (include/pgm_b200.h): device-resident update rows / output structs
    // automatic tap changer (PGM_TapChangingStrategy): 0 disabled, 1 any_valid_tap, 2 min_voltage_tap, 3 max_voltage_tap,
    // 4 fast_any_tap
    int32_t tap_strategy{0};
};
constexpr uint32_t kFlagResidentInput = 1u;  // the update rows of this batch are already in HBM (previous call, same buffers)
constexpr uint32_t kFlagResidentOutput = 2u; // leave the output structs in HBM (no copy to the caller's buffers)

// what() of the exception the reference throws for a failed scenario (common/exception.hpp:82-110: SparseMatrixError,
// IterationDiverge with std::format("{}") of max_dev / err_tol = shortest round-trip decimal)
std::string scenario_failure_text(int32_t status, int64_t max_iter, double max_dev, double err_tol);

struct BatchFailure : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class Model {
  public:
    Model(double system_frequency, InputData const& in);

    // single (update == nullptr) or batch calculation; returns number of failed scenarios
    int64_t calculate(ModelOptions const& opt, UpdateData const* update, OutputData const& out, int32_t* n_iter,
                      int32_t* status);
    void update_permanent(UpdateData const& update);
    void batch_pf_input(UpdateData const& update, bool symmetric, Idx group, double* s_injection, double* source_u_ref);

    // introspection for parity tests
    Idx n_math_groups();
    std::vector<int64_t> const& get_index(Idx group, std::string const& name);
    std::vector<double> const& get_real(Idx group, bool symmetric, std::string const& name);
    double timing[8]{}; // pgmb_model_last_timing: [0..5]; [6] device time of the pipeline (pgmb_model_device_pipeline_ms)
    std::string batch_message;

    // PGM_copy_model / PGM_get_indexer (main_model_impl.hpp:218-228) and the element counts the output dataset is checked
    // against; the copy shares no device state with the original (it builds its own engines on first use)
    std::unique_ptr<Model> clone() const;
    void get_indexer(std::string const& component, ID const* ids, Idx size, Idx* indexer) const;
    Idx component_count(std::string const& component) const; // -1: not a component of this library

  private:
    double freq_;
    // static data
    std::vector<NodeInput> node_;
    std::vector<LineInput> line_in_;
    std::vector<LineConst> line_c_;
    std::vector<AsymLineInput> aline_in_;
    std::vector<AsymLineConst> aline_c_;
    std::vector<GenericBranchInput> gb_in_;
    std::vector<GenericBranchConst> gb_c_;
    std::vector<LinkInput> link_in_;
    std::vector<std::array<double, 2>> link_base_i_; // from, to
    std::vector<ThreeWindingConst> t3w_c_;
    std::vector<ThreeWindingState> t3w_st_;
    std::vector<TransformerInput> trafo_in_;
    std::vector<TransformerConst> trafo_c_;
    std::vector<SourceInput> source_in_;
    std::vector<ShuntInput> shunt_in_;
    std::vector<double> shunt_base_y_;
    struct LoadGenStatic {
        ID id;
        Idx node;
        int lb;           // phases of the component
        double direction; // +1 generator, -1 load
        double base_i;
        IntS type;
    };
    std::vector<LoadGenStatic> lg_;
    Idx n_sym_gen_{}, n_asym_gen_{}, n_sym_load_{}, n_asym_load_{};
    // mutable state
    std::vector<BranchState> branch_st_; // branch sequence of the reference: lines, asym_lines, links, generic_branches, transformers
    std::vector<TransformerState> trafo_st_;
    std::vector<SourceState> source_st_;
    std::vector<ShuntState> shunt_st_;
    std::vector<LoadGenState> lg_st_;
    // voltage regulators (component/voltage_regulator.hpp): static part, regulated load_gen (index into lg_), state
    std::vector<VoltageRegulatorInput> reg_in_;
    std::vector<Idx> reg_lg_;
    std::vector<RegulatorState> reg_st_;
    // transformer tap regulators (component/transformer_tap_regulator.hpp): input, regulated transformer (kind 0: transformer,
    // 1: three-winding transformer; index within the kind), rated voltage of the control-side node, state
    std::vector<TransformerTapRegulatorInput> tap_reg_in_;
    struct TapTarget {
        int kind;
        Idx index;
        double u_rated;
    };
    std::vector<TapTarget> tap_reg_target_;
    std::vector<TapRegulatorState> tap_reg_st_;
    // id lookup
    std::unordered_map<ID, Idx> node_idx_, line_idx_, trafo_idx_, shunt_idx_, source_idx_, lg_idx_, reg_idx_, aline_idx_, gb_idx_, link_idx_,
        t3w_idx_, tap_reg_idx_;
    std::unordered_map<ID, int> all_ids_;

    // caches
    bool topo_valid_{false};
    bool param_valid_[2]{false, false};
    TopologyResult topo_;
    struct GroupEngines { // [sym, asym]; a copy of the model starts without engines and builds its own (own CUDA stream)
        std::unique_ptr<Engine> engine[2];
        GroupEngines() = default;
        GroupEngines(GroupEngines const&) {}
        GroupEngines& operator=(GroupEngines const&) {
            engine[0].reset();
            engine[1].reset();
            return *this;
        }
        GroupEngines(GroupEngines&&) = default;
        GroupEngines& operator=(GroupEngines&&) = default;
    };
    std::vector<GroupEngines> engines_;
    int device_{0};
    // In-process multi-GPU (job_dispatch.hpp:131-172 spreads a batch over host threads; here over devices): device k > 0 is
    // driven by a replica of this model (own engines, streams and device tables on that GPU) that is kept across calls and
    // rebuilt when the permanent state changed.  A copy of the model starts without replicas.
    uint64_t state_version_{0};
    struct Replicas {
        struct Item {
            std::unique_ptr<Model> model;
            uint64_t version;
        };
        std::vector<Item> list;
        Replicas() = default;
        Replicas(Replicas const&) {}
        Replicas& operator=(Replicas const&) {
            list.clear();
            return *this;
        }
        Replicas(Replicas&&) = default;
        Replicas& operator=(Replicas&&) = default;
    };
    Replicas replicas_;
    std::map<std::string, std::vector<int64_t>> index_cache_;
    std::map<std::string, std::vector<double>> real_cache_;

    Idx n_line() const { return static_cast<Idx>(line_in_.size()); }
    Idx n_trafo() const { return static_cast<Idx>(trafo_in_.size()); }
    Idx n_aline() const { return static_cast<Idx>(aline_in_.size()); }
    Idx n_gb() const { return static_cast<Idx>(gb_in_.size()); }
    Idx n_link() const { return static_cast<Idx>(link_in_.size()); }
    Idx n_t3w() const { return static_cast<Idx>(t3w_c_.size()); }
    Idx off_aline() const { return n_line(); }
    Idx off_link() const { return n_line() + n_aline(); }
    Idx off_gb() const { return n_line() + n_aline() + n_link(); }
    Idx off_trafo() const { return n_line() + n_aline() + n_link() + n_gb(); }
    Idx n_branch_comp() const { return off_trafo() + n_trafo(); }
    // per branch of the sequence: id, end nodes (sequence numbers), base currents, rating (> 0: sn, loading = max_s / sn;
    // < 0: -i_n, loading = max_i / i_n; +inf: loading 0)
    struct BranchInfo {
        ID id;
        Idx from, to;
        double base_i_from, base_i_to, rating;
    };
    BranchInfo branch_info(Idx b) const;
    Idx node_seq(ID id) const;
    void prepare_topology();
    template <int B> void prepare_engines();
    template <int B> void param_arrays(Idx group, std::vector<double>& bp, std::vector<double>& sp, std::vector<double>& srcp) const;
    // per-scenario inputs of the grids with voltage regulators: regulator parameters of the current state (per group, shared
    // by the scenarios of one engine call) and the status of every load_gen (appended per scenario)
    struct RegulatorInput {
        std::vector<std::vector<double>> param;     // [group][n_reg][4]
        std::vector<std::vector<int8_t>> lg_status; // [group][n_scn][n_lg]
    };
    template <int B>
    void gather_pf_input(std::vector<std::vector<double>>& sinj, std::vector<std::vector<double>>& uref,
                         RegulatorInput* reg = nullptr) const;
    template <int B> void check_regulators(ModelOptions const& opt) const;

    struct Saved {
        std::vector<std::pair<Idx, BranchState>> branch;
        std::vector<std::pair<Idx, TransformerState>> trafo;
        std::vector<std::pair<Idx, SourceState>> source;
        std::vector<std::pair<Idx, ShuntState>> shunt;
        std::vector<std::pair<Idx, LoadGenState>> lg;
        std::vector<std::pair<Idx, RegulatorState>> reg;
        std::vector<std::pair<Idx, ThreeWindingState>> t3w;
        std::vector<std::pair<Idx, TapRegulatorState>> tap_reg;
        bool topo{false}, param{false};
    };
    void apply_scenario(UpdateData const& u, Idx s, Saved* saved);
    void restore(Saved const& saved);
    void mark(bool topo, bool param, Saved* saved);
    void set_load_power(Idx i, double const* p, double const* q);

    template <int B>
    int64_t calculate_impl(ModelOptions const& opt, UpdateData const* update, OutputData const& out, int32_t* n_iter,
                           int32_t* status);
    // solver output of one engine call per math group: [0] bus voltages, [2] branch flows, [3] sources, [4] shunts, [5] load_gens
    struct BlockSolution {
        std::vector<std::vector<double>> so[6];
        std::vector<std::vector<int8_t>> reg_out;
        std::vector<int32_t> status, n_iter;
        std::vector<double> max_dev;
    };
    template <int B>
    void solve_block(ModelOptions const& opt, Idx n_scn, std::vector<std::vector<double>> const& sinj,
                     std::vector<std::vector<double>> const& uref, RegulatorInput const* reg, BlockSolution& sol);
    template <int B>
    int64_t run_block(ModelOptions const& opt, Idx n_scn, std::vector<std::vector<double>> const& sinj,
                      std::vector<std::vector<double>> const& uref, OutputData const& out, Idx first_scenario,
                      int32_t* n_iter, int32_t* status, RegulatorInput const* reg = nullptr);
    // ---- automatic tap changer (model_tap.cpp; optimizer/tap_position_optimizer.hpp) ----
    // one scenario in the model's current state: ranks the regulated transformers, searches their tap positions with repeated
    // power flows on the GPU, writes the outputs of the final power flow and puts the tap positions back
    template <int B>
    int64_t run_tap_optimizer(ModelOptions const& opt, OutputData const& out, Idx scenario, int32_t* n_iter, int32_t* status);
    // lockstep search over a load-profile batch with one regulated transformer: one batched power flow per search step
    template <int B>
    int64_t run_tap_lockstep(ModelOptions const& opt, UpdateData const& update, OutputData const& out, int32_t* n_iter, int32_t* status,
                             std::vector<Idx>& exact);
    struct TapRanked;
    std::vector<std::vector<TapRanked>> rank_tap_regulators() const;
    std::vector<int64_t> tap_rank_table() const; // introspection: (kind, index, rank group) in the order of the search
    std::vector<IntS> tap_positions_out_; // per tap regulator: tap position found by the last run_tap_optimizer (na: not regulated)
    // ---- branch-outage batches on the shared symbolic pattern (N-1 studies) ----
    // A scenario that opens a few fully connected branches, or moves transformer taps, has the base grid's equations with those
    // branches' admittance contributions changed: it is solved on the base topology's pattern with the contributions replaced
    // (Engine::set_overlay) instead of rebuilding topology, ordering and pattern for it as the reference does
    // (main_model_impl.hpp:139-160 -> rebuild_topology); buses that lose their supply are masked.  Same equations in another
    // elimination order: results agree to rounding, not bit for bit.  Other scenarios (a second supplied island, more than
    // kMaxOutageSlots branches, ...) take the exact per-scenario route.
    static constexpr int kMaxOutageSlots = 8; // switched branches per scenario the overlay carries (N-k)
    // Scenarios that CLOSE a branch which is open in the base state: the pattern of the base grid has no entries for it, so the
    // batch runs on a copy of the model in which every branch some scenario closes is closed (the union grid); on that copy a
    // scenario is the union grid with branches switched OFF -- the ones it does not close (implicit, outage_base_state_) and the
    // ones it opens itself -- which is the N-k overlay again.
    struct BranchSwitch {
        Idx branch;
        bool from, to;
    };
    std::vector<BranchSwitch> outage_base_state_; // union copy only: the base state of the branches the copy closed
    std::shared_ptr<Model> union_model_;
    std::vector<Idx> union_key_;
    uint64_t union_version_{0};
    std::vector<Idx> closing_branches(UpdateData const& update);
    Model* union_model(std::vector<Idx> const& closing);
    Model* outage_host(UpdateData const& update); // this, or the union copy when scenarios close branches
public:
    // introspection of the host planning (CPU tests): per scenario [0] route (0 shared pattern, 1 own topology), [1] overlay
    // slots in use, [2] buses that lose their supply, [3] planned on the union grid
    void outage_plan_summary(UpdateData const& update, bool symmetric, int64_t* out);
private:
    struct OutagePlan {
        int n_slot{1};                    // branch slots per scenario: the arrays below are [n_scn][n_slot]
        std::vector<int64_t> math_branch; // -1 = unused slot
        std::vector<double> bparam;       // [n_scn][n_slot][4][B*B][2]
        std::vector<int32_t> comp;
        std::vector<uint8_t> energized;
        std::vector<int32_t> dead_off; // per scenario: mask of the buses that lose their supply (bridge outages), -1 = none
        std::vector<uint8_t> dead;     // [n_mask][n_bus]
        std::vector<Idx> exact;
    };
    // DFS over the fully connected branches: bridges and, for each bridge, the subtree it cuts off
    struct BridgeInfo {
        std::vector<char> bridge;     // per branch
        std::vector<Idx> child;       // per branch: the DFS child end of a bridge
        std::vector<Idx> disc, size, n_source, root, order; // per node: discovery time, subtree size / active sources, DFS root; node by time
        std::vector<Idx> self_source;                  // per node: active sources on the node itself
        std::vector<Idx> adj_ptr, adj_node, adj_edge;  // the graph of the fully connected branches (CSR by node)
    };
    BridgeInfo bridge_analysis() const;
    OutagePlan const* outage_plan_{nullptr};
    // device pass over a GROUP of scenarios (model.cpp: grouped route): scenario i of the pass delivers its output rows to
    // scenario out_scatter_[i] of the caller's buffers (which are then the whole batch's buffers, not a slice)
    Idx const* out_scatter_{nullptr};
    template <int B> bool plan_outage_batch(UpdateData const& update, OutagePlan& plan) const;
    // ---- device path (model_device.cpp): updates applied and output structs written by CUDA kernels ----
    struct DeviceSide;
    std::shared_ptr<DeviceSide> dev_;
    bool device_path_eligible(UpdateData const& update) const;
    void fetch_resident_rows(int slot, size_t row_bytes, Idx count, Idx index, Idx n_scn, void* dst) const;
    int last_pass_parts_{1}; // parts the last run_batch_device_one split its batch into (device-memory budget)
    int64_t run_batch_device(ModelOptions const& opt, int phases, UpdateData const& update, OutputData const& out, int32_t* n_iter,
                             int32_t* status);
    int64_t run_batch_device_one(ModelOptions const& opt, int phases, UpdateData const& update, OutputData const& out,
                                 int32_t* n_iter, int32_t* status, Idx first_scenario);
    int64_t run_batch_device_part(ModelOptions const& opt, int phases, UpdateData const& update, OutputData const& out,
                                  int32_t* n_iter, int32_t* status, Idx first_scenario);
    template <int B>
    void write_output(Idx n_scn, Idx first_scenario, OutputData const& out,
                      std::vector<std::vector<double>> const (&so)[6], std::vector<std::vector<int8_t>> const& reg_out,
                      std::vector<std::vector<int8_t>> const* lg_status) const;
};

} // namespace pgmb
