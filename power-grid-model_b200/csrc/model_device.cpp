// Device path of the model level: for batches that only change loads / generators (and source references), the raw update
// rows are copied to the GPU once, applied there for every scenario, solved, and the caller's output structs are written by
// kernels straight into a device image of the caller's buffers, which is then copied back with one DMA per component.
// Host work per batch: one pass over the ids of scenario 0 (update element -> load_gen mapping), pinned-memory detection,
// kernel launches.  Replaces, for such batches, the per-scenario update -> prepare_power_flow_input -> output_result ->
// restore loop of the reference (job_dispatch.hpp:88-138, job_adapter.hpp:127-139).
#include "model.hpp"

#include <thread>

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace pgmb {

namespace {
using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

bool is_device_accessible_host(void const* p) { // pinned (page-locked) host memory?
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}
} // namespace

struct Model::DeviceSide {
    // static tables (valid while the topology / permanent state is unchanged)
    DevBuf<int32_t> node_id, node_bus, node_app_ptr, node_app, branch_id, branch_math, app_id, app_math, lg_upd_pos;
    DevBuf<double> node_u_rated, branch_base_i, branch_rating, app_base_i, app_dir, lg_base_s, lg_scale;
    DevBuf<uint8_t> branch_energized, app_status, lg_base_status;
    DevBuf<int8_t> app_kind, lg_phases, lg_upd_buf;
    DevBuf<int32_t> lg_upd_id;
    DevModelTables t{};
    Idx n_app_first[6]{}; // first appliance index of shunt, source, sym_gen, asym_gen, sym_load, asym_load
    // per-batch buffers
    DevBuf<unsigned char> upd[4];
    DevBuf<unsigned char> out[13];
    DevBuf<double> src_res;
    // voltage regulators: component tables and the per-scenario flags of the math regulators
    DevBuf<int32_t> reg_id, reg_math;
    DevBuf<uint8_t> reg_status;
    DevBuf<int8_t> reg_flags;
    DevBuf<int32_t> flag;
    // pinned staging for pageable caller buffers
    unsigned char* pinned{nullptr};
    size_t pinned_size{0};
    // chunk pipeline: one stream per chunk in flight, fork event on the engine stream, solve brackets per chunk
    static constexpr int kMaxChunk = 8;
    cudaStream_t cs[kMaxChunk]{};
    cudaEvent_t fork{}, ev_a[kMaxChunk]{}, ev_b[kMaxChunk]{}, ev_end[kMaxChunk]{};
    cudaEvent_t ev_p0{}, ev_p1{}; // the whole pipeline of one part on the engine stream (fork .. join of every chunk)
    bool streams_ready{false};
    int device{-1}; // the GPU the streams, events and tables live on
    size_t resident_rows[4]{}; // bytes of update rows per component uploaded by the last call (PGMB_FLAG_RESIDENT_INPUT)
    void ensure_streams(int dev) {
        PGMB_CUDA(cudaSetDevice(dev));
        if (streams_ready) return;
        device = dev;
        for (auto& q : cs) PGMB_CUDA(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
        PGMB_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        for (auto& ev : ev_a) PGMB_CUDA(cudaEventCreate(&ev));
        for (auto& ev : ev_b) PGMB_CUDA(cudaEventCreate(&ev));
        for (auto& ev : ev_end) PGMB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        PGMB_CUDA(cudaEventCreate(&ev_p0));
        PGMB_CUDA(cudaEventCreate(&ev_p1));
        streams_ready = true;
    }
    void drain() noexcept { // nothing of this pipeline may still read or write caller / staging memory after an error
        if (!streams_ready) return;
        if (device >= 0) cudaSetDevice(device);
        for (auto& q : cs) cudaStreamSynchronize(q);
    }
    ~DeviceSide() {
        if (device >= 0) cudaSetDevice(device);
        if (pinned != nullptr) cudaFreeHost(pinned);
        if (streams_ready) {
            for (auto& q : cs) cudaStreamSynchronize(q); // the buffers go back to the pool: nothing may still use them
            for (auto& q : cs) cudaStreamDestroy(q);
            cudaEventDestroy(fork);
            for (auto& ev : ev_a) cudaEventDestroy(ev);
            for (auto& ev : ev_b) cudaEventDestroy(ev);
            for (auto& ev : ev_end) cudaEventDestroy(ev);
            cudaEventDestroy(ev_p0);
            cudaEventDestroy(ev_p1);
        }
    }
    unsigned char* staging(size_t bytes) {
        if (bytes > pinned_size) {
            if (pinned != nullptr) cudaFreeHost(pinned);
            pinned = nullptr;
            PGMB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&pinned), bytes));
            pinned_size = bytes;
        }
        return pinned;
    }
};

// rows `index` of one output component (slot as in the request table of run_batch_device_part) for every scenario of the last
// single-device pass that left its output structs in HBM (kFlagResidentOutput): a strided copy of n_scn rows
void Model::fetch_resident_rows(int slot, size_t row_bytes, Idx count, Idx index, Idx n_scn, void* dst) const {
    if (!dev_ || dev_->device < 0 || slot < 0 || slot >= 13) throw InvalidArgument("internal: no device-resident output to read");
    PGMB_CUDA(cudaSetDevice(dev_->device));
    PGMB_CUDA(cudaMemcpy2D(dst, row_bytes, dev_->out[slot].get() + static_cast<size_t>(index) * row_bytes, static_cast<size_t>(count) * row_bytes,
                           row_bytes, static_cast<size_t>(n_scn), cudaMemcpyDeviceToHost));
}

bool Model::device_path_eligible(UpdateData const& u) const {
    if (topo_.math.size() != 1) return false;
    if (n_t3w() != 0) return false; // three-winding transformer output is converted on the host (write_output)
    ComponentBuffer const* bufs[4] = {&u.sym_gen, &u.asym_gen, &u.sym_load, &u.asym_load};
    for (auto const* b : bufs) {
        if (b->data != nullptr && (b->indptr != nullptr || b->n < 0)) return false; // sparse batches: host path
    }
    return u.n_scenarios > 0;
}

namespace {
// views of the caller's update / output buffers for the scenarios [s0, s0 + ns) of a uniform batch
struct OutPart {
    size_t row;
    Idx count;
};
void slice_batch(UpdateData const& update, OutputData const& out, OutPart const (&outs)[13], Idx s0, Idx ns, UpdateData& u, OutputData& o) {
    size_t const urow[4] = {sizeof(SymLoadGenUpdate), sizeof(AsymLoadGenUpdate), sizeof(SymLoadGenUpdate), sizeof(AsymLoadGenUpdate)};
    u = update;
    u.n_scenarios = ns;
    ComponentBuffer* ub[4] = {&u.sym_gen, &u.asym_gen, &u.sym_load, &u.asym_load};
    for (int b = 0; b != 4; ++b)
        if (ub[b]->data != nullptr) ub[b]->data = static_cast<unsigned char const*>(ub[b]->data) + static_cast<size_t>(s0) * ub[b]->n * urow[b];
    if (u.source.data != nullptr) {
        if (u.source.indptr != nullptr) {
            u.source.indptr += s0; // sparse: scenario s of the part is scenario s0 + s of the caller's index pointer
        } else {
            u.source.data = static_cast<unsigned char const*>(u.source.data) + static_cast<size_t>(s0) * u.source.n * sizeof(SourceUpdate);
        }
    }
    o = out;
    void** ohost[13] = {&o.node, &o.line, &o.transformer, &o.shunt, &o.source, &o.sym_gen, &o.asym_gen, &o.sym_load, &o.asym_load,
                        &o.voltage_regulator, &o.asym_line, &o.generic_branch, &o.link};
    for (int k = 0; k != 13; ++k)
        if (*ohost[k] != nullptr) *ohost[k] = static_cast<unsigned char*>(*ohost[k]) + static_cast<size_t>(s0) * outs[k].count * outs[k].row;
}
} // namespace

// returns number of failed scenarios, or -1 when the batch turned out not to be uniform (caller falls back to the host path)
//
// In-process multi-GPU: the reference fans one batch out over host threads inside the call (job_dispatch.hpp:131-172); here the
// batch is cut into contiguous scenario blocks, one per device (opt.n_devices / PGMB_DEVICES, starting at opt.device).  Device 0
// of the call is driven by this model on the calling thread, every further device by a replica of the model on its own host
// thread (own engines: the symbolic structures are rebuilt from the same math topology and uploaded to that device; own
// streams, tables and staging).  The blocks write disjoint slices of the caller's buffers; the per-scenario status arrays and
// the batch messages are concatenated in scenario order.  No collective, no peer traffic.
int64_t Model::run_batch_device(ModelOptions const& opt, int phases, UpdateData const& update, OutputData const& out, int32_t* n_iter,
                                int32_t* status) {
    Idx const n_scn = update.n_scenarios;
    int n_dev = opt.n_devices;
    if (n_dev <= 1) {
        char const* env = std::getenv("PGMB_DEVICES");
        n_dev = env != nullptr ? std::atoi(env) : 1;
    }
    int available = 0;
    if (cudaGetDeviceCount(&available) != cudaSuccess) available = 0;
    n_dev = std::max(1, std::min(n_dev, available - opt.device));
    Idx min_per_dev = 256;
    if (char const* env = std::getenv("PGMB_MIN_SCENARIOS_PER_DEVICE")) min_per_dev = std::max<Idx>(1, std::atoll(env));
    n_dev = static_cast<int>(std::max<Idx>(1, std::min<Idx>(n_dev, n_scn / min_per_dev)));
    if (n_dev == 1) return run_batch_device_one(opt, phases, update, out, n_iter, status, 0);

    bool const sym = phases == 1;
    size_t const row_node = sym ? sizeof(NodeOutput<1>) : sizeof(NodeOutput<3>);
    size_t const row_branch = sym ? sizeof(BranchOutput<1>) : sizeof(BranchOutput<3>);
    size_t const row_app = sym ? sizeof(ApplianceOutput<1>) : sizeof(ApplianceOutput<3>);
    OutPart const outs[13] = {{row_node, static_cast<Idx>(node_.size())}, {row_branch, n_line()}, {row_branch, n_trafo()},
                              {row_app, static_cast<Idx>(shunt_in_.size())}, {row_app, static_cast<Idx>(source_in_.size())},
                              {row_app, n_sym_gen_}, {row_app, n_asym_gen_}, {row_app, n_sym_load_}, {row_app, n_asym_load_},
                              {sizeof(VoltageRegulatorOutput), static_cast<Idx>(reg_in_.size())}, {row_branch, n_aline()}, {row_branch, n_gb()},
                              {row_branch, n_link()}};
    // replicas for devices 1 .. n_dev - 1 (kept across calls; rebuilt when the permanent state has changed since)
    if (replicas_.list.size() < static_cast<size_t>(n_dev - 1)) replicas_.list.resize(n_dev - 1);
    for (int k = 1; k != n_dev; ++k) {
        auto& r = replicas_.list[k - 1];
        if (!r.model || r.version != state_version_ || r.model->device_ != opt.device + k) {
            r.model = clone();
            r.model->device_ = opt.device + k;
            r.version = state_version_;
        }
    }
    // contiguous blocks, cut at multiples of 32 scenarios (whole tiles whatever the tile width)
    Idx const block = ((n_scn + n_dev - 1) / n_dev + 31) / 32 * 32;
    std::vector<int64_t> failed(n_dev, 0);
    std::vector<std::exception_ptr> error(n_dev);
    auto run_on = [&](int k) {
        try {
            Idx const s0 = std::min<Idx>(n_scn, block * k), ns = std::min<Idx>(n_scn, block * (k + 1)) - s0;
            if (ns <= 0) return;
            UpdateData u;
            OutputData o;
            slice_batch(update, out, outs, s0, ns, u, o);
            Model& m = k == 0 ? *this : *replicas_.list[k - 1].model;
            ModelOptions mo = opt;
            mo.device = opt.device + k;
            mo.n_devices = 1;
            if (k != 0) {
                m.outage_plan_ = outage_plan_;
                m.batch_message.clear();
                for (double& t : m.timing) t = 0.0;
                if (phases == 1) {
                    m.prepare_engines<1>();
                } else {
                    m.prepare_engines<3>();
                }
            }
            failed[k] = m.run_batch_device_one(mo, phases, u, o, n_iter ? n_iter + s0 : nullptr, status ? status + s0 : nullptr, s0);
        } catch (...) {
            error[k] = std::current_exception();
        }
        if (k != 0) replicas_.list[k - 1].model->outage_plan_ = nullptr;
    };
    {
        std::vector<std::thread> pool;
        for (int k = 1; k != n_dev; ++k) pool.emplace_back(run_on, k);
        run_on(0);
        for (auto& th : pool) th.join();
    }
    PGMB_CUDA(cudaSetDevice(opt.device));
    for (auto const& ex : error)
        if (ex) std::rethrow_exception(ex);
    int64_t total = 0;
    for (int k = 0; k != n_dev; ++k) {
        if (failed[k] < 0) return -1;
        total += failed[k];
    }
    for (int k = 1; k != n_dev; ++k) {
        Model const& m = *replicas_.list[k - 1].model;
        batch_message += m.batch_message;
        for (int i = 0; i != 7; ++i)
            if (i != 5) timing[i] = std::max(timing[i], m.timing[i]); // the devices work side by side
    }
    return total;
}

// One device.  A batch whose working set would not fit the device is cut into parts that are run one after another; the parts
// see offset views of the caller's update / output buffers.  first_scenario = number of the block's first scenario in the
// caller's batch (messages, outage plan).
int64_t Model::run_batch_device_one(ModelOptions const& opt, int phases, UpdateData const& update, OutputData const& out,
                                    int32_t* n_iter, int32_t* status, Idx first_scenario) {
    bool const sym = phases == 1;
    Engine& e = *engines_[0].engine[sym ? 0 : 1];
    MathTopology const& m = topo_.math[0];
    size_t const N = 2 * static_cast<size_t>(phases);
    size_t const row_node = sym ? sizeof(NodeOutput<1>) : sizeof(NodeOutput<3>);
    size_t const row_branch = sym ? sizeof(BranchOutput<1>) : sizeof(BranchOutput<3>);
    size_t const row_app = sym ? sizeof(ApplianceOutput<1>) : sizeof(ApplianceOutput<3>);
    struct Part {
        void* const* host;
        size_t row;
        Idx count;
    };
    Part const outs[13] = {{&out.node, row_node, static_cast<Idx>(node_.size())},
                          {&out.line, row_branch, n_line()},
                          {&out.transformer, row_branch, n_trafo()},
                          {&out.shunt, row_app, static_cast<Idx>(shunt_in_.size())},
                          {&out.source, row_app, static_cast<Idx>(source_in_.size())},
                          {&out.sym_gen, row_app, n_sym_gen_},
                          {&out.asym_gen, row_app, n_asym_gen_},
                          {&out.sym_load, row_app, n_sym_load_},
                          {&out.asym_load, row_app, n_asym_load_},
                          {&out.voltage_regulator, sizeof(VoltageRegulatorOutput), static_cast<Idx>(reg_in_.size())},
                          {&out.asym_line, row_branch, n_aline()},
                          {&out.generic_branch, row_branch, n_gb()},
                          {&out.link, row_branch, n_link()}};
    ComponentBuffer const* ubufs[4] = {&update.sym_gen, &update.asym_gen, &update.sym_load, &update.asym_load};
    size_t const urow[4] = {sizeof(SymLoadGenUpdate), sizeof(AsymLoadGenUpdate), sizeof(SymLoadGenUpdate), sizeof(AsymLoadGenUpdate)};
    size_t per_scn = static_cast<size_t>(e.pattern().nnz_lu) * N * N * 8 + 6 * static_cast<size_t>(m.n_bus) * N * 8 +
                     static_cast<size_t>(m.n_load_gen()) * (N * 8 + 1) +
                     (static_cast<size_t>(e.wide_plan().max_upd) * N * N + static_cast<size_t>(e.wide_plan().max_lower + e.wide_plan().max_entries) * N) * 8;
    for (Part const& o : outs)
        if (*o.host != nullptr) per_scn += o.row * static_cast<size_t>(o.count);
    for (int b = 0; b != 4; ++b)
        if (ubufs[b]->data != nullptr) per_scn += urow[b] * static_cast<size_t>(ubufs[b]->n);
    // total memory of the device, queried once: cudaMemGetInfo costs milliseconds in a process with many allocations
    static size_t total_cache[64] = {};
    int const dev_slot = e.device() % 64;
    PGMB_CUDA(cudaSetDevice(e.device())); // every allocation / stream below belongs to the engine's device
    if (total_cache[dev_slot] == 0) {
        size_t free_b = 0, total_b = 0;
        PGMB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        total_cache[dev_slot] = total_b;
    }
    size_t budget = total_cache[dev_slot] / 10 * 6; // the pool may hold freed blocks, so the total is the yardstick
    if (char const* env = std::getenv("PGMB_MAX_BATCH_BYTES")) budget = static_cast<size_t>(std::atoll(env));
    Idx const n_scn = update.n_scenarios;
    Idx max_scn = static_cast<Idx>(std::max<size_t>(32, budget / per_scn / 32 * 32));
    last_pass_parts_ = n_scn <= max_scn ? 1 : static_cast<int>((n_scn + max_scn - 1) / max_scn);
    if (n_scn <= max_scn) return run_batch_device_part(opt, phases, update, out, n_iter, status, first_scenario);
    OutPart parts[13];
    for (int k = 0; k != 13; ++k) parts[k] = {outs[k].row, outs[k].count};
    int64_t failed = 0;
    for (Idx s0 = 0; s0 < n_scn; s0 += max_scn) {
        Idx const ns = std::min(max_scn, n_scn - s0);
        UpdateData u;
        OutputData o;
        slice_batch(update, out, parts, s0, ns, u, o);
        if (out_scatter_ != nullptr) o = out; // scattered delivery addresses the caller's buffers by scenario number
        ModelOptions mo = opt;
        mo.flags = 0; // parts reuse the device buffers: nothing stays resident
        int64_t const r = run_batch_device_part(mo, phases, u, o, n_iter ? n_iter + s0 : nullptr, status ? status + s0 : nullptr,
                                                first_scenario + s0);
        if (r < 0) return -1;
        failed += r;
    }
    return failed;
}

int64_t Model::run_batch_device_part(ModelOptions const& opt, int phases, UpdateData const& update, OutputData const& out,
                                     int32_t* n_iter, int32_t* status, Idx first_scenario) {
    auto t0 = Clock::now();
    bool const sym = phases == 1;
    Engine& e = *engines_[0].engine[sym ? 0 : 1];
    cudaStream_t const st = e.stream();
    MathTopology const& m = topo_.math[0];
    Idx const n_scn = update.n_scenarios;
    Idx const nn = static_cast<Idx>(node_.size());
    Idx const n_lg_math = m.n_load_gen();
    // the engine's device is the current device of this thread for everything below (pool allocations key on it); device-side
    // state that lives on another GPU (the caller switched devices between calls) is dropped
    PGMB_CUDA(cudaSetDevice(e.device()));
    if (dev_ && dev_->device >= 0 && dev_->device != e.device()) dev_.reset();
    if (!dev_) dev_ = std::make_shared<DeviceSide>();
    DeviceSide& d = *dev_;
    d.ensure_streams(e.device());
    bool const resident_in = (opt.flags & kFlagResidentInput) != 0;
    bool const resident_out = (opt.flags & kFlagResidentOutput) != 0;

    // ---- tables (rebuilt per call: a few thousand elements; the permanent state may have changed) ----
    {
        std::vector<int32_t> node_id(nn), node_bus(nn), app_ptr(nn + 1, 0), app;
        std::vector<double> u_rated(nn);
        std::vector<std::vector<int32_t>> per_node(nn);
        for (size_t i = 0; i != source_in_.size(); ++i)
            if (topo_.source[i].group == 0) per_node[node_idx_.at(source_in_[i].node)].push_back(static_cast<int32_t>(topo_.source[i].pos));
        Idx const o_sg = 0, o_ag = n_sym_gen_, o_sl = n_sym_gen_ + n_asym_gen_, o_al = o_sl + n_sym_load_;
        auto add_lg = [&](Idx begin, Idx count) {
            for (Idx i = begin; i != begin + count; ++i)
                if (topo_.load_gen[i].group == 0) per_node[lg_[i].node].push_back(static_cast<int32_t>((1u << 28) | topo_.load_gen[i].pos));
        };
        add_lg(o_sl, n_sym_load_);
        add_lg(o_sg, n_sym_gen_);
        add_lg(o_al, n_asym_load_);
        add_lg(o_ag, n_asym_gen_);
        for (Idx i = 0; i != nn; ++i) {
            node_id[i] = node_[i].id;
            u_rated[i] = node_[i].u_rated;
            node_bus[i] = topo_.node[i].group == 0 ? static_cast<int32_t>(topo_.node[i].pos) : -1;
            app.insert(app.end(), per_node[i].begin(), per_node[i].end());
            app_ptr[i + 1] = static_cast<int32_t>(app.size());
        }
        Idx const nb = n_branch_comp();
        std::vector<int32_t> b_id(nb), b_math(nb);
        std::vector<double> b_base(2 * nb), b_rating(nb);
        std::vector<uint8_t> b_en(nb);
        for (Idx i = 0; i != nb; ++i) {
            BranchInfo const info = branch_info(i);
            b_id[i] = info.id;
            b_math[i] = topo_.branch[i].group == 0 ? static_cast<int32_t>(topo_.branch[i].pos) : -1;
            b_base[2 * i] = info.base_i_from;
            b_base[2 * i + 1] = info.base_i_to;
            b_rating[i] = info.rating; // > 0: sn ; < 0: -i_n ; +inf: loading 0
            b_en[i] = (branch_st_[i].from_status || branch_st_[i].to_status) ? 1 : 0;
        }
        Idx const n_app = static_cast<Idx>(shunt_in_.size() + source_in_.size() + lg_.size());
        std::vector<int32_t> a_id(n_app), a_math(n_app);
        std::vector<int8_t> a_kind(n_app);
        std::vector<double> a_base(n_app), a_dir(n_app);
        std::vector<uint8_t> a_status(n_app);
        Idx k = 0;
        d.n_app_first[0] = k;
        for (size_t i = 0; i != shunt_in_.size(); ++i, ++k) {
            a_id[k] = shunt_in_[i].id;
            a_math[k] = topo_.shunt[i].group == 0 ? static_cast<int32_t>(topo_.shunt[i].pos) : -1;
            a_kind[k] = 0;
            a_base[k] = kBasePower3p / node_[node_idx_.at(shunt_in_[i].node)].u_rated / kSqrt3;
            a_dir[k] = -1.0;
            a_status[k] = shunt_st_[i].status ? 1 : 0;
        }
        d.n_app_first[1] = k;
        for (size_t i = 0; i != source_in_.size(); ++i, ++k) {
            a_id[k] = source_in_[i].id;
            a_math[k] = topo_.source[i].group == 0 ? static_cast<int32_t>(topo_.source[i].pos) : -1;
            a_kind[k] = 1;
            a_base[k] = kBasePower3p / node_[node_idx_.at(source_in_[i].node)].u_rated / kSqrt3;
            a_dir[k] = 1.0;
            a_status[k] = source_st_[i].status ? 1 : 0;
        }
        d.n_app_first[2] = k;
        d.n_app_first[3] = k + n_sym_gen_;
        d.n_app_first[4] = k + n_sym_gen_ + n_asym_gen_;
        d.n_app_first[5] = k + n_sym_gen_ + n_asym_gen_ + n_sym_load_;
        for (size_t i = 0; i != lg_.size(); ++i, ++k) {
            a_id[k] = lg_[i].id;
            a_math[k] = topo_.load_gen[i].group == 0 ? static_cast<int32_t>(topo_.load_gen[i].pos) : -1;
            a_kind[k] = 2;
            a_base[k] = lg_[i].base_i;
            a_dir[k] = lg_[i].direction;
            a_status[k] = lg_st_[i].status ? 1 : 0;
        }
        // load update mapping from the ids of scenario 0 (an independent batch repeats them in every scenario)
        std::vector<int8_t> lg_phases(n_lg_math, 1), lg_buf(n_lg_math, -1);
        std::vector<int32_t> lg_pos(n_lg_math, 0), lg_id(n_lg_math, 0);
        std::vector<double> lg_base(n_lg_math * 6, 0.0), lg_scale(n_lg_math, 0.0);
        std::vector<uint8_t> lg_st(n_lg_math, 0);
        for (size_t i = 0; i != lg_.size(); ++i) {
            Coupling const c = topo_.load_gen[i];
            if (c.group != 0) continue;
            lg_phases[c.pos] = static_cast<int8_t>(lg_[i].lb);
            for (int p = 0; p != 3; ++p) {
                lg_base[(c.pos * 3 + p) * 2] = lg_st_[i].s[p].real();
                lg_base[(c.pos * 3 + p) * 2 + 1] = lg_st_[i].s[p].imag();
            }
            lg_scale[c.pos] = lg_[i].direction / (lg_[i].lb == 1 ? kBasePower3p : kBasePower1p);
            lg_st[c.pos] = lg_st_[i].status ? 1 : 0;
        }
        ComponentBuffer const* bufs[4] = {&update.sym_gen, &update.asym_gen, &update.sym_load, &update.asym_load};
        Idx const first[4] = {0, n_sym_gen_, n_sym_gen_ + n_asym_gen_, n_sym_gen_ + n_asym_gen_ + n_sym_load_};
        Idx const count[4] = {n_sym_gen_, n_asym_gen_, n_sym_load_, n_asym_load_};
        size_t const row_size[4] = {sizeof(SymLoadGenUpdate), sizeof(AsymLoadGenUpdate), sizeof(SymLoadGenUpdate), sizeof(AsymLoadGenUpdate)};
        for (int bfr = 0; bfr != 4; ++bfr) {
            if (bufs[bfr]->data == nullptr) continue;
            auto const* base = static_cast<unsigned char const*>(bufs[bfr]->data);
            std::vector<char> seen(count[bfr], 0);
            for (Idx kpos = 0; kpos != bufs[bfr]->n; ++kpos) {
                ID id;
                std::memcpy(&id, base + kpos * row_size[bfr], sizeof(ID));
                Idx seq;
                if (id == kNaID) {
                    if (bufs[bfr]->n != count[bfr]) throw InvalidArgument("update without ids must cover every element of the component");
                    seq = first[bfr] + kpos;
                } else {
                    auto it = lg_idx_.find(id);
                    if (it == lg_idx_.end() || it->second < first[bfr] || it->second >= first[bfr] + count[bfr])
                        throw InvalidArgument("The id cannot be found: " + std::to_string(id) + "\n");
                    seq = it->second;
                }
                if (seen[seq - first[bfr]]) return -1; // same element updated twice in one scenario: host path keeps the order
                seen[seq - first[bfr]] = 1;
                Coupling const c = topo_.load_gen[seq];
                if (c.group != 0) continue;
                lg_buf[c.pos] = static_cast<int8_t>(bfr);
                lg_pos[c.pos] = static_cast<int32_t>(kpos);
                lg_id[c.pos] = id;
            }
        }
        d.node_id.upload(node_id, st);
        d.node_bus.upload(node_bus, st);
        d.node_app_ptr.upload(app_ptr, st);
        d.node_app.upload(app, st);
        d.node_u_rated.upload(u_rated, st);
        d.branch_id.upload(b_id, st);
        d.branch_math.upload(b_math, st);
        d.branch_base_i.upload(b_base, st);
        d.branch_rating.upload(b_rating, st);
        d.branch_energized.upload(b_en, st);
        d.app_id.upload(a_id, st);
        d.app_math.upload(a_math, st);
        d.app_kind.upload(a_kind, st);
        d.app_base_i.upload(a_base, st);
        d.app_dir.upload(a_dir, st);
        d.app_status.upload(a_status, st);
        d.lg_phases.upload(lg_phases, st);
        d.lg_upd_buf.upload(lg_buf, st);
        d.lg_upd_pos.upload(lg_pos, st);
        d.lg_upd_id.upload(lg_id, st);
        d.lg_base_s.upload(lg_base, st);
        d.lg_base_status.upload(lg_st, st);
        d.lg_scale.upload(lg_scale, st);
        std::vector<int32_t> r_id(reg_in_.size()), r_math(reg_in_.size());
        std::vector<uint8_t> r_status(reg_in_.size());
        for (size_t i = 0; i != reg_in_.size(); ++i) {
            r_id[i] = reg_in_[i].id;
            r_math[i] = topo_.voltage_regulator[i].group == 0 ? static_cast<int32_t>(topo_.voltage_regulator[i].pos) : -1;
            r_status[i] = reg_st_[i].status ? 1 : 0;
        }
        d.reg_id.upload(r_id, st);
        d.reg_math.upload(r_math, st);
        d.reg_status.upload(r_status, st);
        PGMB_CUDA(cudaStreamSynchronize(st)); // host vectors die at the end of this scope
        d.t = DevModelTables{static_cast<int32_t>(nn), static_cast<int32_t>(nb), static_cast<int32_t>(n_app),
                             d.node_id.get(), d.node_u_rated.get(), d.node_bus.get(), d.node_app_ptr.get(), d.node_app.get(),
                             d.branch_id.get(), d.branch_math.get(), d.branch_base_i.get(), d.branch_rating.get(),
                             d.branch_energized.get(), d.app_id.get(), d.app_math.get(), d.app_kind.get(), d.app_base_i.get(),
                             d.app_dir.get(), d.app_status.get(), d.lg_phases.get(), d.lg_upd_buf.get(), d.lg_upd_pos.get(), d.lg_upd_id.get(),
                             d.lg_base_s.get(), d.lg_base_status.get(), d.lg_scale.get()};
    }

    // per-scenario source references (u_ref / u_ref_angle updates are allowed on this path)
    std::vector<double> uref;
    bool uref_shared = true;
    if (update.source.data != nullptr) {
        uref_shared = false;
        std::vector<std::vector<double>> sinj_unused(1), u(1);
        for (Idx s = 0; s != n_scn; ++s) {
            Saved saved;
            UpdateData only_source{};
            only_source.n_scenarios = update.n_scenarios;
            only_source.source = update.source;
            apply_scenario(only_source, s, &saved);
            size_t const off = u[0].size();
            u[0].resize(off + m.n_source() * 2);
            for (size_t i = 0; i != source_in_.size(); ++i) {
                Coupling const c = topo_.source[i];
                if (c.group != 0) continue;
                cplx const v = source_u_ref(source_st_[i]);
                u[0][off + c.pos * 2] = v.real();
                u[0][off + c.pos * 2 + 1] = v.imag();
            }
            restore(saved);
        }
        uref = std::move(u[0]);
    } else {
        uref.assign(m.n_source() * 2, 0.0);
        for (size_t i = 0; i != source_in_.size(); ++i) {
            Coupling const c = topo_.source[i];
            if (c.group != 0) continue;
            cplx const v = source_u_ref(source_st_[i]);
            uref[c.pos * 2] = v.real();
            uref[c.pos * 2 + 1] = v.imag();
        }
    }
    timing[0] += ms_since(t0);

    // ---- chunk pipeline: H2D of the raw update rows -> apply -> solve -> output structs -> D2H, one stream per chunk, so the
    //      PCIe transfers of one chunk overlap the kernels of the others (both copy engines + the SMs busy at once) ----
    t0 = Clock::now();
    e.set_method_hint(opt.method);
    e.stage_device(n_scn, uref.data(), uref_shared);
    if (e.has_regulators()) { // VoltageRegulator::calc_param of the permanent state (regulator updates take the host route)
        std::vector<double> rp(m.n_voltage_regulator() * 4, 0.0);
        for (size_t r = 0; r != reg_in_.size(); ++r) {
            Coupling const c = topo_.voltage_regulator[r];
            if (c.group != 0) continue;
            rp[c.pos * 4] = reg_st_[r].status ? 1.0 : 0.0;
            rp[c.pos * 4 + 1] = reg_st_[r].u_ref;
            rp[c.pos * 4 + 2] = reg_st_[r].q_min / kBasePower3p;
            rp[c.pos * 4 + 3] = reg_st_[r].q_max / kBasePower3p;
        }
        e.set_regulators(rp.data());
        d.reg_flags.ensure(static_cast<size_t>(n_scn) * m.n_voltage_regulator() * 2 + 1);
    }
    if (outage_plan_ != nullptr) { // branch-outage overlay of this part's scenarios (model.hpp: OutagePlan)
        OutagePlan const& plan = *outage_plan_;
        size_t const bb2 = static_cast<size_t>(phases) * phases * 2;
        size_t const K = static_cast<size_t>(plan.n_slot), f = static_cast<size_t>(first_scenario);
        e.set_overlay(n_scn, plan.math_branch.data() + f * K, plan.bparam.data() + f * K * 4 * bb2, plan.comp.data() + f * K,
                      plan.energized.data() + f * K, plan.dead_off.data() + f, plan.dead.data(), plan.dead.size(), plan.n_slot);
    }
    SolveOptions const sopt = e.prepare_solve({opt.method, opt.err_tol, static_cast<int32_t>(opt.max_iter)});
    d.flag.ensure(1);
    PGMB_CUDA(cudaMemsetAsync(d.flag.get(), 0, sizeof(int32_t), st));
    int const tw = e.tile_width();
    DevStructure const& ds = e.dev_structure();
    int const force_const_y = e.last_method() == 0 ? 1 : 0;
    int64_t const n_tile = e.dev_batch().n_tile;
    int n_chunk = static_cast<int>(std::min<int64_t>(DeviceSide::kMaxChunk, std::max<int64_t>(1, n_tile / 16)));
    if (resident_in && resident_out) n_chunk = 1; // no transfers to overlap: one launch per kernel
    if (char const* env = std::getenv("PGMB_CHUNKS")) n_chunk = std::max(1, std::min<int>(DeviceSide::kMaxChunk, std::atoi(env)));
    n_chunk = static_cast<int>(std::min<int64_t>(n_chunk, n_tile));

    ComponentBuffer const* ubufs[4] = {&update.sym_gen, &update.asym_gen, &update.sym_load, &update.asym_load};
    size_t const urow[4] = {sizeof(SymLoadGenUpdate), sizeof(AsymLoadGenUpdate), sizeof(SymLoadGenUpdate), sizeof(AsymLoadGenUpdate)};
    for (int bfr = 0; bfr != 4; ++bfr) {
        size_t const bytes = (ubufs[bfr]->data != nullptr && ubufs[bfr]->n != 0) ? static_cast<size_t>(n_scn) * ubufs[bfr]->n * urow[bfr] : 0;
        if (resident_in) {
            if (bytes != d.resident_rows[bfr]) throw InvalidArgument("PGMB_FLAG_RESIDENT_INPUT: the previous call did not upload this batch's update rows");
        } else {
            if (bytes != 0) d.upd[bfr].ensure(bytes);
            d.resident_rows[bfr] = bytes;
        }
    }
    struct Req {
        void* host;
        int slot;
        size_t row;
        Idx count;
    };
    Req const reqs[13] = {
        {out.node, 0, sym ? sizeof(NodeOutput<1>) : sizeof(NodeOutput<3>), nn},
        {out.line, 1, sym ? sizeof(BranchOutput<1>) : sizeof(BranchOutput<3>), n_line()},
        {out.transformer, 2, sym ? sizeof(BranchOutput<1>) : sizeof(BranchOutput<3>), n_trafo()},
        {out.shunt, 3, sym ? sizeof(ApplianceOutput<1>) : sizeof(ApplianceOutput<3>), static_cast<Idx>(shunt_in_.size())},
        {out.source, 4, sym ? sizeof(ApplianceOutput<1>) : sizeof(ApplianceOutput<3>), static_cast<Idx>(source_in_.size())},
        {out.sym_gen, 5, sym ? sizeof(ApplianceOutput<1>) : sizeof(ApplianceOutput<3>), n_sym_gen_},
        {out.asym_gen, 6, sym ? sizeof(ApplianceOutput<1>) : sizeof(ApplianceOutput<3>), n_asym_gen_},
        {out.sym_load, 7, sym ? sizeof(ApplianceOutput<1>) : sizeof(ApplianceOutput<3>), n_sym_load_},
        {out.asym_load, 8, sym ? sizeof(ApplianceOutput<1>) : sizeof(ApplianceOutput<3>), n_asym_load_},
        {out.voltage_regulator, 9, sizeof(VoltageRegulatorOutput), static_cast<Idx>(reg_in_.size())},
        {out.asym_line, 10, sym ? sizeof(BranchOutput<1>) : sizeof(BranchOutput<3>), n_aline()},
        {out.generic_branch, 11, sym ? sizeof(BranchOutput<1>) : sizeof(BranchOutput<3>), n_gb()},
        {out.link, 12, sym ? sizeof(BranchOutput<1>) : sizeof(BranchOutput<3>), n_link()},
    };
    for (Req const& r : reqs) {
        if (r.host != nullptr && r.count != 0) d.out[r.slot].ensure(static_cast<size_t>(n_scn) * r.count * r.row);
    }
    size_t const src_row = sym ? 4 : 12; // doubles per source result
    d.src_res.ensure(static_cast<size_t>(n_scn) * m.n_source() * src_row + 1);
    // The chunks only overlap when the caller's buffers are page-locked: a copy from / to pageable memory blocks the host
    // until the chunk's kernels are done, which would run the chunks one after another, each paying the solver's latency.
    // Pageable OUTPUT buffers (what a client of PGM_calculate that allocates with malloc / numpy hands over) are served through a
    // page-locked staging area owned by the model: every chunk's results go device -> staging asynchronously, and host threads
    // copy a finished chunk into the caller's memory (faulting its fresh pages in parallel) while the next chunks are still in
    // flight.  A direct cudaMemcpy into unfaulted pageable memory runs at a few GB/s on one driver thread.
    // Staging is decided per buffer: a page-locked buffer takes its results directly, a pageable one goes through the staging area
    // (a client may mix both: PGM_create_buffer page-locks large buffers only).
    size_t stage_off[13] = {};
    bool slot_staged[13] = {};
    size_t stage_bytes = 0;
    bool staged = false;
    {
        bool any_pageable = false;
        for (Req const& r : reqs) {
            if (r.host == nullptr || r.count == 0 || resident_out) continue;
            if (!is_device_accessible_host(r.host)) {
                any_pageable = true;
                slot_staged[r.slot] = true;
                stage_off[r.slot] = stage_bytes;
                stage_bytes += (static_cast<size_t>(n_scn) * r.count * r.row + 4095) / 4096 * 4096;
            }
        }
        staged = any_pageable && stage_bytes <= (size_t{3} << 29) && std::getenv("PGMB_NO_STAGING") == nullptr; // at most 1.5 GB page-locked
        if (!staged)
            for (bool& b : slot_staged) b = false;
        // a copy INTO pageable memory returns only when the chunk's kernels are done, which would run the chunks one after another;
        // a copy FROM pageable memory (update rows) returns once the driver has staged the rows, so it does not stop the overlap
        if (!staged && any_pageable && std::getenv("PGMB_CHUNKS") == nullptr) n_chunk = 1;
    }
    unsigned char* const stage = staged ? d.staging(stage_bytes) : nullptr;
    PGMB_CUDA(cudaEventRecord(d.ev_p0, st));
    PGMB_CUDA(cudaEventRecord(d.fork, st));

    try {
    for (int c = 0; c != n_chunk; ++c) {
        // one chunk (device-resident pipeline): everything stays on the engine stream, no fork / join between streams
        cudaStream_t const q = n_chunk == 1 ? st : d.cs[c];
        int64_t const tile_b = n_tile * c / n_chunk, tile_e = n_tile * (c + 1) / n_chunk;
        DevBatch const view = e.batch_view(tile_b, tile_e);
        int64_t const s0 = tile_b * tw, ns = view.n_scn;
        if (q != st) PGMB_CUDA(cudaStreamWaitEvent(q, d.fork, 0));
        DevUpdateBuffers ub{};
        ub.id_mismatch = d.flag.get();
        for (int bfr = 0; bfr != 4; ++bfr) {
            if (ubufs[bfr]->data == nullptr || ubufs[bfr]->n == 0) continue;
            size_t const per = static_cast<size_t>(ubufs[bfr]->n) * urow[bfr];
            if (!resident_in) {
                PGMB_CUDA(cudaMemcpyAsync(d.upd[bfr].get() + s0 * per, static_cast<unsigned char const*>(ubufs[bfr]->data) + s0 * per,
                                          ns * per, cudaMemcpyHostToDevice, q));
            }
            ub.data[bfr] = d.upd[bfr].get() + s0 * per;
            ub.n_per_scenario[bfr] = ubufs[bfr]->n;
        }
        if (sym) {
            launch_apply_load_update_sym(tw, ds, view, d.t, ub, q);
        } else {
            launch_apply_load_update_asym(tw, ds, view, d.t, ub, q);
        }
        PGMB_CUDA(cudaEventRecord(d.ev_a[c], q));
        e.launch_solve(view, sopt, q);
        PGMB_CUDA(cudaEventRecord(d.ev_b[c], q));
        int8_t* const reg_flags = e.has_regulators() ? d.reg_flags.get() + s0 * m.n_voltage_regulator() * 2 : nullptr;
        e.launch_regulator_apply(view, reg_flags, q); // before every kernel that reads the generators' power
        double* const src_res = d.src_res.get() + s0 * m.n_source() * src_row;
        if (sym) {
            launch_source_result_sym(tw, ds, view, force_const_y, src_res, q);
        } else {
            launch_source_result_asym(tw, ds, view, force_const_y, src_res, q);
        }
        for (Req const& r : reqs) {
            if (r.host == nullptr || r.count == 0) continue;
            void* const dst = d.out[r.slot].get() + static_cast<size_t>(s0) * r.count * r.row;
            int const nl = static_cast<int>(n_line()), nt = static_cast<int>(n_trafo());
            int const o_al = static_cast<int>(off_aline()), n_al = static_cast<int>(n_aline()), o_gb = static_cast<int>(off_gb()),
                      n_g = static_cast<int>(n_gb()), o_tr = static_cast<int>(off_trafo()), o_lk = static_cast<int>(off_link()),
                      n_lk = static_cast<int>(n_link());
            int const app_first = (r.slot >= 3 && r.slot < 9) ? static_cast<int>(d.n_app_first[r.slot - 3]) : 0;
            if (sym) {
                switch (r.slot) {
                case 0: launch_pack_node_sym(tw, ds, view, d.t, force_const_y, src_res, dst, q); break;
                case 1: launch_pack_branch_sym(tw, ds, view, d.t, 0, nl, dst, q); break;
                case 2: launch_pack_branch_sym(tw, ds, view, d.t, o_tr, nt, dst, q); break;
                case 10: launch_pack_branch_sym(tw, ds, view, d.t, o_al, n_al, dst, q); break;
                case 11: launch_pack_branch_sym(tw, ds, view, d.t, o_gb, n_g, dst, q); break;
                case 12: launch_pack_branch_sym(tw, ds, view, d.t, o_lk, n_lk, dst, q); break;
                case 9:
                    launch_pack_regulator(ns, static_cast<int>(r.count), static_cast<int>(m.n_voltage_regulator()), d.reg_id.get(),
                                          d.reg_math.get(), d.reg_status.get(), reg_flags, dst, q);
                    break;
                default: launch_pack_appliance_sym(tw, ds, view, d.t, force_const_y, app_first, static_cast<int>(r.count), src_res, dst, q);
                }
            } else {
                switch (r.slot) {
                case 0: launch_pack_node_asym(tw, ds, view, d.t, force_const_y, src_res, dst, q); break;
                case 1: launch_pack_branch_asym(tw, ds, view, d.t, 0, nl, dst, q); break;
                case 2: launch_pack_branch_asym(tw, ds, view, d.t, o_tr, nt, dst, q); break;
                case 10: launch_pack_branch_asym(tw, ds, view, d.t, o_al, n_al, dst, q); break;
                case 11: launch_pack_branch_asym(tw, ds, view, d.t, o_gb, n_g, dst, q); break;
                case 12: launch_pack_branch_asym(tw, ds, view, d.t, o_lk, n_lk, dst, q); break;
                case 9:
                    launch_pack_regulator(ns, static_cast<int>(r.count), static_cast<int>(m.n_voltage_regulator()), d.reg_id.get(),
                                          d.reg_math.get(), d.reg_status.get(), reg_flags, dst, q);
                    break;
                default: launch_pack_appliance_asym(tw, ds, view, d.t, force_const_y, app_first, static_cast<int>(r.count), src_res, dst, q);
                }
            }
        }
        PGMB_CUDA(cudaGetLastError());
        for (Req const& r : reqs) {
            if (r.host == nullptr || r.count == 0 || resident_out) continue;
            size_t const blk = static_cast<size_t>(r.count) * r.row, off = static_cast<size_t>(s0) * blk;
            if (out_scatter_ != nullptr && !slot_staged[r.slot]) { // every scenario's rows to its own place in the caller's batch
                for (int64_t i = 0; i != ns; ++i) {
                    PGMB_CUDA(cudaMemcpyAsync(static_cast<unsigned char*>(r.host) + static_cast<size_t>(out_scatter_[first_scenario + s0 + i]) * blk,
                                              d.out[r.slot].get() + off + static_cast<size_t>(i) * blk, blk, cudaMemcpyDeviceToHost, q));
                }
                continue;
            }
            unsigned char* const dst = slot_staged[r.slot] ? stage + stage_off[r.slot] + off : static_cast<unsigned char*>(r.host) + off;
            PGMB_CUDA(cudaMemcpyAsync(dst, d.out[r.slot].get() + off, static_cast<size_t>(ns) * r.count * r.row,
                                      cudaMemcpyDeviceToHost, q));
        }
        if (q != st) {
            PGMB_CUDA(cudaEventRecord(d.ev_end[c], q));
            PGMB_CUDA(cudaStreamWaitEvent(st, d.ev_end[c], 0)); // join: the engine stream sees the end of every chunk
        }
    }
    PGMB_CUDA(cudaEventRecord(d.ev_p1, st));
    timing[1] += ms_since(t0);

    t0 = Clock::now();
    if (staged) {
        char const* const env_thr = std::getenv("PGMB_COPY_THREADS");
        unsigned const n_copy = env_thr != nullptr ? static_cast<unsigned>(std::max(1, std::min(64, std::atoi(env_thr))))
                                                   : std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        for (int c = 0; c != n_chunk; ++c) {
            PGMB_CUDA(cudaStreamSynchronize(n_chunk == 1 ? st : d.cs[c]));
            int64_t const tile_b = n_tile * c / n_chunk, tile_e = n_tile * (c + 1) / n_chunk;
            int64_t const s0 = tile_b * tw, ns = std::min<int64_t>(tile_e * tw, n_scn) - s0;
            if (ns <= 0) continue;
            struct Piece {
                unsigned char* dst;
                unsigned char const* src;
                size_t bytes;
            };
            std::vector<Piece> pieces;
            size_t total = 0;
            for (Req const& r : reqs) {
                if (r.host == nullptr || r.count == 0 || !slot_staged[r.slot]) continue;
                size_t const off = static_cast<size_t>(s0) * r.count * r.row, bytes = static_cast<size_t>(ns) * r.count * r.row;
                if (out_scatter_ != nullptr) {
                    size_t const blk = static_cast<size_t>(r.count) * r.row;
                    for (int64_t i = 0; i != ns; ++i) {
                        pieces.push_back({static_cast<unsigned char*>(r.host) + static_cast<size_t>(out_scatter_[first_scenario + s0 + i]) * blk,
                                          stage + stage_off[r.slot] + off + static_cast<size_t>(i) * blk, blk});
                    }
                } else {
                    pieces.push_back({static_cast<unsigned char*>(r.host) + off, stage + stage_off[r.slot] + off, bytes});
                }
                total += bytes;
            }
            unsigned const n_thr = total < (size_t{4} << 20) ? 1u : n_copy;
            auto copy_share = [&pieces, n_thr](unsigned t) { // thread t takes the t-th 1/n_thr of every piece, cut at 4 KB
                for (Piece const& p : pieces) {
                    size_t const step = ((p.bytes + n_thr - 1) / n_thr + 4095) / 4096 * 4096; // n_thr * step >= bytes
                    size_t const b = std::min(p.bytes, step * t), e2 = std::min(p.bytes, step * (t + 1));
                    if (e2 > b) std::memcpy(p.dst + b, p.src + b, e2 - b);
                }
            };
            if (n_thr == 1) {
                copy_share(0);
            } else {
                std::vector<std::thread> pool;
                for (unsigned t = 1; t != n_thr; ++t) pool.emplace_back(copy_share, t);
                copy_share(0);
                for (auto& th : pool) th.join();
            }
        }
    }
    for (int c = 0; c != n_chunk; ++c) PGMB_CUDA(cudaStreamSynchronize(n_chunk == 1 ? st : d.cs[c]));
    } catch (...) {
        d.drain(); // copies into caller / staging memory may still be in flight
        throw;
    }
    for (int c = 0; c != n_chunk; ++c) {
        float ms = 0.0f;
        PGMB_CUDA(cudaEventElapsedTime(&ms, d.ev_a[c], d.ev_b[c]));
        timing[2] += ms; // solver time per chunk; chunks overlap, so the sum can exceed the wall time
    }
    {
        PGMB_CUDA(cudaEventSynchronize(d.ev_p1));
        float ms = 0.0f;
        PGMB_CUDA(cudaEventElapsedTime(&ms, d.ev_p0, d.ev_p1));
        timing[6] += ms; // device time of the whole pipeline: (H2D) -> apply -> solve -> results / output structs -> (D2H)
    }
    int32_t mismatch = 0;
    PGMB_CUDA(cudaMemcpyAsync(&mismatch, d.flag.get(), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    std::vector<int32_t> st_local(n_scn), it_local(n_scn);
    e.fetch_status(st_local.data(), it_local.data());
    PGMB_CUDA(cudaStreamSynchronize(st));
    timing[4] += ms_since(t0);
    // ids must repeat in every scenario (independent batch, main_core/update.hpp:58-83); checked by the apply kernel on the
    // rows it reads.  A dependent batch is recomputed by the caller on the per-scenario host path.
    if (mismatch != 0) return -1;

    int64_t failed = 0;
    std::vector<double> dev_local;
    for (Idx s = 0; s != n_scn; ++s) {
        if (n_iter != nullptr) n_iter[s] = it_local[s];
        if (status != nullptr) status[s] = st_local[s];
        if (st_local[s] != 0) {
            ++failed;
            if (dev_local.empty()) { // fetched only when something failed
                dev_local.resize(n_scn);
                PGMB_CUDA(cudaMemcpy(dev_local.data(), e.dev_batch().max_dev, sizeof(double) * n_scn, cudaMemcpyDeviceToHost));
            }
            batch_message += "Error in batch #" + std::to_string(first_scenario + s) + ": " +
                             scenario_failure_text(st_local[s], opt.max_iter, dev_local[s], opt.err_tol) + "\n";
        }
    }
    return failed;
}

} // namespace pgmb
