// Automatic tap changer: the outer loop the reference runs around its power flow when PGM_set_tap_changing_strategy is not
// "disabled" (main_model_impl.hpp:318-348 -> optimizer/tap_position_optimizer.hpp).  Host logic; every evaluation of a tap
// setting is a power flow on the GPU engines of the model (Model::solve_block), with the transformer parameters re-uploaded and
// the symbolic structures kept (a tap change is a parameter change).
//
//   ranking        tap_position_optimizer.hpp:143-433   directed graph over the nodes (regulated transformers: one edge of
//                                                        weight 1 towards the control side; everything else: weight 0 both
//                                                        ways), shortest distance from the sources, rank = distance of the
//                                                        control side
//   search         :442-483, :785-907, :1016-1420       scan (one tap per power flow) or bisection of the tap range, rank by
//                                                        rank, restart of the lower ranks after a change; min / max strategies
//                                                        start from the far end, search, step one further and scan back
//   measurement    :628-731                             |u + z_comp * i| at the control side against u_set +- u_band / 2
//
// The reference asks its caller for a batch scenario by scenario as well (job_dispatch.hpp): a batch with a tap strategy takes
// the model's scenario-by-scenario route, host threads sharing the GPU.
#include "model.hpp"

#include <algorithm>
#include <limits>
#include <numeric>
#include <queue>

namespace pgmb {

namespace {
using Clock = std::chrono::steady_clock;

constexpr int64_t kUnreachable = std::numeric_limits<int64_t>::max();
constexpr int64_t kLastRank = kUnreachable - 1;

struct SolveFailure : std::runtime_error { // IterationDiverge / SparseMatrixError of one power flow
    using std::runtime_error::runtime_error;
};
struct TapSearchFailure : std::runtime_error { // MaxIterationReached of the optimizer (an IterationDiverge in the reference)
    using std::runtime_error::runtime_error;
};

enum class Search { scan, bisect };
enum class Goal { any, fast_any, highest, lowest };

// tap range of one transformer seen as "which way does the voltage at the control side go"
struct TapRange {
    IntS tap_min, tap_max;
    IntS up(IntS pos) const { // one step towards tap_max
        if (pos == tap_max) return pos;
        return static_cast<IntS>(tap_min < tap_max ? pos + 1 : pos - 1);
    }
    IntS down(IntS pos) const { // one step towards tap_min
        if (pos == tap_min) return pos;
        return static_cast<IntS>(tap_min < tap_max ? pos - 1 : pos + 1);
    }
    // a higher tap position is a higher voltage on the tap side, i.e. a lower one across the transformer
    IntS voltage_up(IntS pos, bool control_at_tap_side) const { return control_at_tap_side ? up(pos) : down(pos); }
    IntS voltage_down(IntS pos, bool control_at_tap_side) const { return control_at_tap_side ? down(pos) : up(pos); }
    IntS highest_voltage(bool control_at_tap_side) const { return control_at_tap_side ? tap_max : tap_min; }
    IntS lowest_voltage(bool control_at_tap_side) const { return control_at_tap_side ? tap_min : tap_max; }
    int64_t width() const { return std::abs(static_cast<int64_t>(tap_max) - static_cast<int64_t>(tap_min)); }
};

// bisection state of one transformer (tap_position_optimizer.hpp:785-907); lo / hi are tap positions in numerical order
class Bisection {
  public:
    void reset(IntS pos, TapRange const& range, bool control_at_tap_side) {
        went_down_ = final_check_ = settled_ = false;
        pos_ = pos;
        lo_ = std::min(range.tap_min, range.tap_max);
        hi_ = std::max(range.tap_min, range.tap_max);
        reversed_ = range.tap_max < range.tap_min;
        at_tap_side_ = control_at_tap_side;
    }
    IntS pos() const { return pos_; }
    bool went_down() const { return went_down_; }
    bool settled() const { return settled_; }
    bool exhausted() const { return lo_ >= hi_; }
    void set_pos(IntS pos) { pos_ = pos; }
    void clear_flags() { final_check_ = settled_ = false; }

    // the voltage is outside the band: halve towards the side that brings it back
    void step_towards_band(bool highest, bool above_band) {
        bool const down = (above_band == reversed_) != at_tap_side_;
        if (final_check_) {
            pos_ = down ? lo_ : hi_;
            settled_ = true;
            return;
        }
        went_down_ = down;
        (went_down_ ? hi_ : lo_) = pos_;
        if (lo_ < hi_) pos_ = middle(highest != reversed_);
    }
    // the voltage is inside the band: keep this position as a bound and look for a better one on the preferred side
    void keep_as_bound(bool highest) {
        bool const invert = at_tap_side_ != highest;
        if (reversed_ == invert) {
            lo_ = pos_;
            went_down_ = false;
        } else {
            hi_ = pos_;
            went_down_ = true;
        }
    }
    IntS next_candidate(bool highest, bool previous_down, bool& changed) {
        IntS const candidate = middle((highest != reversed_) != at_tap_side_);
        int const diff = candidate - pos_;
        if (diff == 0) {
            changed = !settled_;
            settled_ = true;
            return candidate;
        }
        if ((diff == 1 && previous_down) || (diff == -1 && !previous_down)) final_check_ = true;
        changed = true;
        pos_ = candidate;
        return candidate;
    }

  private:
    // std::midpoint(a, b) rounds towards a: the first argument is the bound the preference leans to
    IntS middle(bool prefer_higher) const {
        bool const towards_hi = at_tap_side_ != prefer_higher;
        int const a = towards_hi ? hi_ : lo_, b = towards_hi ? lo_ : hi_;
        return static_cast<IntS>(a + (b - a) / 2);
    }
    IntS lo_{}, hi_{}, pos_{};
    bool went_down_{}, final_check_{}, reversed_{}, settled_{}, at_tap_side_{};
};
} // namespace

struct Model::TapRanked {
    Idx regulator; // index into tap_reg_in_
    int kind;      // 0 transformer, 1 three-winding transformer
    Idx index;     // within the kind
    bool control_at_tap_side;
    TapRange range;
};

// tap_position_optimizer.hpp:143-433
std::vector<std::vector<Model::TapRanked>> Model::rank_tap_regulators() const {
    struct Edge {
        Idx from, to;
        int64_t weight;
        int kind; // -1: not regulated
        Idx index;
        ID id;
    };
    std::vector<Edge> edges;
    // regulated transformers of each kind -> control side (only regulators that are switched on count)
    std::unordered_map<Idx, IntS> regulated[2];
    for (size_t r = 0; r != tap_reg_in_.size(); ++r) {
        if (!tap_reg_st_[r].status) continue;
        regulated[tap_reg_target_[r].kind].emplace(tap_reg_target_[r].index, tap_reg_in_[r].control_side);
    }
    auto both_ways = [&](Idx a, Idx b) {
        edges.push_back({a, b, 0, -1, -1, kNaID});
        edges.push_back({b, a, 0, -1, -1, kNaID});
    };
    for (Idx i = 0; i != n_trafo(); ++i) {
        BranchState const& st = branch_st_[off_trafo() + i];
        if (!st.from_status || !st.to_status) continue;
        Idx const from = node_seq(trafo_in_[i].from_node), to = node_seq(trafo_in_[i].to_node);
        if (auto it = regulated[0].find(i); it != regulated[0].end()) {
            bool const control_from = it->second == 0;
            edges.push_back({control_from ? to : from, control_from ? from : to, 1, 0, i, trafo_in_[i].id});
        } else {
            both_ways(from, to);
        }
    }
    for (Idx i = 0; i != n_t3w(); ++i) {
        ThreeWindingTransformerInput const& t = t3w_c_[i].in;
        Idx const node[3] = {node_seq(t.node_1), node_seq(t.node_2), node_seq(t.node_3)};
        auto const it = regulated[1].find(i);
        bool const is_regulated = it != regulated[1].end();
        int const pairs[3][2] = {{0, 1}, {0, 2}, {1, 2}};
        for (auto const& pr : pairs) {
            int const first = pr[0], second = pr[1];
            if (!t3w_st_[i].status[first] || !t3w_st_[i].status[second]) continue;
            bool const tap_at_first = t.tap_side == first;
            if (is_regulated && (tap_at_first || t.tap_side == second)) {
                bool const tap_at_control = it->second == t.tap_side;
                Idx const tap_node = tap_at_first ? node[first] : node[second];
                Idx const other_node = tap_at_first ? node[second] : node[first];
                edges.push_back({tap_at_control ? other_node : tap_node, tap_at_control ? tap_node : other_node, 1, 1, i, t.id});
            } else {
                both_ways(node[first], node[second]);
            }
        }
    }
    for (Idx i = 0; i != n_line(); ++i) {
        if (branch_st_[i].from_status && branch_st_[i].to_status) both_ways(node_seq(line_in_[i].from_node), node_seq(line_in_[i].to_node));
    }
    for (Idx i = 0; i != n_link(); ++i) {
        BranchState const& st = branch_st_[off_link() + i];
        if (st.from_status && st.to_status) both_ways(node_seq(link_in_[i].from_node), node_seq(link_in_[i].to_node));
    }
    // edges grouped by their start node, input order kept inside a node (the reference's compressed-sparse-row graph)
    std::stable_sort(edges.begin(), edges.end(), [](Edge const& a, Edge const& b) { return a.from < b.from; });
    Idx const n = static_cast<Idx>(node_.size());
    std::vector<Idx> first_edge(n + 1, 0);
    for (Edge const& e : edges) ++first_edge[e.from + 1];
    for (Idx v = 0; v != n; ++v) first_edge[v + 1] += first_edge[v];
    std::vector<char> is_source(n, 0);
    for (size_t i = 0; i != source_in_.size(); ++i) is_source[node_seq(source_in_[i].node)] = source_st_[i].status ? 1 : 0;
    std::vector<int64_t> dist(n, kUnreachable);
    for (Idx v = 0; v != n; ++v) {
        if (!is_source[v]) continue;
        using Item = std::pair<int64_t, Idx>;
        std::priority_queue<Item, std::vector<Item>, std::greater<>> queue;
        dist[v] = 0;
        queue.emplace(0, v);
        while (!queue.empty()) {
            auto const [d, u] = queue.top();
            queue.pop();
            if (d != dist[u]) continue;
            for (Idx k = first_edge[u]; k != first_edge[u + 1]; ++k) {
                Edge const& e = edges[k];
                if (dist[u] + e.weight < dist[e.to]) {
                    dist[e.to] = dist[u] + e.weight;
                    queue.emplace(dist[e.to], e.to);
                }
            }
        }
    }
    struct Weighted {
        int64_t rank;
        int kind;
        Idx index;
    };
    std::vector<Weighted> weighted;
    std::vector<ID> wrong_way;
    for (Edge const& e : edges) {
        if (e.kind < 0) continue;
        int64_t const from_rank = dist[e.from], to_rank = dist[e.to];
        if (from_rank == kUnreachable && to_rank == kUnreachable) continue; // not energized
        if (from_rank == kUnreachable || to_rank == kUnreachable) {
            wrong_way.push_back(e.id);
        } else if (from_rank != to_rank - 1) {
            weighted.push_back({kLastRank, e.kind, e.index}); // the control side is already held by a closer transformer
        } else {
            weighted.push_back({to_rank, e.kind, e.index});
        }
    }
    if (!wrong_way.empty()) {
        std::sort(wrong_way.begin(), wrong_way.end());
        wrong_way.erase(std::unique(wrong_way.begin(), wrong_way.end()), wrong_way.end());
        std::string msg = "Automatic tap changer has invalid configuration. The following transformer(s) are being controlled from "
                          "non-source side towards source side:\n  Transformer IDs: ";
        for (size_t i = 0; i != wrong_way.size(); ++i) msg += (i != 0 ? ", " : "") + std::to_string(wrong_way[i]);
        throw InvalidArgument(msg);
    }
    std::stable_sort(weighted.begin(), weighted.end(), [](Weighted const& a, Weighted const& b) { return a.rank < b.rank; });
    std::vector<std::vector<TapRanked>> groups;
    int64_t previous = std::numeric_limits<int64_t>::lowest();
    for (Weighted const& w : weighted) {
        if (w.rank > previous) {
            groups.emplace_back();
            previous = w.rank;
        }
        auto& group = groups.back();
        if (std::any_of(group.begin(), group.end(), [&](TapRanked const& t) { return t.kind == w.kind && t.index == w.index; })) continue;
        Idx regulator = -1;
        for (size_t r = 0; r != tap_reg_in_.size() && regulator < 0; ++r) {
            if (tap_reg_target_[r].kind == w.kind && tap_reg_target_[r].index == w.index) regulator = static_cast<Idx>(r);
        }
        IntS const tap_side = w.kind == 0 ? trafo_in_[w.index].tap_side : t3w_c_[w.index].in.tap_side;
        TapRange const range = w.kind == 0 ? TapRange{trafo_in_[w.index].tap_min, trafo_in_[w.index].tap_max}
                                           : TapRange{t3w_c_[w.index].in.tap_min, t3w_c_[w.index].in.tap_max};
        group.push_back({regulator, w.kind, w.index, tap_reg_in_[regulator].control_side == tap_side, range});
    }
    return groups;
}

std::vector<int64_t> Model::tap_rank_table() const {
    std::vector<int64_t> out;
    auto const groups = rank_tap_regulators();
    for (size_t g = 0; g != groups.size(); ++g) {
        for (auto const& t : groups[g]) {
            out.push_back(t.kind);
            out.push_back(t.index);
            out.push_back(static_cast<int64_t>(g));
        }
    }
    return out;
}

template <int B>
int64_t Model::run_tap_optimizer(ModelOptions const& opt, OutputData const& out, Idx scenario, int32_t* n_iter, int32_t* status) {
    constexpr int c2 = 2 * B;
    Goal const goal = opt.tap_strategy == 2 ? Goal::lowest : opt.tap_strategy == 3 ? Goal::highest
                      : opt.tap_strategy == 4 ? Goal::fast_any : Goal::any;
    // main_model_impl.hpp:337-339: the scan for "any", bisection for everything else
    Search const first_search = goal == Goal::any ? Search::scan : Search::bisect;
    bool const highest = goal == Goal::highest;

    prepare_topology();
    std::vector<std::vector<TapRanked>> const order = rank_tap_regulators();
    auto tap_of = [&](TapRanked const& t) -> IntS& { return t.kind == 0 ? trafo_st_[t.index].tap_pos : t3w_st_[t.index].tap_pos; };
    auto set_tap = [&](TapRanked const& t, IntS pos) { // Transformer::set_tap: limited to the tap range
        IntS const limited = t.kind == 0 ? tap_limit(trafo_c_[t.index], pos) : tap_limit(t3w_c_[t.index], pos);
        if (limited != tap_of(t)) {
            tap_of(t) = limited;
            mark(false, true, nullptr);
        }
    };
    // the tap positions the model holds now come back when the search is over, whatever its outcome
    std::vector<std::pair<TapRanked const*, IntS>> before;
    for (auto const& group : order)
        for (auto const& t : group) before.emplace_back(&t, tap_of(t));
    auto put_back = [&] {
        for (auto const& [t, pos] : before) set_tap(*t, pos);
    };

    BlockSolution sol;
    RegulatorInput reg;
    int32_t total_nr_iterations = 0;
    auto power_flow = [&](int32_t method) {
        ModelOptions o = opt;
        o.method = method;
        check_regulators<B>(o);
        prepare_engines<B>();
        std::vector<std::vector<double>> sinj(topo_.math.size()), uref(topo_.math.size());
        gather_pf_input<B>(sinj, uref, &reg);
        solve_block<B>(o, 1, sinj, uref, &reg, sol);
        total_nr_iterations = std::max(total_nr_iterations, sol.n_iter[0]);
        if (sol.status[0] != 0) throw SolveFailure(scenario_failure_text(sol.status[0], o.max_iter, sol.max_dev[0], o.err_tol));
    };
    // NodeState <=> TransformerTapRegulatorCalcParam (:702-731): -1 below the band, 0 inside, +1 above; false: control side dead
    auto measure = [&](TapRanked const& t, int& cmp) {
        TapRegulatorState const& st = tap_reg_st_[t.regulator];
        IntS const side = tap_reg_in_[t.regulator].control_side;
        Idx node;
        Coupling branch{-1, -1};
        bool from_end = true;
        if (t.kind == 0) {
            node = node_seq(side == 0 ? trafo_in_[t.index].from_node : trafo_in_[t.index].to_node);
            branch = topo_.branch[off_trafo() + t.index];
            from_end = side == 0;
        } else {
            ThreeWindingTransformerInput const& in = t3w_c_[t.index].in;
            node = node_seq(side == 0 ? in.node_1 : side == 1 ? in.node_2 : in.node_3);
            Coupling3 const& c3 = topo_.branch3[t.index];
            branch = {c3.group, c3.group == -1 ? -1 : c3.pos[side]};
        }
        Coupling const bus = topo_.node[node];
        if (bus.group == -1) return false;
        double const u_rated = tap_reg_target_[t.regulator].u_rated;
        double const z_base = u_rated * u_rated / (B == 1 ? kBasePower3p : kBasePower1p);
        cplx const z{std::isnan(st.line_drop_compensation_r) ? 0.0 : st.line_drop_compensation_r,
                     std::isnan(st.line_drop_compensation_x) ? 0.0 : st.line_drop_compensation_x};
        cplx const z_comp = z / z_base;
        double const* u = &sol.so[0][bus.group][bus.pos * c2];
        double const* flow = branch.group == -1 ? nullptr : &sol.so[2][branch.group][branch.pos * 4 * c2];
        double const* i = flow == nullptr ? nullptr : flow + (from_end ? 2 : 3) * c2;
        double v = 0.0;
        for (int p = 0; p != B; ++p) {
            cplx const up{u[2 * p], u[2 * p + 1]};
            cplx const ip = i == nullptr ? cplx{} : cplx{i[2 * p], i[2 * p + 1]};
            v += std::abs(up + z_comp * ip);
        }
        v /= B;
        double const u_set = st.u_set / u_rated, u_band = st.u_band / u_rated;
        double const lower = u_set - 0.5 * u_band, upper = u_set + 0.5 * u_band;
        cmp = v < lower ? -1 : v > upper ? 1 : 0;
        if (!(v >= lower) && !(v < lower)) cmp = 0; // unordered (NaN) reads as "equivalent" in no branch of the reference; keep 0
        return true;
    };

    std::vector<std::vector<Bisection>> bisect(order.size());
    for (size_t g = 0; g != order.size(); ++g) {
        bisect[g].resize(order[g].size());
        for (size_t k = 0; k != order[g].size(); ++k) bisect[g][k].reset(tap_of(order[g][k]), order[g][k].range, order[g][k].control_at_tap_side);
    }
    std::vector<uint64_t> widest(order.size(), 0);
    for (size_t g = 0; g != order.size(); ++g)
        for (auto const& t : order[g]) widest[g] = std::max<uint64_t>(widest[g], static_cast<uint64_t>(t.range.width()));

    // one transformer, one power-flow result: the next tap position to try (true: it changed).  New positions are collected
    // and take effect when the pass over the ranks is over (the reference's update buffer): a pass that throws changes nothing
    std::vector<std::pair<TapRanked const*, IntS>> pending;
    auto adjust_scan = [&](TapRanked const& t) {
        int cmp = 0;
        if (!measure(t, cmp)) return false;
        IntS const now = tap_of(t);
        IntS const next = cmp > 0 ? t.range.voltage_down(now, t.control_at_tap_side)
                          : cmp < 0 ? t.range.voltage_up(now, t.control_at_tap_side) : now;
        if (next == now) return false;
        pending.emplace_back(&t, next);
        return true;
    };
    auto adjust_bisect = [&](TapRanked const& t, Bisection& bs, bool& changed) {
        int cmp = 0;
        if (!measure(t, cmp)) return;
        if (bs.exhausted() || bs.settled()) return;
        if (cmp != 0) bs.step_towards_band(highest, cmp > 0);
        if (IntS const proposed = bs.pos(); proposed != tap_of(t)) {
            bs.set_pos(proposed);
            pending.emplace_back(&t, proposed);
            changed = true;
            return;
        }
        if (goal == Goal::fast_any && cmp == 0) {
            changed = false;
            return;
        }
        bool const previous_down = bs.went_down();
        bs.keep_as_bound(highest);
        IntS const candidate = bs.next_candidate(highest, previous_down, changed);
        if (candidate == tap_of(t) && cmp != 0 && !bs.exhausted()) {
            bs.reset(candidate, t.range, t.control_at_tap_side);
            throw TapSearchFailure("Maximum number of iterations reached! TapPositionOptimizer::binary_search: no valid tap position found "
                                   "between tap " + std::to_string(static_cast<int>(t.range.tap_min)) + " and tap " +
                                   std::to_string(static_cast<int>(t.range.tap_max)) + "\n");
        }
        pending.emplace_back(&t, candidate);
    };
    // iterate (:1048-1105): ranks in order; the first rank that changes a tap ends the pass, the power flow is repeated and the
    // pass starts again from the first rank
    auto iterate = [&](int32_t method, Search search) {
        power_flow(method);
        std::vector<uint64_t> passes(order.size(), 0);
        bool changed = true;
        while (changed) {
            changed = false;
            pending.clear();
            size_t rank = 0;
            for (; rank != order.size(); ++rank) {
                for (size_t k = 0; k != order[rank].size(); ++k) {
                    if (search == Search::scan) {
                        changed = adjust_scan(order[rank][k]) || changed;
                    } else {
                        bool mine = false;
                        adjust_bisect(order[rank][k], bisect[rank][k], mine);
                        changed = mine || changed;
                    }
                }
                if (changed) {
                    std::fill(passes.begin() + static_cast<std::ptrdiff_t>(rank) + 1, passes.end(), 0);
                    ++passes[rank];
                    break;
                }
            }
            if (changed) {
                if (passes[rank] > 2 * widest[rank]) {
                    throw TapSearchFailure("Maximum number of iterations reached! TapPositionOptimizer::iterate " + std::to_string(passes[rank]) +
                                           " iterations reached: " + std::to_string(widest[rank]) + "x2 iterations in rank " +
                                           std::to_string(rank) + "\n");
                }
                for (auto const& [t, pos] : pending) set_tap(*t, pos);
                power_flow(method);
            }
        }
    };
    // iterate_with_fallback (:1029-1046): a power flow that does not converge gets one linear pass to move the taps first
    auto iterate_with_fallback = [&](int32_t method, Search search) {
        try {
            iterate(method, search);
        } catch (SolveFailure const&) {
            iterate(0, search);
            iterate(method, search);
        } catch (TapSearchFailure const&) { // MaxIterationReached is an IterationDiverge in the reference: same second attempt
            iterate(0, search);
            iterate(method, search);
        }
    };

    int64_t failed = 0;
    try {
        // pilot run (:1240-1285): min / max strategies start from the end of the range with the highest / lowest voltage
        if (goal == Goal::highest || goal == Goal::lowest) {
            for (auto const& group : order)
                for (auto const& t : group)
                    set_tap(t, highest ? t.range.highest_voltage(t.control_at_tap_side) : t.range.lowest_voltage(t.control_at_tap_side));
        }
        if (first_search == Search::bisect) {
            for (size_t g = 0; g != order.size(); ++g)
                for (size_t k = 0; k != order[g].size(); ++k) {
                    bisect[g][k].set_pos(tap_of(order[g][k]));
                    bisect[g][k].clear_flags();
                }
        }
        iterate_with_fallback(opt.method, first_search);
        if (goal == Goal::highest || goal == Goal::lowest) {
            // exploit_neighborhood (:1287-1317): one step past the found position, then scan back into the band
            for (auto const& group : order)
                for (auto const& t : group)
                    set_tap(t, highest ? t.range.voltage_up(tap_of(t), t.control_at_tap_side) : t.range.voltage_down(tap_of(t), t.control_at_tap_side));
            iterate_with_fallback(opt.method, Search::scan);
        }
        tap_positions_out_.assign(tap_reg_in_.size(), kNaIntS);
        for (auto const& group : order)
            for (auto const& t : group) tap_positions_out_[t.regulator] = tap_of(t);
        auto const t0 = Clock::now();
        write_output<B>(1, scenario, out, sol.so, sol.reg_out, &reg.lg_status);
        timing[3] += std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
        tap_positions_out_.clear();
        if (n_iter != nullptr) n_iter[scenario] = total_nr_iterations;
        if (status != nullptr) status[scenario] = 0;
    } catch (CudaError const&) {
        tap_positions_out_.clear();
        put_back();
        throw;
    } catch (std::exception const& ex) {
        tap_positions_out_.clear();
        put_back();
        if (status != nullptr) status[scenario] = 3;
        std::string what = ex.what();
        if (what.empty() || what.back() != '\n') what += "\n";
        batch_message += "Error in batch #" + std::to_string(scenario) + ": " + what;
        failed = 1;
    }
    put_back();
    return failed;
}

// ---- lockstep over a batch ------------------------------------------------------------------------------------------------------
// Load-profile batches on a grid with ONE regulated two-winding transformer (the reference benchmark's tap changer,
// fictional_grid_generator.hpp:615-633): every scenario keeps its own tap position and search state, and each step of the search
// is ONE batched power flow of all scenarios on the device pipeline, the transformer's admittances given per scenario through the
// branch overlay of the N-1 route (Engine::set_overlay, model.hpp: OutagePlan).  Intermediate passes bring back node and
// transformer results only; a last pass writes the caller's outputs.  Scenarios that need the reference's fallback (a power flow
// that does not converge, the iteration limit) are returned in `exact` for the scenario-by-scenario search.
// Returns the number of failed scenarios of the last pass, or -1 when the batch does not have this shape.
template <int B>
int64_t Model::run_tap_lockstep(ModelOptions const& opt, UpdateData const& update, OutputData const& out, int32_t* n_iter, int32_t* status,
                                std::vector<Idx>& exact) {
    constexpr size_t bb2 = static_cast<size_t>(B) * B * 2;
    if (opt.method != 1 && opt.method != -128) return -1;
    if (!reg_in_.empty() || update.n_scenarios <= 0 || status == nullptr) return -1;
    if (update.line.data != nullptr || update.transformer.data != nullptr || update.shunt.data != nullptr || update.source.data != nullptr ||
        update.voltage_regulator.data != nullptr || update.asym_line.data != nullptr || update.generic_branch.data != nullptr ||
        update.link.data != nullptr || update.three_winding_transformer.data != nullptr || update.transformer_tap_regulator.data != nullptr) {
        return -1;
    }
    prepare_engines<B>();
    if (!device_path_eligible(update)) return -1;
    MathTopology const& m = topo_.math[0];
    if (std::all_of(m.load_gen_type.begin(), m.load_gen_type.end(), [](int8_t t) { return t == 1; })) return -1; // linear method
    std::vector<std::vector<TapRanked>> const order = rank_tap_regulators();
    if (order.size() != 1 || order[0].size() != 1 || order[0][0].kind != 0) return -1;
    TapRanked const& t = order[0][0];
    Idx const ti = t.index, seq = off_trafo() + ti;
    Coupling const branch = topo_.branch[seq];
    IntS const side = tap_reg_in_[t.regulator].control_side;
    Idx const control_node = node_seq(side == 0 ? trafo_in_[ti].from_node : trafo_in_[ti].to_node);
    if (branch.group != 0 || topo_.node[control_node].group != 0) return -1;

    Goal const goal = opt.tap_strategy == 2 ? Goal::lowest : opt.tap_strategy == 3 ? Goal::highest
                      : opt.tap_strategy == 4 ? Goal::fast_any : Goal::any;
    Search const first_search = goal == Goal::any ? Search::scan : Search::bisect;
    bool const highest = goal == Goal::highest, two_stage = goal == Goal::highest || goal == Goal::lowest;
    Idx const n = update.n_scenarios, nn = static_cast<Idx>(node_.size());
    uint64_t const width = static_cast<uint64_t>(t.range.width());
    TapRegulatorState const& rst = tap_reg_st_[t.regulator];
    double const u_rated = tap_reg_target_[t.regulator].u_rated;
    double const z_base = u_rated * u_rated / (B == 1 ? kBasePower3p : kBasePower1p);
    cplx const z_comp = cplx{std::isnan(rst.line_drop_compensation_r) ? 0.0 : rst.line_drop_compensation_r,
                             std::isnan(rst.line_drop_compensation_x) ? 0.0 : rst.line_drop_compensation_x} / z_base;
    double const lower = rst.u_set / u_rated - 0.5 * rst.u_band / u_rated, upper = rst.u_set / u_rated + 0.5 * rst.u_band / u_rated;

    struct Lane {
        IntS tap;
        Bisection bs;
        uint64_t passes{0};
        int stage{0}; // 0: first search, 1: scan after the step past the found position, 2: finished, 3: scenario-by-scenario route
    };
    std::vector<Lane> lanes(n);
    for (Lane& l : lanes) {
        l.tap = trafo_st_[ti].tap_pos;
        l.bs.reset(l.tap, t.range, t.control_at_tap_side);
        if (two_stage) l.tap = highest ? t.range.highest_voltage(t.control_at_tap_side) : t.range.lowest_voltage(t.control_at_tap_side);
        l.bs.set_pos(l.tap);
        l.bs.clear_flags();
    }
    OutagePlan plan;
    plan.math_branch.assign(n, branch.pos);
    plan.comp.assign(n, static_cast<int32_t>(seq));
    plan.energized.assign(n, 1);
    plan.dead_off.assign(n, -1);
    plan.bparam.assign(n * 4 * bb2, 0.0);
    auto fill_plan = [&] {
        for (Idx s = 0; s != n; ++s) transformer_param<B>(trafo_c_[ti], branch_st_[seq], tap_limit(trafo_c_[ti], lanes[s].tap), &plan.bparam[s * 4 * bb2]);
    };
    // a probe pass runs on one device and leaves node / transformer results in HBM (update rows stay there after the first
    // pass); only the control node's and the transformer's row of every scenario come back
    bool rows_resident = false;
    auto run_pass = [&](bool probe_pass, OutputData const& o, int32_t* it, int32_t* st) {
        fill_plan();
        outage_plan_ = &plan;
        int64_t r = -1;
        try {
            if (probe_pass) {
                ModelOptions po = opt;
                po.n_devices = 1;
                po.flags = kFlagResidentOutput | (rows_resident ? kFlagResidentInput : 0u);
                r = run_batch_device_one(po, B, update, o, it, st, 0);
                rows_resident = true;
            } else {
                r = run_batch_device(opt, B, update, o, it, st);
            }
        } catch (...) {
            outage_plan_ = nullptr;
            throw;
        }
        outage_plan_ = nullptr;
        return r;
    };
    std::vector<NodeOutput<B>> node_probe(n);
    std::vector<BranchOutput<B>> trafo_probe(n);
    std::vector<int32_t> pass_status(n, 0), pass_iter(n, 0);
    unsigned char selects_component = 0; // resident output: the pointers only say which components are produced
    OutputData probe{};
    probe.node = &selects_component;
    probe.transformer = &selects_component;
    auto compare = [&](Idx s) { // -1 below the band, 0 inside, +1 above
        NodeOutput<B> const& no = node_probe[s];
        BranchOutput<B> const& bo = trafo_probe[s];
        double const base_power = B == 1 ? kBasePower3p : kBasePower1p;
        double v = 0.0;
        for (int p = 0; p != B; ++p) {
            cplx const u = std::polar(no.u_pu[p], no.u_angle[p]);
            cplx const s_pu = cplx{side == 0 ? bo.p_from[p] : bo.p_to[p], side == 0 ? bo.q_from[p] : bo.q_to[p]} / base_power;
            cplx const i_pu = std::abs(u) > 0.0 ? std::conj(s_pu / u) : cplx{};
            v += std::abs(u + z_comp * i_pu);
        }
        v /= B;
        return v < lower ? -1 : v > upper ? 1 : 0;
    };
    uint64_t const pass_cap = 8 * (width + 4);
    for (uint64_t pass = 0;; ++pass) {
        bool any_active = false;
        for (Lane const& l : lanes) any_active = any_active || l.stage < 2;
        if (!any_active) break;
        if (pass > pass_cap) {
            for (Lane& l : lanes)
                if (l.stage < 2) l.stage = 3;
            break;
        }
        batch_message.clear();
        if (run_pass(true, probe, pass_iter.data(), pass_status.data()) < 0) return -1;
        if (last_pass_parts_ != 1) return -1; // the batch does not fit one pass: its results are not resident as a whole
        fetch_resident_rows(0, sizeof(NodeOutput<B>), nn, control_node, n, node_probe.data());
        fetch_resident_rows(2, sizeof(BranchOutput<B>), n_trafo(), ti, n, trafo_probe.data());
        for (Idx s = 0; s != n; ++s) {
            Lane& l = lanes[s];
            if (l.stage >= 2) continue;
            if (pass_status[s] != 0) { // IterationDiverge / SparseMatrixError: the reference retries after a linear pass
                l.stage = 3;
                continue;
            }
            int const cmp = compare(s);
            Search const search = l.stage == 0 ? first_search : Search::scan;
            bool changed = false;
            IntS next = l.tap;
            if (search == Search::scan) {
                next = cmp > 0 ? t.range.voltage_down(l.tap, t.control_at_tap_side) : cmp < 0 ? t.range.voltage_up(l.tap, t.control_at_tap_side) : l.tap;
                changed = next != l.tap;
            } else if (!l.bs.exhausted() && !l.bs.settled()) {
                if (cmp != 0) l.bs.step_towards_band(highest, cmp > 0);
                if (IntS const proposed = l.bs.pos(); proposed != l.tap) {
                    next = proposed;
                    changed = true;
                } else if (!(goal == Goal::fast_any && cmp == 0)) {
                    bool const previous_down = l.bs.went_down();
                    l.bs.keep_as_bound(highest);
                    IntS const candidate = l.bs.next_candidate(highest, previous_down, changed);
                    if (candidate == l.tap && cmp != 0 && !l.bs.exhausted()) { // no valid position: MaxIterationReached in the reference
                        l.stage = 3;
                        continue;
                    }
                    next = candidate;
                }
            }
            if (changed) {
                if (++l.passes > 2 * width) {
                    l.stage = 3;
                    continue;
                }
                l.tap = tap_limit(trafo_c_[ti], next);
            } else if (l.stage == 0 && two_stage) {
                l.tap = tap_limit(trafo_c_[ti], highest ? t.range.voltage_up(l.tap, t.control_at_tap_side) : t.range.voltage_down(l.tap, t.control_at_tap_side));
                l.stage = 1;
                l.passes = 0;
            } else {
                l.stage = 2;
            }
        }
    }
    // the caller's outputs with every scenario at its final tap position
    batch_message.clear();
    int64_t const failed = run_pass(false, out, n_iter, status);
    if (failed < 0) return -1;
    exact.clear();
    for (Idx s = 0; s != n; ++s) {
        if (lanes[s].stage == 3 || status[s] != 0) exact.push_back(s);
    }
    if (out.transformer_tap_regulator != nullptr) {
        Idx const n_reg = static_cast<Idx>(tap_reg_in_.size());
        auto* dst = static_cast<TransformerTapRegulatorOutput*>(out.transformer_tap_regulator);
        for (Idx s = 0; s != n; ++s) {
            if (lanes[s].stage == 3 || status[s] != 0) continue;
            dst[s * n_reg + t.regulator].energized = 1;
            dst[s * n_reg + t.regulator].tap_pos = lanes[s].tap;
        }
    }
    return failed;
}

template int64_t Model::run_tap_lockstep<1>(ModelOptions const&, UpdateData const&, OutputData const&, int32_t*, int32_t*, std::vector<Idx>&);
template int64_t Model::run_tap_lockstep<3>(ModelOptions const&, UpdateData const&, OutputData const&, int32_t*, int32_t*, std::vector<Idx>&);
template int64_t Model::run_tap_optimizer<1>(ModelOptions const&, OutputData const&, Idx, int32_t*, int32_t*);
template int64_t Model::run_tap_optimizer<3>(ModelOptions const&, OutputData const&, Idx, int32_t*, int32_t*);

} // namespace pgmb
