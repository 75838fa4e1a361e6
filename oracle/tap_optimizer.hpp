// TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's automatic tap changer,
// power_grid_model/optimizer/tap_position_optimizer.hpp, for the checker side of the parity tests.  Only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference leg may use it.  Pinned by the reference's own validation cases
// (tests/data/power_flow/automatic-tap-regulator/* -> tests/golden/tap_regulator_cases.json, tests/test_oracle_tap_changer.py)
// and the ranking expectations of tests/cpp_unit_tests/optimizer/test_tap_position_optimizer.cpp.
//
// The optimizer is generic over its host (the component model): the host supplies the transformer graph, the state of the
// regulated transformers, a calculator (state -> solver output per math group) and an updater (tap positions).
#pragma once

#include "pf_solvers.hpp"
#include "topology.hpp"

#include <algorithm>
#include <functional>
#include <limits>
#include <queue>
#include <string>
#include <vector>

namespace pgm_oracle::tap {

// PGM_TapChangingStrategy -> OptimizerStrategy (power_grid_model_c/src/model.cpp:150-167)
enum class Strategy : int { any = 1, global_minimum = 2, global_maximum = 3, fast_any = 4 };
enum class SearchMethod { linear_search, binary_search };

constexpr Idx infty = std::numeric_limits<Idx>::max();
constexpr Idx last_rank = infty - 1;

// ---- transformer ranking (tap_position_optimizer.hpp:78-433) -------------------------------------------------------------
struct GraphEdge {
    Idx from, to;
    Idx weight;    // 1: regulated transformer (points at the control side), 0: anything else (given both ways)
    Idx2D regulated; // {kind, index}; {-1, -1}: not regulated.  kind 0: transformer, 1: three-winding transformer
    ID id;
};
struct RankedTransformer {
    Idx2D regulated;
};
using RankedGroups = std::vector<std::vector<Idx2D>>;

inline RankedGroups rank_transformers(Idx n_node, std::vector<GraphEdge> edges, std::vector<char> const& is_source) {
    // compressed-sparse-row graph built from unsorted edges: grouped by source vertex, input order kept (:283-305)
    std::stable_sort(edges.begin(), edges.end(), [](GraphEdge const& a, GraphEdge const& b) { return a.from < b.from; });
    std::vector<Idx> row(n_node + 1, 0);
    for (auto const& e : edges) ++row[e.from + 1];
    for (Idx v = 0; v != n_node; ++v) row[v + 1] += row[v];
    // Dijkstra from every source vertex, distances shared (:308-345)
    std::vector<Idx> distance(n_node, infty);
    for (Idx s = 0; s != n_node; ++s) {
        if (!is_source[s]) continue;
        using Element = std::pair<Idx, Idx>;
        std::priority_queue<Element, std::vector<Element>, std::greater<>> pq;
        distance[s] = 0;
        pq.emplace(0, s);
        while (!pq.empty()) {
            auto [dist, u] = pq.top();
            pq.pop();
            if (dist != distance[u]) continue;
            for (Idx k = row[u]; k != row[u + 1]; ++k) {
                auto const& e = edges[k];
                if (distance[e.from] + e.weight < distance[e.to]) {
                    distance[e.to] = distance[e.from] + e.weight;
                    pq.emplace(distance[e.to], e.to);
                }
            }
        }
    }
    // get_edge_weights (:338-407)
    struct Weighted {
        Idx2D regulated;
        Idx weight;
    };
    std::vector<Weighted> result;
    std::vector<ID> invalid;
    for (auto const& e : edges) {
        if (e.regulated.group == -1) continue;
        Idx const src = distance[e.from], tgt = distance[e.to];
        if (src == infty && tgt == infty) continue;
        if (src == infty || tgt == infty) {
            invalid.push_back(e.id);
        } else if (src != tgt - 1) {
            result.push_back({e.regulated, last_rank});
        } else {
            result.push_back({e.regulated, tgt});
        }
    }
    if (!invalid.empty()) {
        std::sort(invalid.begin(), invalid.end());
        invalid.erase(std::unique(invalid.begin(), invalid.end()), invalid.end());
        std::string msg = "Automatic tap changer has invalid configuration. The following transformer(s) are being controlled "
                          "from non-source side towards source side:\n  Transformer IDs: ";
        for (size_t i = 0; i != invalid.size(); ++i) msg += (i > 0 ? ", " : "") + std::to_string(invalid[i]);
        throw PgmError{msg};
    }
    // rank_transformers (:409-428)
    std::stable_sort(result.begin(), result.end(), [](Weighted const& x, Weighted const& y) { return x.weight < y.weight; });
    RankedGroups groups;
    Idx previous = std::numeric_limits<Idx>::lowest();
    for (auto const& t : result) {
        if (t.weight > previous) {
            groups.emplace_back();
            previous = t.weight;
        }
        auto& group = groups.back();
        bool seen = false;
        for (auto const& x : group) seen = seen || (x.group == t.regulated.group && x.pos == t.regulated.pos);
        if (!seen) group.push_back(t.regulated);
    }
    return groups;
}

// ---- one regulated transformer as the optimizer sees it (:485-535) ---------------------------------------------------------
struct Regulated {
    Idx2D index;       // {kind, index}
    Idx regulator;     // index of its regulator in the model
    IntS tap_min, tap_max;
    bool control_at_tap_side;
};

// :442-483
inline IntS one_step_tap_up(IntS tap_pos, IntS tap_min, IntS tap_max) {
    if (tap_pos == tap_max) return tap_max;
    return tap_min < tap_max ? static_cast<IntS>(tap_pos + 1) : static_cast<IntS>(tap_pos - 1);
}
inline IntS one_step_tap_down(IntS tap_pos, IntS tap_min, IntS tap_max) {
    if (tap_pos == tap_min) return tap_min;
    return tap_min < tap_max ? static_cast<IntS>(tap_pos - 1) : static_cast<IntS>(tap_pos + 1);
}
inline IntS one_step_control_voltage_up(IntS tap_pos, Regulated const& r) {
    return r.control_at_tap_side ? one_step_tap_up(tap_pos, r.tap_min, r.tap_max) : one_step_tap_down(tap_pos, r.tap_min, r.tap_max);
}
inline IntS one_step_control_voltage_down(IntS tap_pos, Regulated const& r) {
    return r.control_at_tap_side ? one_step_tap_down(tap_pos, r.tap_min, r.tap_max) : one_step_tap_up(tap_pos, r.tap_min, r.tap_max);
}

// :785-907
class BinarySearch {
  public:
    BinarySearch() = default;
    BinarySearch(IntS tap_pos, IntS tap_min, IntS tap_max, bool control_at_tap_side) { reset(tap_pos, tap_min, tap_max, control_at_tap_side); }
    IntS get_current_tap() const { return current_; }
    bool get_last_down() const { return last_down_; }
    bool get_inevitable_run() const { return inevitable_run_; }
    bool get_end_of_bs() const { return lower_bound_ >= upper_bound_; }
    void set_current_tap(IntS tap) { current_ = tap; }
    void set_last_check(bool v) { last_check_ = v; }
    void set_inevitable_run(bool v) { inevitable_run_ = v; }

    void recalibrate(bool strategy_max) {
        bool const invert_strategy = control_at_tap_side_ != strategy_max;
        if (tap_reverse_ == invert_strategy) {
            lower_bound_ = current_;
            last_down_ = false;
        } else {
            upper_bound_ = current_;
            last_down_ = true;
        }
    }
    void propose_new_pos(bool strategy_max, bool above_range) {
        bool const is_down = (above_range == tap_reverse_) != control_at_tap_side_;
        if (last_check_) {
            current_ = is_down ? lower_bound_ : upper_bound_;
            inevitable_run_ = true;
        } else {
            last_down_ = is_down;
            adjust(strategy_max);
        }
    }
    IntS repropose_tap(bool strategy_max, bool previous_down, bool& tap_changed) {
        bool const prefer_higher = (strategy_max != tap_reverse_) != control_at_tap_side_;
        IntS const tap_pos = search(prefer_higher);
        int const tap_diff = tap_pos - current_;
        if (tap_diff == 0) {
            if (!inevitable_run_) {
                inevitable_run_ = true;
                tap_changed = true;
            } else {
                tap_changed = false;
            }
            return tap_pos;
        }
        if ((tap_diff == 1 && previous_down) || (tap_diff == -1 && !previous_down)) last_check_ = true;
        tap_changed = true;
        current_ = tap_pos;
        return tap_pos;
    }
    void rewind(IntS tap_pos, IntS tap_min, IntS tap_max) { reset(tap_pos, tap_min, tap_max, control_at_tap_side_); }

  private:
    void reset(IntS tap_pos, IntS tap_min, IntS tap_max, bool control_at_tap_side) {
        last_down_ = false;
        last_check_ = false;
        current_ = tap_pos;
        inevitable_run_ = false;
        lower_bound_ = std::min(tap_min, tap_max);
        upper_bound_ = std::max(tap_min, tap_max);
        tap_reverse_ = tap_max < tap_min;
        control_at_tap_side_ = control_at_tap_side;
    }
    void adjust(bool strategy_max) {
        if (last_down_) {
            upper_bound_ = current_;
        } else {
            lower_bound_ = current_;
        }
        if (lower_bound_ < upper_bound_) current_ = search(strategy_max != tap_reverse_);
    }
    IntS search(bool prefer_higher_in) const {
        bool const prefer_higher = control_at_tap_side_ != prefer_higher_in;
        int const primary = prefer_higher ? upper_bound_ : lower_bound_;
        int const secondary = prefer_higher ? lower_bound_ : upper_bound_;
        return static_cast<IntS>(primary + (secondary - primary) / 2); // std::midpoint: rounds towards its first argument
    }
    IntS lower_bound_{}, upper_bound_{}, current_{};
    bool last_down_{}, last_check_{}, tap_reverse_{}, inevitable_run_{}, control_at_tap_side_{};
};

// ---- the optimizer (:766-1420) -----------------------------------------------------------------------------------------------
// Host: what the optimizer needs from the model.
//   tap_pos(r) / set_taps({(r, pos)...})   state of the regulated transformers
//   calculate(method)                       power flow on the current state; throws IterationDiverge / SparseMatrixError
//   compare(r)                              NodeState <=> regulator band (:702-731): -1 below, 0 inside, +1 above;
//                                           `connected` false when the control side is not part of a math model
struct Comparison {
    bool connected;
    int cmp;
};
template <class Host> class TapPositionOptimizer {
  public:
    TapPositionOptimizer(Host& host, std::vector<std::vector<Regulated>> order, Strategy strategy)
        : host_{host}, order_{std::move(order)}, strategy_{strategy} {
        // main_model_impl.hpp:337-339
        tap_search_ = strategy == Strategy::any ? SearchMethod::linear_search : SearchMethod::binary_search;
    }

    // optimize (:966-984); the host restores the cached tap positions afterwards
    void optimize(CalculationMethod method) {
        opt_prep();
        pilot_run();
        iterate_with_fallback(method, tap_search_);
        if (strategy_ == Strategy::any || strategy_ == Strategy::fast_any) return;
        exploit_neighborhood();
        iterate_with_fallback(method, SearchMethod::linear_search);
    }
    std::vector<std::vector<Regulated>> const& order() const { return order_; }

  private:
    using Updates = std::vector<std::pair<Regulated const*, IntS>>;

    void opt_prep() { // :986-1015
        if (tap_search_ == SearchMethod::binary_search) {
            for (auto const& group : order_) {
                binary_search_.emplace_back();
                for (auto const& r : group) binary_search_.back().emplace_back(host_.tap_pos(r), r.tap_min, r.tap_max, r.control_at_tap_side);
            }
        }
        for (auto const& group : order_) {
            uint64_t widest = 0;
            for (auto const& r : group) widest = std::max<uint64_t>(widest, static_cast<uint64_t>(std::abs(int{r.tap_max} - int{r.tap_min})));
            max_tap_ranges_per_rank_.push_back(widest);
        }
    }
    void regulate_transformers(std::function<IntS(Regulated const&)> const& to_new_tap_pos) { // :1337-1359
        Updates updates;
        for (auto const& group : order_)
            for (auto const& r : group) updates.emplace_back(&r, to_new_tap_pos(r));
        host_.set_taps(updates);
    }
    void pilot_run() { // :1240-1285
        if (strategy_ == Strategy::global_maximum) {
            regulate_transformers([](Regulated const& r) { return r.control_at_tap_side ? r.tap_max : r.tap_min; });
        } else if (strategy_ == Strategy::global_minimum) {
            regulate_transformers([](Regulated const& r) { return r.control_at_tap_side ? r.tap_min : r.tap_max; });
        }
        if (tap_search_ == SearchMethod::binary_search) {
            for (size_t i = 0; i != order_.size(); ++i) {
                for (size_t j = 0; j != order_[i].size(); ++j) {
                    binary_search_[i][j].set_current_tap(host_.tap_pos(order_[i][j]));
                    binary_search_[i][j].set_last_check(false);
                    binary_search_[i][j].set_inevitable_run(false);
                }
            }
        }
    }
    void exploit_neighborhood() { // :1287-1317
        if (strategy_ == Strategy::global_maximum) {
            regulate_transformers([this](Regulated const& r) { return one_step_control_voltage_up(host_.tap_pos(r), r); });
        } else if (strategy_ == Strategy::global_minimum) {
            regulate_transformers([this](Regulated const& r) { return one_step_control_voltage_down(host_.tap_pos(r), r); });
        }
    }
    void iterate_with_fallback(CalculationMethod method, SearchMethod search) { // :1029-1046
        try {
            iterate(method, search);
        } catch (IterationDiverge const&) { // MaxIterationReached included
            iterate(CalculationMethod::linear, search);
            iterate(method, search);
        } catch (SparseMatrixError const&) {
            iterate(CalculationMethod::linear, search);
            iterate(method, search);
        }
    }
    void iterate(CalculationMethod method, SearchMethod search) { // :1048-1105 with RankIteration (:733-764)
        host_.calculate(method);
        bool const strategy_max = strategy_ == Strategy::global_maximum;
        std::vector<uint64_t> iterations_per_rank(order_.size(), 0);
        bool tap_changed = true;
        while (tap_changed) {
            tap_changed = false;
            Updates update_data;
            size_t rank_index = 0;
            for (size_t i = 0; i != order_.size(); ++i) {
                for (size_t j = 0; j != order_[i].size(); ++j) {
                    bool const adjusted = search == SearchMethod::binary_search
                                              ? adjust_transformer_bs(order_[i][j], binary_search_[i][j], strategy_max, update_data)
                                              : adjust_transformer_scan(order_[i][j], update_data);
                    tap_changed = adjusted || tap_changed;
                }
                if (tap_changed) {
                    if (rank_index + 1 < iterations_per_rank.size()) {
                        std::fill(iterations_per_rank.begin() + static_cast<std::ptrdiff_t>(rank_index) + 1, iterations_per_rank.end(), 0);
                    }
                    ++iterations_per_rank[rank_index];
                    break;
                }
                ++rank_index;
            }
            if (tap_changed) {
                if (iterations_per_rank[rank_index] > 2 * max_tap_ranges_per_rank_[rank_index]) {
                    throw MaxIterationReached{"TapPositionOptimizer::iterate " + std::to_string(iterations_per_rank[rank_index]) +
                                              " iterations reached: " + std::to_string(max_tap_ranges_per_rank_[rank_index]) +
                                              "x2 iterations in rank " + std::to_string(rank_index)};
                }
                host_.set_taps(update_data);
                host_.calculate(method);
            }
        }
    }
    bool adjust_transformer_scan(Regulated const& r, Updates& update_data) { // :1134-1165
        Comparison const c = host_.compare(r);
        if (!c.connected) return false;
        IntS const tap_pos = host_.tap_pos(r);
        IntS const new_tap_pos = c.cmp > 0   ? one_step_control_voltage_down(tap_pos, r)
                                 : c.cmp < 0 ? one_step_control_voltage_up(tap_pos, r)
                                             : tap_pos;
        if (new_tap_pos == tap_pos) return false;
        update_data.emplace_back(&r, new_tap_pos);
        return true;
    }
    bool adjust_transformer_bs(Regulated const& r, BinarySearch& current_bs, bool strategy_max, Updates& update_data) { // :1167-1226
        bool tap_changed = false;
        Comparison const c = host_.compare(r);
        if (!c.connected) return false;
        if (current_bs.get_end_of_bs() || current_bs.get_inevitable_run()) return false;
        IntS const tap_pos_now = host_.tap_pos(r);
        if (c.cmp != 0) current_bs.propose_new_pos(strategy_max, c.cmp > 0);
        if (IntS const new_tap_pos = current_bs.get_current_tap(); new_tap_pos != tap_pos_now) {
            current_bs.set_current_tap(new_tap_pos);
            update_data.emplace_back(&r, new_tap_pos);
            return true;
        }
        if (strategy_ == Strategy::fast_any && c.cmp == 0) return false;
        bool const previous_down = current_bs.get_last_down();
        current_bs.recalibrate(strategy_max);
        IntS const tap_pos = current_bs.repropose_tap(strategy_max, previous_down, tap_changed);
        if (tap_pos == tap_pos_now && c.cmp != 0 && !current_bs.get_end_of_bs()) {
            current_bs.rewind(tap_pos, r.tap_min, r.tap_max);
            throw MaxIterationReached{"TapPositionOptimizer::binary_search: no valid tap position found between tap " +
                                      std::to_string(int{r.tap_min}) + " and tap " + std::to_string(int{r.tap_max})};
        }
        update_data.emplace_back(&r, tap_pos);
        return tap_changed;
    }

    Host& host_;
    std::vector<std::vector<Regulated>> order_;
    Strategy strategy_;
    SearchMethod tap_search_;
    std::vector<std::vector<BinarySearch>> binary_search_;
    std::vector<uint64_t> max_tap_ranges_per_rank_;
};

} // namespace pgm_oracle::tap
