// Symbolic stage, part 1 (host, once per topology): connected islands fed by a source, bus ordering, fill-in list and the
// component -> math-model coupling.  Produces the same MathModelTopology as the reference so that every downstream index
// (LU pattern, iteration counts) is bit-identical:
//   Topology::build_topology   (power_grid_model/topology.hpp:137-156): DFS per source over the branch graph,
//       radial island => reversed discovery order, meshed island => tree part in reversed discovery order followed by the
//       cyclic core in minimum-degree order (reorder_node :302-368, minimum_degree_ordering sparse_ordering.hpp:183-219)
//   couple_branch / couple_all_appliance (:370-443, :547-571): per-bus grouping by stable counting sort.
// Own data structures: flat CSR adjacency, explicit DFS stack, degree buckets in one ordered set; vertex ids of the cyclic
// core are compacted monotonically, which preserves every "smallest id first" tie-break of the reference.
#pragma once

#include "symbolic.hpp"

#include <array>
#include <set>
#include <utility>

namespace pgmb {

struct GridGraph {                      // component-level connectivity, component storage order
    Idx n_node{};
    std::vector<std::array<Idx, 2>> branch_node;   // node sequence numbers
    std::vector<std::array<int8_t, 2>> branch_status;
    std::vector<double> branch_shift;              // theta_from - theta_to
    std::vector<Idx> source_node;
    std::vector<int8_t> source_status;
    std::vector<Idx> shunt_node;
    std::vector<Idx> load_gen_node;
    std::vector<int8_t> load_gen_type;
    std::vector<Idx> regulated_load_gen;           // per voltage regulator: sequence number of its load_gen
    // three-way branches (Branch3: three-winding transformers): side k <-> an internal node n_node + b
    // (topology.hpp:137-156 of the reference: the internal nodes are appended to the node list)
    std::vector<std::array<Idx, 3>> branch3_node;
    std::vector<std::array<int8_t, 3>> branch3_status;
    std::vector<std::array<double, 3>> branch3_shift; // theta_node_k - theta_internal
};

struct Coupling {
    Idx group{-1};
    Idx pos{-1};
};

struct Coupling3 {
    Idx group{-1};
    std::array<Idx, 3> pos{-1, -1, -1};
};
struct TopologyResult {
    std::vector<MathTopology> math;
    std::vector<Coupling> node, branch, shunt, load_gen, source, voltage_regulator; // node: user nodes, then the internal nodes
    std::vector<Coupling3> branch3;
};

namespace detail {

// minimum-degree elimination with merging of indistinguishable vertices; returns (order, fill edges)
inline std::pair<std::vector<Idx>, std::vector<std::array<Idx, 2>>> min_degree(std::vector<std::vector<Idx>> adj) {
    Idx const n = static_cast<Idx>(adj.size());
    // symmetrise + sort + unique
    {
        std::vector<std::vector<Idx>> extra(n);
        for (Idx k = 0; k != n; ++k)
            for (Idx const e : adj[k]) extra[e].push_back(k);
        for (Idx k = 0; k != n; ++k) {
            adj[k].insert(adj[k].end(), extra[k].begin(), extra[k].end());
            std::sort(adj[k].begin(), adj[k].end());
            adj[k].erase(std::unique(adj[k].begin(), adj[k].end()), adj[k].end());
        }
    }
    std::vector<char> alive(n, 1);
    Idx n_alive = n;
    std::set<std::pair<Idx, Idx>> by_degree; // (degree, vertex)
    std::vector<Idx> degree(n);
    for (Idx k = 0; k != n; ++k) {
        degree[k] = static_cast<Idx>(adj[k].size());
        by_degree.emplace(degree[k], k);
    }
    auto set_degree = [&](Idx v, Idx d) {
        by_degree.erase({degree[v], v});
        degree[v] = d;
        by_degree.emplace(d, v);
    };
    auto drop = [&](Idx v) { by_degree.erase({degree[v], v}); };
    auto has = [&](Idx from, Idx to) { return std::find(adj[from].begin(), adj[from].end(), to) != adj[from].end(); };
    auto kill = [&](Idx v) {
        if (alive[v]) {
            alive[v] = 0;
            --n_alive;
        }
        adj[v].clear();
    };

    std::vector<Idx> order;
    std::vector<std::array<Idx, 2>> fills;
    for (Idx step = 0; step < n; ++step) {
        Idx const u = by_degree.begin()->second;
        order.push_back(u);
        if (n_alive == 2) {
            Idx first = 0;
            while (!alive[first]) ++first;
            Idx const other = adj[first][0];
            order.push_back(order.back() == first ? other : first);
            break;
        }
        // neighbours with the same closed neighbourhood are eliminated together with u, in adjacency order
        std::vector<Idx> nbs = adj[u];
        std::vector<Idx> closed_u = nbs;
        closed_u.push_back(u);
        std::sort(closed_u.begin(), closed_u.end());
        std::vector<Idx> twins;
        for (Idx const v : nbs) {
            std::vector<Idx> closed_v = adj[v];
            closed_v.push_back(v);
            std::sort(closed_v.begin(), closed_v.end());
            if (closed_v == closed_u) twins.push_back(v);
        }
        order.insert(order.end(), twins.begin(), twins.end());
        std::vector<Idx> removing = twins;
        removing.push_back(u);
        for (Idx const x : removing) {
            if (x != u) nbs.erase(std::remove(nbs.begin(), nbs.end(), x), nbs.end());
            drop(x);
            std::vector<Idx> const around = adj[x];
            for (Idx const e : around) {
                auto& a = adj[e];
                a.erase(std::remove(a.begin(), a.end(), x), a.end());
                if (a.empty()) kill(e);
            }
            kill(x);
        }
        // clique among the remaining neighbours, visited in ascending vertex order; new edges are the fill-ins
        std::vector<Idx> sorted_nbs = nbs;
        std::sort(sorted_nbs.begin(), sorted_nbs.end());
        for (Idx const k : sorted_nbs) {
            for (Idx const e : nbs) {
                if (e == k || has(k, e)) continue;
                if (!alive[k]) {
                    alive[k] = 1;
                    ++n_alive;
                }
                if (!alive[e]) {
                    alive[e] = 1;
                    ++n_alive;
                }
                adj[k].push_back(e);
                adj[e].push_back(k);
                fills.push_back({k, e});
            }
        }
        for (Idx const e : nbs) set_degree(e, static_cast<Idx>(adj[e].size()));
        if (n_alive == 0) break;
    }
    return {order, fills};
}

} // namespace detail

inline TopologyResult build_topology(GridGraph const& g) {
    Idx const n = g.n_node + static_cast<Idx>(g.branch3_node.size()); // user nodes + internal nodes of the three-way branches
    TopologyResult res;
    res.node.assign(n, {});
    res.branch3.assign(g.branch3_node.size(), {});
    res.branch.assign(g.branch_node.size(), {});
    res.shunt.assign(g.shunt_node.size(), {});
    res.load_gen.assign(g.load_gen_node.size(), {});
    res.source.assign(g.source_node.size(), {});

    // directed CSR graph of fully connected branches, edge order per vertex = branch order (from->to then to->from)
    std::vector<Idx> ptr(n + 1, 0);
    auto connected = [&](size_t b) {
        return g.branch_status[b][0] != 0 && g.branch_status[b][1] != 0 && g.branch_node[b][0] != g.branch_node[b][1];
    };
    for (size_t b = 0; b != g.branch_node.size(); ++b) {
        if (!connected(b)) continue;
        ++ptr[g.branch_node[b][0] + 1];
        ++ptr[g.branch_node[b][1] + 1];
    }
    for (size_t b = 0; b != g.branch3_node.size(); ++b)
        for (int m = 0; m != 3; ++m)
            if (g.branch3_status[b][m] != 0) {
                ++ptr[g.branch3_node[b][m] + 1];
                ++ptr[g.n_node + static_cast<Idx>(b) + 1];
            }
    for (Idx i = 0; i != n; ++i) ptr[i + 1] += ptr[i];
    std::vector<Idx> target(ptr.back());
    std::vector<double> shift(ptr.back());
    {
        std::vector<Idx> cur(ptr.begin(), ptr.end() - 1);
        for (size_t b = 0; b != g.branch_node.size(); ++b) {
            if (!connected(b)) continue;
            auto const [i, j] = g.branch_node[b];
            target[cur[i]] = j;
            shift[cur[i]++] = -g.branch_shift[b];
            target[cur[j]] = i;
            shift[cur[j]++] = g.branch_shift[b];
        }
        for (size_t b = 0; b != g.branch3_node.size(); ++b) { // after all two-way branches, like the reference's edge list
            Idx const j = g.n_node + static_cast<Idx>(b);
            for (int m = 0; m != 3; ++m) {
                if (g.branch3_status[b][m] == 0) continue;
                Idx const i = g.branch3_node[b][m];
                target[cur[i]] = j;
                shift[cur[i]++] = -g.branch3_shift[b][m];
                target[cur[j]] = i;
                shift[cur[j]++] = g.branch3_shift[b][m];
            }
        }
    }

    std::vector<double> node_shift(n, 0.0);
    std::vector<Idx> parent(n);
    for (Idx i = 0; i != n; ++i) parent[i] = i;
    std::vector<int8_t> color(n, 0);  // 0 unvisited, 1 on stack, 2 finished
    std::vector<int8_t> in_core(n, 0);

    for (size_t s = 0; s != g.source_node.size(); ++s) {
        if (g.source_status[s] == 0) continue;
        Idx const root = g.source_node[s];
        if (res.node[root].group != -1) continue;
        Idx const group = static_cast<Idx>(res.math.size());
        std::vector<Idx> order; // discovery order
        std::vector<std::array<Idx, 2>> back_edges;
        std::vector<std::array<Idx, 2>> stack; // (vertex, next edge)
        auto discover = [&](Idx v) {
            color[v] = 1;
            res.node[v].group = group;
            order.push_back(v);
            stack.push_back({v, ptr[v]});
        };
        discover(root);
        while (!stack.empty()) {
            Idx const v = stack.back()[0];
            Idx const e = stack.back()[1];
            if (e == ptr[v + 1]) {
                color[v] = 2;
                stack.pop_back();
                continue;
            }
            ++stack.back()[1];
            Idx const w = target[e];
            if (color[w] == 0) {
                node_shift[w] = node_shift[v] + shift[e];
                parent[w] = v;
                discover(w);
            } else if (color[w] == 1 && parent[v] != w) {
                back_edges.push_back({v, w});
            }
        }

        MathTopology topo;
        std::vector<Idx> bus_order;
        if (back_edges.empty()) {
            bus_order.assign(order.rbegin(), order.rend());
            topo.is_radial = true;
        } else {
            topo.is_radial = false;
            for (auto const& be : back_edges) {
                for (Idx v = be[0]; !in_core[v]; v = parent[v]) in_core[v] = 1;
            }
            std::vector<Idx> core;
            for (auto it = order.rbegin(); it != order.rend(); ++it)
                if (!in_core[*it]) bus_order.push_back(*it);
            for (Idx const v : order)
                if (in_core[v]) core.push_back(v);
            if (core.size() < 4) {
                bus_order.insert(bus_order.end(), core.rbegin(), core.rend());
            } else {
                // compact ids monotonically, keep adjacency insertion order: predecessor first, then back edges
                std::vector<Idx> sorted_core = core;
                std::sort(sorted_core.begin(), sorted_core.end());
                auto local = [&](Idx v) {
                    return static_cast<Idx>(std::lower_bound(sorted_core.begin(), sorted_core.end(), v) - sorted_core.begin());
                };
                std::vector<std::vector<Idx>> adj(sorted_core.size());
                for (Idx const v : core)
                    if (parent[v] != v) adj[local(v)].push_back(local(parent[v]));
                for (auto const& be : back_edges) {
                    Idx const a = local(be[0]), b = local(be[1]);
                    if (std::find(adj[a].begin(), adj[a].end(), b) == adj[a].end()) adj[a].push_back(b);
                }
                auto [elim, fills] = detail::min_degree(std::move(adj));
                Idx const offset = static_cast<Idx>(bus_order.size());
                std::vector<Idx> new_pos(sorted_core.size(), -1);
                for (size_t i = 0; i != elim.size(); ++i) {
                    new_pos[elim[i]] = offset + static_cast<Idx>(i);
                    bus_order.push_back(sorted_core[elim[i]]);
                }
                for (auto const& f : fills) {
                    topo.fill_in.push_back(new_pos[f[0]]);
                    topo.fill_in.push_back(new_pos[f[1]]);
                }
            }
        }
        topo.n_bus = static_cast<Idx>(bus_order.size());
        topo.phase_shift.resize(bus_order.size());
        for (size_t i = 0; i != bus_order.size(); ++i) {
            res.node[bus_order[i]].pos = static_cast<Idx>(i);
            topo.phase_shift[i] = node_shift[bus_order[i]];
        }
        topo.slack_bus = res.node[root].pos;
        res.math.push_back(std::move(topo));
    }

    // branches: a branch belongs to the island of a connected side; disconnected side => bus -1
    for (size_t b = 0; b != g.branch_node.size(); ++b) {
        Coupling const ci = res.node[g.branch_node[b][0]], cj = res.node[g.branch_node[b][1]];
        int8_t const si = g.branch_status[b][0], sj = g.branch_status[b][1];
        Idx group = -1;
        if (si != 0 && ci.group != -1) {
            group = ci.group;
        } else if (sj != 0 && cj.group != -1) {
            group = cj.group;
        }
        if (group == -1) continue;
        auto& m = res.math[group];
        res.branch[b] = {group, m.n_branch()};
        m.branch_bus_idx.push_back(si != 0 ? ci.pos : -1);
        m.branch_bus_idx.push_back(sj != 0 ? cj.pos : -1);
    }
    // three-way branches: three math branches side k -> internal node (couple_branch of the reference, second loop)
    for (size_t b = 0; b != g.branch3_node.size(); ++b) {
        Coupling const cj = res.node[g.n_node + static_cast<Idx>(b)];
        Idx group = -1;
        for (int m = 0; m != 3; ++m) {
            Coupling const ci = res.node[g.branch3_node[b][m]];
            if (g.branch3_status[b][m] != 0 && ci.group != -1) group = ci.group;
        }
        if (group == -1) continue;
        auto& mt = res.math[group];
        res.branch3[b].group = group;
        for (int m = 0; m != 3; ++m) {
            res.branch3[b].pos[m] = mt.n_branch();
            mt.branch_bus_idx.push_back(g.branch3_status[b][m] != 0 ? res.node[g.branch3_node[b][m]].pos : -1);
            mt.branch_bus_idx.push_back(cj.pos);
        }
    }
    // appliances grouped per bus, stable in component order
    auto group_by_bus = [&](std::vector<Idx> const& comp_node, std::vector<Coupling>& coupling,
                            std::vector<Idx> MathTopology::*indptr, auto&& include) {
        for (auto& m : res.math) (m.*indptr).assign(m.n_bus + 1, 0);
        for (size_t c = 0; c != comp_node.size(); ++c) {
            Coupling const nc = res.node[comp_node[c]];
            if (include(c) && nc.group != -1) ++(res.math[nc.group].*indptr)[nc.pos + 1];
        }
        std::vector<std::vector<Idx>> cursor(res.math.size());
        for (size_t gi = 0; gi != res.math.size(); ++gi) {
            auto& ip = res.math[gi].*indptr;
            for (Idx i = 0; i != res.math[gi].n_bus; ++i) ip[i + 1] += ip[i];
            cursor[gi].assign(ip.begin(), ip.end() - 1);
        }
        for (size_t c = 0; c != comp_node.size(); ++c) {
            Coupling const nc = res.node[comp_node[c]];
            if (include(c) && nc.group != -1) coupling[c] = {nc.group, cursor[nc.group][nc.pos]++};
        }
    };
    auto all = [](size_t) { return true; };
    group_by_bus(g.shunt_node, res.shunt, &MathTopology::shunts_per_bus, all);
    group_by_bus(g.load_gen_node, res.load_gen, &MathTopology::load_gens_per_bus, all);
    for (auto& m : res.math) m.load_gen_type.assign(m.n_load_gen(), 0);
    for (size_t c = 0; c != g.load_gen_node.size(); ++c)
        if (res.load_gen[c].group != -1) res.math[res.load_gen[c].group].load_gen_type[res.load_gen[c].pos] = g.load_gen_type[c];
    group_by_bus(g.source_node, res.source, &MathTopology::sources_per_bus, [&](size_t c) { return g.source_status[c] != 0; });
    // voltage regulators grouped by the math load_gen they regulate (topology.hpp:594-600); one regulator per load_gen at
    // most, so the position of a regulator is the rank of its load_gen among the regulated ones
    res.voltage_regulator.assign(g.regulated_load_gen.size(), {});
    if (!g.regulated_load_gen.empty()) {
        for (auto& m : res.math) m.load_gen_regulator.assign(m.n_load_gen(), -1);
        for (size_t r = 0; r != g.regulated_load_gen.size(); ++r) {
            Coupling const c = res.load_gen[g.regulated_load_gen[r]];
            if (c.group != -1) res.math[c.group].load_gen_regulator[c.pos] = static_cast<Idx>(r); // component index for now
        }
        for (auto& m : res.math) {
            Idx pos = 0;
            for (Idx& r : m.load_gen_regulator)
                if (r >= 0) {
                    res.voltage_regulator[r] = {static_cast<Idx>(&m - res.math.data()), pos};
                    r = pos++;
                }
        }
    }
    return res;
}

} // namespace pgmb
