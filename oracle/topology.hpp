// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the reference symbolic stage:
//   power_grid_model/topology.hpp         build_topology :137-156, build_sparse_graph :196-237, dfs_search :239-298,
//                                         reorder_node :302-368, couple_branch :370-443, couple_object_components :497-545,
//                                         couple_all_appliance :547-571
//   power_grid_model/sparse_ordering.hpp  DegreeLookup :21-62, remove_vertices_update_degrees :128-180,
//                                         minimum_degree_ordering :183-219
//   power_grid_model/index_mapping.hpp    build_sparse_mapping :27-59 (stable counting sort)
// Boost.Graph semantics restated per SURVEY.md Appendix A: CSR graph keeps the input edge order within a source vertex;
// depth_first_visit = recursive DFS in out-edge order; back_edge fires for gray targets (incl. the edge to the parent).
#pragma once

#include "ybus.hpp"

#include <map>
#include <set>

namespace pgm_oracle {

struct Idx2D {
    Idx group;
    Idx pos;
    bool operator==(Idx2D const&) const = default;
};
using Branch3Idx = std::array<Idx, 3>;

struct ComponentTopology {
    Idx n_node{};
    std::vector<BranchIdx> branch_node_idx;
    std::vector<Branch3Idx> branch3_node_idx;
    IdxVector shunt_node_idx;
    IdxVector source_node_idx;
    IdxVector load_gen_node_idx;
    std::vector<LoadGenType> load_gen_type;
    IdxVector regulated_load_gen_idx; // per voltage regulator: sequence index of the regulated load_gen
    Idx n_node_total() const { return n_node + static_cast<Idx>(branch3_node_idx.size()); }
};
struct ComponentConnections {
    std::vector<std::array<IntS, 2>> branch_connected;
    std::vector<std::array<IntS, 3>> branch3_connected;
    std::vector<double> branch_phase_shift;
    std::vector<std::array<double, 3>> branch3_phase_shift;
    std::vector<IntS> source_connected;
};
struct ComponentToMathCoupling {
    std::vector<Idx2D> node, branch, shunt, load_gen, source, voltage_regulator;
    std::vector<std::pair<Idx, Branch3Idx>> branch3; // group, pos[3]
};

// ---- sparse_ordering.hpp ---------------------------------------------------------------------------------------
namespace ordering {
using Graph = std::map<Idx, IdxVector>;

struct DegreeLookup {
    std::map<Idx, Idx> vertex_to_degree;
    std::map<Idx, std::set<Idx>> degrees_to_vertex;
    void drop_from_bucket(Idx u, Idx degree) {
        auto it = degrees_to_vertex.find(degree);
        if (it == degrees_to_vertex.end()) return;
        it->second.erase(u);
        if (it->second.empty()) degrees_to_vertex.erase(it);
    }
    void set(Idx u, Idx degree) {
        auto it = vertex_to_degree.find(u);
        if (it != vertex_to_degree.end()) {
            drop_from_bucket(u, it->second);
            it->second = degree;
        } else {
            vertex_to_degree.emplace(u, degree);
        }
        degrees_to_vertex[degree].insert(u);
    }
    void erase(Idx u) {
        auto it = vertex_to_degree.find(u);
        if (it == vertex_to_degree.end()) return;
        Idx const degree = it->second;
        vertex_to_degree.erase(it);
        drop_from_bucket(u, degree);
    }
    Idx min_vertex() const { return *degrees_to_vertex.begin()->second.begin(); }
};

inline bool has_edge(Idx from, Idx to, Graph const& d) {
    auto it = d.find(from);
    return it != d.end() && std::find(it->second.begin(), it->second.end(), to) != it->second.end();
}

inline IdxVector eliminate(Idx const u, Graph& d, DegreeLookup& dgd, std::vector<std::pair<Idx, Idx>>& fills) {
    // indistinguishable neighbours: closed neighbourhoods equal
    IdxVector nbs = d.at(u);
    IdxVector closed_u = nbs;
    closed_u.push_back(u);
    std::sort(closed_u.begin(), closed_u.end());
    IdxVector same;
    for (Idx const v : nbs) {
        IdxVector closed_v = d.at(v);
        closed_v.push_back(v);
        std::sort(closed_v.begin(), closed_v.end());
        if (closed_u == closed_v) same.push_back(v);
    }
    IdxVector const alpha = same;
    same.push_back(u);
    for (Idx const uu : same) {
        if (uu != u) std::erase(nbs, uu);
        dgd.erase(uu);
        IdxVector emptied;
        for (Idx const e : d[uu]) {
            auto& adjacents = d[e];
            std::erase(adjacents, uu);
            if (adjacents.empty()) emptied.push_back(e);
        }
        emptied.push_back(uu);
        for (Idx const e : emptied) d.erase(e);
    }
    // make the remaining neighbours a clique; record new edges as fill-ins (clique map iterated in key order)
    Graph clique;
    for (size_t i = 0; i != nbs.size(); ++i) {
        IdxVector others;
        for (size_t j = 0; j != nbs.size(); ++j)
            if (j != i) others.push_back(nbs[j]);
        clique[nbs[i]] = std::move(others);
    }
    for (auto const& [k, adjacent] : clique) {
        auto it = d.find(k);
        for (Idx const e : adjacent) {
            if (!has_edge(k, e, d)) {
                if (it == d.end()) it = d.try_emplace(k).first;
                it->second.push_back(e);
                d[e].push_back(k);
                fills.emplace_back(k, e);
            }
        }
    }
    for (Idx const e : nbs) {
        auto it = d.find(e);
        dgd.set(e, it == d.end() ? 0 : static_cast<Idx>(it->second.size()));
    }
    return alpha;
}

inline std::pair<IdxVector, std::vector<std::pair<Idx, Idx>>> minimum_degree_ordering(Graph d) {
    for (auto const& [k, adjacent] : d) {
        for (Idx const e : adjacent) d[e].push_back(k);
    }
    for (auto& [k, adjacent] : d) {
        std::set<Idx> const uniq{adjacent.begin(), adjacent.end()};
        adjacent.assign(uniq.begin(), uniq.end());
    }
    DegreeLookup dgd;
    for (auto const& [k, adjacent] : d) dgd.set(k, static_cast<Idx>(adjacent.size()));
    Idx const n = static_cast<Idx>(d.size());
    IdxVector alpha;
    std::vector<std::pair<Idx, Idx>> fills;
    for (Idx k = 0; k < n; ++k) {
        Idx const u = dgd.min_vertex();
        alpha.push_back(u);
        if (d.size() == 2) {
            Idx const from = d.begin()->first;
            Idx const to = d.begin()->second[0];
            alpha.push_back(alpha.back() == from ? to : from);
            return {alpha, fills};
        }
        IdxVector const more = eliminate(u, d, dgd, fills);
        alpha.insert(alpha.end(), more.begin(), more.end());
        if (d.empty()) return {alpha, fills};
    }
    return {alpha, fills};
}
} // namespace ordering

// ---- topology.hpp ----------------------------------------------------------------------------------------------
class Topology {
  public:
    Topology(ComponentTopology const& comp_topo, ComponentConnections const& comp_conn)
        : ct_{comp_topo},
          cc_{comp_conn},
          phase_shift_(ct_.n_node_total(), 0.0),
          predecessors_(ct_.n_node_total()),
          node_status_(ct_.n_node_total(), not_processed) {
        for (Idx i = 0; i != ct_.n_node_total(); ++i) predecessors_[i] = i;
    }

    std::pair<std::vector<MathTopology>, ComponentToMathCoupling> build_topology() {
        Idx2D const unknown{-1, -1};
        coup_.node.assign(ct_.n_node_total(), unknown);
        coup_.branch.assign(ct_.branch_node_idx.size(), unknown);
        coup_.branch3.assign(ct_.branch3_node_idx.size(), {-1, {-1, -1, -1}});
        coup_.shunt.assign(ct_.shunt_node_idx.size(), unknown);
        coup_.load_gen.assign(ct_.load_gen_node_idx.size(), unknown);
        coup_.source.assign(ct_.source_node_idx.size(), unknown);
        build_sparse_graph();
        dfs_search();
        couple_branch();
        couple_all_appliance();
        return {std::move(math_), std::move(coup_)};
    }

  private:
    static constexpr Idx not_processed = -1;
    static constexpr Idx in_cycle = -2;
    ComponentTopology const& ct_;
    ComponentConnections const& cc_;
    // CSR graph
    IdxVector adj_ptr_;
    IdxVector adj_target_;
    std::vector<double> adj_shift_;
    std::vector<double> phase_shift_;
    IdxVector predecessors_;
    IdxVector node_status_;
    std::vector<int8_t> color_; // 0 white 1 gray 2 black
    std::vector<MathTopology> math_;
    ComponentToMathCoupling coup_;

    void build_sparse_graph() {
        std::vector<std::array<Idx, 2>> edges;
        std::vector<double> props;
        for (size_t b = 0; b != ct_.branch_node_idx.size(); ++b) {
            auto const [i, j] = ct_.branch_node_idx[b];
            auto const [si, sj] = cc_.branch_connected[b];
            if (si != 0 && sj != 0 && i != j) {
                edges.push_back({i, j});
                props.push_back(-cc_.branch_phase_shift[b]);
                edges.push_back({j, i});
                props.push_back(cc_.branch_phase_shift[b]);
            }
        }
        for (size_t b = 0; b != ct_.branch3_node_idx.size(); ++b) {
            Idx const j_internal = ct_.n_node + static_cast<Idx>(b);
            for (int m = 0; m != 3; ++m) {
                if (cc_.branch3_connected[b][m] != 0) {
                    edges.push_back({ct_.branch3_node_idx[b][m], j_internal});
                    props.push_back(-cc_.branch3_phase_shift[b][m]);
                    edges.push_back({j_internal, ct_.branch3_node_idx[b][m]});
                    props.push_back(cc_.branch3_phase_shift[b][m]);
                }
            }
        }
        Idx const n = ct_.n_node_total();
        adj_ptr_.assign(n + 1, 0);
        for (auto const& e : edges) ++adj_ptr_[e[0] + 1];
        for (Idx i = 0; i != n; ++i) adj_ptr_[i + 1] += adj_ptr_[i];
        adj_target_.resize(edges.size());
        adj_shift_.resize(edges.size());
        IdxVector cursor(adj_ptr_.begin(), adj_ptr_.end() - 1);
        for (size_t e = 0; e != edges.size(); ++e) { // stable placement
            Idx const pos = cursor[edges[e][0]]++;
            adj_target_[pos] = edges[e][1];
            adj_shift_[pos] = props[e];
        }
        color_.assign(n, 0);
    }

    void dfs_search() {
        Idx math_idx = 0;
        for (size_t s = 0; s != ct_.source_node_idx.size(); ++s) {
            if (cc_.source_connected[s] == 0) continue;
            Idx const source_node = ct_.source_node_idx[s];
            if (coup_.node[source_node].group != -1) continue;
            IdxVector dfs_node;
            std::vector<std::pair<Idx, Idx>> back_edges;
            // iterative DFS equivalent to boost::depth_first_visit
            std::vector<std::pair<Idx, Idx>> stack; // (vertex, next edge position)
            auto discover = [&](Idx u) {
                color_[u] = 1;
                coup_.node[u].group = math_idx;
                dfs_node.push_back(u);
                stack.emplace_back(u, adj_ptr_[u]);
            };
            discover(source_node);
            while (!stack.empty()) {
                auto& [u, pos] = stack.back();
                if (pos == adj_ptr_[u + 1]) {
                    color_[u] = 2;
                    stack.pop_back();
                    continue;
                }
                Idx const e = pos++;
                Idx const t = adj_target_[e];
                if (color_[t] == 0) {
                    phase_shift_[t] = phase_shift_[u] + adj_shift_[e];
                    predecessors_[t] = u;
                    discover(t); // invalidates u/pos references, loop re-reads stack.back()
                } else if (color_[t] == 1) {
                    if (predecessors_[u] != t) back_edges.emplace_back(u, t);
                }
            }
            MathTopology topo{};
            if (back_edges.empty()) {
                std::reverse(dfs_node.begin(), dfs_node.end());
                topo.is_radial = true;
            } else {
                topo.fill_in = reorder_node(dfs_node, back_edges);
                topo.is_radial = false;
            }
            topo.phase_shift.resize(dfs_node.size());
            for (size_t i = 0; i != dfs_node.size(); ++i) {
                coup_.node[dfs_node[i]].pos = static_cast<Idx>(i);
                topo.phase_shift[i] = phase_shift_[dfs_node[i]];
            }
            topo.slack_bus = coup_.node[source_node].pos;
            math_.push_back(std::move(topo));
            ++math_idx;
        }
    }

    std::vector<BranchIdx> reorder_node(IdxVector& dfs_node, std::vector<std::pair<Idx, Idx>> const& back_edges) {
        std::vector<BranchIdx> fill_in;
        IdxVector const dfs_copy(dfs_node);
        dfs_node.clear();
        for (auto const& be : back_edges) {
            Idx node = be.first;
            while (node_status_[node] != in_cycle) {
                node_status_[node] = in_cycle;
                node = predecessors_[node];
            }
        }
        for (auto it = dfs_copy.rbegin(); it != dfs_copy.rend(); ++it)
            if (node_status_[*it] == not_processed) dfs_node.push_back(*it);
        IdxVector cyclic;
        for (Idx const x : dfs_copy)
            if (node_status_[x] == in_cycle) cyclic.push_back(x);
        if (cyclic.size() < 4) {
            dfs_node.insert(dfs_node.end(), cyclic.rbegin(), cyclic.rend());
            return fill_in;
        }
        ordering::Graph nn;
        for (Idx const node : cyclic) {
            Idx const pred = predecessors_[node];
            if (pred != node) nn[node] = {pred};
        }
        for (auto const& [from, to] : back_edges) {
            if (!ordering::has_edge(from, to, nn)) nn[from].push_back(to);
        }
        auto [reordered, fills] = ordering::minimum_degree_ordering(std::move(nn));
        Idx const n_non_cyclic = static_cast<Idx>(dfs_node.size());
        std::map<Idx, Idx> permuted;
        for (size_t i = 0; i != reordered.size(); ++i) permuted[reordered[i]] = n_non_cyclic + static_cast<Idx>(i);
        dfs_node.insert(dfs_node.end(), reordered.begin(), reordered.end());
        for (auto [from, to] : fills) fill_in.push_back({permuted[from], permuted[to]});
        return fill_in;
    }

    void couple_branch() {
        auto pos_if = [](IntS status, Idx2D const& m) { return status == 0 ? Idx{-1} : m.pos; };
        for (size_t b = 0; b != ct_.branch_node_idx.size(); ++b) {
            auto const [i, j] = ct_.branch_node_idx[b];
            IntS const si = cc_.branch_connected[b][0];
            IntS const sj = cc_.branch_connected[b][1];
            Idx2D const im = coup_.node[i];
            Idx2D const jm = coup_.node[j];
            Idx group = -1;
            if (si != 0 && im.group != -1) {
                group = im.group;
            } else if (sj != 0 && jm.group != -1) {
                group = jm.group;
            }
            if (group == -1) continue;
            Idx const pos = math_[group].n_branch();
            math_[group].branch_bus_idx.push_back({pos_if(si, im), pos_if(sj, jm)});
            coup_.branch[b] = {group, pos};
        }
        for (size_t b = 0; b != ct_.branch3_node_idx.size(); ++b) {
            auto const& i = ct_.branch3_node_idx[b];
            auto const& st = cc_.branch3_connected[b];
            Idx2D const jm = coup_.node[ct_.n_node + static_cast<Idx>(b)];
            Idx group = -1;
            for (int n = 0; n != 3; ++n)
                if (st[n] != 0 && coup_.node[i[n]].group != -1) group = coup_.node[i[n]].group;
            if (group == -1) continue;
            Branch3Idx pos3{};
            for (int n = 0; n != 3; ++n) {
                Idx const pos = math_[group].n_branch();
                math_[group].branch_bus_idx.push_back({pos_if(st[n], coup_.node[i[n]]), jm.pos});
                pos3[n] = pos;
            }
            coup_.branch3[b] = {group, pos3};
        }
    }

    // stable counting sort of components by bus (index_mapping.hpp build_sparse_mapping); writes indptr + coupling
    template <class Include>
    void couple_objects(IdxVector const& obj_node_idx, IdxVector MathTopology::*indptr_member,
                        std::vector<Idx2D>& coupling, Include include) {
        Idx const n_math = static_cast<Idx>(math_.size());
        std::vector<IdxVector> topo_obj_idx(n_math), topo_comp_idx(n_math);
        for (size_t c = 0; c != obj_node_idx.size(); ++c) {
            if (!include(static_cast<Idx>(c))) continue;
            Idx2D const m = coup_.node[obj_node_idx[c]];
            if (m.group >= 0) {
                topo_obj_idx[m.group].push_back(m.pos);
                topo_comp_idx[m.group].push_back(static_cast<Idx>(c));
            }
        }
        for (Idx g = 0; g != n_math; ++g) {
            Idx const n_bus = math_[g].n_bus();
            IdxVector& indptr = math_[g].*indptr_member;
            indptr.assign(n_bus + 1, 0);
            for (Idx const bus : topo_obj_idx[g]) ++indptr[bus + 1];
            for (Idx i = 0; i != n_bus; ++i) indptr[i + 1] += indptr[i];
            IdxVector cursor(indptr.begin(), indptr.end() - 1);
            for (size_t k = 0; k != topo_obj_idx[g].size(); ++k) {
                Idx const new_pos = cursor[topo_obj_idx[g][k]]++;
                coupling[topo_comp_idx[g][k]] = {g, new_pos};
            }
        }
    }

    void couple_all_appliance() {
        auto all = [](Idx) { return true; };
        couple_objects(ct_.shunt_node_idx, &MathTopology::shunts_per_bus, coup_.shunt, all);
        couple_objects(ct_.load_gen_node_idx, &MathTopology::load_gens_per_bus, coup_.load_gen, all);
        for (auto& m : math_) m.load_gen_type.resize(m.n_load_gen());
        for (size_t i = 0; i != coup_.load_gen.size(); ++i) {
            if (coup_.load_gen[i].group == -1) continue;
            math_[coup_.load_gen[i].group].load_gen_type[coup_.load_gen[i].pos] = ct_.load_gen_type[i];
        }
        couple_objects(ct_.source_node_idx, &MathTopology::sources_per_bus, coup_.source,
                       [this](Idx i) { return cc_.source_connected[i] != 0; });
        couple_voltage_regulators();
    }

    // topology.hpp:594-600: regulators grouped by the math load_gen they regulate (stable counting sort)
    void couple_voltage_regulators() {
        coup_.voltage_regulator.assign(ct_.regulated_load_gen_idx.size(), Idx2D{-1, -1});
        Idx const n_math = static_cast<Idx>(math_.size());
        std::vector<IdxVector> obj(n_math), comp(n_math);
        for (size_t r = 0; r != ct_.regulated_load_gen_idx.size(); ++r) {
            Idx2D const m = coup_.load_gen[ct_.regulated_load_gen_idx[r]];
            if (m.group >= 0) {
                obj[m.group].push_back(m.pos);
                comp[m.group].push_back(static_cast<Idx>(r));
            }
        }
        for (Idx g = 0; g != n_math; ++g) {
            Idx const n_lg = math_[g].n_load_gen();
            IdxVector& indptr = math_[g].voltage_regulators_per_load_gen;
            indptr.assign(n_lg + 1, 0);
            for (Idx const lg : obj[g]) ++indptr[lg + 1];
            for (Idx i = 0; i != n_lg; ++i) indptr[i + 1] += indptr[i];
            IdxVector cursor(indptr.begin(), indptr.end() - 1);
            for (size_t k = 0; k != obj[g].size(); ++k) coup_.voltage_regulator[comp[g][k]] = {g, cursor[obj[g][k]]++};
        }
    }
};

} // namespace pgm_oracle
