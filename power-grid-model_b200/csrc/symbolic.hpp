// Symbolic stage of the engine (host, once per topology): Y-bus CSR, LU pattern with fill-ins, and the elimination
// schedule every scenario shares.
//
// Replaces, for the GPU engine, what the reference computes in
//   YBusStructure::YBusStructure            (math_solver/y_bus.hpp:122-293)   -> LuPattern
//   the index walk of SparseLUSolver::prefactorize (col_position_idx / find_entry,
//                                            math_solver/sparse_lu_solver.hpp:360, 402-414, 437-489, 734-748)
//                                                                              -> EliminationSchedule
// The reference redoes that index walk inside every factorisation; here it is flattened once into integer arrays:
// for every strictly-lower entry (k, c) the list of (U entry (c, j), target entry (k, j)) pairs, and rows grouped into
// dependency levels (row k depends on the rows c < k it has an entry for), so that a thread block can process all rows
// of a level concurrently.  Row-by-row ("IKJ") elimination performs exactly the reference's floating-point operations
// on each entry, in the same order (updates to an entry arrive in ascending pivot order).
#pragma once

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace pgmb {

using Idx = int64_t;

struct MathTopology {
    Idx n_bus{};
    std::vector<double> phase_shift;
    std::vector<Idx> branch_bus_idx; // [n_branch][2]
    std::vector<Idx> fill_in;        // [n_fill][2]
    std::vector<Idx> sources_per_bus, shunts_per_bus, load_gens_per_bus; // indptr
    std::vector<int8_t> load_gen_type;
    // voltage regulator of each load_gen or -1 (at most one, main_core/input.hpp:216-241); empty = grid without regulators.
    // Regulators are numbered in load_gen order = the reference's grouping by regulated object (topology.hpp:594-600).
    std::vector<Idx> load_gen_regulator;
    Idx slack_bus{};
    bool is_radial{};

    Idx n_branch() const { return static_cast<Idx>(branch_bus_idx.size() / 2); }
    Idx n_fill() const { return static_cast<Idx>(fill_in.size() / 2); }
    Idx n_source() const { return sources_per_bus.empty() ? 0 : sources_per_bus.back(); }
    Idx n_shunt() const { return shunts_per_bus.empty() ? 0 : shunts_per_bus.back(); }
    Idx n_load_gen() const { return load_gens_per_bus.empty() ? 0 : load_gens_per_bus.back(); }
    Idx n_voltage_regulator() const {
        Idx n = 0;
        for (Idx const r : load_gen_regulator) n += r >= 0 ? 1 : 0;
        return n;
    }
};

// element kinds follow YBusElementType (common/enum.hpp:79-87): 0..3 = branch ff/ft/tf/tt, 4 = shunt
struct LuPattern {
    Idx n_bus{}, nnz{}, nnz_lu{};
    std::vector<Idx> row_indptr, col_indices, bus_entry;
    std::vector<Idx> y_bus_entry_indptr;
    std::vector<int8_t> element_type;
    std::vector<Idx> element_idx;
    std::vector<Idx> row_indptr_lu, col_indices_lu, diag_lu, map_lu_y_bus, lu_transpose_entry;

    explicit LuPattern(MathTopology const& topo) {
        n_bus = topo.n_bus;
        struct Item {
            Idx row, col;
            int8_t kind; // 0..4 as above, 5 = fill-in
            Idx idx;
        };
        std::vector<Item> items;
        items.reserve(4 * topo.n_branch() + topo.n_shunt() + 2 * topo.n_fill());
        for (Idx b = 0; b != topo.n_branch(); ++b) {
            Idx const side[2] = {topo.branch_bus_idx[2 * b], topo.branch_bus_idx[2 * b + 1]};
            for (int k = 0; k != 4; ++k) {
                Idx const r = side[k / 2], c = side[k % 2];
                if (r != -1 && c != -1) items.push_back({r, c, static_cast<int8_t>(k), b});
            }
        }
        for (Idx bus = 0; bus != n_bus; ++bus)
            for (Idx s = topo.shunts_per_bus[bus]; s != topo.shunts_per_bus[bus + 1]; ++s) items.push_back({bus, bus, 4, s});
        for (Idx f = 0; f != topo.n_fill(); ++f) {
            Idx const i = topo.fill_in[2 * f], j = topo.fill_in[2 * f + 1];
            if (i != -1 && j != -1) {
                items.push_back({i, j, 5, f});
                items.push_back({j, i, 5, f});
            }
        }
        // the reference's two-pass stable counting sort == one stable sort on (row, col): contributions to one entry
        // keep their insertion order, which fixes the floating-point summation order of the admittance
        std::stable_sort(items.begin(), items.end(),
                         [](Item const& x, Item const& y) { return x.row != y.row ? x.row < y.row : x.col < y.col; });

        row_indptr.assign(n_bus + 1, 0);
        row_indptr_lu.assign(n_bus + 1, 0);
        bus_entry.assign(n_bus, -1);
        diag_lu.assign(n_bus, -1);
        y_bus_entry_indptr.push_back(0);
        if (items.empty()) { // single bus without branch or shunt: one artificial diagonal entry
            if (n_bus != 1) throw std::invalid_argument("math topology has a bus without any admittance entry");
            row_indptr = {0, 1};
            col_indices = {0};
            bus_entry = {0};
            y_bus_entry_indptr = {0, 0};
            row_indptr_lu = {0, 1};
            col_indices_lu = {0};
            diag_lu = {0};
            map_lu_y_bus = {0};
            lu_transpose_entry = {0};
            nnz = nnz_lu = 1;
            return;
        }
        for (size_t i = 0; i != items.size();) {
            Idx const r = items[i].row, c = items[i].col;
            bool const fill = items[i].kind == 5;
            size_t j = i;
            while (j != items.size() && items[j].row == r && items[j].col == c) ++j;
            col_indices_lu.push_back(c);
            ++row_indptr_lu[r + 1];
            if (fill) {
                if (j - i != 1) throw std::invalid_argument("fill-in duplicates an existing Y-bus entry");
                map_lu_y_bus.push_back(-1);
            } else {
                map_lu_y_bus.push_back(static_cast<Idx>(col_indices.size()));
                if (r == c) {
                    bus_entry[r] = static_cast<Idx>(col_indices.size());
                    diag_lu[r] = static_cast<Idx>(col_indices_lu.size()) - 1;
                }
                col_indices.push_back(c);
                ++row_indptr[r + 1];
                for (size_t e = i; e != j; ++e) {
                    element_type.push_back(items[e].kind);
                    element_idx.push_back(items[e].idx);
                }
                y_bus_entry_indptr.push_back(static_cast<Idx>(element_type.size()));
            }
            i = j;
        }
        for (Idx r = 0; r != n_bus; ++r) {
            row_indptr[r + 1] += row_indptr[r];
            row_indptr_lu[r + 1] += row_indptr_lu[r];
            if (diag_lu[r] < 0) throw std::invalid_argument("math topology has a bus without a diagonal entry");
        }
        nnz = row_indptr.back();
        nnz_lu = row_indptr_lu.back();
        lu_transpose_entry.resize(nnz_lu);
        for (Idx r = 0; r != n_bus; ++r)
            for (Idx k = row_indptr_lu[r]; k != row_indptr_lu[r + 1]; ++k) lu_transpose_entry[k] = find_lu(col_indices_lu[k], r);
    }

    Idx find_lu(Idx row, Idx col) const {
        auto const first = col_indices_lu.begin() + row_indptr_lu[row];
        auto const last = col_indices_lu.begin() + row_indptr_lu[row + 1];
        auto const it = std::lower_bound(first, last, col);
        if (it == last || *it != col) {
            throw std::invalid_argument("LU pattern is not structurally symmetric / closed under fill-in at (" +
                                        std::to_string(row) + ", " + std::to_string(col) + ")");
        }
        return static_cast<Idx>(it - col_indices_lu.begin());
    }
};

struct EliminationSchedule {
    // for LU entry e: pairs [upd_ptr[e], upd_ptr[e+1]) -- non-empty only for strictly-lower entries (k, c):
    //   upd_u = entry (c, j), upd_a = entry (k, j), for every j > c in row c, ascending j
    std::vector<int32_t> upd_ptr, upd_u, upd_a;
    // rows grouped by dependency level, ascending row inside a level
    std::vector<int32_t> level_ptr, level_rows;
    std::vector<int32_t> row_level;
    int32_t max_row_entries{};
    int64_t n_block_updates{};

    explicit EliminationSchedule(LuPattern const& p) {
        Idx const n = p.n_bus;
        upd_ptr.assign(p.nnz_lu + 1, 0);
        row_level.assign(n, 0);
        for (Idx k = 0; k != n; ++k) {
            max_row_entries = std::max<int32_t>(max_row_entries, static_cast<int32_t>(p.row_indptr_lu[k + 1] - p.row_indptr_lu[k]));
            int32_t level = 0;
            for (Idx e = p.row_indptr_lu[k]; e != p.diag_lu[k]; ++e) {
                Idx const c = p.col_indices_lu[e];
                level = std::max(level, row_level[c] + 1);
                for (Idx ue = p.diag_lu[c] + 1; ue != p.row_indptr_lu[c + 1]; ++ue) {
                    upd_u.push_back(static_cast<int32_t>(ue));
                    upd_a.push_back(static_cast<int32_t>(p.find_lu(k, p.col_indices_lu[ue])));
                }
                upd_ptr[e + 1] = static_cast<int32_t>(upd_u.size());
            }
            // entries from the diagonal on carry no updates
            for (Idx e = p.diag_lu[k]; e != p.row_indptr_lu[k + 1]; ++e) upd_ptr[e + 1] = static_cast<int32_t>(upd_u.size());
            row_level[k] = level;
        }
        n_block_updates = static_cast<int64_t>(upd_u.size());
        int32_t const n_level = n == 0 ? 0 : *std::max_element(row_level.begin(), row_level.end()) + 1;
        level_ptr.assign(n_level + 1, 0);
        for (Idx k = 0; k != n; ++k) ++level_ptr[row_level[k] + 1];
        for (int32_t l = 0; l != n_level; ++l) level_ptr[l + 1] += level_ptr[l];
        level_rows.resize(n);
        std::vector<int32_t> cursor(level_ptr.begin(), level_ptr.end() - 1);
        for (Idx k = 0; k != n; ++k) level_rows[cursor[row_level[k]]++] = static_cast<int32_t>(k);
    }
    int32_t n_level() const { return static_cast<int32_t>(level_ptr.size()) - 1; }
};

// Row programs: the per-row index data of both sweeps packed into one int32 blob in execution (level) order, so that a
// thread block can stage it in shared memory with one bulk copy (TMA) and never chase global index arrays.
//   blob = [level_ptr (n_level + 1)] [task_off (n_bus)] [records ...]        (all offsets in words from blob start)
//   record (6 words header): row | k_diag | ky_diag | n_lower + (n_upper << 12) + (tree << 24) | lg_start + (n_lg << 24)
//                            | src_start + (n_src << 24)
//     tree rows only:  n_lower x (c, ky, k_diag_of_c, k_of_U(c,row))  then  n_upper x (k, j, ky)
// A "tree row" is a row whose Schur updates all land on its own diagonal block (every lower neighbour c has the single
// upper entry (c, row)) and that has at most one upper entry: all rows of a radial grid, and the tree part of a meshed one.
// Such a row is processed entirely in registers; other rows use the generic path on the global index arrays.
struct RowProgram {
    std::vector<int32_t> words;
    int32_t n_tree_rows{};

    RowProgram(LuPattern const& p, EliminationSchedule const& sch, MathTopology const& topo) {
        Idx const n = p.n_bus;
        int32_t const n_level = sch.n_level();
        words.assign(static_cast<size_t>(n_level) + 1 + n, 0);
        for (int32_t l = 0; l <= n_level; ++l) words[l] = sch.level_ptr[l];
        for (Idx i = 0; i != n; ++i) {
            Idx const row = sch.level_rows[i];
            words[n_level + 1 + i] = static_cast<int32_t>(words.size());
            Idx const rb = p.row_indptr_lu[row], re = p.row_indptr_lu[row + 1], dg = p.diag_lu[row];
            Idx const n_lower = dg - rb, n_upper = re - dg - 1;
            Idx const lg0 = topo.load_gens_per_bus[row], n_lg = topo.load_gens_per_bus[row + 1] - lg0;
            Idx const s0 = topo.sources_per_bus[row], n_src = topo.sources_per_bus[row + 1] - s0;
            bool tree = n_upper <= 1 && n_lower < 4096 && n_lg < 128 && n_src < 128 && lg0 < (1 << 24) && s0 < (1 << 24);
            for (Idx e = rb; e != dg && tree; ++e) {
                tree = sch.upd_ptr[e + 1] - sch.upd_ptr[e] == 1 && sch.upd_a[sch.upd_ptr[e]] == dg;
            }
            words.push_back(static_cast<int32_t>(row));
            words.push_back(static_cast<int32_t>(dg));
            words.push_back(static_cast<int32_t>(p.map_lu_y_bus[dg]));
            words.push_back(static_cast<int32_t>((tree ? n_lower : 0) | ((tree ? n_upper : 0) << 12) | ((tree ? 1 : 0) << 24)));
            words.push_back(static_cast<int32_t>(tree ? (lg0 | (n_lg << 24)) : 0));
            words.push_back(static_cast<int32_t>(tree ? (s0 | (n_src << 24)) : 0));
            if (!tree) continue;
            ++n_tree_rows;
            for (Idx e = rb; e != dg; ++e) {
                Idx const c = p.col_indices_lu[e];
                words.push_back(static_cast<int32_t>(c));
                words.push_back(static_cast<int32_t>(p.map_lu_y_bus[e]));
                words.push_back(static_cast<int32_t>(p.diag_lu[c]));
                words.push_back(sch.upd_u[sch.upd_ptr[e]]);
            }
            for (Idx e = dg + 1; e != re; ++e) {
                words.push_back(static_cast<int32_t>(e));
                words.push_back(static_cast<int32_t>(p.col_indices_lu[e]));
                words.push_back(static_cast<int32_t>(p.map_lu_y_bus[e]));
            }
        }
        while (words.size() % 4 != 0) words.push_back(0); // 16-byte granularity for the bulk copy
    }
};

// Wide rows: a row with many lower entries (a hub of a meshed grid: the ring-closing / feeding nodes collect hundreds of
// children) is a long sequential elimination when one thread owns the row.  For such rows the children are eliminated by all
// threads of the block together: the L block of every child and its update terms are computed in parallel, and each target
// entry then subtracts its terms in ascending child order -- the same operations on the same numbers as the row-by-row loop.
// Children whose own block receives terms from earlier children (fill-in inside the row) are ordered in sub-levels.
//   table (8 words per wide row): row, n_sub, off_sub_ptr, off_order, off_in_ptr, off_in_idx, n_upd, 0   (offsets into data)
//   order:   lower-entry positions (entry - row begin) sorted by sub-level; sub_ptr[n_sub + 1] indexes it
//   in_ptr[n_entries + 1], in_idx: for every entry of the row the indices (update number - first update of the row) of the
//   terms it receives, ascending (= ascending child)
struct WideRowPlan {
    static constexpr Idx min_lower = 24;
    std::vector<int32_t> level_ptr;  // wide rows per dependency level: [n_level + 1] into table rows
    std::vector<int32_t> table;      // 8 words per wide row
    std::vector<int32_t> data;
    std::vector<uint8_t> is_wide;    // per row
    int32_t max_upd{}, max_lower{}, max_entries{};

    WideRowPlan(LuPattern const& p, EliminationSchedule const& sch) {
        Idx const n = p.n_bus;
        int32_t const n_level = sch.n_level();
        is_wide.assign(n, 0);
        level_ptr.assign(n_level + 1, 0);
        for (int32_t lv = 0; lv != n_level; ++lv) {
            int32_t const n_rows = sch.level_ptr[lv + 1] - sch.level_ptr[lv];
            for (int32_t i = sch.level_ptr[lv]; i != sch.level_ptr[lv + 1]; ++i) {
                Idx const row = sch.level_rows[i];
                Idx const rb = p.row_indptr_lu[row], re = p.row_indptr_lu[row + 1], dg = p.diag_lu[row];
                Idx const n_lower = dg - rb;
                bool const wide = n_lower >= min_lower && (n_rows <= 8 || n_lower >= 128);
                if (!wide) continue;
                is_wide[row] = 1;
                int32_t const upd_base = sch.upd_ptr[rb];
                int32_t const n_upd = sch.upd_ptr[dg] - upd_base;
                Idx const n_entries = re - rb;
                // incoming terms per entry
                std::vector<std::vector<int32_t>> incoming(n_entries);
                std::vector<int32_t> src_child(n_upd);
                for (Idx e = rb; e != dg; ++e)
                    for (int32_t q = sch.upd_ptr[e]; q != sch.upd_ptr[e + 1]; ++q) {
                        incoming[sch.upd_a[q] - rb].push_back(q - upd_base);
                        src_child[q - upd_base] = static_cast<int32_t>(e - rb);
                    }
                // sub-levels of the lower entries
                std::vector<int32_t> depth(n_lower, 0);
                int32_t n_sub = 1;
                for (Idx j = 0; j != n_lower; ++j) {
                    for (int32_t q : incoming[j]) depth[j] = std::max(depth[j], depth[src_child[q]] + 1);
                    n_sub = std::max(n_sub, depth[j] + 1);
                }
                std::vector<int32_t> sub_ptr(n_sub + 1, 0), order(n_lower);
                for (Idx j = 0; j != n_lower; ++j) ++sub_ptr[depth[j] + 1];
                for (int32_t d = 0; d != n_sub; ++d) sub_ptr[d + 1] += sub_ptr[d];
                {
                    std::vector<int32_t> cursor(sub_ptr.begin(), sub_ptr.end() - 1);
                    for (Idx j = 0; j != n_lower; ++j) order[cursor[depth[j]]++] = static_cast<int32_t>(j);
                }
                table.push_back(static_cast<int32_t>(row));
                table.push_back(n_sub);
                table.push_back(static_cast<int32_t>(data.size()));
                data.insert(data.end(), sub_ptr.begin(), sub_ptr.end());
                table.push_back(static_cast<int32_t>(data.size()));
                data.insert(data.end(), order.begin(), order.end());
                table.push_back(static_cast<int32_t>(data.size()));
                int32_t run = 0;
                for (Idx k = 0; k != n_entries; ++k) {
                    data.push_back(run);
                    run += static_cast<int32_t>(incoming[k].size());
                }
                data.push_back(run);
                table.push_back(static_cast<int32_t>(data.size()));
                for (Idx k = 0; k != n_entries; ++k) data.insert(data.end(), incoming[k].begin(), incoming[k].end());
                table.push_back(n_upd);
                table.push_back(0);
                max_upd = std::max(max_upd, n_upd);
                max_lower = std::max<int32_t>(max_lower, static_cast<int32_t>(n_lower));
                max_entries = std::max<int32_t>(max_entries, static_cast<int32_t>(n_entries));
                ++level_ptr[lv + 1];
            }
        }
        for (int32_t lv = 0; lv != n_level; ++lv) level_ptr[lv + 1] += level_ptr[lv];
        if (data.empty()) data.push_back(0);
    }
    int32_t n_wide() const { return static_cast<int32_t>(table.size() / 8); }
};

// Path programs (radial grids: every row is a tree row).  The elimination tree is cut into PATHS: maximal chains
// child -> parent in which the parent continues the chain of exactly one non-leaf child (its "carry" child).  One thread
// walks a whole path with the carried child's factor, U block, permutation and right-hand side in registers, so the
// dependent chain of a feeder needs neither block barriers nor global round trips between consecutive rows.  Paths are
// grouped in STAGES: a path may start once every other (non-carried) child of its rows is complete; leaves are stage 0.
//   blob = header[8] | leaf records (8 words) | rec_off[n_rec] | stage_ptr[n_stage + 1] | paths (first_rec, n_rows) | records
//   header (12 words): n_leaf, n_rec, n_stage, off_leaf, off_rec_off, off_stage_ptr, off_path, n_path, off_chain,
//                      smem_words (prefix staged in shared memory), 0, 0
//   chain record (8 words per non-leaf row, same index as rec_off): row, k_d, k_u, k_a (block towards the carry child or
//                 -1), k_s (precomputed leaf term or -1), pattern, record offset, parent row
//   pattern: 1 nothing left to eliminate in the chain, 2 carry child only, 3 carry child then the precomputed leaf term,
//            0 anything else (generic loop over the row record)
//   leaf record : row, k_d, ky_d, k_u, j, ky_u, lg0 | n_lg << 24, s0 | n_src << 24            (k_u = -1: no parent)
//   row record  : same 8 words, then n_lower | carry_e << 12 | post_e << 20 (0xff = none), then per lower entry
//                 c, ky, k_diag(c), k_U(c,row) | kind << 28
//   kind: 0 leaf child eliminated when the row is built, 1 carry child (registers), 2 leaf child whose update term is
//         precomputed when the row is built and subtracted in the chain (keeps the reference's order of subtractions),
//         3 any other child: eliminated in the chain from global memory.
// Records of a path are consecutive, bottom row first; stage s (>= 1) owns paths [stage_ptr[s-1], stage_ptr[s]).
struct PathProgram {
    std::vector<int32_t> words;
    bool valid{false};
    int32_t n_stage{};
    int32_t n_path{};
    int32_t max_path_rows{};
    int32_t smem_words{}; // header + stages + paths + chain records

    PathProgram(LuPattern const& p, EliminationSchedule const& sch, MathTopology const& topo, RowProgram const& rows) {
        Idx const n = p.n_bus;
        if (n == 0 || rows.n_tree_rows != n || p.nnz_lu >= (Idx{1} << 28)) return;
        auto n_lower_of = [&](Idx r) { return p.diag_lu[r] - p.row_indptr_lu[r]; };
        std::vector<int32_t> stage(n, 0), path_of(n, -1);
        std::vector<std::vector<Idx>> paths;   // rows bottom-up
        std::vector<int32_t> path_stage;
        std::vector<Idx> carry(n, -1);
        for (Idx r = 0; r != n; ++r) {
            if (n_lower_of(r) == 0) continue;
            Idx cc = -1;
            for (Idx e = p.row_indptr_lu[r]; e != p.diag_lu[r]; ++e) {
                Idx const c = p.col_indices_lu[e];
                if (n_lower_of(c) == 0) continue;
                if (cc < 0 || stage[c] > stage[cc] || (stage[c] == stage[cc] && sch.row_level[c] > sch.row_level[cc])) cc = c;
            }
            int32_t s_side = 1;
            for (Idx e = p.row_indptr_lu[r]; e != p.diag_lu[r]; ++e) {
                Idx const c = p.col_indices_lu[e];
                if (n_lower_of(c) == 0 || c == cc) continue;
                s_side = std::max(s_side, stage[c] + 1);
            }
            Idx const carry_e = cc < 0 ? -1 : (std::lower_bound(p.col_indices_lu.begin() + p.row_indptr_lu[r],
                                                                 p.col_indices_lu.begin() + p.diag_lu[r], cc) -
                                               (p.col_indices_lu.begin() + p.row_indptr_lu[r]));
            if (cc >= 0 && s_side <= stage[cc] && carry_e < 255) {
                stage[r] = stage[cc];
                carry[r] = cc;
                path_of[r] = path_of[cc];
                paths[path_of[r]].push_back(r);
            } else {
                stage[r] = cc >= 0 ? std::max(s_side, stage[cc] + 1) : s_side;
                path_of[r] = static_cast<int32_t>(paths.size());
                paths.push_back({r});
                path_stage.push_back(stage[r]);
            }
        }
        n_stage = 1;
        for (int32_t s : path_stage) n_stage = std::max(n_stage, s + 1);
        n_path = static_cast<int32_t>(paths.size());
        // order paths by stage, longest first inside a stage (the longest chains start first on the slots)
        std::vector<int32_t> order(paths.size());
        for (size_t i = 0; i != order.size(); ++i) order[i] = static_cast<int32_t>(i);
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            if (path_stage[a] != path_stage[b]) return path_stage[a] < path_stage[b];
            return paths[a].size() > paths[b].size();
        });
        std::vector<Idx> leaves;
        for (Idx r = 0; r != n; ++r)
            if (n_lower_of(r) == 0) leaves.push_back(r);
        Idx const n_rec = n - static_cast<Idx>(leaves.size());

        auto head = [&](Idx row) {
            Idx const dg = p.diag_lu[row], re = p.row_indptr_lu[row + 1];
            Idx const lg0 = topo.load_gens_per_bus[row], n_lg = topo.load_gens_per_bus[row + 1] - lg0;
            Idx const s0 = topo.sources_per_bus[row], n_src = topo.sources_per_bus[row + 1] - s0;
            bool const has_u = re - dg - 1 == 1;
            words.push_back(static_cast<int32_t>(row));
            words.push_back(static_cast<int32_t>(dg));
            words.push_back(static_cast<int32_t>(p.map_lu_y_bus[dg]));
            words.push_back(has_u ? static_cast<int32_t>(dg + 1) : -1);
            words.push_back(has_u ? static_cast<int32_t>(p.col_indices_lu[dg + 1]) : -1);
            words.push_back(has_u ? static_cast<int32_t>(p.map_lu_y_bus[dg + 1]) : -1);
            words.push_back(static_cast<int32_t>(lg0 | (n_lg << 24)));
            words.push_back(static_cast<int32_t>(s0 | (n_src << 24)));
        };
        // the latency-critical part (header, stages, paths, chain records) comes first: only this prefix is staged in shared
        // memory, the records of the parallel build passes are read through L1
        words.assign(12, 0);
        Idx const off_stage_ptr = static_cast<Idx>(words.size());
        words.resize(words.size() + n_stage + 1, 0);
        Idx const off_path = static_cast<Idx>(words.size());
        words.resize(words.size() + 2 * paths.size(), 0);
        while (words.size() % 4 != 0) words.push_back(0); // chain records are read as 16-byte vectors
        Idx const off_chain = static_cast<Idx>(words.size());
        words.resize(words.size() + 8 * n_rec, 0);
        Idx const smem_words = static_cast<Idx>(words.size());
        Idx const off_leaf = static_cast<Idx>(words.size());
        for (Idx r : leaves) head(r);
        Idx const off_rec_off = static_cast<Idx>(words.size());
        words.resize(words.size() + n_rec, 0);
        Idx rec_idx = 0;
        for (size_t oi = 0; oi != order.size(); ++oi) {
            auto const& rows_of_path = paths[order[oi]];
            words[off_stage_ptr + path_stage[order[oi]]] += 1; // counts per stage (stage >= 1), prefix-summed below
            words[off_path + 2 * oi] = static_cast<int32_t>(rec_idx);
            words[off_path + 2 * oi + 1] = static_cast<int32_t>(rows_of_path.size());
            max_path_rows = std::max<int32_t>(max_path_rows, static_cast<int32_t>(rows_of_path.size()));
            for (Idx r : rows_of_path) {
                Idx const this_rec = rec_idx;
                words[off_rec_off + rec_idx++] = static_cast<int32_t>(words.size());
                int32_t const this_off = static_cast<int32_t>(words.size());
                head(r);
                Idx const rb = p.row_indptr_lu[r], dg = p.diag_lu[r];
                size_t const w_n = words.size();
                words.push_back(0);
                int32_t carry_e = 0xff, post_e = 0xff;
                bool seen_nonleaf = false;
                for (Idx e = rb; e != dg; ++e) {
                    Idx const c = p.col_indices_lu[e];
                    int32_t kind;
                    if (n_lower_of(c) == 0) {
                        if (!seen_nonleaf) {
                            kind = 0;
                        } else if (post_e == 0xff && e - rb < 255) {
                            kind = 2;
                            post_e = static_cast<int32_t>(e - rb);
                        } else {
                            kind = 3;
                        }
                    } else {
                        seen_nonleaf = true;
                        if (c == carry[r]) {
                            kind = 1;
                            carry_e = static_cast<int32_t>(e - rb);
                        } else {
                            kind = 3;
                        }
                    }
                    words.push_back(static_cast<int32_t>(c));
                    words.push_back(static_cast<int32_t>(p.map_lu_y_bus[e]));
                    words.push_back(static_cast<int32_t>(p.diag_lu[c]));
                    words.push_back(static_cast<int32_t>(sch.upd_u[sch.upd_ptr[e]] | (kind << 28)));
                }
                words[w_n] = static_cast<int32_t>((dg - rb) | (carry_e << 12) | (post_e << 20));
                // chain record
                std::vector<int32_t> seq; // kinds left for the chain, in entry order
                for (Idx e = rb; e != dg; ++e) {
                    int32_t const kind = (words[w_n + 1 + 4 * (e - rb) + 3] >> 28) & 3;
                    if (kind != 0) seq.push_back(kind);
                }
                int32_t pattern = 0;
                if (seq.empty()) pattern = 1;
                else if (seq.size() == 1 && seq[0] == 1) pattern = 2;
                else if (seq.size() == 2 && seq[0] == 1 && seq[1] == 2) pattern = 3;
                Idx const re = p.row_indptr_lu[r + 1];
                bool const has_u = re - dg - 1 == 1;
                int32_t* cr = &words[off_chain + 8 * this_rec];
                cr[0] = static_cast<int32_t>(r);
                cr[1] = static_cast<int32_t>(dg);
                cr[2] = has_u ? static_cast<int32_t>(dg + 1) : -1;
                cr[3] = carry_e != 0xff ? static_cast<int32_t>(rb + carry_e) : -1;
                cr[4] = post_e != 0xff ? static_cast<int32_t>(rb + post_e) : -1;
                cr[5] = pattern;
                cr[6] = this_off;
                cr[7] = has_u ? static_cast<int32_t>(p.col_indices_lu[dg + 1]) : -1;
            }
        }
        // stage_ptr: stage s (>= 1) owns paths [stage_ptr[s - 1], stage_ptr[s]) ; entry 0 counts nothing
        {
            int32_t run = 0;
            for (int32_t s = 0; s <= n_stage; ++s) {
                run += words[off_stage_ptr + s];
                words[off_stage_ptr + s] = run;
            }
        }
        words[0] = static_cast<int32_t>(leaves.size());
        words[1] = static_cast<int32_t>(n_rec);
        words[2] = n_stage;
        words[3] = static_cast<int32_t>(off_leaf);
        words[4] = static_cast<int32_t>(off_rec_off);
        words[5] = static_cast<int32_t>(off_stage_ptr);
        words[6] = static_cast<int32_t>(off_path);
        words[7] = n_path;
        words[8] = static_cast<int32_t>(off_chain);
        words[9] = static_cast<int32_t>(smem_words);
        this->smem_words = static_cast<int32_t>(smem_words);
        while (words.size() % 4 != 0) words.push_back(0);
        valid = true;
    }
};

} // namespace pgmb
