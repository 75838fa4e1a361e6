"""Host-side symbolic stage of the product (LuPattern + elimination schedule in libpgm_b200.so) against the oracle and
the reference's golden index vectors -- bit-exact integer parity.  Needs no GPU (engine created with device = -1)."""
import numpy as np
import pytest

import oracle_lib as orc
import pgm_b200
from grids import random_grid, three_bus_grid

NAMES = ("row_indptr", "col_indices", "bus_entry", "row_indptr_lu", "col_indices_lu", "diag_lu", "map_lu_y_bus",
         "lu_transpose_entry", "y_bus_entry_indptr")


def _engine(n_bus, branch_bus_idx, fill_in, shunts_indptr):
    z = np.zeros(n_bus + 1, np.int64)
    return pgm_b200.Engine(symmetric=True, phase_shift=np.zeros(n_bus), branch_bus_idx=branch_bus_idx, sources_per_bus=z,
                           shunts_per_bus=shunts_indptr, load_gens_per_bus=z, load_gen_type=[], fill_in=fill_in, device=-1)


@pytest.mark.parametrize("case", [
    (4, [[1, 0], [1, 2], [2, 3], [2, 2], [3, 2], [0, 1], [2, -1]], [], [0, 1, 1, 1, 2]),   # test_y_bus.cpp:60-157
    (3, [[0, 1], [1, 2], [2, 0], [1, 1], [2, 2]], [], [0, 0, 0, 0]),                         # test_y_bus.cpp:276-376
    (3, [[1, 0], [0, 2]], [[1, 2]], [0, 0, 0, 0]),                                           # test_y_bus.cpp:378-425
    (1, [], [], [0, 0]), (1, [[0, 0]], [], [0, 0]),                                          # test_y_bus.cpp:231-274
])
def test_pattern_matches_oracle_on_reference_cases(case):
    n_bus, bb, fill, sh = case
    eng = _engine(n_bus, bb, fill, sh)
    bag = orc.ybus_structure(n_bus, bb, fill, sh)
    for name in NAMES:
        assert eng.index(name).tolist() == bag.i64(name).tolist(), name


def test_golden_fill_in_pattern():
    eng = _engine(3, [[1, 0], [0, 2]], [[1, 2]], [0, 0, 0, 0])
    assert eng.index("row_indptr_lu").tolist() == [0, 3, 6, 9]
    assert eng.index("map_lu_y_bus").tolist() == [0, 1, 2, 3, 4, -1, 5, -1, 6]
    assert eng.index("diag_lu").tolist() == [0, 4, 8]
    assert eng.index("lu_transpose_entry").tolist() == [0, 3, 6, 1, 4, 7, 2, 5, 8]


@pytest.mark.parametrize("n_node,extra,seed", [(30, 0, 1), (200, 0, 2), (60, 15, 3), (300, 40, 4)])
def test_pattern_and_admittance_match_oracle_on_random_grids(n_node, extra, seed):
    grid = random_grid(n_node, extra, seed)
    eng = pgm_b200.Engine.from_grid(grid, device=-1)
    out = orc.math_pf(grid, "linear")
    assert out["status"] == 0, out["error"]
    for name in ("row_indptr", "col_indices", "row_indptr_lu", "col_indices_lu", "diag_lu", "map_lu_y_bus"):
        assert eng.index(name).tolist() == out["bag"].i64(name).tolist(), name
    # admittance: same contributions summed in the same order => bit-identical
    assert np.array_equal(eng.admittance().ravel(), out["bag"].c128("admittance"))


def test_schedule_levels_respect_dependencies():
    grid = random_grid(300, 40, 5)
    eng = pgm_b200.Engine.from_grid(grid, device=-1)
    indptr, indices, diag = eng.index("row_indptr_lu"), eng.index("col_indices_lu"), eng.index("diag_lu")
    level_ptr, level_rows = eng.index("level_ptr"), eng.index("level_rows")
    level_of = np.empty(len(diag), np.int64)
    for lv in range(len(level_ptr) - 1):
        level_of[level_rows[level_ptr[lv]:level_ptr[lv + 1]]] = lv
    assert sorted(level_rows.tolist()) == list(range(len(diag)))
    for k in range(len(diag)):
        for e in range(indptr[k], diag[k]):
            assert level_of[indices[e]] < level_of[k]


def test_three_bus_structure():
    grid, _ = three_bus_grid(True)
    eng = pgm_b200.Engine.from_grid(grid, device=-1)
    assert eng.index("row_indptr_lu").tolist() == [0, 2, 5, 7]
    assert eng.index("level_ptr").tolist() == [0, 1, 2, 3]


@pytest.mark.parametrize("n_node,seed", [(2, 1), (40, 2), (300, 3), (1500, 4)])
def test_path_program_covers_a_radial_grid(n_node, seed):
    """PathProgram (symbolic.hpp): every non-leaf row sits in exactly one path, a path is a child -> parent chain, and every
    child that is not carried in registers belongs to an earlier stage (or is a leaf)."""
    grid = random_grid(n_node, 0, seed)
    eng = pgm_b200.Engine.from_grid(grid, device=-1)
    w = eng.index("path_program")
    assert w.size > 0
    rp, ci, dg = eng.index("row_indptr_lu"), eng.index("col_indices_lu"), eng.index("diag_lu")
    n = len(rp) - 1
    n_leaf, n_rec, n_stage, off_leaf, off_ro, off_sp, off_path, n_path, off_chain = (int(x) for x in w[:9])
    leaves = [int(w[off_leaf + 8 * i]) for i in range(n_leaf)]
    assert sorted(leaves) == [r for r in range(n) if dg[r] == rp[r]]
    sp = w[off_sp:off_sp + n_stage + 1]
    assert sp[0] == 0 and sp[n_stage] == n_path
    stage_of = {r: 0 for r in leaves}
    seen = set()
    for st in range(1, n_stage + 1):
        for p in range(sp[st - 1], sp[st]):
            first, n_rows = int(w[off_path + 2 * p]), int(w[off_path + 2 * p + 1])
            rows = [int(w[off_chain + 8 * (first + i)]) for i in range(n_rows)]
            for i, r in enumerate(rows):
                assert r not in seen
                seen.add(r)
                stage_of[r] = st
                cr = w[off_chain + 8 * (first + i):off_chain + 8 * (first + i) + 8]
                assert cr[1] == dg[r]
                parent = int(ci[dg[r] + 1]) if rp[r + 1] - dg[r] == 2 else -1
                assert cr[7] == parent and (i + 1 == n_rows or rows[i + 1] == parent)
                rec = int(w[off_ro + first + i])
                assert rec == cr[6] and w[rec] == r
                n_lower = int(w[rec + 8]) & 0xfff
                assert n_lower == dg[r] - rp[r]
                for e in range(n_lower):
                    c, kind = int(w[rec + 9 + 4 * e]), (int(w[rec + 9 + 4 * e + 3]) >> 28) & 3
                    assert c == ci[rp[r] + e]
                    if kind == 1:
                        assert i > 0 and rows[i - 1] == c and cr[3] == rp[r] + e
                    elif kind in (0, 2):
                        assert stage_of[c] == 0
                    else:
                        assert stage_of[c] < st
    assert len(seen) == n_rec == n - n_leaf


def test_path_program_absent_on_meshed_grid():
    grid = random_grid(60, 10, 5)
    assert pgm_b200.Engine.from_grid(grid, device=-1).index("path_program").size == 0


def test_wide_row_plan_of_a_hub_grid():
    """WideRowPlan (symbolic.hpp): a star-of-rings grid has a hub row with many lower entries; its children are ordered in
    sub-levels that respect the updates between them, and every entry lists the terms it receives in ascending child order"""
    rng = np.random.default_rng(5)
    n_spoke, ring = 30, 4
    edges, n = [], 1
    for s in range(n_spoke):  # hub 0 -- a -- b -- c -- d -- hub: rings through the hub
        nodes = list(range(n, n + ring))
        n += ring
        edges += [(0, nodes[0])] + [(nodes[i], nodes[i + 1]) for i in range(ring - 1)] + [(nodes[-1], 0)]
    bag = orc.topology(n, edges, [[1, 1]] * len(edges), [0.0] * len(edges), [0], [1])
    assert bag.i64("n_math")[0] == 1
    bb = bag.i64("g0.branch_bus_idx").reshape(-1, 2)
    fill = bag.i64("g0.fill_in").reshape(-1, 2)
    eng = _engine(n, bb, fill, np.zeros(n + 1, np.int64))
    table = eng.index("wide_table").reshape(-1, 8)
    assert len(table) >= 1
    data, rp, ci, dg = eng.index("wide_data"), eng.index("row_indptr_lu"), eng.index("col_indices_lu"), eng.index("diag_lu")
    upd_ptr, upd_a = eng.index("upd_ptr"), eng.index("upd_a")
    for row, n_sub, off_sub, off_order, off_in_ptr, off_in_idx, n_upd, _ in table:
        rb, re_, d = rp[row], rp[row + 1], dg[row]
        n_lower, n_entries = d - rb, re_ - rb
        assert n_lower >= 24 and n_upd == upd_ptr[d] - upd_ptr[rb]
        sub_ptr = data[off_sub:off_sub + n_sub + 1]
        order = data[off_order:off_order + n_lower]
        assert sorted(order.tolist()) == list(range(n_lower)) and sub_ptr[0] == 0 and sub_ptr[-1] == n_lower
        depth = np.empty(n_lower, int)
        for sl in range(n_sub):
            depth[order[sub_ptr[sl]:sub_ptr[sl + 1]]] = sl
        in_ptr = data[off_in_ptr:off_in_ptr + n_entries + 1]
        in_idx = data[off_in_idx:off_in_idx + in_ptr[-1]]
        assert in_ptr[-1] == n_upd
        src_child = np.empty(n_upd, int)
        for e in range(rb, d):
            src_child[upd_ptr[e] - upd_ptr[rb]:upd_ptr[e + 1] - upd_ptr[rb]] = e - rb
        for pos in range(n_entries):
            terms = in_idx[in_ptr[pos]:in_ptr[pos + 1]]
            assert (np.diff(terms) > 0).all()                       # ascending update number = ascending child
            assert all(upd_a[upd_ptr[rb] + q] == rb + pos for q in terms)
            if pos < n_lower:                                        # a child only receives terms from earlier sub-levels
                assert all(depth[src_child[q]] < depth[pos] for q in terms)
