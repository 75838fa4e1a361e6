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
