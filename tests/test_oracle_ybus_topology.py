"""Pins the oracle's symbolic stage against the reference's golden index vectors:
tests/cpp_unit_tests/math_solver/test_y_bus.cpp:60-157 (4-bus), :276-376 (cyclic 3-bus), :378-425 (fill-in),
tests/cpp_unit_tests/test_topology.cpp:125-378 (14-node grid with 3 math models), :380-471 (cycle reorder),
tests/cpp_unit_tests/test_sparse_ordering.cpp:21-36."""
import numpy as np

import oracle_lib as orc
from grids import three_bus_grid


def L(x):
    return np.asarray(x).tolist()


def test_ybus_four_bus_structure_and_admittance():
    branch_bus_idx = [[1, 0], [1, 2], [2, 3], [2, 2], [3, 2], [0, 1], [2, -1]]
    bag = orc.ybus_structure(4, branch_bus_idx, [], [0, 1, 1, 1, 2])
    assert L(bag.i64("row_indptr")) == [0, 2, 5, 8, 10]
    assert L(bag.i64("col_indices")) == [0, 1, 0, 1, 2, 1, 2, 3, 2, 3]
    assert L(bag.i64("bus_entry")) == [0, 3, 6, 9]
    assert L(bag.i64("lu_transpose_entry")) == [0, 2, 1, 3, 5, 4, 6, 8, 7, 9]
    assert L(bag.i64("y_bus_entry_indptr")) == [0, 3, 5, 7, 10, 11, 12, 20, 22, 24, 27]
    assert L(bag.i64("map_lu_y_bus")) == list(range(10))
    # admittance sums through the math-level entry (test_y_bus.cpp:117-133)
    bp = np.array([[1j, 2j, 3j, 4j], [5, 6, 7, 8], [9j, 10j, 11j, 12j], [21j, 22j, 22j, 21j], [13, 14, 15, 16],
                   [17, 18, 19, 20], [1000j, 0, 0, 0]], dtype=complex)
    grid = orc.MathGrid(True, [0.0] * 4, branch_bus_idx, [0, 1, 1, 1, 1], [0, 1, 1, 1, 2], [0, 0, 0, 0, 0], [],
                        bp, [100j, 200j], [[1 - 1j, 1 - 1j]], [1.0], [])
    out = orc.math_pf(grid, "linear")
    adm = out["bag"].c128("admittance")
    expected = [4j + 17 + 100j, 18 + 3j, 19 + 2j, 20 + 1j + 5, 6, 7, 8 + (21j + 22j + 22j + 21j) + 9j + 16 + 1000j,
                10j + 15, 11j + 14, 12j + 13 + 200j]
    np.testing.assert_allclose(adm, expected, atol=1e-12)


def test_ybus_branch_and_shunt_flow():
    """test_y_bus.cpp:198-226: flows for u = {1,2,3,4} -- checked through the result extraction of a linear PF whose
    solution we do not control, so recompute from the returned u instead."""
    grid, _ = three_bus_grid(True)
    out = orc.math_pf(grid, "newton_raphson", 1e-12)
    u = out["u"][:, 0]
    bp = grid.branch_param[:, :, 0, 0]
    for b, (f, t) in enumerate(grid.branch_bus_idx):
        i_f = bp[b, 0] * u[f] + bp[b, 1] * u[t]
        i_t = bp[b, 2] * u[f] + bp[b, 3] * u[t]
        np.testing.assert_allclose(out["branch"][b, :, 0], [u[f] * np.conj(i_f), u[t] * np.conj(i_t), i_f, i_t], atol=1e-14)
    i_sh = -grid.shunt_param[0, 0, 0] * u[2]
    np.testing.assert_allclose(out["shunt"][0, :, 0], [u[2] * np.conj(i_sh), i_sh], atol=1e-14)


def test_ybus_one_bus_systems():
    bag = orc.ybus_structure(1, [], [], [0, 0])
    assert L(bag.i64("row_indptr")) == [0, 1] and L(bag.i64("col_indices")) == [0]
    assert L(bag.i64("y_bus_entry_indptr")) == [0, 0] and L(bag.i64("lu_transpose_entry")) == [0]
    bag = orc.ybus_structure(1, [[0, 0]], [], [0, 0])
    assert L(bag.i64("row_indptr")) == [0, 1] and L(bag.i64("y_bus_entry_indptr")) == [0, 4]
    assert L(bag.i64("bus_entry")) == [0]


def test_ybus_cyclic_three_bus():
    bag = orc.ybus_structure(3, [[0, 1], [1, 2], [2, 0], [1, 1], [2, 2]], [], [0, 0, 0, 0])
    assert L(bag.i64("row_indptr")) == [0, 3, 6, 9]
    assert L(bag.i64("col_indices")) == [0, 1, 2, 0, 1, 2, 0, 1, 2]
    assert L(bag.i64("bus_entry")) == [0, 4, 8]
    assert L(bag.i64("lu_transpose_entry")) == [0, 3, 6, 1, 4, 7, 2, 5, 8]
    assert L(bag.i64("y_bus_entry_indptr")) == [0, 2, 3, 4, 5, 11, 12, 13, 14, 20]


def test_ybus_fill_in():
    bag = orc.ybus_structure(3, [[1, 0], [0, 2]], [[1, 2]], [0, 0, 0, 0])
    assert L(bag.i64("row_indptr")) == [0, 3, 5, 7]
    assert L(bag.i64("col_indices")) == [0, 1, 2, 0, 1, 0, 2]
    assert L(bag.i64("bus_entry")) == [0, 4, 6]
    assert L(bag.i64("lu_transpose_entry")) == [0, 3, 6, 1, 4, 7, 2, 5, 8]
    assert L(bag.i64("y_bus_entry_indptr")) == [0, 2, 3, 4, 5, 6, 7, 8]
    assert L(bag.i64("row_indptr_lu")) == [0, 3, 6, 9]
    assert L(bag.i64("col_indices_lu")) == [0, 1, 2, 0, 1, 2, 0, 1, 2]
    assert L(bag.i64("map_lu_y_bus")) == [0, 1, 2, 3, 4, -1, 5, -1, 6]
    assert L(bag.i64("diag_lu")) == [0, 4, 8]


def test_sparse_ordering():
    graph = {0: [3, 5], 1: [4, 5, 8], 2: [4, 5, 6], 3: [6, 7], 4: [6, 8], 6: [7, 8, 9], 7: [8, 9], 8: [9]}
    alpha, fills = orc.min_degree(graph)
    assert alpha == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
    assert fills == [(3, 5), (4, 5), (5, 8), (5, 6), (5, 7)]


def _pairs(a):
    return [tuple(x) for x in np.asarray(a).reshape(-1, 2).tolist()]


def test_topology_fourteen_nodes():
    branch_node_idx = [[0, 1], [0, 3], [3, 4], [7, 8], [6, 7], [4, 2], [5, 4], [4, 5], [1, 1]]
    branch3_node_idx = [[1, 3, 2], [11, 7, 8], [10, 6, 5], [4, 12, 12], [13, 13, 13]]
    branch_connected = [[1, 1], [1, 1], [1, 0], [1, 1], [0, 1], [0, 1], [1, 1], [1, 1], [1, 1]]
    branch3_connected = [[1, 1, 1], [1, 1, 1], [0, 1, 1], [1, 1, 1], [1, 1, 1]]
    bag = orc.topology(
        14, branch_node_idx, branch_connected, [0.0, 1.0, 0, 0, 0, 0, 0, 0, 0],
        source_node_idx=[0, 5, 9, 3, 13], source_connected=[1, 1, 0, 0, 1],
        shunt_node_idx=[3, 6, 9], load_gen_node_idx=[0, 11, 5, 1], load_gen_type=[0, 0, 2, 1],
        branch3_node_idx=branch3_node_idx, branch3_connected=branch3_connected,
        branch3_phase_shift=[[0.0, -1.0, 0.0]] + [[0.0] * 3] * 4,
    )
    assert bag.i64("n_math")[0] == 3
    assert _pairs(bag.i64("coup.node")) == [
        (0, 1), (0, 2), (0, 0), (0, 4), (1, 4), (1, 5), (1, 0), (-1, -1), (-1, -1), (-1, -1), (-1, -1), (-1, -1), (1, 2),
        (2, 1), (0, 3), (-1, -1), (1, 1), (1, 3), (2, 0)]
    assert _pairs(bag.i64("coup.source")) == [(0, 0), (1, 0), (-1, -1), (-1, -1), (2, 0)]
    assert _pairs(bag.i64("coup.branch")) == [(0, 0), (0, 1), (0, 2), (-1, -1), (-1, -1), (0, 3), (1, 0), (1, 1), (0, 4)]
    assert bag.i64("coup.branch3").reshape(-1, 4).tolist() == [
        [0, 5, 6, 7], [-1, -1, -1, -1], [1, 2, 3, 4], [1, 5, 6, 7], [2, 0, 1, 2]]
    assert _pairs(bag.i64("coup.load_gen")) == [(0, 0), (-1, -1), (1, 0), (0, 1)]
    assert _pairs(bag.i64("coup.shunt")) == [(0, 0), (1, 0), (-1, -1)]

    def dense(indptr):  # sparse group indptr -> element->group list
        indptr = np.asarray(indptr)
        return np.repeat(np.arange(len(indptr) - 1), np.diff(indptr)).tolist()

    # math model 0
    assert bag.i64("g0.slack_bus")[0] == 1
    assert dense(bag.i64("g0.sources_per_bus")) == [1] and len(bag.i64("g0.sources_per_bus")) == 6
    assert _pairs(bag.i64("g0.branch_bus_idx")) == [(1, 2), (1, 4), (4, -1), (-1, 0), (2, 2), (2, 3), (4, 3), (0, 3)]
    assert L(bag.f64("g0.phase_shift")) == [0.0, 0.0, 0.0, 0.0, -1.0]
    assert dense(bag.i64("g0.load_gens_per_bus")) == [1, 2]
    assert L(bag.i64("g0.load_gen_type")) == [0, 1]
    assert dense(bag.i64("g0.shunts_per_bus")) == [4]
    assert _pairs(bag.i64("g0.fill_in")) == [(2, 4)]
    # math model 1
    assert bag.i64("g1.slack_bus")[0] == 5
    assert dense(bag.i64("g1.sources_per_bus")) == [5]
    assert _pairs(bag.i64("g1.branch_bus_idx")) == [(5, 4), (4, 5), (-1, 1), (0, 1), (5, 1), (4, 3), (2, 3), (2, 3)]
    assert L(bag.f64("g1.phase_shift")) == [0.0] * 6
    assert dense(bag.i64("g1.load_gens_per_bus")) == [5] and L(bag.i64("g1.load_gen_type")) == [2]
    assert dense(bag.i64("g1.shunts_per_bus")) == [0]
    assert _pairs(bag.i64("g1.fill_in")) == []
    # math model 2
    assert bag.i64("g2.slack_bus")[0] == 1
    assert _pairs(bag.i64("g2.branch_bus_idx")) == [(1, 0), (1, 0), (1, 0)]
    assert L(bag.f64("g2.phase_shift")) == [0.0, 0.0]


def test_cycle_reorder_nine_nodes():
    branches = [[0, 3], [0, 5], [1, 4], [1, 5], [1, 8], [2, 4], [2, 5], [2, 6], [3, 6], [3, 7], [4, 6], [4, 8], [6, 7],
                [6, 8], [6, 9], [7, 8], [7, 9], [8, 9]]
    bag = orc.topology(10, branches, [[1, 1]] * 18, [0.0] * 18, [0], [1])
    assert _pairs(bag.i64("coup.node")) == [(0, i) for i in range(10)]
    assert _pairs(bag.i64("g0.fill_in")) == [(3, 5), (4, 5), (5, 8), (5, 6), (5, 7)]


def test_cycle_reorder_seven_nodes():
    branches = [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [0, 5], [1, 4], [6, 0], [6, 2], [5, 1], [3, 1], [6, 1], [2, 1]]
    bag = orc.topology(7, branches, [[1, 1]] * 13, [0.0] * 13, [0], [1])
    assert _pairs(bag.i64("coup.node")) == [(0, 0), (0, 3), (0, 1), (0, 2), (0, 4), (0, 5), (0, 6)]
    assert _pairs(bag.i64("g0.fill_in")) == [(5, 6), (2, 6), (4, 6)]
