"""Host side of the product's model level (components -> per-unit parameters, topology, coupling, PF input) against the
oracle: integer arrays bit-exact, parameters bit-exact (same formulas, same operation order).  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as orc
import pgm_b200
import validation_cases as vc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INDEX_NAMES = ("slack_bus", "is_radial", "branch_bus_idx", "fill_in", "sources_per_bus", "shunts_per_bus",
               "load_gens_per_bus", "load_gen_type", "voltage_regulators_per_load_gen", "row_indptr", "col_indices", "bus_entry", "row_indptr_lu",
               "col_indices_lu", "diag_lu", "map_lu_y_bus", "lu_transpose_entry")
COUPLING = ("coup.node", "coup.branch", "coup.shunt", "coup.load_gen", "coup.source", "coup.voltage_regulator")
REAL_NAMES = ("branch_param", "shunt_param", "source_param", "s_injection", "source_u_ref")


def assert_model_parity(input_data, sym_options=(True, False)):
    ref = orc.Model(input_data)
    prod = pgm_b200.PowerGridModel(input_data)
    for sym in sym_options:
        bag = ref.export_math(sym)
        n_math = int(bag.i64("n_math")[0])
        assert prod.n_math_groups() == n_math
        n_node = len(input_data["node"])
        for name in COUPLING:
            got = prod.math_index(0, name) if n_math else None
            if got is not None:
                exp = bag.i64(name)
                if name == "coup.node":
                    exp = exp[: 2 * n_node]
                assert got.tolist() == exp.tolist(), name
        for g in range(n_math):
            for name in INDEX_NAMES:
                assert prod.math_index(g, name).tolist() == bag.i64(f"g{g}.{name}").tolist(), (g, name)
            assert np.array_equal(prod.math_real(g, sym, "phase_shift"), bag.f64(f"g{g}.phase_shift"))
            for name in REAL_NAMES:
                got, exp = prod.math_real(g, sym, name), bag.f64(f"g{g}.{name}")
                assert got.shape == exp.shape, (g, name)
                assert np.array_equal(got, exp, equal_nan=True), (g, sym, name, np.nanmax(np.abs(got - exp)))


@pytest.mark.parametrize("name", sorted(vc.load_cases()))
def test_validation_case_math_models_match_oracle(name):
    inp = vc.to_numpy(vc.load_cases()[name]["input"], "input")
    if len(inp.get("generic_branch", ())):  # no asymmetric parameters (GenericBranch::asym_calc_param throws NotImplementedError)
        assert_model_parity(inp, sym_options=(True,))
        with pytest.raises(pgm_b200.PgmB200Error, match="not yet implemented"):
            pgm_b200.PowerGridModel(inp).math_real(0, False, "branch_param")
    else:
        assert_model_parity(inp)


@pytest.mark.parametrize("rings", [False, True])
def test_fictional_grid_1500(rings):
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=rings, has_lv_ring=rings, **pgm_b200.BENCHMARK_OPTION)
    counts = {k: len(v) for k, v in grid.input_data.items()}
    # sizes derived independently in SURVEY.md section 8(d) / BASELINE.md section 3 (seed 0)
    if not rings:
        assert counts == {"node": 2605, "line": 2600, "transformer": 7, "shunt": 4, "source": 1, "sym_load": 197, "asym_load": 1200}
    else:
        assert counts == {"node": 1804, "line": 1851, "transformer": 6, "shunt": 4, "source": 1, "sym_load": 198, "asym_load": 800}
    assert_model_parity(grid.input_data)
    model = pgm_b200.PowerGridModel(grid.input_data)
    assert len(model.math_index(0, "col_indices_lu")) == (7813 if not rings else 7492)
    assert len(model.math_index(0, "fill_in")) == (0 if not rings else 2000)


def test_fictional_grid_50k_ringed_sizes_and_ordering():
    opt = dict(pgm_b200.BENCHMARK_OPTION, n_node_total_specified=50000)
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **opt)
    counts = {k: len(v) for k, v in grid.input_data.items()}
    assert counts == {"node": 53068, "line": 77275, "transformer": 70, "shunt": 4, "source": 1, "sym_load": 134, "asym_load": 26400}
    assert_model_parity(grid.input_data, sym_options=(True,))


def test_batch_generator_is_deterministic_and_scaled():
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    a, b = grid.batch_update(7, seed=0), grid.batch_update(7, seed=0)
    assert a["sym_load"].tobytes() == b["sym_load"].tobytes() and a["asym_load"].tobytes() == b["asym_load"].tobytes()
    sym_in = grid.input_data["sym_load"]
    ratio = a["sym_load"]["p_specified"] / sym_in["p_specified"][None, :]
    assert (ratio >= 0).all() and (ratio < 1).all() and (a["sym_load"]["id"] == sym_in["id"][None, :]).all()
    assert (a["sym_load"]["status"] == pgm_b200.structs.NA_INT_S).all()
    # different batch seeds differ
    assert grid.batch_update(7, seed=1)["sym_load"].tobytes() != a["sym_load"].tobytes()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pgm_b200.h")).read()
    names = re.findall(r"PGMB_API\s+[\w\s\*]+?\b(pgmb_\w+)\s*\(", header)
    assert len(names) >= 24
    lib = C.CDLL(os.path.join(ROOT, "power-grid-model_b200", "libpgm_b200.so"))
    for n in names:
        assert hasattr(lib, n), n


def test_no_cpu_fallback_without_gpu():
    if pgm_b200.lib().pgmb_device_count() > 0:
        pytest.skip("GPU present")
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=100, n_mv_feeder=2, n_node_per_mv_feeder=3, n_lv_feeder=2,
                                  n_connection_per_lv_feeder=3)
    model = pgm_b200.PowerGridModel(grid.input_data)
    with pytest.raises(pgm_b200.PgmB200Error) as e:
        model.calculate_power_flow()
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_invalid_input_is_reported_not_thrown():
    node = pgm_b200.structs.initialize_array("input", "node", 2)
    node["id"] = [1, 1]
    node["u_rated"] = 10e3
    with pytest.raises(pgm_b200.PgmB200Error) as e:
        pgm_b200.PowerGridModel({"node": node})
    assert "Conflicting id" in str(e.value)


@pytest.mark.parametrize("rings", [False, True])
def test_bridge_analysis_of_the_n1_route(rings):
    """Host logic of the shared-pattern N-1 route: a branch is a bridge iff removing it splits its connected component, and the
    subtree size the DFS reports for it is the size of one of the two parts (checked by brute force with scipy)."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components

    grid = pgm_b200.FictionalGrid(seed=2, n_node_total_specified=120, n_mv_feeder=3, n_node_per_mv_feeder=4, n_lv_feeder=3,
                                  n_connection_per_lv_feeder=6, has_mv_ring=rings, has_lv_ring=rings)
    inp = grid.input_data
    model = pgm_b200.PowerGridModel(inp)
    is_bridge = model.math_index(0, "branch_is_bridge")
    cut = model.math_index(0, "bridge_cut_size")
    node_pos = {int(i): k for k, i in enumerate(inp["node"]["id"])}
    ends = [(node_pos[int(b["from_node"])], node_pos[int(b["to_node"])]) for comp in ("line", "transformer") for b in inp[comp]]
    n = len(node_pos)
    assert len(is_bridge) == len(ends)

    def components(skip):
        rows = [f for k, (f, t) in enumerate(ends) if k != skip]
        cols = [t for k, (f, t) in enumerate(ends) if k != skip]
        return connected_components(sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n)), directed=False)

    n0, _ = components(-1)
    assert n0 == 1
    for k in range(len(ends)):
        nc, labels = components(k)
        assert bool(is_bridge[k]) == (nc == 2), k
        if nc == 2:
            assert cut[k] in (int((labels == 0).sum()), int((labels == 1).sum())), k
    assert is_bridge.all() if not rings else (0 < is_bridge.sum() < len(ends))
