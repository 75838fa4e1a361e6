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


def _supplied(inp, line_state, trafo_state):
    """nodes connected to the source node over branches that are closed on both sides (brute force, scipy)"""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components

    node_pos = {int(i): k for k, i in enumerate(inp["node"]["id"])}
    rows, cols = [], []
    for comp, state in (("line", line_state), ("transformer", trafo_state)):
        for b, (f, t) in zip(inp[comp], state):
            if f and t:
                rows.append(node_pos[int(b["from_node"])])
                cols.append(node_pos[int(b["to_node"])])
    n = len(node_pos)
    _, labels = connected_components(sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n)), directed=False)
    return labels == labels[node_pos[int(inp["source"]["node"][0])]]


@pytest.mark.parametrize("seed", [3, 4, 5])
@pytest.mark.parametrize("ties", [False, True])
def test_host_planning_of_branch_switching_batches(ties, seed):
    """Host side of the shared-pattern route for N-k / reconfiguration / tap batches, no device needed: for random
    scenarios (1-5 branches opened fully or on one side, taps moved, open ties closed) the plan puts every scenario on the shared
    pattern, uses one overlay slot per branch whose state or tap differs from the pattern's grid, and masks exactly the buses a
    brute-force connectivity search finds without supply.  ties=True: three lines are open in the base state and scenarios close
    them, so the plan is made on the union grid, where a tie that stays open takes a slot."""
    grid = pgm_b200.FictionalGrid(seed=seed, n_node_total_specified=200, n_mv_feeder=3, n_node_per_mv_feeder=5, n_lv_feeder=3,
                                  n_connection_per_lv_feeder=6, has_mv_ring=True, has_lv_ring=True)
    inp = {k: v.copy() for k, v in grid.input_data.items()}
    lines, trafos = inp["line"], inp["transformer"]
    rng = np.random.default_rng(11 + seed)
    tie_idx = np.zeros(0, int)
    if ties:
        not_bridge = np.flatnonzero(pgm_b200.PowerGridModel(inp).math_index(0, "branch_is_bridge")[: len(lines)] == 0)
        tie_idx = not_bridge[[1, len(not_bridge) // 2, len(not_bridge) - 2]]
        lines["from_status"][tie_idx[0]] = 0
        lines["to_status"][tie_idx[0]] = 0
        lines["to_status"][tie_idx[1]] = 0
        lines["from_status"][tie_idx[2]] = 0
        lines["to_status"][tie_idx[2]] = 0
        assert _supplied(inp, list(zip(lines["from_status"], lines["to_status"])), list(zip(trafos["from_status"], trafos["to_status"]))).all()
    model = pgm_b200.PowerGridModel(inp)
    n_scn = 60
    l_rows, t_rows, expect_slots, expect_dark = [], [], [], []
    for s in range(n_scn):
        l_state = [[int(f), int(t)] for f, t in zip(lines["from_status"], lines["to_status"])]
        t_state = [[int(f), int(t)] for f, t in zip(trafos["from_status"], trafos["to_status"])]
        taps = trafos["tap_pos"].astype(int).copy()
        pick = rng.choice(np.setdiff1d(np.arange(len(lines)), tie_idx), size=int(rng.integers(0, 5)), replace=False)
        closed = rng.choice(tie_idx, size=int(rng.integers(0, len(tie_idx) + 1)), replace=False) if ties else np.zeros(0, int)
        lu = pgm_b200.structs.initialize_array("update", "line", len(pick) + len(closed))
        for k, li in enumerate(pick):
            side = rng.integers(0, 3)  # both sides, from only, to only
            lu["id"][k] = lines["id"][li]
            if side != 2:
                lu["from_status"][k] = 0
                l_state[li][0] = 0
            if side != 1:
                lu["to_status"][k] = 0
                l_state[li][1] = 0
        for k, li in enumerate(closed):
            lu["id"][len(pick) + k] = lines["id"][li]
            lu["from_status"][len(pick) + k] = 1
            lu["to_status"][len(pick) + k] = 1
            l_state[li] = [1, 1]
        n_t = int(rng.integers(0, 3))
        tp = rng.choice(np.arange(1, len(trafos)), size=n_t, replace=False)
        tu = pgm_b200.structs.initialize_array("update", "transformer", n_t)
        for k, ti in enumerate(tp):
            tu["id"][k] = trafos["id"][ti]
            if rng.random() < 0.5:
                tu["from_status"][k] = 0
                t_state[ti][0] = 0
            else:
                taps[ti] = int(np.clip(taps[ti] + rng.integers(-2, 3), min(trafos["tap_min"][ti], trafos["tap_max"][ti]),
                                       max(trafos["tap_min"][ti], trafos["tap_max"][ti])))
                tu["tap_pos"][k] = taps[ti]
        l_rows.append(lu)
        t_rows.append(tu)
        # slots: branches whose state differs from the grid the pattern was built for (ties closed), or whose tap moved
        pattern_l = [[1, 1] if li in tie_idx else [int(lines["from_status"][li]), int(lines["to_status"][li])] for li in range(len(lines))]
        slots = sum(a != b for a, b in zip(l_state, pattern_l))
        slots += sum(list(a) != [int(f), int(t)] or taps[ti] != trafos["tap_pos"][ti]
                     for ti, (a, f, t) in enumerate(zip(t_state, trafos["from_status"], trafos["to_status"])))
        expect_slots.append(slots)
        expect_dark.append(int((~_supplied(inp, l_state, t_state)).sum()))
    update = {"line": {"data": np.concatenate(l_rows), "indptr": np.cumsum([0] + [len(x) for x in l_rows])},
              "transformer": {"data": np.concatenate(t_rows), "indptr": np.cumsum([0] + [len(x) for x in t_rows])}}
    for sym in (True, False):
        plan = model.outage_plan(update, symmetric=sym)
        assert (plan[:, 3] == int(ties)).all()
        shared = plan[:, 0] == 0
        assert shared.sum() >= n_scn - 2  # (a scenario with more than eight changed branches would take its own topology)
        assert np.array_equal(plan[shared, 1], np.array(expect_slots)[shared]), (plan[:, 1], expect_slots)
        assert np.array_equal(plan[shared, 2], np.array(expect_dark)[shared]), (plan[:, 2], expect_dark)
    assert 0 < (np.array(expect_dark) > 0).sum() < n_scn


def test_host_planning_sends_what_the_pattern_cannot_hold_to_the_own_topology_route():
    """Edge cases of the planning (CPU): a scenario that closes a transformer whose LV grid is dark in the base state (its nodes are
    in no math model: the union grid cannot hold it), one that changes more than eight branches, one with an unknown id -- each
    takes the own-topology route alone; a scenario that names a branch without changing anything needs no slot."""
    grid = pgm_b200.FictionalGrid(seed=3, n_node_total_specified=200, n_mv_feeder=3, n_node_per_mv_feeder=5, n_lv_feeder=3,
                                  n_connection_per_lv_feeder=6, has_mv_ring=True, has_lv_ring=False)
    inp = {k: v.copy() for k, v in grid.input_data.items()}
    lines, trafos = inp["line"], inp["transformer"]
    inp["transformer"]["from_status"][2] = 0  # this transformer's LV grid is dark in the base state
    inp["transformer"]["to_status"][2] = 0
    model = pgm_b200.PowerGridModel(inp)
    assert (model.math_index(0, "coup.node").reshape(-1, 2)[:, 0] == -1).any()  # the dark LV grid
    n_scn = 12
    l_rows = [pgm_b200.structs.initialize_array("update", "line", 1) for _ in range(n_scn)]
    t_rows = [pgm_b200.structs.initialize_array("update", "transformer", 0) for _ in range(n_scn)]
    live = [k for k in range(len(lines)) if k % 3 == 0][: n_scn + 10]
    for s in range(n_scn):
        l_rows[s]["id"], l_rows[s]["from_status"], l_rows[s]["to_status"] = lines["id"][live[s]], 0, 0
    # 3: closes the transformer towards the dark grid
    t_rows[3] = pgm_b200.structs.initialize_array("update", "transformer", 1)
    t_rows[3]["id"], t_rows[3]["from_status"], t_rows[3]["to_status"] = trafos["id"][2], 1, 1
    # 5: nine lines at once
    l_rows[5] = pgm_b200.structs.initialize_array("update", "line", 9)
    l_rows[5]["id"], l_rows[5]["from_status"], l_rows[5]["to_status"] = lines["id"][live[12:21]], 0, 0
    # 7: unknown id
    l_rows[7]["id"] = 987654321
    # 9: names a line, changes nothing
    l_rows[9]["from_status"], l_rows[9]["to_status"] = 1, -128
    update = {"line": {"data": np.concatenate(l_rows), "indptr": np.cumsum([0] + [len(x) for x in l_rows])},
              "transformer": {"data": np.concatenate(t_rows), "indptr": np.cumsum([0] + [len(x) for x in t_rows])}}
    plan = model.outage_plan(update)
    assert (plan[:, 3] == 0).all()  # nothing the union grid could take
    assert plan[:, 0].tolist() == [1 if s in (3, 5, 7) else 0 for s in range(n_scn)]
    assert plan[9].tolist() == [0, 0, 0, 0]
    assert all(plan[s, 1] == 1 for s in range(n_scn) if s not in (3, 5, 7, 9))
