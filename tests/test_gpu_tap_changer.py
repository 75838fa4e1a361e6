"""Automatic tap changer (model_tap.cpp) against the reference's own validation cases
(tests/data/power_flow/automatic-tap-regulator/*, stored in tests/golden/tap_regulator_cases.json): tap positions, voltages and
flows of the final power flow; cases the reference expects to end in MaxIterationReached must fail here as well."""
import json
import os

import numpy as np
import pytest

import pgm_b200
import validation_cases as vc

pytestmark = pytest.mark.gpu

CASES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tap_regulator_cases.json")))["cases"]
RUNS = [(n, s, m, b) for n, c in sorted(CASES.items()) for s, m, b in vc.case_runs(c)]


@pytest.mark.parametrize("name,sym,method,batch", RUNS)
def test_tap_regulator_validation_case(name, sym, method, batch):
    case = CASES[name]
    params = case["params"]
    model = pgm_b200.PowerGridModel(vc.to_numpy(case["input"], "input"))
    kw = dict(symmetric=sym, calculation_method=method, tap_changing_strategy=params["tap_changing_strategy"])
    kind = "sym_output" if sym else "asym_output"
    update = None
    if batch:
        update = vc.batch_update_arrays(vc.to_numpy(case["update_batch"], "update"))
        kw["update_data"] = update
    if "raises" in params:
        with pytest.raises(pgm_b200.PgmB200Error, match="Maximum number of iterations reached"):
            model.calculate_power_flow(**kw)
        # the search left the model as it was: the plain power flow still runs and the taps are the input's
        plain = model.calculate_power_flow(symmetric=sym, calculation_method=method, update_data=update)
        assert (plain["transformer_tap_regulator"]["tap_pos"] == -128).all()
        return
    result = model.calculate_power_flow(**kw)
    rtol, atol = params["rtol"], params["atol"]
    if batch:
        expected = vc.to_numpy(case[kind + "_batch"], kind)
        for s, exp in enumerate(expected):
            vc.compare_result({c: result[c][s] for c in exp}, exp, rtol, atol)
    else:
        vc.compare_result(result, vc.to_numpy(case[kind], kind), rtol, atol)
    # a second run gives the same answer: the tap positions of the model were put back
    again = model.calculate_power_flow(**kw)
    for c in result:
        for f in result[c].dtype.names:
            assert np.array_equal(result[c][f], again[c][f], equal_nan=result[c][f].dtype.kind == "f"), (c, f)


def test_disabled_strategy_gives_null_regulator_output():
    case = CASES["automatic-tap-regulator/single-trafo-any-valid-tap"]
    model = pgm_b200.PowerGridModel(vc.to_numpy(case["input"], "input"))
    out = model.calculate_power_flow()["transformer_tap_regulator"]
    assert (out["energized"] == 0).all() and (out["tap_pos"] == -128).all() and (out["id"] >= 0).all()


PGM_RUNS = [r for r in RUNS if r[0].split("/")[-1] in (
    "pgm-automatic-tap-line-drop-max", "auto-tap-changer-with-transformer-into-itself", "step-up-transformer-fast-any-tap",
    "auto-tap-changer-meshed-any-max-iter", "trafo-control-tap-same-side-min-voltage-tap")]


@pytest.mark.parametrize("name,sym,method,batch", PGM_RUNS)
def test_tap_changer_through_pgm_calculate(name, sym, method, batch):
    """the same search behind the reference's C API names: PGM_set_tap_changing_strategy + PGM_calculate"""
    from pgm_b200 import pgm_core

    case = CASES[name]
    params = case["params"]
    strategy = pgm_b200.TAP_STRATEGIES[params["tap_changing_strategy"]]
    model = pgm_core.PowerGridModel(vc.to_numpy(case["input"], "input"))
    kind = "sym_output" if sym else "asym_output"
    kw = dict(symmetric=sym, calculation_method=method, tap_changing_strategy=strategy)
    if batch:
        kw["update_data"] = vc.batch_update_arrays(vc.to_numpy(case["update_batch"], "update"))
    if "raises" in params:
        with pytest.raises(pgm_core.PowerGridError, match="Maximum number of iterations reached"):
            model.calculate_power_flow(**kw)
        return
    res = model.calculate_power_flow(**kw)
    assert "transformer_tap_regulator" in res
    if batch:
        for s, exp in enumerate(vc.to_numpy(case[kind + "_batch"], kind)):
            vc.compare_result({k: v[s] for k, v in res.items()}, exp, params["rtol"], params["atol"])
    else:
        vc.compare_result(res, vc.to_numpy(case[kind], kind), params["rtol"], params["atol"])


def _regulated_benchmark_grid(extra_lv_regulators):
    """BASELINE configs[1] grid with the reference benchmark's tap changer (fictional_grid_generator.hpp:615-633: one regulator
    on the station transformer); optionally one more regulator on every MV/LV transformer (second rank)"""
    grid = pgm_b200.FictionalGrid(seed=0, has_tap_changer=True, **pgm_b200.BENCHMARK_OPTION)
    data = dict(grid.input_data)
    if extra_lv_regulators:
        trafo, node = data["transformer"], data["node"]
        reg = pgm_b200.structs.initialize_array("input", "transformer_tap_regulator", len(trafo) - 1)
        for k, t in enumerate(trafo[1:]):
            u_rated = float(node["u_rated"][node["id"] == t["to_node"]][0])
            reg[k] = (int(data["transformer_tap_regulator"]["id"][0]) + 1 + k, t["id"], 1, 1, 1.02 * u_rated,
                      float(t["tap_size"]) * u_rated / float(t["u2"]) + 0.01 * u_rated, np.nan, np.nan)
        data["transformer_tap_regulator"] = np.concatenate([data["transformer_tap_regulator"], reg])
    return grid, data


@pytest.mark.parametrize("sym,extra", [(True, False), (True, True), (False, False), (False, True)])
@pytest.mark.parametrize("strategy", ["any_valid_tap", "min_voltage_tap", "max_voltage_tap", "fast_any_tap"])
def test_tap_changer_on_the_benchmark_grid_equals_the_oracle(strategy, sym, extra):
    """load-profile batch on the 2605-bus benchmark grid: every scenario's search ends at the oracle's tap positions, and the
    final power flow agrees to the solver tolerance"""
    import oracle_lib as orc
    import parity

    grid, data = _regulated_benchmark_grid(extra)
    n_scn = 24 if sym else 8
    update = grid.batch_update(n_scn, seed=3)
    comps = ["node", "transformer", "transformer_tap_regulator", "source"]
    ref = orc.Model(data).calculate(sym=sym, update=update, threading=0, tap_changing_strategy=strategy, output_components=comps)
    assert ref["n_failed"] == 0, ref["error"]
    model = pgm_b200.PowerGridModel(data)
    res = model.calculate_power_flow(symmetric=sym, update_data=update, tap_changing_strategy=strategy, output_component_types=comps)
    assert (model.status == 0).all()
    assert np.array_equal(res["transformer_tap_regulator"]["tap_pos"], ref["transformer_tap_regulator"]["tap_pos"])
    assert np.array_equal(res["transformer_tap_regulator"]["energized"], ref["transformer_tap_regulator"]["energized"])
    parity.compare_outputs(res, ref, ["node", "transformer", "source"])


@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("strategy", ["any_valid_tap", "min_voltage_tap", "max_voltage_tap", "fast_any_tap"])
def test_lockstep_search_equals_the_scenario_by_scenario_search(strategy, sym, monkeypatch):
    """one regulated transformer + a load-profile batch: the batch searches in lockstep (one batched power flow per step, the
    transformer's admittances per scenario through the branch overlay); PGMB_TAP_EXACT=1 runs every scenario's own search"""
    import parity

    grid, data = _regulated_benchmark_grid(False)
    n_scn = 96 if sym else 16
    update = grid.batch_update(n_scn, seed=5)
    model = pgm_b200.PowerGridModel(data)
    kw = dict(symmetric=sym, update_data=update, tap_changing_strategy=strategy)
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    lock = model.calculate_power_flow(**kw)
    launches_lock = int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0
    monkeypatch.setenv("PGMB_TAP_EXACT", "1")
    exact = model.calculate_power_flow(**kw)
    launches_exact = int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0 - launches_lock
    assert np.array_equal(lock["transformer_tap_regulator"]["tap_pos"], exact["transformer_tap_regulator"]["tap_pos"])
    assert (lock["transformer_tap_regulator"]["energized"] == 1).all()
    parity.compare_outputs(lock, exact, [c for c in lock if c != "transformer_tap_regulator"])
    assert launches_lock < launches_exact  # a handful of batched passes against several power flows per scenario


@pytest.mark.parametrize("seed", range(12))
def test_random_regulated_grids_equal_the_oracle(seed):
    """small fictional grids with a regulator on every transformer, random set points / bands / line-drop compensation, random
    regulators switched off: for every strategy the product ends at the oracle's tap positions, or both report the search as
    failed for the same scenarios"""
    import oracle_lib as orc

    rng = np.random.default_rng(500 + seed)
    grid = pgm_b200.FictionalGrid(seed=seed, n_node_total_specified=int(rng.integers(80, 300)), n_mv_feeder=int(rng.integers(2, 4)),
                                  n_node_per_mv_feeder=int(rng.integers(3, 6)), n_lv_feeder=int(rng.integers(2, 4)),
                                  n_connection_per_lv_feeder=int(rng.integers(4, 10)), has_mv_ring=bool(seed % 2), has_lv_ring=bool(seed % 3 == 0))
    data = dict(grid.input_data)
    trafo, node = data["transformer"], data["node"]
    reg = pgm_b200.structs.initialize_array("input", "transformer_tap_regulator", len(trafo))
    first_id = 1 + max(int(a["id"].max()) for a in data.values() if len(a))
    for k, t in enumerate(trafo):
        u_rated = float(node["u_rated"][node["id"] == t["to_node"]][0])
        step = float(t["tap_size"]) * u_rated / float(t["u2"])
        reg[k] = (first_id + k, t["id"], int(rng.random() > 0.15), 1, rng.uniform(0.97, 1.06) * u_rated, step * rng.uniform(0.6, 2.5),
                  rng.choice([np.nan, 0.0, 0.05]), rng.choice([np.nan, 0.0, 0.1]))
    data["transformer_tap_regulator"] = reg
    n_scn = 6
    update = grid.batch_update(n_scn, seed=seed)
    comps = ["node", "transformer_tap_regulator"]
    model, oracle = pgm_b200.PowerGridModel(data), orc.Model(data)
    for strategy in ("any_valid_tap", "min_voltage_tap", "max_voltage_tap", "fast_any_tap"):
        ref = oracle.calculate(sym=True, update=update, threading=0, tap_changing_strategy=strategy, output_components=comps)
        res = model.calculate_power_flow(update_data=update, tap_changing_strategy=strategy, output_component_types=comps,
                                         continue_on_batch_error=True)
        assert np.array_equal(model.status != 0, ref["status"] != 0), (strategy, model.status, ref["status"], ref["error"][:300])
        ok = ref["status"] == 0
        assert np.array_equal(res["transformer_tap_regulator"]["tap_pos"][ok], ref["transformer_tap_regulator"]["tap_pos"][ok]), strategy
        if ok.any():
            assert np.max(np.abs(res["node"]["u_pu"][ok] - ref["node"]["u_pu"][ok])) < 1e-9, strategy
