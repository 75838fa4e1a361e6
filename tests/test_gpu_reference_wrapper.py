"""GPU: the reference's UNCHANGED Python wrapper (power_grid_model.PowerGridModel.calculate_power_flow, ctypes over PGM_*) with
libpgm_b200.so as its core library -- the drop-in claim of BASELINE.json's north_star, end to end:
 (a) the reference's power-flow validation cases against their golden outputs with the reference's tolerances,
 (b) the API model scenarios (40 V / 70 V) incl. batch errors surfacing as the wrapper's PowerGridBatchError,
 (c) the 1000-scenario benchmark batch: the bytes pgm_b200.PowerGridModel produces.
Skipped when the wrapper is not in baseline/_ref (tools/install_reference_wrapper.py puts it there; it is not product source)."""
import numpy as np
import pytest

import pgm_b200
import reference_wrapper
import validation_cases as vc

pgm = reference_wrapper.load()
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(pgm is None, reason="reference Python wrapper not available (tools/install_reference_wrapper.py)")]

CASES = vc.load_cases()
GPU_METHODS = {"newton_raphson", "iterative_current", "linear", "linear_current"}
RUNS = [(n, s, m, b) for n, c in sorted(CASES.items()) for s, m, b in vc.case_runs(c) if m in GPU_METHODS]


@pytest.mark.parametrize("name,sym,method,is_batch", RUNS, ids=[f"{n}-{'sym' if s else 'asym'}-{m}-{'batch' if b else 'single'}" for n, s, m, b in RUNS])
def test_validation_case_through_the_reference_wrapper(name, sym, method, is_batch):
    case = CASES[name]
    params = case["params"]
    model = pgm.PowerGridModel(vc.to_numpy(case["input"], "input"))
    kind = "sym_output" if sym else "asym_output"
    if not is_batch:
        res = model.calculate_power_flow(symmetric=sym, calculation_method=method)
        vc.compare_result({str(k.value if hasattr(k, "value") else k): v for k, v in res.items()}, vc.to_numpy(case[kind], kind),
                          params["rtol"], params["atol"])
    else:
        updates = vc.to_numpy(case["update_batch"], "update")
        expected = vc.to_numpy(case[kind + "_batch"], kind)
        res = model.calculate_power_flow(symmetric=sym, calculation_method=method, update_data=vc.batch_update_arrays(updates))
        res = {str(k.value if hasattr(k, "value") else k): v for k, v in res.items()}
        for s, exp in enumerate(expected):
            vc.compare_result({k: v[s] for k, v in res.items()}, exp, params["rtol"], params["atol"])


def test_api_model_scenarios_and_batch_error():
    from power_grid_model.errors import IterationDiverge, PowerGridBatchError

    node = pgm.initialize_array("input", "node", 1)
    node["id"], node["u_rated"] = 0, 100.0
    source = pgm.initialize_array("input", "source", 1)
    source["id"], source["node"], source["status"], source["u_ref"], source["sk"], source["rx_ratio"] = 1, 0, 1, 1.0, 1000.0, 0.0
    load = pgm.initialize_array("input", "sym_load", 1)
    load["id"], load["node"], load["status"], load["type"], load["p_specified"], load["q_specified"] = 2, 0, 1, 2, 0.0, 500.0
    model = pgm.PowerGridModel({"node": node, "source": source, "sym_load": load})
    assert model.calculate_power_flow()["node"]["u"][0] == pytest.approx(50.0)
    upd = pgm.initialize_array("update", "sym_load", (2, 1))
    upd["id"], upd["q_specified"] = 2, [[100.0], [300.0]]
    res = model.calculate_power_flow(update_data={"sym_load": upd}, output_component_types=["node"])
    assert res["node"]["u"][:, 0] == pytest.approx([90.0, 70.0])
    # columnar output through the wrapper's own data-format machinery
    col = model.calculate_power_flow(update_data={"sym_load": upd}, output_component_types={"node": ["u_pu"]})
    assert col["node"]["u_pu"][:, 0] == pytest.approx([0.9, 0.7])
    with pytest.raises(IterationDiverge):
        model.calculate_power_flow(max_iterations=1, error_tolerance=1e-100, symmetric=False)
    upd["id"][1, 0] = 99
    with pytest.raises(PowerGridBatchError) as e:
        model.calculate_power_flow(update_data={"sym_load": upd})
    assert e.value.failed_scenarios.tolist() == [1] and e.value.succeeded_scenarios.tolist() == [0]
    res = model.calculate_power_flow(update_data={"sym_load": upd}, continue_on_batch_error=True, output_component_types=["node"])
    assert res["node"]["u"][0, 0] == pytest.approx(90.0)


def test_benchmark_batch_through_the_reference_wrapper():
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    update = grid.batch_update(1000, seed=0)
    theirs = pgm.PowerGridModel(grid.input_data).calculate_power_flow(update_data=update)
    ours = pgm_b200.PowerGridModel(grid.input_data).calculate_power_flow(update_data=update)
    theirs = {str(k.value if hasattr(k, "value") else k): v for k, v in theirs.items()}
    assert set(theirs) == set(ours)
    for comp in ours:
        assert theirs[comp].tobytes() == ours[comp].tobytes(), comp


import json  # noqa: E402
import os  # noqa: E402

TAP_CASES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tap_regulator_cases.json")))["cases"]
TAP_RUNS = [(n, s, m, b) for n, c in sorted(TAP_CASES.items()) for s, m, b in vc.case_runs(c)]


@pytest.mark.parametrize("name,sym,method,is_batch", TAP_RUNS,
                         ids=[f"{n}-{'sym' if s else 'asym'}-{m}-{'batch' if b else 'single'}" for n, s, m, b in TAP_RUNS])
def test_tap_regulator_case_through_the_reference_wrapper(name, sym, method, is_batch):
    """the reference's automatic-tap-regulator validation cases through its unchanged wrapper:
    calculate_power_flow(tap_changing_strategy=...) -> PGM_set_tap_changing_strategy + PGM_calculate of libpgm_b200.so"""
    from power_grid_model.errors import PowerGridError

    case = TAP_CASES[name]
    params = case["params"]
    model = pgm.PowerGridModel(vc.to_numpy(case["input"], "input"))
    kind = "sym_output" if sym else "asym_output"
    kw = dict(symmetric=sym, calculation_method=method, tap_changing_strategy=params["tap_changing_strategy"])
    if is_batch:
        kw["update_data"] = vc.batch_update_arrays(vc.to_numpy(case["update_batch"], "update"))
    if "raises" in params:
        with pytest.raises(PowerGridError, match="Maximum number of iterations reached"):
            model.calculate_power_flow(**kw)
        return
    res = model.calculate_power_flow(**kw)
    res = {str(k.value if hasattr(k, "value") else k): v for k, v in res.items()}
    assert "transformer_tap_regulator" in res
    if is_batch:
        for s, exp in enumerate(vc.to_numpy(case[kind + "_batch"], kind)):
            vc.compare_result({k: v[s] for k, v in res.items()}, exp, params["rtol"], params["atol"])
    else:
        vc.compare_result(res, vc.to_numpy(case[kind], kind), params["rtol"], params["atol"])
