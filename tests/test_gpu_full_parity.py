"""Full-size parity of every BASELINE config: EVERY scenario of the batch against the oracle, not a sample
(iteration counts equal, node voltages within 1e-9 pu, currents / powers within 1e-6 relative: tests/parity.py).

 C2  radial 1500-node-spec grid, symmetric NR, 1000 / 1000 scenarios
 C3  ringed grid, asymmetric NR, 1000 / 1000
 C4  radial grid, iterative_current / linear / linear_current, 10 000 scenarios per method (in slices, to bound host memory)
 C5  ringed 53 068-node grid, asymmetric NR, 512 single-line outages incl. bridges (shared pattern) against the oracle's
     per-scenario topology rebuild
 N-1 overlay route vs the reference route (PGMB_N1_EXACT=1) over 1000 outages: number of scenarios whose iteration count
     differs (expected 0; the overlay keeps the base elimination order)
"""
import os

import numpy as np
import pytest

import oracle_lib as orc
import parity
import pgm_b200

pytestmark = pytest.mark.gpu

ALL = ["node", "line", "transformer", "shunt", "source", "sym_load", "asym_load"]


def _full(grid, update, sym, method="newton_raphson", comps=ALL, max_iter=20):
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(symmetric=sym, update_data=update, calculation_method=method, output_component_types=comps,
                                     max_iterations=max_iter)
    ref = orc.Model(grid.input_data).calculate(sym=sym, update=update, threading=0, method=method, output_components=comps,
                                               max_iter=max_iter)
    assert ref["n_failed"] == 0
    return parity.compare_batch(res, model.n_iter, model.status, ref, comps)


def test_config2_every_scenario():
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    out = _full(grid, grid.batch_update(1000, seed=0), True)
    assert out["scenarios"] == 1000
    print("C2 full parity:", out)


def test_config3_every_scenario():
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **pgm_b200.BENCHMARK_OPTION)
    out = _full(grid, grid.batch_update(1000, seed=0), False)
    assert out["scenarios"] == 1000
    print("C3 full parity:", out)


@pytest.mark.parametrize("method", ["iterative_current", "linear", "linear_current"])
def test_config4_every_scenario(method):
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    n_scn, piece = 10000, 2500
    update = grid.batch_update(n_scn, seed=4)
    model = pgm_b200.PowerGridModel(grid.input_data)
    oracle = orc.Model(grid.input_data)
    # the whole batch in one GPU call (node output), then every slice with all components against the oracle
    whole = model.calculate_power_flow(update_data=update, calculation_method=method, output_component_types=["node"], max_iterations=100)
    n_iter_whole = model.n_iter.copy()
    assert (model.status == 0).all()
    worst = {"max_du_pu": 0.0, "max_rel": 0.0}
    for s0 in range(0, n_scn, piece):
        upd = {k: np.ascontiguousarray(v[s0:s0 + piece]) for k, v in update.items()}
        res = model.calculate_power_flow(update_data=upd, calculation_method=method, output_component_types=ALL, max_iterations=100)
        ref = oracle.calculate(sym=True, update=upd, threading=0, method=method, output_components=ALL, max_iter=100)
        assert ref["n_failed"] == 0
        out = parity.compare_batch(res, model.n_iter, model.status, ref, ALL)
        assert np.array_equal(n_iter_whole[s0:s0 + piece], ref["n_iter"])
        assert np.array_equal(whole["node"]["u_pu"][s0:s0 + piece], res["node"]["u_pu"])  # slicing does not change a bit
        worst = {k: max(worst[k], out[k]) for k in worst}
    print(f"C4 {method} full parity over {n_scn} scenarios:", worst)


def _outage_update(lines, picks):
    upd = pgm_b200.structs.initialize_array("update", "line", (len(picks), 1))
    upd["id"][:, 0] = lines["id"][picks]
    upd["from_status"][:, 0] = 0
    upd["to_status"][:, 0] = 0
    return upd


def test_config5_n1_every_scenario_on_the_50k_grid():
    opt = dict(pgm_b200.BENCHMARK_OPTION)
    opt["n_node_total_specified"] = 50000
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **opt)
    lines = grid.input_data["line"]
    n_scn = 512
    picks = np.random.default_rng(5).choice(len(lines), n_scn, replace=False)
    upd = _outage_update(lines, picks)
    comps = ["node", "line", "source"]
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(symmetric=False, update_data={"line": upd}, output_component_types=comps)
    assert (model.status == 0).all()
    dark = res["node"]["energized"] == 0
    assert dark.any()  # the picks include bridges
    ref = orc.Model(grid.input_data).calculate(sym=False, update={"line": upd}, threading=0, output_components=comps)
    assert ref["n_failed"] == 0
    out = parity.compare_batch(res, model.n_iter, model.status, ref, comps)
    print(f"C5 full parity over {n_scn} outages ({int(dark.any(axis=1).sum())} darken nodes):", out)


@pytest.mark.parametrize("sym", [True, False])
def test_n1_overlay_iteration_counts_equal_the_reference_route(sym, monkeypatch):
    """1000 single-line outages on the ringed 1804-node grid: the shared-pattern overlay and the reference's own-topology
    route (PGMB_N1_EXACT=1) must agree in every scenario's iteration count; voltages to rounding"""
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **pgm_b200.BENCHMARK_OPTION)
    lines = grid.input_data["line"]
    n_scn = 1000
    picks = np.random.default_rng(11).choice(len(lines), n_scn, replace=False)
    upd = _outage_update(lines, picks)
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(symmetric=sym, update_data={"line": upd}, output_component_types=["node"])
    it_overlay, st_overlay = model.n_iter.copy(), model.status.copy()
    monkeypatch.setenv("PGMB_N1_EXACT", "1")
    exact = model.calculate_power_flow(symmetric=sym, update_data={"line": upd}, output_component_types=["node"], threading=0)
    monkeypatch.delenv("PGMB_N1_EXACT")
    differ = int((it_overlay != model.n_iter).sum())
    print(f"N-1 {'sym' if sym else 'asym'}: {differ} of {n_scn} scenarios differ in iteration count between overlay and exact route")
    assert np.array_equal(st_overlay, model.status)
    assert differ == 0
    assert np.array_equal(res["node"]["energized"], exact["node"]["energized"])
    assert np.max(np.abs(res["node"]["u_pu"] - exact["node"]["u_pu"])) < 1e-9
    # and both equal the oracle's iteration counts (the oracle rebuilds the topology per scenario like the reference)
    ref = orc.Model(grid.input_data).calculate(sym=sym, update={"line": upd}, threading=0, output_components=["node"])
    assert np.array_equal(it_overlay, ref["n_iter"])


def test_resident_pipeline_flags_deliver_the_same_bytes():
    """bench.py's device-resident mode (PGMB_FLAG_RESIDENT_INPUT / _OUTPUT): the output structs it leaves in HBM are the ones
    an ordinary call delivers"""
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    update = grid.batch_update(200, seed=3)
    model = pgm_b200.PowerGridModel(grid.input_data)
    plain = model.calculate_power_flow(update_data=update)
    out = {c: pgm_b200.pinned_empty(a.shape, a.dtype) for c, a in plain.items()}
    for a in out.values():
        a.view(np.uint8)[...] = 0
    model.calculate_power_flow(update_data=update, output_buffers=out, output_component_types=list(out), flags=pgm_b200.FLAG_RESIDENT_OUTPUT)
    assert all(not a.view(np.uint8).any() for a in out.values())  # nothing was delivered
    model.calculate_power_flow(update_data=update, output_buffers=out, output_component_types=list(out),
                               flags=pgm_b200.FLAG_RESIDENT_INPUT | pgm_b200.FLAG_RESIDENT_OUTPUT)
    assert model.timing()["device_pipeline"] > 0.0
    model.calculate_power_flow(update_data=update, output_buffers=out, output_component_types=list(out), flags=pgm_b200.FLAG_RESIDENT_INPUT)
    for c in plain:
        assert out[c].tobytes() == plain[c].tobytes(), c
    other = grid.batch_update(100, seed=3)
    with pytest.raises(pgm_b200.PgmB200Error, match="RESIDENT_INPUT"):
        model.calculate_power_flow(update_data=other, flags=pgm_b200.FLAG_RESIDENT_INPUT)


@pytest.mark.skipif(pgm_b200.lib().pgmb_device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("sym,method", [(True, "newton_raphson"), (False, "newton_raphson"), (True, "iterative_current")])
def test_in_process_multi_gpu_equals_one_gpu(sym, method):
    """one PGM_calculate-level call spread over the GPUs of the box (n_devices): byte-identical to the one-GPU result"""
    n_dev = min(8, pgm_b200.lib().pgmb_device_count())
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=not sym, has_lv_ring=not sym, **pgm_b200.BENCHMARK_OPTION)
    n_scn = 1000 * n_dev if sym else 1024
    update = grid.batch_update(n_scn, seed=8)
    comps = ["node", "line", "source"]
    model = pgm_b200.PowerGridModel(grid.input_data)
    one = model.calculate_power_flow(symmetric=sym, calculation_method=method, update_data=update, output_component_types=comps)
    it_one = model.n_iter.copy()
    many = model.calculate_power_flow(symmetric=sym, calculation_method=method, update_data=update, output_component_types=comps,
                                      n_devices=n_dev)
    assert np.array_equal(it_one, model.n_iter)
    for c in comps:
        assert one[c].tobytes() == many[c].tobytes(), c


@pytest.mark.skipif(pgm_b200.lib().pgmb_device_count() < 2, reason="needs two GPUs")
def test_in_process_multi_gpu_failures_and_n1():
    """failed scenarios keep their batch numbers across devices; an N-1 batch (outage plan) is sliced per device"""
    n_dev = min(8, pgm_b200.lib().pgmb_device_count())
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **pgm_b200.BENCHMARK_OPTION)
    n_scn = 512 * n_dev
    update = grid.batch_update(n_scn, seed=2)
    bad = [5, n_scn - 7]
    for s in bad:
        update["sym_load"]["p_specified"][s] *= 1e6
    model = pgm_b200.PowerGridModel(grid.input_data)
    model.calculate_power_flow(update_data=update, output_component_types=["node"], n_devices=n_dev, continue_on_batch_error=True)
    assert sorted(np.nonzero(model.status)[0].tolist()) == bad
    assert all(f"Error in batch #{s}:" in str(model.batch_error) for s in bad)
    lines = grid.input_data["line"]
    picks = np.random.default_rng(3).choice(len(lines), n_scn, replace=n_scn > len(lines))
    upd = _outage_update(lines, picks)
    one = model.calculate_power_flow(symmetric=True, update_data={"line": upd}, output_component_types=["node", "line"])
    it_one = model.n_iter.copy()
    many = model.calculate_power_flow(symmetric=True, update_data={"line": upd}, output_component_types=["node", "line"], n_devices=n_dev)
    assert np.array_equal(it_one, model.n_iter)
    assert one["node"].tobytes() == many["node"].tobytes() and one["line"].tobytes() == many["line"].tobytes()
