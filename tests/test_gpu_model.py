"""GPU parity tests at the PGM_calculate seam: pgm_b200.PowerGridModel (CUDA through the C-ABI) against
 (a) the reference's golden validation outputs (tests/golden/power_flow_cases.json) with the reference's tolerances,
 (b) the oracle on the benchmark grids (BASELINE configs), with the north_star bars: iteration counts identical,
     voltages within 1e-9 pu, powers / currents within 1e-6 relative."""
import numpy as np
import pytest

import oracle_lib as orc
import pgm_b200
import validation_cases as vc

pytestmark = pytest.mark.gpu

CASES = vc.load_cases()
# methods / symmetries the GPU engine implements so far
GPU_METHODS = {"newton_raphson", "iterative_current", "linear", "linear_current"}
RUNS = [(n, s, m, b) for n, c in sorted(CASES.items()) for s, m, b in vc.case_runs(c) if m in GPU_METHODS]


@pytest.mark.parametrize("name,sym,method,is_batch", RUNS, ids=[f"{n}-{'sym' if s else 'asym'}-{m}-{'batch' if b else 'single'}" for n, s, m, b in RUNS])
def test_reference_validation_case(name, sym, method, is_batch):
    case = CASES[name]
    params = case["params"]
    model = pgm_b200.PowerGridModel(vc.to_numpy(case["input"], "input"))
    kind = "sym_output" if sym else "asym_output"
    if not is_batch:
        res = model.calculate_power_flow(symmetric=sym, calculation_method=method)
        vc.compare_result(res, vc.to_numpy(case[kind], kind), params["rtol"], params["atol"])
    else:
        updates = vc.to_numpy(case["update_batch"], "update")
        expected = vc.to_numpy(case[kind + "_batch"], kind)
        res = model.calculate_power_flow(symmetric=sym, calculation_method=method, update_data=vc.batch_update_arrays(updates))
        for s, exp in enumerate(expected):
            vc.compare_result({k: v[s] for k, v in res.items()}, exp, params["rtol"], params["atol"])
        if kind in case:  # model unchanged after the batch
            res = model.calculate_power_flow(symmetric=sym, calculation_method=method)
            vc.compare_result(res, vc.to_numpy(case[kind], kind), params["rtol"], params["atol"])


def _compare_with_oracle(res, ref, n_scn):
    for comp, arr in res.items():
        exp = ref[comp]
        for name in arr.dtype.names:
            a, e = arr[name], exp[name]
            if a.dtype.kind in "iu":
                assert np.array_equal(a, e), (comp, name)
            elif name in ("u_pu",):
                assert np.max(np.abs(a - e)) < 1e-9, (comp, name, np.max(np.abs(a - e)))
            elif name == "u_angle":
                assert np.max(np.abs(np.angle(np.exp(1j * (a - e))))) < 1e-9, (comp, name)
            else:
                # 1e-6 relative, plus an absolute floor of 1e-9 pu (1 MVA base => 1e-3 W/var/VA; currents: 1e-7 A) for
                # flows that are ~0 by cancellation (e.g. the no-load side of an LV transformer)
                atol = 1e-3 if name[0] in "pqs" and name != "pf" else (1e-7 if name.startswith("i") else 1e-9)
                if name == "pf":  # p / s: the absolute floor of p carries over (a pure-reactive shunt has p ~ 0)
                    atol = 1e-9 + 1e-3 / np.maximum(np.abs(exp["s"]), 1e-300)
                err = np.abs(a - e) - (1e-6 * np.abs(e) + atol)
                assert np.max(err) <= 0, (comp, name, np.max(np.abs(a - e)))


@pytest.mark.parametrize("rings", [False, True])
def test_benchmark_grid_batch_matches_oracle(rings):
    """BASELINE config 1/2 (radial) and the sym variant of config 3 (rings): 1500-node fictional grid, load-profile batch"""
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=rings, has_lv_ring=rings, **pgm_b200.BENCHMARK_OPTION)
    n_scn = 48
    update = grid.batch_update(n_scn, seed=0)
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(symmetric=True, update_data=update)
    ref = orc.Model(grid.input_data).calculate(sym=True, update=update, threading=0)
    assert ref["n_failed"] == 0
    assert np.array_equal(model.n_iter, ref["n_iter"]), (model.n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, n_scn)
    # single calculation (config 1) + model unchanged by the batch
    single = model.calculate_power_flow(symmetric=True)
    ref1 = orc.Model(grid.input_data).calculate(sym=True)
    assert model.n_iter[0] == ref1["n_iter"][0]
    _compare_with_oracle({k: v[None] for k, v in single.items()}, ref1, 1)


def test_asymmetric_benchmark_grid_batch_matches_oracle():
    """BASELINE config 3 at reduced size: ringed fictional grid, asymmetric newton_raphson, three-phase load profile"""
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5, n_mv_feeder=3,
                                  has_mv_ring=True, has_lv_ring=True)
    n_scn = 13
    update = grid.batch_update(n_scn, seed=1)
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(symmetric=False, update_data=update)
    ref = orc.Model(grid.input_data).calculate(sym=False, update=update, threading=0)
    assert ref["n_failed"] == 0
    assert np.array_equal(model.n_iter, ref["n_iter"]), (model.n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, n_scn)
    single = model.calculate_power_flow(symmetric=False)
    ref1 = orc.Model(grid.input_data).calculate(sym=False)
    assert model.n_iter[0] == ref1["n_iter"][0]
    _compare_with_oracle({k: v[None] for k, v in single.items()}, ref1, 1)


@pytest.mark.parametrize("rings", [False, True])
@pytest.mark.parametrize("q_lim", [None, 6e6])
def test_voltage_regulators_on_the_benchmark_grid_match_oracle(rings, q_lim):
    """PV buses: 12 regulated generators on the 1500-node grid, with and without reactive-power limits
    (with limits some buses switch PV -> PQ at the iteration-2 check, lower and upper), generators switched off per scenario,
    two generators sharing a bus.  Iteration counts, limit flags and the allocated Q must equal the oracle's."""
    import grids

    n_scn = 24
    _, inp, update = grids.regulated_benchmark_grid(rings, q_lim=q_lim, n_scn=n_scn)
    model = pgm_b200.PowerGridModel(inp)
    res = model.calculate_power_flow(symmetric=True, update_data=update)
    assert model.timing()["output"] == 0.0  # the device pipeline ran (the host-staged route converts outputs on the host)
    ref = orc.Model(inp).calculate(sym=True, update=update, threading=0)
    assert ref["n_failed"] == 0
    assert np.array_equal(model.n_iter, ref["n_iter"]), (model.n_iter, ref["n_iter"])
    if q_lim is not None:
        assert {0, 1, 2} <= set(np.unique(ref["voltage_regulator"]["limit_violated"]).tolist())
    _compare_with_oracle(res, ref, n_scn)
    # single calculation on the permanent state (diverges for the radial grid with limits -- in the oracle as well)
    ref1 = orc.Model(inp).calculate(sym=True)
    if ref1["n_failed"] == 0:
        single = model.calculate_power_flow(symmetric=True)
        assert model.n_iter[0] == ref1["n_iter"][0]
        _compare_with_oracle({k: v[None] for k, v in single.items()}, ref1, 1)
    else:
        with pytest.raises(pgm_b200.BatchError, match="Iteration failed to converge"):
            model.calculate_power_flow(symmetric=True)
    # a regulator update in the batch (u_ref, status): scenario-by-scenario route
    upd2 = {"voltage_regulator": pgm_b200.structs.initialize_array("update", "voltage_regulator", (4, 2))}
    upd2["voltage_regulator"]["id"] = inp["voltage_regulator"]["id"][[2, 3]]
    upd2["voltage_regulator"]["u_ref"][:, 0] = [1.0, 1.01, 1.02, np.nan]
    upd2["voltage_regulator"]["status"][:, 1] = [0, 1, 0, -128]
    res2 = model.calculate_power_flow(symmetric=True, update_data=upd2, continue_on_batch_error=True)
    ref2 = orc.Model(inp).calculate(sym=True, update=upd2, threading=0)
    ok = ref2["status"] == 0 if "status" in ref2 else np.ones(4, bool)
    assert np.array_equal(model.status == 0, ok)
    assert np.array_equal(model.n_iter[ok], ref2["n_iter"][ok])
    if ok.any():
        _compare_with_oracle({k: v[ok] for k, v in res2.items()}, {k: v[ok] for k, v in ref2.items() if k in res2}, int(ok.sum()))
    assert ok.any() or q_lim is not None  # without limits the regulator-update scenarios converge
    with pytest.raises(pgm_b200.PgmB200Error):
        model.calculate_power_flow(symmetric=True, calculation_method="iterative_current")


def test_voltage_regulators_asymmetric_match_oracle():
    import grids

    option = dict(n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5, n_mv_feeder=3)
    n_scn = 6
    _, inp, update = grids.regulated_benchmark_grid(True, n_gen=6, option=option, n_scn=n_scn)
    model = pgm_b200.PowerGridModel(inp)
    res = model.calculate_power_flow(symmetric=False, update_data=update)
    ref = orc.Model(inp).calculate(sym=False, update=update, threading=0)
    assert ref["n_failed"] == 0
    assert np.array_equal(model.n_iter, ref["n_iter"]), (model.n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, n_scn)


@pytest.mark.parametrize("method", ["iterative_current", "linear", "linear_current"])
def test_asymmetric_other_methods_match_oracle(method):
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5,
                                  n_mv_feeder=3, has_mv_ring=True, has_lv_ring=True)
    n_scn = 9
    update = grid.batch_update(n_scn, seed=2)
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(symmetric=False, update_data=update, calculation_method=method)
    ref = orc.Model(grid.input_data).calculate(sym=False, update=update, threading=0, method=method)
    assert ref["n_failed"] == 0
    assert np.array_equal(model.n_iter, ref["n_iter"]), (model.n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, n_scn)


@pytest.mark.parametrize("sym", [True, False])
def test_wide_rows_of_the_ringed_benchmark_grid(monkeypatch, sym):
    """the ringed 1804-bus grid has hub rows with ~400 lower entries that the whole thread block eliminates together
    (symbolic.hpp WideRowPlan); results must equal the one-thread-per-row elimination bit for bit, and the oracle"""
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **pgm_b200.BENCHMARK_OPTION)
    n_scn = 6
    update = grid.batch_update(n_scn, seed=3)
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(symmetric=sym, update_data=update)
    n_iter = model.n_iter.copy()
    ref = orc.Model(grid.input_data).calculate(sym=sym, update=update, threading=0)
    assert ref["n_failed"] == 0 and np.array_equal(n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, n_scn)
    monkeypatch.setenv("PGMB_WIDE", "0")
    plain = pgm_b200.PowerGridModel(grid.input_data).calculate_power_flow(symmetric=sym, update_data=update)
    for comp in res:
        for name in res[comp].dtype.names:
            assert np.array_equal(res[comp][name], plain[comp][name], equal_nan=True), (comp, name)


def test_batch_properties_at_full_size():
    """config 2 at its full size (1000 scenarios): properties that need no oracle run per scenario:
    power balance (sum of node injections == losses), all scenarios converge, scenario order independence."""
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    update = grid.batch_update(1000, seed=0)
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(symmetric=True, update_data=update)
    assert (model.status == 0).all() and model.n_iter.min() >= 1 and model.n_iter.max() <= 20
    # losses = sum of branch p_from + p_to >= 0 and equals minus the sum of all appliance powers at the nodes
    loss = res["line"]["p_from"].sum(1) + res["line"]["p_to"].sum(1) + res["transformer"]["p_from"].sum(1) + res["transformer"]["p_to"].sum(1)
    node_p = res["node"]["p"].sum(1)
    shunt_p = res["shunt"]["p"].sum(1)
    assert (loss > 0).all()
    assert np.max(np.abs(node_p - loss - shunt_p) / np.abs(res["source"]["p"][:, 0])) < 1e-6
    # reversed scenario order gives the reversed results, bit for bit
    rev = {k: np.ascontiguousarray(v[::-1]) for k, v in update.items()}
    res_rev = model.calculate_power_flow(symmetric=True, update_data=rev, output_component_types=["node"])
    assert np.array_equal(res_rev["node"]["u_pu"][::-1], res["node"]["u_pu"])


def test_failed_scenarios_are_isolated():
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    update = grid.batch_update(20, seed=0)
    good = pgm_b200.PowerGridModel(grid.input_data).calculate_power_flow(update_data=update, output_component_types=["node"])
    update["sym_load"]["p_specified"][7] *= 1e6  # absurd load: diverges
    model = pgm_b200.PowerGridModel(grid.input_data)
    with pytest.raises(pgm_b200.BatchError) as e:
        model.calculate_power_flow(update_data=update, output_component_types=["node"])
    assert "Error in batch #7" in str(e.value)
    res = model.calculate_power_flow(update_data=update, output_component_types=["node"], continue_on_batch_error=True)
    assert model.status[7] != 0 and (np.delete(model.status, 7) == 0).all()
    keep = np.arange(20) != 7
    assert np.array_equal(res["node"]["u_pu"][keep], good["node"]["u_pu"][keep])


def test_pinned_buffers_and_reused_outputs():
    """page-locked update buffers + model-owned page-locked outputs (the overlapped pipeline) give the same bytes as plain
    numpy buffers (single chunk)"""
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    update = grid.batch_update(300, seed=5)
    model = pgm_b200.PowerGridModel(grid.input_data)
    plain = model.calculate_power_flow(update_data=update)
    pinned = {}
    for k, v in update.items():
        pinned[k] = pgm_b200.pinned_empty(v.shape, v.dtype)
        pinned[k][...] = v
    first = model.calculate_power_flow(update_data=pinned, reuse_output_buffers=True)
    for comp in plain:
        for name in plain[comp].dtype.names:
            assert np.array_equal(plain[comp][name], first[comp][name], equal_nan=True), (comp, name)
    again = model.calculate_power_flow(update_data=pinned, reuse_output_buffers=True)
    assert all(again[c] is first[c] or again[c].ctypes.data == first[c].ctypes.data for c in first)


def test_mixed_outage_batch_matches_oracle():
    """Shared-pattern outages mixed with scenarios that need their own topology: a transformer outage that de-energises a
    whole LV grid (bridge), two lines at once, a tap change together with a switched line -- asymmetric and symmetric."""
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5,
                                  n_mv_feeder=3, has_mv_ring=True, has_lv_ring=True)
    lines, trafos = grid.input_data["line"], grid.input_data["transformer"]
    n_scn = 10
    line_rows, trafo_rows = [], []
    for s in range(n_scn):
        lu = pgm_b200.structs.initialize_array("update", "line", 2 if s == 4 else 1)
        lu["id"] = lines["id"][[(11 * s + 3) % len(lines), (11 * s + 17) % len(lines)][: len(lu)]]
        lu["from_status"] = 0
        lu["to_status"] = 0
        tu = pgm_b200.structs.initialize_array("update", "transformer", 1 if s in (2, 6) else 0)
        if s == 2:  # bridge: the source transformer -- everything behind it loses its supply
            tu["id"] = trafos["id"][0]
            tu["from_status"] = 0
            tu["to_status"] = 0
            lu = lu[:0]
        if s == 6:
            tu["id"] = trafos["id"][0]
            tu["tap_pos"] = trafos["tap_pos"][0] + 1
        line_rows.append(lu)
        trafo_rows.append(tu)
    update = {"line": {"data": np.concatenate(line_rows), "indptr": np.cumsum([0] + [len(x) for x in line_rows])},
              "transformer": {"data": np.concatenate(trafo_rows), "indptr": np.cumsum([0] + [len(x) for x in trafo_rows])}}
    for sym in (True, False):
        model = pgm_b200.PowerGridModel(grid.input_data)
        res = model.calculate_power_flow(symmetric=sym, update_data=update)
        ref = orc.Model(grid.input_data).calculate(sym=sym, update=update, threading=0)
        assert ref["n_failed"] == 0 and np.array_equal(model.n_iter, ref["n_iter"]), (model.n_iter, ref["n_iter"])
        assert (ref["node"]["energized"][2] == 0).any() and (res["node"]["energized"][2] == ref["node"]["energized"][2]).all()
        _compare_with_oracle(res, ref, n_scn)


@pytest.mark.parametrize("sym", [True, False])
def test_outages_with_a_load_profile_match_oracle(sym):
    """contingency x load profile: every scenario switches a line AND carries its own loads; the outages share the base
    pattern, the load updates are applied on the device like in any load batch"""
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5,
                                  n_mv_feeder=3, has_mv_ring=True, has_lv_ring=True)
    n_scn = 14
    update = grid.batch_update(n_scn, seed=4)
    lines = grid.input_data["line"]
    upd = pgm_b200.structs.initialize_array("update", "line", (n_scn, 1))
    upd["id"][:, 0] = lines["id"][(np.arange(n_scn) * 9 + 2) % len(lines)]
    upd["from_status"][:, 0] = 0
    upd["to_status"][:, 0] = 0
    update["line"] = upd
    model = pgm_b200.PowerGridModel(grid.input_data)
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    res = model.calculate_power_flow(symmetric=sym, update_data=update)
    assert int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0 < 4 * n_scn
    ref = orc.Model(grid.input_data).calculate(sym=sym, update=update, threading=0)
    assert ref["n_failed"] == 0 and np.array_equal(model.n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, n_scn)


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("threads", [1, 4])
def test_branch_switching_batch_on_host_threads(threads, exact, monkeypatch):
    """N-1 style batch (config 5 shape): every scenario switches another line off.  exact=True (PGMB_N1_EXACT): every scenario
    gets its own topology like in the reference, dispatched over host threads with one model copy each
    (job_dispatch.hpp:88-160).  exact=False (default): scenarios whose branch is not a bridge share the base pattern (branch
    outage overlay), the others take the exact route.  Either way results and per-scenario errors must equal the oracle's."""
    if exact:
        monkeypatch.setenv("PGMB_N1_EXACT", "1")
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5,
                                  n_mv_feeder=3, has_mv_ring=True, has_lv_ring=True)
    lines = grid.input_data["line"]
    n_scn = 12
    upd = pgm_b200.structs.initialize_array("update", "line", (n_scn, 1))
    upd["id"][:, 0] = lines["id"][np.arange(n_scn) * 7 % len(lines)]
    upd["from_status"][:, 0] = 0
    upd["to_status"][:, 0] = 0
    upd["to_status"][3, 0] = 1      # one side stays connected
    upd["from_status"][5, 0] = -128  # only the to side opens
    upd["from_status"][7, 0] = 1     # no change at all
    upd["to_status"][7, 0] = 1
    model = pgm_b200.PowerGridModel(grid.input_data)
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    res = model.calculate_power_flow(update_data={"line": upd}, threading=threads)
    launches = int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0
    ref = orc.Model(grid.input_data).calculate(sym=True, update={"line": upd}, threading=0)
    assert ref["n_failed"] == 0 and np.array_equal(model.n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, n_scn)
    if not exact:
        assert launches < 4 * n_scn, launches  # one shared-pattern batch, not a solver launch per scenario
    base = model.calculate_power_flow()  # the model itself is untouched by the batch
    ref0 = orc.Model(grid.input_data).calculate(sym=True)
    _compare_with_oracle({k: v[None] for k, v in base.items()}, ref0, 1)


def test_config3_properties_at_full_size():
    """BASELINE config 3 at its full size (ringed grid, asymmetric NR, 1000 scenarios): every scenario converges, the
    three-phase power balance closes (sum of node injections = branch losses + shunt consumption per phase), a sample of
    scenarios equals the oracle, and the reversed batch gives the reversed results bit for bit."""
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **pgm_b200.BENCHMARK_OPTION)
    update = grid.batch_update(1000, seed=0)
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(symmetric=False, update_data=update)
    assert (model.status == 0).all() and 1 <= model.n_iter.min() and model.n_iter.max() <= 20
    loss = sum(res[c][k].sum(1) for c in ("line", "transformer") for k in ("p_from", "p_to"))  # (n_scn, 3)
    assert (loss.sum(1) > 0).all()
    balance = res["node"]["p"].sum(1) - loss - res["shunt"]["p"].sum(1)
    assert np.max(np.abs(balance)) / np.max(np.abs(res["source"]["p"])) < 1e-6
    pick = [0, 333, 999]
    sample = {k: np.ascontiguousarray(v[pick]) for k, v in update.items()}
    ref = orc.Model(grid.input_data).calculate(sym=False, update=sample, threading=0)
    assert np.array_equal(model.n_iter[pick], ref["n_iter"])
    _compare_with_oracle({k: v[pick] for k, v in res.items()}, ref, len(pick))
    rev = {k: np.ascontiguousarray(v[::-1]) for k, v in update.items()}
    res_rev = model.calculate_power_flow(symmetric=False, update_data=rev, output_component_types=["node"])
    assert np.array_equal(res_rev["node"]["u_pu"][::-1], res["node"]["u_pu"])


@pytest.mark.parametrize("method", ["iterative_current", "linear"])
def test_config4_properties_at_large_batch(method):
    """BASELINE config 4 shape (radial grid, iterative current / linear, a large time series): all scenarios converge, a
    sample equals the oracle (iteration counts included), chunked and unchunked execution agree bit for bit"""
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    n_scn = 4000
    update = grid.batch_update(n_scn, seed=4)
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(update_data=update, calculation_method=method, output_component_types=["node", "source"])
    assert (model.status == 0).all()
    pick = [0, 1777, 3999]
    sample = {k: np.ascontiguousarray(v[pick]) for k, v in update.items()}
    ref = orc.Model(grid.input_data).calculate(sym=True, update=sample, threading=0, method=method, output_components=["node", "source"])
    assert np.array_equal(model.n_iter[pick], ref["n_iter"])
    _compare_with_oracle({k: v[pick] for k, v in res.items()}, ref, len(pick))
    head = {k: np.ascontiguousarray(v[:500]) for k, v in update.items()}
    part = model.calculate_power_flow(update_data=head, calculation_method=method, output_component_types=["node"])
    assert np.array_equal(part["node"]["u_pu"], res["node"]["u_pu"][:500])


@pytest.mark.parametrize("sym,n_scn", [(True, 16), (False, 4)])
def test_50k_node_ringed_grid_matches_oracle(sym, n_scn):
    """the grid of BASELINE config 5 (53 068 nodes, MV and LV rings, 26 628 fill-ins, wide hub rows): load-profile batch
    against the oracle -- index arithmetic and scratch sizing at scale"""
    opt = dict(pgm_b200.BENCHMARK_OPTION)
    opt["n_node_total_specified"] = 50000
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **opt)
    update = grid.batch_update(n_scn, seed=1)
    model = pgm_b200.PowerGridModel(grid.input_data)
    res = model.calculate_power_flow(symmetric=sym, update_data=update, output_component_types=["node", "line", "source"])
    assert (model.status == 0).all()
    pick = [0, n_scn - 1]
    sample = {k: np.ascontiguousarray(v[pick]) for k, v in update.items()}
    ref = orc.Model(grid.input_data).calculate(sym=sym, update=sample, threading=0, output_components=["node", "line", "source"])
    assert np.array_equal(model.n_iter[pick], ref["n_iter"])
    _compare_with_oracle({k: v[pick] for k, v in res.items()}, ref, len(pick))


def test_config5_n1_batch_on_the_50k_node_grid():
    """BASELINE config 5 shape: asymmetric N-1 batch on the 53 068-node ringed grid.  All scenarios share the base grid's
    symbolic pattern (branch-outage overlay, bridge outages with masked buses).  Checked: (1) a sample of scenarios against
    the oracle, which rebuilds the topology per scenario like the reference (iteration counts identical, results within the
    north_star tolerances); (2) at full batch size the properties of an outage: the switched line carries nothing and is not
    energised, every scenario converges, one kernel batch instead of a solver launch per scenario."""
    opt = dict(pgm_b200.BENCHMARK_OPTION)
    opt["n_node_total_specified"] = 50000
    grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **opt)
    lines = grid.input_data["line"]
    n_scn = 96
    pick_lines = np.random.default_rng(5).choice(len(lines), n_scn, replace=False)
    upd = pgm_b200.structs.initialize_array("update", "line", (n_scn, 1))
    upd["id"][:, 0] = lines["id"][pick_lines]
    upd["from_status"][:, 0] = 0
    upd["to_status"][:, 0] = 0
    model = pgm_b200.PowerGridModel(grid.input_data)
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    res = model.calculate_power_flow(symmetric=False, update_data={"line": upd}, output_component_types=["node", "line", "asym_load"])
    launches = int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0
    assert (model.status == 0).all()
    assert launches < n_scn, launches
    rows = np.arange(n_scn)
    switched = res["line"][rows, pick_lines]
    assert (switched["energized"] == 0).all()
    for name in ("p_from", "q_from", "i_from", "s_from", "p_to", "q_to", "i_to", "s_to", "loading"):
        assert np.all(switched[name] == 0.0), name
    dark = res["node"]["energized"] == 0
    assert dark.any() and not dark.all(axis=1).any()  # some outages are bridges; no scenario loses the whole grid
    assert np.all(res["node"]["u_pu"][dark] == 0.0) and np.all(res["node"]["u_pu"][~dark] > 0.5)
    pick = [0, 17, int(np.argmax(dark.sum(axis=1)))]  # the last one is the bridge that darkens most nodes
    sample = {"line": np.ascontiguousarray(upd[pick])}
    ref = orc.Model(grid.input_data).calculate(sym=False, update=sample, threading=0, output_components=["node", "line", "asym_load"])
    assert ref["n_failed"] == 0 and np.array_equal(model.n_iter[pick], ref["n_iter"])
    _compare_with_oracle({k: v[pick] for k, v in res.items()}, ref, len(pick))


def test_batch_larger_than_the_memory_budget_is_split(monkeypatch):
    """a batch whose working set exceeds the device budget runs in parts over offset views of the caller's buffers"""
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5,
                                  n_mv_feeder=3)
    update = grid.batch_update(150, seed=6)
    src = pgm_b200.structs.initialize_array("update", "source", (150, 1))
    src["id"][:, 0] = grid.input_data["source"]["id"][0]
    src["u_ref"][:, 0] = 1.0 + 0.0002 * np.arange(150)
    update["source"] = src
    model = pgm_b200.PowerGridModel(grid.input_data)
    whole = model.calculate_power_flow(update_data=update)
    monkeypatch.setenv("PGMB_MAX_BATCH_BYTES", "4000000")  # ~ 64 scenarios per part
    model2 = pgm_b200.PowerGridModel(grid.input_data)
    parts = model2.calculate_power_flow(update_data=update)
    assert np.array_equal(model.n_iter, model2.n_iter)
    for comp in whole:
        for name in whole[comp].dtype.names:
            assert np.array_equal(whole[comp][name], parts[comp][name], equal_nan=True), (comp, name)
    update["sym_load"]["p_specified"][101] *= 1e6  # the failing scenario keeps its batch position in the message
    with pytest.raises(pgm_b200.BatchError) as e:
        model2.calculate_power_flow(update_data=update)
    assert "Error in batch #101" in str(e.value)


def test_edge_batches():
    """empty batch, single-scenario batch, partial updates (a subset of the loads, status-only rows, NaN = keep)"""
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=60, n_connection_per_lv_feeder=3, n_lv_feeder=2, n_node_per_mv_feeder=3,
                                  n_mv_feeder=2)
    inp = grid.input_data
    model = pgm_b200.PowerGridModel(inp)
    base = model.calculate_power_flow()
    full = grid.batch_update(4, seed=9)
    empty = {k: v[:0] for k, v in full.items()}
    res0 = model.calculate_power_flow(update_data=empty)
    assert all(a.shape[0] == 0 for a in res0.values())
    one = {k: np.ascontiguousarray(v[:1]) for k, v in full.items()}
    ref = orc.Model(inp).calculate(sym=True, update=one)
    _compare_with_oracle(model.calculate_power_flow(update_data=one), ref, 1)
    # partial: two sym loads only; scenario 0 switches one off, scenario 1 keeps p (NaN) and changes q, scenario 2 all NaN
    ids = inp["sym_load"]["id"][:2]
    upd = pgm_b200.structs.initialize_array("update", "sym_load", (3, 2))
    upd["id"][:] = ids
    upd["status"][0, 0] = 0
    upd["q_specified"][1, 1] = 1234.5
    ref = orc.Model(inp).calculate(sym=True, update={"sym_load": upd})
    res = model.calculate_power_flow(update_data={"sym_load": upd})
    assert np.array_equal(model.n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, 3)
    assert np.array_equal(res["node"]["u_pu"][2], base["node"]["u_pu"])  # all-NaN scenario = the base state
    assert res["sym_load"]["energized"][0, 0] == 0 and res["sym_load"]["p"][0, 0] == 0.0


@pytest.mark.parametrize("rings", [False, True])
@pytest.mark.parametrize("q_lim", [None, 2e5])
def test_regulated_grid_symmetric_kernels_equal_the_block_kernel(q_lim, rings, monkeypatch):
    """symmetric grids with voltage regulators run the PV instantiations of the path kernel (radial grids, nr_sym_v3) and the
    level kernel (meshed grids, nr_sym_v2); PGMB_REG_PATH=0 selects the generic block kernel with B = 1.  Same decisions in the
    same order: identical iteration counts, limit flags and values."""
    import grids

    n_scn = 64
    _, inp, update = grids.regulated_benchmark_grid(rings, q_lim=q_lim, n_scn=n_scn)
    a = pgm_b200.PowerGridModel(inp)
    res_path = a.calculate_power_flow(symmetric=True, update_data=update, continue_on_batch_error=True)
    st_path, it_path, t_path = a.status.copy(), a.n_iter.copy(), a.timing()["solve_kernel"]
    monkeypatch.setenv("PGMB_REG_PATH", "0")
    b = pgm_b200.PowerGridModel(inp)
    res_block = b.calculate_power_flow(symmetric=True, update_data=update, continue_on_batch_error=True)
    assert np.array_equal(st_path, b.status) and np.array_equal(it_path, b.n_iter)
    ok = st_path == 0
    assert ok.any()
    for c in res_path:
        for f in res_path[c].dtype.names:
            x, y = res_path[c][f][ok], res_block[c][f][ok]
            assert np.array_equal(x, y, equal_nan=x.dtype.kind == "f"), (c, f, float(np.nanmax(np.abs(x - y))) if x.dtype.kind == "f" else None)
    print(f"regulated {'ringed' if rings else 'radial'} grid, {n_scn} scenarios: {'level' if rings else 'path'} kernel {t_path:.2f} ms, "
          f"block kernel {b.timing()['solve_kernel']:.2f} ms")


@pytest.mark.parametrize("sym", [True, False])
def test_scenarios_with_the_same_switching_state_run_as_one_group(sym, monkeypatch):
    """a batch that revisits a few switching states (two lines out / a transformer tap moved / a shunt off) with different loads:
    scenarios with byte-identical structural updates share one topology + engine build and run as ONE device batch; the
    scenario-by-scenario route (PGMB_GROUP_SCENARIOS=0) and the oracle give the same results.  (These particular states -- closed
    lines opened, a tap moved -- also fit the branch overlay, which takes the whole batch as one; PGMB_N1_EXACT=1 keeps it out so
    that the grouped route is what runs here.  The overlay's own result is compared at the end.)"""
    monkeypatch.setenv("PGMB_N1_EXACT", "1")
    grid = pgm_b200.FictionalGrid(seed=2, has_mv_ring=True, has_lv_ring=True, n_node_total_specified=300, n_mv_feeder=3,
                                  n_node_per_mv_feeder=5, n_lv_feeder=3, n_connection_per_lv_feeder=10)
    n_scn, n_state = 48, 4
    update = dict(grid.batch_update(n_scn, seed=4))
    lines, trafos = grid.input_data["line"], grid.input_data["transformer"]
    line_upd = pgm_b200.structs.initialize_array("update", "line", (n_scn, 2))
    trafo_upd = pgm_b200.structs.initialize_array("update", "transformer", (n_scn, 1))
    ring_lines = lines["id"][-4:]  # ring closures: switching them keeps the grid connected
    for s in range(n_scn):
        state = s % n_state
        line_upd["id"][s] = ring_lines[[0, 1]] if state < 2 else ring_lines[[2, 3]]
        line_upd["from_status"][s] = 0 if state in (0, 2) else 1
        line_upd["to_status"][s] = 0 if state in (0, 2) else 1
        trafo_upd["id"][s, 0] = trafos["id"][1]
        trafo_upd["tap_pos"][s, 0] = [0, 1, -1, 2][state]
    update["line"], update["transformer"] = line_upd, trafo_upd
    ref = orc.Model(grid.input_data).calculate(sym=sym, update=update, threading=0)
    assert ref["n_failed"] == 0
    model = pgm_b200.PowerGridModel(grid.input_data)
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    grouped = model.calculate_power_flow(symmetric=sym, update_data=update)
    launches_grouped = int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0
    assert (model.status == 0).all() and np.array_equal(model.n_iter, ref["n_iter"])
    _compare_with_oracle(grouped, ref, n_scn)
    monkeypatch.setenv("PGMB_GROUP_SCENARIOS", "0")
    single = model.calculate_power_flow(symmetric=sym, update_data=update)
    launches_single = int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0 - launches_grouped
    assert np.array_equal(model.n_iter, ref["n_iter"])
    _compare_with_oracle(single, ref, n_scn)
    assert launches_grouped < launches_single
    monkeypatch.delenv("PGMB_N1_EXACT")
    monkeypatch.delenv("PGMB_GROUP_SCENARIOS")
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    overlay = model.calculate_power_flow(symmetric=sym, update_data=update)
    assert int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0 < launches_grouped
    assert np.array_equal(model.n_iter, ref["n_iter"])
    _compare_with_oracle(overlay, ref, n_scn)
    # the model is unchanged afterwards
    base = model.calculate_power_flow(symmetric=sym)
    base_ref = orc.Model(grid.input_data).calculate(sym=sym)
    _compare_with_oracle({k: v[None] for k, v in base.items()}, base_ref, 1)


def _multi_outage_update(grid, n_scn, seed):
    """every scenario opens 2-4 branches (lines and transformers, some only on one side); scenario 0: two lines that share a node
    (one Y-bus diagonal entry replaced once for both), scenario 1: a line and the transformer it hangs on"""
    rng = np.random.default_rng(seed)
    lines, trafos = grid.input_data["line"], grid.input_data["transformer"]
    line_rows, trafo_rows = [], []
    for s in range(n_scn):
        k = int(rng.integers(2, 5))
        n_t = int(rng.integers(0, 2)) if s != 0 else 0
        li = rng.choice(len(lines), size=k - n_t, replace=False)
        if s == 0:
            node = lines["from_node"][5]
            li = np.flatnonzero((lines["from_node"] == node) | (lines["to_node"] == node))[:2]
            assert len(li) == 2
        lu = pgm_b200.structs.initialize_array("update", "line", len(li))
        lu["id"] = lines["id"][li]
        lu["from_status"] = 0
        lu["to_status"] = rng.choice([0, 0, 0, 1, -128], size=len(li))
        tu = pgm_b200.structs.initialize_array("update", "transformer", n_t)
        if n_t:
            tu["id"] = trafos["id"][rng.integers(1, len(trafos))]
            tu["from_status"] = 0
            tu["to_status"] = 0
        line_rows.append(lu)
        trafo_rows.append(tu)
    return {"line": {"data": np.concatenate(line_rows), "indptr": np.cumsum([0] + [len(x) for x in line_rows])},
            "transformer": {"data": np.concatenate(trafo_rows), "indptr": np.cumsum([0] + [len(x) for x in trafo_rows])}}


@pytest.mark.parametrize("parts", [False, True])
@pytest.mark.parametrize("sym", [True, False])
def test_multi_branch_outages_share_the_base_pattern(sym, parts, monkeypatch):
    """N-k batches (2-4 switched branches per scenario, with a load profile): the scenarios run as ONE device batch on the base
    grid's pattern (overlay with several branch slots per scenario; parts of the grid that lose their supply are masked) and
    equal the oracle, which rebuilds the topology per scenario; the same batch with one slot (PGMB_OUTAGE_SLOTS=1: every
    scenario on its own topology) gives the same results.  parts=True: a device-memory budget that splits the batch into two
    passes (32 + 8 scenarios), each with its slice of the overlay."""
    if parts:
        monkeypatch.setenv("PGMB_MAX_BATCH_BYTES", "1")
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5,
                                  n_mv_feeder=3, has_mv_ring=True, has_lv_ring=True)
    n_scn = 40
    update = grid.batch_update(n_scn, seed=2)
    update.update(_multi_outage_update(grid, n_scn, seed=21))
    model = pgm_b200.PowerGridModel(grid.input_data)
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    res = model.calculate_power_flow(symmetric=sym, update_data=update)
    launches = int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0
    n_iter = model.n_iter.copy()
    assert launches < 2 * n_scn, launches  # one batch, not a solve per scenario
    ref = orc.Model(grid.input_data).calculate(sym=sym, update=update, threading=0)
    assert ref["n_failed"] == 0
    dark = (ref["node"]["energized"] == 0).any(axis=1)
    assert 0 < dark.sum() < n_scn  # some scenarios cut a part of the grid off, some do not
    assert np.array_equal(n_iter, ref["n_iter"]), (n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, n_scn)
    monkeypatch.setenv("PGMB_OUTAGE_SLOTS", "1")
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    one = pgm_b200.PowerGridModel(grid.input_data).calculate_power_flow(symmetric=sym, update_data=update)
    assert int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0 > 2 * n_scn
    _compare_with_oracle(one, ref, n_scn)


@pytest.mark.parametrize("sym", [True, False])
def test_scenarios_that_close_open_branches_run_on_the_union_grid(sym, monkeypatch):
    """Reconfiguration batch: the base grid has three open tie lines; scenarios close one or two of them (fully or on one side),
    open other lines at the same time, or change nothing.  The base pattern has no entries for the open ties, so the batch runs
    as ONE device batch on the union grid (all ties closed) with the not-closed ties switched off per scenario through the
    N-k overlay; results equal the oracle, which rebuilds the topology per scenario."""
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5,
                                  n_mv_feeder=3, has_mv_ring=True, has_lv_ring=True)
    data = {k: v.copy() for k, v in grid.input_data.items()}
    lines = data["line"]
    not_bridge = np.flatnonzero(np.asarray(pgm_b200.PowerGridModel(data).math_index(0, "branch_is_bridge"))[: len(lines)] == 0)
    # three ties from different rings (removing all of them keeps the grid connected: checked through the oracle below)
    ties = not_bridge[[0, len(not_bridge) // 2, len(not_bridge) - 1]]
    lines["from_status"][ties[0]] = 0
    lines["to_status"][ties[0]] = 0
    lines["from_status"][ties[1]] = 0  # half open in the base state
    lines["to_status"][ties[2]] = 0
    lines["from_status"][ties[2]] = 0
    n_scn = 24
    rng = np.random.default_rng(8)
    rows = []
    for s in range(n_scn):
        kind = s % 6
        lu = pgm_b200.structs.initialize_array("update", "line", {0: 1, 1: 2, 2: 2, 3: 0, 4: 1, 5: 1}[kind])
        if kind == 0:    # close one tie
            lu["id"], lu["from_status"], lu["to_status"] = lines["id"][ties[s % 3]], 1, 1
        elif kind == 1:  # close a tie and open another line
            lu["id"][0], lu["from_status"][0], lu["to_status"][0] = lines["id"][ties[s % 3]], 1, 1
            other = rng.choice(np.setdiff1d(np.arange(len(lines)), ties))
            lu["id"][1], lu["from_status"][1], lu["to_status"][1] = lines["id"][other], 0, 0
        elif kind == 2:  # close two ties
            lu["id"] = lines["id"][ties[[s % 3, (s + 1) % 3]]]
            lu["from_status"], lu["to_status"] = 1, 1
        elif kind == 4:  # one side only: a tie that was fully open stays open, the half-open one closes
            lu["id"], lu["from_status"], lu["to_status"] = lines["id"][ties[s % 3]], 1, -128
        elif kind == 5:  # open a line, no tie closed
            lu["id"], lu["from_status"], lu["to_status"] = lines["id"][rng.choice(np.setdiff1d(np.arange(len(lines)), ties))], 0, 0
        rows.append(lu)
    update = grid.batch_update(n_scn, seed=3)
    update["line"] = {"data": np.concatenate(rows), "indptr": np.cumsum([0] + [len(x) for x in rows])}
    ref = orc.Model(data).calculate(sym=sym, update=update, threading=0)
    assert ref["n_failed"] == 0
    model = pgm_b200.PowerGridModel(data)
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    res = model.calculate_power_flow(symmetric=sym, update_data=update)
    launches = int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0
    assert launches < 2 * n_scn, launches
    assert np.array_equal(model.n_iter, ref["n_iter"]), (model.n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, n_scn)
    # the model itself is unchanged: the base state afterwards, and the same batch again (the union copy is reused)
    single = model.calculate_power_flow(symmetric=sym)
    _compare_with_oracle({k: v[None] for k, v in single.items()}, orc.Model(data).calculate(sym=sym), 1)
    again = model.calculate_power_flow(symmetric=sym, update_data=update)
    for comp in res:
        for name in res[comp].dtype.names:
            assert np.array_equal(res[comp][name], again[comp][name], equal_nan=True), (comp, name)
    monkeypatch.setenv("PGMB_NO_UNION_GRID", "1")
    one = pgm_b200.PowerGridModel(data).calculate_power_flow(symmetric=sym, update_data=update)
    _compare_with_oracle(one, ref, n_scn)


@pytest.mark.parametrize("sym", [True, False])
def test_tap_positions_per_scenario_share_the_base_pattern(sym, monkeypatch):
    """A batch that moves transformer taps per scenario (1-3 transformers each, positions beyond the range are clamped, some
    scenarios open a line as well, all with their own loads): a tap position is one more set of branch parameters, so the scenarios run as ONE device batch
    through the branch overlay; equal to the oracle (per-scenario parameter rebuild) and to the own-topology route."""
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5,
                                  n_mv_feeder=3, has_mv_ring=True, has_lv_ring=True)
    trafos, lines = grid.input_data["transformer"], grid.input_data["line"]
    n_scn = 30
    rng = np.random.default_rng(12)
    t_rows, l_rows = [], []
    for s in range(n_scn):
        pick = rng.choice(len(trafos), size=int(rng.integers(1, 4)), replace=False)
        if s == 0:
            pick = np.array([0])  # the source transformer: a bridge whose parameters change while it stays connected
        tu = pgm_b200.structs.initialize_array("update", "transformer", len(pick))
        tu["id"] = trafos["id"][pick]
        tu["tap_pos"] = [rng.integers(int(min(trafos["tap_min"][k], trafos["tap_max"][k])) - 1, int(max(trafos["tap_min"][k], trafos["tap_max"][k])) + 2) for k in pick]
        lu = pgm_b200.structs.initialize_array("update", "line", 1 if s % 4 == 1 else 0)
        if len(lu):
            lu["id"], lu["from_status"], lu["to_status"] = lines["id"][rng.integers(len(lines))], 0, 0
        t_rows.append(tu)
        l_rows.append(lu)
    update = grid.batch_update(n_scn, seed=6)
    update["transformer"] = {"data": np.concatenate(t_rows), "indptr": np.cumsum([0] + [len(x) for x in t_rows])}
    update["line"] = {"data": np.concatenate(l_rows), "indptr": np.cumsum([0] + [len(x) for x in l_rows])}
    ref = orc.Model(grid.input_data).calculate(sym=sym, update=update, threading=0)
    assert ref["n_failed"] == 0
    model = pgm_b200.PowerGridModel(grid.input_data)
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    res = model.calculate_power_flow(symmetric=sym, update_data=update)
    assert int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0 < 2 * n_scn
    assert np.array_equal(model.n_iter, ref["n_iter"]), (model.n_iter, ref["n_iter"])
    _compare_with_oracle(res, ref, n_scn)
    monkeypatch.setenv("PGMB_N1_EXACT", "1")
    one = pgm_b200.PowerGridModel(grid.input_data).calculate_power_flow(symmetric=sym, update_data=update)
    _compare_with_oracle(one, ref, n_scn)
