"""GPU tests through the reference's own C API names (PGM_create_model / PGM_update_model / PGM_copy_model / PGM_calculate of
libpgm_b200.so, driven by pgm_b200.pgm_core the way the reference's Python wrapper drives its core):
 (a) the scenarios of the reference's tests/native_api_tests/test_api_model.cpp with their closed-form answers,
 (b) the reference's power-flow validation cases (tests/golden/power_flow_cases.json) with the reference's tolerances,
 (c) the 1000-scenario benchmark batch: the same bytes as the pgmb_model_* seam produces (one engine behind both)."""
import numpy as np
import pytest

import pgm_b200
import validation_cases as vc
from pgm_b200 import pgm_core
from pgm_b200.structs import initialize_array
from test_pgm_capi import _api_model_input

pytestmark = pytest.mark.gpu


def _update(source_u_ref=None, q=None, line_status=((0, 1), (0, 0))):
    upd = {}
    if source_u_ref is not None:
        s = initialize_array("update", "source", 1)
        s["id"], s["u_ref"] = 1, source_u_ref
        upd["source"] = s
    if q is not None:
        l = initialize_array("update", "sym_load", 1)
        l["id"], l["q_specified"] = 2, q
        upd["sym_load"] = l
    ln = initialize_array("update", "line", 2)
    ln["id"] = [5, 6]
    ln["from_status"], ln["to_status"] = [line_status[0][0], line_status[1][0]], [line_status[0][1], line_status[1][1]]
    upd["line"] = ln
    return upd


def _check_node(node, u0):
    assert node["id"].tolist() == [0, 4] and node["energized"].tolist() == [1, 0]
    assert node["u"][0] == pytest.approx(u0) and node["u_pu"][0] == pytest.approx(u0 / 100.0)
    assert node["u_angle"][0] == pytest.approx(0.0, abs=1e-12)
    assert node["u"][1] == 0.0 and node["u_pu"][1] == 0.0 and node["u_angle"][1] == 0.0


def test_api_model_single_update_copy():
    """test_api_model.cpp:225-253: 50 V; after the permanent update 40 V; a copy made before the update still gives 50 V"""
    model = pgm_core.PowerGridModel(_api_model_input())
    _check_node(model.calculate_power_flow(output_component_types=["node"])["node"], 50.0)
    copy = model.copy()
    model.update(update_data=_update(source_u_ref=0.5, q=100.0))
    _check_node(model.calculate_power_flow(output_component_types=["node"])["node"], 40.0)
    _check_node(copy.calculate_power_flow(output_component_types=["node"])["node"], 50.0)
    # every valid tap changing strategy is the plain power flow on a model without tap regulators (test_api_model.cpp:462-465)
    _check_node(copy.calculate_power_flow(output_component_types=["node"], tap_changing_strategy=3)["node"], 50.0)


def test_api_model_batch():
    """test_api_model.cpp:289-314: sparse source buffer (only scenario 0 updates it), dense load / line buffers"""
    model = pgm_core.PowerGridModel(_api_model_input())
    s0, s1 = _update(source_u_ref=0.5, q=100.0), _update(q=300.0, line_status=((0, 0), (0, 0)))
    batch = {"source": {"data": s0["source"], "indptr": np.array([0, 1, 1])},
             "sym_load": np.stack([s0["sym_load"], s1["sym_load"]]), "line": np.stack([s0["line"], s1["line"]])}
    res = model.calculate_power_flow(update_data=batch, output_component_types=["node"])["node"]
    assert res.shape == (2, 2)
    _check_node(res[0], 40.0)
    _check_node(res[1], 70.0)
    _check_node(model.calculate_power_flow(output_component_types=["node"])["node"], 50.0)  # model restored


def test_api_model_calculation_errors():
    """test_api_model.cpp:468-546: a single calculation that does not converge is PGM_regular_error; a batch with a bad id in
    scenario 1 is PGM_batch_error with that scenario listed and scenario 0 calculated"""
    model = pgm_core.PowerGridModel(_api_model_input())
    with pytest.raises(pgm_core.PowerGridError, match="Iteration failed to converge after") as e:
        model.calculate_power_flow(max_iterations=1, error_tolerance=1e-100, symmetric=False, threading=1)
    assert not isinstance(e.value, pgm_core.PowerGridBatchError)
    s0, s1 = _update(source_u_ref=0.5, q=100.0), _update(q=300.0)
    s1["line"]["id"] = [99, 6]
    batch = {"source": {"data": s0["source"], "indptr": np.array([0, 1, 1])},
             "sym_load": np.stack([s0["sym_load"], s1["sym_load"]]), "line": np.stack([s0["line"], s1["line"]])}
    with pytest.raises(pgm_core.PowerGridBatchError, match="The id cannot be found:") as e:
        model.calculate_power_flow(update_data=batch, output_component_types=["node"])
    assert e.value.failed_scenarios.tolist() == [1] and "The id cannot be found:" in e.value.error_messages[0]
    res = model.calculate_power_flow(update_data=batch, output_component_types=["node"], continue_on_batch_error=True)
    assert model.batch_error.failed_scenarios.tolist() == [1]
    _check_node(res["node"][0], 40.0)
    # non-converging scenarios of a batch: all listed, with the solver's message
    with pytest.raises(pgm_core.PowerGridBatchError) as e:
        model.calculate_power_flow(update_data={"sym_load": np.stack([s0["sym_load"], s1["sym_load"]])}, max_iterations=1,
                                   error_tolerance=1e-100)
    assert e.value.failed_scenarios.tolist() == [0, 1]
    assert all("Iteration failed to converge after" in m for m in e.value.error_messages)
    # empty batch: nothing to do, no error
    empty = model.calculate_power_flow(update_data={"sym_load": np.zeros((0, 1), dtype=s0["sym_load"].dtype)},
                                       output_component_types=["node"])
    assert empty["node"].shape == (0, 2)


CASES = vc.load_cases()
GPU_METHODS = {"newton_raphson", "iterative_current", "linear", "linear_current"}
RUNS = [(n, s, m, b) for n, c in sorted(CASES.items()) for s, m, b in vc.case_runs(c) if m in GPU_METHODS]


@pytest.mark.parametrize("name,sym,method,is_batch", RUNS, ids=[f"{n}-{'sym' if s else 'asym'}-{m}-{'batch' if b else 'single'}" for n, s, m, b in RUNS])
def test_reference_validation_case_through_pgm_calculate(name, sym, method, is_batch):
    case = CASES[name]
    params = case["params"]
    model = pgm_core.PowerGridModel(vc.to_numpy(case["input"], "input"))
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


@pytest.mark.parametrize("sym", [True, False])
def test_benchmark_batch_equals_the_model_seam(sym):
    """BASELINE config 2 shape (and its asymmetric twin at 64 scenarios): PGM_calculate and pgmb_model_calculate are two doors
    to one engine, so the output bytes are identical"""
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    n_scn = 1000 if sym else 64
    update = grid.batch_update(n_scn, seed=0)
    a = pgm_core.PowerGridModel(grid.input_data).calculate_power_flow(symmetric=sym, update_data=update)
    b = pgm_b200.PowerGridModel(grid.input_data).calculate_power_flow(symmetric=sym, update_data=update)
    assert set(a) == set(b) and a["node"].shape == (n_scn, len(grid.input_data["node"]))
    for comp in a:
        assert a[comp].tobytes() == b[comp].tobytes(), comp
    assert np.all(a["node"]["energized"] == 1) and np.all(np.abs(a["node"]["u_pu"] - 1.0) < 0.2)


def _columns(arr):
    """structured array -> attribute -> contiguous column (the reference's columnar data format)"""
    return {name: np.ascontiguousarray(arr[name]) for name in arr.dtype.names}


def test_columnar_buffers_equal_row_buffers():
    """test_api_model.cpp:27-50 builds its datasets from a mix of row-based and columnar, dense and sparse buffers: columnar
    input (node, line), columnar dense batch update (line), sparse row update (source), columnar output -- same answers."""
    data = _api_model_input()
    model = pgm_core.PowerGridModel({"node": _columns(data["node"]), "line": _columns(data["line"]), "source": data["source"],
                                     "sym_load": data["sym_load"]})
    s0, s1 = _update(source_u_ref=0.5, q=100.0), _update(q=300.0, line_status=((0, 0), (0, 0)))
    line = np.stack([s0["line"], s1["line"]])
    batch = {"source": {"data": s0["source"], "indptr": np.array([0, 1, 1])},
             "sym_load": np.stack([s0["sym_load"], s1["sym_load"]]),
             "line": {k: np.ascontiguousarray(line[k]) for k in ("id", "from_status", "to_status")}}
    res = model.calculate_power_flow(update_data=batch, output_component_types={"node": ["id", "energized", "u", "u_pu", "u_angle"]})
    node = res["node"]
    assert node["id"].tolist() == [[0, 4], [0, 4]] and node["energized"].tolist() == [[1, 0], [1, 0]]
    assert node["u"][:, 0] == pytest.approx([40.0, 70.0]) and node["u_pu"][:, 0] == pytest.approx([0.4, 0.7])
    assert (node["u"][:, 1] == 0).all() and np.abs(node["u_angle"]).max() < 1e-12
    # sparse columnar update: only scenario 0 carries a source row
    batch["source"] = {"data": {"id": np.array([1], dtype=np.int32), "u_ref": np.array([0.5])}, "indptr": np.array([0, 1, 1])}
    res2 = model.calculate_power_flow(update_data=batch, output_component_types=["node"])["node"]
    assert res2["u"].tobytes() == node["u"].tobytes()


def test_model_update_optional_id_row_and_columnar():
    """test_api_model.cpp:549-650 ("Model update optional id"): a uniform update that covers every element may leave the ids
    out -- as NaN ids in a row buffer or by not supplying the id column at all"""
    node = initialize_array("input", "node", 1)
    node["id"], node["u_rated"] = 0, 100.0
    source = initialize_array("input", "source", 1)
    source["id"], source["node"], source["status"], source["u_ref"], source["sk"], source["rx_ratio"] = 1, 0, 1, 1.0, 1000.0, 0.0
    load = initialize_array("input", "sym_load", 1)
    load["id"], load["node"], load["status"], load["type"], load["p_specified"], load["q_specified"] = 2, 0, 1, 2, 0.0, 500.0
    inputs = {"node": node, "source": source, "sym_load": load}
    q = np.array([[100.0], [300.0]])
    row = initialize_array("update", "sym_load", (2, 1))
    row["q_specified"] = q  # ids stay NaN
    a = pgm_core.PowerGridModel(inputs).calculate_power_flow(update_data={"sym_load": row}, output_component_types=["node"])
    b = pgm_core.PowerGridModel({k: _columns(v) for k, v in inputs.items()}).calculate_power_flow(
        update_data={"sym_load": {"q_specified": q}}, output_component_types=["node"])
    row["id"] = 2
    c = pgm_core.PowerGridModel(inputs).calculate_power_flow(update_data={"sym_load": row}, output_component_types=["node"])
    assert a["node"]["u"][:, 0] == pytest.approx([90.0, 70.0])
    assert a["node"].tobytes() == b["node"].tobytes() == c["node"].tobytes()


def test_benchmark_batch_columnar_output():
    """columnar output of the benchmark batch (node voltages only -- the reference's advice for large batches,
    docs/user_manual/performance-guide.md): the same numbers as the row buffers hold"""
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    update = grid.batch_update(64, seed=0)
    model = pgm_core.PowerGridModel(grid.input_data)
    rows = model.calculate_power_flow(update_data=update, output_component_types=["node"])["node"]
    cols = model.calculate_power_flow(update_data={k: _columns(v) for k, v in update.items()},
                                      output_component_types={"node": ["u_pu", "u_angle"], "line": ["i_from"]})
    assert cols["node"]["u_pu"].tobytes() == rows["u_pu"].tobytes() and cols["node"]["u_angle"].tobytes() == rows["u_angle"].tobytes()
    assert cols["line"]["i_from"].shape == (64, len(grid.input_data["line"])) and (cols["line"]["i_from"] >= 0).all()


def test_api_model_multi_dimension():
    """tests/native_api_tests/test_api_model_multi_dimension.cpp: u_ref x p_specified x q_specified as a cartesian product of three
    columnar update datasets without ids; source current = |p + jq| / (sqrt3 u_rated u_ref) in the flattened (i, j, k) order"""
    node = initialize_array("input", "node", 1)
    node["id"], node["u_rated"] = 0, 10e3
    source = initialize_array("input", "source", 1)
    source["id"], source["node"], source["status"], source["u_ref"], source["sk"] = 1, 0, 1, 1.0, 1e20
    load = initialize_array("input", "sym_load", 1)
    load["id"], load["node"], load["status"], load["type"], load["p_specified"], load["q_specified"] = 2, 0, 1, 0, 0.0, 0.0
    model = pgm_core.PowerGridModel({"node": node, "source": source, "sym_load": load})
    u_ref, p, q = np.array([0.9, 1.0, 1.1]), np.array([1e6, 2e6, 3e6, 4e6]), np.array([0.1e6, 0.2e6, 0.3e6, 0.4e6, 0.5e6])
    dims = [{"source": {"u_ref": u_ref.reshape(-1, 1)}}, {"sym_load": {"p_specified": p.reshape(-1, 1)}},
            {"sym_load": {"q_specified": q.reshape(-1, 1)}}]
    res = model.calculate_power_flow(update_data=dims, output_component_types={"source": ["i"]})
    expected = np.abs(p[None, :, None] + 1j * q[None, None, :]) / (np.sqrt(3) * 10e3 * u_ref[:, None, None])
    assert res["source"]["i"].shape == (60, 1)
    assert res["source"]["i"][:, 0] == pytest.approx(expected.reshape(-1), rel=1e-6)
    # the model itself is untouched by the outer dimensions
    single = model.calculate_power_flow(output_component_types=["source"])
    assert single["source"]["i"][0] == pytest.approx(0.0, abs=1e-6)
    # failures keep their flattened scenario numbers: with max_iter 1 nothing converges to 1e-100 -- except where the start value
    # is the answer to the last bit (u_ref = 1.0 behind an sk = 1e20 source)
    with pytest.raises(pgm_core.PowerGridBatchError) as e:
        model.calculate_power_flow(update_data=dims, output_component_types={"source": ["i"]}, max_iterations=1, error_tolerance=1e-100)
    failed = e.value.failed_scenarios.tolist()
    assert failed == sorted(set(failed)) and set(range(20)) | set(range(40, 60)) <= set(failed) <= set(range(60))
    assert len(e.value.error_messages) == len(failed) and all("Iteration failed to converge" in m for m in e.value.error_messages)


def test_model_copies_calculate_concurrently_from_two_threads():
    """SURVEY 8b "Threading": one handle per thread, one model per thread (PGM_copy_model); ctypes releases the GIL, so the two
    PGM_calculate calls overlap on the device (each model drives its own CUDA streams).  Same bytes as the serial run."""
    import threading

    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    updates = [grid.batch_update(200, seed=s) for s in (0, 1)]
    base = pgm_core.PowerGridModel(grid.input_data)
    serial = [base.calculate_power_flow(update_data=u, output_component_types=["node", "line"]) for u in updates]
    models = [base.copy(), base.copy()]
    results, errors = [None, None], []

    def work(k):
        try:
            for _ in range(4):
                results[k] = models[k].calculate_power_flow(update_data=updates[k], output_component_types=["node", "line"])
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(2):
        for comp in ("node", "line"):
            assert results[k][comp].tobytes() == serial[k][comp].tobytes(), (k, comp)


def test_buffers_from_pgm_create_buffer_are_page_locked_and_filled_directly():
    """A client that allocates its output through PGM_create_buffer (the reference's C++ wrapper / benchmark do) gets page-locked
    memory on a GPU box; PGM_calculate fills it directly, with the same bytes as any other buffer"""
    import ctypes as C
    import time

    c, h = pgm_core.core(), pgm_core.Handle()
    c.PGM_b200_buffer_is_page_locked.argtypes = [C.c_void_p]
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    n_scn = 1000
    update = grid.batch_update(n_scn, seed=0)
    model = pgm_core.PowerGridModel(grid.input_data)
    expected = model.calculate_power_flow(update_data=update, output_component_types=["node", "line"])
    views, ptrs = {}, []
    for comp in ("node", "line"):
        meta = c.PGM_meta_get_component_by_name(h.h, b"sym_output", comp.encode())
        n = n_scn * len(grid.input_data[comp])
        ptr = c.PGM_create_buffer(h.h, meta, n)
        assert ptr and c.PGM_b200_buffer_is_page_locked(ptr) == 1
        ptrs.append(ptr)
        dt = pgm_b200.structs.SYM_OUTPUT[comp]
        views[comp] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * dt.itemsize,)).view(dt).reshape(n_scn, -1)
    small = c.PGM_create_buffer(h.h, c.PGM_meta_get_component_by_name(h.h, b"input", b"node"), 4)
    assert small and c.PGM_b200_buffer_is_page_locked(small) == 0
    c.PGM_destroy_buffer(small)
    out = pgm_core._Dataset(h, "sym_output", views, mutable=True, is_batch=True, batch_size=n_scn)
    upd = pgm_core._Dataset(h, "update", update, mutable=False, is_batch=True, batch_size=n_scn)
    opt = c.PGM_create_options(h.h)
    times = []
    for _ in range(4):
        t = time.perf_counter()
        c.PGM_calculate(h.h, model.ptr, opt, out.ptr, upd.ptr)
        times.append(time.perf_counter() - t)
        h.check()
    for comp in views:
        assert views[comp].tobytes() == expected[comp].tobytes(), comp
    print(f"PGM_calculate into PGM_create_buffer memory: {min(times) * 1e3:.2f} ms per {n_scn} scenarios (node + line)")
    c.PGM_destroy_options(opt)
    del out, views
    for ptr in ptrs:
        c.PGM_destroy_buffer(ptr)
        assert c.PGM_b200_buffer_is_page_locked(ptr) == 0
