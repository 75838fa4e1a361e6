"""CPU tests of the reference-named C API (include/pgm_b200_capi.h, csrc/capi_pgm.cpp): every declared symbol is exported, the
handle / options / dataset calls behave like the reference's (error codes and messages of power_grid_model_c/src/handle.cpp,
dataset.cpp, auxiliary/dataset.hpp:233-243, 587-625; tests/native_api_tests/test_api_model.cpp "Test get indexer", "Input error
handling").  No compute call succeeds without a GPU: PGM_calculate must report the missing device, never fall back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pgm_b200
from pgm_b200 import pgm_core
from pgm_b200.structs import initialize_array

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_pgm_symbol():
    header = open(os.path.join(ROOT, "include", "pgm_b200_capi.h")).read()
    names = re.findall(r"PGM_API\s+[\w\s\*]+?\b(PGM_\w+)\s*\(", header)
    assert len(names) == 36, len(names)
    lib = C.CDLL(os.path.join(ROOT, "power-grid-model_b200", "libpgm_b200.so"))
    for n in names:
        assert hasattr(lib, n), n
    assert b"pgm_b200" in pgm_core.core().PGM_version()


def _api_model_input(load_id=2):
    """the grid of tests/native_api_tests/test_api_model.cpp:27-52"""
    node = initialize_array("input", "node", 2)
    node["id"], node["u_rated"] = [0, 4], 100.0
    line = initialize_array("input", "line", 2)
    line["id"], line["from_node"], line["to_node"] = [5, 6], [0, 4], [4, 0]
    line["from_status"], line["to_status"] = [0, 0], [1, 0]
    line["r1"], line["x1"], line["c1"], line["tan1"] = 1.0, 1.0, 0.0, 0.0  # never connected; the reference leaves them NaN
    source = initialize_array("input", "source", 1)
    source["id"], source["node"], source["status"], source["u_ref"], source["sk"], source["rx_ratio"] = 1, 0, 1, 1.0, 1000.0, 0.0
    load = initialize_array("input", "sym_load", 1)
    load["id"], load["node"], load["status"], load["type"], load["p_specified"], load["q_specified"] = load_id, 0, 1, 2, 0.0, 500.0
    return {"node": node, "line": line, "source": source, "sym_load": load}


def test_handle_is_cleared_by_every_call_and_reports_dataset_errors():
    c = pgm_core.core()
    h = pgm_core.Handle()
    assert c.PGM_error_code(h.h) == 0 and c.PGM_error_message(h.h) == b"" and c.PGM_n_failed_scenarios(h.h) == 0
    assert c.PGM_create_dataset_const(h.h, b"input", 0, 2) is None
    assert c.PGM_error_code(h.h) == pgm_core.PGM_REGULAR_ERROR
    assert c.PGM_error_message(h.h) == b"Dataset error: For non-batch dataset, batch size should be one!\n"
    assert c.PGM_create_dataset_const(h.h, b"update", 1, -1) is None
    assert b"Batch size cannot be negative" in c.PGM_error_message(h.h)
    assert c.PGM_create_dataset_const(h.h, b"no_such_dataset", 0, 1) is None
    assert c.PGM_error_code(h.h) == pgm_core.PGM_REGULAR_ERROR
    ds = c.PGM_create_dataset_const(h.h, b"update", 1, 3)
    assert ds and c.PGM_error_code(h.h) == 0  # the successful call cleared the error
    buf = initialize_array("update", "sym_load", 6)
    indptr = np.array([0, 1, 3, 6], dtype=np.int64)
    add = c.PGM_dataset_const_add_buffer
    add(h.h, ds, b"sym_load", 2, 5, None, buf.ctypes.data)
    assert b"total_elements should be equal to elements_per_scenario * batch_size" in c.PGM_error_message(h.h)
    add(h.h, ds, b"sym_load", -1, 6, None, buf.ctypes.data)
    assert b"indptr should be supplied" in c.PGM_error_message(h.h)
    add(h.h, ds, b"sym_load", 2, 6, indptr.ctypes.data, buf.ctypes.data)
    assert b"For a uniform buffer, indptr should be nullptr" in c.PGM_error_message(h.h)
    add(h.h, ds, b"sym_load", -1, 5, indptr.ctypes.data, buf.ctypes.data)
    assert b"begin with 0 and end with total_elements" in c.PGM_error_message(h.h)
    bad = np.array([0, 4, 3, 6], dtype=np.int64)
    add(h.h, ds, b"sym_load", -1, 6, bad.ctypes.data, buf.ctypes.data)
    assert b"non-decreasing" in c.PGM_error_message(h.h)
    add(h.h, ds, b"no_such_component", 2, 6, None, buf.ctypes.data)
    assert c.PGM_error_code(h.h) == pgm_core.PGM_REGULAR_ERROR
    add(h.h, ds, b"sym_load", -1, 6, indptr.ctypes.data, buf.ctypes.data)
    assert c.PGM_error_code(h.h) == 0
    add(h.h, ds, b"sym_load", 2, 6, None, buf.ctypes.data)
    assert b"Cannot have duplicated components" in c.PGM_error_message(h.h)
    c.PGM_dataset_const_add_attribute_buffer(h.h, ds, b"sym_load", b"id", buf.ctypes.data)
    assert c.PGM_error_code(h.h) == pgm_core.PGM_REGULAR_ERROR and b"row-based" in c.PGM_error_message(h.h)
    c.PGM_clear_error(h.h)
    assert c.PGM_error_code(h.h) == 0 and c.PGM_error_message(h.h) == b""
    c.PGM_destroy_dataset_const(ds)
    # a null handle is tolerated everywhere, like in the reference
    assert c.PGM_error_code(None) == 0 and c.PGM_error_message(None) is None
    assert c.PGM_create_dataset_const(None, b"input", 0, 2) is None


def test_get_indexer_and_construction_errors():
    """test_api_model.cpp:255-287 ("Test get indexer"), :319-343 ("Construction error"), :345-372 ("Update error")"""
    node = initialize_array("input", "node", 3)
    node["id"], node["u_rated"] = [1, 2, 3], 10e3
    model = pgm_core.PowerGridModel({"node": node})
    assert model.get_indexer("node", [2, 1, 3, 2]).tolist() == [1, 0, 2, 1]
    with pytest.raises(pgm_core.PowerGridError, match="The id cannot be found: 4"):
        model.get_indexer("node", [2, 1, 3, 4])
    with pytest.raises(pgm_core.PowerGridError, match="Wrong type for object with id 2"):
        model.get_indexer("sym_load", [2, 1, 3, 2])
    with pytest.raises(pgm_core.PowerGridError, match="Conflicting id detected:"):
        pgm_core.PowerGridModel(_api_model_input(load_id=0))
    model = pgm_core.PowerGridModel(_api_model_input())
    assert model.get_indexer("sym_load", [2]).tolist() == [0] and model.get_indexer("line", [6, 5]).tolist() == [1, 0]
    upd = initialize_array("update", "source", 1)
    upd["id"], upd["u_ref"] = 99, 0.5
    with pytest.raises(pgm_core.PowerGridError, match="The id cannot be found:"):
        model.update(update_data={"source": upd})
    copy = model.copy()
    assert copy.get_indexer("node", [4, 0]).tolist() == [1, 0]


def test_unsupported_components_and_options_are_refused_loudly():
    data = _api_model_input()
    link = np.zeros(1, dtype=np.dtype([("id", "<i4"), ("from_node", "<i4"), ("to_node", "<i4"), ("from_status", "i1"), ("to_status", "i1")]))
    with pytest.raises(pgm_core.PowerGridError, match="'link' is not built by pgm_b200"):
        pgm_core.PowerGridModel({**data, "link": link})
    # sensors / faults may be present: power flow ignores them
    pgm_core.PowerGridModel({**data, "sym_voltage_sensor": np.zeros(1, dtype=np.dtype([("id", "<i4"), ("pad", "V28")]))})
    model = pgm_core.PowerGridModel(data)
    with pytest.raises(pgm_core.PowerGridError, match="CalculationType is not implemented for"):
        model.calculate_power_flow(calculation_type=-128)
    with pytest.raises(pgm_core.PowerGridError, match="get_optimizer_type is not implemented for"):
        model.calculate_power_flow(tap_changing_strategy=-128)
    with pytest.raises(pgm_core.PowerGridError, match="power_flow only"):
        model.calculate_power_flow(calculation_type=1)
    with pytest.raises(pgm_core.PowerGridError, match="calculation method is invalid"):
        model.calculate_power_flow(calculation_method="iterative_linear")
    c, h = pgm_core.core(), model.handle
    opt = c.PGM_create_options(h.h)
    out = np.zeros(2, dtype=pgm_b200.structs.SYM_OUTPUT["node"])
    ds = pgm_core._Dataset(h, "asym_output", {}, mutable=True, is_batch=False, batch_size=1)
    c.PGM_calculate(h.h, model.ptr, opt, ds.ptr, None)
    assert b"does not match the calculation symmetry" in c.PGM_error_message(h.h)
    ds = pgm_core._Dataset(h, "sym_output", {"node": out[:1]}, mutable=True, is_batch=False, batch_size=1)
    c.PGM_calculate(h.h, model.ptr, opt, ds.ptr, None)
    assert b"must hold exactly the model's 2 elements" in c.PGM_error_message(h.h)
    upd = pgm_core._Dataset(h, "update", {}, mutable=False, is_batch=False, batch_size=1)
    ds = pgm_core._Dataset(h, "sym_output", {"node": out}, mutable=True, is_batch=False, batch_size=1)
    c.PGM_calculate(h.h, model.ptr, opt, ds.ptr, upd.ptr)
    assert b"Both batch_dataset and output_dataset should be a batch" in c.PGM_error_message(h.h)
    c.PGM_destroy_options(opt)


def test_pgm_calculate_has_no_cpu_fallback():
    if pgm_b200.lib().pgmb_device_count() > 0:
        pytest.skip("GPU present")
    model = pgm_core.PowerGridModel(_api_model_input())
    with pytest.raises(pgm_core.PowerGridError, match="no CPU fallback"):
        model.calculate_power_flow()
