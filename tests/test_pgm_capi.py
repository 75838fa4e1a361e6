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
    names.remove("PGM_b200_buffer_is_page_locked")  # the one extension
    assert len(names) == 83, len(names)  # the reference's whole function surface (83 functions)
    lib = C.CDLL(os.path.join(ROOT, "power-grid-model_b200", "libpgm_b200.so"))
    for n in names:
        assert hasattr(lib, n), n
    assert b"pgm_b200" in pgm_core.core().PGM_version()


def test_library_exports_the_dataset_definition_constants():
    """dataset_definitions.h: 838 PGM_def_* pointer constants in the reference; every one declared in
    include/pgm_b200_dataset_definitions.h is exported and points at the meta data PGM_meta_get_*_by_name returns"""
    header = open(os.path.join(ROOT, "include", "pgm_b200_dataset_definitions.h")).read()
    names = re.findall(r"\b(PGM_def_\w+);", header)
    assert len(names) >= 838 and len(set(names)) == len(names)
    lib = C.CDLL(os.path.join(ROOT, "power-grid-model_b200", "libpgm_b200.so"))
    for n in names:
        assert hasattr(lib, n), n
    c, h = pgm_core.core(), pgm_core.Handle()
    value = lambda n: C.c_void_p.in_dll(lib, n).value  # noqa: E731
    assert value("PGM_def_input") == c.PGM_meta_get_dataset_by_name(h.h, b"input")
    assert value("PGM_def_sym_output_line") == c.PGM_meta_get_component_by_name(h.h, b"sym_output", b"line")
    assert value("PGM_def_update_asym_load_q_specified") == c.PGM_meta_get_attribute_by_name(h.h, b"update", b"asym_load", b"q_specified")


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
    assert c.PGM_error_message(h.h) == b"Dataset error: Cannot add attribute buffers to row-based dataset!\n"
    # columnar component: null row pointer, one buffer per attribute (auxiliary/dataset.hpp:627-646)
    ids = np.zeros(6, dtype=np.int32)
    add(h.h, ds, b"asym_load", 2, 6, None, None)
    assert c.PGM_error_code(h.h) == 0
    add_attr = c.PGM_dataset_const_add_attribute_buffer
    add_attr(h.h, ds, b"asym_load", b"id", ids.ctypes.data)
    assert c.PGM_error_code(h.h) == 0
    add_attr(h.h, ds, b"asym_load", b"id", ids.ctypes.data)
    assert b"Cannot have duplicated attribute buffers" in c.PGM_error_message(h.h)
    add_attr(h.h, ds, b"asym_load", b"no_such_attribute", ids.ctypes.data)
    assert b"Cannot find attribute with name: no_such_attribute" in c.PGM_error_message(h.h)
    add_attr(h.h, ds, b"asym_load", b"p_specified", None)
    assert b"Attribute buffer data pointer cannot be null for non-empty component" in c.PGM_error_message(h.h)
    add_attr(h.h, ds, b"shunt", b"id", ids.ctypes.data)
    assert b"Cannot find component 'shunt'" in c.PGM_error_message(h.h)
    # dataset info (dataset.h:27-138)
    info = c.PGM_dataset_const_get_info(h.h, ds)
    assert c.PGM_dataset_info_name(h.h, info) == b"update" and c.PGM_dataset_info_is_batch(h.h, info) == 1
    assert c.PGM_dataset_info_batch_size(h.h, info) == 3 and c.PGM_dataset_info_n_components(h.h, info) == 2
    assert [c.PGM_dataset_info_component_name(h.h, info, i) for i in range(2)] == [b"sym_load", b"asym_load"]
    assert [c.PGM_dataset_info_elements_per_scenario(h.h, info, i) for i in range(2)] == [-1, 2]
    assert [c.PGM_dataset_info_total_elements(h.h, info, i) for i in range(2)] == [6, 6]
    assert [c.PGM_dataset_info_has_attribute_indications(h.h, info, i) for i in range(2)] == [0, 1]
    assert [c.PGM_dataset_info_n_attribute_indications(h.h, info, i) for i in range(2)] == [0, 1]
    assert c.PGM_dataset_info_attribute_name(h.h, info, 1, 0) == b"id"
    assert c.PGM_dataset_info_component_name(h.h, info, 2) is None and b"Index out of range" in c.PGM_error_message(h.h)
    # cartesian product linked list: no self reference, no cycle (test_api_model_multi_dimension.cpp:105-110)
    ds2 = c.PGM_create_dataset_const(h.h, b"update", 1, 2)
    link = c.PGM_dataset_const_set_next_cartesian_product_dimension
    link(h.h, ds, ds)
    assert b"Cannot create cyclic cartesian product dimension linked list" in c.PGM_error_message(h.h)
    link(h.h, ds, ds2)
    assert c.PGM_error_code(h.h) == 0
    link(h.h, ds2, ds)
    assert b"Cannot create cyclic cartesian product dimension linked list" in c.PGM_error_message(h.h)
    link(h.h, ds, None)
    assert c.PGM_error_code(h.h) == 0
    c.PGM_destroy_dataset_const(ds2)
    c.PGM_clear_error(h.h)
    assert c.PGM_error_code(h.h) == 0 and c.PGM_error_message(h.h) == b""
    c.PGM_destroy_dataset_const(ds)
    # a null handle is tolerated everywhere, like in the reference
    assert c.PGM_error_code(None) == 0 and c.PGM_error_message(None) is None
    assert c.PGM_create_dataset_const(None, b"input", 0, 2) is None


def test_get_indexer_and_construction_errors():
    """test_api_model.cpp:255-287 ("Test get indexer"), :319-343 ("Construction error"), :345-372 ("Update error")"""
    # the reference's test builds this model from columnar buffers (test_api_model.cpp:259-265)
    model = pgm_core.PowerGridModel({"node": {"id": np.array([1, 2, 3], dtype=np.int32), "u_rated": np.full(3, 10e3)}})
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
    regulator = np.zeros(1, dtype=np.dtype([("id", "<i4"), ("regulated_object", "<i4"), ("status", "i1"), ("control_side", "i1"),
                                            ("u_set", "<f8"), ("u_band", "<f8"), ("line_drop_compensation_r", "<f8"),
                                            ("line_drop_compensation_x", "<f8")], align=True))
    # a tap regulator must regulate a transformer (main_core/input.hpp:168-214): object 0 is a node, 12345 does not exist
    regulator["id"] = 777
    with pytest.raises(pgm_core.PowerGridError, match="transformer_tap_regulator regulator is not supported for object with ID 0"):
        pgm_core.PowerGridModel({**data, "transformer_tap_regulator": regulator})
    regulator["regulated_object"] = 12345
    with pytest.raises(pgm_core.PowerGridError, match="The id cannot be found: 12345"):
        pgm_core.PowerGridModel({**data, "transformer_tap_regulator": regulator})
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


def test_meta_data_tables_match_the_struct_layouts():
    """PGM_meta_* (meta_data.h): the tables generated from the reference's definition files against the independently written
    numpy dtypes of pgm_b200.structs (which the kernels' structs are static_asserted against) for every component the engine
    builds, plus the shape of the whole table (datasets / components in the reference's order)."""
    meta = pgm_core.power_grid_meta_data()
    assert list(meta) == ["input", "update", "sym_output", "asym_output", "sc_output"]
    comps = list(meta["input"])
    assert len(comps) == 22 and comps[:6] == ["node", "line", "asym_line", "link", "generic_branch", "transformer"]
    assert comps[-3:] == ["fault", "transformer_tap_regulator", "voltage_regulator"]
    assert all(list(meta[d]) == comps for d in meta)
    st = pgm_b200.structs
    n_checked = 0
    for ds, table in (("input", st.INPUT), ("update", st.UPDATE), ("sym_output", st.SYM_OUTPUT), ("asym_output", st.ASYM_OUTPUT)):
        for comp, dt in table.items():
            got = meta[ds][comp]
            assert got.names == dt.names and got.itemsize == dt.itemsize, (ds, comp)
            for name in dt.names:
                assert got.fields[name][1] == dt.fields[name][1] and got.fields[name][0] == dt.fields[name][0], (ds, comp, name)
            n_checked += 1
    assert n_checked == 15 + 14 + 15 + 15
    # sizes the reference states for components outside the engine (auxiliary/static_asserts, SURVEY appendix B conventions)
    assert meta["input"]["link"].itemsize == 16 and meta["update"]["node"].itemsize == 4
    assert meta["input"]["three_winding_transformer"].names[:4] == ("id", "node_1", "node_2", "node_3")
    assert meta["sym_output"]["three_winding_transformer"].names[-4:] == ("p_3", "q_3", "i_3", "s_3")
    assert meta["sc_output"]["node"].fields["u_pu"][0].shape == (3,)
    c, h = pgm_core.core(), pgm_core.Handle()
    assert c.PGM_is_little_endian(h.h) == 1
    assert c.PGM_meta_get_dataset_by_name(h.h, b"nope") is None and b"Cannot find dataset with name: nope!" in c.PGM_error_message(h.h)
    assert c.PGM_meta_get_component_by_name(h.h, b"input", b"nope") is None
    assert b"Cannot find component with name: nope!" in c.PGM_error_message(h.h)
    assert c.PGM_meta_get_attribute_by_name(h.h, b"input", b"node", b"nope") is None
    assert b"Cannot find attribute with name: nope!" in c.PGM_error_message(h.h)
    assert c.PGM_meta_get_dataset_by_idx(h.h, 5) is None and b"Index out of range!" in c.PGM_error_message(h.h)
    node = c.PGM_meta_get_component_by_name(h.h, b"input", b"node")
    assert c.PGM_meta_component_alignment(h.h, node) == 8 and c.PGM_meta_component_size(h.h, node) == 16
    assert c.PGM_meta_get_attribute_by_idx(h.h, node, 2) is None and b"Index out of range!" in c.PGM_error_message(h.h)


def test_buffer_create_set_nan_set_get_value():
    """buffer.h: the calls test_api_buffer.cpp exercises -- NaN fill, strided set / get of one attribute"""
    c, h = pgm_core.core(), pgm_core.Handle()
    comp = c.PGM_meta_get_component_by_name(h.h, b"input", b"asym_load")
    dt = pgm_core.power_grid_meta_data()["input"]["asym_load"]
    ptr = c.PGM_create_buffer(h.h, comp, 4)
    assert ptr and ptr % 8 == 0
    view = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(4 * dt.itemsize,)).view(dt)
    view.view(np.uint8)[:] = 0
    c.PGM_buffer_set_nan(h.h, comp, ptr, 1, 2)
    assert view["id"].tolist() == [0, -2**31, -2**31, 0] and view["status"].tolist() == [0, -128, -128, 0]
    assert np.isnan(view["p_specified"][1:3]).all() and (view["p_specified"][[0, 3]] == 0).all()
    attr_id = c.PGM_meta_get_attribute_by_name(h.h, b"input", b"asym_load", b"id")
    attr_q = c.PGM_meta_get_attribute_by_name(h.h, b"input", b"asym_load", b"q_specified")
    ids = np.array([10, 11, 12, 13], dtype=np.int32)
    c.PGM_buffer_set_value(h.h, attr_id, ptr, ids.ctypes.data, 0, 4, -1)
    q = np.arange(24, dtype=np.float64).reshape(4, 6)  # rows of 6 doubles, the attribute takes the first 3: stride 48
    c.PGM_buffer_set_value(h.h, attr_q, ptr, q.ctypes.data, 1, 2, 48)
    assert view["id"].tolist() == [10, 11, 12, 13]
    assert view["q_specified"][1].tolist() == [6.0, 7.0, 8.0] and view["q_specified"][2].tolist() == [12.0, 13.0, 14.0]
    assert (view["q_specified"][0] == 0).all()
    back = np.full((4, 3), -1.0)
    c.PGM_buffer_get_value(h.h, attr_q, ptr, back.ctypes.data, 1, 2, -1)
    assert back[1].tolist() == [6.0, 7.0, 8.0] and back[2].tolist() == [12.0, 13.0, 14.0] and (back[[0, 3]] == -1).all()
    c.PGM_buffer_set_value(h.h, attr_q, ptr, None, 0, 1, -1)
    assert c.PGM_error_code(h.h) == pgm_core.PGM_REGULAR_ERROR
    c.PGM_destroy_buffer(ptr)


def test_serialization_errors_use_their_own_error_code():
    """(de)serializer failures are PGM_serialization_error (handle.hpp:70-89), not PGM_regular_error"""
    c, h = pgm_core.core(), pgm_core.Handle()
    assert c.PGM_create_deserializer_from_null_terminated_string(h.h, b"{}", 0) is None
    assert c.PGM_error_code(h.h) == pgm_core.PGM_SERIALIZATION_ERROR and b"Key version not found" in c.PGM_error_message(h.h)
    assert c.PGM_create_serializer(h.h, None, 0) is None and c.PGM_error_code(h.h) == pgm_core.PGM_SERIALIZATION_ERROR
    assert c.PGM_create_deserializer_from_null_terminated_string(h.h, b"{}", 1) is None  # msgpack needs the binary entry point
    assert c.PGM_error_code(h.h) == pgm_core.PGM_SERIALIZATION_ERROR
    c.PGM_destroy_serializer(None)
    c.PGM_destroy_deserializer(None)
