"""JSON / msgpack (de)serialization behind the reference's C API names (serialization.h; csrc/capi_pgm_serialization.cpp):
the reference's own dataset files (kept verbatim in tests/golden/power_flow_cases.json) are read into buffers and compared
with an independent loader; datasets are written and read back in every format / layout; the reference's unchanged Python
wrapper runs its json_* / msgpack_* utilities on top of the library.  No GPU needed."""
import json
import os

import numpy as np
import pytest

import pgm_b200
import reference_wrapper
import validation_cases as vc
from pgm_b200 import pgm_core

CASES = dict(vc.load_cases())
# the automatic-tap-regulator cases bring transformer_tap_regulator datasets (input, update batches, outputs)
CASES.update(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tap_regulator_cases.json")))["cases"])


def _same(a, b):
    assert a.dtype == b.dtype and a.shape == b.shape
    for n in a.dtype.names:
        assert np.array_equal(a[n], b[n], equal_nan=True), n


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_dataset_files_are_read_like_the_independent_loader(name):
    case = CASES[name]
    kind, data = pgm_core.deserialize(json.dumps(case["input"]))
    expected = vc.to_numpy(case["input"], "input")
    assert kind == "input" and set(data) == {c for c in expected if len(expected[c])} | {c for c in data if c not in expected}
    for comp, arr in expected.items():
        if len(arr):
            _same(data[comp], arr)
    for key in ("sym_output", "asym_output"):
        if key in case:
            kind, out = pgm_core.deserialize(json.dumps(case[key]))
            assert kind == key
            for comp, arr in vc.to_numpy(case[key], key).items():
                _same(out[comp], arr)
    if "update_batch" in case:
        kind, upd = pgm_core.deserialize(json.dumps(case["update_batch"]))
        assert kind == "update"
        expected = vc.batch_update_arrays(vc.to_numpy(case["update_batch"], "update"))
        for comp, exp in expected.items():
            got = upd[comp]
            if isinstance(got, dict):
                assert np.array_equal(got["indptr"], exp["indptr"])
                _same(got["data"], exp["data"])
            else:  # every scenario holds the same number of elements: a uniform (n_scenarios, n) buffer
                _same(got.reshape(-1), exp["data"])


@pytest.mark.parametrize("fmt", [pgm_core.JSON, pgm_core.MSGPACK])
@pytest.mark.parametrize("compact", [False, True])
def test_round_trip_in_every_format(fmt, compact):
    grid = pgm_b200.FictionalGrid(seed=3, n_node_total_specified=60, n_connection_per_lv_feeder=3, n_lv_feeder=2, n_node_per_mv_feeder=3, n_mv_feeder=2)
    blob = pgm_core.serialize("input", grid.input_data, fmt, use_compact_list=compact)
    kind, back = pgm_core.deserialize(blob, fmt)
    assert kind == "input"
    for comp, arr in grid.input_data.items():
        _same(back[comp], arr)
    update = grid.batch_update(5, seed=1)  # three-phase values, NaN status columns
    update["asym_load"]["p_specified"][2, 1, 0] = np.inf
    update["asym_load"]["q_specified"][3, 0, 2] = -np.inf
    blob = pgm_core.serialize("update", update, fmt, use_compact_list=compact)
    kind, back = pgm_core.deserialize(blob, fmt)
    assert kind == "update"
    for comp, arr in update.items():
        _same(back[comp], arr)
    if fmt == pgm_core.JSON:
        doc = json.loads(blob)
        assert doc["version"] == "1.0" and doc["is_batch"] is True and len(doc["data"]) == 5
        assert ("asym_load" in doc["attributes"]) == compact
        assert '"inf"' in blob and '"-inf"' in blob


def test_sparse_batches_and_unknown_attributes():
    text = json.dumps({"version": "1.0", "type": "update", "is_batch": True, "attributes": {"sym_load": ["id", "p_specified"]},
                       "data": [{"sym_load": [[7, 1.5], [8, None]]}, {}, {"sym_load": [{"id": 9, "q_specified": "inf", "not_an_attribute": 1}],
                                                                         "line": [{"id": 3, "from_status": 0}]}]})
    kind, upd = pgm_core.deserialize(text)
    assert kind == "update" and upd["sym_load"]["indptr"].tolist() == [0, 2, 2, 3] and upd["line"]["indptr"].tolist() == [0, 0, 0, 1]
    sl = upd["sym_load"]["data"]
    assert sl["id"].tolist() == [7, 8, 9] and sl["p_specified"][0] == 1.5 and np.isnan(sl["p_specified"][1]) and np.isinf(sl["q_specified"][2])
    assert sl["status"].tolist() == [-128, -128, -128] and upd["line"]["data"]["to_status"][0] == -128
    again = pgm_core.deserialize(pgm_core.serialize("update", upd, use_compact_list=True))[1]
    assert again["sym_load"]["indptr"].tolist() == [0, 2, 2, 3]
    _same(again["sym_load"]["data"], sl)


@pytest.mark.parametrize("text,message", [
    ("{", "Insufficient bytes"), ("[1]", "Json root should be a map"), ('{"type": "input", "is_batch": false, "attributes": {}, "data": {}}', "Key version not found"),
    ('{"version": "1.0", "type": "input", "is_batch": true, "attributes": {}, "data": {}}', "does not match is_batch"),
    ('{"version": "1.0", "type": "input", "is_batch": false, "attributes": {}, "data": {"nodes": []}}', "Cannot find component"),
    ('{"version": "1.0", "type": "input", "is_batch": false, "attributes": {}, "data": {"node": [[1, 2.0]]}}', "predefined attributes"),
    ('{"version": "1.0", "type": "input", "is_batch": false, "attributes": {}, "data": {"node": [{"id": 1.5}]}}', "Expect an integer"),
    ('{"version": "1.0", "type": "input", "is_batch": false, "attributes": {}, "data": {"node": [{"id": 5000000000}]}}', "overflows"),
])
def test_malformed_documents_are_serialization_errors(text, message):
    with pytest.raises(pgm_core.PowerGridError, match=message):
        pgm_core.deserialize(text)


pgm = reference_wrapper.load()


@pytest.mark.skipif(pgm is None, reason="reference Python wrapper not available (tools/install_reference_wrapper.py)")
def test_the_reference_wrapper_serializes_through_this_library():
    from power_grid_model.utils import json_deserialize, json_serialize, msgpack_deserialize, msgpack_serialize

    case = CASES["dummy-test"]
    data = json_deserialize(json.dumps(case["input"]))
    expected = vc.to_numpy(case["input"], "input")
    for comp, arr in expected.items():
        _same(np.asarray(data[comp]), arr)
    for compact in (False, True):
        back = json_deserialize(json_serialize(data, use_compact_list=compact))
        blob = msgpack_deserialize(msgpack_serialize(data, use_compact_list=compact))
        for comp in data:
            _same(np.asarray(back[comp]), np.asarray(data[comp]))
            _same(np.asarray(blob[comp]), np.asarray(data[comp]))
    batch = json_deserialize(json.dumps(CASES["dummy-test-batch"]["update_batch"]))
    assert "sym_load" in {str(k) for k in batch}
