"""The reference's UNCHANGED Python wrapper on top of libpgm_b200.so (CPU part): it imports (every PGM_* symbol it binds
resolves), builds its numpy dtypes from our PGM_meta_* tables, creates a model, indexes it, maps our error messages to its own
exception classes -- and a calculation without a GPU fails loudly instead of falling back."""
import numpy as np
import pytest

import pgm_b200
import reference_wrapper

pgm = reference_wrapper.load()
pytestmark = pytest.mark.skipif(pgm is None, reason="reference Python wrapper not available (tools/install_reference_wrapper.py)")


def _grid():
    node = pgm.initialize_array("input", "node", 2)
    node["id"], node["u_rated"] = [0, 4], 100.0
    source = pgm.initialize_array("input", "source", 1)
    source["id"], source["node"], source["status"], source["u_ref"], source["sk"], source["rx_ratio"] = 1, 0, 1, 1.0, 1000.0, 0.0
    load = pgm.initialize_array("input", "sym_load", 1)
    load["id"], load["node"], load["status"], load["type"], load["p_specified"], load["q_specified"] = 2, 0, 1, 2, 0.0, 500.0
    return {"node": node, "source": source, "sym_load": load}


def test_wrapper_dtypes_come_from_our_meta_tables():
    st = pgm_b200.structs
    for ds, table in (("input", st.INPUT), ("update", st.UPDATE), ("sym_output", st.SYM_OUTPUT), ("asym_output", st.ASYM_OUTPUT)):
        for comp, dt in table.items():
            theirs = pgm.initialize_array(ds, comp, 1).dtype
            assert theirs == dt, (ds, comp)
    arr = pgm.initialize_array("input", "asym_load", 2)
    assert (arr["id"] == np.iinfo(np.int32).min).all() and np.isnan(arr["p_specified"]).all() and arr["p_specified"].shape == (2, 3)


def test_wrapper_creates_indexes_copies_and_reports_errors():
    model = pgm.PowerGridModel(_grid())
    assert {str(k): v for k, v in model.all_component_count.items()} == {"node": 2, "source": 1, "sym_load": 1}
    assert model.get_indexer("node", np.array([4, 0, 4])).tolist() == [1, 0, 1]
    import copy

    assert copy.copy(model).get_indexer("sym_load", np.array([2])).tolist() == [0]
    bad = _grid()
    bad["sym_load"]["id"] = 0
    from power_grid_model.errors import ConflictID, IDNotFound

    with pytest.raises(ConflictID):
        pgm.PowerGridModel(bad)
    upd = pgm.initialize_array("update", "source", 1)
    upd["id"], upd["u_ref"] = 99, 0.5
    with pytest.raises(IDNotFound):
        model.update(update_data={"source": upd})


def test_wrapper_calculation_needs_the_gpu():
    if pgm_b200.lib().pgmb_device_count() > 0:
        pytest.skip("GPU present")
    from power_grid_model.errors import PowerGridError

    with pytest.raises(PowerGridError, match="no CPU fallback"):
        pgm.PowerGridModel(_grid()).calculate_power_flow()
