"""Loader + comparison for the reference's power-flow validation cases stored in tests/golden/power_flow_cases.json.
The comparison restates tests/unit/utils.py:259-330 of the reference (u_angle relative to the first entry and compared as
a phasor; per-attribute absolute tolerances given as regex patterns)."""
import json
import os
import re

import numpy as np

from pgm_b200 import structs

_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "power_flow_cases.json")
_cache = {}


def load_cases():
    if "cases" not in _cache:
        _cache["cases"] = json.load(open(_FIXTURE))["cases"]
    return _cache["cases"]


def _num(v):
    if isinstance(v, str):
        return {"inf": np.inf, "+inf": np.inf, "-inf": -np.inf, "nan": np.nan}[v.lower()]
    return v


def _fill(arr, rows, attributes):
    for i, row in enumerate(rows):
        if isinstance(row, dict):
            items = row.items()
        else:
            items = zip(attributes, row)
        for k, v in items:
            if k not in arr.dtype.names or v is None:
                continue
            if isinstance(v, list):
                arr[k][i] = [np.nan if x is None else _num(x) for x in v]
            else:
                arr[k][i] = _num(v)


def to_numpy(dataset, kind):
    """kind: input | update | sym_output | asym_output.  Returns dict comp->array (single) or a list of dicts (batch)."""
    attrs = dataset.get("attributes", {})

    def one(d):
        out = {}
        for comp, rows in d.items():
            table = {"input": structs.INPUT, "update": structs.UPDATE, "sym_output": structs.SYM_OUTPUT,
                     "asym_output": structs.ASYM_OUTPUT}[kind]
            if comp not in table:
                continue
            arr = structs.initialize_array(kind, comp, len(rows))
            _fill(arr, rows, attrs.get(comp, []))
            out[comp] = arr
        return out

    if dataset["is_batch"]:
        return [one(d) for d in dataset["data"]]
    return one(dataset["data"])


def batch_update_arrays(update_list):
    """list of per-scenario dicts -> dict comp -> {'data', 'indptr'} (sparse batch buffers)"""
    comps = sorted({c for s in update_list for c in s})
    out = {}
    for c in comps:
        chunks = [s.get(c, np.zeros(0, dtype=structs.UPDATE[c])) for s in update_list]
        indptr = np.cumsum([0] + [len(x) for x in chunks])
        out[c] = {"data": np.concatenate(chunks) if chunks else np.zeros(0, structs.UPDATE[c]), "indptr": indptr}
    return out


def compare_result(actual, expected, rtol, atol):
    """actual / expected: dict comp -> structured array (1-D over components)."""
    for key, expected_data in expected.items():
        for col_name in expected_data.dtype.names:
            actual_col = actual[key][col_name]
            expected_col = expected_data[col_name]
            if expected_col.dtype == np.float64:
                expect_all_nan = np.all(np.isnan(expected_col))
            elif expected_col.dtype == np.int8:
                expect_all_nan = np.all(expected_col == np.iinfo("i1").min)
            else:
                expect_all_nan = np.all(expected_col == np.iinfo("i4").min)
            if expect_all_nan:
                continue
            if col_name == "u_angle":
                actual_col = actual_col - actual_col.ravel()[0]
                expected_col = expected_col - expected_col.ravel()[0]
            if col_name.endswith("_angle"):
                mag = col_name[: -len("_angle")]
                if np.all(np.isnan(expected_data[mag])):
                    continue
                actual_col = actual[key][mag] * np.exp(1j * actual_col)
                expected_col = expected_data[mag] * np.exp(1j * expected_col)
            a = atol
            if isinstance(atol, dict):
                a = atol["default"]
                for pattern, tol in atol.items():
                    if re.fullmatch(pattern, col_name):
                        a = tol
                        break
            ok = np.isclose(actual_col, expected_col, rtol=rtol, atol=a) | np.isnan(expected_col)
            assert np.all(ok), (
                f"Not all values match for {key}.{col_name} (rtol={rtol}, atol={a})\nActual:   {actual_col}\n"
                f"Expected: {expected_col}\nDiff:     {actual_col - expected_col}"
            )


def case_runs(case):
    """yield (sym, method, is_batch) combinations the case defines golden output for"""
    methods = case["params"]["calculation_method"]
    if isinstance(methods, str):
        methods = [methods]
    for m in methods:
        if m == "iterative_linear":
            continue
        for sym in (True, False):
            p = "sym_output" if sym else "asym_output"
            if p in case:
                yield sym, m, False
            if p + "_batch" in case and "update_batch" in case:
                yield sym, m, True
