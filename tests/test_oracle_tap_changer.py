"""The oracle's automatic tap changer (oracle/tap_optimizer.hpp, oracle/model.hpp) pinned on the reference's own validation
cases (tests/data/power_flow/automatic-tap-regulator/* in tests/golden/tap_regulator_cases.json): tap positions, voltages and
flows; the cases the reference expects to end in MaxIterationReached must fail here too."""
import json
import os

import numpy as np
import pytest

import oracle_lib as orc
import validation_cases as vc

CASES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tap_regulator_cases.json")))["cases"]
META = ("n_iter", "status", "error", "n_failed")
RUNS = [(n, s, m, b) for n, c in sorted(CASES.items()) for s, m, b in vc.case_runs(c)]


@pytest.mark.parametrize("name,sym,method,batch", RUNS)
def test_oracle_tap_regulator_validation_case(name, sym, method, batch):
    case = CASES[name]
    params = case["params"]
    model = orc.Model(vc.to_numpy(case["input"], "input"))
    kw = dict(sym=sym, method=method, tap_changing_strategy=params["tap_changing_strategy"])
    kind = "sym_output" if sym else "asym_output"
    if batch:
        kw["update"] = vc.batch_update_arrays(vc.to_numpy(case["update_batch"], "update"))
    res = model.calculate(**kw)
    if "raises" in params:
        assert (res["status"] != 0).all() and "Maximum number of iterations reached" in res["error"]
        plain = model.calculate(sym=sym, method=method, update=kw.get("update"))  # the model is as it was
        assert (plain["status"] == 0).all() and (plain["transformer_tap_regulator"]["tap_pos"] == -128).all()
        return
    assert (res["status"] == 0).all(), res["error"]
    if batch:
        for s, exp in enumerate(vc.to_numpy(case[kind + "_batch"], kind)):
            vc.compare_result({c: res[c][s] for c in exp}, exp, params["rtol"], params["atol"])
    else:
        vc.compare_result({c: res[c][0] for c in res if c not in META}, vc.to_numpy(case[kind], kind),
                          params["rtol"], params["atol"])
    again = model.calculate(**kw)
    for c in res:
        if c not in META:
            for f in res[c].dtype.names:
                assert np.array_equal(res[c][f], again[c][f], equal_nan=res[c][f].dtype.kind == "f"), (c, f)
