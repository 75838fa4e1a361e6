"""Pins the oracle end to end (components -> topology -> Y-bus -> solver -> output structs, incl. batch update /
restore) against the reference's golden validation outputs (tests/data/power_flow/**, committed as
tests/golden/power_flow_cases.json by tests/golden/make_validation_fixtures.py)."""
import numpy as np
import pytest

import oracle_lib as orc
import validation_cases as vc

CASES = vc.load_cases()
RUNS = [(name, sym, m, b) for name, case in sorted(CASES.items()) for sym, m, b in vc.case_runs(case)]


@pytest.mark.parametrize("name,sym,method,is_batch", RUNS, ids=[f"{n}-{'sym' if s else 'asym'}-{m}-{'batch' if b else 'single'}" for n, s, m, b in RUNS])
def test_validation_case(name, sym, method, is_batch):
    case = CASES[name]
    params = case["params"]
    model = orc.Model(vc.to_numpy(case["input"], "input"))
    kind = "sym_output" if sym else "asym_output"
    if not is_batch:
        res = model.calculate(sym=sym, method=method, err_tol=1e-8, max_iter=20)
        assert res["n_failed"] == 0, res["error"]
        expected = vc.to_numpy(case[kind], kind)
        vc.compare_result({k: v[0] for k, v in res.items() if isinstance(v, np.ndarray) and v.ndim == 2}, expected,
                          params["rtol"], params["atol"])
    else:
        updates = vc.to_numpy(case["update_batch"], "update")
        expected = vc.to_numpy(case[kind + "_batch"], kind)
        for threading in (-1, 2):
            res = model.calculate(sym=sym, method=method, update=vc.batch_update_arrays(updates), threading=threading)
            assert res["n_failed"] == 0, res["error"]
            for s, exp in enumerate(expected):
                vc.compare_result({k: v[s] for k, v in res.items() if isinstance(v, np.ndarray) and v.ndim == 2}, exp,
                                  params["rtol"], params["atol"])
        # the model is unchanged after a batch (update -> restore per scenario)
        if kind in case:
            res = model.calculate(sym=sym, method=method)
            vc.compare_result({k: v[0] for k, v in res.items() if isinstance(v, np.ndarray) and v.ndim == 2},
                              vc.to_numpy(case[kind], kind), params["rtol"], params["atol"])
