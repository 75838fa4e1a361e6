"""The reference's power-flow validation cases that must END IN A NAMED EXCEPTION (params.raises / params.xfail.raises:
IDNotFound, InvalidCalculationMethod, UnsupportedLoadGenTypeForVoltageRegulator,
UnsupportedVoltageRegulatorSourceCombinationError, UnsupportedRegulatorCombinationError), committed as
tests/golden/error_cases.json with the text each exception carries (common/exception.hpp): the oracle and the product refuse
them with the reference's message.  The product's checks run on the host before any device work, so these run on CPU."""
import json
import os

import pytest

import oracle_lib as orc
import pgm_b200
import validation_cases as vc

CASES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "error_cases.json")))["cases"]
# (case, method, symmetric)
RUNS = []
for _name, _case in sorted(CASES.items()):
    _methods = _case["params"]["calculation_method"]
    for _m in (_methods if isinstance(_methods, list) else [_methods]):
        RUNS.append((_name, _m))
POWER_FLOW_METHODS = ("newton_raphson", "linear", "iterative_current", "linear_current")


def test_the_fixture_holds_the_five_cases():
    assert sorted(CASES) == ["non-existent-id-update-batch", "pv-node/pv-node-calc-methods", "pv-node/pv-node-load-gen-const-i-y",
                             "pv-node/pv-node-source", "pv-node/pv-node-tap-regulator"]


@pytest.mark.parametrize("name,method", RUNS, ids=[f"{n}-{m}" for n, m in RUNS])
def test_oracle_refuses_with_the_reference_text(name, method):
    case = CASES[name]
    if method not in POWER_FLOW_METHODS:
        pytest.skip("a state-estimation method name: refused by the option parser, not by the calculation")
    model = orc.Model(vc.to_numpy(case["input"], "input"))
    kw = {}
    if "update_batch" in case:
        kw["update"] = vc.batch_update_arrays(vc.to_numpy(case["update_batch"], "update"))
    if "tap_changing_strategy" in case["params"]:
        kw["tap_changing_strategy"] = case["params"]["tap_changing_strategy"]
    try:
        res = model.calculate(sym=True, method=method, **kw)
        message = res["error"] if res["n_failed"] else ""
    except orc.OracleError as ex:
        message = str(ex)
    assert case["error_text"] in message, message


@pytest.mark.parametrize("name,method", RUNS, ids=[f"{n}-{m}" for n, m in RUNS])
def test_product_refuses_with_the_reference_text(name, method):
    case = CASES[name]
    if method not in POWER_FLOW_METHODS:
        pytest.skip("a state-estimation method name: refused by the option parser, not by the calculation")
    if name == "non-existent-id-update-batch":
        pytest.skip("batch updates are resolved scenario by scenario on the device route: the unknown id of a batch is covered by "
                    "tests/test_gpu_pgm_capi.py (PowerGridBatchError, 'The id cannot be found:')")
    model = pgm_b200.PowerGridModel(vc.to_numpy(case["input"], "input"))
    kw = {}
    if "update_batch" in case:
        kw["update_data"] = vc.batch_update_arrays(vc.to_numpy(case["update_batch"], "update"))
    if "tap_changing_strategy" in case["params"]:
        kw["tap_changing_strategy"] = case["params"]["tap_changing_strategy"]
    with pytest.raises(pgm_b200.PgmB200Error) as info:
        model.calculate_power_flow(symmetric=True, calculation_method=method, **kw)
    assert case["error_text"] in str(info.value), str(info.value)
