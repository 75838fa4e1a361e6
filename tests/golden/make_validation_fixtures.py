"""Collects the reference's power-flow validation cases (tests/data/power_flow/**) whose components are in the PF subset
this repo implements into tests/golden/power_flow_cases.json, and the cases of the automatic tap changer
(automatic-tap-regulator/*, params.tap_changing_strategy; cases that expect MaxIterationReached keep their `raises` entry) into
tests/golden/tap_regulator_cases.json, and the cases in which the reference itself expects a named exception (params.raises /
params.xfail.raises other than AssertionError) into tests/golden/error_cases.json together with the text that exception carries
(common/exception.hpp).

Run here (the build container has /root/reference; the GPU box does not):
    python tests/golden/make_validation_fixtures.py
Each case keeps: params (methods, rtol, atol), input, optional update_batch and the reference's golden outputs
(sym_output / asym_output / *_batch). Attributes are stored as the reference stores them (row dicts, or compact rows
with an `attributes` table).  `inf`-like strings are kept verbatim and decoded by the test loader.
"""
import json
import os
import sys

REF = "/root/reference/tests/data/power_flow"
SUPPORTED = {"node", "line", "transformer", "source", "shunt", "sym_load", "sym_gen", "asym_load", "asym_gen",
             "voltage_regulator", "asym_line", "generic_branch", "link", "three_winding_transformer"}
IGNORED_INPUT = {"fault", "sym_voltage_sensor", "sym_power_sensor", "asym_voltage_sensor", "asym_power_sensor"}  # not used by PF
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "power_flow_cases.json")
OUT_TAP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tap_regulator_cases.json")
TAP_SUPPORTED = SUPPORTED | {"transformer_tap_regulator"}
OUT_ERR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "error_cases.json")
# what() of the exception classes the cases name (power_grid_model/common/exception.hpp:124-127, 161-186, 233-236)
ERROR_TEXT = {
    "IDNotFound": "The id cannot be found: ",
    "InvalidCalculationMethod": "The calculation method is invalid for this calculation!",
    "UnsupportedLoadGenTypeForVoltageRegulator": "Unsupported load_gen type for voltage regulators ",
    "UnsupportedVoltageRegulatorSourceCombinationError": "Nodes with a source and a voltage regulated load/generator are not supported when both are enabled. Found at node with id ",
    "UnsupportedRegulatorCombinationError": "The combination of voltage regulators and transformer tap regulators is not supported in the same model.",
}


def main():
    cases = {}
    tap_cases = {}
    error_cases = {}
    skipped = {}
    for dirpath, _, files in sorted(os.walk(REF)):
        if "params.json" not in files or "input.json" not in files:
            continue
        name = os.path.relpath(dirpath, REF)
        params = json.load(open(os.path.join(dirpath, "params.json")))
        inp = json.load(open(os.path.join(dirpath, "input.json")))
        comps = set(inp["data"].keys())
        is_tap = "tap_changing_strategy" in params
        unsupported = comps - (TAP_SUPPORTED if is_tap else SUPPORTED) - IGNORED_INPUT
        if unsupported:
            skipped[name] = "unsupported components: " + ", ".join(sorted(unsupported))
            continue
        expected = (params.get("raises") or params.get("xfail") or {}).get("raises")
        if expected in ERROR_TEXT and not (is_tap and "raises" in params):
            case = {"params": params, "input": inp, "error_text": ERROR_TEXT[expected]}
            p = os.path.join(dirpath, "update_batch.json")
            if os.path.exists(p):
                case["update_batch"] = json.load(open(p))
            error_cases[name] = case
            skipped[name] = "expects " + expected + ": kept in error_cases.json"
            continue
        if "xfail" in params or ("raises" in params and not is_tap):
            skipped[name] = "expects an error / known failure of the reference"
            continue
        case = {"params": params, "input": inp}
        for f in ("update_batch", "sym_output", "asym_output", "sym_output_batch", "asym_output_batch"):
            p = os.path.join(dirpath, f + ".json")
            if os.path.exists(p):
                case[f] = json.load(open(p))
        (tap_cases if is_tap else cases)[name] = case
    json.dump({"source": "PowerGridModel/power-grid-model tests/data/power_flow (MPL-2.0)", "cases": cases, "skipped": skipped},
              open(OUT, "w"), separators=(",", ":"))
    json.dump({"source": "PowerGridModel/power-grid-model tests/data/power_flow/automatic-tap-regulator (MPL-2.0)", "cases": tap_cases},
              open(OUT_TAP, "w"), separators=(",", ":"))
    json.dump({"source": "PowerGridModel/power-grid-model tests/data/power_flow (MPL-2.0)", "cases": error_cases},
              open(OUT_ERR, "w"), separators=(",", ":"))
    print(f"{len(error_cases)} error cases -> {OUT_ERR} ({os.path.getsize(OUT_ERR)} bytes)")
    print(f"{len(cases)} cases -> {OUT} ({os.path.getsize(OUT)} bytes); skipped {len(skipped)}")
    print(f"{len(tap_cases)} tap changer cases -> {OUT_TAP} ({os.path.getsize(OUT_TAP)} bytes)")
    for k, v in skipped.items():
        print("  skipped", k, ":", v)


if __name__ == "__main__":
    sys.exit(main())
