"""numpy dtypes of the PF-relevant component structs at the power-grid-model C-API boundary.

Layouts follow the reference's generated headers (natural C alignment):
``power_grid_model/auxiliary/input.hpp``, ``update.hpp``, ``output.hpp`` (source of truth:
``code_generation/data/attribute_classes/{input,update,output}.json``); the reference's Python wrapper builds the
same dtypes at run time from ``PGM_meta_*`` (``src/power_grid_model/_core/power_grid_meta.py``).
"""
import numpy as np

NA_INT_ID = np.iinfo(np.int32).min
NA_INT_S = np.iinfo(np.int8).min


def _dt(fields):
    return np.dtype(fields, align=True)


_branch_head = [("id", "i4"), ("from_node", "i4"), ("to_node", "i4"), ("from_status", "i1"), ("to_status", "i1")]
_appliance_head = [("id", "i4"), ("node", "i4"), ("status", "i1")]

INPUT = {
    "node": _dt([("id", "i4"), ("u_rated", "f8")]),
    "line": _dt(_branch_head + [(n, "f8") for n in ("r1", "x1", "c1", "tan1", "r0", "x0", "c0", "tan0", "i_n")]),
    "transformer": _dt(
        _branch_head
        + [(n, "f8") for n in ("u1", "u2", "sn", "uk", "pk", "i0", "p0", "i0_zero_sequence", "p0_zero_sequence")]
        + [(n, "i1") for n in ("winding_from", "winding_to", "clock", "tap_side", "tap_pos", "tap_min", "tap_max", "tap_nom")]
        + [
            (n, "f8")
            for n in (
                "tap_size", "uk_min", "uk_max", "pk_min", "pk_max",
                "r_grounding_from", "x_grounding_from", "r_grounding_to", "x_grounding_to",
            )
        ]
    ),
    "source": _dt(_appliance_head + [(n, "f8") for n in ("u_ref", "u_ref_angle", "sk", "rx_ratio", "z01_ratio")]),
    "shunt": _dt(_appliance_head + [(n, "f8") for n in ("g1", "b1", "g0", "b0")]),
    "sym_load": _dt(_appliance_head + [("type", "i1"), ("p_specified", "f8"), ("q_specified", "f8")]),
    "asym_load": _dt(_appliance_head + [("type", "i1"), ("p_specified", "f8", (3,)), ("q_specified", "f8", (3,))]),
}
# transformer_tap_regulator (automatic tap changer, calculate_power_flow(tap_changing_strategy=...)); control_side: 0 from /
# side_1, 1 to / side_2, 2 side_3
INPUT["transformer_tap_regulator"] = _dt([("id", "i4"), ("regulated_object", "i4"), ("status", "i1"), ("control_side", "i1"),
                                          ("u_set", "f8"), ("u_band", "f8"), ("line_drop_compensation_r", "f8"),
                                          ("line_drop_compensation_x", "f8")])
# voltage_regulator (PV buses with reactive-power limits, Newton-Raphson only)
INPUT["voltage_regulator"] = _dt([("id", "i4"), ("regulated_object", "i4"), ("status", "i1"), ("u_ref", "f8"), ("q_min", "f8"),
                                  ("q_max", "f8")])
INPUT["sym_gen"] = INPUT["sym_load"]
INPUT["asym_gen"] = INPUT["asym_load"]
# further branch components (auxiliary/input.hpp:99-165)
INPUT["asym_line"] = _dt(_branch_head + [(n, "f8") for n in (
    "r_aa", "r_ba", "r_bb", "r_ca", "r_cb", "r_cc", "r_na", "r_nb", "r_nc", "r_nn",
    "x_aa", "x_ba", "x_bb", "x_ca", "x_cb", "x_cc", "x_na", "x_nb", "x_nc", "x_nn",
    "c_aa", "c_ba", "c_bb", "c_ca", "c_cb", "c_cc", "c0", "c1", "i_n")])
INPUT["generic_branch"] = _dt(_branch_head + [(n, "f8") for n in ("r1", "x1", "g1", "b1", "k", "theta", "sn")])
INPUT["link"] = _dt(_branch_head)
INPUT["three_winding_transformer"] = _dt(
    [("id", "i4"), ("node_1", "i4"), ("node_2", "i4"), ("node_3", "i4"), ("status_1", "i1"), ("status_2", "i1"), ("status_3", "i1")]
    + [(n, "f8") for n in ("u1", "u2", "u3", "sn_1", "sn_2", "sn_3", "uk_12", "uk_13", "uk_23", "pk_12", "pk_13", "pk_23", "i0", "p0")]
    + [(n, "i1") for n in ("winding_1", "winding_2", "winding_3", "clock_12", "clock_13", "tap_side", "tap_pos", "tap_min", "tap_max",
                            "tap_nom")]
    + [(n, "f8") for n in ("tap_size", "uk_12_min", "uk_12_max", "uk_13_min", "uk_13_max", "uk_23_min", "uk_23_max", "pk_12_min",
                            "pk_12_max", "pk_13_min", "pk_13_max", "pk_23_min", "pk_23_max", "r_grounding_1", "x_grounding_1",
                            "r_grounding_2", "x_grounding_2", "r_grounding_3", "x_grounding_3")])

UPDATE = {
    "line": _dt([("id", "i4"), ("from_status", "i1"), ("to_status", "i1")]),
    "transformer": _dt([("id", "i4"), ("from_status", "i1"), ("to_status", "i1"), ("tap_pos", "i1")]),
    "source": _dt([("id", "i4"), ("status", "i1")] + [(n, "f8") for n in ("u_ref", "u_ref_angle", "sk", "rx_ratio", "z01_ratio")]),
    "shunt": _dt([("id", "i4"), ("status", "i1")] + [(n, "f8") for n in ("g1", "b1", "g0", "b0")]),
    "sym_load": _dt([("id", "i4"), ("status", "i1"), ("p_specified", "f8"), ("q_specified", "f8")]),
    "asym_load": _dt([("id", "i4"), ("status", "i1"), ("p_specified", "f8", (3,)), ("q_specified", "f8", (3,))]),
}
UPDATE["transformer_tap_regulator"] = _dt([("id", "i4"), ("status", "i1"), ("u_set", "f8"), ("u_band", "f8"),
                                           ("line_drop_compensation_r", "f8"), ("line_drop_compensation_x", "f8")])
UPDATE["voltage_regulator"] = _dt([("id", "i4"), ("status", "i1"), ("u_ref", "f8"), ("q_min", "f8"), ("q_max", "f8")])
UPDATE["sym_gen"] = UPDATE["sym_load"]
UPDATE["asym_gen"] = UPDATE["asym_load"]
UPDATE["asym_line"] = UPDATE["line"]
UPDATE["generic_branch"] = UPDATE["line"]
UPDATE["link"] = UPDATE["line"]
UPDATE["three_winding_transformer"] = _dt([("id", "i4"), ("status_1", "i1"), ("status_2", "i1"), ("status_3", "i1"), ("tap_pos", "i1")])


def _real(sym):
    return ("f8",) if sym else ("f8", (3,))


def output_dtypes(sym: bool):
    r = _real(sym)
    node = _dt([("id", "i4"), ("energized", "i1")] + [(n, *r) for n in ("u_pu", "u", "u_angle", "p", "q")])
    branch = _dt(
        [("id", "i4"), ("energized", "i1"), ("loading", "f8")]
        + [(n, *r) for n in ("p_from", "q_from", "i_from", "s_from", "p_to", "q_to", "i_to", "s_to")]
    )
    appliance = _dt([("id", "i4"), ("energized", "i1")] + [(n, *r) for n in ("p", "q", "i", "s", "pf")])
    branch3 = _dt([("id", "i4"), ("energized", "i1")] + [(n, "f8") for n in ("loading_1", "loading_2", "loading_3", "loading")]
                  + [(f"{n}_{k}", *r) for k in (1, 2, 3) for n in ("p", "q", "i", "s")])
    return {
        "node": node, "line": branch, "transformer": branch, "asym_line": branch, "generic_branch": branch, "link": branch,
        "three_winding_transformer": branch3,
        "shunt": appliance, "source": appliance,
        "sym_gen": appliance, "asym_gen": appliance, "sym_load": appliance, "asym_load": appliance,
        "voltage_regulator": _dt([("id", "i4"), ("energized", "i1"), ("limit_violated", "i1")]),
        "transformer_tap_regulator": _dt([("id", "i4"), ("energized", "i1"), ("tap_pos", "i1")]),
    }


SYM_OUTPUT = output_dtypes(True)
ASYM_OUTPUT = output_dtypes(False)

# component storage order of the reference (all_components.hpp:36-39), PF subset
COMPONENT_ORDER = ("node", "line", "transformer", "shunt", "source", "sym_gen", "asym_gen", "sym_load", "asym_load",
                   "voltage_regulator", "asym_line", "generic_branch", "link", "three_winding_transformer",
                   "transformer_tap_regulator")
UPDATABLE = ("line", "transformer", "shunt", "source", "sym_gen", "asym_gen", "sym_load", "asym_load", "voltage_regulator",
             "asym_line", "generic_branch", "link", "three_winding_transformer", "transformer_tap_regulator")


def initialize_array(kind: str, component: str, shape, sym: bool = True):
    """Array filled with the reference's null values (NaN / INT_MIN), like power_grid_model.initialize_array."""
    table = {"input": INPUT, "update": UPDATE, "sym_output": SYM_OUTPUT, "asym_output": ASYM_OUTPUT}[kind]
    arr = np.zeros(shape, dtype=table[component])
    for name in arr.dtype.names:
        base = arr.dtype[name].base
        if base == np.float64:
            arr[name] = np.nan
        elif base == np.int32:
            arr[name] = NA_INT_ID
        else:
            arr[name] = NA_INT_S
    return arr


assert INPUT["asym_line"].itemsize == 248 and INPUT["generic_branch"].itemsize == 72
assert INPUT["link"].itemsize == 16 and INPUT["three_winding_transformer"].itemsize == 304
assert UPDATE["three_winding_transformer"].itemsize == 8
assert SYM_OUTPUT["three_winding_transformer"].itemsize == 136 and ASYM_OUTPUT["three_winding_transformer"].itemsize == 328
assert INPUT["voltage_regulator"].itemsize == 40 and UPDATE["voltage_regulator"].itemsize == 32
assert SYM_OUTPUT["voltage_regulator"].itemsize == 8
assert INPUT["transformer_tap_regulator"].itemsize == 48 and UPDATE["transformer_tap_regulator"].itemsize == 40
assert SYM_OUTPUT["transformer_tap_regulator"].itemsize == 8
assert INPUT["line"].itemsize == 88 and INPUT["transformer"].itemsize == 168 and INPUT["source"].itemsize == 56
assert UPDATE["sym_load"].itemsize == 24 and UPDATE["asym_load"].itemsize == 56
assert SYM_OUTPUT["node"].itemsize == 48 and ASYM_OUTPUT["node"].itemsize == 128
assert SYM_OUTPUT["line"].itemsize == 80 and ASYM_OUTPUT["line"].itemsize == 208
