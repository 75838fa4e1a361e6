"""The reference's benchmark grid (tests/benchmark_cpp/fictional_grid_generator.hpp) through the C-ABI generator."""
import ctypes as C

import numpy as np

from . import _lib, structs
from ._lib import check, lib

# tests/benchmark_cpp/benchmark.cpp:257-263 (release configuration)
BENCHMARK_OPTION = dict(n_node_total_specified=1500, n_mv_feeder=20, n_node_per_mv_feeder=10, n_lv_feeder=10,
                        n_connection_per_lv_feeder=40)


class FictionalGrid:
    def __init__(self, *, n_node_total_specified=1500, n_mv_feeder=20, n_node_per_mv_feeder=10, n_lv_feeder=10,
                 n_connection_per_lv_feeder=40, has_mv_ring=False, has_lv_ring=False, has_tap_changer=False, seed=0):
        opt = _lib.GridOptionC(n_node_total_specified, n_mv_feeder, n_node_per_mv_feeder, n_lv_feeder,
                               n_connection_per_lv_feeder, int(has_mv_ring), int(has_lv_ring))
        self._h = C.c_void_p()
        check(lib().pgmb_fictional_grid_create(C.byref(opt), C.c_uint32(seed), C.byref(self._h)))
        self.input_data = {}
        for c in ("node", "line", "transformer", "shunt", "source", "sym_load", "asym_load"):
            ptr, n = C.c_void_p(), C.c_int64()
            check(lib().pgmb_fictional_grid_get(self._h, c.encode(), C.byref(ptr), C.byref(n)))
            dt = structs.INPUT[c]
            if n.value:
                buf = (C.c_char * (n.value * dt.itemsize)).from_address(ptr.value)
                self.input_data[c] = np.frombuffer(buf, dtype=dt).copy()
            else:
                self.input_data[c] = np.zeros(0, dt)
        if has_tap_changer and len(self.input_data["transformer"]):
            # generate_tap_changer (fictional_grid_generator.hpp:615-633): one regulator on the first transformer, control side
            # "to", set point 1.1 x rated voltage, band = one tap step + 5 % of the rated voltage
            trafo = self.input_data["transformer"][0]
            node = self.input_data["node"]
            u_rated = float(node["u_rated"][node["id"] == trafo["to_node"]][0])
            reg = structs.initialize_array("input", "transformer_tap_regulator", 1)
            reg["id"] = 1 + max(int(a["id"].max()) for a in self.input_data.values() if len(a))
            reg["regulated_object"], reg["status"], reg["control_side"] = trafo["id"], 1, 1
            reg["u_set"], reg["u_band"] = 1.1 * u_rated, float(trafo["tap_size"]) + 0.05 * u_rated
            self.input_data["transformer_tap_regulator"] = reg

    def close(self):
        if getattr(self, "_h", None):
            lib().pgmb_fictional_grid_destroy(self._h)
            self._h = None

    __del__ = close

    def batch_update(self, batch_size, seed=0):
        """generate_batch_input: load-profile update dataset {sym_load, asym_load} of shape (batch_size, n)"""
        sym = np.zeros((batch_size, len(self.input_data["sym_load"])), structs.UPDATE["sym_load"])
        asym = np.zeros((batch_size, len(self.input_data["asym_load"])), structs.UPDATE["asym_load"])
        check(lib().pgmb_fictional_grid_batch(self._h, C.c_int64(batch_size), C.c_uint32(seed),
                                              sym.ctypes.data_as(C.c_void_p), asym.ctypes.data_as(C.c_void_p)))
        return {"sym_load": sym, "asym_load": asym}
