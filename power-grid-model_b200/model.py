"""Model-level host API: the reference's ``PowerGridModel`` (src/power_grid_model/_core/power_grid_model.py) for the PF
component subset, same method and argument names, backed by libpgm_b200.so (CUDA).  Datasets are dicts
``component -> numpy structured array`` with the dtypes of :mod:`pgm_b200.structs` (== the reference's)."""
import ctypes as C

import numpy as np

from . import _lib, structs
from ._lib import BatchError, PgmB200Error, check, lib


def _buffer(arr):
    return _lib.ComponentBufferC(len(arr), None, arr.ctypes.data if len(arr) else None)


class PowerGridModel:
    def __init__(self, input_data: dict, system_frequency: float = 50.0):
        self._keep = {}
        self._pinned_out = {}
        inp = _lib.InputDataC()
        self._counts = {}
        for c in structs.COMPONENT_ORDER:
            arr = input_data.get(c)
            arr = np.zeros(0, structs.INPUT[c]) if arr is None else np.ascontiguousarray(arr, dtype=structs.INPUT[c])
            self._keep[c] = arr
            self._counts[c] = len(arr)
            setattr(inp, c, _buffer(arr))
        unknown = set(input_data) - set(structs.COMPONENT_ORDER)
        if unknown:
            raise PgmB200Error(_lib.PGMB_ERR_INVALID, f"components not supported by pgm_b200: {sorted(unknown)}")
        self._h = C.c_void_p()
        check(lib().pgmb_model_create(C.c_double(system_frequency), C.byref(inp), C.byref(self._h)))
        self.batch_error = None

    def close(self):
        if getattr(self, "_h", None):
            lib().pgmb_model_destroy(self._h)
            self._h = None

    __del__ = close

    # -- updates ----------------------------------------------------------------------------------------------------
    def _update_struct(self, update_data, batch):
        upd = _lib.UpdateDataC()
        keep = []
        n_scn = None
        for c, val in update_data.items():
            if c not in structs.UPDATABLE:
                raise PgmB200Error(_lib.PGMB_ERR_INVALID, f"component {c} cannot be updated")
            if isinstance(val, dict):  # sparse: {"data": 1-D, "indptr": (n_scn + 1,)}
                data = np.ascontiguousarray(val["data"], dtype=structs.UPDATE[c])
                indptr = np.ascontiguousarray(val["indptr"], dtype=np.int64)
                keep += [data, indptr]
                ns = len(indptr) - 1
                buf = _lib.ComponentBufferC(-1, indptr.ctypes.data, data.ctypes.data if len(data) else None)
                if not len(data):
                    continue
            else:
                data = np.ascontiguousarray(val, dtype=structs.UPDATE[c])
                if not batch:
                    data = data.reshape(1, -1)
                if data.ndim != 2:
                    raise PgmB200Error(_lib.PGMB_ERR_INVALID, f"batch update of {c} must be 2-D (n_scenarios, n_elements)")
                keep.append(data)
                ns = data.shape[0]
                if data.shape[1] == 0:
                    n_scn = ns if n_scn is None else n_scn
                    continue
                buf = _lib.ComponentBufferC(data.shape[1], None, data.ctypes.data)
            if n_scn is not None and ns != n_scn:
                raise PgmB200Error(_lib.PGMB_ERR_INVALID, "inconsistent batch sizes in update data")
            n_scn = ns
            setattr(upd, c, buf)
        upd.n_scenarios = 0 if n_scn is None else n_scn
        return upd, keep

    def update(self, *, update_data: dict):
        """Permanent update (PGM_update_model)."""
        upd, keep = self._update_struct(update_data, batch=False)
        check(lib().pgmb_model_update(self._h, C.byref(upd)))

    # -- calculation ------------------------------------------------------------------------------------------------
    def calculate_power_flow(self, *, symmetric=True, error_tolerance=1e-8, max_iterations=20,
                             calculation_method="newton_raphson", update_data=None, threading=-1,
                             output_component_types=None, continue_on_batch_error=False, device=0, output_buffers=None,
                             reuse_output_buffers=False, n_devices=1, flags=0, tap_changing_strategy=0):
        """Same contract as the reference: without ``update_data`` a single calculation returning 1-D arrays, with it a
        batch returning (n_scenarios, n_elements) arrays.  ``threading`` is accepted for signature compatibility; the
        scenario loop of load / source-reference batches runs on the GPU whatever ``threading`` says; batches that switch
        branches or change parameters are dispatched over host threads with one model copy each (``threading``: -1 / 0 = all
        cores, n = n threads), every copy driving its own GPU stream.  ``output_buffers``: caller-owned arrays per component (like the C API; use
        ``pgm_b200.pinned_empty`` for page-locked ones).  ``reuse_output_buffers=True``: results are written into page-locked
        arrays owned by the model and reused by the next call with the same shapes (no page faults, transfers overlap the
        solver) -- copy what must outlive the next calculation.  ``n_devices``: GPUs the batch is spread over inside the call
        (contiguous scenario blocks, starting at ``device``).  ``flags``: ``pgm_b200.FLAG_RESIDENT_INPUT / _OUTPUT`` (device-resident
        update rows / output structs, include/pgm_b200.h).  ``tap_changing_strategy``: "disabled" / 0, "any_valid_tap" / 1,
        "min_voltage_tap" / 2, "max_voltage_tap" / 3, "fast_any_tap" / 4 -- the automatic tap changer of the reference around
        every scenario's power flow (needs ``transformer_tap_regulator`` components)."""
        if isinstance(tap_changing_strategy, str):
            tap_changing_strategy = _lib.TAP_STRATEGIES[tap_changing_strategy]
        if isinstance(calculation_method, str):
            calculation_method = _lib.METHODS[calculation_method]
        opt = _lib.OptionsC(int(calculation_method), int(bool(symmetric)), float(error_tolerance), int(max_iterations),
                            int(n_devices), int(device), int(threading), int(flags), int(tap_changing_strategy))
        upd = None
        n_scn = 1
        keep = None
        if update_data is not None:
            upd, keep = self._update_struct(update_data, batch=True)
            n_scn = upd.n_scenarios
        table = structs.SYM_OUTPUT if symmetric else structs.ASYM_OUTPUT
        comps = output_component_types
        if comps is None:
            comps = [c for c in structs.COMPONENT_ORDER if self._counts[c]]
        out = _lib.OutputDataC()
        result = {}
        for c in comps:
            if output_buffers is not None and c in output_buffers:  # caller-owned (e.g. pinned) buffers, like the C API
                arr = output_buffers[c]
                assert arr.dtype == table[c] and arr.shape == (n_scn, self._counts[c]) and arr.flags.c_contiguous
            elif reuse_output_buffers:
                key = (c, n_scn, bool(symmetric))
                if key not in self._pinned_out:
                    self._pinned_out[key] = _lib.pinned_empty((n_scn, self._counts[c]), table[c])
                arr = self._pinned_out[key]
            else:
                arr = np.empty((n_scn, self._counts[c]), dtype=table[c])
            result[c] = arr
            if arr.size:
                setattr(out, c, arr.ctypes.data)
        self.n_iter = np.zeros(max(n_scn, 1), np.int32)
        self.status = np.zeros(max(n_scn, 1), np.int32)
        rc = lib().pgmb_model_calculate(self._h, C.byref(opt), C.byref(upd) if upd is not None else None, C.byref(out),
                                        self.n_iter.ctypes.data_as(C.c_void_p), self.status.ctypes.data_as(C.c_void_p))
        self.batch_error = None
        if rc == _lib.PGMB_ERR_BATCH:
            err = BatchError(lib().pgmb_last_error().decode(), self.status.copy())
            if update_data is None or not continue_on_batch_error:
                raise err
            self.batch_error = err
        else:
            check(rc)
        if update_data is None:
            result = {c: a[0] for c, a in result.items()}
        return result

    def timing(self):
        """milliseconds of the last calculate: host prepare, pipeline enqueue, solver kernels (sum over overlapping chunks),
        unused, pipeline drain + status, total (include/pgm_b200.h: pgmb_model_last_timing)"""
        t = (C.c_double * 6)()
        check(lib().pgmb_model_last_timing(self._h, t))
        out = dict(zip(("prepare", "enqueue", "solve_kernel", "output", "drain", "total"), list(t)))
        ms = C.c_double()
        check(lib().pgmb_model_device_pipeline_ms(self._h, C.byref(ms)))
        out["device_pipeline"] = ms.value
        return out

    def batch_pf_input(self, update_data, symmetric=True, group=0):
        """PowerFlowInput of every scenario for one math group: (s_injection (n_scn, n_load_gen, B), u_ref (n_scn, n_source))"""
        upd, keep = self._update_struct(update_data, batch=True)
        B = 1 if symmetric else 3
        n_lg = int(self.math_index(group, "load_gens_per_bus")[-1])
        n_src = int(self.math_index(group, "sources_per_bus")[-1])
        s = np.zeros((upd.n_scenarios, n_lg, B), np.complex128)
        u = np.zeros((upd.n_scenarios, n_src), np.complex128)
        check(lib().pgmb_model_batch_pf_input(self._h, C.byref(upd), C.c_int32(int(symmetric)), C.c_int64(group),
                                              s.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p)))
        return s, u

    # -- introspection (parity tests) ---------------------------------------------------------------------------------
    def outage_plan(self, update_data, symmetric=True):
        """host planning of a branch-switching batch: (n_scn, 4) = route (0 shared pattern / 1 own topology), overlay slots,
        buses without supply, planned on the union grid"""
        upd, keep = self._update_struct(update_data, batch=True)
        plan = np.zeros((upd.n_scenarios, 4), np.int64)
        check(lib().pgmb_model_outage_plan(self._h, C.byref(upd), C.c_int32(int(symmetric)), plan.ctypes.data_as(C.c_void_p)))
        return plan

    def n_math_groups(self):
        return int(lib().pgmb_model_n_math_groups(self._h))

    def math_index(self, group, name):
        ptr, n = C.POINTER(C.c_int64)(), C.c_int64()
        check(lib().pgmb_model_get_index(self._h, C.c_int64(group), name.encode(), C.byref(ptr), C.byref(n)))
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int64)

    def math_real(self, group, symmetric, name):
        ptr, n = C.POINTER(C.c_double)(), C.c_int64()
        check(lib().pgmb_model_get_real(self._h, C.c_int64(group), C.c_int32(int(symmetric)), name.encode(), C.byref(ptr), C.byref(n)))
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float64)
