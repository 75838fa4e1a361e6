"""ctypes binding over the reference's own C API names (include/pgm_b200_capi.h), written the way the reference's Python
wrapper drives its core library (src/power_grid_model/_core/power_grid_core.py: one handle, every call takes it first,
errors read back from it; _core/power_grid_model.py: datasets built from numpy row buffers, PGM_calculate).
It exists to show -- and to test -- that a client of `PGM_create_model / PGM_update_model / PGM_calculate` runs unchanged
against libpgm_b200.so.  Row-based numpy buffers only."""
import ctypes as C

import numpy as np

from . import structs
from ._lib import LIB_PATH, lib

PGM_NO_ERROR, PGM_REGULAR_ERROR, PGM_BATCH_ERROR, PGM_SERIALIZATION_ERROR = range(4)
CALCULATION_METHOD = {"default_method": -128, "linear": 0, "newton_raphson": 1, "iterative_linear": 2,
                      "iterative_current": 3, "linear_current": 4, "iec60909": 5}


class PowerGridError(RuntimeError):
    """PGM_regular_error"""


class PowerGridBatchError(PowerGridError):
    """PGM_batch_error: failed_scenarios / error_messages like the reference's PowerGridBatchError"""

    def __init__(self, message, failed_scenarios, error_messages):
        super().__init__(message)
        self.failed_scenarios = failed_scenarios
        self.error_messages = error_messages


_core = None


def core():
    global _core
    if _core is None:
        lib()  # loud failure with the build hint when the library is missing
        l = C.CDLL(LIB_PATH)
        P, I, D, S = C.c_void_p, C.c_int64, C.c_double, C.c_char_p
        sig = {
            "PGM_create_handle": (P, []), "PGM_destroy_handle": (None, [P]), "PGM_error_code": (I, [P]),
            "PGM_error_message": (S, [P]), "PGM_n_failed_scenarios": (I, [P]), "PGM_failed_scenarios": (C.POINTER(I), [P]),
            "PGM_batch_errors": (C.POINTER(S), [P]), "PGM_clear_error": (None, [P]), "PGM_version": (S, []),
            "PGM_create_options": (P, [P]), "PGM_destroy_options": (None, [P]),
            "PGM_set_calculation_type": (None, [P, P, I]), "PGM_set_calculation_method": (None, [P, P, I]),
            "PGM_set_symmetric": (None, [P, P, I]), "PGM_set_err_tol": (None, [P, P, D]), "PGM_set_max_iter": (None, [P, P, I]),
            "PGM_set_threading": (None, [P, P, I]), "PGM_set_short_circuit_voltage_scaling": (None, [P, P, I]),
            "PGM_set_tap_changing_strategy": (None, [P, P, I]), "PGM_set_experimental_features": (None, [P, P, I]),
            "PGM_create_dataset_const": (P, [P, S, I, I]), "PGM_create_dataset_const_from_mutable": (P, [P, P]),
            "PGM_destroy_dataset_const": (None, [P]), "PGM_dataset_const_add_buffer": (None, [P, P, S, I, I, P, P]),
            "PGM_dataset_const_add_attribute_buffer": (None, [P, P, S, S, P]),
            "PGM_dataset_const_set_next_cartesian_product_dimension": (None, [P, P, P]),
            "PGM_create_dataset_mutable": (P, [P, S, I, I]), "PGM_destroy_dataset_mutable": (None, [P]),
            "PGM_dataset_mutable_add_buffer": (None, [P, P, S, I, I, P, P]),
            "PGM_dataset_mutable_add_attribute_buffer": (None, [P, P, S, S, P]),
            "PGM_meta_n_datasets": (I, [P]), "PGM_meta_get_dataset_by_idx": (P, [P, I]), "PGM_meta_get_dataset_by_name": (P, [P, S]),
            "PGM_meta_dataset_name": (S, [P, P]), "PGM_meta_n_components": (I, [P, P]),
            "PGM_meta_get_component_by_idx": (P, [P, P, I]), "PGM_meta_get_component_by_name": (P, [P, S, S]),
            "PGM_meta_component_name": (S, [P, P]), "PGM_meta_component_size": (C.c_size_t, [P, P]),
            "PGM_meta_component_alignment": (C.c_size_t, [P, P]), "PGM_meta_n_attributes": (I, [P, P]),
            "PGM_meta_get_attribute_by_idx": (P, [P, P, I]), "PGM_meta_get_attribute_by_name": (P, [P, S, S, S]),
            "PGM_meta_attribute_name": (S, [P, P]), "PGM_meta_attribute_ctype": (I, [P, P]),
            "PGM_meta_attribute_offset": (C.c_size_t, [P, P]), "PGM_is_little_endian": (C.c_int, [P]),
            "PGM_create_buffer": (P, [P, P, I]), "PGM_destroy_buffer": (None, [P]), "PGM_buffer_set_nan": (None, [P, P, P, I, I]),
            "PGM_buffer_set_value": (None, [P, P, P, P, I, I, I]), "PGM_buffer_get_value": (None, [P, P, P, P, I, I, I]),
            "PGM_dataset_info_name": (S, [P, P]), "PGM_dataset_info_is_batch": (I, [P, P]), "PGM_dataset_info_batch_size": (I, [P, P]),
            "PGM_dataset_info_n_components": (I, [P, P]), "PGM_dataset_info_component_name": (S, [P, P, I]),
            "PGM_dataset_info_elements_per_scenario": (I, [P, P, I]), "PGM_dataset_info_total_elements": (I, [P, P, I]),
            "PGM_dataset_info_has_attribute_indications": (I, [P, P, I]),
            "PGM_dataset_info_n_attribute_indications": (I, [P, P, I]), "PGM_dataset_info_attribute_name": (S, [P, P, I, I]),
            "PGM_dataset_const_get_info": (P, [P, P]), "PGM_dataset_mutable_get_info": (P, [P, P]),
            "PGM_create_model": (P, [P, D, P]), "PGM_update_model": (None, [P, P, P]), "PGM_copy_model": (P, [P, P]),
            "PGM_get_indexer": (None, [P, P, S, I, P, P]), "PGM_calculate": (None, [P, P, P, P, P]),
            "PGM_destroy_model": (None, [P]),
            "PGM_create_deserializer_from_binary_buffer": (P, [P, C.c_char_p, I, I]),
            "PGM_create_deserializer_from_null_terminated_string": (P, [P, S, I]),
            "PGM_deserializer_get_dataset": (P, [P, P]), "PGM_deserializer_parse_to_buffer": (None, [P, P]),
            "PGM_destroy_deserializer": (None, [P]), "PGM_create_serializer": (P, [P, P, I]),
            "PGM_serializer_get_to_binary_buffer": (None, [P, P, I, C.POINTER(C.c_void_p), C.POINTER(I)]),
            "PGM_serializer_get_to_zero_terminated_string": (S, [P, P, I, I]), "PGM_destroy_serializer": (None, [P]),
            "PGM_dataset_writable_get_info": (P, [P, P]), "PGM_dataset_writable_set_buffer": (None, [P, P, S, P, P]),
            "PGM_dataset_writable_set_attribute_buffer": (None, [P, P, S, S, P]),
            "PGM_create_dataset_const_from_writable": (P, [P, P]),
        }
        for name, (res, args) in sig.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _core = l
    return _core


def create_buffer(dataset: str, component: str, shape):
    """numpy view of a buffer allocated with PGM_create_buffer (buffer.h: the library owns the memory until PGM_destroy_buffer).
    On a GPU box this library hands out page-locked memory for large buffers, so PGM_calculate transfers into it directly."""
    import weakref

    c = core()
    h = Handle()
    meta = c.PGM_meta_get_component_by_name(h.h, dataset.encode(), component.encode())
    h.check()
    table = {"input": structs.INPUT, "update": structs.UPDATE, "sym_output": structs.SYM_OUTPUT, "asym_output": structs.ASYM_OUTPUT}[dataset]
    dt = table[component]
    n = int(np.prod(shape))
    ptr = c.PGM_create_buffer(h.h, meta, max(n, 1))
    h.check()
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(n, 1) * dt.itemsize,))[:n * dt.itemsize].view(dt).reshape(shape)
    weakref.finalize(arr.base if arr.base is not None else arr, c.PGM_destroy_buffer, ptr)
    return arr


class Handle:
    def __init__(self):
        self.h = core().PGM_create_handle()

    def __del__(self):
        if _core is not None and self.h:
            _core.PGM_destroy_handle(self.h)
            self.h = None

    def check(self, batch_error_ok=False):
        """assert_no_error of the reference's wrapper (_core/error_handling.py)"""
        c = core()
        code = c.PGM_error_code(self.h)
        if code == PGM_NO_ERROR:
            return None
        msg = c.PGM_error_message(self.h).decode()
        if code == PGM_BATCH_ERROR:
            n = c.PGM_n_failed_scenarios(self.h)
            failed = np.array([c.PGM_failed_scenarios(self.h)[i] for i in range(n)], dtype=np.int64)
            errs = c.PGM_batch_errors(self.h)
            err = PowerGridBatchError(msg, failed, [errs[i].decode() for i in range(n)])
            c.PGM_clear_error(self.h)
            if batch_error_ok:
                return err
            raise err
        c.PGM_clear_error(self.h)
        raise PowerGridError(msg)


_CTYPES = {0: "<i4", 1: "i1", 2: "<f8", 3: ("<f8", (3,))}
_meta_cache = None


def power_grid_meta_data():
    """dataset -> component -> numpy dtype, read through PGM_meta_* the way the reference's wrapper does
    (_core/power_grid_meta.py: names, formats, offsets, itemsize from the library, nothing hard-coded)"""
    global _meta_cache
    if _meta_cache is None:
        c, h = core(), Handle()
        meta = {}
        for d in range(c.PGM_meta_n_datasets(h.h)):
            ds = c.PGM_meta_get_dataset_by_idx(h.h, d)
            comps = {}
            for k in range(c.PGM_meta_n_components(h.h, ds)):
                comp = c.PGM_meta_get_component_by_idx(h.h, ds, k)
                attrs = [c.PGM_meta_get_attribute_by_idx(h.h, comp, a) for a in range(c.PGM_meta_n_attributes(h.h, comp))]
                comps[c.PGM_meta_component_name(h.h, comp).decode()] = np.dtype({
                    "names": [c.PGM_meta_attribute_name(h.h, a).decode() for a in attrs],
                    "formats": [_CTYPES[c.PGM_meta_attribute_ctype(h.h, a)] for a in attrs],
                    "offsets": [c.PGM_meta_attribute_offset(h.h, a) for a in attrs],
                    "itemsize": c.PGM_meta_component_size(h.h, comp), "aligned": True})
            meta[c.PGM_meta_dataset_name(h.h, ds).decode()] = comps
        h.check()
        _meta_cache = meta
    return _meta_cache


def _is_sparse(val):
    return isinstance(val, dict) and set(val) == {"data", "indptr"}


class _Dataset:
    """Const (input / update) or mutable (output) dataset over numpy row buffers; keeps the arrays alive."""

    def __init__(self, handle, name, data, *, mutable, is_batch, batch_size):
        c = core()
        self.handle, self.mutable, self.keep, self.batch_size = handle, mutable, [], batch_size
        create = c.PGM_create_dataset_mutable if mutable else c.PGM_create_dataset_const
        self.ptr = create(handle.h, name.encode(), int(is_batch), int(batch_size))
        handle.check()
        add = c.PGM_dataset_mutable_add_buffer if mutable else c.PGM_dataset_const_add_buffer
        add_attr = c.PGM_dataset_mutable_add_attribute_buffer if mutable else c.PGM_dataset_const_add_attribute_buffer
        for comp, val in data.items():
            indptr = None
            if _is_sparse(val):
                indptr = np.ascontiguousarray(val["indptr"], dtype=np.int64)
                self.keep.append(indptr)
                val = val["data"]
            if isinstance(val, dict):  # columnar: attribute -> array, dense (n_scenarios, n[, 3]) or the flat data of a sparse buffer
                cols = {a: (v if mutable else np.ascontiguousarray(v)) for a, v in val.items()}
                self.keep.append(cols)
                first = next(iter(cols.values()))
                if indptr is not None:
                    per_scenario, total = -1, first.shape[0]
                elif is_batch:
                    per_scenario, total = first.shape[1], first.shape[0] * first.shape[1]
                else:
                    per_scenario, total = first.shape[0], first.shape[0]
                add(handle.h, self.ptr, comp.encode(), per_scenario, total, None if indptr is None else indptr.ctypes.data, None)
                handle.check()
                for a, v in cols.items():
                    assert v.flags.c_contiguous
                    add_attr(handle.h, self.ptr, comp.encode(), a.encode(), v.ctypes.data)
                    handle.check()
                continue
            if indptr is not None:  # sparse row buffer
                arr = np.ascontiguousarray(val)
                self.keep.append(arr)
                add(handle.h, self.ptr, comp.encode(), -1, arr.size, indptr.ctypes.data, arr.ctypes.data)
            else:
                arr = val if mutable else np.ascontiguousarray(val)
                assert arr.flags.c_contiguous
                self.keep.append(arr)
                per_scenario = arr.shape[-1] if arr.ndim else 1
                add(handle.h, self.ptr, comp.encode(), per_scenario, arr.size, None, arr.ctypes.data)
            handle.check()

    def __del__(self):
        if _core is not None and self.ptr:
            (_core.PGM_destroy_dataset_mutable if self.mutable else _core.PGM_destroy_dataset_const)(self.ptr)
            self.ptr = None


class PowerGridModel:
    """The reference's PowerGridModel surface for power flow, over PGM_* symbols."""

    def __init__(self, input_data, system_frequency=50.0, _ptr=None, _counts=None):
        self.handle = Handle()
        if _ptr is not None:
            self.ptr, self._counts = _ptr, _counts
            return
        ds = _Dataset(self.handle, "input", input_data, mutable=False, is_batch=False, batch_size=1)
        self.ptr = core().PGM_create_model(self.handle.h, float(system_frequency), ds.ptr)
        self.handle.check()
        self._counts = {c: len(next(iter(a.values())) if isinstance(a, dict) else a) for c, a in input_data.items()}

    def __del__(self):
        if _core is not None and getattr(self, "ptr", None):
            _core.PGM_destroy_model(self.ptr)
            self.ptr = None

    def copy(self):
        h = Handle()
        ptr = core().PGM_copy_model(h.h, self.ptr)
        h.check()
        return PowerGridModel(None, _ptr=ptr, _counts=dict(self._counts))

    def update(self, *, update_data):
        ds = _Dataset(self.handle, "update", update_data, mutable=False, is_batch=False, batch_size=1)
        core().PGM_update_model(self.handle.h, self.ptr, ds.ptr)
        self.handle.check()

    def get_indexer(self, component_type, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        indexer = np.empty(ids.shape, dtype=np.int64)
        core().PGM_get_indexer(self.handle.h, self.ptr, component_type.encode(), ids.size, ids.ctypes.data, indexer.ctypes.data)
        self.handle.check()
        return indexer

    def calculate_power_flow(self, *, symmetric=True, error_tolerance=1e-8, max_iterations=20,
                             calculation_method="newton_raphson", update_data=None, threading=-1,
                             output_component_types=None, continue_on_batch_error=False, tap_changing_strategy=0,
                             calculation_type=0, output_buffers=None):
        c = core()
        opt = c.PGM_create_options(self.handle.h)
        try:
            method = CALCULATION_METHOD[calculation_method] if isinstance(calculation_method, str) else int(calculation_method)
            c.PGM_set_calculation_type(self.handle.h, opt, int(calculation_type))
            c.PGM_set_calculation_method(self.handle.h, opt, method)
            c.PGM_set_symmetric(self.handle.h, opt, int(bool(symmetric)))
            c.PGM_set_err_tol(self.handle.h, opt, float(error_tolerance))
            c.PGM_set_max_iter(self.handle.h, opt, int(max_iterations))
            c.PGM_set_threading(self.handle.h, opt, int(threading))
            c.PGM_set_tap_changing_strategy(self.handle.h, opt, int(tap_changing_strategy))
            batch, n_scn, upd = update_data is not None, 1, None
            if isinstance(update_data, (list, tuple)):  # cartesian product of batch dimensions, outermost first
                dims = []
                for data in update_data:
                    size = np.asarray(next(iter(next(iter(data.values())).values())) if isinstance(next(iter(data.values())), dict)
                                      and not _is_sparse(next(iter(data.values()))) else next(iter(data.values()))).shape[0]
                    dims.append(_Dataset(self.handle, "update", data, mutable=False, is_batch=True, batch_size=size))
                for outer, inner in zip(dims, dims[1:]):
                    c.PGM_dataset_const_set_next_cartesian_product_dimension(self.handle.h, outer.ptr, inner.ptr)
                    self.handle.check()
                upd, n_scn = dims[0], int(np.prod([d.batch_size for d in dims]))
                self._dims = dims
            elif batch:
                sizes = {(len(v["indptr"]) - 1) if _is_sparse(v) else
                         np.asarray(next(iter(v.values())) if isinstance(v, dict) else v).shape[0] for v in update_data.values()}
                if len(sizes) > 1:
                    raise PowerGridError("inconsistent batch sizes in update data")
                n_scn = sizes.pop() if sizes else 0
                upd = _Dataset(self.handle, "update", update_data, mutable=False, is_batch=True, batch_size=n_scn)
            table = structs.SYM_OUTPUT if symmetric else structs.ASYM_OUTPUT
            comps = output_component_types or [k for k in structs.COMPONENT_ORDER if self._counts.get(k)]
            shape = (lambda k: (n_scn, self._counts.get(k, 0))) if batch else (lambda k: (self._counts.get(k, 0),))
            if isinstance(comps, dict):  # component -> attribute names (columnar output) or None (row output)
                result = {k: (np.zeros(shape(k), dtype=table[k]) if attrs is None else
                              {a: np.zeros(shape(k) + table[k][a].shape, dtype=table[k][a].base) for a in attrs})
                          for k, attrs in comps.items()}
            elif output_buffers is not None:  # caller-owned row buffers (e.g. from create_buffer: page-locked when large)
                result = {k: output_buffers[k] for k in comps}
                assert all(result[k].dtype == table[k] and result[k].shape == shape(k) for k in comps)
            else:
                result = {k: np.zeros(shape(k), dtype=table[k]) for k in comps}
            out = _Dataset(self.handle, "sym_output" if symmetric else "asym_output", result, mutable=True, is_batch=batch,
                           batch_size=n_scn)
            c.PGM_calculate(self.handle.h, self.ptr, opt, out.ptr, upd.ptr if upd is not None else None)
            self.batch_error = self.handle.check(batch_error_ok=continue_on_batch_error)
            return result
        finally:
            c.PGM_destroy_options(opt)


# ---- (de)serialization (serialization.h), the way the reference's wrapper drives it (_core/serialization.py) -----------------
JSON, MSGPACK = 0, 1


def _check_serialization(h):
    c = core()
    code = c.PGM_error_code(h.h)
    if code != PGM_NO_ERROR:
        msg = c.PGM_error_message(h.h).decode()
        c.PGM_clear_error(h.h)
        raise PowerGridError(msg)


def deserialize(data, serialization_format=JSON):
    """bytes / str in the reference's dataset format -> (dataset type, dict component -> structured array, or
    {"data", "indptr"} for a component whose scenarios hold different numbers of elements)"""
    c, h = core(), Handle()
    raw = data.encode() if isinstance(data, str) else bytes(data)
    des = c.PGM_create_deserializer_from_binary_buffer(h.h, raw, len(raw), serialization_format)
    _check_serialization(h)
    try:
        ds = c.PGM_deserializer_get_dataset(h.h, des)
        info = c.PGM_dataset_writable_get_info(h.h, ds)
        name = c.PGM_dataset_info_name(h.h, info).decode()
        is_batch = bool(c.PGM_dataset_info_is_batch(h.h, info))
        batch_size = c.PGM_dataset_info_batch_size(h.h, info)
        table = power_grid_meta_data()[name]  # dtypes of every component of the reference, built from PGM_meta_*
        out, keep = {}, []
        for k in range(c.PGM_dataset_info_n_components(h.h, info)):
            comp = c.PGM_dataset_info_component_name(h.h, info, k).decode()
            per, total = c.PGM_dataset_info_elements_per_scenario(h.h, info, k), c.PGM_dataset_info_total_elements(h.h, info, k)
            arr = np.zeros(total, dtype=table[comp])
            indptr = np.zeros(batch_size + 1, np.int64) if per < 0 else None
            keep += [arr, indptr]
            c.PGM_dataset_writable_set_buffer(h.h, ds, comp.encode(), None if indptr is None else indptr.ctypes.data, arr.ctypes.data)
            h.check()
            if per < 0:
                out[comp] = {"data": arr, "indptr": indptr}
            else:
                out[comp] = arr.reshape(batch_size, per) if is_batch else arr
        c.PGM_deserializer_parse_to_buffer(h.h, des)
        _check_serialization(h)
        return name, out
    finally:
        c.PGM_destroy_deserializer(des)


def serialize(dataset_type, data, serialization_format=JSON, use_compact_list=False, indent=2, is_batch=None):
    """dict component -> array (2-D for a batch) or {"data", "indptr"} -> str (JSON) / bytes (msgpack)"""
    c, h = core(), Handle()
    first = next(iter(data.values()))
    if is_batch is None:
        is_batch = isinstance(first, dict) or np.asarray(first).ndim == 2
    batch_size = 1
    if is_batch:
        batch_size = (len(first["indptr"]) - 1) if isinstance(first, dict) else np.asarray(first).shape[0]
    ds = _Dataset(h, dataset_type, data, mutable=False, is_batch=is_batch, batch_size=batch_size)
    ser = c.PGM_create_serializer(h.h, ds.ptr, serialization_format)
    _check_serialization(h)
    try:
        if serialization_format == JSON:
            text = c.PGM_serializer_get_to_zero_terminated_string(h.h, ser, int(use_compact_list), indent)
            _check_serialization(h)
            return text.decode()
        ptr, size = C.c_void_p(), I64(0)
        c.PGM_serializer_get_to_binary_buffer(h.h, ser, int(use_compact_list), C.byref(ptr), C.byref(size))
        _check_serialization(h)
        return C.string_at(ptr, size.value)
    finally:
        c.PGM_destroy_serializer(ser)


I64 = C.c_int64
