"""ctypes wrapper over oracle/liboracle.so (the CPU restatement; test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

import pgm_b200
from pgm_b200 import structs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "liboracle.so")

STATUS_OK, STATUS_DIVERGED, STATUS_SINGULAR, STATUS_ERROR = 0, 1, 2, 3
METHODS = {"linear": 0, "newton_raphson": 1, "iterative_current": 3, "linear_current": 4, "default": -128}


def _load():
    if not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(_SO)
    lib.orc_bag_new.restype = C.c_void_p
    lib.orc_bag_free.argtypes = [C.c_void_p]
    lib.orc_bag_error.restype = C.c_char_p
    lib.orc_bag_error.argtypes = [C.c_void_p]
    lib.orc_model_create.restype = C.c_void_p
    lib.orc_model_destroy.argtypes = [C.c_void_p]
    lib.orc_model_calculate.restype = C.c_int64
    lib.orc_hardware_concurrency.restype = C.c_int64
    return lib


lib = _load()

_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


def _p(arr, ctype=None):
    if arr is None:
        return None
    return arr.ctypes.data_as(C.c_void_p)


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64).reshape(-1))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


def _i8(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int8).reshape(-1))


def _cplx(a):
    """complex array -> interleaved float64 view (copy)"""
    a = np.ascontiguousarray(np.asarray(a, dtype=np.complex128))
    return a.view(np.float64).reshape(-1).copy()


class Bag:
    def __init__(self):
        self.h = C.c_void_p(lib.orc_bag_new())

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_bag_free(self.h)
            self.h = None

    def i64(self, name):
        ptr, n = _i64p(), C.c_int64()
        if lib.orc_bag_get_i64(self.h, name.encode(), C.byref(ptr), C.byref(n)) != 0:
            raise KeyError(name)
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int64)

    def f64(self, name):
        ptr, n = _f64p(), C.c_int64()
        if lib.orc_bag_get_f64(self.h, name.encode(), C.byref(ptr), C.byref(n)) != 0:
            raise KeyError(name)
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float64)

    def c128(self, name):
        return self.f64(name).view(np.complex128)

    @property
    def error(self):
        return lib.orc_bag_error(self.h).decode()


def sparse_lu_solve(n_block, indptr, indices, diag, data, rhs, use_pivot_perturbation=False, prefactorize_separately=False):
    """data: (nnz, N, N) row/col matrices (real or complex); rhs: (n, N). Returns (status, x, lu_data, bag)."""
    data = np.asarray(data)
    is_complex = np.iscomplexobj(data) or np.iscomplexobj(rhs)
    dt = np.complex128 if is_complex else np.float64
    n = len(indptr) - 1
    N = n_block
    d = np.asarray(data, dtype=dt).reshape(-1, N, N)
    d_cm = np.array(np.transpose(d, (0, 2, 1)), order="C", copy=True)  # column-major blocks
    r = np.array(np.asarray(rhs, dtype=dt).reshape(n, N), order="C", copy=True)
    x = np.zeros_like(r)
    bag = Bag()
    st = lib.orc_sparse_lu_solve(
        bag.h, C.c_int(N), C.c_int(int(is_complex)), C.c_int64(n), _p(_i64(indptr)), _p(_i64(indices)), _p(_i64(diag)),
        _p(d_cm), _p(r), _p(x), C.c_int(int(use_pivot_perturbation)), C.c_int(int(prefactorize_separately)),
    )
    return st, x, np.transpose(d_cm, (0, 2, 1)), bag


def min_degree(graph: dict):
    keys = sorted(graph)
    indptr = np.cumsum([0] + [len(graph[k]) for k in keys])
    adj = [v for k in keys for v in graph[k]]
    bag = Bag()
    st = lib.orc_min_degree(bag.h, C.c_int64(len(keys)), _p(_i64(keys)), _p(_i64(indptr)), _p(_i64(adj)))
    assert st == 0, bag.error
    return bag.i64("alpha").tolist(), [tuple(x) for x in bag.i64("fills").reshape(-1, 2).tolist()]


def ybus_structure(n_bus, branch_bus_idx, fill_in, shunts_per_bus_indptr):
    bb = _i64(branch_bus_idx)
    fi = _i64(fill_in)
    bag = Bag()
    st = lib.orc_ybus_structure(
        bag.h, C.c_int64(n_bus), C.c_int64(len(bb) // 2), _p(bb), C.c_int64(len(fi) // 2), _p(fi), _p(_i64(shunts_per_bus_indptr))
    )
    assert st == 0, bag.error
    return bag


def topology(n_node, branch_node_idx, branch_connected, branch_phase_shift, source_node_idx, source_connected,
             shunt_node_idx=(), load_gen_node_idx=(), load_gen_type=(), branch3_node_idx=(), branch3_connected=(),
             branch3_phase_shift=()):
    bn = _i64(branch_node_idx)
    b3 = _i64(branch3_node_idx)
    bag = Bag()
    st = lib.orc_topology(
        bag.h, C.c_int64(n_node), C.c_int64(len(bn) // 2), _p(bn), _p(_i8(branch_connected)), _p(_f64(branch_phase_shift)),
        C.c_int64(len(b3) // 3), _p(b3), _p(_i8(branch3_connected)), _p(_f64(branch3_phase_shift)),
        C.c_int64(len(source_node_idx)), _p(_i64(source_node_idx)), _p(_i8(source_connected)),
        C.c_int64(len(shunt_node_idx)), _p(_i64(shunt_node_idx)),
        C.c_int64(len(load_gen_node_idx)), _p(_i64(load_gen_node_idx)), _p(_i8(load_gen_type)),
    )
    assert st == 0, bag.error
    return bag


class MathGrid:
    """One math sub-grid as flat arrays (the MathModelTopology / MathModelParam / PowerFlowInput seam)."""

    def __init__(self, sym, phase_shift, branch_bus_idx, sources_per_bus, shunts_per_bus, load_gens_per_bus,
                 load_gen_type, branch_param, shunt_param, source_param, source_u_ref, s_injection, fill_in=()):
        self.sym = bool(sym)
        self.B = 1 if sym else 3
        B = self.B
        self.phase_shift = _f64(phase_shift)
        self.n_bus = len(self.phase_shift)
        self.branch_bus_idx = _i64(branch_bus_idx).reshape(-1, 2)
        self.fill_in = _i64(fill_in).reshape(-1, 2)
        self.sources_per_bus = _i64(sources_per_bus)
        self.shunts_per_bus = _i64(shunts_per_bus)
        self.load_gens_per_bus = _i64(load_gens_per_bus)
        self.load_gen_type = _i8(load_gen_type)
        self.branch_param = np.asarray(branch_param, dtype=np.complex128).reshape(-1, 4, B, B)
        self.shunt_param = np.asarray(shunt_param, dtype=np.complex128).reshape(-1, B, B)
        self.source_param = np.asarray(source_param, dtype=np.complex128).reshape(-1, 2)  # y1, y0
        self.source_u_ref = np.asarray(source_u_ref, dtype=np.complex128).reshape(-1)
        self.s_injection = np.asarray(s_injection, dtype=np.complex128).reshape(-1, B)


def math_pf(grid: MathGrid, method="newton_raphson", err_tol=1e-8, max_iter=20):
    bag = Bag()
    st = lib.orc_math_pf(
        bag.h, C.c_int(int(grid.sym)), C.c_int64(grid.n_bus), _p(grid.phase_shift), C.c_int64(len(grid.branch_bus_idx)),
        _p(_i64(grid.branch_bus_idx)), C.c_int64(len(grid.fill_in)), _p(_i64(grid.fill_in)), _p(grid.sources_per_bus),
        _p(grid.shunts_per_bus), _p(grid.load_gens_per_bus), _p(grid.load_gen_type), _p(_cplx(grid.branch_param)),
        _p(_cplx(grid.shunt_param)), _p(_cplx(grid.source_param)), _p(_cplx(grid.source_u_ref)),
        _p(_cplx(grid.s_injection)), C.c_int(METHODS[method]), C.c_double(err_tol), C.c_int64(max_iter),
    )
    out = {"status": st, "error": bag.error, "bag": bag}
    if st == 0:
        B = grid.B
        out["u"] = bag.c128("u").reshape(-1, B)
        out["bus_injection"] = bag.c128("bus_injection").reshape(-1, B)
        out["branch"] = bag.c128("branch").reshape(-1, 4, B)  # s_f, s_t, i_f, i_t
        out["source"] = bag.c128("source").reshape(-1, 2, B)  # s, i
        out["shunt"] = bag.c128("shunt").reshape(-1, 2, B)
        out["load_gen"] = bag.c128("load_gen").reshape(-1, 2, B)
        out["num_iter"] = int(bag.i64("num_iter")[0])
    return out


# ---- component-level model -------------------------------------------------------------------------------------
ORACLE_COMPONENT_ORDER = structs.COMPONENT_ORDER
ORACLE_UPDATABLE = structs.UPDATABLE


class _ModelInput(C.Structure):
    _fields_ = [(f, t) for c in ORACLE_COMPONENT_ORDER for f, t in ((f"n_{c}", C.c_int64), (c, C.c_void_p))]


class _UpdateBuffer(C.Structure):
    _fields_ = [("n_per_scenario", C.c_int64), ("indptr", C.c_void_p), ("data", C.c_void_p)]


class _BatchUpdate(C.Structure):
    _fields_ = [("n_scenarios", C.c_int64)] + [(c, _UpdateBuffer) for c in ORACLE_UPDATABLE]


class _BatchOutput(C.Structure):
    _fields_ = [(c, C.c_void_p) for c in ORACLE_COMPONENT_ORDER]


class Model:
    """Component-level oracle model. `input_data`: dict component -> structured array (structs.INPUT dtypes)."""

    def __init__(self, input_data: dict, system_frequency=50.0):
        self._keep = {}
        mi = _ModelInput()
        self.counts = {}
        for c in ORACLE_COMPONENT_ORDER:
            arr = input_data.get(c)
            n = 0 if arr is None else len(arr)
            self.counts[c] = n
            setattr(mi, f"n_{c}", n)
            if n:
                arr = np.ascontiguousarray(arr, dtype=structs.INPUT[c])
                self._keep[c] = arr
                setattr(mi, c, arr.ctypes.data)
        bag = Bag()
        self.h = lib.orc_model_create(bag.h, C.c_double(system_frequency), C.byref(mi))
        if not self.h:
            raise RuntimeError(bag.error)
        self.h = C.c_void_p(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_model_destroy(self.h)
            self.h = None

    def export_math(self, sym=True):
        bag = Bag()
        st = lib.orc_model_export_math(bag.h, self.h, C.c_int(int(sym)))
        assert st == 0, bag.error
        return bag

    def tap_rank(self):
        """rows (kind, index, rank group) of the regulated transformers in the order the tap changer visits them"""
        bag = Bag()
        st = lib.orc_model_tap_rank(bag.h, self.h)
        if st != 0:
            raise RuntimeError(bag.error)
        return bag.i64("tap_rank").reshape(-1, 3)

    def calculate(self, sym=True, method="newton_raphson", err_tol=1e-8, max_iter=20, threading=-1, update=None,
                  output_components=None, reuse_ic_factorization=False, out=None, tap_changing_strategy=0):
        """update: None or dict component -> array of shape (n_scn, n_per) or {"data": flat, "indptr": ...}.
        Returns dict with per-component output arrays (n_scn, n_comp), n_iter, status, error.
        out: optional dict of preallocated output arrays to write into (the caller reuses them between calls, as a client of
        the reference does with its output dataset)."""
        n_scn = 1
        bu = None
        keep = []
        if update is not None:
            bu = _BatchUpdate()
            n_scn = None
            for c, val in update.items():
                if isinstance(val, dict):
                    data = np.ascontiguousarray(val["data"], dtype=structs.UPDATE[c])
                    indptr = _i64(val["indptr"])
                    keep += [data, indptr]
                    ns = len(indptr) - 1
                    buf = _UpdateBuffer(-1, indptr.ctypes.data, data.ctypes.data)
                else:
                    data = np.ascontiguousarray(val, dtype=structs.UPDATE[c])
                    assert data.ndim == 2
                    keep.append(data)
                    ns = data.shape[0]
                    buf = _UpdateBuffer(data.shape[1], None, data.ctypes.data if data.size else None)
                assert n_scn in (None, ns)
                n_scn = ns
                setattr(bu, c, buf)
            bu.n_scenarios = n_scn
        table = structs.SYM_OUTPUT if sym else structs.ASYM_OUTPUT
        comps = output_components if output_components is not None else [c for c in ORACLE_COMPONENT_ORDER if self.counts[c]]
        bo = _BatchOutput()
        result = {}
        for c in comps:
            if out is not None and c in out:
                arr = out[c]
                assert arr.dtype == table[c] and arr.shape == (n_scn, self.counts[c]) and arr.flags.c_contiguous
            else:
                arr = np.zeros((n_scn, self.counts[c]), dtype=table[c])
            result[c] = arr
            if arr.size:
                setattr(bo, c, arr.ctypes.data)
        n_iter = np.zeros(n_scn, dtype=np.int64)
        status = np.zeros(n_scn, dtype=np.int32)
        bag = Bag()
        failed = lib.orc_model_calculate(
            bag.h, self.h, C.c_int(int(sym)), C.c_int(METHODS[method]), C.c_double(err_tol), C.c_int64(max_iter),
            C.c_int64(threading), C.c_int(int(reuse_ic_factorization)), C.byref(bu) if bu is not None else None,
            C.byref(bo), _p(n_iter), _p(status), C.c_int({"disabled": 0, "any_valid_tap": 1, "min_voltage_tap": 2, "max_voltage_tap": 3,
                                                          "fast_any_tap": 4}.get(tap_changing_strategy, tap_changing_strategy)),
        )
        result["n_iter"] = n_iter
        result["status"] = status
        result["error"] = bag.error
        result["n_failed"] = failed
        return result
