"""Loader of libpgm_b200.so (the C-ABI declared in include/pgm_b200.h).  There is no fallback: a missing library is an
ImportError-like RuntimeError with the build hint, a missing GPU surfaces as PGMB_ERR_CUDA from the calls."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGMB_LIB", os.path.join(_HERE, "libpgm_b200.so"))  # PGMB_LIB: experiment builds

PGMB_OK, PGMB_ERR_INVALID, PGMB_ERR_CUDA, PGMB_ERR_BATCH, PGMB_ERR_INTERNAL = range(5)
METHODS = {"default_method": -128, "linear": 0, "newton_raphson": 1, "iterative_current": 3, "linear_current": 4}
TAP_STRATEGIES = {"disabled": 0, "any_valid_tap": 1, "min_voltage_tap": 2, "max_voltage_tap": 3, "fast_any_tap": 4}


class PgmB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


class BatchError(PgmB200Error):
    """Analogue of PowerGridBatchError: some scenarios failed, the others are valid."""

    def __init__(self, message, status):
        super().__init__(PGMB_ERR_BATCH, message)
        self.status = status


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C power-grid-model_b200/csrc). pgm_b200 has no CPU fallback."
            )
        l = C.CDLL(LIB_PATH)
        l.pgmb_last_error.restype = C.c_char_p
        l.pgmb_version.restype = C.c_char_p
        l.pgmb_kernel_launch_count.restype = C.c_uint64
        l.pgmb_host_alloc.argtypes = [C.c_uint64, C.POINTER(C.c_void_p)]
        l.pgmb_host_free.argtypes = [C.c_void_p]
        l.pgmb_model_n_math_groups.restype = C.c_int64
        _lib = l
    return _lib


def check(rc, allow_batch=False):
    if rc == PGMB_OK or (allow_batch and rc == PGMB_ERR_BATCH):
        return rc
    raise PgmB200Error(rc, lib().pgmb_last_error().decode())


class MathTopologyC(C.Structure):
    _fields_ = [
        ("n_bus", C.c_int64), ("phase_shift", C.c_void_p), ("n_branch", C.c_int64), ("branch_bus_idx", C.c_void_p),
        ("n_fill_in", C.c_int64), ("fill_in", C.c_void_p), ("sources_per_bus", C.c_void_p),
        ("shunts_per_bus", C.c_void_p), ("load_gens_per_bus", C.c_void_p), ("load_gen_type", C.c_void_p),
        ("voltage_regulators_per_load_gen", C.c_void_p),
    ]


class MathParamC(C.Structure):
    _fields_ = [("branch_param", C.c_void_p), ("shunt_param", C.c_void_p), ("source_param", C.c_void_p)]


class RunOptionsC(C.Structure):
    _fields_ = [("method", C.c_int32), ("err_tol", C.c_double), ("max_iter", C.c_int64)]


class PfInputC(C.Structure):
    _fields_ = [("n_scenarios", C.c_int64), ("source_u_ref", C.c_void_p), ("source_is_shared", C.c_int32),
                ("s_injection", C.c_void_p), ("voltage_regulator", C.c_void_p), ("load_gen_status", C.c_void_p),
                ("method_hint", C.c_int32), ("method_hint_valid", C.c_int32)]


class SolverOutputC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("u", "bus_injection", "branch", "source", "shunt", "load_gen", "status", "n_iter", "max_dev",
                                               "voltage_regulator")]


class ComponentBufferC(C.Structure):
    _fields_ = [("n", C.c_int64), ("indptr", C.c_void_p), ("data", C.c_void_p)]


_COMPS = ("node", "line", "transformer", "shunt", "source", "sym_gen", "asym_gen", "sym_load", "asym_load", "voltage_regulator",
          "asym_line", "generic_branch", "link", "three_winding_transformer", "transformer_tap_regulator")


class InputDataC(C.Structure):
    _fields_ = [(c, ComponentBufferC) for c in _COMPS]


class UpdateDataC(C.Structure):
    _fields_ = [("n_scenarios", C.c_int64)] + [(c, ComponentBufferC) for c in _COMPS[1:]]


class OutputDataC(C.Structure):
    _fields_ = [(c, C.c_void_p) for c in _COMPS]


class OptionsC(C.Structure):
    _fields_ = [("calculation_method", C.c_int32), ("symmetric", C.c_int32), ("err_tol", C.c_double),
                ("max_iter", C.c_int64), ("n_devices", C.c_int32), ("first_device", C.c_int32), ("threading", C.c_int32),
                ("flags", C.c_uint32), ("tap_changing_strategy", C.c_int32)]


FLAG_RESIDENT_INPUT = 1
FLAG_RESIDENT_OUTPUT = 2


class GridOptionC(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_node_total_specified", "n_mv_feeder", "n_node_per_mv_feeder", "n_lv_feeder",
                                         "n_connection_per_lv_feeder")] + [("has_mv_ring", C.c_int32), ("has_lv_ring", C.c_int32)]


class _PinnedBlock:
    """owner of one cudaHostAlloc block, exposed through the array interface: numpy views keep it alive, the block is
    released when the last view is garbage-collected"""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        check(lib().pgmb_host_alloc(C.c_uint64(nbytes), C.byref(self.ptr)))
        self.__array_interface__ = {"data": (self.ptr.value, False), "shape": (nbytes,), "typestr": "|u1", "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                lib().pgmb_host_free(self.ptr)
        except Exception:  # interpreter shutdown
            pass


def pinned_empty(shape, dtype):
    """numpy array in page-locked host memory (pgmb_host_alloc): update / output buffers of this kind let the batch
    calculation overlap its PCIe transfers with the solver kernels"""
    import numpy as np

    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    if n == 0:
        return np.empty(shape, dtype)
    return np.asarray(_PinnedBlock(n)).view(dtype).reshape(shape)
