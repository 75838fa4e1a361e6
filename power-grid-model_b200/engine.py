"""Engine-level host API (math-solver seam): one math sub-grid, many scenarios, on one GPU.

Mirrors how the reference's math-solver tests drive a solver (tests/cpp_unit_tests/math_solver/test_math_solver_pf.hpp):
build topology + parameters, then run power flow on PowerFlowInput -- here for a whole batch in one call."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check


def _arr(a, dtype):
    return np.ascontiguousarray(np.asarray(a, dtype=dtype).reshape(-1))


def _cplx_flat(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.complex128)).view(np.float64).reshape(-1).copy()


class Engine:
    def __init__(self, *, symmetric, phase_shift, branch_bus_idx, sources_per_bus, shunts_per_bus, load_gens_per_bus,
                 load_gen_type, fill_in=(), device=0, voltage_regulators_per_load_gen=None):
        self.symmetric = bool(symmetric)
        self.B = 1 if symmetric else 3
        self._keep = [
            _arr(phase_shift, np.float64), _arr(branch_bus_idx, np.int64), _arr(fill_in, np.int64),
            _arr(sources_per_bus, np.int64), _arr(shunts_per_bus, np.int64), _arr(load_gens_per_bus, np.int64),
            _arr(load_gen_type, np.int8),
        ]
        ps, bb, fi, spb, shb, lgb, lgt = self._keep
        self.n_bus = len(ps)
        self.n_branch = len(bb) // 2
        self.n_source = int(spb[-1])
        self.n_shunt = int(shb[-1])
        self.n_load_gen = int(lgb[-1])
        self.n_regulator = 0
        vr_ptr = None
        if voltage_regulators_per_load_gen is not None and len(voltage_regulators_per_load_gen):
            vr = _arr(voltage_regulators_per_load_gen, np.int64)
            assert len(vr) == self.n_load_gen + 1
            self._keep.append(vr)
            self.n_regulator = int(vr[-1])
            vr_ptr = vr.ctypes.data
        topo = _lib.MathTopologyC(self.n_bus, ps.ctypes.data, self.n_branch, bb.ctypes.data, len(fi) // 2, fi.ctypes.data,
                                  spb.ctypes.data, shb.ctypes.data, lgb.ctypes.data, lgt.ctypes.data, vr_ptr)
        self._h = C.c_void_p()
        check(lib().pgmb_engine_create(C.byref(topo), C.c_int32(int(symmetric)), C.c_int32(device), C.byref(self._h)))

    @classmethod
    def from_grid(cls, grid, device=0):
        """grid: object with the attributes of tests/oracle_lib.MathGrid"""
        eng = cls(symmetric=grid.sym, phase_shift=grid.phase_shift, branch_bus_idx=grid.branch_bus_idx,
                  sources_per_bus=grid.sources_per_bus, shunts_per_bus=grid.shunts_per_bus,
                  load_gens_per_bus=grid.load_gens_per_bus, load_gen_type=grid.load_gen_type, fill_in=grid.fill_in,
                  device=device, voltage_regulators_per_load_gen=getattr(grid, "voltage_regulators_per_load_gen", None))
        eng.set_param(grid.branch_param, grid.shunt_param, grid.source_param)
        return eng

    def close(self):
        if getattr(self, "_h", None):
            lib().pgmb_engine_destroy(self._h)
            self._h = None

    __del__ = close

    def set_param(self, branch_param, shunt_param, source_param):
        bp, sp, so = _cplx_flat(branch_param), _cplx_flat(shunt_param), _cplx_flat(source_param)
        assert bp.size == self.n_branch * 4 * self.B * self.B * 2 and so.size == self.n_source * 4
        p = _lib.MathParamC(bp.ctypes.data, sp.ctypes.data, so.ctypes.data)
        check(lib().pgmb_engine_set_param(self._h, C.byref(p)))

    def index(self, name):
        ptr, n = C.POINTER(C.c_int64)(), C.c_int64()
        check(lib().pgmb_engine_get_index(self._h, name.encode(), C.byref(ptr), C.byref(n)))
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int64)

    def admittance(self):
        ptr, n = C.POINTER(C.c_double)(), C.c_int64()
        check(lib().pgmb_engine_get_admittance(self._h, C.byref(ptr), C.byref(n)))
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy().view(np.complex128).reshape(-1, self.B, self.B)

    # -- running ---------------------------------------------------------------------------------------------------
    def _input(self, s_injection, source_u_ref, voltage_regulator=None, load_gen_status=None, method=None):
        s = np.ascontiguousarray(np.asarray(s_injection, dtype=np.complex128))
        if s.ndim == 2 and self.B == 1:
            s = s.reshape(s.shape[0], s.shape[1], 1)
        n_scn = s.shape[0]
        assert s.shape == (n_scn, self.n_load_gen, self.B), s.shape
        u = np.ascontiguousarray(np.asarray(source_u_ref, dtype=np.complex128))
        shared = u.ndim == 1
        assert u.shape[-1] == self.n_source
        vr = ls = None
        if self.n_regulator:
            # (n_regulator, 4): status, u_ref, q_min, q_max per unit -- shared by the scenarios of the call
            vr = np.ascontiguousarray(np.asarray(voltage_regulator, dtype=np.float64).reshape(self.n_regulator, 4))
            if load_gen_status is not None:
                ls = np.ascontiguousarray(np.asarray(load_gen_status, dtype=np.int8).reshape(n_scn, self.n_load_gen))
        self._in_keep = (s, u, vr, ls)
        return _lib.PfInputC(n_scn, u.ctypes.data, int(shared), s.ctypes.data, vr.ctypes.data if vr is not None else None,
                             ls.ctypes.data if ls is not None else None, _lib.METHODS[method] if method else 0,
                             1 if method else 0), n_scn

    def _output(self, n_scn, full=True):
        B = self.B
        o = {
            "u": np.zeros((n_scn, self.n_bus, B), np.complex128),
            "status": np.zeros(n_scn, np.int32), "n_iter": np.zeros(n_scn, np.int32), "max_dev": np.zeros(n_scn, np.float64),
        }
        if full:
            o.update({
                "bus_injection": np.zeros((n_scn, self.n_bus, B), np.complex128),
                "branch": np.zeros((n_scn, self.n_branch, 4, B), np.complex128),
                "source": np.zeros((n_scn, self.n_source, 2, B), np.complex128),
                "shunt": np.zeros((n_scn, self.n_shunt, 2, B), np.complex128),
                "load_gen": np.zeros((n_scn, self.n_load_gen, 2, B), np.complex128),
            })
        if self.n_regulator:
            o["voltage_regulator"] = np.zeros((n_scn, self.n_regulator, 2), np.int8)  # limit_violated, generator_status
        c = _lib.SolverOutputC(*[o[k].ctypes.data if k in o and o[k].size else None for k in
                                 ("u", "bus_injection", "branch", "source", "shunt", "load_gen", "status", "n_iter", "max_dev",
                                  "voltage_regulator")])
        return o, c

    def run(self, s_injection, source_u_ref, method="newton_raphson", err_tol=1e-8, max_iter=20, full_output=True,
            voltage_regulator=None, load_gen_status=None):
        """s_injection: (n_scn, n_load_gen[, B]) complex; source_u_ref: (n_source,) shared or (n_scn, n_source).
        Returns dict of arrays; failed scenarios are flagged in out['status'] (no exception)."""
        inp, n_scn = self._input(s_injection, source_u_ref, voltage_regulator, load_gen_status, method)
        out, outc = self._output(n_scn, full_output)
        opt = _lib.RunOptionsC(_lib.METHODS[method], err_tol, max_iter)
        check(lib().pgmb_engine_run(self._h, C.byref(opt), C.byref(inp), C.byref(outc)), allow_batch=True)
        return out

    def stage(self, s_injection, source_u_ref, voltage_regulator=None, load_gen_status=None, method=None):
        """method: the method the staged batch will be solved with (optional hint for the tile width)"""
        inp, n_scn = self._input(s_injection, source_u_ref, voltage_regulator, load_gen_status, method)
        check(lib().pgmb_engine_stage(self._h, C.byref(inp)))
        self._n_staged = n_scn

    def solve_staged(self, method="newton_raphson", err_tol=1e-8, max_iter=20):
        """Runs the solver kernels on the staged batch; returns kernel milliseconds (CUDA events on the engine stream)."""
        opt = _lib.RunOptionsC(_lib.METHODS[method], err_tol, max_iter)
        ms = C.c_float()
        check(lib().pgmb_engine_solve_staged(self._h, C.byref(opt), C.byref(ms)))
        return ms.value

    def fetch(self, full_output=False):
        out, outc = self._output(self._n_staged, full_output)
        check(lib().pgmb_engine_fetch(self._h, C.byref(outc)))
        return out


class SparseLU:
    """Batched block-sparse LU solve on the GPU with the reference's pivot perturbation + iterative refinement
    (math_solver/sparse_lu_solver.hpp: SparseLUSolver; include/pgm_b200.h: pgmb_sparse_lu_*).  One pattern, many systems."""

    def __init__(self, row_indptr, col_indices, diag_lu, block_size=1, is_complex=False, device=0):
        self._ip = np.ascontiguousarray(row_indptr, dtype=np.int64)
        self._ix = np.ascontiguousarray(col_indices, dtype=np.int64)
        self._dg = np.ascontiguousarray(diag_lu, dtype=np.int64)
        self.n, self.nnz, self.block_size, self.is_complex = len(self._ip) - 1, int(self._ip[-1]), int(block_size), bool(is_complex)
        self._h = C.c_void_p()
        check(lib().pgmb_sparse_lu_create(C.c_int64(self.n), self._ip.ctypes.data_as(C.c_void_p), self._ix.ctypes.data_as(C.c_void_p),
                                          self._dg.ctypes.data_as(C.c_void_p), C.c_int32(self.block_size), C.c_int32(int(self.is_complex)),
                                          C.c_int32(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().pgmb_sparse_lu_destroy(self._h)
            self._h = None

    __del__ = close

    def solve(self, data, rhs, use_pivot_perturbation=False):
        """data (n_batch, nnz, N, N) column-major blocks (i.e. data[b, k, c, r]) or (n_batch, nnz) for N = 1; rhs (n_batch, n, N).
        Returns dict: x, status (0 ok / 2 singular), perturbed, n_solves, lu, perm (n_batch, n, 2, N)."""
        dt = np.complex128 if self.is_complex else np.float64
        N = self.block_size
        data = np.ascontiguousarray(data, dtype=dt).reshape(-1, self.nnz, N, N)
        nb = data.shape[0]
        rhs = np.ascontiguousarray(rhs, dtype=dt).reshape(nb, self.n, N)
        out = {"x": np.zeros_like(rhs), "status": np.zeros(nb, np.int32), "perturbed": np.zeros(nb, np.int32),
               "n_solves": np.zeros(nb, np.int32), "lu": np.zeros_like(data), "perm": np.zeros((nb, self.n, 2, N), np.int8)}
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = lib().pgmb_sparse_lu_solve(self._h, C.c_int64(nb), p(data), p(rhs), C.c_int32(int(use_pivot_perturbation)), p(out["x"]),
                                        p(out["status"]), p(out["perturbed"]), p(out["n_solves"]), p(out["lu"]), p(out["perm"]))
        if rc != _lib.PGMB_ERR_BATCH:
            check(rc)
        return out
