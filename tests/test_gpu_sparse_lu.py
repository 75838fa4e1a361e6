"""SURVEY section 8 row a16 on the device: the batched sparse LU solver with pivot perturbation + iterative refinement
(pgmb_sparse_lu_*, csrc/sparse_lu.cu) against the reference's known-answer systems
(tests/cpp_unit_tests/math_solver/test_sparse_lu_solver.cpp:69-251 and the ill-conditioned cases :468-545) and against the
oracle (oracle/sparse_lu.hpp) on random systems of every scalar / block type the power-flow path uses."""
import numpy as np
import pytest

import oracle_lib as orc
import pgm_b200
from pgm_b200.engine import SparseLU

pytestmark = pytest.mark.gpu

ROW_INDPTR = [0, 3, 6, 9]
COL_INDICES = [0, 1, 2, 0, 1, 2, 0, 1, 2]
DIAG_LU = [0, 4, 8]


def _cm(data, N):  # (nnz, N, N) row/col matrices -> column-major blocks
    return np.ascontiguousarray(np.transpose(np.asarray(data).reshape(-1, N, N), (0, 2, 1)))


def test_scalar_3x3_with_fill_ins_known_answer():
    data = np.array([4, 1, 5, 3, 7, 0, 2, 0, 6], dtype=float)  # test_sparse_lu_solver.cpp:69-100
    out = SparseLU(ROW_INDPTR, COL_INDICES, DIAG_LU).solve(data[None], np.array([[21.0, 2.0, 18.0]]))
    assert out["status"][0] == 0 and out["n_solves"][0] == 1 and out["perturbed"][0] == 0
    np.testing.assert_allclose(out["x"].ravel(), [3, -1, 2], atol=1e-12)


@pytest.mark.parametrize("block", [False, True])
def test_ill_conditioned_system_needs_pivot_perturbation(block):
    """test_sparse_lu_solver.cpp:468-545: SparseMatrixError without perturbation, x = {8, 0, 10, 0} with it"""
    if not block:
        indptr, indices, diag, N = [0, 4, 8, 12, 16], list(range(4)) * 4, [0, 5, 10, 15], 1
        data = np.array([0, 0, 0, -1, 0, -1, 0, 0, 0, 0, 5, 1, -1, 0, 1, -9], dtype=float)
        rhs = np.array([0, 0, 50, 2], dtype=float)
    else:
        indptr, indices, diag, N = [0, 2, 4], [0, 1, 0, 1], [0, 3], 2
        data = np.array([[[0, 0], [0, -1]], [[0, -1], [0, 0]], [[0, 0], [-1, 0]], [[5, 1], [1, -9]]], dtype=float)
        rhs = np.array([[0, 0], [50, 2]], dtype=float)
    solver = SparseLU(indptr, indices, diag, block_size=N)
    dev = _cm(data, N)[None]
    out = solver.solve(dev, rhs[None], use_pivot_perturbation=False)
    assert out["status"][0] == 2  # PGMB_SCN_SINGULAR == SparseMatrixError
    out = solver.solve(dev, rhs[None], use_pivot_perturbation=True)
    assert out["status"][0] == 0 and out["perturbed"][0] == 1 and 2 <= out["n_solves"][0] <= 6
    np.testing.assert_allclose(out["x"].ravel(), [8, 0, 10, 0], atol=1e-8)
    st, x, *_ = orc.sparse_lu_solve(N, indptr, indices, diag, data, rhs, use_pivot_perturbation=True, prefactorize_separately=True)
    assert st == 0
    np.testing.assert_allclose(out["x"].ravel(), x.ravel(), atol=1e-12)


@pytest.mark.parametrize("N,cplx", [(1, False), (2, False), (3, False), (6, False), (1, True), (3, True)])
@pytest.mark.parametrize("use_pp", [False, True])
def test_random_batches_against_the_oracle(N, cplx, use_pp):
    """a batch of random systems on a sparse pattern with fill-ins: solution, factors and block permutations equal the oracle's"""
    rng = np.random.default_rng(100 * N + cplx)
    n = 9
    # arrow + band pattern, closed under fill-in when eliminated in order (last row / column dense)
    rows = [sorted({i, max(i - 1, 0), min(i + 1, n - 1), n - 1}) for i in range(n - 1)] + [list(range(n))]
    indptr = np.cumsum([0] + [len(r) for r in rows])
    indices = [c for r in rows for c in r]
    diag = [int(indptr[i]) + rows[i].index(i) for i in range(n)]
    nnz, n_batch = len(indices), 37
    data = rng.normal(size=(n_batch, nnz, N, N))
    rhs = rng.normal(size=(n_batch, n, N))
    if cplx:
        data = data + 1j * rng.normal(size=data.shape)
        rhs = rhs + 1j * rng.normal(size=rhs.shape)
    for i in range(n):
        data[:, diag[i]] += 6 * np.fliplr(np.eye(N))  # regular, with off-diagonal pivots inside the blocks
    if use_pp:  # a few systems get an exactly singular first pivot block: solvable only with the perturbation
        data[::5, diag[0]] = 0.0
    out = SparseLU(indptr, indices, diag, block_size=N, is_complex=cplx).solve(
        np.ascontiguousarray(np.swapaxes(data, 2, 3)), rhs, use_pivot_perturbation=use_pp)
    for b in range(n_batch):
        st, x, lu, bag = orc.sparse_lu_solve(N, indptr, indices, diag, data[b], rhs[b], use_pivot_perturbation=use_pp,
                                            prefactorize_separately=True)
        assert (st != 0) == (out["status"][b] != 0), b
        if st != 0:
            continue
        scale = max(1.0, float(np.max(np.abs(x))))
        if use_pp and b % 5 == 0:  # perturbed pivot of 1e-13 * norm: the factors are huge, only the refined solution is comparable
            np.testing.assert_allclose(out["x"][b], x, rtol=0, atol=1e-8 * scale, err_msg=f"system {b}")
            continue
        np.testing.assert_allclose(out["x"][b], x, rtol=0, atol=1e-10 * scale, err_msg=f"system {b}")
        if N > 1:
            perm = bag.i64("perm").reshape(2, n, N)  # all p, then all q
            assert np.array_equal(out["perm"][b, :, 0, :], perm[0]) and np.array_equal(out["perm"][b, :, 1, :], perm[1]), b
        np.testing.assert_allclose(np.swapaxes(out["lu"][b], 1, 2), lu, rtol=1e-9, atol=1e-9)
    if use_pp:
        assert out["perturbed"][::5].all() and (out["n_solves"][::5] >= 2).all()
        assert not out["perturbed"][1::5].any() and (out["n_solves"][1::5] == 1).all()
