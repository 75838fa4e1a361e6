"""Pins the oracle's block-sparse LU against the reference's known-answer tests
(tests/cpp_unit_tests/math_solver/test_sparse_lu_solver.cpp:69-251 matrices, :253-333 dense LU, :335-466 sparse solve,
:468-545 ill-conditioned system with/without pivot perturbation)."""
import numpy as np
import pytest

import oracle_lib as orc

ROW_INDPTR = [0, 3, 6, 9]
COL_INDICES = [0, 1, 2, 0, 1, 2, 0, 1, 2]
DIAG_LU = [0, 4, 8]


def test_scalar_3x3_with_fill_ins():
    data = np.array([4.0, 1.0, 5.0, 3.0, 7.0, 0.0, 2.0, 0.0, 6.0])
    rhs = np.array([21.0, 2.0, 18.0])
    for separately in (False, True):
        st, x, lu, _ = orc.sparse_lu_solve(1, ROW_INDPTR, COL_INDICES, DIAG_LU, data, rhs, prefactorize_separately=separately)
        assert st == 0
        np.testing.assert_allclose(x.ravel(), [3.0, -1.0, 2.0], atol=1e-12)


def test_scalar_singular():
    data = np.array([0.0, 1.0, 5.0, 3.0, 7.0, 0.0, 2.0, 0.0, 6.0])
    st, *_ = orc.sparse_lu_solve(1, ROW_INDPTR, COL_INDICES, DIAG_LU, data, [21.0, 2.0, 18.0])
    assert st == orc.STATUS_SINGULAR


BLOCKS_WITH_FILL_INS = np.array([
    [[0, 1], [100, 0]], [[1, 2], [7, -1]], [[3, 4], [5, 6]],
    [[1, 2], [-3, 4]], [[0, 200], [3, 1]], [[0, 0], [0, 0]],
    [[5, 6], [-7, 8]], [[0, 0], [0, 0]], [[1, 0], [0, 100]],
], dtype=float)


def test_block_2x2_with_fill_ins():
    rhs = np.array([[38, 356], [-389, 2], [44, 611]], dtype=float)
    for separately in (False, True):
        st, x, lu, bag = orc.sparse_lu_solve(2, ROW_INDPTR, COL_INDICES, DIAG_LU, BLOCKS_WITH_FILL_INS, rhs,
                                             prefactorize_separately=separately)
        assert st == 0
        np.testing.assert_allclose(x, [[3, 4], [-1, -2], [5, 6]], atol=1e-10)


def test_block_2x2_without_fill_ins():
    indptr, indices, diag = [0, 2, 5, 7], [0, 1, 0, 1, 2, 1, 2], [0, 3, 6]
    data = np.array([
        [[0, 200], [3, 1]], [[1, 2], [-3, 4]],
        [[1, 2], [7, -1]], [[0, 1], [100, 0]], [[3, 4], [5, 6]],
        [[5, 6], [-7, 8]], [[1, 0], [0, 100]],
    ], dtype=float)
    rhs = np.array([[-389, 2], [38, 356], [44, 611]], dtype=float)
    st, x, *_ = orc.sparse_lu_solve(2, indptr, indices, diag, data, rhs)
    assert st == 0
    np.testing.assert_allclose(x, [[-1, -2], [3, 4], [5, 6]], atol=1e-10)


def test_block_four_node_meshed():
    indptr = [0, 3, 7, 10, 14]
    indices = [0, 1, 3, 0, 1, 2, 3, 1, 2, 3, 0, 1, 2, 3]
    diag = [0, 4, 8, 13]
    data = np.array([
        [[20, 1], [2, 21]], [[1, 0], [2, -1]], [[-1, 2], [0, 1]],
        [[0, 1], [-2, 1]], [[22, -1], [1, 23]], [[2, 1], [-1, 0]], [[0, 0], [0, 0]],
        [[1, -2], [0, 1]], [[24, 2], [-1, 25]], [[-2, 1], [1, 2]],
        [[1, 1], [-1, 2]], [[0, 0], [0, 0]], [[0, -1], [2, 1]], [[26, -2], [1, 27]],
    ], dtype=float)
    rhs = np.array([[21, 40], [-17, 64], [82, -47], [55, 38]], dtype=float)
    st, x, *_ = orc.sparse_lu_solve(2, indptr, indices, diag, data, rhs)
    assert st == 0
    np.testing.assert_allclose(x, [[1, 2], [-1, 3], [4, -2], [2, 1]], atol=1e-10)


def test_block_pseudo_singular():
    data = BLOCKS_WITH_FILL_INS.copy()
    data[0][0, 1] = 0.0
    st, *_ = orc.sparse_lu_solve(2, ROW_INDPTR, COL_INDICES, DIAG_LU, data, np.zeros((3, 2)))
    assert st == orc.STATUS_SINGULAR


def test_one_block_row_and_column_pivoting():
    """100 at (1,1) forces a row and a column swap: L U = P A Q with non-identity P and Q."""
    a = np.array([[[1.0, 2.0], [3.0, 100.0]]])
    st, x, lu, bag = orc.sparse_lu_solve(2, [0, 1], [0], [0], a, np.array([[5.0, 203.0]]))
    assert st == 0
    np.testing.assert_allclose(x, [[1.0, 2.0]], atol=1e-12)
    perm = bag.i64("perm")
    assert perm.tolist() == [1, 0, 1, 0]  # p then q
    lower = np.tril(lu[0], -1) + np.eye(2)
    upper = np.triu(lu[0])
    P = np.zeros((2, 2))
    Q = np.zeros((2, 2))
    for i in range(2):
        P[perm[i], i] = 1
        Q[perm[2 + i], i] = 1
    np.testing.assert_allclose(lower @ upper, P @ a[0] @ Q, atol=1e-12)


def test_dense_lu_3x3_first_maximum_in_column_major_order():
    """Pivot search = cwiseAbs2().maxCoeff(): ties resolved by column-major visiting order."""
    a = np.array([[[2.0, -2.0, 0.0], [2.0, 1.0, 1.0], [0.0, 1.0, 2.0]]])
    st, x, lu, bag = orc.sparse_lu_solve(3, [0, 1], [0], [0], a, np.array([[0.0, 4.0, 3.0]]))
    assert st == 0
    np.testing.assert_allclose(a[0] @ x[0], [0.0, 4.0, 3.0], atol=1e-12)
    perm = bag.i64("perm")
    # first pivot: |2| ties at (0,0), (1,0), (0,1), (2,2): column-major first -> (0,0): no swap in step 0
    assert perm[0] == 0 and perm[3] == 0


@pytest.mark.parametrize("block", [False, True])
def test_ill_conditioned_needs_pivot_perturbation(block):
    if not block:
        indptr = [0, 4, 8, 12, 16]
        indices = list(range(4)) * 4
        diag = [0, 5, 10, 15]
        data = np.array([0, 0, 0, -1, 0, -1, 0, 0, 0, 0, 5, 1, -1, 0, 1, -9], dtype=float)
        rhs = np.array([0, 0, 50, 2], dtype=float)
        n = 1
    else:
        indptr, indices, diag = [0, 2, 4], [0, 1, 0, 1], [0, 3]
        data = np.array([[[0, 0], [0, -1]], [[0, -1], [0, 0]], [[0, 0], [-1, 0]], [[5, 1], [1, -9]]], dtype=float)
        rhs = np.array([[0, 0], [50, 2]], dtype=float)
        n = 2
    st, *_ = orc.sparse_lu_solve(n, indptr, indices, diag, data, rhs, use_pivot_perturbation=False, prefactorize_separately=True)
    assert st == orc.STATUS_SINGULAR
    st, x, *_ = orc.sparse_lu_solve(n, indptr, indices, diag, data, rhs, use_pivot_perturbation=True, prefactorize_separately=True)
    assert st == 0
    np.testing.assert_allclose(x.ravel(), [8, 0, 10, 0], atol=1e-8)


def test_random_block_systems_against_dense_solve():
    rng = np.random.default_rng(0)
    for N, cplx in ((2, False), (6, False), (3, True), (1, True)):
        n = 7
        # dense pattern, diagonally dominant enough to be regular but with off-diagonal pivots inside blocks
        indptr = [i * n for i in range(n + 1)]
        indices = list(range(n)) * n
        diag = [i * n + i for i in range(n)]
        data = rng.normal(size=(n * n, N, N))
        if cplx:
            data = data + 1j * rng.normal(size=(n * n, N, N))
        for i in range(n):
            data[i * n + i] += 10 * np.fliplr(np.eye(N))
        rhs = rng.normal(size=(n, N)) + (1j * rng.normal(size=(n, N)) if cplx else 0)
        dense = np.block([[data[i * n + j] for j in range(n)] for i in range(n)])
        st, x, *_ = orc.sparse_lu_solve(N, indptr, indices, diag, data, rhs)
        assert st == 0
        np.testing.assert_allclose(x.ravel(), np.linalg.solve(dense, rhs.ravel()), rtol=1e-9, atol=1e-9)
