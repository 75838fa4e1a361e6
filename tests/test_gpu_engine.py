"""GPU parity tests at the math-solver seam: libpgm_b200.so (CUDA, through the C-ABI) against the oracle on the same
inputs.  Bars (BASELINE.json north_star): iteration counts identical, node voltages within 1e-9 pu, currents / powers
within 1e-6 relative (here: 1e-9 absolute per-unit on everything, far tighter)."""
import numpy as np
import pytest

import oracle_lib as orc
import pgm_b200
from grids import random_grid, random_scenarios, three_bus_grid

pytestmark = pytest.mark.gpu
KEYS = ("u", "bus_injection", "branch", "source", "shunt", "load_gen")
U_TOL = 1e-9


def oracle_batch(grid, s, u_ref, method="newton_raphson", err_tol=1e-8, max_iter=20):
    outs = []
    for k in range(s.shape[0]):
        grid.s_injection = s[k].reshape(-1, grid.B)
        grid.source_u_ref = np.atleast_1d(u_ref[k] if np.ndim(u_ref) == 2 else u_ref).astype(complex)
        outs.append(orc.math_pf(grid, method, err_tol, max_iter))
    return outs


def compare(gpu, oracle_outs, tol=U_TOL):
    for k, ref in enumerate(oracle_outs):
        assert gpu["status"][k] == ref["status"], (k, gpu["status"][k], ref["status"], ref["error"])
        if ref["status"] != 0:
            continue
        assert gpu["n_iter"][k] == ref["num_iter"], (k, gpu["n_iter"][k], ref["num_iter"])
        for key in KEYS:
            if key in gpu:
                err = np.max(np.abs(gpu[key][k] - ref[key])) if ref[key].size else 0.0
                assert err < tol, (k, key, err)


def test_three_bus_known_answer():
    grid, expected = three_bus_grid(True)
    eng = pgm_b200.Engine.from_grid(grid)
    out = eng.run(grid.s_injection[None], grid.source_u_ref, err_tol=1e-12, max_iter=20)
    assert out["status"][0] == 0
    for key in KEYS:
        assert np.max(np.abs(out[key][0] - expected[key])) < 1e-12, key
    ref = orc.math_pf(grid, "newton_raphson", 1e-12, 20)
    assert out["n_iter"][0] == ref["num_iter"]


def test_three_bus_asymmetric_known_answer():
    """reference test_math_solver_pf_nr 'Test asymmetric' closed form, three-phase 6 x 6 block kernel"""
    grid, expected = three_bus_grid(False)
    eng = pgm_b200.Engine.from_grid(grid)
    s = np.repeat(grid.s_injection[None], 9, axis=0)
    out = eng.run(s, grid.source_u_ref, err_tol=1e-12, max_iter=20)
    ref = orc.math_pf(grid, "newton_raphson", 1e-12, 20)
    assert (out["status"] == 0).all() and (out["n_iter"] == ref["num_iter"]).all()
    for key in KEYS:
        for k in (0, 8):
            assert np.max(np.abs(out[key][k] - expected[key])) < 1e-12, key
            assert np.max(np.abs(out[key][k] - ref[key])) < 1e-13, key
    grid, _ = three_bus_grid(False, diverge=True)
    out = pgm_b200.Engine.from_grid(grid).run(grid.s_injection[None], grid.source_u_ref, err_tol=1e-12, max_iter=20)
    assert out["status"][0] == 1 and out["n_iter"][0] == 20
    grid, _ = three_bus_grid(False, singular=True)
    out = pgm_b200.Engine.from_grid(grid).run(grid.s_injection[None], grid.source_u_ref, err_tol=1e-12, max_iter=20)
    assert out["status"][0] == 2


@pytest.mark.parametrize("method", ["iterative_current", "linear", "linear_current"])
def test_three_bus_asymmetric_other_methods(method):
    """three-phase complex 3 x 3 block kernels (cplx_block.cu) against the oracle; the linear method against the
    closed form of the const-impedance variant (test_math_solver_pf_linear 'Test asymmetric')"""
    grid, expected = three_bus_grid(False, const_z=(method == "linear"))
    eng = pgm_b200.Engine.from_grid(grid)
    s = np.repeat(grid.s_injection[None], 5, axis=0)
    out = eng.run(s, grid.source_u_ref, method=method, err_tol=1e-12, max_iter=100)
    ref = orc.math_pf(grid, method, 1e-12, 100)
    assert (out["status"] == 0).all() and (out["n_iter"] == ref["num_iter"]).all()
    for key in KEYS:
        assert np.max(np.abs(out[key][4] - ref[key])) < 1e-12, key
        if method == "linear":
            assert np.max(np.abs(out[key][0] - expected[key])) < 1e-11, key


def test_generic_block_kernel_equals_symmetric_kernel(monkeypatch):
    """the generic (2B x 2B) kernel instantiated for B = 1 must reproduce the specialised symmetric kernel bit for bit"""
    grid = random_grid(120, 15, 9, n_source=2)
    s, u_ref = random_scenarios(grid, 21, 9)
    a = pgm_b200.Engine.from_grid(grid).run(s, u_ref)
    monkeypatch.setenv("PGMB_KERNEL", "0")
    b = pgm_b200.Engine.from_grid(grid).run(s, u_ref)
    assert np.array_equal(a["n_iter"], b["n_iter"]) and np.array_equal(a["u"], b["u"])


@pytest.mark.parametrize("rings", [True, False])
def test_row_split_block_kernel_equals_block_kernel(monkeypatch, rings):
    """the opt-in asymmetric kernel with the row task split over six threads (nr_block6.cu, PGMB_BLOCK6=1: block row per thread,
    6 x 6 full-pivot factorisation by warp shuffles, hub rows on the cooperative code) must reproduce the one-thread-per-row
    kernel bit for bit -- every output value of every component, and the iteration counts"""
    grid = pgm_b200.FictionalGrid(seed=0, n_node_total_specified=300, n_connection_per_lv_feeder=5, n_lv_feeder=4, n_node_per_mv_feeder=5,
                                  n_mv_feeder=3, has_mv_ring=rings, has_lv_ring=rings)
    update = grid.batch_update(21, seed=3)
    model = pgm_b200.PowerGridModel(grid.input_data)
    a = model.calculate_power_flow(symmetric=False, update_data=update)
    it_a = model.n_iter.copy()
    monkeypatch.setenv("PGMB_BLOCK6", "1")
    b = model.calculate_power_flow(symmetric=False, update_data=update)
    assert np.array_equal(it_a, model.n_iter) and (model.status == 0).all()
    for comp in a:
        for name in a[comp].dtype.names:
            assert np.array_equal(a[comp][name], b[comp][name], equal_nan=True), (comp, name)


@pytest.mark.parametrize("n_node,seed,n_scn", [(60, 11, 9), (400, 12, 33), (1500, 13, 40)])
def test_path_kernel_equals_level_kernel_on_radial_grids(monkeypatch, n_node, seed, n_scn):
    """radial grids run the path kernel (nr_sym_v3.cu); it must reproduce the level-scheduled kernel bit for bit"""
    grid = random_grid(n_node, 0, seed, n_source=1)
    s, u_ref = random_scenarios(grid, n_scn, seed)
    eng = pgm_b200.Engine.from_grid(grid)
    assert eng.index("path_program").size > 0
    a = eng.run(s, u_ref)
    monkeypatch.setenv("PGMB_KERNEL", "2")
    b = pgm_b200.Engine.from_grid(grid).run(s, u_ref)
    assert np.array_equal(a["n_iter"], b["n_iter"]) and np.array_equal(a["status"], b["status"])
    for key in KEYS:
        assert np.array_equal(a[key], b[key]), key
    compare(a, oracle_batch(grid, s, u_ref))


def test_three_bus_diverge_and_singular_status():
    grid, _ = three_bus_grid(True, diverge=True)
    eng = pgm_b200.Engine.from_grid(grid)
    out = eng.run(grid.s_injection[None], grid.source_u_ref, err_tol=1e-12, max_iter=20)
    assert out["status"][0] == 1 and out["n_iter"][0] == 20
    grid, _ = three_bus_grid(True, singular=True)
    eng = pgm_b200.Engine.from_grid(grid)
    out = eng.run(grid.s_injection[None], grid.source_u_ref, err_tol=1e-12, max_iter=20)
    assert out["status"][0] == 2


def test_mixed_batch_keeps_good_scenarios():
    """one diverging scenario must not disturb its tile neighbours (BatchCalculationError semantics)"""
    grid, expected = three_bus_grid(True)
    eng = pgm_b200.Engine.from_grid(grid)
    s = np.repeat(grid.s_injection[None], 11, axis=0)
    s[4, 6, 0] = 1e6
    out = eng.run(s, grid.source_u_ref, err_tol=1e-12, max_iter=20)
    assert out["status"].tolist() == [0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    for k in range(11):
        if k != 4:
            assert np.max(np.abs(out["u"][k] - expected["u"])) < 1e-12


@pytest.mark.parametrize("n_node,extra,seed,n_scn", [(40, 0, 1, 5), (250, 0, 2, 37), (80, 20, 3, 19), (300, 45, 4, 70)])
def test_random_grids_match_oracle(n_node, extra, seed, n_scn):
    grid = random_grid(n_node, extra, seed, n_source=2 if seed % 2 == 0 else 1)
    s, u_ref = random_scenarios(grid, n_scn, seed)
    eng = pgm_b200.Engine.from_grid(grid)
    out = eng.run(s, u_ref)
    compare(out, oracle_batch(grid, s, u_ref))
    assert (out["status"] == 0).all()


@pytest.mark.parametrize("tile", ["4", "8", "16", "32"])
def test_every_tile_width(tile, monkeypatch):
    monkeypatch.setenv("PGMB_TILE", tile)
    grid = random_grid(120, 10, 7)
    s, u_ref = random_scenarios(grid, 45, 7)
    out = pgm_b200.Engine.from_grid(grid).run(s, u_ref)
    compare(out, oracle_batch(grid, s, u_ref))


def test_heavy_load_iteration_counts():
    """stress the convergence decision: heavier loading => more iterations, still identical counts"""
    grid = random_grid(150, 5, 11)
    s, u_ref = random_scenarios(grid, 40, 11, scale=3.0)
    out = pgm_b200.Engine.from_grid(grid).run(s, u_ref, max_iter=30)
    refs = oracle_batch(grid, s, u_ref, max_iter=30)
    compare(out, refs)
    assert len({r["num_iter"] for r in refs if r["status"] == 0}) > 1


def test_staged_solve_is_idempotent_and_timed():
    grid = random_grid(200, 0, 13)
    s, u_ref = random_scenarios(grid, 64, 13)
    eng = pgm_b200.Engine.from_grid(grid)
    eng.stage(s, u_ref)
    ms1 = eng.solve_staged()
    a = eng.fetch()
    ms2 = eng.solve_staged()
    b = eng.fetch()
    assert ms1 > 0 and ms2 > 0
    assert np.array_equal(a["u"], b["u"]) and np.array_equal(a["n_iter"], b["n_iter"])


@pytest.mark.parametrize("method", ["iterative_current", "linear", "linear_current"])
def test_three_bus_other_methods(method):
    grid, expected = three_bus_grid(True)
    eng = pgm_b200.Engine.from_grid(grid)
    out = eng.run(grid.s_injection[None], grid.source_u_ref, method=method, err_tol=1e-12, max_iter=100)
    ref = orc.math_pf(grid, method, 1e-12, 100)
    assert out["status"][0] == 0 and ref["status"] == 0
    assert out["n_iter"][0] == ref["num_iter"]
    for key in KEYS:
        assert np.max(np.abs(out[key][0] - ref[key])) < 1e-12, key
        if method == "iterative_current":
            assert np.max(np.abs(out[key][0] - expected[key])) < 1e-11, key


def test_three_bus_const_z_is_forced_linear():
    grid, expected = three_bus_grid(True, const_z=True)
    grid.load_gen_type[:] = 1
    eng = pgm_b200.Engine.from_grid(grid)
    out = eng.run(grid.s_injection[None], grid.source_u_ref, method="newton_raphson")
    for key in KEYS:
        assert np.max(np.abs(out[key][0] - expected[key])) < 1e-8, key


@pytest.mark.parametrize("method", ["iterative_current", "linear", "linear_current"])
@pytest.mark.parametrize("n_node,extra,seed,n_scn", [(250, 0, 2, 37), (300, 45, 4, 70)])
def test_random_grids_other_methods(method, n_node, extra, seed, n_scn):
    grid = random_grid(n_node, extra, seed, n_source=2 if seed % 2 == 0 else 1)
    s, u_ref = random_scenarios(grid, n_scn, seed)
    eng = pgm_b200.Engine.from_grid(grid)
    out = eng.run(s, u_ref, method=method, max_iter=100)
    compare(out, oracle_batch(grid, s, u_ref, method=method, max_iter=100))
    assert (out["status"] == 0).all()


@pytest.mark.parametrize("method", ["iterative_current", "linear"])
def test_singular_and_diverge_other_methods(method):
    grid, _ = three_bus_grid(True, singular=True)
    out = pgm_b200.Engine.from_grid(grid).run(grid.s_injection[None], grid.source_u_ref, method=method, err_tol=1e-12)
    assert out["status"][0] == 2
    if method == "iterative_current":
        grid, _ = three_bus_grid(True, diverge=True)
        out = pgm_b200.Engine.from_grid(grid).run(grid.s_injection[None], grid.source_u_ref, method=method, err_tol=1e-12, max_iter=20)
        assert out["status"][0] == 1 and out["n_iter"][0] == 20
