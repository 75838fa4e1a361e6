import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, pgm_b200, oracle_lib as orc
opt = dict(pgm_b200.BENCHMARK_OPTION); opt["n_node_total_specified"] = 50000
grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **opt)
print({k: len(v) for k, v in grid.input_data.items()})
for sym, n_scn in ((True, 64), (False, 16)):
    update = grid.batch_update(n_scn, seed=1)
    model = pgm_b200.PowerGridModel(grid.input_data)
    t0 = time.perf_counter(); res = model.calculate_power_flow(symmetric=sym, update_data=update, output_component_types=["node", "line"]); dt = time.perf_counter() - t0
    t1 = time.perf_counter(); res = model.calculate_power_flow(symmetric=sym, update_data=update, output_component_types=["node", "line"]); dt2 = time.perf_counter() - t1
    pick = [0, n_scn - 1]
    ref = orc.Model(grid.input_data).calculate(sym=sym, update={k: np.ascontiguousarray(v[pick]) for k, v in update.items()}, threading=0, output_components=["node", "line"])
    du = np.max(np.abs(res["node"]["u_pu"][pick] - ref["node"]["u_pu"]))
    dp = np.max(np.abs(res["line"]["p_from"][pick] - ref["line"]["p_from"]) / (1e-6 * np.abs(ref["line"]["p_from"]) + 1e-3))
    print(f"sym={sym} n_scn={n_scn}: first {dt:.2f}s second {dt2:.3f}s timing {model.timing()} status_ok {(model.status == 0).all()} n_iter gpu {model.n_iter[pick]} oracle {ref['n_iter']} max|du| {du:.2e} p err/tol {dp:.2e}")
