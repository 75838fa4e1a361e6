import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, pgm_b200
os.environ["PGMB_N1_EXACT"] = "1"
grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **pgm_b200.BENCHMARK_OPTION)
n = int(os.environ.get('N_SCN', '256'))
lines = grid.input_data["line"]
upd = pgm_b200.structs.initialize_array("update", "line", (n, 1))
upd["id"][:, 0] = lines["id"][np.random.default_rng(0).choice(len(lines), n, replace=False)]
upd["from_status"][:, 0] = 0
upd["to_status"][:, 0] = 0
model = pgm_b200.PowerGridModel(grid.input_data)
for thr in [int(x) for x in os.environ.get('THREADS', '1,16').split(',')]:
    for _ in range(2):
        t0 = time.perf_counter()
        model.calculate_power_flow(update_data={"line": upd}, output_component_types=["node"], threading=thr, continue_on_batch_error=True)
        dt = time.perf_counter() - t0
    print("threads", thr, "ms per scenario", round(1e3 * dt / n, 3), "timing", {k: round(v, 2) for k, v in model.timing().items()})
