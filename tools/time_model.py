"""Wall time of PowerGridModel.calculate_power_flow for a benchmark configuration (host buffers in / out)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pgm_b200

n_scn = int(os.environ.get("N_SCN", "1000"))
rings = os.environ.get("RINGS", "0") == "1"
sym = os.environ.get("ASYM", "0") != "1"
method = os.environ.get("METHOD", "newton_raphson")
grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=rings, has_lv_ring=rings, **pgm_b200.BENCHMARK_OPTION)
update = grid.batch_update(n_scn, seed=0)
model = pgm_b200.PowerGridModel(grid.input_data)
reuse = os.environ.get("REUSE", "0") == "1"
if reuse:  # page-locked update buffers as well
    pinned = {}
    for k, v in update.items():
        pinned[k] = pgm_b200.pinned_empty(v.shape, v.dtype)
        pinned[k][...] = v
    update = pinned
for rep in range(3):
    t0 = time.perf_counter()
    res = model.calculate_power_flow(symmetric=sym, update_data=update, calculation_method=method, reuse_output_buffers=reuse)
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {1e3 * dt:.1f} ms  {n_scn / dt:.0f} scenarios/s  timing {model.timing()}")
print("rings", rings, "sym", sym, "method", method, "mean iter", model.n_iter.mean())
