"""Wall time of BASELINE config 2 (1000 scenarios, radial benchmark grid, symmetric NR, all outputs) through
 (a) the reference's unchanged Python wrapper on libpgm_b200.so (baseline/_ref, pageable numpy outputs allocated per call),
 (b) pgm_b200.pgm_core (PGM_* symbols, pageable outputs),
 (c) pgm_b200.PowerGridModel with page-locked reused outputs (what bench.py's e2e measures)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pgm_b200  # noqa: E402
import reference_wrapper  # noqa: E402
from pgm_b200 import pgm_core  # noqa: E402

grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
update = grid.batch_update(1000, seed=0)


def best(f, n=7):
    f()
    f()
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t)
    return min(ts) * 1e3, sorted(ts)[len(ts) // 2] * 1e3


rows = []
pgm = reference_wrapper.load()
if pgm is not None:
    m = pgm.PowerGridModel(grid.input_data)
    rows.append(("reference wrapper (unchanged) -> PGM_calculate, all outputs", best(lambda: m.calculate_power_flow(update_data=update))))
    rows.append(("reference wrapper, node output only", best(lambda: m.calculate_power_flow(update_data=update, output_component_types=["node"]))))
m2 = pgm_core.PowerGridModel(grid.input_data)
rows.append(("pgm_core -> PGM_calculate, all outputs", best(lambda: m2.calculate_power_flow(update_data=update))))
m3 = pgm_b200.PowerGridModel(grid.input_data)
rows.append(("pgm_b200.PowerGridModel, page-locked reused outputs", best(lambda: m3.calculate_power_flow(update_data=update, reuse_output_buffers=True))))
for name, (lo, med) in rows:
    print(f"{name:70s} best {lo:8.2f} ms  median {med:8.2f} ms  ({1000 / med * 1e3:9.0f} scenarios/s)")
