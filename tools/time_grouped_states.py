#!/usr/bin/env python
"""A batch that revisits a few switching states (ring lines out / in, a transformer tap moved) with different load profiles:
scenarios with identical structural updates run as one device batch per state (model.cpp: grouped route) against the
scenario-by-scenario route (PGMB_GROUP_SCENARIOS=0) and the oracle on all host cores."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import oracle_lib as orc  # noqa: E402
import pgm_b200  # noqa: E402

n_scn, n_state = int(os.environ.get("N_SCN", "1000")), 4
grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **pgm_b200.BENCHMARK_OPTION)
update = dict(grid.batch_update(n_scn, seed=0))
lines, trafos = grid.input_data["line"], grid.input_data["transformer"]
line_upd = pgm_b200.structs.initialize_array("update", "line", (n_scn, 2))
trafo_upd = pgm_b200.structs.initialize_array("update", "transformer", (n_scn, 1))
ring = lines["id"][-4:]
for s in range(n_scn):
    state = s % n_state
    line_upd["id"][s] = ring[[0, 1]] if state < 2 else ring[[2, 3]]
    line_upd["from_status"][s] = line_upd["to_status"][s] = 0 if state in (0, 2) else 1
    trafo_upd["id"][s, 0] = trafos["id"][1]
    trafo_upd["tap_pos"][s, 0] = [0, 1, -1, 2][state]
update["line"], update["transformer"] = line_upd, trafo_upd
model = pgm_b200.PowerGridModel(grid.input_data)
kw = dict(symmetric=True, update_data=update, output_component_types=["node", "line"], reuse_output_buffers=True)


def best(fn, reps=3):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        t.append(time.perf_counter() - t0)
    return min(t), r


t_overlay, res = best(lambda: model.calculate_power_flow(**kw))  # these states fit the branch overlay: one device batch
u_overlay = res["node"]["u_pu"].copy()
os.environ["PGMB_N1_EXACT"] = "1"  # without the overlay: one device batch per switching state
t_grouped, res = best(lambda: model.calculate_power_flow(**kw))
u_grouped = res["node"]["u_pu"].copy()
os.environ["PGMB_GROUP_SCENARIOS"] = "0"
t_single, res = best(lambda: model.calculate_power_flow(**kw), reps=2)
u_single = res["node"]["u_pu"].copy()
t_cpu, ref = best(lambda: orc.Model(grid.input_data).calculate(sym=True, update=update, threading=0, output_components=["node", "line"]), reps=2)
print(f"{n_scn} scenarios over {n_state} switching states (1804-bus ringed grid, sym NR): overlay {1e3 * t_overlay:.1f} ms "
      f"(max |du| vs oracle {np.max(np.abs(u_overlay - ref['node']['u_pu'])):.1e}), grouped {1e3 * t_grouped:.1f} ms, "
      f"scenario by scenario {1e3 * t_single:.1f} ms, oracle ({orc.lib.orc_hardware_concurrency()} threads) {1e3 * t_cpu:.1f} ms; "
      f"max |du| grouped vs single {np.max(np.abs(u_grouped - u_single)):.1e}, vs oracle {np.max(np.abs(u_grouped - ref['node']['u_pu'])):.1e}")
