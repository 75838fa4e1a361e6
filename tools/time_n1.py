"""N-k batch (K lines switched off per scenario, default 1): wall time of pgm_b200 against the oracle port (config 5 shape).
PGMB_OUTAGE_SLOTS=1 with K > 1 shows the scenario-by-scenario route the overlay with several branch slots replaces."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pgm_b200

n_scn = int(os.environ.get("N_SCN", "64"))
sym = os.environ.get("ASYM", "0") != "1"
nodes = int(os.environ.get("NODES", "1500"))
opt = dict(pgm_b200.BENCHMARK_OPTION)
opt["n_node_total_specified"] = nodes
grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **opt)
lines = grid.input_data["line"]
rng = np.random.default_rng(0)
k_branch = int(os.environ.get("K", "1"))
upd = pgm_b200.structs.initialize_array("update", "line", (n_scn, k_branch))
if k_branch == 1:
    upd["id"][:, 0] = lines["id"][rng.choice(len(lines), n_scn, replace=False)]
else:
    for s in range(n_scn):
        upd["id"][s] = lines["id"][rng.choice(len(lines), k_branch, replace=False)]
upd["from_status"] = 0
upd["to_status"] = 0
update = {"line": upd}
model = pgm_b200.PowerGridModel(grid.input_data)
model.calculate_power_flow(symmetric=sym)
for _ in range(2):  # the second call reuses the device buffers and the page-locked output of the first
    t0 = time.perf_counter()
    res = model.calculate_power_flow(symmetric=sym, update_data=update, output_component_types=["node"], continue_on_batch_error=True,
                                     threading=int(os.environ.get("THREADS", "-1")), reuse_output_buffers=True)
    dt = time.perf_counter() - t0
print(f"pgm_b200: {n_scn} N-{k_branch} scenarios on {len(grid.input_data['node'])} nodes sym={sym}: {1e3 * dt / n_scn:.3f} ms/scenario ({1e3 * dt:.1f} ms wall, second call), failed {int((model.status != 0).sum())}")
if os.environ.get("ORACLE", "1") == "1":
    import oracle_lib as orc
    om = orc.Model(grid.input_data)
    t0 = time.perf_counter()
    ref = om.calculate(sym=sym, update=update, threading=0, output_components=["node"])
    dt = time.perf_counter() - t0
    print(f"oracle (all cores): {1e3 * dt / n_scn:.1f} ms/scenario; n_iter equal {np.array_equal(ref['n_iter'], model.n_iter)}")
    ok = model.status == 0
    print("max |du_pu|", float(np.max(np.abs(res["node"]["u_pu"][ok] - ref["node"]["u_pu"][ok]))))
print("timing", {k: round(v, 2) for k, v in model.timing().items()})
