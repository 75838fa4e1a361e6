#!/usr/bin/env python
"""Automatic tap changer on the reference benchmark's shape (benchmark.cpp:333-422: fictional grid + one regulator on the station
transformer, load-profile batch, symmetric Newton-Raphson): the lockstep search on the GPU (one batched power flow per search
step), the scenario-by-scenario search on the GPU (PGMB_TAP_EXACT=1), and the oracle on all host cores.  Tap positions of the
three are compared."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import pgm_b200  # noqa: E402

n_scn = int(os.environ.get("N_SCN", "1000"))
with_oracle = os.environ.get("ORACLE", "1") == "1"
sym = os.environ.get("ASYM", "0") != "1"
grid = pgm_b200.FictionalGrid(seed=0, has_tap_changer=True, **pgm_b200.BENCHMARK_OPTION)
update = grid.batch_update(n_scn, seed=0)
comps = ["node", "transformer", "transformer_tap_regulator"]
model = pgm_b200.PowerGridModel(grid.input_data)
if with_oracle:
    import oracle_lib as orc

    oracle = orc.Model(grid.input_data)


def best(fn, reps=3):
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        t.append(time.perf_counter() - t0)
    return min(t), r


for strategy in ("any_valid_tap", "fast_any_tap", "min_voltage_tap", "max_voltage_tap"):
    kw = dict(symmetric=sym, update_data=update, tap_changing_strategy=strategy, output_component_types=comps, reuse_output_buffers=True)
    os.environ.pop("PGMB_TAP_EXACT", None)
    t_lock, r_lock = best(lambda: model.calculate_power_flow(**kw), reps=4)  # results live in the model's page-locked buffers
    r_lock = {k: v.copy() for k, v in r_lock.items()}
    timing = model.timing()
    os.environ["PGMB_TAP_EXACT"] = "1"
    t_exact, r_exact = best(lambda: model.calculate_power_flow(**kw), reps=2)
    r_exact = {k: v.copy() for k, v in r_exact.items()}
    os.environ.pop("PGMB_TAP_EXACT", None)
    taps = r_lock["transformer_tap_regulator"]["tap_pos"]
    line = (f"{strategy:16s} {n_scn} scenarios: lockstep {1e3 * t_lock:8.1f} ms ({n_scn / t_lock:9.0f}/s)  scenario-by-scenario "
            f"{1e3 * t_exact:8.1f} ms ({n_scn / t_exact:8.0f}/s)  taps equal {np.array_equal(taps, r_exact['transformer_tap_regulator']['tap_pos'])}"
            f"  max |du| {np.max(np.abs(r_lock['node']['u_pu'] - r_exact['node']['u_pu'])):.1e}")
    if with_oracle:
        cpu_out = {c: np.zeros((n_scn, len(grid.input_data[c])), (pgm_b200.structs.SYM_OUTPUT if sym else pgm_b200.structs.ASYM_OUTPUT)[c])
                   for c in comps}
        t_cpu, r_cpu = best(lambda: oracle.calculate(sym=sym, update=update, threading=0, tap_changing_strategy=strategy,
                                                     output_components=comps, out=cpu_out), reps=2)
        line += (f"  | oracle {1e3 * t_cpu:8.1f} ms ({n_scn / t_cpu:7.0f}/s, {orc.lib.orc_hardware_concurrency()} threads)  taps equal "
                 f"{np.array_equal(taps, r_cpu['transformer_tap_regulator']['tap_pos'])}  speed-up {t_cpu / t_lock:.1f}x")
    line += f"  [lockstep call: solver kernels {timing['solve_kernel']:.1f} ms, total {timing['total']:.1f} ms]"
    print(line + f"  tap positions {dict(zip(*[x.tolist() for x in np.unique(taps, return_counts=True)]))}")
