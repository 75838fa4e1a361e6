#!/usr/bin/env python
"""In-process multi-GPU strong scaling: ONE calculate call (pgmb_model_calculate, n_devices = 1 .. all) on one batch.
   config 2 at 8000 scenarios (sym NR), config 4 (100 000 scenarios, iterative_current / linear, node output),
   config 3 at 2048 scenarios (asym NR).  Outputs of every n_devices are compared byte for byte with the one-GPU result.
   Wall clock around the call (host buffers in / out, page-locked, reused), best of 3."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import pgm_b200  # noqa: E402


def run(name, grid, n_scn, sym, method, comps, max_iter=20):
    n_gpu = pgm_b200.lib().pgmb_device_count()
    update = grid.batch_update(n_scn, seed=0)
    host_update = {k: pgm_b200.pinned_empty(v.shape, v.dtype) for k, v in update.items()}
    for k, v in update.items():
        host_update[k][...] = v
    table = pgm_b200.structs.SYM_OUTPUT if sym else pgm_b200.structs.ASYM_OUTPUT
    out = {c: pgm_b200.pinned_empty((n_scn, len(grid.input_data[c])), table[c]) for c in comps}
    model = pgm_b200.PowerGridModel(grid.input_data)
    ref_bytes, rows = None, []
    for n_dev in [d for d in (1, 2, 4, 8) if d <= n_gpu]:
        best = None
        for _ in range(4):
            t0 = time.perf_counter()
            model.calculate_power_flow(symmetric=sym, calculation_method=method, update_data=host_update, output_buffers=out,
                                       output_component_types=list(comps), n_devices=n_dev, max_iterations=max_iter)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        assert (model.status == 0).all()
        b = b"".join(out[c].tobytes() for c in comps)
        if ref_bytes is None:
            ref_bytes = b
        rows.append({"n_devices": n_dev, "ms": 1e3 * best, "scenarios_per_s": n_scn / best, "identical_to_1gpu": b == ref_bytes,
                     "d2h_gbs": sum(a.nbytes for a in out.values()) / best / 1e9})
    print(json.dumps({"config": name, "n_scenarios": n_scn, "outputs": list(comps), "runs": rows}))


def main():
    radial = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    ringed = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **pgm_b200.BENCHMARK_OPTION)
    full = ["node", "line", "transformer", "shunt", "source", "sym_load", "asym_load"]
    import psutil

    big = 100000 if psutil.virtual_memory().available > (96 << 30) else 25000  # 100k scenarios need ~20 GB of page-locked buffers
    run("configs[1] x 8000, sym NR, all outputs", radial, 8000, True, "newton_raphson", full)
    run("configs[1] x 8000, sym NR, node output", radial, 8000, True, "newton_raphson", ["node"])
    run(f"configs[3] {big} iterative_current, node output", radial, big, True, "iterative_current", ["node"], max_iter=100)
    run(f"configs[3] {big} linear, node output", radial, big, True, "linear", ["node"])
    run("configs[2] x 2048, asym NR, node output", ringed, 2048, False, "newton_raphson", ["node"])


if __name__ == "__main__":
    main()
