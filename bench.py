#!/usr/bin/env python
"""Benchmark of the batch power-flow hot path (BASELINE.json metric: batch power-flow scenarios/sec, NR, fp64).

Workload (config.workload): BASELINE configs[1] -- the reference's fictional 1500-node-spec radial grid (seed 0: 2605
nodes, tests/benchmark_cpp/benchmark.cpp:257-263), symmetric Newton-Raphson, err_tol 1e-8, max_iter 20, 1000 load-profile
update scenarios per GPU per step (generate_batch_input, seed 0 + rank).  One step = one pass over the batch.

  value     scenarios/s with inputs resident in HBM: solver kernel(s) + device-side result kernels, CUDA events
  e2e       scenarios/s through the public model API with HOST update buffers in and HOST output structs out
  roofline  algorithmic HBM bytes of the NR kernel / its measured duration, against MEASURED_PEAKS.json
  cpu_baseline  the oracle (CPU restatement of the reference path) on the box's host cores, all threads

`--impl reference` times the reference's CPU path instead (the oracle port: the reference itself cannot be built in this
image) on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SCN = 1000
ERR_TOL = 1e-8
MAX_ITER = 20


def measured_peak_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs: burst copy figure, the only HBM figure there)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the solver kernel from the committed ncu --set full capture"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def algorithmic_bytes_per_solve(n_bus, nnz_lu, n_lg, b=1):
    """A_iter = 3*B_J + 10*n_bus*2b*8 + n_lg*2b*8 (one linear solve with a fresh matrix)"""
    b_j = nnz_lu * (2 * b) ** 2 * 8
    return 3 * b_j + 10 * n_bus * 2 * b * 8 + n_lg * 2 * b * 8


def run_reference(args, rank, world):
    """CPU arm: the oracle port with the reference's dispatch shape (threading = 0: all cores, stride scheduling)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    import pgm_b200

    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    cores = int(orc.lib.orc_hardware_concurrency())
    sample = max(64, min(N_SCN, 16 * cores))  # bounded sample of the 1000-scenario batch per step
    update = {k: v[:sample] for k, v in grid.batch_update(N_SCN, seed=0).items()}
    model = orc.Model(grid.input_data)
    for _ in range(args.warmup):
        model.calculate(sym=True, update=update, threading=0, err_tol=ERR_TOL, max_iter=MAX_ITER)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = model.calculate(sym=True, update=update, threading=0, err_tol=ERR_TOL, max_iter=MAX_ITER)
    dt = time.perf_counter() - t0
    assert res["n_failed"] == 0
    value = sample * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "batch power-flow scenarios/sec (NR, fp64)", "value": value, "unit": "scenarios/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(sample),
        "cpu_baseline": {"value": value, "unit": "scenarios/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} of the {N_SCN} scenarios per step, all {cores} host threads (reference threading=0)"},
        "e2e": {"value": value, "unit": "scenarios/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def other_configs(pgm_b200, np, device):
    """kernel milliseconds (CUDA events, second of two runs) of BASELINE configs 3 and 4 on one GPU; informational"""

    def staged_engine(grid, sym, n_scn, seed, method=None):
        model = pgm_b200.PowerGridModel(grid.input_data)
        eng = pgm_b200.Engine(symmetric=sym, phase_shift=model.math_real(0, sym, "phase_shift"),
                              branch_bus_idx=model.math_index(0, "branch_bus_idx"), sources_per_bus=model.math_index(0, "sources_per_bus"),
                              shunts_per_bus=model.math_index(0, "shunts_per_bus"), load_gens_per_bus=model.math_index(0, "load_gens_per_bus"),
                              load_gen_type=model.math_index(0, "load_gen_type"), fill_in=model.math_index(0, "fill_in"), device=device)
        eng.set_param(model.math_real(0, sym, "branch_param").view(np.complex128), model.math_real(0, sym, "shunt_param").view(np.complex128),
                      model.math_real(0, sym, "source_param").view(np.complex128))
        s_inj, u_ref = model.batch_pf_input(grid.batch_update(n_scn, seed=seed), symmetric=sym)
        eng.stage(s_inj, u_ref, method=method)
        return eng

    out = {}
    ringed = pgm_b200.FictionalGrid(seed=0, has_mv_ring=True, has_lv_ring=True, **pgm_b200.BENCHMARK_OPTION)
    eng = staged_engine(ringed, False, 1000, 0)
    ms = [eng.solve_staged(err_tol=ERR_TOL, max_iter=MAX_ITER) for _ in range(2)][-1]
    out["configs[2] ringed grid, asymmetric newton_raphson, 1000 scenarios"] = {"kernel_ms": ms, "scenarios_per_s": 1000 / ms * 1e3}
    radial = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    for method in ("iterative_current", "linear"):
        eng = staged_engine(radial, True, 12500, 0, method)
        ms = [eng.solve_staged(method=method, err_tol=ERR_TOL, max_iter=MAX_ITER) for _ in range(2)][-1]
        out[f"configs[3] radial grid, {method}, 12500 scenarios (one GPU's share of 100k over 8)"] = {
            "kernel_ms": ms, "scenarios_per_s": 12500 / ms * 1e3}
    # configs[4] shape at the 1500-node size: asymmetric N-1 batch (one line switched off per scenario) through the public
    # API; scenarios share the base grid's symbolic pattern (branch-outage overlay), node output only; second of two calls
    n1 = 1000
    lines = ringed.input_data["line"]
    upd = pgm_b200.structs.initialize_array("update", "line", (n1, 1))
    upd["id"][:, 0] = lines["id"][np.random.default_rng(0).choice(len(lines), n1, replace=False)]
    upd["from_status"][:, 0] = 0
    upd["to_status"][:, 0] = 0
    model = pgm_b200.PowerGridModel(ringed.input_data)
    for _ in range(2):
        t0 = time.perf_counter()
        model.calculate_power_flow(symmetric=False, update_data={"line": upd}, output_component_types=["node"],
                                   reuse_output_buffers=True, device=device)
        wall = 1e3 * (time.perf_counter() - t0)
    out["configs[4] shape, ringed 1804-node grid, asymmetric N-1 (1000 single-line outages, shared pattern), public API"] = {
        "wall_ms": wall, "kernel_ms": model.timing()["solve_kernel"], "scenarios_per_s": n1 / wall * 1e3,
        "failed": int((model.status != 0).sum())}
    return out


def workload_config(n_scn_per_gpu):
    return {"workload": "configs[1]: fictional radial grid n_node_total_specified=1500 (seed 0: 2605 nodes, 2600 lines, "
                        "7 transformers, 197 sym_load, 1200 asym_load), symmetric newton_raphson, err_tol 1e-8, max_iter 20",
            "scenarios_per_gpu_per_step": n_scn_per_gpu, "batch": "load-profile updates (generate_batch_input)",
            "l2_policy": "per-step working set (Jacobian/LU factors 250 MB per 1000 scenarios) exceeds the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch

    import pgm_b200

    if not torch.cuda.is_available() or pgm_b200.lib().pgmb_device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: pgm_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    cpu_binding = pgm_b200.distributed.bind_process_to_device_cpus(local_rank) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- build the workload: every rank owns its own scenarios (weak scaling, no data-path collective) ----
    grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
    update = grid.batch_update(N_SCN, seed=rank)
    model = pgm_b200.PowerGridModel(grid.input_data)
    n_bus = len(grid.input_data["node"])
    nnz_lu = len(model.math_index(0, "col_indices_lu"))
    n_lg = len(grid.input_data["sym_load"]) + len(grid.input_data["asym_load"])
    calc = dict(symmetric=True, calculation_method="newton_raphson", error_tolerance=ERR_TOL, max_iterations=MAX_ITER,
                device=local_rank)

    # ---- device-resident arm: engine level, inputs staged in HBM once ----
    eng = pgm_b200.Engine(symmetric=True, phase_shift=model.math_real(0, True, "phase_shift"),
                          branch_bus_idx=model.math_index(0, "branch_bus_idx"), sources_per_bus=model.math_index(0, "sources_per_bus"),
                          shunts_per_bus=model.math_index(0, "shunts_per_bus"), load_gens_per_bus=model.math_index(0, "load_gens_per_bus"),
                          load_gen_type=model.math_index(0, "load_gen_type"), fill_in=model.math_index(0, "fill_in"), device=local_rank)
    eng.set_param(model.math_real(0, True, "branch_param").view(np.complex128), model.math_real(0, True, "shunt_param").view(np.complex128),
                  model.math_real(0, True, "source_param").view(np.complex128))
    s_inj, u_ref = model.batch_pf_input(update)  # host-side PowerFlowInput of every scenario
    eng.stage(s_inj, u_ref)
    for _ in range(args.warmup):
        eng.solve_staged(err_tol=ERR_TOL, max_iter=MAX_ITER)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    kernel_ms = 0.0
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        kernel_ms += eng.solve_staged(err_tol=ERR_TOL, max_iter=MAX_ITER)  # CUDA events on the engine stream
    barrier()
    wall_dev = time.perf_counter() - t0
    gpu_launches = int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0  # counted by the library's launchers
    out = eng.fetch(full_output=False)
    assert (out["status"] == 0).all()
    mean_iter = float(out["n_iter"].mean())
    dev_time = max_over_ranks(kernel_ms / 1e3)
    value = world * N_SCN * args.steps / dev_time

    # ---- end-to-end arm: public model API, HOST (pinned) update buffers in / HOST (pinned) output structs out ----
    def pinned(shape, dtype):
        n = int(np.prod(shape)) * dtype.itemsize
        return torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=True).numpy()[:n].view(dtype).reshape(shape)

    host_update = {}
    for k, v in update.items():
        host_update[k] = pinned(v.shape, v.dtype)
        host_update[k][...] = v
    out_dtypes = pgm_b200.structs.SYM_OUTPUT
    host_out = {c: pinned((N_SCN, len(grid.input_data[c])), out_dtypes[c]) for c in
                ("node", "line", "transformer", "shunt", "source", "sym_load", "asym_load")}
    calc["output_buffers"] = host_out
    for _ in range(args.warmup):
        model.calculate_power_flow(update_data=host_update, **calc)
    barrier()
    launches0 = int(pgm_b200.lib().pgmb_kernel_launch_count())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = model.calculate_power_flow(update_data=host_update, **calc)
    barrier()
    e2e_time = max_over_ranks(time.perf_counter() - t0)
    e2e_launches = int(pgm_b200.lib().pgmb_kernel_launch_count()) - launches0
    clocks = sampler.stop() if rank == 0 else None
    timing = model.timing()
    h2d = sum(v.nbytes for v in update.values())
    d2h = sum(v.nbytes for v in res.values())

    # ---- informational, N = 1 only, after the timed regions: the same batch through the reference's own C API names
    #      (PGM_create_model / PGM_calculate, pgm_b200.pgm_core) with PAGEABLE numpy buffers allocated per call -- what an unchanged
    #      client of the reference's wrapper hands over; the library stages them through page-locked memory ----
    drop_in = None
    if world == 1:
        from pgm_b200 import pgm_core

        capi_model = pgm_core.PowerGridModel(grid.input_data)
        comps = list(host_out)
        ts = []
        for _ in range(5):
            t1 = time.perf_counter()
            capi_model.calculate_power_flow(update_data=update, output_component_types=comps)
            ts.append(time.perf_counter() - t1)
        med = sorted(ts[2:])[1]
        drop_in = {"value": N_SCN / med, "unit": "scenarios/s", "ms_per_step": 1e3 * med,
                   "path": "PGM_calculate (reference C API names), pageable host buffers allocated per call, all outputs"}

    # ---- the other BASELINE configs, solver kernels only, rank 0 (reported beside the bench line, not part of it; after
    #      every timed region so that they cannot disturb it) ----
    other = None
    if rank == 0 and os.environ.get("PGMB_BENCH_OTHER", "1") == "1":
        other = other_configs(pgm_b200, np, local_rank)

    if rank == 0:
        peak, peak_kind = measured_peak_hbm()
        a_solve = algorithmic_bytes_per_solve(n_bus, nnz_lu, n_lg)
        bytes_per_launch = N_SCN * (mean_iter + 1.0) * a_solve
        launch_s = (kernel_ms / 1e3) / args.steps
        achieved = bytes_per_launch / launch_s / 1e9
        # CPU baseline on a bounded sample (oracle port, all host threads)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as orc
        cores = int(orc.lib.orc_hardware_concurrency())
        sample = max(64, min(N_SCN, 16 * cores))
        cpu_model = orc.Model(grid.input_data)
        cpu_update = {k: v[:sample] for k, v in update.items()}
        cpu_model.calculate(sym=True, update=cpu_update, threading=0)
        reps = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 10.0 and reps < 1000:
            cpu_model.calculate(sym=True, update=cpu_update, threading=0, err_tol=ERR_TOL, max_iter=MAX_ITER)
            reps += 1
        cpu_value = sample * reps / (time.perf_counter() - t0)
        print(json.dumps({
            "metric": "batch power-flow scenarios/sec (NR, fp64)", "value": value, "unit": "scenarios/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_time / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(N_SCN), mean_nr_iterations=mean_iter, tile_width=os.environ.get("PGMB_TILE", "auto"),
                           **({"rank0_cpu_binding": f"{len(cpu_binding)} CPUs local to the GPU (NVML affinity)"} if cpu_binding else {})),
            "e2e": {"value": world * N_SCN * args.steps / e2e_time, "unit": "scenarios/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_time / args.steps, "last_step_breakdown_ms": timing,
                    "gpu_launches": e2e_launches, **({"drop_in_pageable": drop_in} if drop_in else {})},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(), "peak_kind": peak_kind, "kernel": "nr_sym_v3_kernel (path kernel of radial grids; the whole NR loop of a launch)",
                         "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": 1e3 * launch_s},
            "cpu_baseline": {"value": cpu_value, "unit": "scenarios/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} of the {N_SCN} scenarios x {reps} repeats, all {cores} host threads (reference threading=0)"},
            "clocks": clocks, "device_wall_ms_per_step": 1e3 * wall_dev / args.steps, "other_configs_kernel_only": other,
        }))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
