"""Latency of one symmetric / asymmetric power flow (BASELINE configs[0]) through the model API, GPU vs the oracle port."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc
import pgm_b200

grid = pgm_b200.FictionalGrid(seed=0, **pgm_b200.BENCHMARK_OPTION)
model = pgm_b200.PowerGridModel(grid.input_data)
om = orc.Model(grid.input_data)
for sym in (True, False):
    model.calculate_power_flow(symmetric=sym)
    om.calculate(sym=sym)
    t0 = time.perf_counter()
    for _ in range(20):
        model.calculate_power_flow(symmetric=sym)
    gpu = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(5):
        om.calculate(sym=sym)
    cpu = (time.perf_counter() - t0) / 5
    print(f"single scenario sym={sym}: pgm_b200 {1e3 * gpu:.2f} ms (solve kernel {model.timing()['solve_kernel']:.2f} ms), oracle port 1 thread {1e3 * cpu:.2f} ms")
