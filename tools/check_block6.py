"""Row-split asymmetric kernel (PGMB_BLOCK6=1) against the one-thread-per-row kernel: bit-identical voltages, same iteration counts,
and the kernel time of both on BASELINE configs[2]."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pgm_b200

n_scn = int(os.environ.get("N_SCN", "1000"))
rings = os.environ.get("RINGS", "1") == "1"
grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=rings, has_lv_ring=rings, **pgm_b200.BENCHMARK_OPTION)
update = grid.batch_update(n_scn, seed=0)
model = pgm_b200.PowerGridModel(grid.input_data)
sym = False
eng = pgm_b200.Engine(symmetric=sym, phase_shift=model.math_real(0, sym, "phase_shift"),
                      branch_bus_idx=model.math_index(0, "branch_bus_idx"), sources_per_bus=model.math_index(0, "sources_per_bus"),
                      shunts_per_bus=model.math_index(0, "shunts_per_bus"), load_gens_per_bus=model.math_index(0, "load_gens_per_bus"),
                      load_gen_type=model.math_index(0, "load_gen_type"), fill_in=model.math_index(0, "fill_in"))
eng.set_param(model.math_real(0, sym, "branch_param").view(np.complex128), model.math_real(0, sym, "shunt_param").view(np.complex128),
              model.math_real(0, sym, "source_param").view(np.complex128))
s_inj, u_ref = model.batch_pf_input(update, symmetric=sym)
res = {}
for flag in ("0", "1"):
    os.environ["PGMB_BLOCK6"] = flag
    eng.stage(s_inj, u_ref)
    ms = [eng.solve_staged() for _ in range(3)]
    out = eng.fetch()
    res[flag] = out
    print(f"PGMB_BLOCK6={flag}: solve ms {[round(x, 2) for x in ms]} mean iter {out['n_iter'].mean():.3f} failed {(out['status'] != 0).sum()}")
same_iter = np.array_equal(res["0"]["n_iter"], res["1"]["n_iter"])
du = np.abs(res["0"]["u"] - res["1"]["u"]).max()
print("iterations equal:", same_iter, " max |du| between the kernels:", du, " bit-identical:", np.array_equal(res["0"]["u"], res["1"]["u"]))
