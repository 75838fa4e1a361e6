"""Small driver for ncu: stage the BASELINE configs[1] batch once and run the solver kernel a few times."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pgm_b200

n_scn = int(os.environ.get("N_SCN", "1000"))
reps = int(os.environ.get("REPS", "3"))
rings = os.environ.get("RINGS", "0") == "1"
sym = os.environ.get("ASYM", "0") != "1"
method = os.environ.get("METHOD", "newton_raphson")
grid = pgm_b200.FictionalGrid(seed=0, has_mv_ring=rings, has_lv_ring=rings, **pgm_b200.BENCHMARK_OPTION)
update = grid.batch_update(n_scn, seed=0)
model = pgm_b200.PowerGridModel(grid.input_data)
eng = pgm_b200.Engine(symmetric=sym, phase_shift=model.math_real(0, sym, "phase_shift"),
                      branch_bus_idx=model.math_index(0, "branch_bus_idx"), sources_per_bus=model.math_index(0, "sources_per_bus"),
                      shunts_per_bus=model.math_index(0, "shunts_per_bus"), load_gens_per_bus=model.math_index(0, "load_gens_per_bus"),
                      load_gen_type=model.math_index(0, "load_gen_type"), fill_in=model.math_index(0, "fill_in"))
eng.set_param(model.math_real(0, sym, "branch_param").view(np.complex128), model.math_real(0, sym, "shunt_param").view(np.complex128),
              model.math_real(0, sym, "source_param").view(np.complex128))
s_inj, u_ref = model.batch_pf_input(update, symmetric=sym)
eng.stage(s_inj, u_ref)
ms = [eng.solve_staged(method=method) for _ in range(reps)]
out = eng.fetch()
print("solve ms:", ms, "mean iter", out["n_iter"].mean(), "levels", len(eng.index("level_ptr")) - 1,
      "method", method, "n_scn", n_scn, "tile", os.environ.get("PGMB_TILE", "auto"), "slots", os.environ.get("PGMB_SLOTS", "auto"))
