import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracle_lib as orc, pgm_b200, grids
rings = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
_, inp, update = grids.regulated_benchmark_grid(rings, q_lim=6e6, n_scn=24)
model = pgm_b200.PowerGridModel(inp)
res = model.calculate_power_flow(symmetric=True, update_data=update, continue_on_batch_error=True)
ref = orc.Model(inp).calculate(sym=True, update=update, threading=0)
print("batch status", model.status.tolist()); print("n_iter", model.n_iter.tolist()); print("ref  ", ref["n_iter"].tolist())
print("viol", res["voltage_regulator"]["limit_violated"][:4].tolist()); print("ref ", ref["voltage_regulator"]["limit_violated"][:4].tolist())
try:
    single = model.calculate_power_flow(symmetric=True)
    print("single ok", model.n_iter[0])
except Exception as e:
    print("single failed", e)
ref1 = orc.Model(inp).calculate(sym=True)
print("ref single", ref1["n_iter"], ref1["voltage_regulator"]["limit_violated"].tolist())
