#!/bin/bash
# Every number of DESIGN.md section 7 in one run on one B200 (about 3 minutes).  Output: plain text, one block per BASELINE config.
set -u
echo "== config 1: single scenario latency (radial 2605-bus grid)"; python tools/time_single.py 2>&1 | tail -2
echo "== config 2: bench.py (1000 load-profile scenarios, symmetric NR)"; PGMB_BENCH_OTHER=0 python bench.py 2>/dev/null | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())
print('value %.0f scenarios/s (%.3f ms/step) | e2e %.0f scenarios/s (%.2f ms/step) | roofline frac %.3f | cpu_baseline %.0f scenarios/s on %d cores' % (
    d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']))"
echo "== config 2 at 8000 scenarios (kernel only)"; N_SCN=8000 REPS=3 python tools/profile_solve.py 2>&1 | tail -1
echo "== config 3: ringed grid, asymmetric NR, 1000 scenarios: kernel, then the public API with page-locked buffers"
ASYM=1 RINGS=1 REPS=3 python tools/profile_solve.py 2>&1 | tail -1
ASYM=1 RINGS=1 REUSE=1 python tools/time_model.py 2>&1 | tail -2
echo "== config 3 grid, symmetric NR (level kernel)"; RINGS=1 REPS=3 python tools/profile_solve.py 2>&1 | tail -1
echo "== config 4: radial grid, 12 500 scenarios (one GPU's share of 100 k over 8): iterative current / linear / linear current, public API"
for m in iterative_current linear linear_current; do METHOD=$m N_SCN=12500 REUSE=1 python tools/time_model.py 2>&1 | tail -2 | head -1; done
echo "== config 5 shape: N-1 (one line out per scenario), shared pattern; 1804-bus grid 1000 scenarios, 53 068-bus grid"
ORACLE=0 NODES=1500 N_SCN=1000 python tools/time_n1.py 2>&1 | head -1
ORACLE=0 ASYM=1 NODES=1500 N_SCN=1000 python tools/time_n1.py 2>&1 | head -1
ORACLE=0 NODES=50000 N_SCN=1024 python tools/time_n1.py 2>&1 | head -1
ORACLE=0 ASYM=1 NODES=50000 N_SCN=512 python tools/time_n1.py 2>&1 | head -1
echo "== config 5 shape against the CPU port (16 host threads), 256 scenarios on the 1804-bus grid"
NODES=1500 N_SCN=256 python tools/time_n1.py 2>&1 | head -3
ASYM=1 NODES=1500 N_SCN=256 python tools/time_n1.py 2>&1 | head -3
echo "== drop-in: config 2 through the reference's unchanged Python wrapper / PGM_calculate (pageable buffers) vs page-locked reused buffers"
python tools/time_reference_wrapper.py 2>&1 | tail -4
