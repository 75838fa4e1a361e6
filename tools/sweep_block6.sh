#!/bin/bash
# row-split asymmetric kernel: thread counts, phases, against the one-thread-per-row kernel (BASELINE configs[2])
for thr in 512 768 1024; do
  echo "== PGMB_BLOCK6_THREADS=$thr"
  PGMB_BLOCK6_THREADS=$thr python tools/check_block6.py 2>&1 | tail -4
done
echo "== phases (768 threads)"
PGMB_BLOCK6=1 PGMB_BLOCK6_THREADS=768 PGMB_DEBUG_PHASES=1 ASYM=1 RINGS=1 REPS=2 python tools/profile_solve.py 2>&1 | tail -4
echo "== radial grid, asym"
RINGS=0 PGMB_BLOCK6_THREADS=768 python tools/check_block6.py 2>&1 | tail -4
