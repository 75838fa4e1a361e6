#!/bin/bash
# tile-width / slot sweep of the asymmetric block kernel on BASELINE configs[2] (ringed grid, 1000 scenarios)
for cfg in "8 32" "4 32" "4 64" "4 16" "8 16" "16 16" "2 64"; do
  set -- $cfg
  ASYM=1 RINGS=1 PGMB_TILE=$1 PGMB_SLOTS=$2 REPS=3 python tools/profile_solve.py 2>&1 | tail -1
done
