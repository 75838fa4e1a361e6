#!/bin/bash
# tile-width / slot sweep of the symmetric path kernel on BASELINE configs[1] (1000 scenarios)
for cfg in "8 64" "4 64" "4 128" "8 32" "4 32" "16 32" "16 16"; do
  set -- $cfg
  PGMB_TILE=$1 PGMB_SLOTS=$2 REPS=5 python tools/profile_solve.py 2>&1 | tail -1
done
