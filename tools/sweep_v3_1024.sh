#!/bin/bash
# experiment: nr_sym_v3 compiled with -DV3_THREADS=1024 (64 registers): more slots per tile / two blocks per SM
for cfg in "4 128" "8 128" "8 64" "4 64" "4 256"; do
  set -- $cfg
  PGMB_TILE=$1 PGMB_SLOTS=$2 REPS=5 python tools/profile_solve.py 2>&1 | tail -1
done
