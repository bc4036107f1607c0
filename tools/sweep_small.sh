#!/bin/bash
# band-height / tile-height sweep for small work (one process per setting: the env is read once)
for cfg in "4k 1" "4k 8" "8k 1"; do
  set -- $cfg
  for r in 8 12 16 24 32 48; do
    FOV360_SAT_BAND_ROWS=$r python tools/stage_bench.py --workload $1 --batch $2 --steps 30 --tag "band=$r" 2>&1 | grep -E "band=|sat_onepass"
  done
  for r in 8 16 32; do
    FOV360_INTERP_ROWS=$r python tools/stage_bench.py --workload $1 --batch $2 --steps 30 --tag "irows=$r" 2>&1 | grep -E "irows=|interpolate"
  done
done
