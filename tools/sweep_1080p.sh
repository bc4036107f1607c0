#!/bin/bash
# SAT tile geometry for one 1080p frame (135 tiles at the default 24-row bands: less than one per SM)
for r in 6 8 12 16 24; do
  for w in 0 2 3 5; do
    FOV360_SAT_BAND_ROWS=$r FOV360_SAT_WARPS=$w python tools/stage_bench.py --workload 1080p --batch 1 --steps 40 --tag "band=$r warps=$w" 2>&1 | grep -E "band=|sat_onepass"
  done
done
for r in 8 16 32; do
  FOV360_INTERP_ROWS=$r python tools/stage_bench.py --workload 1080p --batch 1 --steps 40 --tag "irows=$r" 2>&1 | grep -E "irows=|interpolate"
done
