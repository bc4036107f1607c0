#!/bin/bash
# Tile-height sweep of the log-polar inverse warp (rows per CTA, FOV360_LP_ROWS).  Run on a GPU box:
#   tools/sweep_tile_rows.sh > gpurun_out/tile_rows.log
# (profiles/r02_tile_rows.txt also holds the interpolate_rect part of the first sweep, taken with an
# experimental build whose <32> instantiation accepted any tile height at run time: heights between
# the instantiated 8 / 16 / 32 were never faster, so that build was dropped.)
cd "$(dirname "$0")/.."
for r in 0 4 6 8 10 12 16 20 24 32 40 50 64; do
  echo "-- FOV360_LP_ROWS=$r (0 = the library's own choice)"
  FOV360_LP_ROWS=$r python tools/logpolar_stats.py --sizes "" --time 1080p,4k,8k 2>&1 | grep interpolate
done
