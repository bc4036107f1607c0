set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1 2; do FOV360_SAT_VARIANT=$v python tools/stage_bench.py --tag var$v; done
FOV360_SAT_TMA_STORE=1 python tools/stage_bench.py --tag tma
FOV360_SAT_DEBUG_NOWAIT=1 python tools/stage_bench.py --tag nowait
FOV360_SAT_BAND_ROWS=64 python tools/stage_bench.py --tag band64
FOV360_SAT_BAND_ROWS=16 python tools/stage_bench.py --tag band16
python tools/stage_bench.py --workload 4k --tag 4k
