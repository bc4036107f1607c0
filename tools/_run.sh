timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/stage_bench.py --batch 16 --tag clean
./tools/sat_trace.bin 16
