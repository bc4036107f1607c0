for r in 16 24 32; do for nw in 5 6; do echo "R=$r NW=$nw"; FOV360_SAT_BAND_ROWS=$r FOV360_SAT_WARPS=$nw timeout 120 python tools/stage_bench.py --workload 4k --batch 8 --steps 50 | grep "onepass"; done; done
timeout 120 python tools/stage_bench.py --workload 4k --batch 8 --steps 50
timeout 120 python tools/stage_bench.py --workload 1080p --batch 16 --steps 50
