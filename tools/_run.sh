for nw in 4 5 6; do FOV360_SAT_WARPS=$nw timeout 120 python tools/stage_bench.py --batch 16 --tag NW$nw | grep "onepass"; done
for v in 1 3; do FOV360_SAT_VARIANT=$v FOV360_SAT_TMA_STORE=0 timeout 120 python tools/stage_bench.py --batch 16 --tag var$v | grep "onepass"; done
FOV360_SAT_POLICY=1 timeout 120 python tools/stage_bench.py --batch 16 --tag pol1 | grep "onepass"
FOV360_SAT_POLICY=3 timeout 120 python tools/stage_bench.py --batch 16 --tag pol3 | grep "onepass"
