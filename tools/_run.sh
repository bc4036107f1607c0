timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python tools/stage_bench.py --tag base
for v in c8 m5 c2; do FOV360_LIB=tools/variants/libfov360_$v.so timeout 120 python tools/stage_bench.py --tag $v | grep -v "onepass\|sample"; done
