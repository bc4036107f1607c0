timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python tools/stage_bench.py --tag hint | grep "fps\|sample"
FOV360_SAMPLE_NO_SRC=1 timeout 120 python tools/stage_bench.py --tag nohint| grep "fps\|sample"
for v in s5 s6; do FOV360_LIB=tools/variants/libfov360_$v.so timeout 120 python tools/stage_bench.py --tag $v | grep "fps\|sample"; FOV360_SAMPLE_NO_SRC=1 FOV360_LIB=tools/variants/libfov360_$v.so timeout 120 python tools/stage_bench.py --tag ${v}nohint | grep "fps\|sample"; done
