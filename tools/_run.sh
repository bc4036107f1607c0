timeout 120 python tools/stage_bench.py --tag b4 | grep "fps\|onepass"
for v in b2 b3 b6; do FOV360_LIB=tools/variants/libfov360_$v.so timeout 120 python tools/stage_bench.py --tag $v | grep "onepass"; done
FOV360_SAT_BAND_ROWS=64 timeout 120 python tools/stage_bench.py --tag band64 | grep "onepass"
FOV360_SAT_VARIANT=1 timeout 120 python tools/stage_bench.py --tag var1 | grep "onepass"
