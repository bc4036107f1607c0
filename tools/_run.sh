set -x
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench_v5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -c 3 -o gpurun_out/prof_r01_v5_all -f python tools/profile_step.py --batch 8 --steps 1 > gpurun_out/ncu_v5.log 2>&1
