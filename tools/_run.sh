timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err
timeout 900 ncu --set full --clock-control none --import-source on -c 3 -o gpurun_out/prof_r01_v8_all -f python tools/profile_step.py --batch 16 --steps 1 > gpurun_out/ncu_v8.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v8.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench_v8.log 2>&1
