timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
timeout 900 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_v7.json; cut -c1-400 gpurun_out/bench_v7.json
