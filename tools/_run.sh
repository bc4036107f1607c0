timeout 900 python -m pytest tests -m gpu -x -q -k "image_sampler or golden or dropin" 2>&1 | tail -3
timeout 200 python tools/logpolar_bench.py --workload 4k 2>/dev/null | head -1
