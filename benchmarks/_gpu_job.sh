timeout 600 python -m pytest tests/test_gpu_ragged.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -15
timeout 600 python benchmarks/bench_configs.py --config 2 --steps 5 2>&1 | tail -3
