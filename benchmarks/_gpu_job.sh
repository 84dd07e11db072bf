timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_rank.py tests/test_gpu_ranker.py -x -q 2>&1 | tail -12
python benchmarks/bench_configs.py --config 1 --steps 20 | tail -n 1
