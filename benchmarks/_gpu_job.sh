timeout 600 python -m pytest tests/test_gpu_simmatrix_tc.py -x -q -s 2>&1 | tail -25
timeout 600 python benchmarks/bench_configs.py --config 2 --steps 5 2>&1 | tail -3
