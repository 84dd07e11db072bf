timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -5
for i in 1 2; do timeout 300 python benchmarks/sweep_topk.py --batches 4096 --algos gemm --steps 8 2>&1 | tail -n 1; done
timeout 300 python benchmarks/sweep_topk.py --rows 1250000 --batches 4096 --algos gemm --steps 10 2>&1 | tail -n 1
timeout 300 python benchmarks/sweep_topk.py --batches 128,1024 --algos gemm --steps 5 2>&1 | tail -n 2
