timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -5
timeout 300 python benchmarks/sweep_topk.py --batches 128,1024,4096 --algos gemm --steps 5 2>&1 | tee gpurun_out/sweep_gemm.jsonl
