timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_rank.py tests/test_gpu_tcstream.py -x -q 2>&1 | tail -5
for i in 1 2; do
SS_GEMM_NO_SHARE=1 timeout 300 python benchmarks/sweep_topk.py --batches 4096 --algos gemm --steps 8 2>&1 | tail -n 1
timeout 300 python benchmarks/sweep_topk.py --batches 4096 --algos gemm --steps 8 2>&1 | tail -n 1
done
