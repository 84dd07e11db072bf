timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 benchmarks/peer_check.py 2>&1 | grep -v "OMP\|\*\*\*" | tail -4
for ex in nccl peer; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 8 --batch 1 --steps 200 --no-cpu-baseline --exchange $ex 2> gpurun_out/bench_p8_$ex.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$ex N8 B1 value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'kernel_ms',round(d['roofline']['kernel_ms'],4))"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 8 --rows 100000000 --dim 384 --k 100 --batch 16 --dtype fp16 --no-secondary --no-cpu-baseline --steps 100 --exchange $ex 2>> gpurun_out/bench_p8_$ex.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$ex N8 cfg5 value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'kernel_ms',round(d['roofline']['kernel_ms'],4))"
done
