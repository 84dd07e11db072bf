python bench.py --steps 10 --no-cpu-baseline 2> gpurun_out/bench_g1.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['extra']['query_batch_1']
print('N1 B4096 value',round(d['value']),'e2e',round(d['e2e']['value']),d['e2e'].get('api'))
print('N1 B1 value',round(e['value'],1),'e2e',round(e['e2e']['value'],1), e['e2e'].get('api'))"
tail -n 3 gpurun_out/bench_g1.err
python bench.py --rows 12500000 --dim 384 --k 100 --batch 16 --dtype fp16 --no-secondary --no-cpu-baseline --steps 50 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5 shard value',round(d['value']),'e2e',round(d['e2e']['value']),d['e2e'].get('api'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 2> gpurun_out/bench_g2.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['extra']['query_batch_1']
print('N2 B4096 value',round(d['value']),'e2e',round(d['e2e']['value']),d['e2e'].get('api'))
print('N2 B1 value',round(e['value'],1),'e2e',round(e['e2e']['value'],1), e['e2e'].get('api'))"
grep -v OMP gpurun_out/bench_g2.err | tail -n 3
