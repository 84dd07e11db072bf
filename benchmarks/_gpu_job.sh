# B=1 under bench conditions: K1 (CUDA cores) vs K7 (tensor cores), cold and right after a hot GEMM run
python bench.py --batch 1 --algo stream --steps 100 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cold stream', d['value'], d['roofline']['frac'], d['clocks'])"
python bench.py --batch 1 --algo tcstream --steps 100 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cold tcstream', d['value'], d['roofline']['frac'], d['clocks'])"
python - <<'PY'
import torch, sys, os, time
sys.path.insert(0, os.getcwd())
from semanticsearch_b200 import similarity
g = torch.Generator(device="cuda").manual_seed(1)
C = torch.empty((10_000_000, 768), dtype=torch.bfloat16, device="cuda")
for a in range(0, 10_000_000, 1 << 20):
    e = min(10_000_000, a + (1 << 20)); C[a:e] = torch.randn((e - a, 768), generator=g, device="cuda").to(torch.bfloat16)
Q = torch.randn((4096, 768), generator=g, device="cuda").to(torch.bfloat16)
q1 = Q[:1].contiguous()
def t(fn, n):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for rnd in range(2):
    print("gemm hot ms", t(lambda: similarity.cosine_topk(C, Q, 10), 30))
    print(" hot stream   B=1 ms", t(lambda: similarity.cosine_topk(C, q1, 10, algo="stream"), 100))
    print("gemm hot ms", t(lambda: similarity.cosine_topk(C, Q, 10), 30))
    print(" hot tcstream B=1 ms", t(lambda: similarity.cosine_topk(C, q1, 10, algo="tcstream"), 100))
PY
