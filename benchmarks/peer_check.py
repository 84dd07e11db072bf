#!/usr/bin/env python
"""torchrun check: the NVLink peer-memory exchange returns exactly what NCCL all-gather + merge returns.
    python -m torch.distributed.run --nproc-per-node N benchmarks/peer_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticsearch_b200.sharded import PeerExchange, ShardedCorpus, shard_bounds  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n, d = 400_000, 384
g = torch.Generator(device=dev).manual_seed(1)          # same corpus on every rank, each keeps its rows
C = torch.randn((n, d), generator=g, device=dev).to(torch.float16)
lo, hi = shard_bounds(n, world, rank)
ok = True
for b, k in ((1, 10), (16, 100), (300, 10)):
    Q = torch.randn((b, d), generator=g, device=dev).to(torch.float16)
    ref = ShardedCorpus(C[lo:hi].contiguous(), lo)
    peer = PeerExchange(dev, max_queries=b, k=k)
    fused = ShardedCorpus(C[lo:hi].contiguous(), lo, peer_exchange=peer)
    for it in range(5):   # several searches: both parities of the double buffer, increasing seq
        s0, i0 = ref.search(Q, k)
        s1, i1 = fused.search(Q, k)
        torch.cuda.synchronize()
        same = torch.equal(i0, i1) and torch.equal(s0, s1)
        ok = ok and same
    # against a single-GPU search of the whole corpus
    from semanticsearch_b200 import similarity
    sf, jf = similarity.cosine_topk(C, Q, k)
    ok = ok and torch.equal(jf, i1)
    peer.close()
    if rank == 0:
        print(f"B={b} k={k}: peer exchange == NCCL path == single-GPU search: {ok}", flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
