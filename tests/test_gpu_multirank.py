"""Multi-rank parity on real GPUs: with >= 2 devices visible, two ranks shard one corpus and every
exchange path — NCCL all-gather + merge, NVLink peer push + merge, and the CUDA-graph replay of the peer path — must
return exactly the single-GPU answer, bit for bit, including a tie between rows that live in different shards
(the lower global row index wins on any number of GPUs)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _corpus(n, d, dtype):
    g = torch.Generator().manual_seed(123)
    C = torch.randn((n, d), generator=g).to(dtype)
    # cross-shard ties: rows of the second shard duplicate rows of the first one (and one zero row)
    C[n // 2 + 17] = C[41]
    C[n // 2 + 999] = C[7]
    C[n - 1] = C[n // 2 - 1]
    C[100] = 0
    return C


def _queries(b, d, dtype, C):
    g = torch.Generator().manual_seed(321)
    Q = torch.randn((b, d), generator=g).to(dtype)
    Q[0] = C[41]      # its best match is the duplicated pair (41, n/2 + 17): index 41 must come first
    if b > 2:
        Q[1] = C[7]
        Q[2] = C[C.shape[0] // 2 - 1]
    return Q


def _worker(rank, world, port, n, d, cases, out_dir):
    import torch.distributed as dist
    from semanticsearch_b200 import sharded, similarity
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        for ci, (dtype, b, k) in enumerate(cases):
            C = _corpus(n, d, dtype)
            Q = _queries(b, d, dtype, C).to(dev)
            lo, hi = sharded.shard_bounds(n, world, rank)
            shard = C[lo:hi].to(dev).contiguous()
            peer = sharded.PeerExchange(dev, b, k)
            corpus = sharded.ShardedCorpus(shard, lo, peer_exchange=peer)
            res = {}
            res["nccl"] = corpus.search(Q, k, exchange="nccl")
            res["peer"] = corpus.search(Q, k, exchange="peer")
            gs = sharded.GraphedSearch(corpus, b, k)
            for rep in range(3):
                s, i = gs(Q)
                res[f"graph{rep}"] = (s.clone(), i.clone())
            torch.cuda.synchronize()
            assert peer.status() == 0
            if rank == 0:
                want_s, want_i = similarity.cosine_topk(C.to(dev), Q, k)     # the whole corpus on one GPU
                for name, (s, i) in res.items():
                    assert torch.equal(i, want_i), (ci, name, "indices differ from the single-GPU result")
                    assert torch.equal(s, want_s), (ci, name, "scores differ from the single-GPU result")
                assert int(want_i[0, 0]) == 41 and int(want_i[0, 1]) == n // 2 + 17      # the cross-shard tie, lower index first
                if b > 2:
                    assert int(want_i[1, 0]) == 7 and int(want_i[1, 1]) == n // 2 + 999
            # every rank holds the same answer
            mine = res["graph2"][1].contiguous()
            both = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(both, mine)
            assert all(torch.equal(x, mine) for x in both)
            del gs
            peer.close()
        if rank == 0:
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_search_equals_single_gpu_search(tmp_path):
    import torch.multiprocessing as mp
    cases = [(torch.bfloat16, 1, 10), (torch.float16, 16, 100), (torch.bfloat16, 256, 10), (torch.float32, 3, 10)]
    mp.spawn(_worker, args=(2, _free_port(), 400_000, 128, cases, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
