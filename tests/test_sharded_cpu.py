"""Host-side logic of the multi-GPU path on CPU: shard bounds, document partitioning, and the
all-gather + merge plumbing of ShardedCorpus with world_size 2 over gloo.  The CUDA kernels are
replaced by injected oracle-based callables (the product defaults are the CUDA ops)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import keys as okeys
from oracle import rank_oracle as ro
from semanticsearch_b200.sharded import ShardedCorpus, partition_documents, shard_bounds


def test_shard_bounds_cover_rows_exactly():
    for n, w in ((10_000_000, 8), (10, 3), (7, 8), (1, 1), (100, 4)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_partition_documents_balances_cost():
    rng = np.random.default_rng(0)
    sizes = rng.integers(16, 513, size=2000)
    for power in (1, 2):
        parts = partition_documents(sizes, 8, power=power)
        assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
        loads = [float((sizes[p].astype(float) ** power).sum()) for p in parts]
        assert max(loads) / min(loads) < 1.01


def _oracle_local_search(local_rows, queries, k, index_base=0, return_keys=True, algo="auto"):
    s, i = ro.cosine_topk_ref(queries.numpy(), local_rows.numpy(), k)
    pad = k - s.shape[1]
    if pad:
        s = np.concatenate([s, np.full((s.shape[0], pad), -np.inf, np.float32)], axis=1)
        i = np.concatenate([i, np.full((i.shape[0], pad), -1, np.int64)], axis=1)
    gi = np.where(i >= 0, i + index_base, 0)
    keys = okeys.pack_keys(s, gi)
    keys = np.where(i >= 0, keys, 0)
    return torch.from_numpy(s), torch.from_numpy(np.where(i >= 0, i + index_base, -1)), torch.from_numpy(keys)


def _oracle_merge(gathered, k):
    merged = okeys.merge_keys(gathered.numpy(), k)
    s, i = okeys.unpack_keys(merged)
    return torch.from_numpy(s), torch.from_numpy(i), torch.from_numpy(merged)


def _worker(rank, world, port, n, d, b, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    C = rng.standard_normal((n, d)).astype(np.float32)
    C[n - 1] = C[1]  # a tie that straddles the two shards
    Q = rng.standard_normal((b, d)).astype(np.float32)
    lo, hi = shard_bounds(n, world, rank)
    corpus = ShardedCorpus(torch.from_numpy(C[lo:hi]), lo, local_search=_oracle_local_search, merge=_oracle_merge)
    s, i = corpus.search(torch.from_numpy(Q), k)
    ref_s, ref_i = ro.cosine_topk_ref(Q, C, k)
    ok = np.array_equal(i.numpy(), ref_i) and np.allclose(s.numpy(), ref_s, atol=1e-6)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_world2_gloo():
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, 501, 32, 3, 10, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_key_packing_roundtrip_and_order():
    s = np.array([[0.5, -0.25, 0.0, -0.0, 1.0, -1.0, np.inf, -np.inf]], dtype=np.float32)
    i = np.arange(8)[None, :]
    k = okeys.pack_keys(s, i)
    s2, i2 = okeys.unpack_keys(k)
    assert np.array_equal(s2.view(np.uint32), s.view(np.uint32)) and np.array_equal(i2, i)
    order = np.argsort(k.view(np.uint64)[0])[::-1]
    assert list(s[0][order]) == sorted(s[0].tolist(), reverse=True)
    # equal scores: the lower index is the larger key
    k2 = okeys.pack_keys(np.array([0.3, 0.3], np.float32), np.array([7, 3]))
    assert k2.view(np.uint64)[1] > k2.view(np.uint64)[0]
