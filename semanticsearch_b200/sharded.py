"""Corpus sharding across the GPUs of one box.

Each rank owns a contiguous block of corpus rows.  A search is: local fused cosine/top-k on the
rank's shard (global row indices baked into the packed keys) -> NCCL all-gather of ``B x k``
keys per rank (80 B ... 328 KB, latency-bound) -> the same k-way merge kernel on every rank.
Ties resolve to the lower global row index, so the result is independent of the GPU count.
The reference is single-process (no collective anywhere); this is the distributed form of
Tool/rank_chunks_optimized.py:215-216,225.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block ``[lo, hi)`` of ``rank`` (first ``n_rows % world`` ranks get one extra)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def partition_documents(sizes, world_size: int, power: int = 1):
    """Static LPT partition of ragged documents by cost ``n**power`` (power=2 for the n x n
    similarity pass, 1 for the adjacent-similarity pass).  Returns a list of index lists; no
    cross-GPU traffic is ever needed for ragged batches."""
    import numpy as np
    sizes = np.asarray(sizes, dtype=np.int64)
    cost = sizes.astype(np.float64) ** power
    order = np.argsort(-cost, kind="stable")
    loads = [0.0] * world_size
    parts = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        parts[r].append(int(i))
        loads[r] += float(cost[i])
    return [sorted(p) for p in parts]


class PeerExchange:
    """Top-k key exchange over NVLink peer memory (CUDA IPC) fused with the merge — the latency-bound
    replacement of "NCCL all-gather + merge" for the ranks of one box (csrc/peer_exchange.cu)."""

    def __init__(self, device: torch.device, max_queries: int, k: int, group: Optional[dist.ProcessGroup] = None):
        import ctypes

        from . import _lib
        self.lib = _lib.load()
        self._check = _lib.check
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.max_queries, self.k = int(max_queries), int(k)
        self.device = device
        nbytes = self.lib.ss_peer_buffer_bytes(self.world, self.max_queries, self.k)
        if nbytes == 0:
            raise ValueError("unsupported world size / max_queries / k for the peer exchange")
        with torch.cuda.device(device):
            ptr = ctypes.c_void_p()
            handle = (ctypes.c_ubyte * 64)()
            self._check(self.lib.ss_peer_alloc(nbytes, ctypes.byref(ptr), handle), "ss_peer_alloc")
            self.local_ptr = ptr.value
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self.peer_ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.peer_ptrs.append(self.local_ptr)
                    continue
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                out = ctypes.c_void_p()
                self._check(self.lib.ss_peer_open(buf, ctypes.byref(out)), "ss_peer_open")
                self.peer_ptrs.append(out.value)
            self.bases = torch.tensor(self.peer_ptrs, dtype=torch.int64, device=device)
        dist.barrier(group=group)  # every buffer is mapped everywhere before the first push

    def exchange_merge(self, keys: torch.Tensor):
        """``keys``: this rank's int64 ``[B, k]`` packed keys -> global ``(scores, indices)`` on every rank.  The sequence
        number of the exchange lives on the device, so the call can be captured in a CUDA graph and replayed (every rank
        must issue the same number of exchanges, in the same order)."""
        b, k = keys.shape
        if k != self.k or b > self.max_queries:
            raise ValueError(f"exchange was sized for k={self.k}, <= {self.max_queries} queries")
        dev = keys.device
        with torch.cuda.device(dev):
            scores = torch.empty((b, k), dtype=torch.float32, device=dev)
            idx = torch.empty((b, k), dtype=torch.int64, device=dev)
            st = self.lib.ss_topk_peer_exchange_merge_auto(keys.data_ptr(), b, k, self.rank, self.world, self.bases.data_ptr(),
                                                           self.max_queries, self.local_ptr, None, scores.data_ptr(),
                                                           idx.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            self._check(st, "ss_topk_peer_exchange_merge_auto")
        return scores, idx

    def status(self) -> int:
        """0 = healthy; r + 1 = a merge gave up waiting for rank r (its result must be discarded).  Synchronises."""
        import ctypes
        out = ctypes.c_int()
        with torch.cuda.device(self.device):
            self._check(self.lib.ss_peer_status(self.local_ptr, ctypes.byref(out)), "ss_peer_status")
        return int(out.value)

    def close(self):
        torch.cuda.synchronize(self.device)
        bad = self.status()
        dist.barrier(group=self.group)  # nobody unmaps while a peer may still push
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank and p:
                self.lib.ss_peer_close(p)
        self.lib.ss_peer_free(self.local_ptr)
        self.peer_ptrs = []
        if bad:
            raise RuntimeError(f"peer exchange: a merge timed out waiting for rank {bad - 1}")


class ShardedCorpus:
    """One rank's shard of a row-sharded corpus, resident in HBM."""

    def __init__(self, local_rows: torch.Tensor, row_offset: int, group: Optional[dist.ProcessGroup] = None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 peer_exchange: Optional["PeerExchange"] = None):
        self.local_rows = local_rows
        self.row_offset = int(row_offset)
        self.group = group
        self.peer_exchange = peer_exchange
        if local_search is None or merge is None:
            from . import similarity
            local_search = local_search or similarity.cosine_topk
            merge = merge or similarity.topk_merge
        self._local_search = local_search
        self._merge = merge
        from . import similarity as _sim
        # the shard is resident and immutable for the life of this object: K2 may keep its row norms between searches
        self._resident = _sim.ResidentIndex() if local_search is _sim.cosine_topk else None

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def search(self, queries: torch.Tensor, k: int, algo: str = "auto", exchange: str = "auto", resident=None):
        """Global top-k for ``queries`` (replicated on every rank): ``(scores, indices)``.  ``exchange``: "nccl" = all-gather
        of the per-rank keys + merge kernel, "peer" = push over NVLink peer memory fused with the merge (needs a
        ``PeerExchange``), "auto" = peer when one was given and fits, else NCCL.  Both return the same bits."""
        token = resident if resident is not None else self._resident
        extra = {"resident": token} if token is not None else {}
        scores, idx, keys = self._local_search(self.local_rows, queries, k, index_base=self.row_offset,
                                               return_keys=True, algo=algo, **extra)
        world = self.world_size
        if world == 1:
            return scores, idx
        b, kk = keys.shape
        fits = self.peer_exchange is not None and kk == self.peer_exchange.k and b <= self.peer_exchange.max_queries
        if exchange == "peer" and not fits:
            raise ValueError("no PeerExchange of this size was attached to the corpus")
        if exchange != "nccl" and fits:
            return self.peer_exchange.exchange_merge(keys.contiguous())
        gathered = torch.empty((world * b, kk), dtype=keys.dtype, device=keys.device)  # rank-major concatenation
        dist.all_gather_into_tensor(gathered, keys.contiguous(), group=self.group)
        scores, idx, _ = self._merge(gathered.view(world, b, kk), k)
        return scores, idx


class GraphedSearch:
    """``ShardedCorpus.search`` for one fixed (batch, dim, dtype, k) captured in a CUDA graph: the kernels of a search
    replay as ONE graph launch, which removes the per-call Python / ctypes / launch latency (tens of microseconds: 7 % of
    a single-query step, 10 % of a config-5 step).  On several GPUs the exchange is the kernel-only NVLink peer path (the
    corpus needs a ``PeerExchange``; an NCCL all-gather inside the captured region hung at 2 ranks with torch 2.11 / NCCL
    2.28) and every rank must construct and call its GraphedSearch in lock-step.  Queries are copied into a static
    input buffer; the returned tensors are static outputs that the next call overwrites.  The graph owns its K2
    workspace (through the corpus's ResidentIndex), so eager searches elsewhere in the process cannot invalidate it."""

    def __init__(self, corpus: ShardedCorpus, batch: int, k: int, algo: str = "auto", warmup: int = 3):
        world = corpus.world_size
        if world > 1 and (corpus.peer_exchange is None or corpus.peer_exchange.k != int(k) or batch > corpus.peer_exchange.max_queries):
            raise RuntimeError("GraphedSearch on several GPUs needs a PeerExchange sized for this batch and k")
        rows = corpus.local_rows
        self.corpus, self.k, self.algo = corpus, int(k), algo
        self.exchange = "peer" if world > 1 else "auto"
        self.q = torch.zeros((batch, rows.shape[1]), dtype=rows.dtype, device=rows.device)
        from . import similarity as _sim
        self._resident = _sim.ResidentIndex() if corpus._resident is not None else None   # the graph's own workspaces
        side = torch.cuda.Stream(device=rows.device)
        side.wait_stream(torch.cuda.current_stream(rows.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):  # sizes every workspace before capture
                corpus.search(self.q, self.k, algo=algo, exchange=self.exchange, resident=self._resident)
        torch.cuda.current_stream(rows.device).wait_stream(side)
        torch.cuda.synchronize(rows.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.scores, self.indices = corpus.search(self.q, self.k, algo=algo, exchange=self.exchange, resident=self._resident)
        if self._resident is not None:
            self._resident.frozen = True

    def __call__(self, queries: torch.Tensor):
        self.q.copy_(queries, non_blocking=True)
        self.graph.replay()
        return self.scores, self.indices
