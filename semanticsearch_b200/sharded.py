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


class ShardedCorpus:
    """One rank's shard of a row-sharded corpus, resident in HBM."""

    def __init__(self, local_rows: torch.Tensor, row_offset: int, group: Optional[dist.ProcessGroup] = None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None):
        self.local_rows = local_rows
        self.row_offset = int(row_offset)
        self.group = group
        if local_search is None or merge is None:
            from . import similarity
            local_search = local_search or similarity.cosine_topk
            merge = merge or similarity.topk_merge
        self._local_search = local_search
        self._merge = merge

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def search(self, queries: torch.Tensor, k: int, algo: str = "auto"):
        """Global top-k for ``queries`` (replicated on every rank): ``(scores, indices)``."""
        scores, idx, keys = self._local_search(self.local_rows, queries, k, index_base=self.row_offset,
                                               return_keys=True, algo=algo)
        world = self.world_size
        if world == 1:
            return scores, idx
        b, kk = keys.shape
        gathered = torch.empty((world * b, kk), dtype=keys.dtype, device=keys.device)  # rank-major concatenation
        dist.all_gather_into_tensor(gathered, keys.contiguous(), group=self.group)
        scores, idx, _ = self._merge(gathered.view(world, b, kk), k)
        return scores, idx
