"""Host-side operators over the C ABI: torch is used for device memory, streams and nothing else.

Every function here takes CUDA tensors, passes raw device pointers + the current stream to
``libsemsearch_b200.so`` and returns CUDA tensors.  No function has a CPU code path.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.SS_F32, torch.bfloat16: _lib.SS_BF16, torch.float16: _lib.SS_F16}

_workspaces: Dict[Tuple[int, int, str], torch.Tensor] = {}


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; expected float32, bfloat16 or float16") from None


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError("semanticsearch_b200 operators require CUDA tensors (there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("all tensors must live on the same CUDA device")
    return dev


def _stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def workspace(dev: torch.device, nbytes: int, tag: str = "default") -> torch.Tensor:
    """Grow-only scratch buffer (the caller-provided workspace of the C ABI), one per (device, STREAM, tag): kernels of
    different streams never share ticket counters or partial lists, and a buffer is only ever replaced by a call on the
    stream that uses it (stream order protects the old contents).  Searches that must keep a buffer alive across calls
    (a CUDA graph, a resident index) own theirs through a ``ResidentIndex`` instead."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, int(torch.cuda.current_stream(dev).cuda_stream), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
        _workspaces[key] = buf
    return buf


# Dispatch thresholds for 16-bit corpora (measured on B200, profiles/):
#   queries <  TC_MIN_BATCH              -> K1 CUDA-core streaming kernel (HBM-bound up to a few queries)
#   TC_MIN_BATCH <= queries < GEMM_MIN_BATCH, or k > 16 -> K7 tensor-core streaming kernel (HBM-bound)
#   queries >= GEMM_MIN_BATCH and k <= 16 -> K2 tensor-core GEMM kernel (tensor-bound)
TC_MIN_BATCH = 2
GEMM_MIN_BATCH = 96
_ALGOS = ("auto", "stream", "gemm", "tcstream", "small")
# K9 (score matrix in the workspace + per-query selection) takes over from K1 when several query groups
# would each re-read a small corpus: config 1 (100 x 10k x 384 fp32) runs in 0.43 ms through K1
SMALL_MIN_BATCH = 9
SMALL_MAX_ROWS = 1 << 17
SMALL_MAX_SCORES = 1 << 26
_TCSTREAM_WS_BUDGET = 1 << 30  # bytes of partial top-k lists per K7 call


def _lowp_eligible(corpus: torch.Tensor, queries: torch.Tensor) -> bool:
    return (corpus.dtype in (torch.bfloat16, torch.float16) and queries.dtype == corpus.dtype
            and corpus.shape[1] % 8 == 0 and corpus.data_ptr() % 16 == 0 and queries.data_ptr() % 16 == 0
            and corpus.shape[0] < 2 ** 31 - 128)


def _gemm_eligible(corpus: torch.Tensor, queries: torch.Tensor, k: int) -> bool:
    return _lowp_eligible(corpus, queries) and k <= 16


def _tcstream_eligible(corpus: torch.Tensor, queries: torch.Tensor, k: int) -> bool:
    return _lowp_eligible(corpus, queries) and k <= 1024


def choose_algo(corpus: torch.Tensor, queries: torch.Tensor, k: int) -> str:
    b = queries.shape[0]
    if b >= GEMM_MIN_BATCH and _gemm_eligible(corpus, queries, k):
        return "gemm"
    if b >= TC_MIN_BATCH and _tcstream_eligible(corpus, queries, k):
        return "tcstream"
    n = corpus.shape[0]
    if b >= SMALL_MIN_BATCH and n <= SMALL_MAX_ROWS and b * n <= SMALL_MAX_SCORES and k <= 4096:
        return "small"
    return "stream"


class ResidentIndex:
    """Token of a corpus that stays resident and unchanged in HBM (``sharded.ShardedCorpus`` owns one, and so does every
    ``sharded.GraphedSearch``).  It carries its own workspaces — nobody else touches them, so a CUDA graph captured from a
    search with this token stays valid whatever other searches the process runs — and the K2 workspace keeps the corpus's
    inverse row norms at its head: while buffer and corpus (pointer, shape, dtype) are those of the token's previous K2
    call, the next call skips the norm pre-pass (one read of the whole corpus)."""

    __slots__ = ("bufs", "state", "frozen")

    def __init__(self):
        self.bufs: Dict[str, torch.Tensor] = {}   # grow-only uint8 workspaces by tag
        self.state = None    # (workspace data_ptr, corpus data_ptr, rows, dim, dtype) of the last K2 call
        self.frozen = False  # set once a CUDA graph has baked the buffer addresses in

    @property
    def ws(self):
        return self.bufs.get("gemm")

    def workspace(self, dev: torch.device, nbytes: int, tag: str = "gemm") -> torch.Tensor:
        buf = self.bufs.get(tag)
        if buf is None or buf.numel() < nbytes or buf.device != dev:
            if self.frozen:
                raise RuntimeError("this index's workspaces are referenced by a captured CUDA graph and cannot grow; "
                                   "search with the batch size and k the graph was built for")
            buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
            self.bufs[tag] = buf
            if tag == "gemm":
                self.state = None
        return buf

    def invalidate(self) -> None:
        """Call after the corpus bytes change."""
        self.state = None


def cosine_topk(corpus: torch.Tensor, queries: torch.Tensor, k: int, *, index_base: int = 0,
                return_keys: bool = False, algo: str = "auto", resident: "ResidentIndex | None" = None):
    """Top-k cosine similarity of each query row against every corpus row.

    Device form of ``cosine_similarity(q, C)[0]`` + ``np.argsort(-s)[:k]``
    (Tool/rank_chunks_optimized.py:215-216,225).  Returns ``(scores fp32 [B,k], indices int64
    [B,k])`` best-first, ties resolved to the lower index; with ``return_keys`` also the packed
    int64 keys used for multi-GPU merging.  ``resident``: token of a corpus that stays unchanged in HBM, see
    ``ResidentIndex``.
    """
    dev = _require_cuda(corpus, queries)
    if corpus.dim() != 2 or queries.dim() != 2 or corpus.shape[1] != queries.shape[1]:
        raise ValueError(f"shape mismatch: corpus {tuple(corpus.shape)} vs queries {tuple(queries.shape)}")
    if not corpus.is_contiguous() or not queries.is_contiguous():
        raise ValueError("corpus and queries must be contiguous row-major tensors")
    n, d = corpus.shape
    b = queries.shape[0]
    k = int(k)
    if k <= 0 or n <= 0 or b <= 0:
        raise ValueError("k, corpus rows and query rows must be positive")
    if algo not in _ALGOS:
        raise ValueError(f"algo must be one of {_ALGOS}")
    if algo == "auto":
        algo = choose_algo(corpus, queries, k)
    if algo == "gemm" and not _gemm_eligible(corpus, queries, k):
        raise ValueError("the tensor-core GEMM path needs bf16/fp16 corpus and queries of one dtype, k <= 16, dim % 8 == 0")
    if algo == "tcstream" and not _tcstream_eligible(corpus, queries, k):
        raise ValueError("the tensor-core streaming path needs bf16/fp16 corpus and queries of one dtype, k <= 1024, dim % 8 == 0")
    if k > n:
        # fewer rows than requested: search with k = n and pad the tail (score -inf, index -1, key 0 = the empty slot of
        # the merge kernels), whichever kernel ran
        res = cosine_topk(corpus, queries, n, index_base=index_base, return_keys=True, algo=algo, resident=resident)
        pad = k - n
        scores = torch.cat([res[0], torch.full((b, pad), float("-inf"), dtype=torch.float32, device=dev)], dim=1)
        idx = torch.cat([res[1], torch.full((b, pad), -1, dtype=torch.int64, device=dev)], dim=1)
        keys = torch.cat([res[2], torch.zeros((b, pad), dtype=torch.int64, device=dev)], dim=1)
        return (scores, idx, keys) if return_keys else (scores, idx)
    lib = _lib.load()
    with torch.cuda.device(dev):
        scores = torch.empty((b, k), dtype=torch.float32, device=dev)
        idx = torch.empty((b, k), dtype=torch.int64, device=dev)
        keys = torch.empty((b, k), dtype=torch.int64, device=dev) if return_keys else None
        keys_ptr = keys.data_ptr() if keys is not None else None
        if algo == "gemm":
            need = lib.ss_cosine_topk_gemm_workspace_bytes(n, d, b, k)
            ws = resident.workspace(dev, need, "gemm") if resident is not None else workspace(dev, need, "gemm")
            state = (ws.data_ptr(), corpus.data_ptr(), n, d, corpus.dtype)
            valid = resident is not None and resident.state == state
            st = lib.ss_cosine_topk_gemm_resident(
                corpus.data_ptr(), n, d, _dtype_code(corpus), queries.data_ptr(), b, k, int(index_base), ws.data_ptr(),
                ws.numel(), int(valid), keys_ptr, scores.data_ptr(), idx.data_ptr(), _stream_ptr(dev))
            _lib.check(st, "ss_cosine_topk_gemm_resident")
            if resident is not None:
                resident.state = state
        elif algo == "tcstream":
            # per-CTA partial lists cost ~148 * k * 8 bytes per query: very large batches go through in slices
            step = max(64, min(b, (_TCSTREAM_WS_BUDGET // (160 * k * 8)) // 64 * 64))
            for q0 in range(0, b, step):
                q1 = min(b, q0 + step)
                need = lib.ss_cosine_topk_tcstream_workspace_bytes(n, d, q1 - q0, k)
                ws = resident.workspace(dev, need, "tcstream") if resident is not None else workspace(dev, need)
                st = lib.ss_cosine_topk_tcstream(
                    corpus.data_ptr(), n, d, _dtype_code(corpus), queries[q0:q1].data_ptr(), q1 - q0, k, int(index_base),
                    ws.data_ptr(), ws.numel(), keys[q0:q1].data_ptr() if keys is not None else None,
                    scores[q0:q1].data_ptr(), idx[q0:q1].data_ptr(), _stream_ptr(dev))
                _lib.check(st, "ss_cosine_topk_tcstream")
        elif algo == "small":
            need = lib.ss_cosine_topk_small_workspace_bytes(n, b)
            ws = resident.workspace(dev, need, "small") if resident is not None else workspace(dev, need)
            st = lib.ss_cosine_topk_small(
                corpus.data_ptr(), n, d, _dtype_code(corpus), queries.data_ptr(), b, _dtype_code(queries), k,
                int(index_base), ws.data_ptr(), ws.numel(), keys_ptr, scores.data_ptr(), idx.data_ptr(), _stream_ptr(dev))
            _lib.check(st, "ss_cosine_topk_small")
        else:
            need = lib.ss_cosine_topk_stream_workspace_bytes(n, d, _dtype_code(corpus), b, k)
            ws = resident.workspace(dev, need, "stream") if resident is not None else workspace(dev, need)
            st = lib.ss_cosine_topk_stream(
                corpus.data_ptr(), n, d, _dtype_code(corpus), queries.data_ptr(), b, _dtype_code(queries), k,
                int(index_base), ws.data_ptr(), ws.numel(), keys_ptr, scores.data_ptr(), idx.data_ptr(), _stream_ptr(dev))
            _lib.check(st, "ss_cosine_topk_stream")
    if return_keys:
        return scores, idx, keys
    return scores, idx


def cosine_scores(corpus: torch.Tensor, queries: torch.Tensor) -> torch.Tensor:
    """Full ``[B, N]`` float32 cosine matrix — ``cosine_similarity(Q, C)`` of
    Tool/rank_chunks_optimized.py:215-216 for callers that need every score (RRF ranks)."""
    dev = _require_cuda(corpus, queries)
    if corpus.dim() != 2 or queries.dim() != 2 or corpus.shape[1] != queries.shape[1]:
        raise ValueError(f"shape mismatch: corpus {tuple(corpus.shape)} vs queries {tuple(queries.shape)}")
    if not corpus.is_contiguous() or not queries.is_contiguous():
        raise ValueError("corpus and queries must be contiguous row-major tensors")
    n, d = corpus.shape
    b = queries.shape[0]
    lib = _lib.load()
    with torch.cuda.device(dev):
        need = lib.ss_cosine_topk_stream_workspace_bytes(n, d, _dtype_code(corpus), b, 1)
        ws = workspace(dev, need)
        out = torch.empty((b, n), dtype=torch.float32, device=dev)
        st = lib.ss_cosine_scores(corpus.data_ptr(), n, d, _dtype_code(corpus), queries.data_ptr(), b, _dtype_code(queries),
                                  ws.data_ptr(), ws.numel(), out.data_ptr(), _stream_ptr(dev))
        _lib.check(st, "ss_cosine_scores")
    return out


def rank_order(scores: torch.Tensor):
    """``(order, rank1)`` for each row of ``scores [B, N]``: ``order`` = ``np.argsort(-scores)``
    (ties: lower index first), ``rank1[b, i]`` = 1-based rank of element i
    (Tool/rank_chunks_optimized.py:225-235)."""
    dev = _require_cuda(scores)
    if scores.dim() != 2 or scores.dtype != torch.float32 or not scores.is_contiguous():
        raise ValueError("scores must be a contiguous float32 [B, N] tensor")
    b, n = scores.shape
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws = workspace(dev, lib.ss_rank_order_workspace_bytes(b, n), "sort")
        order = torch.empty((b, n), dtype=torch.int32, device=dev)
        rank1 = torch.empty((b, n), dtype=torch.int32, device=dev)
        st = lib.ss_rank_order(scores.data_ptr(), b, n, ws.data_ptr(), ws.numel(), order.data_ptr(), rank1.data_ptr(),
                               _stream_ptr(dev))
        _lib.check(st, "ss_rank_order")
    return order, rank1


def segmented_rank_rrf(chunks: torch.Tensor, offsets: torch.Tensor, queries: torch.Tensor, bm25: Optional[torch.Tensor] = None,
                       k_rrf: float = 60.0, upper_percentile: float = 80.0, lower_percentile: float = 20.0,
                       max_group_rows: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Ranking of every query group of a block in one launch (Tool/rank_chunks_optimized.py:215-250,518-519).

    ``chunks`` fp32 ``[rows, d]`` (groups concatenated), ``offsets`` int32 ``[G + 1]`` on the device, ``queries``
    fp32 ``[G, d]``, ``bm25`` fp32 ``[rows]`` (host-computed lexical scores) or None.  Returns device tensors:
    ``cosine`` fp32, ``rank_cosine`` / ``rank_bm25`` int32 (1-based inside the group), ``rrf`` fp64, ``order`` int32
    (local rows by fused score, best first) and ``thresholds`` fp64 ``[G, 2]`` (upper, lower percentile of rrf)."""
    dev = _require_cuda(chunks, offsets, queries)
    if chunks.dtype != torch.float32 or queries.dtype != torch.float32 or not chunks.is_contiguous() or not queries.is_contiguous():
        raise ValueError("chunks and queries must be contiguous float32 tensors")
    if offsets.dtype != torch.int32 or offsets.dim() != 1 or offsets.numel() != queries.shape[0] + 1:
        raise ValueError("offsets must be an int32 [groups + 1] tensor")
    if chunks.shape[1] != queries.shape[1]:
        raise ValueError("chunks and queries differ in dimension")
    rows, d = chunks.shape
    g = queries.shape[0]
    if max_group_rows is None:
        max_group_rows = int((offsets[1:] - offsets[:-1]).max().item()) if g else 0
    if bm25 is not None and (bm25.dtype != torch.float32 or bm25.numel() != rows or not bm25.is_cuda):
        raise ValueError("bm25 must be a CUDA float32 tensor with one score per chunk row")
    lib = _lib.load()
    with torch.cuda.device(dev):
        out = {"cosine": torch.empty(rows, dtype=torch.float32, device=dev),
               "rank_cosine": torch.empty(rows, dtype=torch.int32, device=dev),
               "rank_bm25": torch.empty(rows, dtype=torch.int32, device=dev) if bm25 is not None else None,
               "rrf": torch.empty(rows, dtype=torch.float64, device=dev),
               "order": torch.empty(rows, dtype=torch.int32, device=dev),
               "thresholds": torch.empty((g, 2), dtype=torch.float64, device=dev)}
        st = lib.ss_segmented_rank_rrf(chunks.data_ptr(), d, offsets.data_ptr(), g, max(1, int(max_group_rows)), queries.data_ptr(),
                                       bm25.data_ptr() if bm25 is not None else None, float(k_rrf), float(upper_percentile),
                                       float(lower_percentile), out["cosine"].data_ptr(), out["rank_cosine"].data_ptr(),
                                       out["rank_bm25"].data_ptr() if bm25 is not None else None, out["rrf"].data_ptr(),
                                       out["order"].data_ptr(), out["thresholds"].data_ptr(), _stream_ptr(dev))
        _lib.check(st, "ss_segmented_rank_rrf")
    return out


def topk_merge(keys: torch.Tensor, k_out: Optional[int] = None):
    """Merge ``keys[P, B, k]`` (P best-first lists per query) into the global top ``k_out``.

    Used after the NCCL all-gather of per-GPU results; returns ``(scores, indices, keys)``.
    """
    dev = _require_cuda(keys)
    if keys.dim() != 3 or keys.dtype != torch.int64 or not keys.is_contiguous():
        raise ValueError("keys must be a contiguous int64 tensor of shape [lists, queries, k]")
    p, b, k_in = keys.shape
    k_out = int(k_out or k_in)
    lib = _lib.load()
    with torch.cuda.device(dev):
        scores = torch.empty((b, k_out), dtype=torch.float32, device=dev)
        idx = torch.empty((b, k_out), dtype=torch.int64, device=dev)
        out_keys = torch.empty((b, k_out), dtype=torch.int64, device=dev)
        st = lib.ss_topk_merge(keys.data_ptr(), p, b, k_in, k_in, b * k_in, k_out, out_keys.data_ptr(),
                               scores.data_ptr(), idx.data_ptr(), _stream_ptr(dev))
        _lib.check(st, "ss_topk_merge")
    return scores, idx, out_keys


def row_inv_norms(rows: torch.Tensor, zero_value: float = 1.0) -> torch.Tensor:
    """``1/||row||`` in fp32; zero rows map to ``zero_value`` (1.0 = sklearn's rule)."""
    dev = _require_cuda(rows)
    if rows.dim() != 2 or not rows.is_contiguous():
        raise ValueError("rows must be a contiguous 2-D tensor")
    lib = _lib.load()
    with torch.cuda.device(dev):
        out = torch.empty(rows.shape[0], dtype=torch.float32, device=dev)
        st = lib.ss_row_inv_norms(rows.data_ptr(), rows.shape[0], rows.shape[1], _dtype_code(rows), float(zero_value),
                                  out.data_ptr(), _stream_ptr(dev))
        _lib.check(st, "ss_row_inv_norms")
    return out
