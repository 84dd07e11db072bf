"""Host-side operators for ragged (per-document) batches: K3 similarity matrices, K4 grouping
threshold pass, K5 splitter passes.  Documents are concatenated row-wise; a ``RaggedPlan`` holds
the CSR offsets and the derived index arrays on the device.  No CPU code path exists here.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .similarity import _dtype_code, _require_cuda, _stream_ptr

KNN_WIDTH = 33
TC_MAX_DIM = 768  # largest embedding width the 3xTF32 similarity kernel is dispatched for


@dataclass
class RaggedPlan:
    offsets: np.ndarray        # int32 [D+1] (host)
    s_offsets: np.ndarray      # int64 [D+1] (host) prefix sums of n^2
    tile_prefix: np.ndarray    # int32 [D+1] (host)
    total_tiles: int
    max_rows: int
    offsets_d: torch.Tensor
    s_offsets_d: torch.Tensor
    tile_prefix_d: torch.Tensor
    units128_d: Optional[torch.Tensor] = None   # int32 [units, 4] work list of the tensor-core kernel (built on first use)
    host_chunks: Optional[tuple] = None         # grouping_pass_host's document chunks and their plans (built on first use)

    @property
    def n_docs(self) -> int:
        return len(self.offsets) - 1

    @property
    def total_rows(self) -> int:
        return int(self.offsets[-1])

    @property
    def total_s(self) -> int:
        return int(self.s_offsets[-1])

    def sizes(self) -> np.ndarray:
        return np.diff(self.offsets)


def make_plan(sizes: Sequence[int], device) -> RaggedPlan:
    """Index arrays for a batch of documents with ``sizes[d]`` sentences each."""
    sizes = np.asarray(sizes, dtype=np.int64)
    if sizes.ndim != 1 or sizes.size == 0 or (sizes < 0).any():
        raise ValueError("sizes must be a non-empty 1-D array of non-negative sentence counts")
    if int(sizes.sum()) >= 2 ** 31:
        raise ValueError("more than 2^31 rows in one ragged batch")
    offsets = np.zeros(sizes.size + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(sizes)
    s_off = np.zeros(sizes.size + 1, dtype=np.int64)
    tile_prefix = np.zeros(sizes.size + 1, dtype=np.int32)
    total = ctypes.c_int64()
    mx = ctypes.c_int()
    lib = _lib.load()
    st = lib.ss_segmented_plan_host(offsets.ctypes.data, int(sizes.size), s_off.ctypes.data, tile_prefix.ctypes.data,
                                    ctypes.byref(total), ctypes.byref(mx))
    _lib.check(st, "ss_segmented_plan_host")
    dev = torch.device(device)
    return RaggedPlan(offsets, s_off, tile_prefix, int(total.value), int(mx.value),
                      torch.from_numpy(offsets).to(dev), torch.from_numpy(s_off).to(dev),
                      torch.from_numpy(tile_prefix).to(dev))


def _units128(plan: RaggedPlan, dev) -> torch.Tensor:
    """Work list {doc, tile row, tile col, 0} of 128 x 128 upper-triangular tiles, uploaded once per plan."""
    if plan.units128_d is None or plan.units128_d.device != dev:
        lib = _lib.load()
        total = ctypes.c_int64()
        _lib.check(lib.ss_segmented_plan128_host(plan.offsets.ctypes.data, plan.n_docs, None, 0, ctypes.byref(total)),
                   "ss_segmented_plan128_host")
        units = np.zeros((max(int(total.value), 1), 4), dtype=np.int32)
        _lib.check(lib.ss_segmented_plan128_host(plan.offsets.ctypes.data, plan.n_docs, units.ctypes.data, int(total.value),
                                                 ctypes.byref(total)), "ss_segmented_plan128_host")
        plan.units128_d = torch.from_numpy(units[: int(total.value)]).to(dev)
    return plan.units128_d


def segmented_simmatrix(E: torch.Tensor, plan: RaggedPlan, out: Optional[torch.Tensor] = None, algo: str = "auto",
                        validate: bool = False) -> torch.Tensor:
    """All per-document ``S = En @ En.T`` blocks, packed (Method/semantic_common.py:158-191).

    ``algo``: "tc" = tcgen05 kernel (3-term split product, needs dim % 4 == 0), "ffma" = CUDA-core fp32 kernel,
    "auto" = "tc" whenever its layout constraints hold.  The tensor-core kernel scales every row by a power of two taken
    from its first 32 elements; ``validate=True`` reads back its range flag (one stream synchronisation) and recomputes the
    batch with the fp32 kernel in the never-observed case that a later element of a row is > 2^14 times larger."""
    dev = _require_cuda(E)
    if E.dtype != torch.float32 or E.dim() != 2 or not E.is_contiguous():
        raise ValueError("E must be a contiguous float32 [total_rows, dim] tensor")
    if E.shape[0] != plan.total_rows:
        raise ValueError(f"E has {E.shape[0]} rows, plan expects {plan.total_rows}")
    if algo not in ("auto", "tc", "ffma"):
        raise ValueError("algo must be 'auto', 'tc' or 'ffma'")
    # The tensor core truncates when it accumulates, a bias that grows with dim and with |S| (measured
    # 5.6e-6 at dim = 768 on the diagonal, 2.9e-6 at 384): past 768 dimensions the CUDA-core fp32 kernel
    # keeps a safe margin to the 1e-5 parity bound.
    tc_ok = E.shape[1] % 4 == 0 and E.shape[1] <= TC_MAX_DIM and E.data_ptr() % 16 == 0 and plan.total_rows > 0
    if algo == "tc" and not tc_ok:
        raise ValueError(f"the tensor-core similarity kernel needs dim % 4 == 0, dim <= {TC_MAX_DIM} and 16-byte aligned rows")
    use_tc = tc_ok if algo == "auto" else algo == "tc"
    lib = _lib.load()
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty(plan.total_s, dtype=torch.float32, device=dev)
        elif out.numel() < plan.total_s or out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError("out must be a CUDA float32 tensor with at least plan.total_s elements")
        if use_tc:
            units = _units128(plan, dev)
            flag = torch.zeros(1, dtype=torch.int32, device=dev) if validate else None
            st = lib.ss_segmented_simmatrix_tc(E.data_ptr(), E.shape[0], E.shape[1], plan.offsets_d.data_ptr(),
                                               plan.s_offsets_d.data_ptr(), units.data_ptr(), units.shape[0], out.data_ptr(),
                                               flag.data_ptr() if flag is not None else None, _stream_ptr(dev))
            _lib.check(st, "ss_segmented_simmatrix_tc")
            if flag is None or int(flag.item()) == 0:
                return out
        st = lib.ss_segmented_simmatrix(E.data_ptr(), E.shape[1], plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(),
                                        plan.tile_prefix_d.data_ptr(), plan.n_docs, plan.total_tiles, out.data_ptr(),
                                        _stream_ptr(dev))
        _lib.check(st, "ss_segmented_simmatrix")
    return out


def group_threshold_pass(S: torch.Tensor, plan: RaggedPlan, tau: float = 0.15, knn_mode: int = 0,
                         symmetric: bool = False, out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """Grouping threshold pass for every document (Method/Semantic_Grouping_Optimized.py:100-115,
    270-283,343-360).  Returns device tensors: sim_sharp (packed), centrality [rows] f64,
    doc_stats [D,8] f64 = (mu, sigma, q80, q65, q60, 0.1*std, count, k), knn_idx/knn_val [rows,33].
    ``symmetric=True`` promises ``S == S.T`` bit for bit (the output of ``segmented_simmatrix``): quantiles and moments of
    the positive values are then computed from the strict upper triangle (half the histogram and gather work)."""
    dev = _require_cuda(S)
    if S.dtype != torch.float32 or not S.is_contiguous() or S.numel() < plan.total_s:
        raise ValueError("S must be the packed float32 output of segmented_simmatrix")
    lib = _lib.load()
    with torch.cuda.device(dev):
        if out is not None:   # caller-owned device buffers at least as large as this plan needs (grouping_pass_host's slots)
            sharp, cent, stats = out["sim_sharp"][: plan.total_s], out["centrality"][: plan.total_rows], out["doc_stats"][: plan.n_docs]
            kidx, kval = out["knn_idx"][: plan.total_rows], out["knn_val"][: plan.total_rows]
        else:
            sharp = torch.empty(plan.total_s, dtype=torch.float32, device=dev)
            cent = torch.empty(plan.total_rows, dtype=torch.float64, device=dev)
            stats = torch.empty((plan.n_docs, 8), dtype=torch.float64, device=dev)
            kidx = torch.empty((plan.total_rows, KNN_WIDTH), dtype=torch.int32, device=dev)
            kval = torch.empty((plan.total_rows, KNN_WIDTH), dtype=torch.float32, device=dev)
        st = lib.ss_group_threshold_pass(S.data_ptr(), plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), plan.n_docs,
                                         float(tau), int(knn_mode), int(bool(symmetric)), sharp.data_ptr(), cent.data_ptr(), stats.data_ptr(),
                                         kidx.data_ptr(), kval.data_ptr(), _stream_ptr(dev))
        _lib.check(st, "ss_group_threshold_pass")
    return {"sim_sharp": sharp, "centrality": cent, "doc_stats": stats, "knn_idx": kidx, "knn_val": kval}


def group_block_sums(sharp: torch.Tensor, plan: RaggedPlan, groups_per_doc: Sequence[Sequence[Sequence[int]]]):
    """Block sums of ``sim_sharp`` over member lists for every document of a batch — the device form of the reference's
    ``_mean_between`` / ``_mean_within`` loops and reassignment means (Method/Semantic_Grouping_Optimized.py:118-130,
    566-588).  ``groups_per_doc[d]`` = the member lists (document-local sentence indices, repeats allowed) of document d.
    Returns one ``(rowsum float64 [n, G], block float64 [G, G])`` pair of host arrays per document:
    ``rowsum[x, g] = sum(sharp[x, members_g])``, ``block[a, b] = sum(sharp[np.ix_(members_a, members_b)])``."""
    dev = _require_cuda(sharp)
    if sharp.dtype != torch.float32 or not sharp.is_contiguous() or sharp.numel() < plan.total_s:
        raise ValueError("sharp must be the packed float32 sim_sharp output of group_threshold_pass")
    if len(groups_per_doc) != plan.n_docs:
        raise ValueError("one list of member lists per document")
    sizes = plan.sizes()
    gcount = np.array([len(g) for g in groups_per_doc], dtype=np.int64)
    group_prefix = np.zeros(plan.n_docs + 1, dtype=np.int32)
    group_prefix[1:] = np.cumsum(gcount)
    flat = [np.asarray(m, dtype=np.int32).reshape(-1) for gs in groups_per_doc for m in gs]
    member_prefix = np.zeros(len(flat) + 1, dtype=np.int32)
    if flat:
        member_prefix[1:] = np.cumsum([m.size for m in flat])
    members = np.concatenate(flat) if flat and member_prefix[-1] > 0 else np.zeros(1, dtype=np.int32)
    rs_off = np.zeros(plan.n_docs + 1, dtype=np.int64)
    rs_off[1:] = np.cumsum(sizes.astype(np.int64) * gcount)
    blk_off = np.zeros(plan.n_docs + 1, dtype=np.int64)
    blk_off[1:] = np.cumsum(gcount * gcount)
    out = []
    if plan.total_rows == 0 or int(group_prefix[-1]) == 0:
        return [(np.zeros((int(n), int(g))), np.zeros((int(g), int(g)))) for n, g in zip(sizes, gcount)]
    lib = _lib.load()
    with torch.cuda.device(dev):
        row_doc = torch.repeat_interleave(torch.arange(plan.n_docs, dtype=torch.int32, device=dev),
                                          (plan.offsets_d[1:] - plan.offsets_d[:-1]).to(torch.int64))
        packed = np.concatenate([group_prefix.view(np.int32), member_prefix, members])   # one H2D copy for the three index arrays
        packed_d = torch.from_numpy(packed).to(dev)
        gp_d = packed_d[: group_prefix.size]
        mp_d = packed_d[group_prefix.size: group_prefix.size + member_prefix.size]
        mem_d = packed_d[group_prefix.size + member_prefix.size:]
        offs_d = torch.from_numpy(np.concatenate([rs_off[:-1], blk_off[:-1]])).to(dev)
        rowsum = torch.empty(max(int(rs_off[-1]), 1), dtype=torch.float64, device=dev)
        block = torch.empty(max(int(blk_off[-1]), 1), dtype=torch.float64, device=dev)
        st = lib.ss_group_block_sums(sharp.data_ptr(), plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), plan.n_docs,
                                     plan.total_rows, row_doc.data_ptr(), gp_d.data_ptr(), mp_d.data_ptr(), mem_d.data_ptr(),
                                     offs_d.data_ptr(), offs_d[plan.n_docs:].data_ptr(), rowsum.data_ptr(), block.data_ptr(),
                                     _stream_ptr(dev))
        _lib.check(st, "ss_group_block_sums")
        rs_h, blk_h = rowsum.cpu().numpy(), block.cpu().numpy()
    for d in range(plan.n_docs):
        n, g = int(sizes[d]), int(gcount[d])
        out.append((rs_h[rs_off[d]:rs_off[d + 1]].reshape(n, g), blk_h[blk_off[d]:blk_off[d + 1]].reshape(g, g)))
    return out


class DocBlockSums:
    """``group_block_sums`` bound to ONE document whose ``sim_sharp`` stays on the device: the host clustering stage calls
    it a handful of times per document (once per merge / refine / reassign phase, once per bisection), each call being
    one small H2D copy of the member lists, two launches and one D2H copy of the sums."""

    def __init__(self, sharp_doc: torch.Tensor, n: int):
        dev = _require_cuda(sharp_doc)
        if sharp_doc.dtype != torch.float32 or not sharp_doc.is_contiguous() or sharp_doc.numel() < n * n:
            raise ValueError("sharp_doc must be a contiguous float32 CUDA tensor holding the document's n x n sim_sharp")
        self.sharp, self.n, self.dev = sharp_doc, int(n), dev
        with torch.cuda.device(dev):
            self.offsets_d = torch.tensor([0, self.n], dtype=torch.int32, device=dev)
            self.s_offsets_d = torch.tensor([0, self.n * self.n], dtype=torch.int64, device=dev)
            self.row_doc = torch.zeros(max(self.n, 1), dtype=torch.int32, device=dev)
            self.zero_off = torch.zeros(1, dtype=torch.int64, device=dev)
        self.calls = 0

    def __call__(self, groups: Sequence[Sequence[int]]):
        """``(rowsum float64 [n, G], block float64 [G, G])`` for the member lists ``groups`` (repeats allowed)."""
        g = len(groups)
        n = self.n
        if g == 0 or n == 0:
            return np.zeros((n, g)), np.zeros((g, g))
        sizes = [len(m) for m in groups]
        packed = np.zeros(2 + g + 1 + max(sum(sizes), 1), dtype=np.int32)
        packed[1] = g                                   # group_prefix = [0, G]
        packed[3:3 + g] = np.cumsum(sizes)              # member_prefix = [0, ...]
        if sum(sizes):
            packed[3 + g:3 + g + sum(sizes)] = np.concatenate([np.asarray(m, dtype=np.int32).reshape(-1) for m in groups if len(m)])
        lib = _lib.load()
        with torch.cuda.device(self.dev):
            pk = torch.from_numpy(packed).to(self.dev)
            out = torch.empty(n * g + g * g, dtype=torch.float64, device=self.dev)
            st = lib.ss_group_block_sums(self.sharp.data_ptr(), self.offsets_d.data_ptr(), self.s_offsets_d.data_ptr(), 1, n,
                                         self.row_doc.data_ptr(), pk.data_ptr(), pk[2:].data_ptr(), pk[3 + g:].data_ptr(),
                                         self.zero_off.data_ptr(), self.zero_off.data_ptr(), out.data_ptr(),
                                         out[n * g:].data_ptr(), _stream_ptr(self.dev))
            _lib.check(st, "ss_group_block_sums")
            h = out.cpu().numpy()
        self.calls += 1
        return h[: n * g].reshape(n, g), h[n * g:].reshape(g, g)


def group_coassociation(labels) -> torch.Tensor:
    """Consensus matrix of a sweep of labelings (Method/Semantic_Grouping_Optimized.py:231-241): ``labels`` int
    ``[L, n]`` (host or device) -> float64 ``[n, n]`` device tensor, ``C[i, j]`` = share of labelings that put i and
    j together, zero diagonal."""
    lab = torch.as_tensor(np.asarray(labels) if not isinstance(labels, torch.Tensor) else labels)
    if lab.dim() != 2 or lab.shape[0] < 1 or lab.shape[1] < 1:
        raise ValueError("labels must be a non-empty [labelings, n] integer array")
    lab = lab.to(device="cuda", dtype=torch.int32).contiguous()
    dev = _require_cuda(lab)
    lib = _lib.load()
    with torch.cuda.device(dev):
        out = torch.empty((lab.shape[1], lab.shape[1]), dtype=torch.float64, device=dev)
        _lib.check(lib.ss_group_coassociation(lab.data_ptr(), lab.shape[0], lab.shape[1], out.data_ptr(), _stream_ptr(dev)),
                   "ss_group_coassociation")
    return out


STAT_KEYS = ("min", "max", "mean", "std", "p10", "p25", "p50", "p75", "p80", "p85", "p90", "p95")


def similarity_distribution(S: torch.Tensor, plan: RaggedPlan, eps: float = 1e-5) -> torch.Tensor:
    """Per-document statistics of the strict upper triangle of S (Method/semantic_common.py:250-270).
    Returns a float64 ``[D, 13]`` tensor: count, then the values of ``STAT_KEYS``."""
    dev = _require_cuda(S)
    if S.dtype != torch.float32 or not S.is_contiguous() or S.numel() < plan.total_s:
        raise ValueError("S must be the packed float32 output of segmented_simmatrix")
    lib = _lib.load()
    with torch.cuda.device(dev):
        out = torch.empty((plan.n_docs, 13), dtype=torch.float64, device=dev)
        st = lib.ss_similarity_distribution(S.data_ptr(), plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), plan.n_docs,
                                            float(eps), out.data_ptr(), _stream_ptr(dev))
        _lib.check(st, "ss_similarity_distribution")
    return out


def diameter_split(S: torch.Tensor, plan: RaggedPlan, threshold: float):
    """Diameter-bounded splitting of every document (data_process/simple_chunk_controller.py:571-594 on the similarity
    matrix of :614): returns ``(span_ends int32 [rows], n_spans int32 [D], diameter float64 [D])`` device tensors —
    document ``d`` is split into the spans that end (exclusively) at ``span_ends[offsets[d] : offsets[d] + n_spans[d]]``;
    ``diameter[d] = 1 - min off-diagonal similarity`` of the whole document."""
    dev = _require_cuda(S)
    if S.dtype != torch.float32 or not S.is_contiguous() or S.numel() < plan.total_s:
        raise ValueError("S must be the packed float32 output of segmented_simmatrix")
    lib = _lib.load()
    with torch.cuda.device(dev):
        ends = torch.zeros(max(plan.total_rows, 1), dtype=torch.int32, device=dev)
        n_spans = torch.empty(plan.n_docs, dtype=torch.int32, device=dev)
        diam = torch.empty(plan.n_docs, dtype=torch.float64, device=dev)
        st = lib.ss_diameter_split(S.data_ptr(), plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), plan.n_docs,
                                   max(plan.max_rows, 1), float(threshold), ends.data_ptr(), n_spans.data_ptr(), diam.data_ptr(),
                                   _stream_ptr(dev))
        _lib.check(st, "ss_diameter_split")
    return ends, n_spans, diam


def c99_rank_matrix(S: torch.Tensor, plan: RaggedPlan, use_local_rank: bool = False, mask_size: int = 11,
                    symmetric: bool = False) -> torch.Tensor:
    """C99 rank transform of every document's S (Method/Semantic_Splitter_Optimized.py:171-192), packed like S.
    ``symmetric=True`` promises ``S == S.T`` bit for bit (the output of ``segmented_simmatrix``): the global mode then
    takes the column ranks as the transposed row ranks instead of sorting every column."""
    dev = _require_cuda(S)
    if S.dtype != torch.float32 or not S.is_contiguous() or S.numel() < plan.total_s:
        raise ValueError("S must be the packed float32 output of segmented_simmatrix")
    lib = _lib.load()
    with torch.cuda.device(dev):
        R = torch.empty(plan.total_s, dtype=torch.float32, device=dev)
        ws = torch.empty(plan.total_rows + plan.total_rows // 16 + 2 * plan.n_docs + 8, dtype=torch.int32, device=dev)
        st = lib.ss_c99_rank_matrix(S.data_ptr(), plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), plan.n_docs,
                                    plan.total_rows, max(plan.max_rows, 1),
                                    int(bool(use_local_rank)) | (2 if (symmetric and not use_local_rank) else 0), int(mask_size),
                                    ws.data_ptr(), R.data_ptr(), _stream_ptr(dev))
        _lib.check(st, "ss_c99_rank_matrix")
    return R


C99_CUTS_MAX_ROWS = 4096  # longest document the device cut search takes (ss_c99_divisive_cuts)


def c99_divisive_cuts(R: torch.Tensor, plan: RaggedPlan, min_chunk, max_cuts=None, min_gain: float = 0.01,
                      stop_by_gain: bool = True, want_profile: bool = False):
    """Greedy C99 cut search on every document's rank matrix (Method/Semantic_Splitter_Optimized.py:194-238).

    ``min_chunk`` / ``max_cuts``: one int for all documents or a per-document sequence (``max_cuts`` ``None`` or
    negative = unlimited).  Returns ``(cuts int32 [rows], n_cuts int32 [D], profile float64 [rows] | None)``:
    document ``d``'s cuts, in pick order, are ``cuts[offsets[d] : offsets[d] + n_cuts[d]]`` and its density
    profile ``profile[offsets[d] : offsets[d] + n_cuts[d] + 1]``."""
    dev = _require_cuda(R)
    if R.dtype != torch.float32 or not R.is_contiguous() or R.numel() < plan.total_s:
        raise ValueError("R must be the packed float32 output of c99_rank_matrix")
    if plan.max_rows > C99_CUTS_MAX_ROWS:
        raise ValueError(f"the device cut search takes documents of at most {C99_CUTS_MAX_ROWS} sentences")

    def per_doc(value, name, default):
        if value is None:
            return None, default
        if np.ndim(value) == 0:
            return None, int(value)
        arr = np.ascontiguousarray(value, dtype=np.int32)
        if arr.shape != (plan.n_docs,):
            raise ValueError(f"{name} must be a scalar or one value per document")
        return torch.from_numpy(arr).to(dev), default

    mc_d, mc_all = per_doc(min_chunk, "min_chunk", 0)
    if (mc_d is None and mc_all < 1) or (mc_d is not None and int(mc_d.min()) < 1):
        raise ValueError("min_chunk must be >= 1")
    mx_d, mx_all = per_doc(max_cuts, "max_cuts", -1)
    sizes = plan.sizes().astype(np.int64)
    sat_off = np.zeros(plan.n_docs + 1, dtype=np.int64)
    sat_off[1:] = np.cumsum((sizes + 1) ** 2)
    lib = _lib.load()
    with torch.cuda.device(dev):
        sat = torch.empty(int(sat_off[-1]), dtype=torch.float64, device=dev)
        sat_off_d = torch.from_numpy(sat_off[:-1].copy()).to(dev)
        cuts = torch.zeros(max(plan.total_rows, 1), dtype=torch.int32, device=dev)
        n_cuts = torch.empty(plan.n_docs, dtype=torch.int32, device=dev)
        profile = torch.zeros(max(plan.total_rows, 1), dtype=torch.float64, device=dev) if want_profile else None
        st = lib.ss_c99_divisive_cuts(R.data_ptr(), plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), sat_off_d.data_ptr(),
                                      plan.n_docs, max(plan.max_rows, 1), mc_d.data_ptr() if mc_d is not None else None, mc_all,
                                      mx_d.data_ptr() if mx_d is not None else None, mx_all, float(min_gain), int(bool(stop_by_gain)),
                                      sat.data_ptr(), cuts.data_ptr(), n_cuts.data_ptr(),
                                      profile.data_ptr() if profile is not None else None, _stream_ptr(dev))
        _lib.check(st, "ss_c99_divisive_cuts")
    return cuts, n_cuts, profile


def adjacent_cosine(E: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``adj[r] = cos(E[r], E[r+1])`` for the whole concatenated matrix
    (Method/Semantic_Splitter_Optimized.py:140-152,412); the last row gets 0."""
    dev = _require_cuda(E)
    if E.dim() != 2 or not E.is_contiguous():
        raise ValueError("E must be a contiguous 2-D tensor")
    lib = _lib.load()
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty(E.shape[0], dtype=torch.float32, device=dev)
        elif out.dtype != torch.float32 or out.numel() != E.shape[0] or not out.is_cuda or not out.is_contiguous():
            raise ValueError("out must be a contiguous CUDA float32 tensor with one element per row of E")
        st = lib.ss_segmented_adjacent_cosine(E.data_ptr(), E.shape[0], E.shape[1], _dtype_code(E), out.data_ptr(),
                                              _stream_ptr(dev))
        _lib.check(st, "ss_segmented_adjacent_cosine")
    return out


def segmented_percentile(adj: torch.Tensor, plan: RaggedPlan, pct: float = 95.0, want_stats: bool = True, out=None):
    """Per-document percentile threshold of ``1 - adj`` + breakpoint flags (BASELINE.json cfg 3)
    and, optionally, the splitter's robust statistics (Splitter:340-356,417-437).
    Returns ``(thr [D] f64, flags [rows] u8, stats [D,4] f64 | None, smooth [rows] f32 | None)``;
    stats columns are (median, MAD+1e-9, P25, P75) of the median-of-3 smoothed series.
    ``adj`` is updated in place: the slot of each document's last sentence is set to 0.
    ``out``: optional ``(thr, flags, stats, smooth)`` tuple of a previous call to overwrite."""
    dev = _require_cuda(adj)
    if adj.dtype != torch.float32 or adj.dim() != 1 or not adj.is_contiguous() or adj.numel() != plan.total_rows:
        raise ValueError("adj must be the float32 [total_rows] output of adjacent_cosine")
    lib = _lib.load()
    with torch.cuda.device(dev):
        if out is not None:   # (thr, flags, stats, smooth) buffers of a previous call with the same plan shape
            thr, flags, stats, smooth = out
            if thr.numel() != plan.n_docs or flags.numel() != plan.total_rows or (want_stats and (stats is None or smooth is None)):
                raise ValueError("out does not match the plan")
        else:
            thr = torch.empty(plan.n_docs, dtype=torch.float64, device=dev)
            flags = torch.empty(plan.total_rows, dtype=torch.uint8, device=dev)
            stats = torch.empty((plan.n_docs, 4), dtype=torch.float64, device=dev) if want_stats else None
            smooth = torch.empty(plan.total_rows, dtype=torch.float32, device=dev) if want_stats else None
        st = lib.ss_segmented_percentile(adj.data_ptr(), plan.offsets_d.data_ptr(), plan.n_docs, max(plan.max_rows, 1),
                                         float(pct), thr.data_ptr(), flags.data_ptr(),
                                         stats.data_ptr() if want_stats else None,
                                         smooth.data_ptr() if want_stats else None, _stream_ptr(dev))
        _lib.check(st, "ss_segmented_percentile")
    return thr, flags, stats, smooth


def knn_graph_from_lists(knn_idx: np.ndarray, knn_val: np.ndarray, floor: float) -> np.ndarray:
    """Dense symmetric ``W`` of one document from its neighbour lists — the tail of
    ``_build_knn_graph`` (Grouping:277-283): drop self, keep ``val >= floor``, ``max(W, W.T)``.
    Pure indexing on a few hundred entries per document; stays on the host next to the
    sequential clustering that consumes ``W``."""
    n = knn_idx.shape[0]
    W = np.zeros((n, n), dtype=float)
    rows = np.repeat(np.arange(n), knn_idx.shape[1])
    cols = knn_idx.reshape(-1).astype(np.int64)
    vals = knn_val.reshape(-1).astype(float)
    keep = (cols >= 0) & (cols != rows) & (vals >= floor)
    W[rows[keep], cols[keep]] = vals[keep]
    return np.maximum(W, W.T)


# ----------------------------------------------------------------------------------------------------------------
# Host-buffer operators: the calls a pipeline makes when its embeddings sit in host memory (the reference's
# numpy arrays).  Inputs are host tensors (pinned memory makes the copies asynchronous), outputs are pinned host
# tensors; every function raises without a CUDA device.
# ----------------------------------------------------------------------------------------------------------------
def _pinned_like(t: torch.Tensor) -> torch.Tensor:
    return torch.empty(t.shape, dtype=t.dtype, pin_memory=True)


def _to_host(tensors: Dict[str, torch.Tensor], out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    host = out if out is not None else {k: _pinned_like(v) for k, v in tensors.items()}
    for k, v in tensors.items():
        host[k].copy_(v, non_blocking=True)
    return host


_HOST_CHUNK_BYTES = 256 << 20   # embeddings + similarity matrix of one pipeline chunk of grouping_pass_host (128 MB ... 1.5 GB measured within 10 %)
_side_streams: Dict[int, tuple] = {}


def _copy_streams(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    return _side_streams[key]


def _host_chunks(plan: RaggedPlan, dim: int, chunk_bytes: int, dev):
    """Contiguous document ranges of at most ``chunk_bytes`` of embeddings + similarity matrix each, with their own plans
    (cached on the parent plan: a caller that reuses a plan pays for them once)."""
    key = (dim, chunk_bytes)
    if plan.host_chunks is not None and plan.host_chunks[0] == key:
        return plan.host_chunks[1]
    sizes = np.diff(plan.offsets).astype(np.int64)
    cost = sizes * dim * 4 + sizes * sizes * 4
    chunks, d0, acc = [], 0, 0
    for d in range(plan.n_docs):
        if d > d0 and acc + int(cost[d]) > chunk_bytes:
            chunks.append((d0, d))
            d0, acc = d, 0
        acc += int(cost[d])
    chunks.append((d0, plan.n_docs))
    out = [(a, b, plan if len(chunks) == 1 else make_plan(sizes[a:b], dev)) for a, b in chunks]
    plan.host_chunks = (key, out)
    return out


def grouping_pass_host(E_host: torch.Tensor, sizes: Sequence[int], tau: float = 0.15, knn_mode: int = 0,
                       out: Optional[Dict[str, torch.Tensor]] = None, plan: Optional[RaggedPlan] = None,
                       chunk_bytes: int = _HOST_CHUNK_BYTES) -> Dict[str, torch.Tensor]:
    """``create_similarity_matrix`` + the grouping threshold pass (Method/semantic_common.py:158-191,
    Method/Semantic_Grouping_Optimized.py:100-115,270-283,351-355) for a packed batch of documents whose embeddings are
    in HOST memory: H2D copy, K3, K4, D2H copy of everything the host clustering stage reads (S, sim_sharp,
    centrality, per-document thresholds, neighbour lists).  The step is bound by the host link — it returns about as
    many bytes as it takes — so the batch runs as a pipeline of document chunks: the H2D copy of chunk i + 1, the
    kernels of chunk i and the D2H copy of chunk i - 1 overlap (two copy streams, two device slots), which lets the two
    directions of the link work at the same time.  ``out``: the dict a previous call returned, to reuse its pinned
    buffers.  The call synchronises before returning (the caller reads the results)."""
    if E_host.is_cuda:
        raise ValueError("grouping_pass_host takes host embeddings (use segmented_simmatrix / group_threshold_pass for device tensors)")
    dev = torch.device("cuda", torch.cuda.current_device())
    plan = plan or make_plan(sizes, dev)
    if E_host.dtype != torch.float32 or E_host.dim() != 2 or E_host.shape[0] != plan.total_rows:
        raise ValueError("E_host must be a float32 [total_rows, dim] tensor matching the plan")
    dim = E_host.shape[1]
    chunks = _host_chunks(plan, dim, int(chunk_bytes), dev)
    if len(chunks) == 1:
        E = E_host.to(dev, non_blocking=True)
        S = segmented_simmatrix(E, plan)
        res = group_threshold_pass(S, plan, tau=tau, knn_mode=knn_mode, symmetric=True)   # K3 mirrors its tiles
        res["S"] = S
        host = _to_host(res, out)
        torch.cuda.current_stream(dev).synchronize()
        return host
    shapes = {"S": ((plan.total_s,), torch.float32), "sim_sharp": ((plan.total_s,), torch.float32),
              "centrality": ((plan.total_rows,), torch.float64), "doc_stats": ((plan.n_docs, 8), torch.float64),
              "knn_idx": ((plan.total_rows, KNN_WIDTH), torch.int32), "knn_val": ((plan.total_rows, KNN_WIDTH), torch.float32)}
    host = out if out is not None else {k: torch.empty(shp, dtype=dt, pin_memory=True) for k, (shp, dt) in shapes.items()}
    max_rows = max(p.total_rows for _, _, p in chunks)
    max_s = max(p.total_s for _, _, p in chunks)
    max_docs = max(p.n_docs for _, _, p in chunks)
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = _copy_streams(dev)
    slots = []
    for _ in range(2):
        slots.append({"E": torch.empty((max_rows, dim), dtype=torch.float32, device=dev),
                      "S": torch.empty(max_s, dtype=torch.float32, device=dev),
                      "sim_sharp": torch.empty(max_s, dtype=torch.float32, device=dev),
                      "centrality": torch.empty(max_rows, dtype=torch.float64, device=dev),
                      "doc_stats": torch.empty((max_docs, 8), dtype=torch.float64, device=dev),
                      "knn_idx": torch.empty((max_rows, KNN_WIDTH), dtype=torch.int32, device=dev),
                      "knn_val": torch.empty((max_rows, KNN_WIDTH), dtype=torch.float32, device=dev),
                      "copied": torch.cuda.Event(), "computed": torch.cuda.Event(), "drained": torch.cuda.Event(), "used": False})
    s_in.wait_stream(cur)    # the slots were allocated on the caller's stream
    s_out.wait_stream(cur)
    for i, (d0, d1, sub) in enumerate(chunks):
        slot = slots[i & 1]
        r0, r1 = int(plan.offsets[d0]), int(plan.offsets[d1])
        e0, e1 = int(plan.s_offsets[d0]), int(plan.s_offsets[d1])
        if r1 == r0:   # a chunk of empty documents
            host["doc_stats"][d0:d1].zero_()
            continue
        if slot["used"]:
            s_in.wait_event(slot["computed"])    # the kernels of chunk i - 2 have read this slot's embeddings
        with torch.cuda.stream(s_in):
            slot["E"][: r1 - r0].copy_(E_host[r0:r1], non_blocking=True)
            slot["copied"].record(s_in)
        cur.wait_event(slot["copied"])
        if slot["used"]:
            cur.wait_event(slot["drained"])      # ... and its results have left for the host
        segmented_simmatrix(slot["E"][: r1 - r0], sub, out=slot["S"])
        group_threshold_pass(slot["S"], sub, tau=tau, knn_mode=knn_mode, symmetric=True, out=slot)
        slot["computed"].record(cur)
        s_out.wait_event(slot["computed"])
        with torch.cuda.stream(s_out):
            host["S"][e0:e1].copy_(slot["S"][: e1 - e0], non_blocking=True)
            host["sim_sharp"][e0:e1].copy_(slot["sim_sharp"][: e1 - e0], non_blocking=True)
            host["centrality"][r0:r1].copy_(slot["centrality"][: r1 - r0], non_blocking=True)
            host["doc_stats"][d0:d1].copy_(slot["doc_stats"][: d1 - d0], non_blocking=True)
            host["knn_idx"][r0:r1].copy_(slot["knn_idx"][: r1 - r0], non_blocking=True)
            host["knn_val"][r0:r1].copy_(slot["knn_val"][: r1 - r0], non_blocking=True)
            slot["drained"].record(s_out)
        slot["used"] = True
    cur.wait_stream(s_out)
    cur.wait_stream(s_in)
    cur.synchronize()
    return host


def splitter_breakpoints_host(E_host: torch.Tensor, sizes: Sequence[int], pct: float = 95.0, want_stats: bool = False,
                              out: Optional[Dict[str, torch.Tensor]] = None, plan: Optional[RaggedPlan] = None) -> Dict[str, torch.Tensor]:
    """Adjacent-sentence cosine + per-document percentile breakpoints (Method/Semantic_Splitter_Optimized.py:140-152,412;
    BASELINE.json config 3) for a packed batch of documents in HOST memory: H2D, K5a, K5b, D2H of ``adj`` (fp32 per row),
    ``thr`` (fp64 per document) and ``flags`` (u8 per row)."""
    if E_host.is_cuda:
        raise ValueError("splitter_breakpoints_host takes host embeddings")
    dev = torch.device("cuda", torch.cuda.current_device())
    plan = plan or make_plan(sizes, dev)
    E = E_host.to(dev, non_blocking=True)
    adj = adjacent_cosine(E)
    thr, flags, stats, smooth = segmented_percentile(adj, plan, pct, want_stats=want_stats)
    res = {"adj": adj, "thr": thr, "flags": flags}
    if want_stats:
        res.update({"stats": stats, "smooth": smooth})
    host = _to_host(res, out)
    torch.cuda.current_stream(dev).synchronize()
    return host


class SplitterStream:
    """Config 3 at corpus scale: batches of documents stream from pinned HOST memory through a double-buffered
    H2D -> K5a -> K5b -> D2H pipeline (copy engine and SMs overlap; two device buffers, two result buffers).  A corpus of
    1 M documents x 384 fp32 is 405 GB: it never fits the device, and the pipeline is bound by the host link.

        stream = SplitterStream(max_rows, dim, max_docs)
        for E_host_pinned, sizes in batches:
            prev = stream.submit(E_host_pinned, sizes)      # results of the batch submitted two calls ago, or None
        for res in stream.drain(): ...
    """

    def __init__(self, max_rows: int, dim: int, max_docs: int, pct: float = 95.0, device=None):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.pct = float(pct)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.slots = []
        for _ in range(2):
            self.slots.append({
                "E": torch.empty((max_rows, dim), dtype=torch.float32, device=self.dev),
                "adj": torch.empty(max_rows, dtype=torch.float32, device=self.dev),
                "h_thr": torch.empty(max_docs, dtype=torch.float64, pin_memory=True),
                "h_flags": torch.empty(max_rows, dtype=torch.uint8, pin_memory=True),
                "copied": torch.cuda.Event(), "computed": torch.cuda.Event(), "done": torch.cuda.Event(),
                "busy": False, "plan": None, "rows": 0,
            })
        self.turn = 0

    def _finish(self, slot):
        slot["done"].synchronize()
        slot["busy"] = False
        plan = slot["plan"]
        return {"thr": slot["h_thr"][: plan.n_docs], "flags": slot["h_flags"][: slot["rows"]], "plan": plan}

    def submit(self, E_host: torch.Tensor, sizes: Sequence[int], plan: Optional[RaggedPlan] = None):
        slot = self.slots[self.turn]
        self.turn ^= 1
        ready = self._finish(slot) if slot["busy"] else None
        plan = plan or make_plan(sizes, self.dev)
        rows = plan.total_rows
        if rows > slot["E"].shape[0] or plan.n_docs > slot["h_thr"].numel():
            raise ValueError("batch larger than the stream's buffers")
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot["computed"])          # the kernels that read this buffer last time are done
            slot["E"][:rows].copy_(E_host[:rows], non_blocking=True)
            slot["copied"].record(self.copy_stream)
        main.wait_event(slot["copied"])
        adjacent_cosine(slot["E"][:rows], out=slot["adj"][:rows])
        thr, flags, _, _ = segmented_percentile(slot["adj"][:rows], plan, self.pct, want_stats=False)
        slot["computed"].record(main)
        slot["h_thr"][: plan.n_docs].copy_(thr, non_blocking=True)
        slot["h_flags"][:rows].copy_(flags, non_blocking=True)
        slot["done"].record(main)
        slot.update(busy=True, plan=plan, rows=rows)
        return ready

    def drain(self):
        out = []
        for _ in range(2):
            slot = self.slots[self.turn]
            self.turn ^= 1
            if slot["busy"]:
                out.append(self._finish(slot))
        return out


def c99_cuts_host(E_host: torch.Tensor, sizes: Sequence[int], min_chunk, use_local_rank: bool = False, mask_size: int = 11,
                  min_gain: float = 0.01, plan: Optional[RaggedPlan] = None, chunk_bytes: int = _HOST_CHUNK_BYTES) -> Dict[str, torch.Tensor]:
    """The C99 leg of the splitter (Method/Semantic_Splitter_Optimized.py:169-238) for a packed batch of documents in
    HOST memory: H2D, similarity matrices, rank transform (global, or the reference preset's 11 x 11 local rank,
    data_process/simple_chunk_controller.py:1451), divisive cut search, D2H of the cuts.  The step takes megabytes in per
    byte out, so long batches run as document chunks whose H2D copy (a side stream, two device slots) overlaps the
    kernels of the chunk before."""
    if E_host.is_cuda:
        raise ValueError("c99_cuts_host takes host embeddings")
    dev = torch.device("cuda", torch.cuda.current_device())
    plan = plan or make_plan(sizes, dev)
    if E_host.dtype != torch.float32 or E_host.dim() != 2 or E_host.shape[0] != plan.total_rows:
        raise ValueError("E_host must be a float32 [total_rows, dim] tensor matching the plan")
    chunks = _host_chunks(plan, E_host.shape[1], int(chunk_bytes), dev)
    if len(chunks) == 1:
        E = E_host.to(dev, non_blocking=True)
        S = segmented_simmatrix(E, plan)
        R = c99_rank_matrix(S, plan, use_local_rank=use_local_rank, mask_size=mask_size, symmetric=not use_local_rank)
        cuts, n_cuts, _ = c99_divisive_cuts(R, plan, min_chunk, min_gain=min_gain)
        host = _to_host({"cuts": cuts, "n_cuts": n_cuts})
        torch.cuda.current_stream(dev).synchronize()
        return host
    per_doc_min = None if np.ndim(min_chunk) == 0 else np.ascontiguousarray(min_chunk, dtype=np.int32)
    if per_doc_min is not None and per_doc_min.shape != (plan.n_docs,):
        raise ValueError("min_chunk must be a scalar or one value per document")
    cur = torch.cuda.current_stream(dev)
    s_in, _ = _copy_streams(dev)
    max_rows = max(p.total_rows for _, _, p in chunks)
    slots = [{"E": torch.empty((max_rows, E_host.shape[1]), dtype=torch.float32, device=dev), "copied": torch.cuda.Event(),
              "computed": torch.cuda.Event(), "used": False} for _ in range(2)]
    cuts = torch.empty(plan.total_rows, dtype=torch.int32, device=dev)
    n_cuts = torch.zeros(plan.n_docs, dtype=torch.int32, device=dev)
    s_in.wait_stream(cur)

    def load(i):
        d0, d1, _ = chunks[i]
        r0, r1 = int(plan.offsets[d0]), int(plan.offsets[d1])
        slot = slots[i & 1]
        if slot["used"]:
            s_in.wait_event(slot["computed"])   # the kernels of chunk i - 2 have read this slot
        with torch.cuda.stream(s_in):
            if r1 > r0:
                slot["E"][: r1 - r0].copy_(E_host[r0:r1], non_blocking=True)
            slot["copied"].record(s_in)

    load(0)
    for i, (d0, d1, sub) in enumerate(chunks):
        if i + 1 < len(chunks):
            load(i + 1)                          # its copy runs under this chunk's kernels
        slot = slots[i & 1]
        r0, r1 = int(plan.offsets[d0]), int(plan.offsets[d1])
        cur.wait_event(slot["copied"])
        if r1 > r0:
            S = segmented_simmatrix(slot["E"][: r1 - r0], sub)
            R = c99_rank_matrix(S, sub, use_local_rank=use_local_rank, mask_size=mask_size, symmetric=not use_local_rank)
            c, nc, _ = c99_divisive_cuts(R, sub, min_chunk if per_doc_min is None else per_doc_min[d0:d1], min_gain=min_gain)
            cuts[r0:r1].copy_(c[: r1 - r0])
            n_cuts[d0:d1].copy_(nc)
        slot["computed"].record(cur)
        slot["used"] = True
    host = _to_host({"cuts": cuts, "n_cuts": n_cuts})
    cur.wait_stream(s_in)
    cur.synchronize()
    return host
