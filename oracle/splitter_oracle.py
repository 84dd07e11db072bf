"""TEST INFRASTRUCTURE — oracle for the semantic-splitter similarity passes.

Restates the dense arithmetic of Method/Semantic_Splitter_Optimized.py: ``_embed``'s
normalisation ``:140-152``, adjacent-sentence similarity ``:412``, ``_median_smooth``
``:340-356``, the robust MAD/IQR sigmoid ``:417-437`` (+ valley tau ``:465``), and the C99
similarity / rank matrices ``:169-192`` and the divisive cut search on the rank matrix ``:194-264``.  BASELINE.json config 3's "95th-percentile
breakpoints" has no counterpart in the reference; it is pinned to numpy as
``d = 1 - adj``, ``thr = np.percentile(d, 95)``, ``breakpoints = where(d > thr)``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from .simmatrix_oracle import normalize_rows_1e9


def adjacent_sims_ref(E: np.ndarray) -> np.ndarray:
    """Splitter:412 — ``float(En[i] @ En[i+1])`` for i < n-1 (fp32 dot widened to fp64)."""
    En = normalize_rows_1e9(np.asarray(E))
    n = En.shape[0]
    return np.array([float(En[i] @ En[i + 1]) for i in range(n - 1)], dtype=float)


def median_smooth_ref(arr, window: int = 3) -> List[float]:
    """Splitter:340-356 — odd window, edge-replicated padding, ``np.median`` per window."""
    w = int(window)
    arr = list(arr)
    if w <= 1:
        return arr
    if w % 2 == 0:
        w += 1
    n = len(arr)
    if n == 0 or w > max(1, n):
        return arr
    half = w // 2
    padded = [arr[0]] * half + arr + [arr[-1]] * half
    return [float(np.median(padded[i:i + w])) for i in range(n)]


def robust_stats_ref(adj_sims, smooth_window: int = 3) -> Dict[str, object]:
    """Splitter:417-437,465 (auto_params path): smoothed series, median, MAD+1e-9, IQR,
    ``tau = max(IQR/2, 0.05)``, ``sigmoid((x-med)/MAD / tau)``, ``valley_tau = max(IQR/2, 0.06)``."""
    adj_base = median_smooth_ref(adj_sims, smooth_window) if smooth_window and smooth_window > 1 else list(adj_sims)
    x = np.array(adj_base, dtype=float)
    med = float(np.median(x)) if x.size else 0.0
    mad = float(np.median(np.abs(x - med)) + 1e-9) if x.size else 1e-9
    iqr = float(np.percentile(x, 75) - np.percentile(x, 25)) if x.size else 0.0
    tau = max(iqr / 2.0, 0.05)
    z = (x - med) / mad
    with np.errstate(over="ignore"):
        sig = 1.0 / (1.0 + np.exp(-(z / tau)))
    return {"adj_base": x, "median": med, "mad": mad, "iqr": iqr, "tau": tau,
            "adj_for_valley": sig, "valley_tau": max(iqr / 2.0, 0.06)}


def p95_breakpoints_ref(adj_sims, pct: float = 95.0) -> Tuple[float, np.ndarray]:
    """BASELINE.json config 3 rule: distance = 1 - adj; threshold = np.percentile(d, pct)
    (linear interpolation); breakpoints = indices with d > threshold."""
    d = 1.0 - np.asarray(adj_sims, dtype=float)
    if d.size == 0:
        return float("nan"), np.zeros(0, dtype=np.int64)
    thr = float(np.percentile(d, pct))
    return thr, np.nonzero(d > thr)[0]


def segmented_splitter_pass_ref(E: np.ndarray, offsets: np.ndarray, pct: float = 95.0):
    """Whole ragged batch: per document adjacent sims, P-th percentile distance threshold and
    breakpoint flags, concatenated in the layout the CUDA pass emits (one slot per row; the
    last row of each document carries adj = 0 / flag = 0)."""
    E = np.asarray(E)
    total = E.shape[0]
    adj = np.zeros(total, dtype=np.float32)
    flags = np.zeros(total, dtype=np.uint8)
    thr = np.full(len(offsets) - 1, np.nan, dtype=np.float64)
    for di, (a, b) in enumerate(zip(offsets[:-1], offsets[1:])):
        a, b = int(a), int(b)
        if b - a < 2:
            continue
        sims = adjacent_sims_ref(E[a:b])
        adj[a:b - 1] = sims.astype(np.float32)
        t, bp = p95_breakpoints_ref(sims, pct)
        thr[di] = t
        flags[a + bp] = 1
    return adj, thr, flags


def c99_similarity_ref(En: np.ndarray) -> np.ndarray:
    """Splitter:169 — ``S = embs @ embs.T`` on already-normalised embeddings."""
    En = np.asarray(En)
    return En @ En.T


def c99_global_rank_ref(S: np.ndarray) -> np.ndarray:
    """Splitter:189-192 — R[i,j] = #{k: S[i,k] < S[i,j]} + #{k: S[k,j] < S[i,j]} as fp32."""
    S = np.asarray(S)
    row_less = (S[:, None, :] < S[:, :, None]).sum(axis=2).astype(np.int32)
    col_less = (S.T[:, None, :] < S.T[:, :, None]).sum(axis=2).astype(np.int32).T
    return (row_less + col_less).astype(np.float32)


def c99_local_rank_ref(S: np.ndarray, mask_size: int = 11) -> np.ndarray:
    """Splitter:171-186 — fraction of the clipped m x m window strictly below S[i,j]."""
    S = np.asarray(S)
    n = S.shape[0]
    m = max(3, int(mask_size) | 1)
    half = m // 2
    R = np.zeros_like(S, dtype=np.float32)
    for i in range(n):
        i0, i1 = max(0, i - half), min(n, i + half + 1)
        for j in range(n):
            j0, j1 = max(0, j - half), min(n, j + half + 1)
            win = S[i0:i1, j0:j1]
            R[i, j] = float((win < S[i, j]).sum()) / float(win.size if win.size else 1)
    return R


def _block_mean_f32(R: np.ndarray, a: int, b: int, default: float) -> float:
    """``float(R[a:b, a:b].mean())`` — an fp32 numpy mean widened, as in Splitter:215-222."""
    blk = R[a:b, a:b]
    return float(blk.mean()) if blk.size > 0 else default


def c99_inside_density_ref(R: np.ndarray, segments) -> float:
    """Splitter:194-204 — summed fp32 block sums over summed block areas."""
    tot, area = 0.0, 0
    for a, b in segments:
        if b > a:
            blk = R[a:b, a:b]
            tot += float(blk.sum()) if blk.size > 0 else 0.0
            area += (b - a) * (b - a)
    return tot / float(area) if area > 0 else 0.0


def c99_divisive_ref(R: np.ndarray, min_chunk_size: int = 3, max_cuts=None, min_gain: float = 0.01, stopping: str = "gain",
                     knee_c: float = 1.2, smooth_window: int = 3):
    """Splitter:205-264 — greedy divisive search.  Returns ``(boundaries, cuts_in_pick_order, D_series)``.

    Every round scans the segment list in list order (a split pops the segment and appends its two halves at
    the END of the list, :232-233) and every admissible cut ``a+m <= c <= b-m`` in ascending order; strict
    ``>`` keeps the first best.  ``gain = 0.5 * (mean(left) + mean(right)) - mean(whole)`` with fp32 block
    means.  Stop: no candidate, ``max_cuts`` reached, or (``stopping == 'gain'``) best gain below
    ``max(min_gain, 0.1 * |mean(whole of the best segment)|)``.  ``'profile'`` keeps cutting and then keeps
    the cuts before the first smoothed density increment below ``mean - knee_c * std`` (:239-264).
    """
    R = np.asarray(R)
    n = R.shape[0]
    m = int(min_chunk_size)
    if n < 2 * m:
        return [], [], []
    mode = stopping.lower()
    seg_list = [(0, n)]
    picked = []
    series = [c99_inside_density_ref(R, seg_list)]
    while True:
        top = (-1e9, None, None, 0.0)  # gain, cut, list index, mean of the whole segment
        for li, (a, b) in enumerate(seg_list):
            if b - a < 2 * m:
                continue
            whole = _block_mean_f32(R, a, b, 0.0)
            for c in range(a + m, b - m + 1):
                g = 0.5 * (_block_mean_f32(R, a, c, whole) + _block_mean_f32(R, c, b, whole)) - whole
                if g > top[0]:
                    top = (g, c, li, whole)
        g_best, cut, li, whole = top
        floor = max(float(min_gain), 0.1 * abs(whole))
        if cut is None or (max_cuts is not None and len(picked) >= int(max_cuts)):
            break
        if mode == "gain" and g_best < floor:
            break
        a, b = seg_list.pop(int(li))
        seg_list.extend([(a, cut), (cut, b)])
        picked.append(int(cut))
        series.append(c99_inside_density_ref(R, sorted(seg_list)))
    if mode != "profile" or not picked:
        return sorted(set(picked)), picked, series
    return c99_profile_knee_ref(picked, series, knee_c, smooth_window), picked, series


def c99_profile_knee_ref(picked, series, knee_c: float = 1.2, smooth_window: int = 3):
    """Splitter:241-264 — keep the cuts made before the first sharp drop of the smoothed density increments."""
    steps = np.diff(np.array(series, dtype=float))
    if steps.size == 0:
        return sorted(set(picked))
    w = max(1, int(smooth_window))
    sm = np.convolve(steps, np.ones(w, dtype=float) / float(w), mode="same") if (w > 1 and steps.size >= w) else steps
    limit = float(sm.mean()) - float(knee_c) * float(sm.std() + 1e-9)
    for i, v in enumerate(sm, start=1):
        if v < limit:
            return sorted(set(picked[: min(max(1, i) - 1, len(picked))]))
    return sorted(set(picked))


# ---- float64 statement of the divisive search (exact block sums from a summed-area table): what the device kernel
# ---- K10 must reproduce bit for bit; the reference itself compares fp32 ndarray.mean() values (c99_divisive_ref above)
class BlockSumsF64:
    """Summed-area table over the rank matrix: any square block sum in O(1)."""

    def __init__(self, R: np.ndarray):
        n = R.shape[0]
        self.sat = np.zeros((n + 1, n + 1), dtype=np.float64)
        self.sat[1:, 1:] = np.cumsum(np.cumsum(R.astype(np.float64), axis=0), axis=1)

    def mean(self, a: int, b: int, default: float = 0.0) -> float:
        """Mean of R[a:b, a:b]; ``default`` for an empty block (reference :216-217)."""
        if b <= a:
            return default
        s = self.sat
        return float(s[b, b] - s[a, b] - s[b, a] + s[a, a]) / float((b - a) * (b - a))

    def total(self, a: int, b: int) -> float:
        s = self.sat
        return float(s[b, b] - s[a, b] - s[b, a] + s[a, a])


def c99_divisive_f64_ref(R: np.ndarray, min_chunk: int, max_cuts: Optional[int], min_gain: float, stopping: str, knee_c: float,
                   smooth_window: int) -> List[int]:
    """Reference :194-264 — repeatedly take the cut with the largest inside-density gain."""
    n = R.shape[0]
    blocks = BlockSumsF64(R)
    segs: List[Tuple[int, int]] = [(0, n)]
    cuts: List[int] = []

    def inside_density(segments) -> float:
        tot, area = 0.0, 0
        for a, b in segments:
            if b <= a:
                continue
            tot += blocks.total(a, b)
            area += (b - a) * (b - a)
        return tot / float(area) if area > 0 else 0.0

    profile = [inside_density(segs)]
    by_gain = stopping.lower() == "gain"
    while True:
        best_gain, best_pos, best_idx, best_mean_all = -1e9, None, None, 0.0
        for idx, (a, b) in enumerate(segs):
            if (b - a) < 2 * min_chunk:
                continue
            mean_all = blocks.mean(a, b)
            for c in range(a + min_chunk, b - min_chunk + 1):
                gain = 0.5 * (blocks.mean(a, c, mean_all) + blocks.mean(c, b, mean_all)) - mean_all
                if gain > best_gain:
                    best_gain, best_pos, best_idx, best_mean_all = gain, c, idx, mean_all
        thr = max(float(min_gain), 0.1 * abs(best_mean_all))
        if best_pos is None or (max_cuts is not None and len(cuts) >= int(max_cuts)):
            break
        if by_gain and best_gain < thr:
            break
        a, b = segs.pop(int(best_idx))
        segs += [(a, best_pos), (best_pos, b)]
        cuts.append(int(best_pos))
        profile.append(inside_density(sorted(segs)))
    if stopping.lower() != "profile" or not cuts:
        return sorted(set(cuts))
    return c99_profile_knee_ref(cuts, profile, knee_c, smooth_window)


