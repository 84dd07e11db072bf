"""TEST INFRASTRUCTURE — oracle for the semantic-grouping "threshold pass".

Restates the data-parallel part of ``semantic_grouping_main``
(Method/Semantic_Grouping_Optimized.py): sigmoid sharpening ``:100-108``, diagonal removal
``:110-113``, centrality ``:115``, quantile thresholds on positive entries ``:351-355``,
``:458-462``, ``:534-537``, reassign delta ``:559-561``, and ``_build_knn_graph`` ``:270-283``.
The sequential clustering that follows (eigh, k-means, Louvain, split/merge) is host logic
and is not restated here.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np


def sharpen_ref(S: np.ndarray, tau: Optional[float] = None) -> Tuple[np.ndarray, np.ndarray, float, float]:
    """(sim_sharp fp32 with zero diagonal, centrality fp64, mu, sigma).

    Grouping:102-106: mu/sigma are numpy fp32 reductions over ALL n*n entries (diagonal
    included) widened to Python floats; the sigmoid stays in the matrix dtype (fp32).
    Grouping:111 zeroes the diagonal; ``:115`` centrality = row sum / max(n-1, 1) as fp64.
    """
    S = np.asarray(S)
    n = S.shape[0]
    mu = float(np.mean(S))
    sigma = float(np.std(S) + 1e-9)
    t = 0.15 if tau is None else float(tau)
    z = (S - mu) / sigma
    with np.errstate(over="ignore"):
        sharp = 1.0 / (1.0 + np.exp(-(z / t)))
    np.fill_diagonal(sharp, 0.0)
    centrality = (sharp.sum(axis=1) / max(n - 1, 1)).astype(float)
    return sharp, centrality, mu, sigma


def positive_values(sharp: np.ndarray) -> np.ndarray:
    """``vals = arr[arr > 0.0]`` after widening to fp64 (Grouping:353-354)."""
    arr = np.asarray(sharp, dtype=float)
    return arr[arr > 0.0]


def thresholds_ref(sharp: np.ndarray, *, edge_floor_default: float = 0.4, tau_merge: float = 0.38,
                   reassign_delta: float = 0.02) -> Dict[str, float]:
    """The four data-derived thresholds of the auto-parameter path."""
    vals = positive_values(sharp)
    if vals.size:
        return {
            "edge_floor": float(np.quantile(vals, 0.80)),       # Grouping:355
            "tau_merge": float(np.quantile(vals, 0.65)),        # Grouping:462
            "global_merge_thr": float(np.quantile(vals, 0.60)),  # Grouping:537
            "reassign_delta": float(np.std(vals)) * 0.1,         # Grouping:561
            "count": int(vals.size),
        }
    return {"edge_floor": float(edge_floor_default), "tau_merge": float(tau_merge),
            "global_merge_thr": 0.5, "reassign_delta": float(reassign_delta), "count": 0}


def k_eff_auto(n: int) -> int:
    """Grouping:347 — ``max(5, min(32, round(0.06 n)))`` with Python's banker's rounding."""
    return int(max(5, min(32, round(n * 0.06))))


def knn_graph_ref(sharp: np.ndarray, kk: int, floor: float) -> np.ndarray:
    """``_build_knn_graph`` (Grouping:270-283): per row keep the top-(k_eff+1) entries (minus
    self) whose value is >= floor, then symmetrise with max.  Ties are broken lower-index
    first here (the reference's non-stable argsort leaves them unspecified)."""
    S = np.asarray(sharp)
    nn = S.shape[0]
    W = np.zeros((nn, nn), dtype=float)
    k_eff = int(max(1, min(kk, nn - 1)))
    for i in range(nn):
        order = np.argsort(-S[i], kind="stable")[: k_eff + 1]
        for j in order:
            j = int(j)
            if j == i:
                continue
            v = float(S[i, j])
            if v >= floor:
                W[i, j] = v
    return np.maximum(W, W.T)


def knn_lists_ref(sharp: np.ndarray, kk: int) -> Tuple[np.ndarray, np.ndarray]:
    """Neighbour lists before the floor filter: (idx ``n x (k_eff+1)`` int32, val fp32),
    ordered value-descending / index-ascending — the layout the CUDA pass emits."""
    S = np.asarray(sharp)
    nn = S.shape[0]
    k_eff = int(max(1, min(kk, nn - 1)))
    width = min(k_eff + 1, nn)
    idx = np.zeros((nn, width), dtype=np.int32)
    val = np.zeros((nn, width), dtype=S.dtype)
    for i in range(nn):
        order = np.argsort(-S[i], kind="stable")[:width]
        idx[i] = order
        val[i] = S[i, order]
    return idx, val


def grouping_pass_ref(S: np.ndarray, tau: Optional[float] = None):
    """Everything the device pass hands to the host clustering stage for one document."""
    sharp, centrality, mu, sigma = sharpen_ref(S, tau)
    thr = thresholds_ref(sharp)
    n = S.shape[0]
    k_all = k_eff_auto(n)
    W = knn_graph_ref(sharp, k_all, thr["edge_floor"])
    return {"sim_sharp": sharp, "centrality": centrality, "mu": mu, "sigma": sigma,
            "thresholds": thr, "k_eff_all": k_all, "W_all": W}


def knn_graph_mismatches(W_a: np.ndarray, W_b: np.ndarray, sharp: np.ndarray, kk: int, tie_tol: float = 0.0):
    """Compare two kNN graphs allowing the reference's unspecified tie order.

    An entry (i, j) may differ only if ``sharp[i, j]`` ties (within ``tie_tol``) with the
    cut-off value — the (k_eff+1)-th largest entry — of row i or of row j, i.e. the argsort
    at Grouping:276 could legitimately have picked either candidate.  Returns the list of
    (i, j) positions that differ for any other reason.
    """
    S = np.asarray(sharp)
    nn = S.shape[0]
    k_eff = int(max(1, min(kk, nn - 1)))
    width = min(k_eff + 1, nn)
    cut = np.sort(S, axis=1)[:, ::-1][:, width - 1]
    bad = []
    for i, j in zip(*np.nonzero(np.asarray(W_a) != np.asarray(W_b))):
        v = S[i, j]
        v2 = S[j, i]
        if abs(float(v) - float(cut[i])) <= tie_tol or abs(float(v2) - float(cut[j])) <= tie_tol:
            continue
        bad.append((int(i), int(j)))
    return bad


def block_sums_ref(sharp: np.ndarray, groups) -> Tuple[np.ndarray, np.ndarray]:
    """numpy statement of ss_group_block_sums for one document: ``rowsum[x, g] = sum(sharp[x, members_g])``,
    ``block[a, b] = sum(sharp[np.ix_(members_a, members_b)])`` in float64 (member lists are multisets).  The reference
    reaches the same sums element by element inside ``_mean_between`` / ``_mean_within`` (Grouping:118-130) and the
    reassignment loop (:566-588)."""
    S = np.asarray(sharp, dtype=np.float64)
    n, g = S.shape[0], len(groups)
    rowsum = np.zeros((n, g), dtype=np.float64)
    block = np.zeros((g, g), dtype=np.float64)
    for j, members in enumerate(groups):
        idx = np.asarray(list(members), dtype=np.int64)
        if idx.size:
            rowsum[:, j] = S[:, idx].sum(axis=1)
    for i, members in enumerate(groups):
        idx = np.asarray(list(members), dtype=np.int64)
        if idx.size:
            block[i, :] = rowsum[idx, :].sum(axis=0)
    return rowsum, block
