"""TEST INFRASTRUCTURE — oracle for the chunk-ranking cosine + ordering stage.

Restates, in plain numpy, what ``Tool/rank_chunks_optimized.py:215-216`` computes through
scikit-learn's ``cosine_similarity`` (third-party, not vendored in the reference; version
unpinned there, 1.9.0 in this image) followed by ``np.argsort(-scores)`` (``:225,232``).

scikit-learn's published algorithm (``sklearn.metrics.pairwise.cosine_similarity`` →
``sklearn.preprocessing.normalize(norm="l2")`` → ``safe_sparse_dot``):
  1. ``norms = sqrt(einsum('ij,ij->i', X, X))`` in the input dtype (fp32 stays fp32);
  2. rows whose norm is exactly 0 are divided by 1.0 instead (``_handle_zeros_in_scale``);
  3. both operands are normalised *first* (copies), then ``Xn @ Yn.T``.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def l2_normalize_rows_sklearn(X: np.ndarray) -> np.ndarray:
    """Row L2 normalisation with sklearn's zero rule (norm 0 -> divide by 1)."""
    X = np.asarray(X)
    norms = np.sqrt(np.einsum("ij,ij->i", X, X))
    norms = np.where(norms == 0.0, np.asarray(1.0, dtype=norms.dtype), norms)
    return X / norms[:, None]


def cosine_similarity_ref(Q: np.ndarray, C: np.ndarray) -> np.ndarray:
    """``cosine_similarity(Q, C)`` as at rank_chunks_optimized.py:216 (returns ``B x N``)."""
    Q = np.atleast_2d(np.asarray(Q))
    C = np.atleast_2d(np.asarray(C))
    return l2_normalize_rows_sklearn(Q) @ l2_normalize_rows_sklearn(C).T


def rank_order_ref(scores: np.ndarray) -> np.ndarray:
    """Full descending order, ``np.argsort(-cosine_scores)`` (rank_chunks_optimized.py:225,232).

    The reference uses numpy's default *non-stable* sort, so the order inside a run of equal
    scores is unspecified there; this restatement fixes it to "lower index first" (stable).
    """
    return np.argsort(-np.asarray(scores), kind="stable")


def cosine_topk_ref(Q: np.ndarray, C: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k of each query row: (scores ``B x k`` fp32-or-input dtype, indices ``B x k`` int64).

    One reference call per query row (rank_chunks_optimized.py:215-216 is called with a
    ``1 x d`` query), top-k = first k entries of the descending order.
    """
    Q = np.atleast_2d(np.asarray(Q))
    C = np.atleast_2d(np.asarray(C))
    k_eff = min(int(k), C.shape[0])
    Cn = l2_normalize_rows_sklearn(C)
    out_s = np.zeros((Q.shape[0], k_eff), dtype=np.result_type(Q.dtype, C.dtype))
    out_i = np.zeros((Q.shape[0], k_eff), dtype=np.int64)
    for b in range(Q.shape[0]):
        s = (l2_normalize_rows_sklearn(Q[b:b + 1]) @ Cn.T)[0]
        order = rank_order_ref(s)[:k_eff]
        out_s[b] = s[order]
        out_i[b] = order
    return out_s, out_i


def cosine_fp64(Q: np.ndarray, C_rows: np.ndarray) -> np.ndarray:
    """fp64 cosine of each query against selected corpus rows — used ONLY to adjudicate ties."""
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
    C_rows = np.atleast_2d(np.asarray(C_rows, dtype=np.float64))
    qn = np.linalg.norm(Q, axis=1, keepdims=True)
    cn = np.linalg.norm(C_rows, axis=1, keepdims=True)
    qn[qn == 0] = 1.0
    cn[cn == 0] = 1.0
    return (Q / qn) @ (C_rows / cn).T


def check_topk_against_oracle(Q, C, got_scores, got_idx, k, score_tol, tie_tol=1e-5):
    """Parity rule from BASELINE.json:north_star.

    * every returned score is within ``score_tol`` of the oracle cosine of the returned row;
    * returned indices equal the oracle's, except where the swapped rows' fp64 cosines differ
      by less than ``tie_tol`` (ties), in which case either member of the tie is accepted.
    Returns a list of human-readable violations (empty == pass).
    """
    Q = np.atleast_2d(np.asarray(Q, dtype=np.float32))
    C = np.asarray(C, dtype=np.float32)
    ref_s, ref_i = cosine_topk_ref(Q, C, k)
    problems = []
    got_scores = np.asarray(got_scores)
    got_idx = np.asarray(got_idx)
    if got_idx.shape != ref_i.shape:
        return [f"shape mismatch {got_idx.shape} vs {ref_i.shape}"]
    for b in range(Q.shape[0]):
        if len(set(got_idx[b].tolist())) != got_idx.shape[1]:
            problems.append(f"q{b}: duplicate indices {got_idx[b].tolist()}")
            continue
        exact = cosine_fp64(Q[b:b + 1], C[got_idx[b]])[0]
        err = np.abs(exact - got_scores[b].astype(np.float64))
        if err.max(initial=0.0) > score_tol:
            problems.append(f"q{b}: score err {err.max():.3e} > {score_tol}")
        if np.any(np.diff(got_scores[b].astype(np.float64)) > 0):
            problems.append(f"q{b}: scores not descending")
        if np.array_equal(got_idx[b], ref_i[b]):
            continue
        ref_exact = cosine_fp64(Q[b:b + 1], C[ref_i[b]])[0]
        for pos in range(ref_i.shape[1]):
            if got_idx[b, pos] == ref_i[b, pos]:
                continue
            # positional mismatch is fine iff the two candidates tie within tie_tol in fp64
            if abs(exact[pos] - ref_exact[pos]) >= tie_tol:
                problems.append(
                    f"q{b} pos{pos}: idx {got_idx[b, pos]} ({exact[pos]:.7f}) vs oracle "
                    f"{ref_i[b, pos]} ({ref_exact[pos]:.7f})")
    return problems
