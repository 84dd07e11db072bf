"""TEST INFRASTRUCTURE — oracle for the per-document sentence x sentence similarity matrix.

Restates the arithmetic of ``create_similarity_matrix`` (Method/semantic_common.py:144-191)
*after* the encoder call: the embeddings ``E`` (``n x d`` fp32) are the input here because the
transformer encoder is outside the hot path.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np


def normalize_rows_1e9(E: np.ndarray) -> np.ndarray:
    """Row L2 normalise with the reference's zero rule: norm==0 -> 1e-9.

    semantic_common.py:158-160 and Semantic_Splitter_Optimized.py:150-152 (identical code).
    """
    E = np.asarray(E)
    norms = np.linalg.norm(E, axis=1, keepdims=True)
    norms[norms == 0] = 1e-9
    return E / norms


def similarity_matrix_ref(E: np.ndarray) -> Optional[np.ndarray]:
    """``S = En @ En.T`` (semantic_common.py:186-191; the CUDA branch ``:163-164`` is the same
    product in fp32 SGEMM).  Returns ``None`` for fewer than 2 rows (``:151-152``)."""
    E = np.asarray(E)
    if E.ndim != 2 or E.shape[0] < 2:
        return None
    En = normalize_rows_1e9(E)
    return En @ En.T


def segmented_similarity_ref(E: np.ndarray, offsets: np.ndarray) -> List[Optional[np.ndarray]]:
    """Per-document matrices for a ragged batch given CSR-style row offsets."""
    out = []
    for a, b in zip(offsets[:-1], offsets[1:]):
        out.append(similarity_matrix_ref(E[int(a):int(b)]))
    return out


PCTS = (10, 25, 50, 75, 80, 85, 90, 95)
STAT_KEYS = ("min", "max", "mean", "std") + tuple(f"p{p}" for p in PCTS)


def analyze_similarity_distribution_ref(S) -> Optional[Dict[str, float]]:
    """semantic_common.py:250-270: stats of the strict upper triangle, dropping values
    >= 1 - 1e-5; percentiles via ``np.percentile`` (linear interpolation)."""
    if not isinstance(S, np.ndarray) or S.ndim != 2 or S.shape[0] < 2:
        return None
    iu = np.triu_indices_from(S, k=1)
    sims = S[iu]
    kept = sims[sims < (1.0 - 1e-5)]
    if kept.size == 0:
        if sims.size > 0:
            mx = float(np.max(sims))
            return {key: mx for key in STAT_KEYS}
        return None
    out = {
        "min": float(np.min(kept)),
        "max": float(np.max(kept)),
        "mean": float(np.mean(kept)),
        "std": float(np.std(kept)),
    }
    for p in PCTS:
        out[f"p{p}"] = float(np.percentile(kept, p))
    return out


def split_indices_by_diameter_ref(sim_matrix: np.ndarray, start: int, end: int, threshold: float):
    """Restatement of the controller's ``_split_indices_by_diameter`` (data_process/simple_chunk_controller.py:571-594):
    a span whose diameter ``1 - min off-diagonal similarity`` exceeds ``threshold`` is cut after the first minimum of its
    adjacent similarities and both halves are split recursively, left first."""
    length = end - start
    if length <= 1:
        return [(start, end)]
    sub = sim_matrix[start:end, start:end]
    if length == 2:
        diam = 1.0 - float(sub[0, 1])
    else:
        mask = ~np.eye(length, dtype=bool)
        diam = 1.0 - float(sub[mask].min())
    if diam <= threshold:
        return [(start, end)]
    adj = [float(sim_matrix[i, i + 1]) for i in range(start, end - 1)]
    if not adj:
        return [(start, end)]
    cut = start + int(np.argmin(np.array(adj))) + 1
    return split_indices_by_diameter_ref(sim_matrix, start, cut, threshold) + split_indices_by_diameter_ref(sim_matrix, cut, end, threshold)
