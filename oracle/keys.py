"""TEST INFRASTRUCTURE — numpy twin of the packed top-k key format used between kernels/ranks
(include/semsearch_b200.h: high 32 bits = order-preserving fp32 bits, low 32 bits =
0xFFFFFFFF - global row index; larger key = better, ties -> lower index; 0 = empty)."""
from __future__ import annotations

import numpy as np


def pack_keys(scores: np.ndarray, indices: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(scores, dtype=np.float32)
    b = s.view(np.uint32).astype(np.uint64)
    neg = (b >> np.uint64(31)) != 0
    ordered = np.where(neg, (~b) & np.uint64(0xFFFFFFFF), b | np.uint64(0x80000000))
    low = np.uint64(0xFFFFFFFF) - np.asarray(indices).astype(np.uint64)
    return ((ordered << np.uint64(32)) | low).view(np.int64)


def unpack_keys(keys: np.ndarray):
    k = np.ascontiguousarray(keys).view(np.uint64)
    ordered = (k >> np.uint64(32)).astype(np.uint32)
    pos = (ordered & np.uint32(0x80000000)) != 0
    bits = np.where(pos, ordered & np.uint32(0x7FFFFFFF), ~ordered)
    scores = bits.astype(np.uint32).view(np.float32)
    idx = (np.uint64(0xFFFFFFFF) - (k & np.uint64(0xFFFFFFFF))).astype(np.int64)
    empty = k == 0
    return np.where(empty, -np.inf, scores).astype(np.float32), np.where(empty, -1, idx)


def merge_keys(keys: np.ndarray, k_out: int) -> np.ndarray:
    """keys[P, B, k] -> [B, k_out]: the k_out largest keys per query, best first."""
    p, b, k = keys.shape
    flat = np.transpose(keys.view(np.uint64), (1, 0, 2)).reshape(b, p * k)
    order = np.argsort(flat, axis=1)[:, ::-1][:, :k_out]
    return np.take_along_axis(flat, order, axis=1).view(np.int64)
