"""Drop-in for the reference's ``Method/semantic_common.py`` on the dense-similarity path.

Same public names and call signatures; the arithmetic runs in libsemsearch_b200's CUDA kernels:

* ``create_similarity_matrix``          -> K3 ``ss_segmented_simmatrix``  (reference :144-191)
* ``analyze_similarity_distribution``   -> ``ss_similarity_distribution`` (reference :250-270)
* ``similarity_matrices_from_embeddings`` (new) batches many documents into one launch.

Differences from the reference, by mandate: only CUDA is honoured (no XLA / DirectML / numpy
branch) and a CUDA failure raises instead of silently falling back to ``embs @ embs.T``.
The OpenIE string helpers of the reference module are outside the hot path and not provided.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..Tool import Sentence_Embedding as _embedding


def embed_text_list(text_list, model_name, batch_size: int = 32, device_preference=None):
    """Module-level encoder hook, patched exactly like the reference's (semantic_common.py:28)."""
    return _embedding.sentence_embedding(text_list, model_name=model_name, batch_size=batch_size,
                                         device_preference=device_preference)


# ---------------- Device & batch utilities ---------------- #

def normalize_device(device: Optional[object]) -> str:
    """Reference :41-61.  Only ``"cuda"`` exists here; anything else maps to ``"cuda"`` when a GPU
    is present so that reference configs (``"dml"``, ``"tpu"``, ``"auto"``) keep working."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("semanticsearch_b200 requires a CUDA device (there is no CPU fallback)")
    return "cuda"


def estimate_optimal_batch_size(sentences: List[str], base_batch_size: int, device: str) -> int:
    """Encoder batch-size heuristic (reference :63-77, CUDA branch)."""
    if not sentences:
        return base_batch_size
    avg_len = sum(len(s) for s in sentences) / max(1, len(sentences))
    for limit, cap, floor in ((50, 256, 64), (100, 128, 32), (200, 64, 16)):
        if avg_len < limit:
            return min(cap, max(base_batch_size, floor), len(sentences))
    return min(32, max(base_batch_size, 8), len(sentences))


optimize_gpu_batch_size = estimate_optimal_batch_size


def embed_sentences_batched(sentences: List[str], model_name: str, base_batch_size: int = 32,
                            device: Optional[object] = None, silent: bool = True) -> np.ndarray:
    """Reference :84-140: batch the sentences through the encoder hook, halve on OOM."""
    if not sentences:
        return np.array([])
    batch_size = max(1, estimate_optimal_batch_size(sentences, base_batch_size, "cuda"))
    chunks: List[np.ndarray] = []
    start = 0
    cur = batch_size
    while start < len(sentences):
        batch = sentences[start:start + cur]
        try:
            emb = embed_text_list(batch, model_name=model_name, batch_size=max(1, cur), device_preference="cuda")
        except Exception as exc:  # OOM back-off only; everything else propagates
            text = str(exc).lower()
            if "out of memory" in text and cur > 1:
                cur = max(1, cur // 2)
                log_msg(silent, f"[embed] OOM, sub-batch -> {cur}", "warning", "common")
                continue
            raise
        if emb is not None and len(emb) > 0:
            chunks.append(np.asarray(emb))
        start += len(batch)
    if not chunks:
        return np.array([])
    return np.vstack(chunks)


# ---------------- Similarity matrix ---------------- #

def pack_document_rows(docs: Sequence) -> "torch.Tensor":
    """Concatenate per-document embedding matrices into one CUDA float32 ``[rows, dim]`` tensor.  Encoder outputs that
    already live in HBM (``SentenceTransformer.encode(convert_to_tensor=True)``) are concatenated on the device — the
    embedding hand-off of SURVEY.md section 8f rank 4, no numpy round trip (reference :107-114); host arrays take one H2D copy."""
    import torch
    if all(isinstance(e, torch.Tensor) and e.is_cuda for e in docs):
        return torch.cat([e.to(torch.float32) for e in docs], dim=0).contiguous()
    rows = [np.ascontiguousarray(e.detach().cpu().numpy() if isinstance(e, torch.Tensor) else e, dtype=np.float32) for e in docs]
    return torch.from_numpy(np.concatenate(rows, axis=0)).cuda()


def similarity_matrices_from_embeddings(doc_embeddings: Sequence[np.ndarray]) -> List[Optional[np.ndarray]]:
    """Batched core of ``create_similarity_matrix``: one kernel launch for all documents.

    ``doc_embeddings[d]`` is ``n_d x dim`` float32 (un-normalised).  Returns one fresh ``n x n``
    float32 ndarray per document, ``None`` for documents with fewer than 2 sentences
    (reference :151-152)."""
    import torch
    from .. import ragged
    sizes = [int(e.shape[0]) if e is not None and getattr(e, "ndim", 0) == 2 else 0 for e in doc_embeddings]
    live = [d for d, n in enumerate(sizes) if n >= 2]
    out: List[Optional[np.ndarray]] = [None] * len(sizes)
    if not live:
        return out
    dim = int(doc_embeddings[live[0]].shape[1])
    if any(int(doc_embeddings[d].shape[1]) != dim for d in live):
        raise ValueError("all documents of one batch must share the embedding dimension")
    plan = ragged.make_plan([sizes[d] for d in live], "cuda")
    E = pack_document_rows([doc_embeddings[d] for d in live])
    S = ragged.segmented_simmatrix(E, plan, validate=True).cpu().numpy()
    for slot, d in enumerate(live):
        n = sizes[d]
        out[d] = S[plan.s_offsets[slot]:plan.s_offsets[slot + 1]].reshape(n, n).copy()
    return out


def create_similarity_matrix(sentences: List[str], model_name: str, batch_size: int = 32,
                             device: Optional[str] = "cuda", silent: bool = True) -> Optional[np.ndarray]:
    """Same contract as the reference (:144-191): ``None`` for fewer than 2 sentences or when the
    encoder returns a mismatching number of rows; otherwise a caller-owned ``n x n`` float32 matrix
    of cosine similarities."""
    if len(sentences) < 2:
        return None
    normalize_device(device)
    embs = embed_sentences_batched(sentences, model_name, base_batch_size=batch_size, device="cuda", silent=silent)
    if embs is None or embs.size == 0 or embs.shape[0] != len(sentences):
        return None
    return similarity_matrices_from_embeddings([np.asarray(embs, dtype=np.float32)])[0]


def split_indices_by_diameter(doc_embeddings: Sequence[np.ndarray], threshold: float) -> List[List[Tuple[int, int]]]:
    """Batched device form of the controller's ``_split_indices_by_diameter`` / ``_enforce_diameter_on_tuples``
    (data_process/simple_chunk_controller.py:571-625): for every document (``n x dim`` embeddings, un-normalised) the
    list of ``(start, end)`` sentence spans whose diameter ``1 - min off-diagonal cosine`` does not exceed ``threshold``,
    obtained by cutting at the lowest adjacent similarity.  Documents with fewer than two sentences keep one span."""
    import torch
    from .. import ragged
    sizes = [int(e.shape[0]) if e is not None and getattr(e, "ndim", 0) == 2 else 0 for e in doc_embeddings]
    out: List[List[Tuple[int, int]]] = [[(0, n)] if n > 0 else [] for n in sizes]
    live = [d for d, n in enumerate(sizes) if n >= 2]
    if not live:
        return out
    rows = [np.ascontiguousarray(doc_embeddings[d], dtype=np.float32) for d in live]
    plan = ragged.make_plan([r.shape[0] for r in rows], "cuda")
    S = ragged.segmented_simmatrix(torch.from_numpy(np.concatenate(rows, axis=0)).cuda(), plan, validate=True)
    ends, n_spans, _ = ragged.diameter_split(S, plan, float(threshold))
    ends_h, n_h = ends.cpu().numpy(), n_spans.cpu().numpy()
    for slot, d in enumerate(live):
        e = [int(x) for x in ends_h[plan.offsets[slot]: plan.offsets[slot] + int(n_h[slot])]]
        out[d] = list(zip([0] + e[:-1], e))
    return out


def analyze_similarity_distribution(sim_matrix) -> Optional[Dict[str, float]]:
    """Reference :250-270, computed on the device (strict upper triangle, values >= 1-1e-5 dropped,
    min/max/mean/std + percentiles 10/25/50/75/80/85/90/95)."""
    if not isinstance(sim_matrix, np.ndarray) or sim_matrix.ndim != 2 or sim_matrix.shape[0] < 2:
        return None
    import torch
    from .. import ragged
    n = sim_matrix.shape[0]
    plan = ragged.make_plan([n], "cuda")
    S = torch.from_numpy(np.ascontiguousarray(sim_matrix, dtype=np.float32).reshape(-1)).cuda()
    st = ragged.similarity_distribution(S, plan).cpu().numpy()[0]
    if st[0] < 0:
        return None
    return {key: float(st[1 + i]) for i, key in enumerate(ragged.STAT_KEYS)}


__all__ = [
    "normalize_device", "estimate_optimal_batch_size", "optimize_gpu_batch_size", "embed_sentences_batched",
    "create_similarity_matrix", "similarity_matrices_from_embeddings", "analyze_similarity_distribution", "split_indices_by_diameter",
    "init_logger", "log_msg",
]

# ---------------- Logging ---------------- #

_LEVELS = {"debug": logging.DEBUG, "info": logging.INFO, "warn": logging.WARNING, "warning": logging.WARNING,
           "error": logging.ERROR}


def init_logger(name: str = "semantic", level: int = logging.INFO) -> logging.Logger:
    logger = logging.getLogger(name)
    if not logger.handlers:
        logger.setLevel(level)
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter("[%(asctime)s][%(levelname)s][%(name)s] %(message)s", "%H:%M:%S"))
        logger.addHandler(handler)
        logger.propagate = False
    return logger


def log_msg(silent: bool, msg: str, level: str = "info", component: str = None):
    """Reference :313-319 — nothing is emitted when ``silent`` is true."""
    if silent:
        return
    name = "semantic" if component is None else f"semantic.{component}"
    init_logger(name).log(_LEVELS.get(level.lower(), logging.INFO), msg)
