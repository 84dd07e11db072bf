"""Drop-in for the reference's ``Tool/rank_chunks_optimized.py`` on the dense path.

Same public names and signatures (``cosine_similarity``, ``OptimizedRanker``,
``rank_and_filter_chunks_optimized``, ``rank_by_cosine_similarity`` / ``rank_by_bm25`` /
``rank_by_rrf``).  The dense stage runs on the GPU through the C ABI:

* ``cosine_similarity(X, Y)``                     -> ``ss_cosine_scores``   (reference :15,:215-216)
* ``np.argsort(-cosine_scores)`` + rank lookup    -> ``ss_rank_order``      (reference :225-235)
* ``OptimizedRanker.top_k(query, chunks, k)`` (new) -> fused cosine + top-k (K1/K2), no score vector

BM25 (lexical) stays on the host; ``rank_query_groups_batched`` fuses the RRF sum, the fused order and the
percentile thresholds of a whole block of query groups into one launch (``ss_segmented_rank_rrf``); the
per-query API keeps the RRF sum and the labelling on the host (SURVEY.md section 8f,
rank 3): they are O(N) bookkeeping on per-query groups.  Equal scores rank lower-index-first where
the reference's non-stable ``argsort`` leaves the order unspecified.
"""
from __future__ import annotations

import csv
import gc
import hashlib
import math
import os
from collections import Counter
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import pandas as pd

from . import Sentence_Embedding as _embedding
from .._lib import DeviceError as _DeviceError

try:
    csv.field_size_limit(2147483647)
except OverflowError:  # pragma: no cover
    csv.field_size_limit(2 ** 31 - 1)


# ----------------------------------------------------------------------------------------------
# Dense stage on the device
# ----------------------------------------------------------------------------------------------
def cosine_similarity(X, Y) -> np.ndarray:
    """``sklearn.metrics.pairwise.cosine_similarity(X, Y)`` for host arrays: rows are L2-normalised
    (zero rows stay zero) and ``Xn @ Yn.T`` is returned as a fresh float32 ``[len(X), len(Y)]`` array."""
    import torch
    from .. import similarity
    X = np.atleast_2d(np.asarray(X, dtype=np.float32))
    Y = np.atleast_2d(np.asarray(Y, dtype=np.float32))
    if X.shape[1] != Y.shape[1]:
        raise ValueError(f"Incompatible dimension for X and Y matrices: X.shape[1] == {X.shape[1]} while Y.shape[1] == {Y.shape[1]}")
    scores = similarity.cosine_scores(torch.from_numpy(np.ascontiguousarray(Y)).cuda(),
                                      torch.from_numpy(np.ascontiguousarray(X)).cuda())
    return scores.cpu().numpy()


def _device_rank_lookup(scores: np.ndarray):
    """(order, 1-based rank lookup) of one score vector, both on the device sort."""
    import torch
    from .. import similarity
    order, rank1 = similarity.rank_order(torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32)[None, :]).cuda())
    return order[0].cpu().numpy().astype(np.int64), rank1[0].cpu().numpy().astype(np.float64)


# ----------------------------------------------------------------------------------------------
# BM25 (lexical side channel) — Okapi BM25 as published, with the rank_bm25 epsilon floor
# ----------------------------------------------------------------------------------------------
class BM25Okapi:
    """Okapi BM25 (k1 = 1.5, b = 0.75) with rank_bm25's ``epsilon`` floor for negative idf values;
    the reference calls ``BM25Okapi(tokenized_chunks, epsilon=0.25)`` (:220)."""

    def __init__(self, corpus: List[List[str]], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.doc_freqs = [Counter(doc) for doc in corpus]
        self.doc_len = np.array([len(doc) for doc in corpus], dtype=float)
        self.corpus_size = len(corpus)
        self.avgdl = float(self.doc_len.sum()) / self.corpus_size if self.corpus_size else 0.0
        df: Counter = Counter()
        for freqs in self.doc_freqs:
            df.update(freqs.keys())
        self.idf: Dict[str, float] = {}
        idf_sum, negative = 0.0, []
        for word, freq in df.items():
            val = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
            self.idf[word] = val
            idf_sum += val
            if val < 0:
                negative.append(word)
        self.average_idf = idf_sum / len(self.idf) if self.idf else 0.0
        for word in negative:
            self.idf[word] = self.epsilon * self.average_idf

    def get_scores(self, query: List[str]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        for q in query:
            q_freq = np.array([freqs.get(q, 0) for freqs in self.doc_freqs], dtype=float)
            score += (self.idf.get(q) or 0) * (q_freq * (self.k1 + 1) /
                                               (q_freq + self.k1 * (1 - self.b + self.b * self.doc_len / self.avgdl)))
        return score


# ----------------------------------------------------------------------------------------------
# Helpers kept from the reference's public surface
# ----------------------------------------------------------------------------------------------
def estimate_memory_usage(file_path: str) -> Dict[str, float]:
    """Reference :28-50 — rough RAM estimate for a TSV (pandas x4, embeddings x1.7)."""
    if not Path(file_path).exists():
        return {"error": "File not found"}
    size = os.path.getsize(file_path)
    est_gb = size * 4.0 / (1024 ** 3)
    return {"file_size_mb": size / (1024 * 1024), "estimated_memory_gb": est_gb, "peak_memory_gb": est_gb * 1.7,
            "recommended_chunk_size": max(10000, min(100000, int(50000 / max(1, est_gb / 2))))}


_ALIASES = (("chunk_text", {"chunk_text", "passage", "text"}), ("query_text", {"query_text", "query", "question"}),
            ("query_id", {"query_id", "qid"}), ("chunk_id", {"chunk_id", "cid", "pid"}), ("label", {"label", "score", "target"}))


def _standardize_columns(df: pd.DataFrame, *, require_query_text: bool = True) -> pd.DataFrame:
    """Reference :70-104 — map column aliases onto the canonical names."""
    mapping = {}
    for canonical, names in _ALIASES:
        hit = next((c for c in (col.strip() for col in df.columns) if c.lower() in names), None)
        if hit is not None or canonical == "chunk_text":
            mapping[hit] = canonical
    return df.rename(columns=mapping)


def _f64_rank_surrogate(scores: np.ndarray) -> np.ndarray:
    """float32 values whose descending order (ties: lower row first) is exactly that of the float64 ``scores``.  The device
    rank kernels take fp32 keys; BM25 scores are float64 in the reference (:219-235) and can differ below fp32 resolution,
    so they are ranked here in float64 (stable) and handed over as ``-rank`` — exact in fp32 for up to 2^24 rows."""
    order = np.argsort(-np.asarray(scores, dtype=np.float64), kind="stable")
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order))
    return (-rank).astype(np.float32)


class DeviceRowStore:
    """HBM row store behind ``OptimizedRanker.chunk_embedding_cache`` (reference :116-139): one fp32 device matrix holds the
    embedding of every cached chunk, ``md5(text) -> slot`` lives on the host.  A chunk crosses the host link at most once —
    and not at all when the encoder returns CUDA tensors; ranking a query against cached chunks is a device-side row
    gather.  Eviction follows the reference: once more than ``cache_size`` chunks are cached, the oldest quarter goes."""

    def __init__(self, cache_size: int):
        self.cache_size = int(cache_size)
        self.slots: Dict[str, int] = {}      # insertion-ordered, like the reference's dict
        self.free: List[int] = []
        self.rows = None                     # torch float32 [capacity, dim] on the device
        self.h2d_rows = 0                    # chunk rows uploaded from host memory so far
        self.d2d_rows = 0                    # chunk rows taken over from CUDA tensors of the encoder

    def __len__(self) -> int:
        return len(self.slots)

    def __contains__(self, key: str) -> bool:
        return key in self.slots

    def keys(self):
        return self.slots.keys()

    def __getitem__(self, key: str) -> np.ndarray:
        return self.rows[self.slots[key]].cpu().numpy()

    def _reserve(self, n_new: int, dim: int, device):
        import torch
        need = len(self.slots) + n_new
        if self.rows is None:
            self.rows = torch.empty((max(1024, 2 * need), dim), dtype=torch.float32, device=device)
            self.free = list(range(self.rows.shape[0] - 1, -1, -1))
        elif self.rows.shape[1] != dim:
            raise RuntimeError(f"embedding width changed from {self.rows.shape[1]} to {dim}")
        if len(self.free) < n_new:
            old = self.rows
            grown = torch.empty((max(2 * old.shape[0], need + 1024), dim), dtype=torch.float32, device=old.device)
            grown[: old.shape[0]] = old
            self.free = list(range(grown.shape[0] - 1, old.shape[0] - 1, -1)) + self.free
            self.rows = grown

    def add(self, keys: List[str], vectors) -> None:
        """Cache ``vectors[i]`` under ``keys[i]`` (keys not yet cached; duplicates inside the call keep the last row)."""
        import torch
        if not keys:
            return
        on_device = isinstance(vectors, torch.Tensor) and vectors.is_cuda
        vec = vectors.to(torch.float32) if on_device else torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32))
        device = vec.device if on_device else torch.device("cuda", torch.cuda.current_device())
        self._reserve(len(keys), vec.shape[1], device)
        slots = []
        for key in keys:
            slot = self.slots.get(key)
            if slot is None:
                slot = self.free.pop()
                self.slots[key] = slot
            slots.append(slot)
        idx = torch.tensor(slots, dtype=torch.int64, device=self.rows.device)
        self.rows.index_copy_(0, idx, vec.to(self.rows.device, non_blocking=True))
        if on_device:
            self.d2d_rows += len(keys)
        else:
            self.h2d_rows += len(keys)

    def gather(self, keys: List[str]):
        """Device matrix ``[len(keys), dim]`` of the cached rows, in the order of ``keys``."""
        import torch
        idx = torch.tensor([self.slots[k] for k in keys], dtype=torch.int64, device=self.rows.device)
        return self.rows.index_select(0, idx)

    def evict_if_needed(self) -> None:
        """Reference :131-139: beyond ``cache_size`` entries, drop the oldest 25 %."""
        if len(self.slots) > self.cache_size:
            for key in list(self.slots.keys())[: len(self.slots) // 4]:
                self.free.append(self.slots.pop(key))


class OptimizedRanker:
    """Per-query hybrid ranker with md5-keyed embedding caches (reference :107-250).  The chunk cache is a
    ``DeviceRowStore``: chunk embeddings stay resident in HBM between queries."""

    def __init__(self, model_name: str = "all-MiniLM-L6-v2", device_preference: str = "dml", cache_size: int = 1000):
        self.model_name = model_name
        self.device_preference = device_preference if device_preference else "dml"
        self.cache_size = cache_size
        self.query_embedding_cache: Dict[str, np.ndarray] = {}
        self.chunk_embedding_cache = DeviceRowStore(cache_size)
        self.sentence_embedding = _embedding.sentence_embedding

    def _get_text_hash(self, text: str) -> str:
        return hashlib.md5(text.encode("utf-8")).hexdigest()

    def _manage_cache_size(self):
        self.chunk_embedding_cache.evict_if_needed()

    def get_query_embedding(self, query: str) -> np.ndarray:
        key = self._get_text_hash(query)
        if key in self.query_embedding_cache:
            return self.query_embedding_cache[key]
        emb = self.sentence_embedding(text_list=[query], model_name=self.model_name, device_preference=self.device_preference)
        if emb is None or emb.shape[0] == 0:
            raise RuntimeError(f"Failed to embed query: {query}")
        vec = emb[0].detach().float().cpu().numpy() if hasattr(emb, "detach") else np.asarray(emb[0])
        self.query_embedding_cache[key] = vec
        return vec

    def get_chunk_embeddings_device(self, chunks: List[str], batch_size: int = 32):
        """CUDA float32 ``[len(chunks), dim]`` matrix of the chunk embeddings: cached rows are gathered on the device, only
        chunks never seen before go through the encoder (and over the host link, unless the encoder returns CUDA tensors)."""
        store = self.chunk_embedding_cache
        keys = [self._get_text_hash(c) for c in chunks]
        seen, miss_keys, miss_text = set(), [], []
        for key, text in zip(keys, chunks):
            if key not in store and key not in seen:
                seen.add(key)
                miss_keys.append(key)
                miss_text.append(text)
        if miss_text:
            new = self.sentence_embedding(text_list=miss_text, model_name=self.model_name, batch_size=batch_size,
                                          device_preference=self.device_preference)
            if new is None or new.shape[0] != len(miss_text):
                raise RuntimeError("Failed to embed some chunks")
            store.add(miss_keys, new)
        rows = store.gather(keys)
        self._manage_cache_size()     # after the gather, like the reference (:196-199)
        return rows

    def get_chunk_embeddings_batch(self, chunks: List[str], batch_size: int = 32) -> np.ndarray:
        """Reference :161-199 — host ndarray ``[len(chunks), dim]`` (a device-to-host copy of the cached rows)."""
        return self.get_chunk_embeddings_device(chunks, batch_size).cpu().numpy()

    def _query_device(self, query: str):
        import torch
        return torch.from_numpy(np.ascontiguousarray(self.get_query_embedding(query), dtype=np.float32).reshape(1, -1)).cuda()

    def top_k(self, query: str, chunks: List[str], k: int = 10):
        """Fused device path for callers that only need the best k chunks: (scores, indices)."""
        from .. import similarity
        s, i = similarity.cosine_topk(self.get_chunk_embeddings_device(chunks).contiguous(), self._query_device(query),
                                      min(k, len(chunks)))
        return s[0].cpu().numpy(), i[0].cpu().numpy()

    def rank_single_query_optimized(self, query: str, chunks_df: pd.DataFrame, text_column: str = "chunk_text",
                                    id_column: str = "chunk_id") -> pd.DataFrame:
        """Reference :201-250: adds cosine_score / bm25_score / rrf_score and sorts by rrf desc."""
        from .. import similarity
        if text_column not in chunks_df.columns:
            raise ValueError(f"Text column '{text_column}' not found in DataFrame")
        chunks = chunks_df[text_column].fillna("").tolist()
        q_dev = self._query_device(query)
        c_dev = self.get_chunk_embeddings_device(chunks).contiguous()
        cos_dev = similarity.cosine_scores(c_dev, q_dev)                      # ss_cosine_scores on the resident rows
        _, cosine_rank_dev = similarity.rank_order(cos_dev)
        cosine_scores = cos_dev[0].cpu().numpy()
        cosine_rank = cosine_rank_dev[0].cpu().numpy().astype(np.float64)
        bm25 = BM25Okapi([c.lower().split() for c in chunks], epsilon=0.25)
        bm25_scores = np.maximum(bm25.get_scores(query.lower().split()), 0.0)
        _, bm25_rank = _device_rank_lookup(_f64_rank_surrogate(bm25_scores))   # float64 order, fp32 keys
        k = 60
        out = chunks_df.copy()
        out["cosine_score"] = cosine_scores
        out["bm25_score"] = bm25_scores
        out["rrf_score"] = 1.0 / (k + cosine_rank) + 1.0 / (k + bm25_rank)
        return out.sort_values(by="rrf_score", ascending=False).reset_index(drop=True)


def rank_query_groups_batched(ranker: "OptimizedRanker", groups: List[tuple], upper_percentile: float = 80,
                              lower_percentile: float = 20, text_column: str = "chunk_text"):
    """Device-batched form of the per-query loop (reference :488-536 calling :201-250 and :517-526).

    ``groups`` = ``[(query_text, chunks_df), ...]``.  Embeddings (cached encoder) and BM25 stay per group on the
    host; cosine scores, both rank lookups, the RRF sum, the fused order and the two percentile thresholds
    of ALL groups come from one ``ss_segmented_rank_rrf`` launch.  Returns one
    ``(ranked_df, pos_threshold, neg_threshold)`` per group, ``ranked_df`` exactly as
    ``rank_single_query_optimized`` builds it (ties: lower row first)."""
    import torch
    from .. import similarity
    if not groups:
        return []
    q_rows, c_rows, bm_rows, bm_keys, sizes = [], [], [], [], []
    for query, df in groups:
        if text_column not in df.columns:
            raise ValueError(f"Text column '{text_column}' not found in DataFrame")
        chunks = df[text_column].fillna("").tolist()
        q_rows.append(np.asarray(ranker.get_query_embedding(query), dtype=np.float32).reshape(-1))
        c_rows.append(ranker.get_chunk_embeddings_device(chunks))           # device rows: cached chunks never cross the link again
        bm25 = BM25Okapi([c.lower().split() for c in chunks], epsilon=0.25)
        bm_rows.append(np.maximum(bm25.get_scores(query.lower().split()), 0.0))
        bm_keys.append(_f64_rank_surrogate(bm_rows[-1]))                    # the reference ranks the float64 scores (:226-235)
        sizes.append(len(chunks))
    if max(sizes) > 8192:
        raise ValueError("a query group holds more than 8192 chunks; rank it with rank_single_query_optimized")
    offsets = np.zeros(len(sizes) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(sizes)
    out = similarity.segmented_rank_rrf(
        torch.cat(c_rows, dim=0).contiguous(), torch.from_numpy(offsets).cuda(),
        torch.from_numpy(np.stack(q_rows)).cuda(), torch.from_numpy(np.concatenate(bm_keys)).cuda(),
        k_rrf=60.0, upper_percentile=float(upper_percentile), lower_percentile=float(lower_percentile), max_group_rows=max(sizes))
    cos, rrf = out["cosine"].cpu().numpy(), out["rrf"].cpu().numpy()
    order, thr = out["order"].cpu().numpy(), out["thresholds"].cpu().numpy()
    results = []
    for gi, (_query, df) in enumerate(groups):
        a, b = int(offsets[gi]), int(offsets[gi + 1])
        ranked = df.copy()
        ranked["cosine_score"] = cos[a:b]
        ranked["bm25_score"] = bm_rows[gi]
        ranked["rrf_score"] = rrf[a:b]
        ranked = ranked.iloc[order[a:b]].reset_index(drop=True)
        results.append((ranked, float(thr[gi, 0]), float(thr[gi, 1])))
    return results


def _label_group(ranked: pd.DataFrame, upper_percentile: int, lower_percentile: int) -> Optional[pd.DataFrame]:
    """Reference :517-526 — keep rows at or above P_upper (label 1) or at or below P_lower (label 0)."""
    scores = ranked["rrf_score"].to_numpy(dtype=float)
    pos_thr, neg_thr = np.percentile(scores, upper_percentile), np.percentile(scores, lower_percentile)
    kept = ranked[(ranked["rrf_score"] >= pos_thr) | (ranked["rrf_score"] <= neg_thr)].copy()
    if kept.empty:
        return None
    kept["label"] = (kept["rrf_score"] >= pos_thr).astype(int)
    return kept


def _process_queries_sequential(df: pd.DataFrame, model_name: str, upper_percentile: int, lower_percentile: int) -> List[pd.DataFrame]:
    """Reference :488-536.  One ranker (one encoder, one CUDA context) serves every query group."""
    ranker = OptimizedRanker(model_name=model_name)
    kept = []
    groups = [(qid, group) for qid, group in df.groupby("query_id") if len(group) >= 2]
    batched = [(qid, group) for qid, group in groups if len(group) <= 8192]
    single = [(qid, group) for qid, group in groups if len(group) > 8192]
    from .._lib import DeviceError
    try:
        # every query group of the block in one device launch
        ranked_all = rank_query_groups_batched(ranker, [(str(g["query_text"].iloc[0]), g) for _qid, g in batched],
                                               upper_percentile, lower_percentile)
        for ranked, pos_thr, neg_thr in ranked_all:
            sel = ranked[(ranked["rrf_score"] >= pos_thr) | (ranked["rrf_score"] <= neg_thr)].copy()
            if not sel.empty:
                sel["label"] = (sel["rrf_score"] >= pos_thr).astype(int)
                kept.append(sel)
    except DeviceError:
        raise  # a failing kernel or CUDA error is never swallowed: there is no CPU fallback
    except Exception as exc:
        # one bad group (encoder failure, missing column, ...) must not take the block down: the reference isolates every
        # query (:488-536), so the groups are ranked one by one and only the failing ones are reported and skipped
        print(f"Batched ranking failed ({exc}); ranking the {len(batched)} groups of this block one by one")
        single = batched + single
    for query_id, group in single:
        try:
            ranked = ranker.rank_single_query_optimized(str(group["query_text"].iloc[0]), group)
            labelled = None if ranked.empty else _label_group(ranked, upper_percentile, lower_percentile)
            if labelled is not None:
                kept.append(labelled)
        except DeviceError:
            raise
        except Exception as exc:
            print(f"Error processing query {query_id}: {exc}")
    return kept


def rank_and_filter_chunks_optimized(chunks_tsv: str, output_dir: Path, original_tsv: str, upper_percentile: int = 80,
                                     lower_percentile: int = 20, model_name: str = "thenlper/gte-base",
                                     max_workers: int = None, chunk_size: int = 50000) -> str:
    """Reference :253-485: stream the chunk TSV, rank every query group, keep the labelled extremes
    and write ``<stem>_rrf_filtered.tsv`` (query_id, chunk_text, label).  Returns ``""`` on I/O
    failure.  ``max_workers`` is accepted for compatibility: the GPU path ranks the groups of a
    block back to back in this process instead of spawning one model per query (:563-580)."""
    if not Path(chunks_tsv).exists():
        return ""
    try:
        mapping = pd.read_csv(original_tsv, sep="\t", usecols=["query_id", "query_text"], quoting=csv.QUOTE_NONE,
                              engine="python", on_bad_lines="warn")
        query_map = dict(mapping.groupby("query_id")["query_text"].first())
    except Exception as exc:
        print(f"Error loading original TSV for query_text mapping: {exc}")
        return ""
    est = estimate_memory_usage(chunks_tsv)
    if "error" not in est and chunk_size > est["recommended_chunk_size"]:
        chunk_size = est["recommended_chunk_size"]
    try:
        head = _standardize_columns(pd.read_csv(chunks_tsv, sep="\t", nrows=5, on_bad_lines="warn", engine="python"),
                                    require_query_text=False)
        if "chunk_text" not in head.columns:
            return ""
    except Exception as exc:
        print(f"Error scanning file: {exc}")
        return ""
    kept: List[pd.DataFrame] = []
    for block in pd.read_csv(chunks_tsv, sep="\t", chunksize=chunk_size, on_bad_lines="warn", engine="python"):
        try:
            block.columns = block.columns.str.strip()
            block = _standardize_columns(block, require_query_text=False)
            if "query_text" not in block.columns and "query_id" in block.columns:
                block["query_text"] = block["query_id"].map(query_map).fillna("")
            kept.extend(_process_queries_sequential(block, model_name, upper_percentile, lower_percentile))
        except _DeviceError:
            raise  # device failures propagate to the caller (the reference would fall back to the CPU; this path has none)
        except Exception as exc:
            print(f"Error processing chunk: {exc}")
    if not kept:
        return ""
    final = pd.concat(kept, ignore_index=True)
    save_path = Path(output_dir) / f"{Path(chunks_tsv).stem}_rrf_filtered.tsv"
    final[["query_id", "chunk_text", "label"]].to_csv(save_path, sep="\t", index=False)
    return str(save_path)


# ---- legacy API (reference :658-705) ------------------------------------------------------------
def rank_by_cosine_similarity(query: str, chunks_df: pd.DataFrame, text_column: str = "chunk_text",
                              model_name: str = "all-MiniLM-L6-v2", batch_size: int = 32, *, device_preference: str = None) -> pd.DataFrame:
    ranker = OptimizedRanker(model_name=model_name, device_preference=device_preference)
    result = ranker.rank_single_query_optimized(query, chunks_df, text_column)
    cols = ["chunk_id", "cosine_score"] + [c for c in chunks_df.columns if c != "chunk_id"]
    return result[cols].sort_values("cosine_score", ascending=False)


def rank_by_bm25(query: str, chunks_df: pd.DataFrame, text_column: str = "chunk_text") -> pd.DataFrame:
    corpus = chunks_df[text_column].fillna("").tolist()
    scores = np.maximum(BM25Okapi([t.lower().split() for t in corpus], epsilon=0.25).get_scores(query.lower().split()), 0.0)
    out = chunks_df.copy()
    out["bm25_score"] = scores
    return out.sort_values("bm25_score", ascending=False).reset_index(drop=True)


def rank_by_rrf(cosine_df: pd.DataFrame, bm25_df: pd.DataFrame, k: int = 60, id_column: str = "chunk_id") -> pd.DataFrame:
    if id_column not in cosine_df.columns or id_column not in bm25_df.columns:
        raise ValueError(f"Column '{id_column}' must exist in both DataFrames")
    cr = cosine_df[[id_column]].copy()
    cr["rank_cosine"] = np.arange(1, len(cr) + 1)
    br = bm25_df[[id_column]].copy()
    br["rank_bm25"] = np.arange(1, len(br) + 1)
    ranks = pd.merge(cr, br, on=id_column, how="outer")
    worst = max(len(cr), len(br)) + k
    ranks["rank_cosine"] = ranks["rank_cosine"].fillna(worst)
    ranks["rank_bm25"] = ranks["rank_bm25"].fillna(worst)
    ranks["rrf_score"] = 1.0 / (k + ranks["rank_cosine"]) + 1.0 / (k + ranks["rank_bm25"])
    out = pd.merge(bm25_df, cosine_df[[id_column, "cosine_score"]], on=id_column, how="outer")
    out = pd.merge(out, ranks[[id_column, "rrf_score"]], on=id_column, how="left")
    return out.sort_values("rrf_score", ascending=False).reset_index(drop=True)
