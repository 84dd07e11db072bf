"""Drop-in for the reference's ``Method/Semantic_Splitter_Optimized.py``.

Same public names, signatures and return conventions (``process_sentence_splitting_with_semantics``
-> ``(chunks, sentences, groups)``, ``semantic_splitter_main`` / ``chunk_passage_text_splitter``
-> ``[(f"{doc_id}_chunk{i}", text, meta_json|None)]``, ``_c99_boundaries`` and
``_valley_boundaries`` importable by name).  The dense arithmetic runs on the GPU:

* adjacent-sentence cosine + median-of-3 smoothing + median / MAD / P25 / P75   -> K5
* the C99 similarity matrix and its rank transform                              -> K3 + ss_c99_rank_matrix

* the greedy divisive cut search on the rank matrix (reference :194-238)            -> K10 (ss_c99_divisive_cuts)

while the sequential boundary logic (valley detection, the C99 profile knee, voting, NMS, soft cap,
boundary snapping, short-segment merge; reference :239-338,447-652) is re-implemented on the host
with the same semantics.  The divisive search evaluates block means through a float64 summed-area
table instead of one ``ndarray.mean()`` per candidate cut; ``c99_boundaries_batch`` runs it for many
documents at once.
"""
from __future__ import annotations

import json
import math
import re
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..Tool import Sentence_Segmenter as _segmenter
from .semantic_common import embed_sentences_batched, normalize_device, pack_document_rows


def extract_sentences_spacy(text: str) -> List[str]:
    return _segmenter.extract_sentences_spacy(text)


_PRE_SUBS = [
    (re.compile(r"^Language:\s*\w+\s+Article\s*Type:\s*[^\s\[\]]*\s*(?:\[Text\])?\s*", re.IGNORECASE), ""),
    (re.compile(r"\s*[\"“”']{0,3}\s*Language:\s*\w+\s+Article\s*Type:\s*[A-Za-z0-9\-]+\.?", re.IGNORECASE), " "),
    (re.compile(r"\[Article by[^\]]*\]\s*"), ""),
    (re.compile(r"\[Report by[^\]]*\]\s*"), ""),
    (re.compile(r"\[From the[^\]]*\]\s*"), ""),
    (re.compile(r"\[Excerpts?\]\s*"), ""),
    (re.compile(r"\[Text\]\s*"), ""),
]


def _preclean(text) -> str:
    """Reference :382-394."""
    if not isinstance(text, str):
        return ""
    s = text
    for rx, rep in _PRE_SUBS:
        s = rx.sub(rep, s)
    return re.sub(r"\s+", " ", s).strip()


# ----------------------------------------------------------------------------------------------
# Device passes
# ----------------------------------------------------------------------------------------------
def _embed(sentences: List[str], model_name: str, device, silent: bool) -> Optional[np.ndarray]:
    """Reference :140-152.  Returns the RAW sentence embeddings: every consumer below normalises
    on the fly inside its kernel, which is what the reference's explicit division achieves."""
    if not sentences:
        return None
    embs = embed_sentences_batched(sentences, model_name, base_batch_size=32, device=device, silent=silent)
    if embs is None or embs.size == 0:
        return None
    return np.asarray(embs, dtype=np.float32)


def splitter_device_pass(doc_embeddings: Sequence[np.ndarray], pct: float = 95.0):
    """K5 for a batch of documents: returns per document ``dict(adj_sims, adj_base, median, mad,
    p25, p75, p95_threshold, breakpoints)`` (``None`` for documents with fewer than 2 sentences)."""
    import torch
    from .. import ragged
    sizes = [int(e.shape[0]) if e is not None and getattr(e, "ndim", 0) == 2 else 0 for e in doc_embeddings]
    live = [d for d, n in enumerate(sizes) if n >= 2]
    out = [None] * len(sizes)
    if not live:
        return out
    plan = ragged.make_plan([sizes[d] for d in live], "cuda")
    E = pack_document_rows([doc_embeddings[d] for d in live])   # host arrays or CUDA tensors straight from the encoder
    adj = ragged.adjacent_cosine(E)
    thr, flags, stats, smooth = ragged.segmented_percentile(adj, plan, pct, want_stats=True)
    adj_h, thr_h, flags_h = adj.cpu().numpy(), thr.cpu().numpy(), flags.cpu().numpy()
    stats_h, smooth_h = stats.cpu().numpy(), smooth.cpu().numpy()
    for slot, d in enumerate(live):
        a, b = plan.offsets[slot], plan.offsets[slot + 1]
        m = b - a - 1
        out[d] = {
            "adj_sims": [float(x) for x in adj_h[a:a + m]],            # fp32 dots widened, like float(E[i] @ E[i+1])
            "adj_base": [float(x) for x in smooth_h[a:a + m]],
            "median": float(stats_h[slot, 0]), "mad": float(stats_h[slot, 1]),
            "p25": float(stats_h[slot, 2]), "p75": float(stats_h[slot, 3]),
            "p95_threshold": float(thr_h[slot]), "breakpoints": np.nonzero(flags_h[a:b])[0],
        }
    return out


_C99_BATCH_BYTES = 6 << 30  # device bytes one slice of a C99 batch may take (S + R + float64 block-sum table)


def c99_boundaries_batch(doc_embeddings: Sequence[np.ndarray], min_chunk_sizes, max_cuts: Optional[int] = None,
                         min_gain: float = 0.01, *, use_local_rank: bool = False, mask_size: int = 11, stopping: str = "gain",
                         knee_c: float = 1.2, smooth_window: int = 3) -> List[List[int]]:
    """``_c99_boundaries`` (reference :155-264) for a batch of documents in three launches per slice: similarity
    matrices (K3), rank transform, divisive cut search (K10).  Only the picked cuts (and, for
    ``stopping='profile'``, the density profile) come back to the host.  ``min_chunk_sizes``: one int or one per
    document.  Documents of up to 4096 sentences are supported (the reference corpus's longest has 3939)."""
    import torch
    from .. import ragged
    n_docs = len(doc_embeddings)
    mins = [int(min_chunk_sizes)] * n_docs if np.ndim(min_chunk_sizes) == 0 else [int(x) for x in min_chunk_sizes]
    if len(mins) != n_docs:
        raise ValueError("min_chunk_sizes must be a scalar or one value per document")
    mode = str(stopping).lower()
    out: List[List[int]] = [[] for _ in range(n_docs)]
    if max_cuts is not None and int(max_cuts) <= 0:
        return out                                       # :225 stops before the first cut
    # min_chunk < 1 searches like min_chunk = 1: the extra candidate cuts of the reference (c = a and c = b, :212) have an
    # empty side, whose block mean is NaN and never wins
    mins = [max(1, m) for m in mins]
    live = []
    for d, e in enumerate(doc_embeddings):
        n = int(e.shape[0])
        if n < 2 * mins[d]:
            continue                                     # :165-166
        if n > ragged.C99_CUTS_MAX_ROWS:
            raise ValueError(f"document {d} has {n} sentences; the device cut search takes at most {ragged.C99_CUTS_MAX_ROWS} "
                             "(there is no host fallback)")
        live.append(d)
    start = 0
    while start < len(live):
        stop, used = start, 0
        while stop < len(live):
            n = int(doc_embeddings[live[stop]].shape[0])
            need = 16 * (n + 1) * (n + 1)
            if stop > start and used + need > _C99_BATCH_BYTES:
                break
            used += need
            stop += 1
        ids = live[start:stop]
        start = stop
        plan = ragged.make_plan([int(doc_embeddings[d].shape[0]) for d in ids], "cuda")
        E = pack_document_rows([doc_embeddings[d] for d in ids])
        S = ragged.segmented_simmatrix(E, plan, validate=True)
        R = ragged.c99_rank_matrix(S, plan, use_local_rank=bool(use_local_rank), mask_size=int(mask_size), symmetric=True)
        del S
        cuts, n_cuts, profile = ragged.c99_divisive_cuts(R, plan, [mins[d] for d in ids], max_cuts, float(min_gain),
                                                         stop_by_gain=(mode == "gain"), want_profile=(mode == "profile"))
        cuts_h, n_h = cuts.cpu().numpy(), n_cuts.cpu().numpy()
        prof_h = profile.cpu().numpy() if profile is not None else None
        for slot, d in enumerate(ids):
            base, cnt = int(plan.offsets[slot]), int(n_h[slot])
            if cnt < 0:
                raise RuntimeError("ss_c99_divisive_cuts skipped a document it was dispatched for")
            picked = [int(x) for x in cuts_h[base:base + cnt]]
            if mode != "profile" or not picked:
                out[d] = sorted(set(picked))
            else:
                out[d] = _profile_knee(picked, prof_h[base:base + cnt + 1], float(knee_c), int(smooth_window))
    return out


# ----------------------------------------------------------------------------------------------
# Host logic
# ----------------------------------------------------------------------------------------------
def _profile_knee(cuts: List[int], profile, knee_c: float, smooth_window: int) -> List[int]:
    """Reference :239-264 — keep the cuts picked before the first sharp drop of the smoothed density increments."""
    deltas = np.diff(np.array(profile, dtype=float))
    if deltas.size == 0:
        return sorted(set(cuts))
    sw = max(1, int(smooth_window))
    dsm = np.convolve(deltas, np.ones(sw, dtype=float) / float(sw), mode="same") if sw > 1 and deltas.size >= sw else deltas
    limit = float(dsm.mean()) - float(knee_c) * float(dsm.std() + 1e-9)
    knee = next((i for i, v in enumerate(dsm, start=1) if v < limit), None)
    if knee is None:
        return sorted(set(cuts))
    return sorted(set(cuts[: min(max(1, int(knee)) - 1, len(cuts))]))


def _c99_boundaries(embs: np.ndarray, min_chunk_size: int = 3, max_cuts: Optional[int] = None, min_gain: float = 0.01, *,
                    use_local_rank: bool = False, mask_size: int = 11, stopping: str = "gain", knee_c: float = 1.2,
                    smooth_window: int = 3) -> List[int]:
    """Reference :155-264.  ``embs`` are sentence embeddings (normalised or not: the kernel
    normalises); the similarity matrix and its rank transform are computed on the GPU."""
    if embs.shape[0] < 2 * int(min_chunk_size):
        return []
    return c99_boundaries_batch([embs], int(min_chunk_size), max_cuts, min_gain, use_local_rank=use_local_rank, mask_size=mask_size,
                                stopping=stopping, knee_c=knee_c, smooth_window=smooth_window)[0]


def _valley_boundaries(adj_sims: List[float], *, triplet_tau: float = 0.12, min_boundary_spacing: int = 2,
                       min_first_boundary_index: int = 5) -> List[int]:
    """Reference :267-338 — local valleys of the adjacent-similarity series, strength z-scored and
    squashed, first-boundary constraint, score-ordered non-maximum suppression."""
    n = len(adj_sims)
    if n < 3:
        return []
    sims = np.array(adj_sims, dtype=float)
    valleys: List[Tuple[int, float]] = []
    i = 1
    while i <= n - 2:
        if not (sims[i] <= sims[i - 1]):
            i += 1
            continue
        j = lo = i
        lo_val = sims[i]
        while j + 1 <= n - 2 and sims[j + 1] <= sims[j]:
            j += 1
            if sims[j] < lo_val:
                lo_val, lo = sims[j], j
        if j < n - 1 and sims[j + 1] >= sims[j]:
            left = max(0.0, float(sims[lo - 1] - sims[lo])) if lo > 0 else 0.0
            right = max(0.0, float(sims[lo + 1] - sims[lo])) if (lo + 1) < n else 0.0
            valleys.append((lo + 1, left + right))
        i = j + 1
    if not valleys:
        return []
    strengths = np.array([s for _, s in valleys], dtype=float)
    z = (strengths - float(strengths.mean())) / float(strengths.std() + 1e-9)
    scores = 1.0 / (1.0 + np.exp(-(z / max(float(triplet_tau), 1e-9))))
    cands = [(int(b), float(sc), float(s)) for (b, s), sc in zip(valleys, scores) if b >= int(min_first_boundary_index)]
    if not cands:
        return []
    cands.sort(key=lambda x: (-x[1], -x[2]))
    spacing = max(1, int(min_boundary_spacing))
    picked: List[int] = []
    for b, _sc, _s in cands:
        if all(abs(b - x) >= spacing for x in picked):
            picked.append(b)
    return sorted(set(picked))


def _median_smooth(arr: List[float], window: int = 3) -> List[float]:
    """Reference :340-356 (host form; the drop-in takes the device result for window 3)."""
    w = int(window)
    if w <= 1:
        return list(arr)
    if w % 2 == 0:
        w += 1
    n = len(arr)
    if n == 0 or w > max(1, n):
        return list(arr)
    half = w // 2
    padded = [arr[0]] * half + list(arr) + [arr[-1]] * half
    return [float(np.median(padded[i:i + w])) for i in range(n)]


def _score_based_nms(boundaries: List[int], score_of: dict, min_spacing: int) -> List[int]:
    """Reference :358-369."""
    if not boundaries:
        return []
    spacing = max(1, int(min_spacing))
    picked: List[int] = []
    for b in sorted(boundaries, key=lambda b: (-float(score_of.get(b, 0.0)), int(b))):
        if all(abs(b - x) >= spacing for x in picked):
            picked.append(b)
    return sorted(set(picked))


def _groups_from(boundaries: List[int], sentences: List[str]):
    chunks, groups, cursor = [], [], 0
    for b in list(boundaries) + [len(sentences)]:
        grp = list(range(cursor, b))
        if grp:
            chunks.append(" ".join(sentences[cursor:b]))
            groups.append(grp)
        cursor = b
    return chunks, groups


def split_from_device_pass(sentences: List[str], embeddings: np.ndarray, dp: Dict, *, min_boundary_spacing: int = 5,
                           min_first_boundary_index: int = 5, **kw) -> Tuple[List[str], List[str], List[List[int]]]:
    """Reference :409-661 on the device outputs of one document."""
    n = len(sentences)
    auto = bool(kw.get("auto_params", True))
    adj_sims = dp["adj_sims"]
    try:
        smooth_w = int(kw.get("smooth_adj_window", 3) or 3)
    except Exception:
        smooth_w = 3
    if smooth_w == 3:
        adj_base = dp["adj_base"]
        med, mad, iqr = dp["median"], dp["mad"], dp["p75"] - dp["p25"]
    else:  # non-default window: order statistics of a handful of numbers on the host
        adj_base = _median_smooth(adj_sims, smooth_w) if smooth_w > 1 else adj_sims
        x = np.array(adj_base, dtype=float)
        med = float(np.median(x))
        mad = float(np.median(np.abs(x - med)) + 1e-9)
        iqr = float(np.percentile(x, 75) - np.percentile(x, 25))
    arr = np.array(adj_base, dtype=float)
    adj_for_valley = adj_base
    if auto:
        z = (arr - med) / mad
        adj_for_valley = (1.0 / (1.0 + np.exp(-(z / max(iqr / 2.0, 0.05))))).tolist()
    elif kw.get("sim_sigmoid_tau") is not None:
        z = (arr - float(arr.mean())) / float(arr.std() + 1e-9)
        adj_for_valley = (1.0 / (1.0 + np.exp(-(z / max(float(kw["sim_sigmoid_tau"]), 1e-9))))).tolist()
    if auto:
        min_boundary_spacing = max(5, int(round(n / 50)))
        min_first_boundary_index = max(min_first_boundary_index, int(round(0.05 * n)))
    c99_bounds = _c99_boundaries(
        embeddings, min_chunk_size=max(3, int(min_boundary_spacing)), max_cuts=None,
        use_local_rank=bool(kw.get("c99_use_local_rank", False)), mask_size=int(kw.get("c99_mask_size", 11) or 11),
        stopping=str(kw.get("c99_stopping", "gain")), knee_c=float(kw.get("c99_knee_c", 1.2) or 1.2),
        smooth_window=int(kw.get("c99_smooth_window", 3) or 3))
    valley_tau = max(iqr / 2.0, 0.06) if auto else float(kw.get("valley_tau", 0.12))
    valley_bounds = _valley_boundaries(adj_for_valley, triplet_tau=valley_tau, min_boundary_spacing=min_boundary_spacing,
                                       min_first_boundary_index=min_first_boundary_index)
    mode = "union_weighted" if auto else str(kw.get("hybrid_mode", "intersection")).lower()
    vote_thr = 0.75 if auto else float(kw.get("vote_thr", 0.8) or 0.8)
    if mode == "union_weighted":
        every = sorted(set(c99_bounds) | set(valley_bounds))
        score_map = {b: (0.5 if b in valley_bounds else 0.0) + (0.5 if b in c99_bounds else 0.0) for b in every}
        boundaries = [b for b in every if score_map[b] >= vote_thr]
    elif mode == "union":
        boundaries = sorted(set(c99_bounds) | set(valley_bounds))
        score_map = {b: (1.0 if (b in c99_bounds and b in valley_bounds) else 0.8 if b in valley_bounds else 0.7) for b in boundaries}
    else:
        try:
            tol = int(kw.get("intersect_snap_tolerance", max(1, int(min_boundary_spacing) - 1)))
        except Exception:
            tol = max(1, int(min_boundary_spacing) - 1)
        vset = sorted(set(valley_bounds))
        boundaries = sorted({c for c in set(c99_bounds) if any(abs(v - c) <= tol for v in vset)})
        score_map = {b: 1.0 for b in boundaries}
    boundaries = _score_based_nms(boundaries, score_map, min_boundary_spacing)
    if mode == "intersection" and not boundaries:
        boundaries = c99_bounds
    chunks, groups = _groups_from(boundaries, sentences)

    # ---- soft cap (reference :538-595) ----------------------------------------------------------
    try:
        cap = kw.get("soft_cap", None)
        cap = int(cap) if cap is not None else None
    except Exception:
        cap = None
    if auto and cap is None:
        cap = max(24, int(round(n * 0.12)))
    if cap and int(cap) > 0:
        cap = int(cap)
        try:
            delta = int(kw.get("soft_cap_delta", 2) or 2)
        except Exception:
            delta = 2
        new_bs: List[int] = []
        prev = 0
        for cut in sorted(boundaries) + [n]:
            while (cut - prev) > cap and (cut - prev) >= 3:
                target = prev + cap
                lo, hi = max(prev + 1, target - delta), min(cut - 1, target + delta)
                if hi <= lo:
                    break
                local = np.array(adj_sims[max(prev, lo - 1):min(cut - 1, hi)], dtype=float)
                if local.size <= 0:
                    break
                pos = max(prev + 1, lo + int(np.argmin(local)))
                if prev == 0 and pos < int(min_first_boundary_index):
                    pos = int(min_first_boundary_index)
                pos = min(max(pos, prev + 1), cut - 1)
                new_bs.append(pos)
                prev = pos
            if cut != n:
                new_bs.append(cut)
            prev = cut
        if new_bs:
            boundaries = sorted(set(x for x in new_bs if 1 <= x < n))
            chunks, groups = _groups_from(boundaries, sentences)

    # ---- snap each boundary to the local minimum of the smoothed series (reference :597-628) -----
    if auto and len(boundaries) > 0:
        win = 2
        snapped = []
        for b in sorted(boundaries):
            lo, hi = max(1, b - win), min(n - 1, b + win)
            local = arr[lo - 1:hi] if hi > lo else np.zeros(0)
            if hi <= lo or local.size == 0:
                snapped.append(b)
                continue
            snapped.append(max(1, min(n - 1, int(lo - 1 + np.argmin(local) + 1))))
        boundaries = sorted(set(snapped))
        chunks, groups = _groups_from(boundaries, sentences)

    # ---- merge short segments (reference :630-652) ------------------------------------------------
    if auto and groups:
        lens = [len(g) for g in groups]
        min_len = max(3, int(round(np.percentile(lens, 10)))) if len(lens) >= 5 else 3
        out_c: List[str] = []
        out_g: List[List[int]] = []
        buf_t, buf_g = None, []
        for ct, gp in zip(chunks, groups):
            if buf_t is None:
                buf_t, buf_g = ct, gp
            elif len(buf_g) < min_len:
                buf_t = (buf_t + " " + ct).strip()
                buf_g = list(range(buf_g[0], gp[-1] + 1))
            else:
                out_c.append(buf_t)
                out_g.append(buf_g)
                buf_t, buf_g = ct, gp
        if buf_t is not None:
            out_c.append(buf_t)
            out_g.append(buf_g)
        chunks, groups = out_c, out_g
    return chunks, sentences, groups


def process_sentence_splitting_with_semantics(
    text: str,
    *,
    embedding_model: str = "sentence-transformers/all-MiniLM-L6-v2",
    device: Optional[str] = None,
    min_boundary_spacing: int = 5,
    min_first_boundary_index: int = 5,
    silent: bool = True,
    **_legacy_kwargs,
) -> Tuple[List[str], List[str], List[List[int]]]:
    """Same contract as the reference (:371-661)."""
    text = _preclean(text)
    sentences = extract_sentences_spacy(text)
    if not sentences:
        return [], [], []
    if len(sentences) <= 1:
        return [" ".join(sentences)], sentences, [list(range(len(sentences)))]
    normalize_device(device)
    embeddings = _embed(sentences, embedding_model, "cuda", silent)
    if embeddings is None:
        return [" ".join(sentences)], sentences, [list(range(len(sentences)))]
    dp = splitter_device_pass([embeddings])[0]
    return split_from_device_pass(sentences, embeddings, dp, min_boundary_spacing=min_boundary_spacing,
                                  min_first_boundary_index=min_first_boundary_index, **_legacy_kwargs)


def semantic_splitter_main(
    doc_id: str,
    passage_text: str,
    embedding_model: str = "sentence-transformers/all-MiniLM-L6-v2",
    device: Optional[str] = None,
    min_boundary_spacing: int = 2,
    min_first_boundary_index: int = 5,
    silent: bool = False,
    collect_metadata: bool = False,
    **legacy_kwargs,
) -> List[Tuple[str, str, Optional[str]]]:
    """Same contract as the reference (:663-721)."""
    chunks, sentences, groups = process_sentence_splitting_with_semantics(
        text=passage_text, embedding_model=embedding_model, device=device, min_boundary_spacing=min_boundary_spacing,
        min_first_boundary_index=min_first_boundary_index, silent=silent, **legacy_kwargs)
    if not chunks:
        return [(f"{doc_id}_fallback", passage_text, None)] if extract_sentences_spacy(passage_text) else []
    out: List[Tuple[str, str, Optional[str]]] = []
    if not collect_metadata:
        return [(f"{doc_id}_chunk{idx}", ctext, None) for idx, ctext in enumerate(chunks)]
    # like the reference, metadata is computed on the sentences of the UNCLEANED passage (:689-692)
    sentences = extract_sentences_spacy(passage_text)
    adj = None
    if sentences and len(sentences) >= 2:
        embs = _embed(sentences, embedding_model, "cuda", True)
        dp = splitter_device_pass([embs])[0] if embs is not None else None
        adj = dp["adj_sims"] if dp is not None else None
    for idx, grp in enumerate(groups):
        ctext = " ".join(sentences[grp[0]: grp[-1] + 1]) if sentences and grp else ""
        if not ctext:
            continue
        cid = f"{doc_id}_chunk{idx}"
        meta = {"chunk_id": cid, "sent_indices": ",".join(str(i) for i in grp), "n": len(grp)}
        if adj is not None and len(grp) > 1:
            sims = [adj[a] for a in grp[:-1] if a < len(adj)]   # contiguous group: E[a].E[a+1]
            if sims:
                m = sum(sims) / len(sims)
                var = sum((x - m) ** 2 for x in sims) / len(sims)
                meta.update({"sim_mean": round(m, 4), "sim_min": round(min(sims), 4), "sim_max": round(max(sims), 4),
                             "sim_std": round(math.sqrt(var), 4)})
        out.append((cid, ctext, json.dumps(meta, ensure_ascii=False)))
    return out


def chunk_passage_text_splitter(
    doc_id: str,
    passage_text: str,
    embedding_model: str = "sentence-transformers/all-MiniLM-L6-v2",
    device: Optional[str] = None,
    min_boundary_spacing: int = 2,
    min_first_boundary_index: int = 3,
    silent: bool = False,
    collect_metadata: bool = False,
    **legacy_kwargs,
) -> List[Tuple[str, str, Optional[str]]]:
    """Controller-facing wrapper (reference :723-744)."""
    return semantic_splitter_main(
        doc_id=doc_id, passage_text=passage_text, embedding_model=embedding_model, device=device,
        min_boundary_spacing=min_boundary_spacing, min_first_boundary_index=min_first_boundary_index, silent=silent,
        collect_metadata=collect_metadata, **legacy_kwargs)
