"""Drop-in for the reference's ``Method/Semantic_Grouping_Optimized.py``.

``semantic_grouping_main`` / ``semantic_chunk_passage_from_grouping_logic`` keep the reference's
signatures, sentinel ids (``_single``, ``_matrix_fail``, ``_fallback``, ``_cluster{i}``) and
metadata keys.  The dense part runs on the GPU:

    embeddings --K3--> S --K4--> sim_sharp, centrality, q80/q65/q60/0.1*std, kNN lists

and the sequential, data-dependent clustering that follows (spectral embedding + seeded k-means,
split / merge / refine / one-pass reassignment; reference :133-588) is re-implemented here on the
host with the same semantics — including its quirks (empty clusters keep their index, a merged
small cluster may appear twice).  It consumes the device outputs in the reference's dtypes:
``sim_matrix``/``sim_sharp`` float32, ``centrality``/``W_all``/thresholds float64.

``group_documents`` batches many documents into one K3 + K4 launch.
"""
from __future__ import annotations

import json
import math
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..Tool import Sentence_Segmenter as _segmenter
from .semantic_common import embed_sentences_batched, log_msg, normalize_device

_HEADER_RE = re.compile(r"\s*[\"“”']{0,3}\s*Language:\s*\w+\s+Article\s*Type:\s*[A-Za-z0-9\-]+\.?\s*", re.IGNORECASE)


def extract_sentences_spacy(text: str) -> List[str]:
    """Module-level hook (patched like the reference's import at :7)."""
    return _segmenter.extract_sentences_spacy(text)


def _preclean(text) -> str:
    """Reference :65-78 — drop 'Language: .. Article Type: ..' residue, collapse whitespace."""
    if not isinstance(text, str):
        return ""
    return re.sub(r"\s+", " ", _HEADER_RE.sub(" ", text)).strip()


# ----------------------------------------------------------------------------------------------
# Device pass
# ----------------------------------------------------------------------------------------------
@dataclass
class DevicePass:
    """What the GPU hands to the host stage for one document (reference dtypes)."""
    sim_matrix: np.ndarray   # n x n float32
    sim_sharp: np.ndarray    # n x n float32, zero diagonal
    centrality: np.ndarray   # n float64
    mu: float
    sigma: float
    q80: float
    q65: float
    q60: float
    reassign_delta: float    # 0.1 * std(positive entries)
    n_positive: int
    k_all: int
    knn_idx: np.ndarray      # n x 33 int32
    knn_val: np.ndarray      # n x 33 float32


def device_pass_batch(doc_embeddings: Sequence[np.ndarray], tau: float = 0.15, knn_mode: int = 0) -> List[Optional[DevicePass]]:
    """K3 + K4 for a batch of documents in two kernel launches (``None`` for n < 2)."""
    import torch
    from .. import ragged
    sizes = [int(e.shape[0]) if e is not None and getattr(e, "ndim", 0) == 2 else 0 for e in doc_embeddings]
    live = [d for d, n in enumerate(sizes) if n >= 2]
    out: List[Optional[DevicePass]] = [None] * len(sizes)
    if not live:
        return out
    rows = [np.ascontiguousarray(doc_embeddings[d], dtype=np.float32) for d in live]
    plan = ragged.make_plan([r.shape[0] for r in rows], "cuda")
    E = torch.from_numpy(np.concatenate(rows, axis=0)).cuda()
    S = ragged.segmented_simmatrix(E, plan)
    res = ragged.group_threshold_pass(S, plan, tau=tau, knn_mode=knn_mode)
    S_h = S.cpu().numpy()
    sharp_h = res["sim_sharp"].cpu().numpy()
    cent_h = res["centrality"].cpu().numpy()
    stats_h = res["doc_stats"].cpu().numpy()
    kidx_h = res["knn_idx"].cpu().numpy()
    kval_h = res["knn_val"].cpu().numpy()
    for slot, d in enumerate(live):
        n = sizes[d]
        a, b = plan.s_offsets[slot], plan.s_offsets[slot + 1]
        r0, r1 = plan.offsets[slot], plan.offsets[slot + 1]
        st = stats_h[slot]
        out[d] = DevicePass(
            sim_matrix=S_h[a:b].reshape(n, n).copy(), sim_sharp=sharp_h[a:b].reshape(n, n).copy(),
            centrality=cent_h[r0:r1].copy(), mu=float(st[0]), sigma=float(st[1]), q80=float(st[2]), q65=float(st[3]),
            q60=float(st[4]), reassign_delta=float(st[5]), n_positive=int(st[6]), k_all=int(st[7]),
            knn_idx=kidx_h[r0:r1].copy(), knn_val=kval_h[r0:r1].copy())
    return out


# ----------------------------------------------------------------------------------------------
# Host stage: block means, spectral clustering, k-means
# ----------------------------------------------------------------------------------------------
def _mean_between(sharp: np.ndarray, A: List[int], B: List[int]) -> float:
    """Reference :118-121 — fp64 mean of sim_sharp over A x B (row-major), 0.0 if either is empty."""
    if not A or not B:
        return 0.0
    return float(np.mean(sharp[np.ix_(A, B)].astype(np.float64).ravel()))


def _mean_within(sharp: np.ndarray, A: List[int]) -> float:
    """Reference :123-130 — fp64 mean over unordered pairs i < j, 1.0 for |A| <= 1."""
    m = len(A)
    if m <= 1:
        return 1.0
    sub = sharp[np.ix_(A, A)].astype(np.float64)
    vals = sub[np.triu_indices(m, 1)]
    return float(np.mean(vals)) if vals.size else 1.0


def _normalized_laplacian(W: np.ndarray) -> np.ndarray:
    """Reference :285-292 — L = I - D^-1/2 W D^-1/2 with isolated vertices mapped to 0."""
    d = np.sum(W, axis=1)
    with np.errstate(divide="ignore"):
        d_inv_sqrt = np.where(d > 0, 1.0 / np.sqrt(d), 0.0)
    D = np.diag(d_inv_sqrt)
    return np.eye(W.shape[0], dtype=float) - (D @ W @ D)


def _kmeans(X: np.ndarray, k: int, n_init: int = 5, max_iter: int = 100, seed: int = 0) -> np.ndarray:
    """Reference :294-316 — Lloyd iterations from ``RandomState(seed).choice`` starts; the restart
    with the strictly smallest inertia wins."""
    rng = np.random.RandomState(seed)
    best_labels, best_inertia = None, float("inf")
    for _ in range(n_init):
        centers = X[rng.choice(X.shape[0], size=k, replace=False)].copy()
        labels = None
        for _ in range(max_iter):
            dists = ((X[:, None, :] - centers[None, :, :]) ** 2).sum(axis=2)
            labels = np.argmin(dists, axis=1)
            new_centers = np.vstack([X[labels == c].mean(axis=0) if np.any(labels == c) else centers[c] for c in range(k)])
            shift = float(np.linalg.norm(new_centers - centers))
            centers = new_centers
            if shift < 1e-6:
                break
        inertia = float(((X - centers[labels]) ** 2).sum())
        if inertia < best_inertia:
            best_inertia, best_labels = inertia, labels.copy()
    return best_labels.astype(int)


def _row_normalize(U: np.ndarray) -> np.ndarray:
    return U / (np.linalg.norm(U, axis=1) + 1e-9)[:, None]


def _auto_k_spectral_labels(W: np.ndarray, kmax: int) -> Optional[np.ndarray]:
    """Reference :318-341 — eigengap choice of K on the normalised Laplacian, then k-means."""
    nn = W.shape[0]
    if nn <= 2 or np.allclose(W, 0.0):
        return None
    try:
        evals, evecs = np.linalg.eigh(_normalized_laplacian(W))
    except Exception:
        return None
    order = np.argsort(evals)
    evals, evecs = evals[order], evecs[:, order]
    kmax_eff = int(max(2, min(kmax, nn - 1)))
    gaps = np.diff(evals[: kmax_eff + 1])
    k = 2 if gaps.size == 0 else max(2, min(int(np.argmax(gaps) + 1), kmax_eff))
    return _kmeans(_row_normalize(evecs[:, :k]), k=k, n_init=5, max_iter=100, seed=0)


def _rmt_filter(S: np.ndarray, keep_eigs: int = 3) -> np.ndarray:
    """Reference :133-165 — keep the top eigen-components, flatten the rest to their mean."""
    try:
        evals, evecs = np.linalg.eigh(0.5 * (S + S.T))
        order = np.argsort(evals)[::-1]
        evals, evecs = evals[order], evecs[:, order]
        k = int(max(1, min(keep_eigs, S.shape[0])))
        if k < len(evals):
            noise = float(np.mean(evals[k:]))
            evals_f = np.array([evals[i] if i < k else noise for i in range(len(evals))], dtype=float)
        else:
            evals_f = evals.astype(float)
        S_f = np.maximum((evecs @ np.diag(evals_f) @ evecs.T).astype(float), 0.0)
    except Exception:
        S_f = np.maximum(S.astype(float), 0.0)
    np.fill_diagonal(S_f, 0.0)
    return S_f


def _modularity_multiscale_labels(S_filtered, gamma_start, gamma_end, gamma_step, edge_floor_local, kmax_cap):
    """Reference :168-268 — Louvain sweep over the resolution, co-association consensus, spectral
    k-means.  Needs networkx + python-louvain; without them it returns None and the caller falls
    back to the spectral engine, exactly like the reference (:265-267)."""
    n_local = int(S_filtered.shape[0])
    if n_local <= 2:
        return None
    A = np.where(S_filtered >= float(edge_floor_local), S_filtered, 0.0).astype(float)
    np.fill_diagonal(A, 0.0)
    if np.allclose(A, 0.0):
        return None
    try:
        import networkx as nx  # type: ignore
        import community as community_louvain  # type: ignore
    except Exception:
        return None
    G = nx.Graph()
    G.add_nodes_from(range(n_local))
    iu, ju = np.nonzero(np.triu(A, 1) > 0.0)
    G.add_weighted_edges_from((int(i), int(j), float(A[i, j])) for i, j in zip(iu, ju))
    if G.number_of_edges() == 0:
        return None
    step = float(gamma_step if gamma_step > 0 else 0.2)
    cur = float(gamma_start)
    label_list = []
    while cur <= float(gamma_end) + 1e-9:
        try:
            part = community_louvain.best_partition(G, weight="weight", resolution=float(cur), random_state=0)
            lab = np.array([int(part.get(i, 0)) for i in range(n_local)], dtype=int)
            k = int(np.max(lab) + 1) if lab.size else 0
            if 2 <= k <= int(max(2, min(kmax_cap, n_local - 1))):
                label_list.append(lab)
        except Exception:
            pass
        cur += step
    if not label_list:
        return None
    try:
        C = np.zeros((n_local, n_local), dtype=float)
        for lab in label_list:
            C += (lab[:, None] == lab[None, :]).astype(float)
        np.fill_diagonal(C, 0.0)
        C = C / float(len(label_list))
        thr = float(np.quantile(C[np.triu_indices(n_local, 1)], 0.5)) if n_local > 1 else 0.0
        Wc = np.where(C >= thr, C, 0.0)
        Wc = np.maximum(Wc, Wc.T)
        if np.allclose(Wc, 0.0):
            return label_list[-1]
        evals, evecs = np.linalg.eigh(_normalized_laplacian(Wc))
        order = np.argsort(evals)
        evals = evals[order]
        gaps = np.diff(evals[: min(len(evals) - 1, kmax_cap) + 1])
        k_final = 2 if gaps.size == 0 else int(max(2, min(kmax_cap, int(np.argmax(gaps) + 1))))
        return _kmeans(_row_normalize(evecs[:, :k_final]), k=k_final, n_init=10, max_iter=200, seed=0)
    except Exception:
        return label_list[-1]


# ----------------------------------------------------------------------------------------------
# Host stage: the sequential pipeline
# ----------------------------------------------------------------------------------------------
def cluster_from_device_pass(dp: DevicePass, *, auto_params: bool = True, knn_k: Optional[int] = None, edge_floor: float = 0.25,
                             spectral_kmax: Optional[int] = None, rmt_keep_eigs: int = 3, mod_gamma_start: float = 0.7,
                             mod_gamma_end: float = 1.6, mod_gamma_step: float = 0.15, cap_soft: Optional[int] = None,
                             small_group_min: int = 2, tau_merge: float = 0.38, reassign_delta: float = 0.02,
                             engine: Optional[str] = None,
                             W_override: Optional[np.ndarray] = None) -> Tuple[List[List[int]], str, np.ndarray]:
    """Reference :343-588 on the device outputs.  Returns (clusters, method_used, W_all).
    ``W_override`` substitutes a pre-built kNN graph (tests pin the host stage with the reference's own W)."""
    from ..ragged import knn_graph_from_lists
    sharp = dp.sim_sharp
    n = sharp.shape[0]
    has_pos = dp.n_positive > 0
    eff_edge_floor = (dp.q80 if has_pos else 0.4) if auto_params else float(edge_floor)
    W_all = W_override if W_override is not None else knn_graph_from_lists(dp.knn_idx, dp.knn_val, eff_edge_floor)

    if auto_params:
        kmax_eff = int(max(2, min(16, max(2, n // 6))))
    else:
        kmax_eff = int(spectral_kmax if spectral_kmax is not None else max(2, min(10, max(2, n // 5))))
    labels = None
    eng = (engine or "rmt").lower().strip()
    method_used = "RMT"
    if eng == "spectral":
        method_used = "SpectralOnly"
        labels = _auto_k_spectral_labels(W_all, kmax=kmax_eff)
    else:
        try:
            labels = _modularity_multiscale_labels(_rmt_filter(sharp, int(max(1, rmt_keep_eigs))), float(mod_gamma_start),
                                                   float(mod_gamma_end), float(mod_gamma_step), eff_edge_floor, kmax_eff)
        except Exception:
            labels = None
        if labels is None:
            method_used = "SpectralFallback"
            labels = _auto_k_spectral_labels(W_all, kmax=kmax_eff)

    if labels is None:
        groups: List[List[int]] = [list(range(n))]
    else:
        groups = [[] for _ in range(int(np.max(labels) + 1))]
        for i, lab in enumerate(labels.tolist()):
            groups[int(lab)].append(i)

    # ---- split over-large clusters (reference :403-442) ---------------------------------------
    if auto_params and cap_soft is None:
        eff_cap_soft = int(max(20, n // 4))
    else:
        eff_cap_soft = int(cap_soft if cap_soft is not None else max(20, n // 3))

    def bisect(members: List[int]):
        if len(members) < 4:
            return None
        try:
            _, evecs = np.linalg.eigh(_normalized_laplacian(W_all[np.ix_(members, members)]))
        except Exception:
            return None
        lab2 = _kmeans(_row_normalize(evecs[:, :2]), k=2, n_init=5, max_iter=100, seed=1)
        left = [m for m, l in zip(members, lab2) if l == 0]
        right = [m for m, l in zip(members, lab2) if l == 1]
        if not left or not right:
            return None
        sep = _mean_between(sharp, left, right) - 0.5 * (_mean_within(sharp, left) + _mean_within(sharp, right))
        return (sorted(left), sorted(right)) if sep < 0.0 else None

    split_groups: List[List[int]] = []
    for g in groups:
        halves = bisect(g) if len(g) > eff_cap_soft else None
        if halves is not None and all(len(x) >= max(2, small_group_min) for x in halves):
            split_groups.extend(list(halves))
        else:
            split_groups.append(sorted(g))
    groups = split_groups

    # ---- merge undersized clusters (reference :444-491) ---------------------------------------
    if auto_params:
        sizes = [len(g) for g in groups]
        min_len = int(max(2, np.percentile(sizes, 10))) if len(sizes) >= 5 else 2
        eff_tau_merge = dp.q65 if has_pos else float(tau_merge)
    else:
        min_len = int(max(2, small_group_min))
        eff_tau_merge = float(tau_merge)
    merged: List[List[int]] = []
    consumed = set()
    for i, g in enumerate(groups):
        if i in consumed:
            continue
        if len(g) >= max(2, int(min_len)):
            merged.append(g)
            continue
        best_j, best_gain = None, 0.0
        for j, h in enumerate(groups):
            if j == i or j in consumed:
                continue
            if _mean_between(sharp, g, h) < float(eff_tau_merge):
                continue
            gain = _mean_within(sharp, sorted(g + h)) - 0.5 * (_mean_within(sharp, g) + _mean_within(sharp, h))
            if gain > best_gain:
                best_gain, best_j = gain, j
        if best_j is not None and best_gain > 0.0:
            consumed.add(best_j)
            merged.append(sorted(groups[best_j] + g))
        else:
            merged.append(g)

    # ---- refine: split loose clusters, merge near-duplicate neighbours (reference :494-553) ------
    try:
        internal = [float(_mean_within(sharp, g)) for g in merged]
        low_thr = float(np.percentile(np.array(internal, dtype=float), 25)) if len(internal) >= 2 else 0.0
        refined: List[List[int]] = []
        for g in merged:
            if len(g) >= 6 and float(_mean_within(sharp, g)) < max(0.5, low_thr):
                halves = bisect(g)
                if halves is not None:
                    parent = float(_mean_within(sharp, g))
                    if float(_mean_within(sharp, halves[0])) > parent and float(_mean_within(sharp, halves[1])) > parent:
                        refined.append(sorted(halves[0]))
                        refined.append(sorted(halves[1]))
                        continue
            refined.append(g)
        global_merge_thr = dp.q60 if has_pos else 0.5
        merged_adj: List[List[int]] = []
        i = 0
        while i < len(refined):
            cur = refined[i]
            j = i + 1
            while j < len(refined):
                inter = _mean_between(sharp, cur, refined[j])
                cmp_thr = 0.9 * min(max(float(_mean_within(sharp, cur)), 1e-6), max(float(_mean_within(sharp, refined[j])), 1e-6))
                if inter >= max(cmp_thr, global_merge_thr):
                    cur = sorted(cur + refined[j])
                    j += 1
                else:
                    break
            merged_adj.append(cur)
            i = j
        merged = merged_adj
    except Exception:
        pass

    # ---- one-pass sentence reassignment (reference :555-588) -----------------------------------
    if len(merged) >= 2:
        delta = (dp.reassign_delta if has_pos else float(reassign_delta)) if auto_params else float(reassign_delta)
        for x in range(n):
            cur = next((cid for cid, g in enumerate(merged) if x in g), None)
            if cur is None:
                continue
            members = [y for y in merged[cur] if y != x]
            best_c = cur
            best_score = float(np.mean(sharp[x, members].astype(np.float64))) if members else 0.0
            for c2, h in enumerate(merged):
                if c2 == cur:
                    continue
                other = float(np.mean(sharp[x, h].astype(np.float64))) if h else 0.0
                if other > best_score + float(delta):
                    best_score, best_c = other, c2
            if best_c != cur:
                merged[cur] = [y for y in merged[cur] if y != x]
                merged[best_c] = sorted(merged[best_c] + [x])
    return merged, method_used, W_all


def _emit(doc_id: str, passage_text: str, sentences: List[str], merged: List[List[int]], method_used: str, dp: DevicePass,
          collect_metadata: bool) -> List[Tuple[str, str, Optional[str]]]:
    """Reference :590-654."""
    n = len(sentences)
    out: List[Tuple[str, str, Optional[str]]] = []
    cent = [float(c) for c in dp.centrality] if collect_metadata else []
    for i, g in enumerate(merged):
        idxs = [idx for idx in sorted(set(g)) if 0 <= idx < n]
        if not idxs:
            continue
        text = " ".join(sentences[idx] for idx in idxs).strip()
        if not text:
            continue
        cid = f"{doc_id}_cluster{i}"
        if not collect_metadata:
            out.append((cid, text, None))
            continue
        meta = {"chunk_id": cid, "sent_indices": ",".join(str(x) for x in sorted(set(g))), "n": len(g), "method_used": method_used}
        if cent and g:
            exemplar = max(g, key=lambda t: cent[t])
            sims = [float(dp.sim_matrix[exemplar, j]) for j in g if j != exemplar]
            if sims:
                m = sum(sims) / len(sims)
                var = sum((x - m) ** 2 for x in sims) / len(sims)
                meta.update({"exemplar": exemplar, "sim_mean": round(m, 4), "sim_min": round(min(sims), 4),
                             "sim_max": round(max(sims), 4), "sim_std": round(math.sqrt(var), 4),
                             "exemplar_centrality": round(cent[exemplar], 4)})
        out.append((cid, text, json.dumps(meta, ensure_ascii=False)))
    if not out:
        return [(f"{doc_id}_fallback", passage_text, None)]
    return out


def _knn_mode(auto_params: bool, knn_k: Optional[int]) -> int:
    if auto_params:
        return 0
    if knn_k is None:
        return -1
    if int(knn_k) > 32:
        raise ValueError("knn_k > 32 is not supported by the device threshold pass")
    return max(1, int(knn_k))


def semantic_grouping_main(
    passage_text: str,
    doc_id: str,
    embedding_model: str,
    *,
    knn_k: Optional[int] = None,
    edge_floor: float = 0.25,
    spectral_kmax: Optional[int] = None,
    rmt_keep_eigs: int = 3,
    mod_gamma_start: float = 0.7,
    mod_gamma_end: float = 1.6,
    mod_gamma_step: float = 0.15,
    cap_soft: Optional[int] = None,
    small_group_min: int = 2,
    tau_merge: float = 0.38,
    reassign_delta: float = 0.02,
    embedding_batch_size: int = 64,
    device: Optional[str] = "cuda",
    silent: bool = False,
    collect_metadata: bool = False,
    sigmoid_tau_group: Optional[float] = None,
    engine: Optional[str] = None,
    **_extra,
) -> List[Tuple[str, str, Optional[str]]]:
    """Same signature and return convention as the reference (:14-42); unknown keyword
    arguments are accepted and ignored."""
    log_msg(silent, f"[grouping] doc={doc_id} model={embedding_model}", "info", "grouping")
    passage_text = _preclean(passage_text)
    sentences = extract_sentences_spacy(passage_text)
    if not sentences:
        return []
    if len(sentences) <= 1:
        return [(f"{doc_id}_single", passage_text, None)]
    normalize_device(device or "cuda")
    embs = embed_sentences_batched(sentences, embedding_model, base_batch_size=embedding_batch_size, device="cuda", silent=silent)
    if embs is None or embs.size == 0 or embs.shape[0] != len(sentences):
        return [(f"{doc_id}_matrix_fail", passage_text, None)]
    auto_params = bool(_extra.get("auto_params", True))
    tau = 0.15 if sigmoid_tau_group is None else float(sigmoid_tau_group)
    dp = device_pass_batch([np.asarray(embs, dtype=np.float32)], tau=tau, knn_mode=_knn_mode(auto_params, knn_k))[0]
    merged, method_used, _ = cluster_from_device_pass(
        dp, auto_params=auto_params, knn_k=knn_k, edge_floor=edge_floor, spectral_kmax=spectral_kmax,
        rmt_keep_eigs=rmt_keep_eigs, mod_gamma_start=mod_gamma_start, mod_gamma_end=mod_gamma_end, mod_gamma_step=mod_gamma_step,
        cap_soft=cap_soft, small_group_min=small_group_min, tau_merge=tau_merge, reassign_delta=reassign_delta, engine=engine)
    chunks = _emit(doc_id, passage_text, sentences, merged, method_used, dp, collect_metadata)
    log_msg(silent, f"[grouping] doc={doc_id} method={method_used} clusters={len(chunks)}", "info", "grouping")
    return chunks


def group_documents(docs: Sequence[Tuple[str, List[str], np.ndarray]], *, collect_metadata: bool = False,
                    sigmoid_tau_group: Optional[float] = None, **params) -> Dict[str, List[Tuple[str, str, Optional[str]]]]:
    """Batched form: ``docs`` = (doc_id, sentences, embeddings) triples; one K3 + K4 launch for the
    whole batch, then the host stage per document."""
    auto_params = bool(params.pop("auto_params", True))
    tau = 0.15 if sigmoid_tau_group is None else float(sigmoid_tau_group)
    passes = device_pass_batch([np.asarray(e, dtype=np.float32) for _, _, e in docs], tau=tau,
                               knn_mode=_knn_mode(auto_params, params.get("knn_k")))
    out = {}
    for (doc_id, sentences, _), dp in zip(docs, passes):
        text = " ".join(sentences)
        if dp is None:
            out[doc_id] = [(f"{doc_id}_single", text, None)] if sentences else []
            continue
        merged, method_used, _ = cluster_from_device_pass(dp, auto_params=auto_params, **params)
        out[doc_id] = _emit(doc_id, text, sentences, merged, method_used, dp, collect_metadata)
    return out


def semantic_chunk_passage_from_grouping_logic(
    doc_id: str,
    passage_text: str,
    embedding_model: str = "thenlper/gte-base",
    *,
    knn_k: Optional[int] = None,
    edge_floor: float = 0.25,
    spectral_kmax: Optional[int] = None,
    rmt_keep_eigs: int = 3,
    mod_gamma_start: float = 0.7,
    mod_gamma_end: float = 1.6,
    mod_gamma_step: float = 0.15,
    cap_soft: Optional[int] = None,
    small_group_min: int = 2,
    tau_merge: float = 0.38,
    reassign_delta: float = 0.02,
    embedding_batch_size: int = 64,
    device: Optional[str] = "cuda",
    silent: bool = False,
    collect_metadata: bool = False,
    sigmoid_tau_group: Optional[float] = None,
    engine: Optional[str] = None,
    **_extra,
) -> List[Tuple[str, str, Optional[str]]]:
    """Controller-facing wrapper (reference :657-705): like the reference it does NOT forward
    ``**_extra`` (so ``auto_params`` stays at its default there too)."""
    return semantic_grouping_main(
        passage_text=passage_text, doc_id=doc_id, embedding_model=embedding_model, cap_soft=cap_soft,
        small_group_min=small_group_min, tau_merge=tau_merge, knn_k=knn_k, edge_floor=edge_floor, spectral_kmax=spectral_kmax,
        rmt_keep_eigs=rmt_keep_eigs, mod_gamma_start=mod_gamma_start, mod_gamma_end=mod_gamma_end, mod_gamma_step=mod_gamma_step,
        reassign_delta=reassign_delta, embedding_batch_size=embedding_batch_size, device=device, silent=silent,
        collect_metadata=collect_metadata, sigmoid_tau_group=sigmoid_tau_group, engine=engine)
