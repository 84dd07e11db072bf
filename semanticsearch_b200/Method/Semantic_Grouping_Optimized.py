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
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..Tool import Sentence_Segmenter as _segmenter
from .semantic_common import embed_sentences_batched, log_msg, normalize_device, pack_document_rows

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
    block_sums: Optional[Callable] = None   # groups -> (rowsum [n, G], block [G, G]) float64: ragged.DocBlockSums on the device


def device_pass_batch(doc_embeddings: Sequence[np.ndarray], tau: float = 0.15, knn_mode: int = 0) -> List[Optional[DevicePass]]:
    """K3 + K4 for a batch of documents in two kernel launches (``None`` for n < 2)."""
    import torch
    from .. import ragged
    sizes = [int(e.shape[0]) if e is not None and getattr(e, "ndim", 0) == 2 else 0 for e in doc_embeddings]
    live = [d for d, n in enumerate(sizes) if n >= 2]
    out: List[Optional[DevicePass]] = [None] * len(sizes)
    if not live:
        return out
    plan = ragged.make_plan([sizes[d] for d in live], "cuda")
    E = pack_document_rows([doc_embeddings[d] for d in live])   # host arrays or CUDA tensors straight from the encoder
    S = ragged.segmented_simmatrix(E, plan, validate=True)   # the drop-ins copy results to the host anyway: one more word
    res = ragged.group_threshold_pass(S, plan, tau=tau, knn_mode=knn_mode, symmetric=True)   # K3's S is bit-symmetric
    S_h = S.cpu().numpy()
    sharp_h = res["sim_sharp"].cpu().numpy()
    cent_h = res["centrality"].cpu().numpy()
    stats_h = res["doc_stats"].cpu().numpy()
    kidx_h = res["knn_idx"].cpu().numpy()
    kval_h = res["knn_val"].cpu().numpy()
    sharp_d = res["sim_sharp"]   # stays on the device: the host stage asks for block sums over it (ss_group_block_sums)
    for slot, d in enumerate(live):
        n = sizes[d]
        a, b = plan.s_offsets[slot], plan.s_offsets[slot + 1]
        r0, r1 = plan.offsets[slot], plan.offsets[slot + 1]
        st = stats_h[slot]
        out[d] = DevicePass(
            sim_matrix=S_h[a:b].reshape(n, n).copy(), sim_sharp=sharp_h[a:b].reshape(n, n).copy(),
            centrality=cent_h[r0:r1].copy(), mu=float(st[0]), sigma=float(st[1]), q80=float(st[2]), q65=float(st[3]),
            q60=float(st[4]), reassign_delta=float(st[5]), n_positive=int(st[6]), k_all=int(st[7]),
            knn_idx=kidx_h[r0:r1].copy(), knn_val=kval_h[r0:r1].copy(),
            block_sums=ragged.DocBlockSums(sharp_d[int(a):int(b)], n))
    return out


# ----------------------------------------------------------------------------------------------
# Host stage: block means, spectral clustering, k-means
# ----------------------------------------------------------------------------------------------
class _GroupMeans:
    """Every block mean the clustering stage needs for one list of groups, from ONE launch of ss_group_block_sums
    (reference `_mean_between` / `_mean_within`, :118-130).  Groups are multisets of sentence indices; sums over unions
    of groups are sums of their blocks, so merged candidates need no further launch.  `sim_sharp` is bit-symmetric with
    a zero diagonal (K3 mirrors its tiles, K4 sharpens element-wise), hence the i < j pairs of a group are half of its
    full block."""

    def __init__(self, block_sums: Callable, groups: Sequence[Sequence[int]]):
        self.sizes = [len(g) for g in groups]
        self.rowsum, self.block = block_sums([list(g) for g in groups])

    def between(self, a: Sequence[int], b: Sequence[int]) -> float:
        """mean over (union of groups a) x (union of groups b); 0.0 if either side is empty (:119-120)."""
        la, lb = sum(self.sizes[i] for i in a), sum(self.sizes[j] for j in b)
        if not la or not lb:
            return 0.0
        return float(sum(self.block[i, j] for i in a for j in b) / (la * lb))

    def within(self, a: Sequence[int]) -> float:
        """mean over the unordered pairs of the union of groups a; 1.0 for fewer than two members (:124-125)."""
        m = sum(self.sizes[i] for i in a)
        if m <= 1:
            return 1.0
        return float(sum(self.block[i, j] for i in a for j in a) / (m * (m - 1)))


def _louvain_available() -> bool:
    try:
        import networkx  # type: ignore # noqa: F401
        import community  # type: ignore # noqa: F401
        return True
    except Exception:
        return False


def _normalized_laplacian(W: np.ndarray) -> np.ndarray:
    """Reference :285-292 — L = I - D^-1/2 W D^-1/2 with isolated vertices mapped to 0."""
    d = np.sum(W, axis=1)
    with np.errstate(divide="ignore"):
        d_inv_sqrt = np.where(d > 0, 1.0 / np.sqrt(d), 0.0)
    # (D @ W @ D)[i, j] = (d_i * W_ij) * d_j: the products with D's zeros add exact zeros, so scaling rows then columns
    # gives the reference's matrix bit for bit without its two n^3 products
    return np.eye(W.shape[0], dtype=float) - ((d_inv_sqrt[:, None] * W) * d_inv_sqrt[None, :])


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
        from ..ragged import group_coassociation
        C = group_coassociation(np.stack(label_list)).cpu().numpy()   # ss_group_coassociation (reference :231-241)
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
                             engine: Optional[str] = None, W_override: Optional[np.ndarray] = None,
                             block_sums: Optional[Callable] = None) -> Tuple[List[List[int]], str, np.ndarray]:
    """Reference :343-588 on the device outputs.  Returns (clusters, method_used, W_all).
    ``W_override`` substitutes a pre-built kNN graph (tests pin the host stage with the reference's own W).
    ``block_sums`` (default: the device pass's own ``ragged.DocBlockSums``) maps a list of member lists to
    ``(rowsum, block)``; every block mean of the merge / refine / reassign phases is derived from its output."""
    from ..ragged import knn_graph_from_lists
    sums = block_sums or dp.block_sums
    if sums is None:
        raise RuntimeError("the clustering stage needs the device block sums of sim_sharp (ss_group_block_sums); "
                           "there is no CPU fallback")
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
        # without networkx + python-louvain the modularity engine returns None whatever its input (reference :192-195,
        # 265-267): the eigendecomposition of the RMT filter is then skipped, the outcome is the same
        if _louvain_available():
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
        """(left, right, within(left), within(right)) when the two spectral halves separate (sep < 0), else None."""
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
        gm = _GroupMeans(sums, [left, right])
        w_left, w_right = gm.within([0]), gm.within([1])
        sep = gm.between([0], [1]) - 0.5 * (w_left + w_right)
        return (sorted(left), sorted(right), w_left, w_right) if sep < 0.0 else None

    split_groups: List[List[int]] = []
    for g in groups:
        halves = bisect(g) if len(g) > eff_cap_soft else None
        if halves is not None and all(len(x) >= max(2, small_group_min) for x in halves[:2]):
            split_groups.extend([halves[0], halves[1]])
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
    gm = _GroupMeans(sums, groups) if any(len(g) < max(2, int(min_len)) for g in groups) else None  # one launch for the phase
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
            if gm.between([i], [j]) < float(eff_tau_merge):
                continue
            gain = gm.within([i, j]) - 0.5 * (gm.within([i]) + gm.within([j]))
            if gain > best_gain:
                best_gain, best_j = gain, j
        if best_j is not None and best_gain > 0.0:
            consumed.add(best_j)
            merged.append(sorted(groups[best_j] + g))
        else:
            merged.append(g)

    # ---- refine: split loose clusters, merge near-duplicate neighbours (reference :494-553) ------
    try:
        gm = _GroupMeans(sums, merged)
        internal = [gm.within([i]) for i in range(len(merged))]
        low_thr = float(np.percentile(np.array(internal, dtype=float), 25)) if len(internal) >= 2 else 0.0
        refined: List[List[int]] = []
        for i, g in enumerate(merged):
            if len(g) >= 6 and internal[i] < max(0.5, low_thr):
                halves = bisect(g)
                if halves is not None and halves[2] > internal[i] and halves[3] > internal[i]:
                    refined.append(halves[0])
                    refined.append(halves[1])
                    continue
            refined.append(g)
        global_merge_thr = dp.q60 if has_pos else 0.5
        gm = _GroupMeans(sums, refined)
        merged_adj: List[List[int]] = []
        i = 0
        while i < len(refined):
            cur = refined[i]
            span = [i]                       # `cur` is the multiset union of refined[i .. j-1]
            j = i + 1
            while j < len(refined):
                inter = gm.between(span, [j])
                cmp_thr = 0.9 * min(max(gm.within(span), 1e-6), max(gm.within([j]), 1e-6))
                if inter >= max(cmp_thr, global_merge_thr):
                    cur = sorted(cur + refined[j])
                    span.append(j)
                    j += 1
                else:
                    break
            merged_adj.append(cur)
            i = j
        merged = merged_adj
    except RuntimeError:
        raise                                # device errors are never swallowed (no CPU fallback)
    except Exception:
        pass

    # ---- one-pass sentence reassignment (reference :555-588) -----------------------------------
    if len(merged) >= 2:
        delta = (dp.reassign_delta if has_pos else float(reassign_delta)) if auto_params else float(reassign_delta)
        n_c = len(merged)
        R = np.array(_GroupMeans(sums, merged).rowsum, dtype=np.float64)   # R[x, c] = sum of sim_sharp[x, members of c]
        occ = np.zeros((n_c, n), dtype=np.int64)                           # occ[c, y] = how often y is listed in cluster c
        for c, g in enumerate(merged):
            for y in g:
                if 0 <= y < n:
                    occ[c, y] += 1
        sizes_c = [len(g) for g in merged]
        for x in range(n):
            holders = np.flatnonzero(occ[:, x])
            if holders.size == 0:
                continue
            cur = int(holders[0])
            m_cur = sizes_c[cur] - int(occ[cur, x])        # members of `cur` other than x; sim_sharp[x, x] == 0
            best_c = cur
            best_score = float(R[x, cur] / m_cur) if m_cur else 0.0
            row = R[x]
            for c2 in range(n_c):
                if c2 == cur:
                    continue
                other = float(row[c2] / sizes_c[c2]) if sizes_c[c2] else 0.0
                if other > best_score + float(delta):
                    best_score, best_c = other, c2
            if best_c != cur:
                removed = int(occ[cur, x])
                merged[cur] = [y for y in merged[cur] if y != x]
                merged[best_c] = sorted(merged[best_c] + [x])
                col = sharp[:, x].astype(np.float64)       # the sums of the two clusters change by x's column
                R[:, cur] -= removed * col
                R[:, best_c] += col
                occ[cur, x] = 0
                occ[best_c, x] += 1
                sizes_c[cur] -= removed
                sizes_c[best_c] += 1
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
    passes = device_pass_batch([e for _, _, e in docs], tau=tau,
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
