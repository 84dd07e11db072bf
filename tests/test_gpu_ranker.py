"""GPU tests for the full-score / full-order kernels and the ranker drop-in
(semanticsearch_b200.Tool.rank_chunks_optimized)."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import rank_oracle as ro

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,n", [(1, 1), (1, 2), (3, 100), (2, 2048), (2, 2049), (1, 5000), (1, 70001), (4, 4096)])
def test_rank_order_matches_stable_argsort(b, n):
    from semanticsearch_b200 import similarity
    rng = np.random.default_rng(b * 100003 + n)
    s = rng.standard_normal((b, n)).astype(np.float32)
    s[:, n // 3] = s[:, 0]            # ties
    if n > 10:
        s[0, 5:9] = 0.25
    order, rank1 = similarity.rank_order(torch.from_numpy(s).cuda())
    order, rank1 = order.cpu().numpy(), rank1.cpu().numpy()
    for q in range(b):
        want = ro.rank_order_ref(s[q])
        np.testing.assert_array_equal(order[q], want)
        lookup = np.zeros(n, dtype=np.int64)
        lookup[want] = np.arange(1, n + 1)
        np.testing.assert_array_equal(rank1[q], lookup)


@pytest.mark.parametrize("dtype,d", [(torch.float32, 384), (torch.bfloat16, 768), (torch.float32, 50)])
def test_cosine_scores_full_matrix(dtype, d):
    from semanticsearch_b200 import similarity
    rng = np.random.default_rng(d)
    C = rng.standard_normal((7001, d)).astype(np.float32)
    Q = rng.standard_normal((5, d)).astype(np.float32)
    C[11] = 0.0
    Ct, Qt = torch.from_numpy(C).cuda().to(dtype), torch.from_numpy(Q).cuda().to(dtype)
    got = similarity.cosine_scores(Ct, Qt).cpu().numpy()
    want = ro.cosine_similarity_ref(Qt.float().cpu().numpy(), Ct.float().cpu().numpy())
    np.testing.assert_allclose(got, want, atol=1e-5, rtol=0)
    assert np.all(got[:, 11] == 0.0)


def test_cosine_similarity_dropin_matches_reference_fixture(golden_dir):
    from semanticsearch_b200.Tool import rank_chunks_optimized as R
    g = np.load(os.path.join(golden_dir, "rank_cosine.npz"))
    for b in range(g["Q"].shape[0]):
        got = R.cosine_similarity(g["Q"][b].reshape(1, -1), g["C"])[0]   # the call at rank_chunks_optimized.py:216
        assert got.dtype == np.float32
        np.testing.assert_allclose(got, g["scores"][b], atol=1e-5, rtol=0)


def test_optimized_ranker_dropin(golden_dir):
    from semanticsearch_b200.Tool import Sentence_Embedding as emb
    from semanticsearch_b200.Tool import rank_chunks_optimized as R
    g = np.load(os.path.join(golden_dir, "rank_cosine.npz"))
    C, Q = g["C"], g["Q"]
    texts = [f"chunk {i:04d} alpha" if i % 2 else f"chunk {i:04d} beta" for i in range(len(C))]
    table = {t: v for t, v in zip(texts, C)}
    table["alpha query"] = Q[0]
    emb.set_embedding_backend(lambda text_list, model_name, batch_size=32, device_preference=None:
                              np.stack([table[t] for t in text_list]).astype(np.float32))
    try:
        ranker = R.OptimizedRanker(model_name="m", device_preference="cuda", cache_size=100000)
        df = pd.DataFrame({"chunk_id": [f"c{i}" for i in range(len(C))], "chunk_text": texts})
        ranked = ranker.rank_single_query_optimized("alpha query", df)
        assert list(ranked.columns) == ["chunk_id", "chunk_text", "cosine_score", "bm25_score", "rrf_score"]
        by_id = ranked.set_index("chunk_id").loc[df["chunk_id"]]
        # the cosine column equals the reference ranker's (rank_chunks_optimized.py:201-250 through the shim)
        np.testing.assert_allclose(by_id["cosine_score"].to_numpy(), g["ranker_cosine_q0"], atol=1e-5, rtol=0)
        cos = by_id["cosine_score"].to_numpy(dtype=np.float32)
        bm = by_id["bm25_score"].to_numpy()
        assert np.all(bm >= 0) and bm[1] > 0 and bm[0] == 0
        cr = np.empty(len(C)); cr[ro.rank_order_ref(cos)] = np.arange(1, len(C) + 1)
        br = np.empty(len(C)); br[ro.rank_order_ref(bm.astype(np.float32))] = np.arange(1, len(C) + 1)
        np.testing.assert_allclose(by_id["rrf_score"].to_numpy(), 1.0 / (60 + cr) + 1.0 / (60 + br), rtol=1e-12)
        assert np.all(np.diff(ranked["rrf_score"].to_numpy()) <= 0)
        s, i = ranker.top_k("alpha query", texts, 10)
        assert list(i) == list(ro.rank_order_ref(cos)[:10])
        with pytest.raises(ValueError):
            ranker.rank_single_query_optimized("alpha query", df.rename(columns={"chunk_text": "x"}))
        legacy = R.rank_by_cosine_similarity("alpha query", df, model_name="m", device_preference="cuda")
        assert legacy["cosine_score"].is_monotonic_decreasing
    finally:
        emb.set_embedding_backend(None)


def test_rank_and_filter_end_to_end(tmp_path):
    from semanticsearch_b200.Tool import Sentence_Embedding as emb
    from semanticsearch_b200.Tool import rank_chunks_optimized as R
    rng = np.random.default_rng(0)
    rows, table = [], {}
    for q in range(3):
        table[f"query text {q}"] = rng.standard_normal(32).astype(np.float32)
        for c in range(40):
            t = f"q{q} chunk {c} words {c % 5}"
            table[t] = rng.standard_normal(32).astype(np.float32)
            rows.append({"query_id": f"Q{q}", "chunk_id": f"Q{q}_c{c}", "chunk_text": t})
    chunks = tmp_path / "chunks.tsv"
    pd.DataFrame(rows).to_csv(chunks, sep="\t", index=False)
    orig = tmp_path / "orig.tsv"
    pd.DataFrame({"query_id": [f"Q{q}" for q in range(3)], "query_text": [f"query text {q}" for q in range(3)]}).to_csv(orig, sep="\t", index=False)
    emb.set_embedding_backend(lambda text_list, model_name, batch_size=32, device_preference=None:
                              np.stack([table[t] for t in text_list]).astype(np.float32))
    try:
        out = R.rank_and_filter_chunks_optimized(str(chunks), tmp_path, str(orig), model_name="m")
        assert out.endswith("chunks_rrf_filtered.tsv")
        res = pd.read_csv(out, sep="\t")
        assert list(res.columns) == ["query_id", "chunk_text", "label"]
        assert set(res["label"]) == {0, 1} and res["query_id"].nunique() == 3
        assert 3 * 14 <= len(res) <= 3 * 20   # ~ top 20% + bottom 20% of 40 chunks per query
        assert R.rank_and_filter_chunks_optimized(str(tmp_path / "missing.tsv"), tmp_path, str(orig)) == ""
    finally:
        emb.set_embedding_backend(None)


def test_segmented_rank_rrf_matches_numpy():
    """K8 against the reference expressions (rank_chunks_optimized.py:215-250,518-519) evaluated with numpy on
    the same inputs: cosine within 1e-5, ranks / fused scores / order / percentile thresholds exact given the
    kernel's own cosine values (ties -> lower row first)."""
    from semanticsearch_b200 import similarity
    rng = np.random.default_rng(12)
    sizes = [2, 3, 40, 257, 1000, 1, 64, 5000]
    d = 96
    C = rng.standard_normal((sum(sizes), d)).astype(np.float32)
    Q = rng.standard_normal((len(sizes), d)).astype(np.float32)
    bm = np.maximum(rng.standard_normal(sum(sizes)), 0.0).astype(np.float32)   # many exact zeros: ties
    C[5] = 0.0
    C[50] = C[49]                                                                # duplicate chunk: cosine tie
    off = np.zeros(len(sizes) + 1, dtype=np.int32)
    off[1:] = np.cumsum(sizes)
    out = similarity.segmented_rank_rrf(torch.from_numpy(C).cuda(), torch.from_numpy(off).cuda(), torch.from_numpy(Q).cuda(),
                                        torch.from_numpy(bm).cuda(), upper_percentile=80, lower_percentile=20)
    torch.cuda.synchronize()
    cos, rc, rb = out["cosine"].cpu().numpy(), out["rank_cosine"].cpu().numpy(), out["rank_bm25"].cpu().numpy()
    rrf, order, thr = out["rrf"].cpu().numpy(), out["order"].cpu().numpy(), out["thresholds"].cpu().numpy()
    for g, n in enumerate(sizes):
        a, b = off[g], off[g + 1]
        want_cos = ro.cosine_similarity_ref(Q[g:g + 1], C[a:b])[0]
        np.testing.assert_allclose(cos[a:b], want_cos, atol=1e-5, rtol=0)
        for scores, ranks in ((cos[a:b], rc[a:b]), (bm[a:b], rb[a:b])):
            lookup = np.empty(n, dtype=np.int64)
            lookup[ro.rank_order_ref(scores)] = np.arange(1, n + 1)              # np.argsort(-s) + rank lookup, stable ties
            np.testing.assert_array_equal(ranks, lookup)
        want_rrf = 1.0 / (60 + rc[a:b].astype(np.float64)) + 1.0 / (60 + rb[a:b].astype(np.float64))
        np.testing.assert_array_equal(rrf[a:b], want_rrf)                        # bit-exact fp64
        np.testing.assert_array_equal(order[a:b], np.argsort(-want_rrf, kind="stable"))
        assert thr[g, 0] == np.percentile(want_rrf, 80) and thr[g, 1] == np.percentile(want_rrf, 20)


def test_batched_group_ranking_equals_per_query_path():
    from semanticsearch_b200.Tool import Sentence_Embedding as emb
    from semanticsearch_b200.Tool import rank_chunks_optimized as R
    rng = np.random.default_rng(3)
    table, groups = {}, []
    for q in range(4):
        table[f"query text {q} words"] = rng.standard_normal(48).astype(np.float32)
        rows = []
        for c in range(30 + 7 * q):
            t = f"q{q} chunk {c} words {c % 4} text"
            table[t] = rng.standard_normal(48).astype(np.float32)
            rows.append({"query_id": f"Q{q}", "chunk_id": f"Q{q}_c{c}", "chunk_text": t})
        groups.append((f"query text {q} words", pd.DataFrame(rows)))
    emb.set_embedding_backend(lambda text_list, model_name, batch_size=32, device_preference=None:
                              np.stack([table[t] for t in text_list]).astype(np.float32))
    try:
        ranker = R.OptimizedRanker(model_name="m", device_preference="cuda", cache_size=100000)
        batched = R.rank_query_groups_batched(ranker, groups, 80, 20)
        for (query, df), (ranked, pos_thr, neg_thr) in zip(groups, batched):
            single = ranker.rank_single_query_optimized(query, df)
            assert list(ranked.columns) == list(single.columns)
            a = ranked.set_index("chunk_id").loc[df["chunk_id"]]
            b = single.set_index("chunk_id").loc[df["chunk_id"]]
            np.testing.assert_allclose(a["cosine_score"].to_numpy(), b["cosine_score"].to_numpy(), atol=1e-6, rtol=0)
            np.testing.assert_array_equal(a["bm25_score"].to_numpy(), b["bm25_score"].to_numpy())
            np.testing.assert_allclose(a["rrf_score"].to_numpy(), b["rrf_score"].to_numpy(), rtol=1e-12)
            assert np.all(np.diff(ranked["rrf_score"].to_numpy()) <= 0)
            labelled = R._label_group(single, 80, 20)
            sel = ranked[(ranked["rrf_score"] >= pos_thr) | (ranked["rrf_score"] <= neg_thr)]
            assert sorted(sel["chunk_id"]) == sorted(labelled["chunk_id"])
    finally:
        emb.set_embedding_backend(None)


def _ranker_with_table(n=300, d=48, seed=5, cuda_encoder=False):
    from semanticsearch_b200.Tool import Sentence_Embedding as emb
    from semanticsearch_b200.Tool import rank_chunks_optimized as R
    rng = np.random.default_rng(seed)
    texts = [f"chunk {i:04d} " + ("alpha" if i % 2 else "beta") for i in range(n)]
    table = {t: rng.standard_normal(d).astype(np.float32) for t in texts}
    table["alpha query"] = rng.standard_normal(d).astype(np.float32)
    table["beta query"] = rng.standard_normal(d).astype(np.float32)
    calls = []

    def backend(text_list, model_name, batch_size=32, device_preference=None):
        calls.append(list(text_list))
        if any(t == "poison" for t in text_list):
            raise RuntimeError("encoder exploded")
        out = np.stack([table[t] for t in text_list]).astype(np.float32)
        return torch.from_numpy(out).cuda() if cuda_encoder else out

    emb.set_embedding_backend(backend)
    return R, R.OptimizedRanker(model_name="m", device_preference="cuda", cache_size=100000), texts, table, calls


def test_device_row_store_serves_cached_chunks_without_host_traffic():
    """§8f-4: a second query over cached chunks encodes nothing and uploads no chunk row; results equal the uncached path."""
    from semanticsearch_b200.Tool import Sentence_Embedding as emb
    R, ranker, texts, table, calls = _ranker_with_table()
    try:
        df = pd.DataFrame({"chunk_id": [f"c{i}" for i in range(len(texts))], "chunk_text": texts})
        first = ranker.rank_single_query_optimized("alpha query", df)
        store = ranker.chunk_embedding_cache
        assert store.h2d_rows == len(texts) and len(store) == len(texts)
        n_calls = len(calls)
        second = ranker.rank_single_query_optimized("beta query", df.iloc[::-1].reset_index(drop=True))
        assert store.h2d_rows == len(texts)                       # no chunk row crossed the link again
        assert [c for c in calls[n_calls:]] == [["beta query"]]   # only the new query went through the encoder
        s, i = ranker.top_k("alpha query", texts[:100], k=7)
        assert store.h2d_rows == len(texts)
        C = np.stack([table[t] for t in texts[:100]])
        want = ro.cosine_topk_ref(table["alpha query"][None, :], C, 7)
        np.testing.assert_array_equal(i, want[1][0])
        np.testing.assert_allclose(s, want[0][0], atol=1e-5, rtol=0)
        # the host view of the cache keeps the reference's contract
        got = ranker.get_chunk_embeddings_batch(texts[5:9] + texts[5:6])
        np.testing.assert_array_equal(got, np.stack([table[t] for t in texts[5:9] + texts[5:6]]))
        fresh = R.OptimizedRanker(model_name="m", device_preference="cuda", cache_size=100000)
        again = fresh.rank_single_query_optimized("beta query", df.iloc[::-1].reset_index(drop=True))
        pd.testing.assert_frame_equal(second, again)
        assert first.shape == second.shape
    finally:
        emb.set_embedding_backend(None)


def test_device_row_store_eviction_and_cuda_encoder_handoff():
    """The oldest quarter leaves when the cache overflows (reference :131-139), evicted chunks are re-encoded on demand, and
    an encoder that returns CUDA tensors never touches the host link."""
    from semanticsearch_b200.Tool import Sentence_Embedding as emb
    R, ranker, texts, table, calls = _ranker_with_table(n=64, cuda_encoder=True)
    try:
        ranker.cache_size = ranker.chunk_embedding_cache.cache_size = 40
        rows = ranker.get_chunk_embeddings_device(texts)             # 64 > 40: 16 oldest evicted after the gather
        store = ranker.chunk_embedding_cache
        assert rows.is_cuda and rows.shape == (64, 48) and len(store) == 48 and store.h2d_rows == 0 and store.d2d_rows == 64
        np.testing.assert_array_equal(rows.cpu().numpy(), np.stack([table[t] for t in texts]))
        n_calls = len(calls)
        again = ranker.get_chunk_embeddings_device(texts[:20])        # 16 evicted chunks come back through the encoder
        assert calls[n_calls:] == [texts[:16]]
        np.testing.assert_array_equal(again.cpu().numpy(), np.stack([table[t] for t in texts[:20]]))
    finally:
        emb.set_embedding_backend(None)


def test_one_poisoned_group_does_not_drop_the_block():
    """a group whose encoding fails is reported and skipped; the other groups of the block are still ranked
    (reference :488-536 isolates every query); a device failure propagates instead."""
    from semanticsearch_b200 import _lib
    from semanticsearch_b200.Tool import Sentence_Embedding as emb
    R, ranker, texts, table, calls = _ranker_with_table(n=40)
    try:
        rows = []
        for qid, qtext in ((1, "alpha query"), (2, "beta query"), (3, "alpha query")):
            for j in range(10):
                t = texts[(qid - 1) * 10 + j]
                rows.append({"query_id": qid, "query_text": qtext, "chunk_id": f"c{qid}_{j}", "chunk_text": "poison" if (qid == 2 and j == 3) else t})
        df = pd.DataFrame(rows)
        kept = R._process_queries_sequential(df, "m", 80, 20)
        got_qids = sorted({int(k["query_id"].iloc[0]) for k in kept})
        assert got_qids == [1, 3]                      # group 2 failed alone
        for k in kept:
            assert set(k["label"]) <= {0, 1} and len(k) >= 2

        def broken(*a, **kw):
            raise _lib.DeviceError("simulated kernel failure")
        orig = R.rank_query_groups_batched
        R.rank_query_groups_batched = broken
        try:
            with pytest.raises(_lib.DeviceError):
                R._process_queries_sequential(df[df["query_id"] != 2], "m", 80, 20)
        finally:
            R.rank_query_groups_batched = orig
    finally:
        emb.set_embedding_backend(None)


@pytest.mark.parametrize("algo,dtype,b", [("stream", torch.float32, 1), ("stream", torch.bfloat16, 3), ("tcstream", torch.bfloat16, 4),
                                          ("gemm", torch.bfloat16, 130), ("small", torch.float32, 12)])
def test_k_larger_than_the_corpus_pads_consistently(algo, dtype, b):
    """k > n_rows returns the n real rows, then score -inf / index -1 / key 0 on every kernel path."""
    from semanticsearch_b200 import similarity
    g = torch.Generator(device="cuda").manual_seed(3)
    C = torch.randn((6, 64), generator=g, device="cuda").to(dtype)
    Q = torch.randn((b, 64), generator=g, device="cuda").to(dtype)
    s, i, keys = similarity.cosine_topk(C, Q, 10, algo=algo, return_keys=True)
    assert s.shape == (b, 10) and bool((i[:, 6:] == -1).all()) and bool(torch.isinf(s[:, 6:]).all()) and bool((keys[:, 6:] == 0).all())
    want = ro.cosine_topk_ref(Q.float().cpu().numpy(), C.float().cpu().numpy(), 6)
    np.testing.assert_array_equal(i[:, :6].cpu().numpy(), want[1])
