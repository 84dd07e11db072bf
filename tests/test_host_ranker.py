"""CPU tests of the host-side pieces of the ranking drop-in (Tool/rank_chunks_optimized.py): BM25 (lexical side channel,
stays on the host), the legacy RRF merge, percentile labelling, cache eviction and error behaviour.  Where the reference
tree is mounted the pandas/numpy functions are compared with the reference's own (its BM25 dependency is stubbed, so BM25
is checked against the published formula instead)."""
import math

import numpy as np
import pandas as pd
import pytest

from oracle import ref_shim
from semanticsearch_b200.Tool import rank_chunks_optimized as R


def test_bm25_okapi_matches_the_published_formula():
    docs = [["a", "b", "a", "c"], ["b", "c"], ["a"], ["d", "d", "d", "b"], ["b"]]
    bm = R.BM25Okapi(docs, epsilon=0.25)
    n = len(docs)
    avgdl = sum(len(d) for d in docs) / n
    df = {w: sum(w in d for d in docs) for w in "abcd"}
    idf = {w: math.log(n - f + 0.5) - math.log(f + 0.5) for w, f in df.items()}
    mean_idf = sum(idf.values()) / len(idf)
    idf = {w: (v if v >= 0 else 0.25 * mean_idf) for w, v in idf.items()}   # rank_bm25's epsilon floor ("b" is in 4 of 5)
    assert idf["b"] == pytest.approx(0.25 * mean_idf) and bm.idf["b"] == pytest.approx(idf["b"])
    query = ["a", "b", "zzz"]
    want = np.zeros(n)
    for w in query:
        for i, d in enumerate(docs):
            f = d.count(w)
            want[i] += idf.get(w, 0.0) * f * 2.5 / (f + 1.5 * (1 - 0.75 + 0.75 * len(d) / avgdl))
    np.testing.assert_allclose(bm.get_scores(query), want, rtol=1e-12, atol=0)


def _frames(rng, n):
    ids = [f"c{i}" for i in range(n)]
    cos = pd.DataFrame({"chunk_id": ids, "cosine_score": rng.random(n).astype(np.float32), "chunk_text": [f"t{i}" for i in range(n)]})
    bm = pd.DataFrame({"chunk_id": ids, "bm25_score": rng.random(n), "chunk_text": [f"t{i}" for i in range(n)]})
    return (cos.sort_values("cosine_score", ascending=False).reset_index(drop=True),
            bm.sort_values("bm25_score", ascending=False).reset_index(drop=True))


def test_rank_by_rrf_known_answer_and_errors():
    rng = np.random.default_rng(3)
    cos, bm = _frames(rng, 7)
    out = R.rank_by_rrf(cos, bm.iloc[:5], k=60)       # two chunks missing from the BM25 side get the worst rank 7 + 60
    rc = {cid: r for r, cid in enumerate(cos["chunk_id"], start=1)}
    rb = {cid: r for r, cid in enumerate(bm["chunk_id"].iloc[:5], start=1)}
    for _, row in out.iterrows():
        want = 1.0 / (60 + rc[row["chunk_id"]]) + 1.0 / (60 + rb.get(row["chunk_id"], 67))
        assert row["rrf_score"] == pytest.approx(want, rel=1e-15)
    assert list(out["rrf_score"]) == sorted(out["rrf_score"], reverse=True)
    with pytest.raises(ValueError):
        R.rank_by_rrf(cos.drop(columns=["chunk_id"]), bm)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_rrf_merge_matches_the_live_reference():
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(4)
    for n in (5, 40, 200):
        cos, bm = _frames(rng, n)
        pd.testing.assert_frame_equal(R.rank_by_rrf(cos, bm), ref.rank.rank_by_rrf(cos, bm))
        pd.testing.assert_frame_equal(R.rank_by_rrf(cos, bm.iloc[: n // 2], k=10), ref.rank.rank_by_rrf(cos, bm.iloc[: n // 2], k=10))


def test_label_group_percentile_rule():
    """Reference :517-526: keep rows with rrf >= P80 (label 1) or <= P20 (label 0); ties at the thresholds are kept."""
    rrf = np.array([0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.07, 0.08, 0.09, 0.10, 0.10])
    ranked = pd.DataFrame({"chunk_id": range(len(rrf)), "rrf_score": rrf})
    kept = R._label_group(ranked, 80, 20)
    hi, lo = np.percentile(rrf, 80), np.percentile(rrf, 20)
    assert set(kept["chunk_id"]) == {i for i, v in enumerate(rrf) if v >= hi or v <= lo}
    assert all((kept["label"] == 1) == (kept["rrf_score"] >= hi))
    same = pd.DataFrame({"chunk_id": [0, 1, 2], "rrf_score": [0.5, 0.5, 0.5]})
    assert list(R._label_group(same, 80, 20)["label"]) == [1, 1, 1]   # every row is >= P80: positive wins, as in the reference


def test_chunk_cache_drops_the_oldest_quarter():
    """Reference :131-139, on the bookkeeping of the device row store (md5 -> slot; the rows themselves live in HBM)."""
    r = R.OptimizedRanker(cache_size=8)
    store = r.chunk_embedding_cache
    for i in range(9):
        store.slots[f"k{i}"] = i
    r._manage_cache_size()
    assert list(store.keys()) == [f"k{i}" for i in range(2, 9)]   # 9 // 4 = 2 oldest entries evicted
    assert sorted(store.free) == [0, 1]                            # their slots are reusable
    r._manage_cache_size()
    assert len(store) == 7 and "k2" in store and "k0" not in store  # at or under the limit: untouched


def test_bm25_float64_order_survives_the_fp32_rank_keys():
    """BM25 scores that differ below fp32 resolution keep their float64 order (reference :226-235 ranks float64 values)."""
    scores = np.array([1.0, 1.0 + 1e-12, 1.0 - 1e-12, 0.5, 1.0 + 1e-12, 0.0])
    key = R._f64_rank_surrogate(scores)
    assert key.dtype == np.float32
    order = np.lexsort((np.arange(len(key)), -key.astype(np.float64)))
    assert list(order) == list(np.argsort(-scores, kind="stable")) == [1, 4, 0, 2, 3, 5]


def test_column_aliases_and_memory_estimate(tmp_path):
    df = pd.DataFrame({"qid": [1], "Question": ["q"], "passage": ["p"], "pid": ["c"], "target": [1]})
    assert list(R._standardize_columns(df).columns) == ["query_id", "query_text", "chunk_text", "chunk_id", "label"]
    assert R.estimate_memory_usage(str(tmp_path / "nope.tsv")) == {"error": "File not found"}
    f = tmp_path / "x.tsv"
    f.write_text("a\tb\n" * 1000)
    est = R.estimate_memory_usage(str(f))
    assert est["peak_memory_gb"] == pytest.approx(est["estimated_memory_gb"] * 1.7) and est["recommended_chunk_size"] == 50000
