"""Pins oracle/*.py (the numpy restatement) to outputs of the unmodified reference.

The committed fixtures in tests/golden/ were produced by ``python -m oracle.gen_golden`` from
the live reference (see that script).  When /root/reference is present (build container) the
same comparisons are repeated against the reference executed on fresh random inputs.
"""
import json
import os

import numpy as np
import pytest

from oracle import grouping_oracle as go
from oracle import rank_oracle as ro
from oracle import simmatrix_oracle as so
from oracle import splitter_oracle as spo
from oracle import ref_shim


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_rank_cosine_matches_reference(golden_dir):
    g = _load(golden_dir, "rank_cosine.npz")
    C, Q, scores = g["C"], g["Q"], g["scores"]
    mine = ro.cosine_similarity_ref(Q, C)
    assert mine.dtype == np.float32
    np.testing.assert_allclose(mine, scores, atol=2e-7, rtol=0)
    # zero row scores exactly 0 under sklearn's zero rule
    assert np.all(mine[:, 17] == 0.0)
    # ranker path (OptimizedRanker.rank_single_query_optimized) returns the same cosine column
    np.testing.assert_allclose(mine[0], g["ranker_cosine_q0"], atol=2e-7, rtol=0)


def test_rank_order_matches_reference_up_to_ties(golden_dir):
    g = _load(golden_dir, "rank_cosine.npz")
    scores, order = g["scores"], g["order"]
    for b in range(scores.shape[0]):
        mine = ro.rank_order_ref(scores[b])
        # same score sequence; index differences only inside runs of equal scores
        np.testing.assert_array_equal(scores[b][mine], scores[b][order[b]])
        diff = np.nonzero(mine != order[b])[0]
        for p in diff:
            assert scores[b][mine[p]] == scores[b][order[b][p]]
    s, i = ro.cosine_topk_ref(g["Q"], g["C"], 10)
    assert s.shape == (6, 10) and i.dtype == np.int64
    assert ro.check_topk_against_oracle(g["Q"], g["C"], s, i, 10, 1e-5) == []


def test_simmatrix_matches_reference(golden_dir):
    g = _load(golden_dir, "grouping.npz")
    for name in ("a", "b", "c", "tiny", "seven"):
        S = so.similarity_matrix_ref(g[f"{name}_E"])
        assert S.dtype == np.float32
        np.testing.assert_allclose(S, g[f"{name}_S"], atol=3e-7, rtol=0)
    assert so.similarity_matrix_ref(g["a_E"][:1]) is None
    # zero sentence vector => zero row and column, including the diagonal
    S = so.similarity_matrix_ref(g["a_E"])
    assert np.all(S[5] == 0) and np.all(S[:, 5] == 0)


def test_similarity_distribution_matches_reference(golden_dir):
    g = _load(golden_dir, "grouping.npz")
    meta = json.loads(str(g["meta_json"]))
    for name in ("a", "b", "c", "tiny", "seven"):
        mine = so.analyze_similarity_distribution_ref(g[f"{name}_S"])
        want = meta[f"{name}_dist"]
        assert set(mine) == set(want)
        for key in want:
            assert mine[key] == pytest.approx(want[key], abs=1e-12), (name, key)


def test_grouping_pass_matches_reference(golden_dir):
    g = _load(golden_dir, "grouping.npz")
    meta = json.loads(str(g["meta_json"]))
    for name in ("a", "b", "c", "tiny", "seven"):
        S = g[f"{name}_S"]
        sc = meta[f"{name}_scalars"]
        res = go.grouping_pass_ref(S)
        assert res["sim_sharp"].dtype == np.float32
        np.testing.assert_array_equal(res["sim_sharp"], g[f"{name}_sim_sharp"])
        np.testing.assert_array_equal(res["centrality"], g[f"{name}_centrality"])
        assert res["mu"] == sc["mu"] and res["sigma"] == sc["sigma"]
        assert res["k_eff_all"] == sc["k_eff_all"]
        assert res["thresholds"]["edge_floor"] == sc["eff_edge_floor"]
        if sc["eff_tau_merge"] is not None:
            assert res["thresholds"]["tau_merge"] == sc["eff_tau_merge"]
        if sc["global_merge_thr"] is not None:
            assert res["thresholds"]["global_merge_thr"] == sc["global_merge_thr"]
        if sc["eff_reassign_delta"] is not None:
            assert res["thresholds"]["reassign_delta"] == sc["eff_reassign_delta"]
        # identical except where the reference's non-stable argsort had a free choice (ties)
        assert go.knn_graph_mismatches(res["W_all"], g[f"{name}_W_all"], res["sim_sharp"], res["k_eff_all"]) == []
        assert np.count_nonzero(res["W_all"]) == pytest.approx(np.count_nonzero(g[f"{name}_W_all"]), rel=0.05)


def test_splitter_pass_matches_reference(golden_dir):
    g = _load(golden_dir, "splitter.npz")
    meta = json.loads(str(g["meta_json"]))
    for name in ("s40", "s97", "s12", "s3"):
        E = g[f"{name}_E"]
        np.testing.assert_array_equal(so.normalize_rows_1e9(E), g[f"{name}_En"])
        adj = spo.adjacent_sims_ref(E)
        np.testing.assert_array_equal(adj, g[f"{name}_adj_sims"])
        st = spo.robust_stats_ref(adj, 3)
        np.testing.assert_array_equal(st["adj_base"], g[f"{name}_adj_base"])
        np.testing.assert_array_equal(st["adj_for_valley"], g[f"{name}_adj_for_valley"])
        assert st["valley_tau"] == meta[name]["valley_tau"]
        if f"{name}_c99_S" in g.files:
            S = spo.c99_similarity_ref(g[f"{name}_En"])
            np.testing.assert_array_equal(S, g[f"{name}_c99_S"])
            if meta[name]["kwargs"].get("c99_use_local_rank"):
                R = spo.c99_local_rank_ref(S, 11)
            else:
                R = spo.c99_global_rank_ref(S)
            np.testing.assert_array_equal(R, g[f"{name}_c99_R"])


def test_c99_divisive_search_matches_reference(golden_dir):
    """The reference's own _c99_boundaries outputs (boundaries, pick order, density profile) for 14 cases."""
    g = _load(golden_dir, "c99_cuts.npz")
    meta = json.loads(str(g["meta_json"]))
    assert len(meta) >= 14
    for name, m in meta.items():
        kw = dict(m["kwargs"])
        S = spo.c99_similarity_ref(g[f"{name}_En"])
        if kw.pop("use_local_rank", False):
            R = spo.c99_local_rank_ref(S, kw.pop("mask_size", 11))
        else:
            R = spo.c99_global_rank_ref(S)
        if f"{name}_R" in g.files:
            np.testing.assert_array_equal(R, g[f"{name}_R"])
        bounds, picked, series = spo.c99_divisive_ref(R, **kw)
        assert bounds == m["bounds"], name
        assert picked == m["cuts"], name
        np.testing.assert_array_equal(np.asarray(series, dtype=np.float64), g[f"{name}_D"])


def test_p95_breakpoints_known_answer():
    adj = np.array([0.9, 0.8, 0.1, 0.85, 0.95, 0.2, 0.7, 0.75, 0.6, 0.65, 0.3], dtype=float)
    thr, bp = spo.p95_breakpoints_ref(adj)
    d = 1.0 - adj
    assert thr == pytest.approx(np.percentile(d, 95))
    assert bp.tolist() == [2]
    thr, bp = spo.p95_breakpoints_ref(np.array([0.5]))
    assert bp.size == 0 and thr == 0.5


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_oracle_against_live_reference():
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(7)
    for trial in range(3):
        n, d = int(rng.integers(8, 90)), int(rng.integers(8, 64))
        E = rng.standard_normal((n, d)).astype(np.float32)
        text, sents = ref_shim.make_doc(E, tag=f"t{trial}")
        S_ref = ref.common.create_similarity_matrix(sents, "m", device="cpu")
        np.testing.assert_array_equal(so.similarity_matrix_ref(E), S_ref)
        C = rng.standard_normal((300, d)).astype(np.float32)
        q = rng.standard_normal((1, d)).astype(np.float32)
        np.testing.assert_allclose(ro.cosine_similarity_ref(q, C), ref.rank.cosine_similarity(q, C), atol=2e-7, rtol=0)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_c99_oracle_against_live_reference():
    """Fresh random documents through the reference's own _c99_boundaries (Splitter:155-264), every run with different
    sizes / chunk floors / stopping rules: boundaries, pick order and density profile must be the reference's."""
    from oracle.gen_golden import capture_locals, topic_doc
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(2024)
    for trial in range(8):
        n = int(rng.integers(6, 70))
        E = topic_doc(rng, n, 24, sent_per_topic=int(rng.integers(3, 10)), noise=0.6)
        En = (E / np.linalg.norm(E, axis=1, keepdims=True)).astype(np.float32)
        kw = {"min_chunk_size": int(rng.integers(1, 6)), "stopping": "profile" if trial % 3 == 2 else "gain",
              "min_gain": float(rng.choice([0.0, 0.01, 2.0])), "use_local_rank": trial % 4 == 3}
        if trial % 5 == 4:
            kw["max_cuts"] = int(rng.integers(1, 4))
        want, grabbed = capture_locals(ref.split._c99_boundaries, ("R", "cuts", "D_series"), {"_c99_boundaries"}, En, **kw)
        loc = grabbed.get("_c99_boundaries", {})
        S = spo.c99_similarity_ref(En)
        R = spo.c99_local_rank_ref(S, 11) if kw["use_local_rank"] else spo.c99_global_rank_ref(S)
        if "R" in loc:
            np.testing.assert_array_equal(R, loc["R"])
        got, picked, series = spo.c99_divisive_ref(R, kw["min_chunk_size"], kw.get("max_cuts"), kw["min_gain"], kw["stopping"])
        assert got == [int(x) for x in want], (trial, kw)
        assert picked == [int(x) for x in loc.get("cuts", [])], (trial, kw)
        np.testing.assert_array_equal(np.asarray(series, dtype=np.float64), np.asarray(loc.get("D_series", []), dtype=np.float64))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_splitter_and_grouping_oracles_against_live_reference():
    """Random documents through process_sentence_splitting_with_semantics and semantic_grouping_main: the locals the
    reference computes on the dense path equal the oracle's restatement bit for bit."""
    from oracle.gen_golden import capture_locals, topic_doc
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(77)
    for trial in range(3):
        n = int(rng.integers(12, 60))
        E = topic_doc(rng, n, 32, sent_per_topic=int(rng.integers(4, 10)), noise=0.6)
        text, _sents = ref_shim.make_doc(E, tag=f"lv{trial}")
        _out, grabbed = capture_locals(ref.split.process_sentence_splitting_with_semantics,
                                       ("embeddings", "adj_sims", "adj_base", "adj_for_valley", "valley_tau"),
                                       {"process_sentence_splitting_with_semantics"}, text, embedding_model="m",
                                       device="cpu", silent=True)
        loc = grabbed["process_sentence_splitting_with_semantics"]
        np.testing.assert_array_equal(so.normalize_rows_1e9(E), np.asarray(loc["embeddings"]))
        adj = spo.adjacent_sims_ref(E)
        np.testing.assert_array_equal(adj, np.asarray(loc["adj_sims"], dtype=np.float64))
        st = spo.robust_stats_ref(adj, 3)
        np.testing.assert_array_equal(st["adj_base"], np.asarray(loc["adj_base"], dtype=np.float64))
        np.testing.assert_array_equal(st["adj_for_valley"], np.asarray(loc["adj_for_valley"], dtype=np.float64))
        assert st["valley_tau"] == float(loc["valley_tau"])
        _out, grabbed = capture_locals(ref.group.semantic_grouping_main, ("sim_sharp", "centrality", "eff_edge_floor", "mu", "sigma"),
                                       {"semantic_grouping_main"}, text, f"doc_lv{trial}", "m", device="cpu", silent=True,
                                       collect_metadata=True)
        loc = grabbed.get("semantic_grouping_main", {})
        if "sim_sharp" not in loc:
            continue
        res = go.grouping_pass_ref(so.similarity_matrix_ref(E))
        np.testing.assert_array_equal(res["sim_sharp"], np.asarray(loc["sim_sharp"]))
        np.testing.assert_array_equal(res["centrality"], np.asarray(loc["centrality"]))
        assert res["thresholds"]["edge_floor"] == float(loc["eff_edge_floor"])
        assert res["mu"] == float(loc["mu"]) and res["sigma"] == float(loc["sigma"])
