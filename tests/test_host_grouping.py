"""CPU tests of the grouping drop-in's host stage: fed with the reference's own intermediate
arrays (tests/golden/grouping.npz, captured from the unmodified reference), the re-implemented
sequential clustering must reproduce the reference's clusters and metadata exactly."""
import json
import os

import numpy as np
import pytest

from oracle import grouping_oracle as go
from semanticsearch_b200.Method import Semantic_Grouping_Optimized as G


def _device_pass_from_golden(g, meta, name):
    sc = meta[f"{name}_scalars"]
    sharp = g[f"{name}_sim_sharp"]
    n = sharp.shape[0]
    idx, val = go.knn_lists_ref(sharp, sc["k_eff_all"])
    kidx = np.full((n, 33), -1, np.int32)
    kval = np.zeros((n, 33), np.float32)
    kidx[:, :idx.shape[1]] = idx
    kval[:, :val.shape[1]] = val
    thr = go.thresholds_ref(sharp)
    return G.DevicePass(sim_matrix=g[f"{name}_S"], sim_sharp=sharp, centrality=g[f"{name}_centrality"], mu=sc["mu"],
                        sigma=sc["sigma"], q80=thr["edge_floor"], q65=thr["tau_merge"], q60=thr["global_merge_thr"],
                        reassign_delta=thr["reassign_delta"], n_positive=thr["count"], k_all=sc["k_eff_all"],
                        knn_idx=kidx, knn_val=kval)


@pytest.mark.parametrize("name", ["a", "b", "c", "tiny", "seven"])
def test_host_stage_reproduces_reference_clusters(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "grouping.npz"))
    meta = json.loads(str(g["meta_json"]))
    dp = _device_pass_from_golden(g, meta, name)
    merged, method, W = G.cluster_from_device_pass(dp, W_override=g[f"{name}_W_all"],
                                                  block_sums=lambda groups: go.block_sums_ref(dp.sim_sharp, groups))
    assert method == meta[f"{name}_scalars"]["method_used"]
    n = dp.sim_sharp.shape[0]
    sentences = [f"s{i}" for i in range(n)]
    chunks = G._emit(f"doc_{name}", "whole", sentences, merged, method, dp, collect_metadata=True)
    want = meta[f"{name}_chunks"]
    assert [c[0] for c in chunks] == [w[0] for w in want]
    for (cid, _text, mjson), (wid, wjson) in zip(chunks, want):
        assert json.loads(mjson) == json.loads(wjson), cid


def test_knn_graph_from_lists_matches_oracle(golden_dir):
    from semanticsearch_b200.ragged import knn_graph_from_lists
    g = np.load(os.path.join(golden_dir, "grouping.npz"))
    meta = json.loads(str(g["meta_json"]))
    for name in ("a", "b", "c", "seven"):
        sharp = g[f"{name}_sim_sharp"]
        sc = meta[f"{name}_scalars"]
        idx, val = go.knn_lists_ref(sharp, sc["k_eff_all"])
        W = knn_graph_from_lists(idx, val, sc["eff_edge_floor"])
        np.testing.assert_array_equal(W, go.knn_graph_ref(sharp, sc["k_eff_all"], sc["eff_edge_floor"]))


def test_sentinels_without_gpu():
    from semanticsearch_b200.Tool import Sentence_Segmenter as seg
    seg.set_sentence_splitter(lambda t: [s for s in t.split("|") if s])
    try:
        assert G.semantic_grouping_main("", "d0", "m", silent=True) == []
        assert G.semantic_grouping_main("only one sentence here", "d1", "m", silent=True) == [("d1_single", "only one sentence here", None)]
    finally:
        seg.set_sentence_splitter(None)


@pytest.mark.skipif(not __import__("oracle.ref_shim", fromlist=["x"]).reference_available(), reason="reference tree not mounted")
def test_host_stage_matches_live_reference_on_random_documents():
    """Fresh random documents through the reference's semantic_grouping_main (CPU): the host stage, fed with the oracle's
    restatement of the device pass and the reference's own kNN graph (its argsort ties are unspecified), must emit the
    same clusters, method and metadata."""
    from oracle import ref_shim, simmatrix_oracle as so
    from oracle.gen_golden import capture_locals, topic_doc
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(909)
    checked = 0
    for trial in range(6):
        n = int(rng.integers(8, 70))
        E = topic_doc(rng, n, 32, sent_per_topic=int(rng.integers(3, 12)), noise=float(rng.choice([0.4, 0.7])))
        text, sents = ref_shim.make_doc(E, tag=f"hg{trial}")
        want, grabbed = capture_locals(ref.group.semantic_grouping_main, ("sim_sharp", "W_all", "k_eff_all", "method_used"),
                                       {"semantic_grouping_main"}, text, f"doc_hg{trial}", "m", device="cpu", silent=True,
                                       collect_metadata=True)
        loc = grabbed.get("semantic_grouping_main", {})
        if "W_all" not in loc:
            continue
        S = so.similarity_matrix_ref(E)
        res = go.grouping_pass_ref(S)
        np.testing.assert_array_equal(res["sim_sharp"], np.asarray(loc["sim_sharp"]))
        idx, val = go.knn_lists_ref(res["sim_sharp"], res["k_eff_all"])
        kidx = np.full((n, 33), -1, np.int32)
        kval = np.zeros((n, 33), np.float32)
        kidx[:, :idx.shape[1]] = idx
        kval[:, :val.shape[1]] = val
        thr = res["thresholds"]
        dp = G.DevicePass(sim_matrix=S, sim_sharp=res["sim_sharp"], centrality=res["centrality"], mu=res["mu"], sigma=res["sigma"],
                          q80=thr["edge_floor"], q65=thr["tau_merge"], q60=thr["global_merge_thr"],
                          reassign_delta=thr["reassign_delta"], n_positive=thr["count"], k_all=res["k_eff_all"],
                          knn_idx=kidx, knn_val=kval)
        merged, method, _W = G.cluster_from_device_pass(dp, W_override=np.asarray(loc["W_all"]),
                                                       block_sums=lambda groups, _s=res["sim_sharp"]: go.block_sums_ref(_s, groups))
        assert method == str(loc["method_used"])
        chunks = G._emit(f"doc_hg{trial}", text, sents, merged, method, dp, collect_metadata=True)
        assert [c[0] for c in chunks] == [w[0] for w in want], trial
        for (cid, _t, mj), (_wid, _wt, wj) in zip(chunks, want):
            assert json.loads(mj) == json.loads(wj), cid
        checked += 1
    assert checked >= 4


def test_host_stage_refuses_to_run_without_device_block_sums(golden_dir):
    """The product path has no host fallback for the block means: a DevicePass without its device callable raises."""
    g = np.load(os.path.join(golden_dir, "grouping.npz"))
    meta = json.loads(str(g["meta_json"]))
    dp = _device_pass_from_golden(g, meta, "a")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.cluster_from_device_pass(dp, W_override=g["a_W_all"])


def test_block_sums_oracle_equals_the_reference_loops():
    """oracle.block_sums_ref against the reference's element-by-element `_mean_between` / `_mean_within` definitions
    (Grouping:118-130), including a repeated member."""
    rng = np.random.default_rng(3)
    A = rng.random((20, 20)).astype(np.float32)
    sharp = np.maximum(A, A.T)
    np.fill_diagonal(sharp, 0.0)
    groups = [[0, 3, 5], [1, 2, 2, 7], [9], []]
    rowsum, block = go.block_sums_ref(sharp, groups)
    gm = G._GroupMeans(lambda gs: go.block_sums_ref(sharp, gs), groups)
    for a, ga in enumerate(groups):
        for b, gb in enumerate(groups):
            want = float(np.mean([float(sharp[i, j]) for i in ga for j in gb])) if ga and gb else 0.0
            assert abs(gm.between([a], [b]) - want) < 1e-12
        vals = [float(sharp[ga[i], ga[j]]) for i in range(len(ga)) for j in range(i + 1, len(ga))]
        want_w = float(np.mean(vals)) if len(ga) > 1 and vals else 1.0
        assert abs(gm.within([a]) - want_w) < 1e-12
    u = sorted(groups[0] + groups[1])
    vals = [float(sharp[u[i], u[j]]) for i in range(len(u)) for j in range(i + 1, len(u))]
    assert abs(gm.within([0, 1]) - float(np.mean(vals))) < 1e-12
    assert rowsum.shape == (20, 4) and block.shape == (4, 4)
