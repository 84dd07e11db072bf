"""GPU tests of the drop-in modules (semanticsearch_b200.Method.*) against outputs captured from the
unmodified reference (tests/golden/*.npz): same call signatures, same chunk / cluster boundaries."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import simmatrix_oracle as so
from oracle import splitter_oracle as spo

pytestmark = pytest.mark.gpu
DELIM = " || "


@pytest.fixture()
def fake_text_stack():
    """Synthetic sentences backed by a lookup-table encoder, as in oracle/ref_shim.py."""
    from semanticsearch_b200.Tool import Sentence_Embedding as emb
    from semanticsearch_b200.Tool import Sentence_Segmenter as seg
    table = {}

    def backend(text_list, model_name, batch_size=32, device_preference=None):
        return np.stack([table[t] for t in text_list]).astype(np.float32)

    def make_doc(E, tag):
        sents = [f"{tag}s{i:05d}x" for i in range(len(E))]
        for s, v in zip(sents, E):
            table[s] = np.asarray(v, np.float32)
        return DELIM.join(sents), sents

    emb.set_embedding_backend(backend)
    seg.set_sentence_splitter(lambda t: [s.strip() for s in t.split(DELIM.strip()) if s.strip()])
    yield make_doc
    emb.set_embedding_backend(None)
    seg.set_sentence_splitter(None)


def test_create_similarity_matrix_and_distribution(golden_dir, fake_text_stack):
    from semanticsearch_b200.Method import semantic_common as sc
    g = np.load(os.path.join(golden_dir, "grouping.npz"))
    meta = json.loads(str(g["meta_json"]))
    for name in ("a", "b", "c", "tiny", "seven"):
        _text, sents = fake_text_stack(g[f"{name}_E"], name)
        S = sc.create_similarity_matrix(sents, "m", batch_size=64, device="cuda", silent=True)
        assert S.dtype == np.float32 and S.shape == g[f"{name}_S"].shape
        np.testing.assert_allclose(S, g[f"{name}_S"], atol=1e-5, rtol=0)
        # statistics of the reference's own S: order statistics exact, moments to fp32 rounding
        got = sc.analyze_similarity_distribution(g[f"{name}_S"])
        want = meta[f"{name}_dist"]
        assert set(got) == set(want)
        for key, val in want.items():
            tol = 1e-6 if key in ("mean", "std") else 0.0
            assert got[key] == pytest.approx(val, abs=tol), (name, key)
    assert sc.create_similarity_matrix(["one"], "m") is None
    assert sc.analyze_similarity_distribution(np.ones((1, 1), np.float32)) is None
    # everything >= 1 - 1e-5 filtered -> every key reports max(sims)
    got = sc.analyze_similarity_distribution(np.ones((4, 4), np.float32))
    assert set(got.values()) == {1.0}


def test_similarity_distribution_random_matches_numpy_exactly():
    from semanticsearch_b200.Method import semantic_common as sc
    rng = np.random.default_rng(3)
    for n in (2, 3, 17, 130, 400):
        E = rng.standard_normal((n, 48)).astype(np.float32)
        S = so.similarity_matrix_ref(E)
        got = sc.analyze_similarity_distribution(S)
        want = so.analyze_similarity_distribution_ref(S)
        for key, val in want.items():
            tol = 1e-6 if key in ("mean", "std") else 0.0
            assert got[key] == pytest.approx(val, abs=tol), (n, key)


def test_grouping_dropin_reproduces_reference_chunks(golden_dir, fake_text_stack):
    from semanticsearch_b200.Method import Semantic_Grouping_Optimized as G
    g = np.load(os.path.join(golden_dir, "grouping.npz"))
    meta = json.loads(str(g["meta_json"]))
    for name in ("a", "b", "c", "tiny", "seven"):
        text, _ = fake_text_stack(g[f"{name}_E"], name)
        out = G.semantic_chunk_passage_from_grouping_logic(f"doc_{name}", text, "m", device="cuda", silent=True,
                                                           collect_metadata=True, output_dir=None)
        want = meta[f"{name}_chunks"]
        assert [c[0] for c in out] == [w[0] for w in want], name
        for (cid, _t, mj), (_wid, wj) in zip(out, want):
            mine, ref = json.loads(mj), json.loads(wj)
            assert mine["sent_indices"] == ref["sent_indices"] and mine["n"] == ref["n"], cid
            assert mine["method_used"] == ref["method_used"]
            for key in ("sim_mean", "sim_min", "sim_max", "sim_std", "exemplar_centrality"):
                if key in ref:
                    assert mine[key] == pytest.approx(ref[key], abs=2e-4)


def _strict_rank_by_sorting(S):
    """#{k: S[i,k] < S[i,j]} + #{k: S[k,j] < S[i,j]} via searchsorted (cheap at large n; checked against the oracle below)."""
    row = np.stack([np.searchsorted(np.sort(r), r, side="left") for r in S])
    col = np.stack([np.searchsorted(np.sort(c), c, side="left") for c in S.T]).T
    return (row + col).astype(np.float32)


def test_c99_rank_kernel_exact_on_own_similarity():
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(9)
    sizes = [12, 40, 97, 130, 33, 2, 300, 257, 512, 64]
    rows = [rng.standard_normal((n, 32)).astype(np.float32) for n in sizes]
    rows[2][5] = rows[2][60]          # duplicated sentences: exact ties inside rows and columns
    rows[2][61] = rows[2][60]
    rows[6][100:110] = rows[6][0]
    rows[3][7] = 0.0                  # a zero embedding: a row / column of zeros
    plan = ragged.make_plan(sizes, "cuda")
    E = torch.from_numpy(np.concatenate(rows)).cuda()
    S = ragged.segmented_simmatrix(E, plan)
    Rg = ragged.c99_rank_matrix(S, plan, use_local_rank=False).cpu().numpy()
    Rl = ragged.c99_rank_matrix(S, plan, use_local_rank=True, mask_size=11).cpu().numpy()
    S_h = S.cpu().numpy()
    for d, n in enumerate(sizes):
        blk = slice(plan.s_offsets[d], plan.s_offsets[d + 1])
        Sd = S_h[blk].reshape(n, n)
        want = spo.c99_global_rank_ref(Sd)
        np.testing.assert_array_equal(_strict_rank_by_sorting(Sd), want)
        np.testing.assert_array_equal(Rg[blk].reshape(n, n), want)
        if n <= 130:
            np.testing.assert_array_equal(Rl[blk].reshape(n, n), spo.c99_local_rank_ref(Sd, 11))
    # both similarity kernels mirror their tiles, so S is symmetric bit for bit and the transposed-row-rank shortcut applies
    for algo in ("tc", "ffma"):
        Sa = ragged.segmented_simmatrix(E, plan, algo=algo)
        Sa_h = Sa.cpu().numpy()
        Rs = ragged.c99_rank_matrix(Sa, plan, symmetric=True).cpu().numpy()
        Rn = ragged.c99_rank_matrix(Sa, plan).cpu().numpy()
        np.testing.assert_array_equal(Rs, Rn)
        for d, n in enumerate(sizes):
            Sd = Sa_h[plan.s_offsets[d]:plan.s_offsets[d + 1]].reshape(n, n)
            np.testing.assert_array_equal(Sd, Sd.T)


@pytest.mark.parametrize("sizes", [[100, 90, 128, 5], [1100, 40, 513], [2048, 7], [2300, 64]])
def test_c99_global_rank_all_size_classes(sizes):
    """Every kernel variant (<= 128, <= 512, <= 2048 sentences: sorting; longer: counting) on non-symmetric input."""
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(sum(sizes))
    plan = ragged.make_plan(sizes, "cuda")
    blocks = [np.round(rng.standard_normal((n, n)), 2).astype(np.float32) for n in sizes]  # rounded: many ties, not symmetric
    S = torch.from_numpy(np.concatenate([b.ravel() for b in blocks])).cuda()
    R = ragged.c99_rank_matrix(S, plan, use_local_rank=False).cpu().numpy()
    for d, n in enumerate(sizes):
        got = R[plan.s_offsets[d]:plan.s_offsets[d + 1]].reshape(n, n)
        np.testing.assert_array_equal(got, _strict_rank_by_sorting(blocks[d]))


def test_splitter_dropin_reproduces_reference_groups(golden_dir, fake_text_stack):
    from semanticsearch_b200.Method import Semantic_Splitter_Optimized as SP
    g = np.load(os.path.join(golden_dir, "splitter.npz"))
    meta = json.loads(str(g["meta_json"]))
    for name in ("s40", "s97", "s12", "s3"):
        E = g[f"{name}_E"]
        text, sents = fake_text_stack(E, name)
        m = meta[name]
        min_chunk = max(3, m["min_boundary_spacing"])
        c99 = SP._c99_boundaries(g[f"{name}_En"], min_chunk_size=min_chunk,
                                 use_local_rank=bool(m["kwargs"].get("c99_use_local_rank", False)))
        assert c99 == m["c99_bounds"], name
        valley = SP._valley_boundaries(list(g[f"{name}_adj_for_valley"]), triplet_tau=m["valley_tau"],
                                       min_boundary_spacing=m["min_boundary_spacing"],
                                       min_first_boundary_index=m["min_first_boundary_index"])
        assert valley == m["valley_bounds"], name
        chunks, sentences, groups = SP.process_sentence_splitting_with_semantics(
            text, embedding_model="m", device="cuda", silent=True, **m["kwargs"])
        assert sentences == sents
        assert [[grp[0], grp[-1]] for grp in groups] == m["groups"], name
        out = SP.chunk_passage_text_splitter(f"doc_{name}", text, "m", device="cuda", silent=True, collect_metadata=True,
                                             **m["kwargs"])
        assert [c[0] for c in out] == [f"doc_{name}_chunk{i}" for i in range(len(m["groups"]))]
        for (cid, _t, mj), grp in zip(out, m["groups"]):
            assert json.loads(mj)["n"] == grp[1] - grp[0] + 1


def test_similarity_matrices_accept_device_embeddings():
    """Embedding hand-off: CUDA tensors in, same matrices out as for numpy input."""
    import torch
    from semanticsearch_b200.Method import semantic_common as sc
    rng = np.random.default_rng(9)
    docs = [rng.standard_normal((n, 384)).astype(np.float32) for n in (5, 130, 1, 40)]
    host = sc.similarity_matrices_from_embeddings(docs)
    dev = sc.similarity_matrices_from_embeddings([torch.from_numpy(d).cuda() for d in docs])
    assert host[2] is None and dev[2] is None
    for a, b in zip(host, dev):
        if a is not None:
            np.testing.assert_array_equal(a, b)


def test_cuda_tensor_handoff_equals_host_arrays():
    """Encoder outputs that already live in HBM feed the grouping, splitter and similarity passes without a host round trip
     and give the same bits as the numpy call."""
    from semanticsearch_b200.Method import Semantic_Grouping_Optimized as G
    from semanticsearch_b200.Method import Semantic_Splitter_Optimized as SP
    from semanticsearch_b200.Method import semantic_common as sc
    rng = np.random.default_rng(77)
    docs = [rng.standard_normal((n, 64)).astype(np.float32) for n in (9, 40, 130, 2)]
    dev = [torch.from_numpy(e).cuda() for e in docs]
    for a, b in zip(sc.similarity_matrices_from_embeddings(docs), sc.similarity_matrices_from_embeddings(dev)):
        assert np.array_equal(a, b)
    for a, b in zip(G.device_pass_batch(docs), G.device_pass_batch(dev)):
        assert np.array_equal(a.sim_sharp, b.sim_sharp) and np.array_equal(a.knn_idx, b.knn_idx) and a.q80 == b.q80
    for a, b in zip(SP.splitter_device_pass(docs), SP.splitter_device_pass(dev)):
        assert a["adj_sims"] == b["adj_sims"] and a["p95_threshold"] == b["p95_threshold"]
    assert SP.c99_boundaries_batch(docs, 3) == SP.c99_boundaries_batch(dev, 3)
