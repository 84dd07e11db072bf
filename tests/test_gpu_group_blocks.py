"""GPU tests of the block sums behind the grouping stage's `_mean_between` / `_mean_within` / reassignment means
(reference Method/Semantic_Grouping_Optimized.py:118-130,566-588) and of the co-association matrix (:231-241),
against the numpy oracle, plus the drop-in's clusters against the reference-captured goldens."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import grouping_oracle as go

pytestmark = pytest.mark.gpu


def _sym_sharp(rng, n):
    a = rng.random((n, n)).astype(np.float32)
    s = np.maximum(a, a.T)
    np.fill_diagonal(s, 0.0)
    return s


def test_block_sums_batch_matches_oracle():
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(11)
    sizes = [2, 7, 33, 130, 64, 1, 257]
    mats = [_sym_sharp(rng, n) for n in sizes]
    plan = ragged.make_plan(sizes, "cuda")
    sharp = torch.from_numpy(np.concatenate([m.ravel() for m in mats])).cuda()
    groups = []
    for n in sizes:
        perm = rng.permutation(n)
        cuts = sorted(set(rng.integers(0, n + 1, size=min(5, n)).tolist()) | {0, n})
        gs = [sorted(perm[a:b].tolist()) for a, b in zip(cuts[:-1], cuts[1:])]
        if n > 3:
            gs.append(sorted(gs[0] + gs[-1][:2]))       # overlapping group (the reference's merged-twice quirk)
            gs.append([int(perm[0]), int(perm[0]), int(perm[1])])   # a repeated member
            gs.append([])                                # empty group
        groups.append(gs)
    got = ragged.group_block_sums(sharp, plan, groups)
    for m, gs, (rowsum, block) in zip(mats, groups, got):
        want_r, want_b = go.block_sums_ref(m, gs)
        np.testing.assert_allclose(rowsum, want_r, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(block, want_b, rtol=1e-13, atol=1e-13)
    # reproducible bit for bit
    again = ragged.group_block_sums(sharp, plan, groups)
    for (r0, b0), (r1, b1) in zip(got, again):
        assert np.array_equal(r0, r1) and np.array_equal(b0, b1)


def test_doc_block_sums_single_document():
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(12)
    m = _sym_sharp(rng, 100)
    f = ragged.DocBlockSums(torch.from_numpy(m.ravel()).cuda(), 100)
    groups = [list(range(0, 40)), list(range(40, 41)), list(range(41, 100)), [5, 5, 50]]
    rowsum, block = f(groups)
    want_r, want_b = go.block_sums_ref(m, groups)
    np.testing.assert_allclose(rowsum, want_r, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(block, want_b, rtol=1e-13, atol=1e-13)
    r0, b0 = f([])
    assert r0.shape == (100, 0) and b0.shape == (0, 0)


def test_coassociation_matches_numpy():
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(13)
    for n, L in ((5, 1), (64, 7), (300, 4)):
        labels = rng.integers(0, 6, size=(L, n))
        C = ragged.group_coassociation(labels).cpu().numpy()
        want = np.zeros((n, n))
        for lab in labels:
            want += (lab[:, None] == lab[None, :]).astype(float)
        np.fill_diagonal(want, 0.0)
        want /= float(L)
        assert np.array_equal(C, want)


def test_device_block_sums_reproduce_reference_clusters(golden_dir):
    """The whole drop-in host stage on top of the DEVICE block sums (sim_sharp of the reference uploaded as is)
    reproduces the reference's clusters and metadata of every golden document."""
    from semanticsearch_b200 import ragged
    from semanticsearch_b200.Method import Semantic_Grouping_Optimized as G
    from test_host_grouping import _device_pass_from_golden
    g = np.load(os.path.join(golden_dir, "grouping.npz"))
    meta = json.loads(str(g["meta_json"]))
    names = sorted({k[:-len("_sim_sharp")] for k in g.files if k.endswith("_sim_sharp")})
    assert len(names) >= 5
    for name in names:
        dp = _device_pass_from_golden(g, meta, name)
        n = dp.sim_sharp.shape[0]
        dp.block_sums = ragged.DocBlockSums(torch.from_numpy(np.ascontiguousarray(dp.sim_sharp).ravel()).cuda(), n)
        merged, method, _ = G.cluster_from_device_pass(dp, W_override=g[f"{name}_W_all"])
        assert method == meta[f"{name}_scalars"]["method_used"]
        chunks = G._emit(f"doc_{name}", "whole", [f"s{i}" for i in range(n)], merged, method, dp, collect_metadata=True)
        want = meta[f"{name}_chunks"]
        assert [c[0] for c in chunks] == [w[0] for w in want], name
        for (cid, _t, mj), (_wid, wj) in zip(chunks, want):
            assert json.loads(mj) == json.loads(wj), cid
