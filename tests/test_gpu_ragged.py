"""GPU parity tests for the ragged-document kernels: K3 segmented similarity matrices, K4 grouping
threshold pass, K5 splitter passes.  All calls go through the C ABI.

Two kinds of checks:
* floating-point outputs (S, sim_sharp, centrality, adjacent similarities) against the numpy
  oracle and the committed reference fixtures, tolerance 1e-5 abs (BASELINE.json north_star);
* selection / order-statistic logic (radix-select quantiles, kNN lists, percentile thresholds,
  breakpoint flags, medians) bit-exactly against numpy applied to the kernel's own fp32 outputs,
  so that no tolerance hides an off-by-one in an index.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import grouping_oracle as go
from oracle import simmatrix_oracle as so
from oracle import splitter_oracle as spo

pytestmark = pytest.mark.gpu
TOL = 1e-5


def topic_docs(rng, sizes, d, noise=0.7):
    rows = []
    for n in sizes:
        n_topics = max(1, int(np.ceil(n / 12)))
        cent = rng.standard_normal((n_topics, d)).astype(np.float32)
        topic = np.minimum(np.arange(n) // 12, n_topics - 1)
        rows.append((cent[topic] + noise * rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32))
    return rows


def run_sim(rows_list):
    from semanticsearch_b200 import ragged
    sizes = [r.shape[0] for r in rows_list]
    E = np.concatenate([r for r in rows_list if r.shape[0] > 0], axis=0) if sum(sizes) else np.zeros((0, 1), np.float32)
    plan = ragged.make_plan(sizes, "cuda")
    Et = torch.from_numpy(E).cuda()
    S = ragged.segmented_simmatrix(Et, plan)
    torch.cuda.synchronize()
    S_host = S.cpu().numpy()
    blocks = [S_host[plan.s_offsets[i]:plan.s_offsets[i + 1]].reshape(sizes[i], sizes[i]) for i in range(len(sizes))]
    return plan, Et, S, blocks


def test_simmatrix_golden_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "grouping.npz"))
    for name in ("a", "b", "c", "tiny", "seven"):
        E = g[f"{name}_E"]
        _, _, _, blocks = run_sim([E])
        np.testing.assert_allclose(blocks[0], g[f"{name}_S"], atol=TOL, rtol=0)
        np.testing.assert_array_equal(blocks[0], blocks[0].T)  # exactly symmetric
    S = run_sim([g["a_E"]])[3][0]
    assert np.all(S[5] == 0) and np.all(S[:, 5] == 0)  # zero sentence vector -> zero row/column


@pytest.mark.parametrize("d", [768, 384, 50, 3])
def test_simmatrix_ragged_batch_vs_oracle(d):
    rng = np.random.default_rng(3 + d)
    sizes = [int(x) for x in rng.integers(16, 513, size=24)] + [1, 2, 0, 64, 65, 128, 63]
    rows = topic_docs(rng, sizes, d)
    _, _, _, blocks = run_sim(rows)
    for E, S in zip(rows, blocks):
        if E.shape[0] >= 2:
            np.testing.assert_allclose(S, so.similarity_matrix_ref(E), atol=TOL, rtol=0)
        elif E.shape[0] == 1:
            assert abs(S[0, 0] - 1.0) < TOL


def test_group_pass_golden_fixture(golden_dir):
    from semanticsearch_b200 import ragged
    g = np.load(os.path.join(golden_dir, "grouping.npz"))
    meta = json.loads(str(g["meta_json"]))
    for name in ("a", "b", "c", "tiny", "seven"):
        plan, _, S, blocks = run_sim([g[f"{name}_E"]])
        out = ragged.group_threshold_pass(S, plan)
        torch.cuda.synchronize()
        n = blocks[0].shape[0]
        sharp = out["sim_sharp"].cpu().numpy().reshape(n, n)
        sc = meta[f"{name}_scalars"]
        np.testing.assert_allclose(sharp, g[f"{name}_sim_sharp"], atol=TOL, rtol=0)
        np.testing.assert_allclose(out["centrality"].cpu().numpy(), g[f"{name}_centrality"], atol=TOL, rtol=0)
        st = out["doc_stats"].cpu().numpy()[0]
        assert st[0] == pytest.approx(sc["mu"], abs=1e-6) and st[1] == pytest.approx(sc["sigma"], abs=1e-6)
        assert st[2] == pytest.approx(sc["eff_edge_floor"], abs=TOL)
        if sc["eff_tau_merge"] is not None:
            assert st[3] == pytest.approx(sc["eff_tau_merge"], abs=TOL)
        if sc["global_merge_thr"] is not None:
            assert st[4] == pytest.approx(sc["global_merge_thr"], abs=TOL)
        if sc["eff_reassign_delta"] is not None:
            assert st[5] == pytest.approx(sc["eff_reassign_delta"], abs=TOL)
        assert int(st[7]) == sc["k_eff_all"]
        W = ragged.knn_graph_from_lists(out["knn_idx"].cpu().numpy(), out["knn_val"].cpu().numpy(), st[2])
        W_ref = g[f"{name}_W_all"]
        # same graph, except edges whose value ties with a row's cut-off or sits at the floor
        cut = np.sort(sharp, axis=1)[:, ::-1][:, min(sc["k_eff_all"], n - 1)]
        for i, j in zip(*np.nonzero((W != 0) != (W_ref != 0))):
            v = sharp[i, j]
            assert min(abs(v - cut[i]), abs(v - cut[j]), abs(v - st[2])) <= 2e-5, (name, i, j, v)
        both = (W != 0) & (W_ref != 0)
        np.testing.assert_allclose(W[both], W_ref[both], atol=TOL, rtol=0)


@pytest.mark.parametrize("symmetric", [False, True])
def test_group_pass_selection_logic_is_exact(symmetric):
    """Quantiles and neighbour lists must equal numpy's on the kernel's own sim_sharp, bit for bit — with and without the
    promise that S is symmetric (K3's output is: the kernel then counts the strict upper triangle only)."""
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(17)
    sizes = [16, 17, 33, 100, 257, 512, 2, 3, 64, 300, 513, 700, 129, 384, 385, 40]  # > 512 rows: the generic (memory-resident) row path
    rows = topic_docs(rng, sizes, 96)
    rows[3][7] = 0.0
    plan, _, S, blocks = run_sim(rows)
    out = ragged.group_threshold_pass(S, plan, symmetric=symmetric)
    torch.cuda.synchronize()
    sharp_all = out["sim_sharp"].cpu().numpy()
    stats = out["doc_stats"].cpu().numpy()
    kidx = out["knn_idx"].cpu().numpy()
    kval = out["knn_val"].cpu().numpy()
    cent = out["centrality"].cpu().numpy()
    for d, n in enumerate(sizes):
        sharp = sharp_all[plan.s_offsets[d]:plan.s_offsets[d + 1]].reshape(n, n)
        assert np.all(np.diag(sharp) == 0)
        np.testing.assert_allclose(sharp, go.sharpen_ref(blocks[d])[0], atol=TOL, rtol=0)
        thr = go.thresholds_ref(sharp)
        assert stats[d, 6] == thr["count"]
        assert stats[d, 2] == thr["edge_floor"], (d, n)
        assert stats[d, 3] == thr["tau_merge"]
        assert stats[d, 4] == thr["global_merge_thr"]
        assert stats[d, 5] == pytest.approx(thr["reassign_delta"], rel=1e-9, abs=1e-12)
        assert int(stats[d, 7]) == go.k_eff_auto(n)
        idx_ref, val_ref = go.knn_lists_ref(sharp, go.k_eff_auto(n))
        w = idx_ref.shape[1]
        rows_d = slice(plan.offsets[d], plan.offsets[d + 1])
        np.testing.assert_array_equal(kidx[rows_d, :w], idx_ref)
        np.testing.assert_array_equal(kval[rows_d, :w], val_ref)
        assert np.all(kidx[rows_d, w:] == -1)
        np.testing.assert_allclose(cent[rows_d], (sharp.sum(axis=1) / max(n - 1, 1)).astype(float), atol=1e-6, rtol=0)
        W = ragged.knn_graph_from_lists(kidx[rows_d], kval[rows_d], stats[d, 2])
        np.testing.assert_array_equal(W, go.knn_graph_ref(sharp, go.k_eff_auto(n), stats[d, 2]))


def run_split(rows_list, dtype=torch.float32, pct=95.0):
    from semanticsearch_b200 import ragged
    sizes = [r.shape[0] for r in rows_list]
    E = np.concatenate(rows_list, axis=0)
    plan = ragged.make_plan(sizes, "cuda")
    Et = torch.from_numpy(E).cuda().to(dtype)
    adj = ragged.adjacent_cosine(Et)
    thr, flags, stats, smooth = ragged.segmented_percentile(adj, plan, pct)
    torch.cuda.synchronize()
    return plan, Et.float().cpu().numpy(), adj.cpu().numpy(), thr.cpu().numpy(), flags.cpu().numpy(), stats.cpu().numpy(), smooth.cpu().numpy()


def test_splitter_golden_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "splitter.npz"))
    for name in ("s40", "s97", "s12", "s3"):
        E = g[f"{name}_E"]
        plan, _, adj, thr, flags, stats, smooth = run_split([E])
        m = E.shape[0] - 1
        np.testing.assert_allclose(adj[:m], g[f"{name}_adj_sims"], atol=TOL, rtol=0)
        np.testing.assert_allclose(smooth[:m], g[f"{name}_adj_base"], atol=TOL, rtol=0)
        ref = spo.robust_stats_ref(g[f"{name}_adj_sims"], 3)
        assert stats[0, 0] == pytest.approx(ref["median"], abs=TOL)
        assert stats[0, 1] == pytest.approx(ref["mad"], abs=TOL)
        assert stats[0, 3] - stats[0, 2] == pytest.approx(ref["iqr"], abs=2 * TOL)
        t_ref, bp_ref = spo.p95_breakpoints_ref(g[f"{name}_adj_sims"])
        assert thr[0] == pytest.approx(t_ref, abs=TOL)


@pytest.mark.parametrize("d,dtype", [(384, torch.float32), (768, torch.float32), (384, torch.bfloat16), (50, torch.float32), (6, torch.float16)])
def test_splitter_ragged_batch(d, dtype):
    rng = np.random.default_rng(5 + d)
    sizes = [int(x) for x in rng.integers(16, 513, size=40)] + [1, 2, 3, 4, 1, 700]
    rows = topic_docs(rng, sizes, d, noise=0.6)
    rows[2][3] = 0.0
    plan, Ef, adj, thr, flags, stats, smooth = run_split(rows, dtype)
    tol = TOL if dtype == torch.float32 else 2e-3
    for di, n in enumerate(sizes):
        a, b = plan.offsets[di], plan.offsets[di + 1]
        m = n - 1
        assert adj[b - 1] == 0.0
        if m < 1:
            assert np.isnan(thr[di])
            continue
        ref_adj = spo.adjacent_sims_ref(Ef[a:b])
        np.testing.assert_allclose(adj[a:a + m], ref_adj, atol=tol if dtype == torch.float32 else 1e-5, rtol=0)
        # order-statistic logic: exact on the kernel's own adjacent similarities
        mine = adj[a:a + m].astype(float)
        t_ref, bp_ref = spo.p95_breakpoints_ref(mine)
        assert thr[di] == t_ref, (di, n)
        np.testing.assert_array_equal(np.nonzero(flags[a:b])[0], bp_ref)
        st = spo.robust_stats_ref(mine, 3)
        np.testing.assert_array_equal(smooth[a:a + m].astype(float), st["adj_base"])
        assert stats[di, 0] == st["median"] and stats[di, 1] == st["mad"]
        assert stats[di, 2] == float(np.percentile(st["adj_base"], 25)) and stats[di, 3] == float(np.percentile(st["adj_base"], 75))


def test_splitter_other_percentiles_and_large_batch():
    rng = np.random.default_rng(99)
    sizes = [int(x) for x in rng.integers(2, 200, size=3000)]
    rows = topic_docs(rng, sizes, 64, noise=0.5)
    for pct in (95.0, 50.0, 0.0, 100.0, 12.5):
        plan, Ef, adj, thr, flags, stats, smooth = run_split(rows, pct=pct)
        for di in rng.integers(0, len(sizes), size=60):
            a, b = plan.offsets[di], plan.offsets[di + 1]
            mine = adj[a:b - 1].astype(float)
            t_ref, bp_ref = spo.p95_breakpoints_ref(mine, pct)
            assert thr[di] == t_ref
            np.testing.assert_array_equal(np.nonzero(flags[a:b])[0], bp_ref)


def test_group_pass_crowded_bin_takes_radix_fallback():
    """Thousands of identical similarities land in one histogram bin (more than the shared-memory
    candidate list holds): the radix-select fallback must still return numpy's quantiles exactly."""
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(23)
    sizes = [100, 80, 300]
    mats = []
    S0 = np.full((100, 100), 0.3, dtype=np.float32)
    np.fill_diagonal(S0, 1.0)
    mats.append(S0)                                   # 9900 identical off-diagonal values
    S1 = np.where(rng.random((80, 80)) < 0.5, 0.2, 0.7).astype(np.float32)
    S1 = np.maximum(S1, S1.T)
    np.fill_diagonal(S1, 1.0)
    mats.append(S1)                                   # two heavy clusters of duplicates
    E = topic_docs(rng, [300], 64)[0]
    mats.append(so.similarity_matrix_ref(E).astype(np.float32))  # ordinary document in the same batch
    plan = ragged.make_plan(sizes, "cuda")
    S = torch.from_numpy(np.concatenate([m.reshape(-1) for m in mats])).cuda()
    out = ragged.group_threshold_pass(S, plan)
    torch.cuda.synchronize()
    sharp_all = out["sim_sharp"].cpu().numpy()
    stats = out["doc_stats"].cpu().numpy()
    kidx = out["knn_idx"].cpu().numpy()
    kval = out["knn_val"].cpu().numpy()
    for d, n in enumerate(sizes):
        sharp = sharp_all[plan.s_offsets[d]:plan.s_offsets[d + 1]].reshape(n, n)
        thr = go.thresholds_ref(sharp)
        assert stats[d, 6] == thr["count"]
        assert stats[d, 2] == thr["edge_floor"], (d, n)
        assert stats[d, 3] == thr["tau_merge"]
        assert stats[d, 4] == thr["global_merge_thr"]
        idx_ref, val_ref = go.knn_lists_ref(sharp, go.k_eff_auto(n))
        w = idx_ref.shape[1]
        rows_d = slice(plan.offsets[d], plan.offsets[d + 1])
        np.testing.assert_array_equal(kidx[rows_d, :w], idx_ref)   # ties -> ascending column index
        np.testing.assert_array_equal(kval[rows_d, :w], val_ref)


def test_splitter_passes_accept_output_buffers():
    """adjacent_cosine / segmented_percentile into caller-owned buffers give the same bits as fresh allocations."""
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(41)
    sizes = [2, 9, 33, 130, 1, 64]
    E = torch.from_numpy(np.concatenate(topic_docs(rng, sizes, 48))).cuda()
    plan = ragged.make_plan(sizes, "cuda")
    adj0 = ragged.adjacent_cosine(E)
    ref = ragged.segmented_percentile(adj0.clone(), plan, 95.0, want_stats=True)
    adj1 = torch.full_like(adj0, 7.0)
    assert ragged.adjacent_cosine(E, out=adj1) is adj1
    bufs = tuple(torch.full_like(t, 3) for t in ref)
    got = ragged.segmented_percentile(adj1, plan, 95.0, want_stats=True, out=bufs)
    for a, b, c in zip(got, ref, bufs):
        assert a is c
        assert torch.equal(torch.nan_to_num(a.double(), nan=-1.0), torch.nan_to_num(b.double(), nan=-1.0))
    with pytest.raises(ValueError):
        ragged.adjacent_cosine(E, out=torch.empty(3, device="cuda"))
    with pytest.raises(ValueError):
        ragged.segmented_percentile(adj1, plan, 95.0, out=(bufs[0][:1], bufs[1], bufs[2], bufs[3]))


def test_group_pass_symmetric_promise_changes_nothing():
    """With a bit-symmetric S (K3's output) the upper-triangle form must return the same bits as the full form: sim_sharp,
    centrality, neighbour lists, mu / sigma / quantiles / count; 0.1 * std only differs by float64 summation order."""
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(23)
    sizes = [int(x) for x in rng.integers(2, 513, size=60)] + [33, 128, 129, 256, 257, 384, 385, 512]
    rows = topic_docs(rng, sizes, 64)
    plan, _, S, _ = run_sim(rows)
    a = ragged.group_threshold_pass(S, plan, symmetric=False)
    b = ragged.group_threshold_pass(S, plan, symmetric=True)
    torch.cuda.synchronize()
    for key in ("sim_sharp", "centrality", "knn_idx", "knn_val"):
        assert torch.equal(a[key], b[key]), key
    sa, sb = a["doc_stats"].cpu().numpy(), b["doc_stats"].cpu().numpy()
    for col in (0, 1, 2, 3, 4, 6, 7):
        assert np.array_equal(sa[:, col], sb[:, col]), col
    np.testing.assert_allclose(sa[:, 5], sb[:, 5], rtol=1e-12, atol=1e-15)


def test_group_pass_heavy_ties_with_symmetric_promise():
    """Thousands of identical similarities (the radix-select fallback) and two heavy clusters of duplicates, upper-triangle form."""
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(29)
    S0 = np.full((100, 100), 0.3, dtype=np.float32)
    np.fill_diagonal(S0, 1.0)
    S1 = np.where(rng.random((80, 80)) < 0.5, 0.2, 0.7).astype(np.float32)
    S1 = np.maximum(S1, S1.T)
    np.fill_diagonal(S1, 1.0)
    S2 = np.where(rng.random((400, 400)) < 0.3, 0.1, 0.9).astype(np.float32)
    S2 = np.maximum(S2, S2.T)
    np.fill_diagonal(S2, 1.0)
    mats = [S0, S1, S2]
    sizes = [m.shape[0] for m in mats]
    plan = ragged.make_plan(sizes, "cuda")
    S = torch.from_numpy(np.concatenate([m.reshape(-1) for m in mats])).cuda()
    out = ragged.group_threshold_pass(S, plan, symmetric=True)
    torch.cuda.synchronize()
    sharp_all = out["sim_sharp"].cpu().numpy()
    stats = out["doc_stats"].cpu().numpy()
    kidx, kval = out["knn_idx"].cpu().numpy(), out["knn_val"].cpu().numpy()
    for d, n in enumerate(sizes):
        sharp = sharp_all[plan.s_offsets[d]:plan.s_offsets[d + 1]].reshape(n, n)
        thr = go.thresholds_ref(sharp)
        assert stats[d, 6] == thr["count"]
        assert stats[d, 2] == thr["edge_floor"] and stats[d, 3] == thr["tau_merge"] and stats[d, 4] == thr["global_merge_thr"]
        idx_ref, val_ref = go.knn_lists_ref(sharp, go.k_eff_auto(n))
        w = idx_ref.shape[1]
        rows_d = slice(plan.offsets[d], plan.offsets[d + 1])
        np.testing.assert_array_equal(kidx[rows_d, :w], idx_ref)
        np.testing.assert_array_equal(kval[rows_d, :w], val_ref)


def test_grouping_pass_host_pipeline_matches_the_single_pass():
    """The host-buffer operator runs long batches as a pipeline of document chunks (H2D / kernels / D2H overlapped on two
    copy streams): every output must equal the one-chunk pass bit for bit, including empty documents, chunk boundaries
    that fall next to them, a reused plan and reused pinned output buffers."""
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(21)
    sizes = [int(x) for x in rng.integers(1, 300, size=90)] + [0, 0, 513, 2, 0, 64, 65]
    E = torch.from_numpy(rng.standard_normal((sum(sizes), 96)).astype(np.float32)).pin_memory()
    one = ragged.grouping_pass_host(E, sizes, chunk_bytes=1 << 40)
    one = {k: v.clone() for k, v in one.items()}
    plan = ragged.make_plan(sizes, "cuda")
    out = None
    for _ in range(2):                                   # second round: cached chunk plans, reused pinned outputs
        out = ragged.grouping_pass_host(E, sizes, out=out, plan=plan, chunk_bytes=1 << 20)
        assert len(plan.host_chunks[1]) > 5
        assert set(out) == set(one)
        for k in one:
            assert out[k].shape == one[k].shape and out[k].dtype == one[k].dtype, k
            assert torch.equal(out[k].view(torch.uint8), one[k].view(torch.uint8)), k
