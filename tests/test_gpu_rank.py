"""GPU parity tests for the chunk-ranking path (K1 streaming cosine + fused top-k, K6 merge).

Every comparison goes through the C ABI (ctypes -> libsemsearch_b200.so) and is checked against
the numpy oracle / the committed reference fixtures.  Tolerances (BASELINE.json north_star):
scores within 1e-5 abs for fp32 inputs, 2e-3 abs for bf16/fp16 inputs; indices identical except
where fp64 cosines tie within 1e-5.
"""
import os

import numpy as np
import pytest
import torch

from oracle import rank_oracle as ro

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
LOWP_TOL = 2e-3


def _run(C, Q, k, dtype=torch.float32, **kw):
    from semanticsearch_b200 import similarity
    Ct = torch.from_numpy(C).cuda().to(dtype).contiguous()
    Qt = torch.from_numpy(Q).cuda().to(dtype).contiguous()
    s, i = similarity.cosine_topk(Ct, Qt, k, **kw)
    torch.cuda.synchronize()
    # the oracle sees exactly the rounded values the kernel saw, upcast to fp32
    return s.cpu().numpy(), i.cpu().numpy(), Ct.float().cpu().numpy(), Qt.float().cpu().numpy()


def test_library_loaded_and_device_is_blackwell():
    from semanticsearch_b200 import _lib
    import ctypes
    lib = _lib.load()
    sm, maj, mnr, smem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
    torch.zeros(1, device="cuda")
    assert lib.ss_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr), ctypes.byref(smem)) == 0
    assert maj.value == 10 and sm.value >= 100


def test_golden_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "rank_cosine.npz"))
    C, Q, ref_scores = g["C"], g["Q"], g["scores"]
    s, i, Cr, Qr = _run(C, Q, 10)
    assert ro.check_topk_against_oracle(Qr, Cr, s, i, 10, FP32_TOL) == []
    # against the reference's own cosine values (sklearn via rank_chunks_optimized.py:216)
    for b in range(Q.shape[0]):
        np.testing.assert_allclose(s[b], ref_scores[b][i[b]], atol=FP32_TOL, rtol=0)
        want = np.sort(ref_scores[b])[::-1][:10]
        np.testing.assert_allclose(s[b], want, atol=FP32_TOL, rtol=0)
    # duplicate rows 3 and 40 tie exactly for the collinear query: lower index first
    assert list(i[5][:2]) == [3, 40] or abs(s[5][0] - s[5][1]) > 0


def test_config1_100q_10k_384_fp32():
    rng1, rng2 = np.random.default_rng(1), np.random.default_rng(2)
    C = rng1.standard_normal((10000, 384)).astype(np.float32)
    Q = rng2.standard_normal((100, 384)).astype(np.float32)
    s, i, Cr, Qr = _run(C, Q, 10)
    assert ro.check_topk_against_oracle(Qr, Cr, s, i, 10, FP32_TOL) == []
    ref_s, ref_i = ro.cosine_topk_ref(Q, C, 10)
    assert (i == ref_i).mean() > 0.999


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, LOWP_TOL), (torch.float16, LOWP_TOL), (torch.float32, FP32_TOL)])
@pytest.mark.parametrize("n,d,b,k", [(50000, 768, 1, 10), (30011, 384, 3, 10), (20000, 384, 16, 100), (4097, 768, 5, 33)])
def test_random_shapes(dtype, tol, n, d, b, k):
    rng = np.random.default_rng(n + d + b + k)
    C = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((b, d)).astype(np.float32)
    s, i, Cr, Qr = _run(C, Q, k, dtype)
    assert ro.check_topk_against_oracle(Qr, Cr, s, i, k, tol) == []


@pytest.mark.parametrize("d,dtype", [(50, torch.float32), (100, torch.bfloat16), (3, torch.float16), (1, torch.float32)])
def test_generic_row_sizes(d, dtype):
    """Rows that are not 16-byte multiples take the plain-load kernel; same results."""
    rng = np.random.default_rng(d)
    C = rng.standard_normal((3000, d)).astype(np.float32)
    Q = rng.standard_normal((4, d)).astype(np.float32)
    k = 7
    s, i, Cr, Qr = _run(C, Q, k, dtype)
    assert ro.check_topk_against_oracle(Qr, Cr, s, i, k, LOWP_TOL if dtype != torch.float32 else FP32_TOL) == []


def test_edge_cases_zero_rows_ties_small_n():
    rng = np.random.default_rng(5)
    d = 64
    C = rng.standard_normal((40, d)).astype(np.float32)
    C[7] = 0.0
    C[20] = C[2]
    C[30] = C[2]
    Q = np.stack([C[2] * 3.0, np.zeros(d, dtype=np.float32)])
    s, i, Cr, Qr = _run(C, Q, 5)
    assert list(i[0][:3]) == [2, 20, 30]            # exact ties resolve to the lower index
    np.testing.assert_allclose(s[0][:3], 1.0, atol=FP32_TOL)
    assert np.all(s[1] == 0.0) and list(i[1]) == [0, 1, 2, 3, 4]  # zero query: all scores 0
    # n < k: trailing slots are empty (-inf / -1)
    s, i, _, _ = _run(C[:3], Q[:1], 5)
    assert list(i[0][3:]) == [-1, -1] and np.all(np.isneginf(s[0][3:]))
    assert sorted(i[0][:3].tolist()) == [0, 1, 2]
    # single row, single query
    s, i, _, _ = _run(C[:1], C[:1], 1)
    assert i[0, 0] == 0 and abs(s[0, 0] - 1.0) < FP32_TOL


def test_index_base_and_sharded_merge_equals_whole():
    from semanticsearch_b200 import similarity
    rng = np.random.default_rng(11)
    n, d, b, k = 60000, 384, 4, 10
    C = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32)).cuda().to(torch.bfloat16)
    Q = torch.from_numpy(rng.standard_normal((b, d)).astype(np.float32)).cuda().to(torch.bfloat16)
    C[40000] = C[123]  # a cross-shard exact tie
    s_all, i_all, k_all = similarity.cosine_topk(C, Q, k, return_keys=True)
    bounds = [0, 17000, 17001, 45000, n]
    parts = []
    for a, e in zip(bounds[:-1], bounds[1:]):
        _, i_p, k_p = similarity.cosine_topk(C[a:e].contiguous(), Q, k, index_base=a, return_keys=True)
        parts.append(k_p)
    s_m, i_m, k_m = similarity.topk_merge(torch.stack(parts).contiguous(), k)
    assert torch.equal(i_m, i_all) and torch.equal(s_m, s_all) and torch.equal(k_m, k_all)


def test_full_size_properties_2M_rows():
    """Size-independent checks at a size the oracle cannot finish quickly: planted winners are
    found at rank 0 with score ~1, results are sorted, and a torch fp32 matmul agrees."""
    from semanticsearch_b200 import similarity
    g = torch.Generator(device="cuda").manual_seed(6)
    n, d, k = 2_000_000, 768, 10
    C = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    Q = torch.randn((2, d), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    C[1_234_567] = Q[0] * 2
    C[1_999_999] = Q[1] * 0.5
    s, i = similarity.cosine_topk(C, Q, k)
    assert i[0, 0].item() == 1_234_567 and i[1, 0].item() == 1_999_999
    assert torch.all(s[:, 0] > 0.999)
    assert torch.all(s[:, 1:] <= s[:, :-1])
    Cf = C.float()
    ref = (torch.nn.functional.normalize(Q.float(), dim=1) @ torch.nn.functional.normalize(Cf, dim=1).T)
    rs, ri = torch.topk(ref, k, dim=1)
    assert torch.allclose(rs, s, atol=LOWP_TOL)
    assert (ri == i).float().mean().item() >= 0.9


def test_graphed_search_replays_the_eager_search():
    """sharded.GraphedSearch (one CUDA-graph launch per search) must return what ShardedCorpus.search returns,
    for every kernel family, across several replays with different queries."""
    from semanticsearch_b200.sharded import GraphedSearch, ShardedCorpus
    g = torch.Generator(device="cuda").manual_seed(5)
    C = torch.randn((60000, 384), generator=g, device="cuda").to(torch.bfloat16)
    corpus = ShardedCorpus(C, 1000)
    for batch, k in ((1, 10), (16, 100), (200, 10)):
        gs = GraphedSearch(corpus, batch, k)
        for _ in range(3):
            Q = torch.randn((batch, 384), generator=g, device="cuda").to(torch.bfloat16)
            s0, i0 = corpus.search(Q, k)
            s1, i1 = gs(Q)
            torch.cuda.synchronize()
            assert torch.equal(i0, i1) and torch.equal(s0, s1)
            assert int(i1.min()) >= 1000


def test_resident_corpus_keeps_its_norms_between_gemm_searches():
    """K2 through ShardedCorpus skips the corpus-norm pre-pass from the second search on: results must not change, two
    corpora on one device must not see each other's norms, and invalidate() forces the pre-pass again."""
    from semanticsearch_b200 import sharded, similarity
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn((30000, 128), generator=g, device="cuda").to(torch.bfloat16)
    B = (torch.randn((30000, 128), generator=g, device="cuda") * 3.0).to(torch.bfloat16)   # different norms, same shape
    Q = torch.randn((256, 128), generator=g, device="cuda").to(torch.bfloat16)
    ca, cb = sharded.ShardedCorpus(A, 0), sharded.ShardedCorpus(B, 0)
    want_a = similarity.cosine_topk(A, Q, 10, algo="gemm")
    want_b = similarity.cosine_topk(B, Q, 10, algo="gemm")
    for _ in range(3):   # first call computes the norms, the later ones reuse them; the corpora alternate
        for corpus, want in ((ca, want_a), (cb, want_b)):
            s, i = corpus.search(Q, 10, algo="gemm")
            assert torch.equal(i, want[1]) and torch.equal(s, want[0])
    assert ca._resident.state is not None and ca._resident.ws is not cb._resident.ws
    # a smaller batch reuses the norms as well (their place in the workspace does not depend on the batch)
    s, i = ca.search(Q[:128].contiguous(), 10, algo="gemm")
    w = similarity.cosine_topk(A, Q[:128].contiguous(), 10, algo="gemm")
    assert torch.equal(i, w[1]) and torch.equal(s, w[0])
    A.mul_(0.5)                       # the corpus bytes change (cosines do not): the owner must say so
    ca._resident.invalidate()
    s, i = ca.search(Q, 10, algo="gemm")
    w = similarity.cosine_topk(A, Q, 10, algo="gemm")
    assert torch.equal(i, w[1]) and torch.equal(s, w[0])


def test_graph_survives_larger_eager_searches_and_streams_do_not_share_scratch():
    """(1) A captured search keeps working after eager searches that need bigger workspaces (the graph owns its buffers);
    (2) two streams running different searches at the same time return what each returns alone (scratch is per stream)."""
    from semanticsearch_b200 import sharded, similarity
    g = torch.Generator(device="cuda").manual_seed(9)
    C = torch.randn((300000, 128), generator=g, device="cuda").to(torch.bfloat16)
    for batch, k in ((1, 10), (16, 100), (256, 10)):
        Q = torch.randn((batch, 128), generator=g, device="cuda").to(torch.bfloat16)
        corpus = sharded.ShardedCorpus(C, 0)
        gs = sharded.GraphedSearch(corpus, batch, k)
        s0, i0 = gs(Q)
        s0, i0 = s0.clone(), i0.clone()
        big = torch.randn((4 * batch + 700, 128), generator=g, device="cuda").to(torch.bfloat16)
        for algo in ("stream", "tcstream", "gemm"):      # eager calls on the same corpus and on the module-level scratch
            if algo == "gemm" and k > 16:
                continue
            similarity.cosine_topk(C, big, k, algo=algo)
        with pytest.raises(RuntimeError, match="captured CUDA graph"):
            corpus.search(big, k, resident=gs._resident)  # the graph's own buffers refuse to grow
        s1, i1 = gs(Q)
        torch.cuda.synchronize()
        assert torch.equal(i0, i1) and torch.equal(s0, s1)
    # two streams
    Qa = torch.randn((8, 128), generator=g, device="cuda").to(torch.bfloat16)
    Qb = torch.randn((24, 128), generator=g, device="cuda").to(torch.bfloat16)
    want_a = similarity.cosine_topk(C, Qa, 10)
    want_b = similarity.cosine_topk(C, Qb, 50)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(10):
        with torch.cuda.stream(sa):
            ra = similarity.cosine_topk(C, Qa, 10)
        with torch.cuda.stream(sb):
            rb = similarity.cosine_topk(C, Qb, 50)
        outs.append((ra, rb))
    torch.cuda.synchronize()
    for ra, rb in outs:
        assert torch.equal(ra[1], want_a[1]) and torch.equal(ra[0], want_a[0])
        assert torch.equal(rb[1], want_b[1]) and torch.equal(rb[0], want_b[0])
