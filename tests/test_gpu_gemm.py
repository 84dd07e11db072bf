"""GPU parity tests for K2 — the tcgen05/TMEM tensor-core cosine kernel with fused top-k.

The oracle is the reference arithmetic (sklearn cosine_similarity + argsort,
Tool/rank_chunks_optimized.py:215-216,225) applied to the same bf16/fp16-rounded values upcast to
fp32.  Tolerance: scores 2e-3 abs for bf16/fp16 inputs, indices identical up to fp64 ties < 1e-5.
"""
import numpy as np
import pytest
import torch

from oracle import rank_oracle as ro

pytestmark = pytest.mark.gpu
LOWP_TOL = 2e-3


def _run(C, Q, k, dtype, algo="gemm", **kw):
    from semanticsearch_b200 import similarity
    Ct = torch.from_numpy(C).cuda().to(dtype).contiguous()
    Qt = torch.from_numpy(Q).cuda().to(dtype).contiguous()
    s, i = similarity.cosine_topk(Ct, Qt, k, algo=algo, **kw)
    torch.cuda.synchronize()
    return s.cpu().numpy(), i.cpu().numpy(), Ct.float().cpu().numpy(), Qt.float().cpu().numpy()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,d,b,k", [(50000, 768, 300, 10), (1000, 384, 130, 10), (4099, 768, 16, 4), (257, 64, 1, 1),
                                     (20000, 72, 128, 16), (300, 768, 5, 8)])
def test_gemm_path_vs_oracle(dtype, n, d, b, k):
    rng = np.random.default_rng(n + d + b + k)
    C = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((b, d)).astype(np.float32)
    s, i, Cr, Qr = _run(C, Q, k, dtype)
    assert ro.check_topk_against_oracle(Qr, Cr, s, i, k, LOWP_TOL) == []
    # in practice the fp32-accumulated scores agree far better than the bf16 tolerance
    ref_s, _ = ro.cosine_topk_ref(Qr, Cr, k)
    assert np.abs(ref_s - s).max() < 2e-5


def test_gemm_matches_stream_kernel_exactly_on_indices():
    rng = np.random.default_rng(42)
    C = rng.standard_normal((120000, 768)).astype(np.float32)
    Q = rng.standard_normal((200, 768)).astype(np.float32)
    s_g, i_g, _, _ = _run(C, Q, 10, torch.bfloat16, "gemm")
    s_s, i_s, _, _ = _run(C, Q, 10, torch.bfloat16, "stream")
    assert (i_g == i_s).mean() > 0.999
    np.testing.assert_allclose(s_g, s_s, atol=2e-5, rtol=0)


def test_gemm_edge_cases():
    rng = np.random.default_rng(5)
    d = 64
    C = rng.standard_normal((700, d)).astype(np.float32)
    C[7] = 0.0
    C[400] = C[2]
    C[600] = C[2]
    Q = np.stack([C[2] * 3.0, np.zeros(d, dtype=np.float32)] + [rng.standard_normal(d).astype(np.float32) for _ in range(30)])
    s, i, Cr, Qr = _run(C, Q, 5, torch.bfloat16)
    assert list(i[0][:3]) == [2, 400, 600]          # exact ties resolve to the lower index, across tiles
    assert np.all(s[1] == 0.0) and list(i[1]) == [0, 1, 2, 3, 4]
    assert ro.check_topk_against_oracle(Qr, Cr, s, i, 5, LOWP_TOL) == []
    # n < k: trailing slots are empty
    s, i, _, _ = _run(C[:3], Q[:2], 5, torch.bfloat16)
    assert list(i[0][3:]) == [-1, -1] and np.all(np.isneginf(s[0][3:]))
    # index_base shifts indices
    s2, i2, _, _ = _run(C, Q, 5, torch.bfloat16, index_base=1000)
    assert np.array_equal(i2, _run(C, Q, 5, torch.bfloat16)[1] + 1000)


def test_gemm_large_batch_properties():
    """B = 1024 over 400k rows: planted winners, sortedness, agreement with a torch fp32 matmul."""
    from semanticsearch_b200 import similarity
    g = torch.Generator(device="cuda").manual_seed(8)
    n, d, b, k = 400_000, 768, 1024, 10
    C = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    Q = torch.randn((b, d), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    plant = torch.randint(0, n, (b,), generator=g, device="cuda")
    plant = torch.unique(plant)
    C[plant] = Q[: plant.numel()] * 1.5
    s, i = similarity.cosine_topk(C, Q, k)  # auto -> tensor-core path
    assert torch.equal(i[: plant.numel(), 0], plant)
    assert torch.all(s[: plant.numel(), 0] > 0.999)
    assert torch.all(s[:, 1:] <= s[:, :-1])
    ref = torch.nn.functional.normalize(Q.float(), dim=1) @ torch.nn.functional.normalize(C.float(), dim=1).T
    rs, ri = torch.topk(ref, k, dim=1)
    assert torch.allclose(rs, s, atol=2e-5)
    assert (ri == i).float().mean().item() > 0.995


def test_full_size_config4_properties():
    """BASELINE config 4 at full size (10 M x 768 bf16, 4096 queries, top-10) through size-independent
    properties: planted winners come first, lists are sorted and duplicate-free, the three kernels agree
    on a query subset, and re-scoring the returned rows in fp32 reproduces the returned scores."""
    from semanticsearch_b200 import similarity
    free, _total = torch.cuda.mem_get_info()
    if free < 24 * (1 << 30):
        pytest.skip("needs ~20 GB of free HBM")
    g = torch.Generator(device="cuda").manual_seed(6)
    n, d, b, k = 10_000_000, 768, 4096, 10
    C = torch.empty((n, d), dtype=torch.bfloat16, device="cuda")
    for a in range(0, n, 1 << 20):
        e = min(n, a + (1 << 20))
        C[a:e] = torch.randn((e - a, d), generator=g, device="cuda").to(torch.bfloat16)
    Q = torch.randn((b, d), generator=g, device="cuda").to(torch.bfloat16)
    plant = torch.unique(torch.randint(0, n, (256,), generator=g, device="cuda"))
    C[plant] = Q[: plant.numel()] * 2.0
    s, i = similarity.cosine_topk(C, Q, k)                       # K2 (CTA pairs)
    assert similarity.choose_algo(C, Q, k) == "gemm"
    assert torch.equal(i[: plant.numel(), 0], plant) and torch.all(s[: plant.numel(), 0] > 0.999)
    assert torch.all(s[:, 1:] <= s[:, :-1])
    assert all(len(set(row)) == k for row in i[:64].tolist())
    # fp32 re-score of the returned rows
    rows = C[i[:32].reshape(-1)].float().view(32, k, d)
    qn = torch.nn.functional.normalize(Q[:32].float(), dim=1)
    ref = torch.einsum("bkd,bd->bk", torch.nn.functional.normalize(rows, dim=2), qn)
    assert torch.allclose(ref, s[:32], atol=2e-5)
    # the streaming kernels return the same lists
    s7, i7 = similarity.cosine_topk(C, Q[:48].contiguous(), k, algo="tcstream")
    assert (i7 == i[:48]).float().mean().item() > 0.995 and torch.allclose(s7, s[:48], atol=2e-5)
    s1, i1 = similarity.cosine_topk(C, Q[:1].contiguous(), k, algo="stream")
    assert (i1 == i[:1]).float().mean().item() >= 0.9 and torch.allclose(s1, s[:1], atol=2e-5)
    # independent oracle over ALL 10 M rows for 32 queries (planted and ordinary ones): a row that every kernel of this
    # repo dropped would still be found here
    from conftest import exhaustive_topk_check
    pick = torch.cat([torch.arange(0, 16, device="cuda"), torch.arange(2000, 2016, device="cuda")])
    exhaustive_topk_check(C, Q[pick], k, s[pick], i[pick], score_tol=2e-3)
    exhaustive_topk_check(C, Q[:16], k, s7[:16], i7[:16], score_tol=2e-3)
    exhaustive_topk_check(C, Q[:1], k, s1, i1, score_tol=2e-3)
