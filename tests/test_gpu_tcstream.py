"""GPU parity tests for K7 — the tensor-core streaming cosine kernel (corpus tile as the MMA M
operand, resident query group as N, in-kernel corpus norms, pooled top-k up to k = 1024).

Oracle: the reference arithmetic (sklearn cosine_similarity + argsort,
Tool/rank_chunks_optimized.py:215-216,225) on the same bf16/fp16-rounded values upcast to fp32.
Tolerance: scores 2e-3 abs for bf16/fp16 inputs (measured < 2e-5), indices identical up to fp64
ties < 1e-5.
"""
import numpy as np
import pytest
import torch

from oracle import rank_oracle as ro

pytestmark = pytest.mark.gpu
LOWP_TOL = 2e-3


def _run(C, Q, k, dtype, algo="tcstream", **kw):
    from semanticsearch_b200 import similarity
    Ct = torch.from_numpy(C).cuda().to(dtype).contiguous()
    Qt = torch.from_numpy(Q).cuda().to(dtype).contiguous()
    s, i = similarity.cosine_topk(Ct, Qt, k, algo=algo, **kw)
    torch.cuda.synchronize()
    return s.cpu().numpy(), i.cpu().numpy(), Ct.float().cpu().numpy(), Qt.float().cpu().numpy()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,d,b,k", [
    (60000, 384, 16, 100),    # BASELINE config 5 shape (per-shard slice)
    (50000, 768, 2, 10),
    (30011, 768, 64, 10),
    (4099, 384, 17, 33),      # ragged last tile, group padded to 32
    (257, 64, 1, 1),
    (20000, 72, 40, 16),      # dim not a multiple of 64: TMA zero-fills the K tail
    (3000, 768, 130, 7),      # more than 64 queries: several corpus passes
    (9000, 128, 5, 1000),     # k close to the 1024 limit, fewer rows per CTA than k
])
def test_tcstream_vs_oracle(dtype, n, d, b, k):
    rng = np.random.default_rng(n + d + b + k)
    C = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((b, d)).astype(np.float32)
    s, i, Cr, Qr = _run(C, Q, k, dtype)
    assert ro.check_topk_against_oracle(Qr, Cr, s, i, k, LOWP_TOL) == []
    ref_s, _ = ro.cosine_topk_ref(Qr, Cr, k)
    assert np.abs(ref_s - s).max() < 2e-5


def test_tcstream_matches_stream_kernel():
    rng = np.random.default_rng(43)
    C = rng.standard_normal((200000, 384)).astype(np.float32)
    Q = rng.standard_normal((16, 384)).astype(np.float32)
    s_t, i_t, _, _ = _run(C, Q, 100, torch.float16, "tcstream")
    s_s, i_s, _, _ = _run(C, Q, 100, torch.float16, "stream")
    assert (i_t == i_s).mean() > 0.995
    np.testing.assert_allclose(s_t, s_s, atol=2e-5, rtol=0)


def test_tcstream_edge_cases():
    rng = np.random.default_rng(5)
    d = 64
    C = rng.standard_normal((700, d)).astype(np.float32)
    C[7] = 0.0
    C[400] = C[2]
    C[600] = C[2]
    Q = np.stack([C[2] * 3.0, np.zeros(d, dtype=np.float32)] + [rng.standard_normal(d).astype(np.float32) for _ in range(10)])
    s, i, Cr, Qr = _run(C, Q, 5, torch.bfloat16)
    assert list(i[0][:3]) == [2, 400, 600]          # exact ties resolve to the lower index, across tiles
    assert np.all(s[1] == 0.0) and list(i[1]) == [0, 1, 2, 3, 4]   # zero query: all scores 0, lowest rows win
    assert ro.check_topk_against_oracle(Qr, Cr, s, i, 5, LOWP_TOL) == []
    # n < k: trailing slots are empty
    s, i, _, _ = _run(C[:3], Q[:2], 5, torch.bfloat16)
    assert list(i[0][3:]) == [-1, -1] and np.all(np.isneginf(s[0][3:]))
    # index_base shifts indices
    s2, i2, _, _ = _run(C, Q, 5, torch.bfloat16, index_base=1000)
    assert np.array_equal(i2, _run(C, Q, 5, torch.bfloat16)[1] + 1000)


def test_tcstream_all_equal_scores_fill_pools():
    """Every row identical: all candidates tie, every tile appends 128 entries per query until the
    threshold key (score, lowest index) closes — exercises pool compaction and the tie rule."""
    d = 128
    row = np.linspace(-1, 1, d).astype(np.float32)
    C = np.tile(row, (40000, 1))
    Q = np.stack([row, -row, row * 0.5])
    s, i, _, _ = _run(C, Q, 100, torch.float16)
    assert np.array_equal(i[0], np.arange(100)) and np.array_equal(i[1], np.arange(100))
    assert np.allclose(s[0], 1.0, atol=1e-3) and np.allclose(s[1], -1.0, atol=1e-3)


def test_tcstream_config5_shard_properties():
    """One 8-way shard of BASELINE config 5 at reduced length (2 M x 384 fp16, 16 queries, top-100):
    planted winners, sortedness, agreement with a torch fp32 matmul."""
    from semanticsearch_b200 import similarity
    g = torch.Generator(device="cuda").manual_seed(9)
    n, d, b, k = 2_000_000, 384, 16, 100
    C = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32).to(torch.float16)
    Q = torch.randn((b, d), generator=g, device="cuda", dtype=torch.float32).to(torch.float16)
    plant = torch.unique(torch.randint(0, n, (b,), generator=g, device="cuda"))
    C[plant] = Q[: plant.numel()] * 0.5
    s, i = similarity.cosine_topk(C, Q, k)  # auto -> tcstream
    assert similarity.choose_algo(C, Q, k) == "tcstream"
    assert torch.equal(i[: plant.numel(), 0], plant)
    assert torch.all(s[:, 1:] <= s[:, :-1])
    Cn = torch.nn.functional.normalize(C.float(), dim=1)
    ref = torch.nn.functional.normalize(Q.float(), dim=1) @ Cn.T
    rs, ri = torch.topk(ref, k, dim=1)
    assert torch.allclose(rs, s, atol=2e-5)
    assert (ri == i).float().mean().item() > 0.99


def test_config5_shard_full_length_against_exhaustive_oracle():
    """One whole 8-way shard of BASELINE config 5 (12.5 M x 384 fp16, 16 queries, top-100) against the exact answer of a
    chunked torch fp32 pass over every row (tests/conftest.py:exhaustive_topk_check), duplicates and a zero row included."""
    from conftest import exhaustive_topk_check
    from semanticsearch_b200 import similarity
    free, _total = torch.cuda.mem_get_info()
    if free < 16 * (1 << 30):
        pytest.skip("needs ~12 GB of free HBM")
    g = torch.Generator(device="cuda").manual_seed(9)
    n, d, b, k = 12_500_000, 384, 16, 100
    C = torch.empty((n, d), dtype=torch.float16, device="cuda")
    for a in range(0, n, 1 << 21):
        e = min(n, a + (1 << 21))
        C[a:e] = torch.randn((e - a, d), generator=g, device="cuda").half()
    Q = torch.randn((b, d), generator=g, device="cuda").half()
    C[5] = Q[0] * 0.25
    C[n - 3] = C[5]            # duplicate of a winner at the far end: the lower index must come first
    C[1234567] = 0
    s, i = similarity.cosine_topk(C, Q, k)
    assert similarity.choose_algo(C, Q, k) == "tcstream"
    assert int(i[0, 0]) == 5 and int(i[0, 1]) == n - 3
    assert torch.all(s[:, 1:] <= s[:, :-1])
    exhaustive_topk_check(C, Q, k, s, i, score_tol=2e-3)
