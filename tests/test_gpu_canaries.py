"""Out-of-bounds canaries for the round-2 kernels (compute-sanitizer is not available on the GPU pool): every output buffer
is over-allocated and pre-filled with a sentinel; after the launch the region past the documented size must be untouched
and the documented region fully written where the contract says so."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PAD = 4096


def _padded(n, dtype, fill):
    t = torch.full((n + PAD,), fill, dtype=dtype, device="cuda")
    return t


def test_canaries_around_the_new_kernels():
    from semanticsearch_b200 import _lib, ragged
    lib = _lib.load()
    rng = np.random.default_rng(5)
    sizes = [1, 2, 33, 97, 130, 64, 257, 5]
    E = torch.from_numpy(rng.standard_normal((sum(sizes), 48)).astype(np.float32)).cuda()
    plan = ragged.make_plan(sizes, "cuda")
    st = torch.cuda.current_stream().cuda_stream
    units = ragged._units128(plan, E.device)

    # K3 (fp16 cross terms): S packed, flag word
    S = _padded(plan.total_s, torch.float32, float("nan"))
    flag = torch.zeros(1 + PAD, dtype=torch.int32, device="cuda")
    _lib.check(lib.ss_segmented_simmatrix_tc(E.data_ptr(), E.shape[0], E.shape[1], plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(),
                                             units.data_ptr(), units.shape[0], S.data_ptr(), flag.data_ptr(), st), "simmatrix_tc")
    torch.cuda.synchronize()
    assert not torch.isnan(S[: plan.total_s]).any() and torch.isnan(S[plan.total_s:]).all() and int(flag.abs().sum()) == 0
    S = S[: plan.total_s].contiguous()

    # K4 with the symmetric promise
    sharp = _padded(plan.total_s, torch.float32, float("nan"))
    cent = _padded(plan.total_rows, torch.float64, float("nan"))
    stats = _padded(plan.n_docs * 8, torch.float64, float("nan"))
    kidx = _padded(plan.total_rows * 33, torch.int32, -7)
    kval = _padded(plan.total_rows * 33, torch.float32, float("nan"))
    _lib.check(lib.ss_group_threshold_pass(S.data_ptr(), plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), plan.n_docs, 0.15, 0, 1,
                                           sharp.data_ptr(), cent.data_ptr(), stats.data_ptr(), kidx.data_ptr(), kval.data_ptr(), st), "group pass")
    torch.cuda.synchronize()
    assert torch.isnan(sharp[plan.total_s:]).all() and not torch.isnan(sharp[: plan.total_s]).any()
    assert torch.isnan(cent[plan.total_rows:]).all() and torch.isnan(stats[plan.n_docs * 8:]).all()
    assert bool((kidx[plan.total_rows * 33:] == -7).all()) and torch.isnan(kval[plan.total_rows * 33:]).all()
    assert bool((kidx[: plan.total_rows * 33] != -7).all())

    # C99 rank (tiled local mode, every compiled window) and cut search
    ws_len = plan.total_rows + plan.total_rows // 16 + 2 * plan.n_docs + 8
    for mask in (3, 5, 7, 9, 11, 13, 15):
        R = _padded(plan.total_s, torch.float32, float("nan"))
        ws = _padded(ws_len, torch.int32, -7)
        _lib.check(lib.ss_c99_rank_matrix(S.data_ptr(), plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), plan.n_docs, plan.total_rows,
                                          plan.max_rows, 1, mask, ws.data_ptr(), R.data_ptr(), st), "c99 rank")
        torch.cuda.synchronize()
        assert torch.isnan(R[plan.total_s:]).all() and not torch.isnan(R[: plan.total_s]).any(), mask
        assert bool((ws[ws_len:] == -7).all()), mask
        assert float(R[: plan.total_s].max()) <= 1.0 and float(R[: plan.total_s].min()) >= 0.0

    # diameter split
    ends = _padded(plan.total_rows, torch.int32, -7)
    nsp = _padded(plan.n_docs, torch.int32, -7)
    diam = _padded(plan.n_docs, torch.float64, float("nan"))
    _lib.check(lib.ss_diameter_split(S.data_ptr(), plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), plan.n_docs, plan.max_rows, 0.5,
                                     ends.data_ptr(), nsp.data_ptr(), diam.data_ptr(), st), "diameter")
    torch.cuda.synchronize()
    assert bool((ends[plan.total_rows:] == -7).all()) and bool((nsp[plan.n_docs:] == -7).all()) and torch.isnan(diam[plan.n_docs:]).all()
    assert bool((nsp[: plan.n_docs] >= 1).all())

    # block sums for a batch
    groups = [[list(range(0, n // 2)), list(range(n // 2, n)), [0, 0]] for n in sizes]
    got = ragged.group_block_sums(sharp[: plan.total_s].contiguous(), plan, groups)
    assert len(got) == len(sizes) and all(np.isfinite(r).all() and np.isfinite(b).all() for r, b in got)


def test_canaries_around_k2_outputs():
    from semanticsearch_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(2)
    for n, b, k in ((70000, 300, 10), (4000, 130, 16), (300, 257, 3)):
        C = torch.randn((n, 64), generator=g, device="cuda").to(torch.bfloat16)
        Q = torch.randn((b, 64), generator=g, device="cuda").to(torch.bfloat16)
        need = lib.ss_cosine_topk_gemm_workspace_bytes(n, 64, b, k)
        ws = torch.full((need + PAD,), 0x5A, dtype=torch.uint8, device="cuda")
        keys = _padded(b * k, torch.int64, -7)
        scores = _padded(b * k, torch.float32, float("nan"))
        idx = _padded(b * k, torch.int64, -7)
        _lib.check(lib.ss_cosine_topk_gemm(C.data_ptr(), n, 64, _lib.SS_BF16, Q.data_ptr(), b, k, 0, ws.data_ptr(), need, keys.data_ptr(),
                                           scores.data_ptr(), idx.data_ptr(), torch.cuda.current_stream().cuda_stream), "gemm")
        torch.cuda.synchronize()
        assert bool((ws[need:] == 0x5A).all()), (n, b, k)
        assert bool((keys[b * k:] == -7).all()) and bool((idx[b * k:] == -7).all()) and torch.isnan(scores[b * k:]).all()
        assert bool((idx[: b * k] >= 0).all()) and bool((idx[: b * k] < n).all())
