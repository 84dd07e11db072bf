"""GPU parity tests for the tensor-core form of K3 (tcgen05 kind::tf32 with hi/lo operand splitting).

Oracle: create_similarity_matrix's arithmetic (Method/semantic_common.py:158-164,186-191) restated in
numpy (oracle/simmatrix_oracle.py) and an fp64 recomputation.  Tolerance 1e-5 abs (BASELINE.json
north_star, fp32 inputs); the measured error of the 3xTF32 product must stay below 6e-6.
"""
import os

import numpy as np
import pytest
import torch

from oracle import simmatrix_oracle as so

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _docs(rng, sizes, d, noise=0.7):
    rows = []
    for n in sizes:
        n_topics = max(1, (n + 11) // 12)
        cent = rng.standard_normal((n_topics, d)).astype(np.float32)
        topic = np.minimum(np.arange(n) // 12, n_topics - 1)
        rows.append((cent[topic] + noise * rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32))
    return rows


def _run(rows_list, algo):
    from semanticsearch_b200 import ragged
    sizes = [r.shape[0] for r in rows_list]
    E = np.concatenate([r for r in rows_list if r.shape[0] > 0], axis=0)
    plan = ragged.make_plan(sizes, "cuda")
    S = ragged.segmented_simmatrix(torch.from_numpy(E).cuda(), plan, algo=algo)
    torch.cuda.synchronize()
    S_host = S.cpu().numpy()
    return [S_host[plan.s_offsets[i]:plan.s_offsets[i + 1]].reshape(sizes[i], sizes[i]) for i in range(len(sizes))]


def _fp64(E):
    E = E.astype(np.float64)
    nrm = np.linalg.norm(E, axis=1, keepdims=True)
    nrm[nrm == 0] = 1e-9
    En = E / nrm
    return En @ En.T


@pytest.mark.parametrize("d", [768, 384, 100, 4])
def test_tc_ragged_batch_vs_oracle(d):
    rng = np.random.default_rng(30 + d)
    sizes = [int(x) for x in rng.integers(16, 513, size=24)] + [1, 2, 0, 64, 65, 127, 128, 129, 256, 257, 511, 512, 513, 700]
    rows = _docs(rng, sizes, d)
    blocks = _run(rows, "tc")
    worst = 0.0
    for E, S in zip(rows, blocks):
        if E.shape[0] == 0:
            continue
        np.testing.assert_array_equal(S, S.T)                     # exactly symmetric
        if E.shape[0] >= 2:
            np.testing.assert_allclose(S, so.similarity_matrix_ref(E), atol=TOL, rtol=0)
        worst = max(worst, float(np.abs(S - _fp64(E)).max()))
    print(f"d={d}: max |S - S_fp64| = {worst:.3e}")
    assert worst < 6e-6, worst


def test_tc_matches_ffma_kernel_and_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "grouping.npz"))
    for name in ("a", "b", "c", "tiny", "seven"):
        E = g[f"{name}_E"]
        S_tc = _run([E], "tc")[0]
        S_ff = _run([E], "ffma")[0]
        np.testing.assert_allclose(S_tc, g[f"{name}_S"], atol=TOL, rtol=0)   # the reference's own output
        np.testing.assert_allclose(S_tc, S_ff, atol=6e-6, rtol=0)
    S = _run([g["a_E"]], "tc")[0]
    assert np.all(S[5] == 0) and np.all(S[:, 5] == 0)  # zero sentence vector -> zero row/column


def test_tc_adversarial_magnitudes():
    """Rows scaled over 12 orders of magnitude and near-duplicate rows: the hi/lo split must stay relative."""
    rng = np.random.default_rng(77)
    n, d = 300, 768
    E = rng.standard_normal((n, d)).astype(np.float32)
    E *= (10.0 ** rng.uniform(-6, 6, size=(n, 1))).astype(np.float32)
    E[10] = E[3] * np.float32(1.0000001)
    E[11] = -E[3]
    E[12] = 0
    S = _run([E], "tc")[0]
    np.testing.assert_allclose(S, _fp64(E), atol=TOL, rtol=0)
    assert abs(S[3, 10] - 1.0) < 6e-6 and abs(S[3, 11] + 1.0) < 6e-6
    assert np.all(S[12] == 0)


def test_tc_packed_windows_of_small_documents():
    """Documents of <= 64 sentences share 128-row tiles (the reference corpus: median 10 sentences): every
    document's block must come out exactly as if it had a tile of its own — symmetric, zero rows zero,
    nothing written across document boundaries."""
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(41)
    sizes = [int(x) for x in np.clip(np.rint(rng.lognormal(np.log(10.0), 1.2, size=300)), 1, 200)] + [0, 64, 64, 1, 0, 2, 63, 65, 0]
    rows = _docs(rng, sizes, 384)
    rows[5][0] = 0.0
    E = np.concatenate([r for r in rows if r.shape[0] > 0], axis=0)
    plan = ragged.make_plan(sizes, "cuda")
    S = torch.full((plan.total_s + 7,), float("nan"), dtype=torch.float32, device="cuda")   # canary past the end
    ragged.segmented_simmatrix(torch.from_numpy(E).cuda(), plan, out=S, algo="tc")
    torch.cuda.synchronize()
    Sh = S.cpu().numpy()
    assert np.all(np.isnan(Sh[plan.total_s:])) and not np.any(np.isnan(Sh[: plan.total_s]))
    for d, n in enumerate(sizes):
        if n == 0:
            continue
        blk = Sh[plan.s_offsets[d]:plan.s_offsets[d + 1]].reshape(n, n)
        np.testing.assert_array_equal(blk, blk.T)
        np.testing.assert_allclose(blk, _fp64(rows[d]), atol=TOL, rtol=0)
    blk5 = Sh[plan.s_offsets[5]:plan.s_offsets[6]].reshape(sizes[5], sizes[5])
    assert np.all(blk5[0] == 0)


@pytest.mark.parametrize("kind", ["all_positive", "near_duplicates", "alternating", "huge_dynamic_range"])
def test_tc_worst_cases_at_d768_stay_inside_the_parity_bound(kind):
    """The 3xTF32 error grows with the width and with |S| (the tensor core truncates when it accumulates): pin the worst
    cases at the widest supported width, d = 768 — all-positive rows (every S near 1, every product of one sign),
    near-duplicates, sign-alternating rows and a 2^20 dynamic range inside one row — under the 1e-5 bound of north_star."""
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(768)
    n, d = 512, 768
    if kind == "all_positive":
        E = np.abs(rng.standard_normal((n, d))).astype(np.float32) + 0.5
    elif kind == "near_duplicates":
        base = rng.standard_normal((1, d)).astype(np.float32)
        E = (base + 1e-3 * rng.standard_normal((n, d))).astype(np.float32)
    elif kind == "alternating":
        E = (np.abs(rng.standard_normal((n, d))) * np.where(np.arange(d) % 2 == 0, 1.0, -1.0)).astype(np.float32)
    else:
        E = (rng.standard_normal((n, d)) * np.exp2(rng.integers(-10, 11, size=(n, d)))).astype(np.float32)
    plan = ragged.make_plan([n, 37], "cuda")
    S = ragged.segmented_simmatrix(torch.from_numpy(np.concatenate([E, E[:37]])).cuda(), plan, algo="tc").cpu().numpy()
    got = S[: n * n].reshape(n, n)
    E64 = E.astype(np.float64)
    E64 /= np.linalg.norm(E64, axis=1, keepdims=True)
    err = float(np.abs(got - E64 @ E64.T).max())
    print(f"{kind}: max |S - S_fp64| = {err:.3e}")
    assert err <= 1e-5
    assert np.array_equal(got, got.T)


def test_tc_row_scaling_of_the_fp16_cross_terms():
    """The cross terms run on fp16 copies scaled per row by the exponent of the row's first 32 elements: rows whose first
    block is tiny / huge / zero compared with the rest, and whole rows at 1e-12 ... 1e12, stay inside the parity bound; a
    row with a later element > 2^14 times its first block raises the range flag and `validate=True` recomputes exactly."""
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(91)
    n, d = 160, 384
    E = rng.standard_normal((n, d)).astype(np.float32)
    E[0, :32] *= 1e-3                      # first block 1000x smaller than the rest (still inside 2^14)
    E[1, :32] *= 1e3                       # first block 1000x larger: the rest shrinks towards fp16's subnormals
    E[2, :32] = 0.0                        # all-zero first block: scale 1
    E[3] *= 1e-12                          # (fp32 sums of squares underflow below ~1e-19, in numpy too)
    E[4] *= 1e12
    E[5, :32] = 0.0
    E[5] *= 1e-8                           # zero first block AND a tiny row
    E[6] = np.abs(E[6]) * 3.0              # all positive
    E[7] = E[6] * np.float32(1.5)          # parallel to row 6
    plan = ragged.make_plan([n], "cuda")
    Ed = torch.from_numpy(E).cuda()
    S = ragged.segmented_simmatrix(Ed, plan, algo="tc", validate=True).cpu().numpy().reshape(n, n)
    err = np.abs(S - _fp64(E))
    print(f"row scaling: max |S - S_fp64| = {err.max():.3e} (rows 0-7: {err[:8].max():.3e})")
    assert err.max() <= 1e-5
    assert abs(S[6, 7] - 1.0) < 6e-6
    # saturation: element 40 of row 9 is 1e6 times the row's first block
    E2 = E.copy()
    E2[9, 40] = 1e6
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    from semanticsearch_b200 import _lib
    lib = _lib.load()
    units = ragged._units128(plan, Ed.device)
    E2d = torch.from_numpy(E2).cuda()
    out = torch.empty(plan.total_s, dtype=torch.float32, device="cuda")
    st = lib.ss_segmented_simmatrix_tc(E2d.data_ptr(), n, d, plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), units.data_ptr(),
                                       units.shape[0], out.data_ptr(), flag.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert st == 0 and int(flag.item()) == 1
    S2 = ragged.segmented_simmatrix(E2d, plan, algo="tc", validate=True).cpu().numpy().reshape(n, n)   # falls to the fp32 kernel
    np.testing.assert_allclose(S2, _fp64(E2), atol=1e-5, rtol=0)
