"""GPU test of the diameter-bounded splitting (a12, data_process/simple_chunk_controller.py:571-594,614) against the
numpy restatement applied to the kernel's own similarity matrices, and the host-facing batch helper."""
import numpy as np
import pytest
import torch

from oracle import simmatrix_oracle as so

pytestmark = pytest.mark.gpu


def _docs(rng, sizes, d):
    out = []
    for n in sizes:
        topics = rng.standard_normal((max(1, n // 6 + 1), d))
        out.append((topics[np.arange(n) // 6] + 0.5 * rng.standard_normal((n, d))).astype(np.float32))
    return out


@pytest.mark.parametrize("threshold", [0.3, 0.6, 0.9, 1.5])
def test_diameter_split_matches_the_controller_recursion(threshold):
    from semanticsearch_b200 import ragged
    rng = np.random.default_rng(31)
    sizes = [2, 3, 7, 1, 24, 40, 100, 260, 5]
    rows = _docs(rng, sizes, 32)
    rows[4][3] = rows[4][4]                      # a duplicated sentence (similarity 1)
    plan = ragged.make_plan(sizes, "cuda")
    S = ragged.segmented_simmatrix(torch.from_numpy(np.concatenate(rows)).cuda(), plan)
    ends, n_spans, diam = ragged.diameter_split(S, plan, threshold)
    S_h, ends_h, n_h, diam_h = S.cpu().numpy(), ends.cpu().numpy(), n_spans.cpu().numpy(), diam.cpu().numpy()
    for d, n in enumerate(sizes):
        Sd = S_h[plan.s_offsets[d]:plan.s_offsets[d + 1]].reshape(n, n)
        want = so.split_indices_by_diameter_ref(Sd, 0, n, threshold)
        e = [int(x) for x in ends_h[plan.offsets[d]:plan.offsets[d] + n_h[d]]]
        assert list(zip([0] + e[:-1], e)) == want, (d, n)
        if n >= 2:
            off = Sd[~np.eye(n, dtype=bool)]
            assert diam_h[d] == 1.0 - float(off.min())
        else:
            assert diam_h[d] == 0.0


def test_split_indices_by_diameter_helper():
    from semanticsearch_b200.Method import semantic_common as sc
    rng = np.random.default_rng(32)
    rows = _docs(rng, [12, 1, 30], 24)
    got = sc.split_indices_by_diameter(rows + [None], 0.5)
    assert got[1] == [(0, 1)] and got[3] == []
    for E, spans in ((rows[0], got[0]), (rows[2], got[2])):
        En = E / np.linalg.norm(E, axis=1, keepdims=True)
        S = (En @ En.T).astype(np.float32)
        assert spans[0][0] == 0 and spans[-1][1] == len(E) and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        for a, b in spans:                        # every final span respects the bound (or is a single sentence)
            if b - a >= 2:
                sub = S[a:b, a:b]
                assert 1.0 - sub[~np.eye(b - a, dtype=bool)].min() <= 0.5 + 1e-5
