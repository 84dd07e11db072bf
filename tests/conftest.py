import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def _has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def exhaustive_topk_check(C, Q, k, got_s, got_i, score_tol, tie_tol=1e-5, chunk=1 << 20):
    """Independent oracle at BASELINE sizes: the exact top-k of every query in ``Q`` over ALL rows of the CUDA corpus ``C``
    from a chunked torch fp32 computation (upcast storage values, L2-normalise rows with sklearn's zero rule, matmul) —
    the reference's arithmetic, no kernel of this repo involved.  Asserts scores within ``score_tol`` and identical
    indices except where the exact scores of the two rows tie within ``tie_tol`` (float64 re-score), and that no row the
    kernel dropped beats its k-th pick."""
    import torch
    dev = C.device
    b = Q.shape[0]
    qn = Q.float()
    qn = qn / torch.where(qn.norm(dim=1, keepdim=True) == 0, torch.ones(1, device=dev), qn.norm(dim=1, keepdim=True))
    best_s = torch.full((b, k), -float("inf"), device=dev)
    best_i = torch.full((b, k), -1, dtype=torch.int64, device=dev)
    for a in range(0, C.shape[0], chunk):
        c = C[a:a + chunk].float()
        nrm = c.norm(dim=1, keepdim=True)
        c = c / torch.where(nrm == 0, torch.ones_like(nrm), nrm)
        sc = qn @ c.T
        ts, ti = torch.topk(sc, min(k, sc.shape[1]), dim=1)
        cat_s, cat_i = torch.cat([best_s, ts], dim=1), torch.cat([best_i, ti + a], dim=1)
        order = torch.argsort(cat_s, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = torch.gather(cat_s, 1, order), torch.gather(cat_i, 1, order)
    assert torch.allclose(got_s, best_s, atol=score_tol, rtol=0), float((got_s - best_s).abs().max())
    diff = (got_i != best_i).nonzero().tolist()
    q64 = Q.double()
    q64 = q64 / q64.norm(dim=1, keepdim=True).clamp_min(1e-300)
    for qi, slot in diff:
        r_got, r_want = int(got_i[qi, slot]), int(best_i[qi, slot])
        rows = C[[r_got, r_want]].double()
        rows = rows / rows.norm(dim=1, keepdim=True).clamp_min(1e-300)
        e = rows @ q64[qi]
        assert abs(float(e[0] - e[1])) <= tie_tol, (qi, slot, r_got, r_want, float(e[0]), float(e[1]))
    return len(diff)
