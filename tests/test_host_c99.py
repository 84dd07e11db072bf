"""CPU tests of the host side of the C99 leg: the float64 summed-area statement of the divisive search (the form the
K10 kernel mirrors) and the profile knee against the reference's own outputs, plus the paths of the batched drop-in
that must not touch the GPU."""
import json
import os

import numpy as np

from oracle import splitter_oracle as spo
from semanticsearch_b200.Method import Semantic_Splitter_Optimized as SP


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "c99_cuts.npz"))
    meta = json.loads(str(g["meta_json"]))
    for name, m in meta.items():
        kw = dict(m["kwargs"])
        S = spo.c99_similarity_ref(g[f"{name}_En"])
        R = spo.c99_local_rank_ref(S, kw.get("mask_size", 11)) if kw.get("use_local_rank") else spo.c99_global_rank_ref(S)
        yield name, m, kw, R, g[f"{name}_D"]


def test_float64_block_sum_search_reproduces_reference_boundaries(golden_dir):
    n_cases = 0
    for name, m, kw, R, _D in _cases(golden_dir):
        got = spo.c99_divisive_f64_ref(R, int(kw.get("min_chunk_size", 3)), kw.get("max_cuts"), float(kw.get("min_gain", 0.01)),
                                kw.get("stopping", "gain"), float(kw.get("knee_c", 1.2)), int(kw.get("smooth_window", 3)))
        assert got == m["bounds"], name
        n_cases += 1
    assert n_cases >= 14


def test_profile_knee_matches_oracle_and_reference(golden_dir):
    for name, m, kw, _R, D in _cases(golden_dir):
        if kw.get("stopping") != "profile" or not m["cuts"]:
            continue
        knee_c, sw = float(kw.get("knee_c", 1.2)), int(kw.get("smooth_window", 3))
        want = spo.c99_profile_knee_ref(m["cuts"], list(D), knee_c, sw)
        assert SP._profile_knee(m["cuts"], D, knee_c, sw) == want == m["bounds"], name
    # a profile with an obvious knee after the second cut: increments 1.0, 1.0, 0.01, 1.0, 1.0
    cuts = [50, 20, 70, 10, 90]
    D = np.cumsum([1.0, 1.0, 1.0, 0.01, 1.0, 1.0])
    assert SP._profile_knee(cuts, D, 1.0, 1) == spo.c99_profile_knee_ref(cuts, list(D), 1.0, 1) == [20, 50]


def test_block_sums_table():
    rng = np.random.default_rng(1)
    R = rng.integers(0, 50, size=(23, 23)).astype(np.float32)
    b = spo.BlockSumsF64(R)
    for a, c in ((0, 23), (3, 9), (22, 23), (5, 5)):
        assert b.total(a, c) == float(R[a:c, a:c].astype(np.float64).sum())
    assert b.mean(5, 5, default=7.5) == 7.5 and b.mean(3, 9) == float(R[3:9, 3:9].astype(np.float64).mean())


def test_batch_paths_that_need_no_gpu():
    docs = [np.zeros((5, 8), np.float32), np.zeros((2, 8), np.float32)]
    assert SP.c99_boundaries_batch(docs, 3) == [[], []]                 # n < 2 * min_chunk (reference :165-166)
    assert SP.c99_boundaries_batch([np.zeros((40, 8), np.float32)], 3, max_cuts=0) == [[]]  # :225 stops before the first cut
    assert SP._c99_boundaries(np.zeros((4, 8), np.float32), min_chunk_size=3) == []
