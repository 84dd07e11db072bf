"""CPU tests of the sequential boundary logic the splitter drop-in keeps on the host (valley detection, median smoothing,
score-ordered non-maximum suppression, text pre-cleaning) against the reference's own functions where the reference tree
is mounted, plus known answers that hold everywhere."""
import numpy as np
import pytest

from oracle import ref_shim
from semanticsearch_b200.Method import Semantic_Splitter_Optimized as SP


def test_median_smooth_known_answers():
    assert SP._median_smooth([3.0, 1.0, 2.0], 1) == [3.0, 1.0, 2.0]
    assert SP._median_smooth([3.0, 1.0, 2.0, 9.0], 3) == [3.0, 2.0, 2.0, 9.0]      # edge-replicated padding
    assert SP._median_smooth([3.0, 1.0, 2.0, 9.0], 2) == [3.0, 2.0, 2.0, 9.0]      # even windows are bumped to odd
    assert SP._median_smooth([1.0, 5.0], 5) == [1.0, 5.0]                            # window longer than the series
    assert SP._median_smooth([], 3) == []


def test_valley_boundaries_known_answer():
    sims = [0.9, 0.8, 0.2, 0.85, 0.9, 0.88, 0.3, 0.35, 0.9, 0.1, 0.95]
    got = SP._valley_boundaries(sims, triplet_tau=0.12, min_boundary_spacing=2, min_first_boundary_index=1)
    assert got == sorted(set(got)) and set(got) <= {3, 7, 10}      # a boundary sits AFTER the sentence that closes a valley
    assert 10 in got or 3 in got
    assert SP._valley_boundaries([0.5, 0.4], min_first_boundary_index=0) == []      # fewer than 3 similarities
    assert SP._valley_boundaries([0.1, 0.2, 0.3, 0.4], min_first_boundary_index=0) == []   # monotone: no valley


def test_score_based_nms_known_answer():
    score = {4: 0.9, 5: 0.8, 9: 0.7, 12: 0.95}
    assert sorted(SP._score_based_nms([4, 5, 9, 12], score, 3)) == [4, 9, 12]
    assert sorted(SP._score_based_nms([4, 5, 9, 12], score, 1)) == [4, 5, 9, 12]
    assert SP._score_based_nms([], score, 3) == []


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_host_boundary_logic_matches_the_live_reference():
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(12)
    for trial in range(40):
        n = int(rng.integers(2, 80))
        # plateaus and repeated values on purpose: the valley walk has <= / >= branches
        sims = np.round(rng.random(n), 1 if trial % 2 else 3).tolist()
        kw = {"triplet_tau": float(rng.choice([0.05, 0.12, 0.3])), "min_boundary_spacing": int(rng.integers(1, 6)),
              "min_first_boundary_index": int(rng.integers(0, 6))}
        assert SP._valley_boundaries(sims, **kw) == ref.split._valley_boundaries(sims, **kw), (trial, kw)
        for w in (1, 2, 3, 5):
            assert SP._median_smooth(sims, w) == ref.split._median_smooth(sims, w)
        cand = sorted(set(int(x) for x in rng.integers(0, max(n, 2), size=min(n, 12))))
        score = {b: float(np.round(rng.random(), 1)) for b in cand}
        spacing = int(rng.integers(1, 5))
        assert SP._score_based_nms(cand, score, spacing) == ref.split._score_based_nms(cand, score, spacing)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_divisive_search_statement_matches_live_reference_on_handmade_rank_matrices():
    """The reference's _c99_boundaries builds R from embeddings; block-structured embeddings give rank matrices with
    many exactly equal block means (identical sentences), the case where first-best tie-breaking matters."""
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(13)
    for trial in range(6):
        topics = rng.standard_normal((int(rng.integers(2, 6)), 16)).astype(np.float32)
        reps = rng.integers(3, 9, size=len(topics))
        En = np.concatenate([np.repeat(t[None, :], r, axis=0) for t, r in zip(topics, reps)])   # identical rows per topic
        En = (En / np.linalg.norm(En, axis=1, keepdims=True)).astype(np.float32)
        for m in (2, 3):
            want = ref.split._c99_boundaries(En, min_chunk_size=m)
            S = En @ En.T
            row = (S[:, None, :] < S[:, :, None]).sum(axis=2)
            col = (S.T[:, None, :] < S.T[:, :, None]).sum(axis=2).T
            R = (row + col).astype(np.float32)
            from oracle import splitter_oracle as spo
            assert spo.c99_divisive_f64_ref(R, m, None, 0.01, "gain", 1.2, 3) == [int(x) for x in want], (trial, m)
