"""TEST INFRASTRUCTURE — CPU restatement ("oracle") of the reference's dense-similarity hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or as the timed CPU baseline.  The product package ``semanticsearch_b200`` never imports
``oracle`` and raises if its CUDA library is missing.

Parity pinning: the reference repository ships no tests, golden vectors or fixtures
(SURVEY.md §4, §8c).  The oracle is therefore pinned against *outputs of the reference itself*,
executed in the build container through ``oracle/ref_shim.py`` by ``oracle/gen_golden.py``;
the resulting vectors are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py`` (which also re-runs the live reference when ``/root/reference``
is present).

Modules (each function cites the reference file:line it restates):

* ``rank_oracle``      — sklearn ``cosine_similarity`` + ``np.argsort`` top-k
                         (Tool/rank_chunks_optimized.py:215-216,225-235)
* ``simmatrix_oracle`` — ``create_similarity_matrix`` / ``analyze_similarity_distribution``
                         (Method/semantic_common.py:144-191,250-270)
* ``grouping_oracle``  — sharpen, centrality, quantile thresholds, kNN graph
                         (Method/Semantic_Grouping_Optimized.py:100-115,270-283,343-360)
* ``splitter_oracle``  — normalise, adjacent similarity, robust stats, C99 rank matrix,
                         P95 breakpoints (Method/Semantic_Splitter_Optimized.py:140-192,340-356,412-437)
"""
