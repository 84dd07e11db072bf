#!/usr/bin/env python
"""Small-shape pass over every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python benchmarks/sanitize_small.py
Shapes are tiny on purpose (the sanitizer slows kernels down ~100x) but hit the ragged tails: last
partial tiles, K tails, padded query groups, pair kernels with an odd number of query blocks."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticsearch_b200 import ragged, similarity  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)


def rnd(n, d, dt):
    return torch.randn((n, d), generator=g, device="cuda").to(dt)


for dt in (torch.bfloat16, torch.float16):
    C = rnd(5000 + 37, 200, dt)
    for b, k, algo in ((1, 10, "stream"), (5, 33, "stream"), (3, 10, "tcstream"), (17, 100, "tcstream"), (70, 5, "tcstream"),
                       (130, 10, "gemm"), (300, 16, "gemm"), (64, 4, "gemm")):
        s, i = similarity.cosine_topk(C, rnd(b, 200, dt), k, algo=algo)
        assert torch.all(s[:, 1:] <= s[:, :-1])
Cf = rnd(3000, 50, torch.float32)
similarity.cosine_topk(Cf, rnd(4, 50, torch.float32), 7)
sc = similarity.cosine_scores(Cf, rnd(2, 50, torch.float32))
similarity.rank_order(sc)
sizes = [1, 2, 17, 130, 257, 64, 0, 300]
E = rnd(sum(sizes), 100, torch.float32)
plan = ragged.make_plan(sizes, "cuda")
for algo in ("tc", "ffma"):
    S = ragged.segmented_simmatrix(E, plan, algo=algo)
S = ragged.segmented_simmatrix(E, plan, algo="tc", validate=True)
out = ragged.group_threshold_pass(S, plan)
out_sym = ragged.group_threshold_pass(S, plan, symmetric=True)
assert torch.equal(out["knn_idx"], out_sym["knn_idx"])
ragged.similarity_distribution(S, plan)
R = ragged.c99_rank_matrix(S, plan)
ragged.c99_rank_matrix(S, plan, symmetric=True)
for mask in (3, 7, 11, 15, 21):                      # tiled local-rank kernel (H = 1..7) and the counting fallback
    ragged.c99_rank_matrix(S, plan, use_local_rank=True, mask_size=mask)
ragged.c99_divisive_cuts(R, plan, 3, want_profile=True)
adj = ragged.adjacent_cosine(E)
ragged.segmented_percentile(adj, plan, 95.0)
# round-2 kernels: block sums / co-association, diameter split, long-document cut search, K2 with k = 16 and tiny corpora
groups = [[list(range(0, n // 2)), list(range(n // 2, n)), [0, 0] if n else []] for n in sizes]
ragged.group_block_sums(out["sim_sharp"], plan, groups)
ragged.group_coassociation(np.random.default_rng(0).integers(0, 4, size=(5, 77)))
ragged.diameter_split(S, plan, 0.4)
big = [2100, 5]
Eb = rnd(sum(big), 32, torch.float32)
pb = ragged.make_plan(big, "cuda")
Rb = ragged.c99_rank_matrix(ragged.segmented_simmatrix(Eb, pb), pb, symmetric=True)
ragged.c99_divisive_cuts(Rb, pb, [40, 2])
for n_small, b in ((300, 130), (6, 130), (257, 300)):
    similarity.cosine_topk(rnd(n_small, 64, torch.bfloat16), rnd(b, 64, torch.bfloat16), 16 if n_small > 16 else 10, algo="gemm")
torch.cuda.synchronize()
print("sanitize_small: ok")
