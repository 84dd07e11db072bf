#!/usr/bin/env python
"""One C99-leg run (S, rank matrix in both modes, cut search) for ncu: python benchmarks/prof_c99.py [--docs D]"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticsearch_b200 import ragged  # noqa: E402
from benchmarks.bench_configs import topic_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=2000)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
sizes = np.random.default_rng(15).integers(16, 513, size=a.docs)
E = topic_rows(sizes, 384, 16, "cuda")
plan = ragged.make_plan(sizes, "cuda")
mins = np.maximum(3, np.maximum(5, np.rint(sizes / 50.0))).astype(np.int32)
for _ in range(a.iters):
    S = ragged.segmented_simmatrix(E, plan)
    Rl = ragged.c99_rank_matrix(S, plan, use_local_rank=True, mask_size=11)
    Rg = ragged.c99_rank_matrix(S, plan, symmetric=True)
    ragged.c99_divisive_cuts(Rg, plan, mins)
torch.cuda.synchronize()
print("ok", plan.total_rows, plan.total_s)
