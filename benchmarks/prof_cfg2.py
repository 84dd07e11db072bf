#!/usr/bin/env python
"""One cfg2-shaped run of K3 (+K4) for ncu: python benchmarks/prof_cfg2.py [--docs D] [--group]"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticsearch_b200 import ragged  # noqa: E402
from benchmarks.bench_configs import topic_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=10000)
ap.add_argument("--group", action="store_true")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
sizes = np.random.default_rng(3).integers(16, 513, size=a.docs)
E = topic_rows(sizes, 768, 4, "cuda")
plan = ragged.make_plan(sizes, "cuda")
S = torch.empty(plan.total_s, dtype=torch.float32, device="cuda")
for _ in range(a.iters):
    ragged.segmented_simmatrix(E, plan, out=S)
    if a.group:
        ragged.group_threshold_pass(S, plan, symmetric=True)
torch.cuda.synchronize()
print("ok", plan.total_rows, plan.total_s)
