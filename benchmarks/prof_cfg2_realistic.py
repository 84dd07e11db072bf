#!/usr/bin/env python
"""cfg2 with the reference corpus's length mix (clipped log-normal, median 10 sentences) for ncu launch lists."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticsearch_b200 import ragged  # noqa: E402
from benchmarks.bench_configs import topic_rows  # noqa: E402

rng = np.random.default_rng(3)
sizes = np.clip(np.rint(rng.lognormal(np.log(10.0), 1.618, size=200_000)), 2, 512).astype(np.int64)
E = topic_rows(sizes, 768, 4, "cuda")
plan = ragged.make_plan(sizes, "cuda")
S = torch.empty(plan.total_s, dtype=torch.float32, device="cuda")
for _ in range(3):
    ragged.segmented_simmatrix(E, plan, out=S)
    ragged.group_threshold_pass(S, plan, symmetric=True)
torch.cuda.synchronize()
print("ok", plan.total_rows, plan.total_s, int((sizes <= 32).sum()), int(((sizes > 32) & (sizes <= 128)).sum()), int((sizes > 128).sum()))
