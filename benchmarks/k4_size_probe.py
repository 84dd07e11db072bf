#!/usr/bin/env python
"""Cost of the grouping threshold pass (K4) per document-size class: python benchmarks/k4_size_probe.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticsearch_b200 import ragged  # noqa: E402

def t(fn, steps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

rng = np.random.default_rng(0)
for lo, hi, D in ((2, 32, 200000), (33, 64, 100000), (65, 128, 50000), (129, 256, 20000), (257, 512, 8000)):
    sizes = rng.integers(lo, hi + 1, size=D)
    plan = ragged.make_plan(sizes, "cuda")
    S = torch.rand(plan.total_s, device="cuda")
    ms = t(lambda: ragged.group_threshold_pass(S, plan))
    print(f"n in [{lo},{hi}] docs={D} sum_n2={plan.total_s/1e6:.0f}M  {ms:.2f} ms  {ms*1e3/D*444:.1f} us per doc-slot  {plan.total_s/ms/1e6:.0f} G elem/s", flush=True)
