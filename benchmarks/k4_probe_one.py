"""One grouping-threshold-pass run on documents of lo..hi sentences (ncu driver): python benchmarks/k4_probe_one.py LO HI DOCS"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticsearch_b200 import ragged
lo, hi, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sizes = np.random.default_rng(0).integers(lo, hi + 1, size=D)
plan = ragged.make_plan(sizes, "cuda")
S = torch.rand(plan.total_s, device="cuda")
for _ in range(3):
    ragged.group_threshold_pass(S, plan)
torch.cuda.synchronize()
print("ok")
