#!/usr/bin/env python
"""Run one cosine top-k configuration a few times (the command wrapped by ncu for profiles/).

    python benchmarks/prof_one.py --rows N --dim D --dtype bf16|fp16|fp32 --k K --batch B --algo auto|stream|tcstream|gemm --iters I
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from semanticsearch_b200 import similarity  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    g = torch.Generator(device="cuda").manual_seed(6)
    C = torch.empty((args.rows, args.dim), dtype=dt, device="cuda")
    for a in range(0, args.rows, 1 << 20):
        e = min(args.rows, a + (1 << 20))
        C[a:e] = torch.randn((e - a, args.dim), generator=g, device="cuda").to(dt)
    Q = torch.randn((args.batch, args.dim), generator=g, device="cuda").to(dt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(args.iters):
        if it == args.iters - 1:
            e0.record()
        s, i = similarity.cosine_topk(C, Q, args.k, algo=args.algo)
    e1.record()
    torch.cuda.synchronize()
    print(f"algo={similarity.choose_algo(C, Q, args.k) if args.algo == 'auto' else args.algo} last_iter_ms={e0.elapsed_time(e1):.4f} "
          f"top1={s[0, 0].item():.6f}@{i[0, 0].item()}")


if __name__ == "__main__":
    main()
