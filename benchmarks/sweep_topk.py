#!/usr/bin/env python
"""Batch-size sweep of the three cosine top-k kernels (K1 stream, K7 tcstream, K2 gemm) on one GPU.

    python benchmarks/sweep_topk.py [--rows N] [--dim D] [--dtype bf16|fp16] [--k K] [--batches 1,2,4,...]

One JSON line per (batch, algo): ms per search, queries/s, and the corpus bytes streamed per second
(2*N*d / time — above the HBM peak means the corpus was served to several query groups per read).
Used to place the dispatch thresholds in semanticsearch_b200/similarity.py.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from semanticsearch_b200 import similarity  # noqa: E402


def cuda_time(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256")
    ap.add_argument("--algos", default="stream,tcstream,gemm")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    g = torch.Generator(device="cuda").manual_seed(6)
    C = torch.empty((args.rows, args.dim), dtype=dt, device="cuda")
    for a in range(0, args.rows, 1 << 20):
        e = min(args.rows, a + (1 << 20))
        C[a:e] = torch.randn((e - a, args.dim), generator=g, device="cuda").to(dt)
    for b in [int(x) for x in args.batches.split(",")]:
        Q = torch.randn((b, args.dim), generator=g, device="cuda").to(dt)
        ref = None
        for algo in args.algos.split(","):
            if algo == "gemm" and args.k > 16:
                continue
            if algo == "stream" and b > 64:
                continue
            try:
                ms = cuda_time(lambda: similarity.cosine_topk(C, Q, args.k, algo=algo), args.steps)
                s, i = similarity.cosine_topk(C, Q, args.k, algo=algo)
            except Exception as exc:  # noqa: BLE001
                print(json.dumps({"batch": b, "algo": algo, "error": str(exc)[:200]}), flush=True)
                continue
            agree = None
            if ref is None:
                ref = i
            else:
                agree = float((ref == i).float().mean().item())
            print(json.dumps({"batch": b, "algo": algo, "k": args.k, "rows": args.rows, "dim": args.dim, "dtype": args.dtype,
                              "ms": round(ms, 4), "qps": round(b / (ms * 1e-3), 1),
                              "corpus_GBps": round(2 * args.rows * args.dim / (ms * 1e-3) / 1e9, 1),
                              "idx_agree_with_first_algo": agree}), flush=True)


if __name__ == "__main__":
    main()
