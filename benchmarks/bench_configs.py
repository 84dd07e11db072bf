#!/usr/bin/env python
"""Secondary benchmarks for BASELINE.json configs 1, 2, 3 and 5 (bench.py covers config 4).

    python benchmarks/bench_configs.py --config 1|2|3|5 [--steps K] [--docs D]

Prints one JSON line per config: throughput, CUDA-event kernel times, fraction of the measured HBM
roofline on the algorithmic bytes of SURVEY.md section 8(d), and the reference's own CPU expression
timed on a bounded sample of the same synthetic workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from semanticsearch_b200 import ragged, similarity  # noqa: E402


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))


def my_documents(sizes, power):
    """Static LPT partition of the documents over the ranks (no cross-GPU traffic)."""
    if WORLD == 1:
        return np.asarray(sizes)
    from semanticsearch_b200.sharded import partition_documents
    part = partition_documents(sizes, WORLD, power=power)[RANK]
    return np.asarray(sizes)[part]


def max_over_ranks(ms):
    if WORLD == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cuda_time(fn, steps, warmup=3, after=None):
    """Mean CUDA-event time of `fn` on the current stream; `after` (optional) joins side streams before the stop event."""
    for _ in range(warmup):
        fn()
    if after:
        after()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    if after:
        after()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def topic_rows(sizes, d, seed, device):
    """SURVEY.md 8(d) cfg 2/3 generator: one topic centroid per ~12 sentences + 0.7 * noise."""
    g = torch.Generator(device=device).manual_seed(seed)
    sizes_t = torch.as_tensor(sizes, device=device)
    total = int(sizes_t.sum())
    doc_of = torch.repeat_interleave(torch.arange(len(sizes), device=device), sizes_t)
    start = torch.cumsum(sizes_t, 0) - sizes_t
    local = torch.arange(total, device=device) - start[doc_of]
    topics_per_doc = (sizes_t + 11) // 12
    topic_base = torch.cumsum(topics_per_doc, 0) - topics_per_doc
    topic = topic_base[doc_of] + local // 12
    cent = torch.randn((int(topics_per_doc.sum()), d), generator=g, device=device)
    out = torch.empty((total, d), dtype=torch.float32, device=device)
    step = 1 << 20
    for a in range(0, total, step):
        b = min(total, a + step)
        out[a:b] = cent[topic[a:b]] + 0.7 * torch.randn((b - a, d), generator=g, device=device)
    return out


def config1(args):
    C = torch.from_numpy(np.random.default_rng(1).standard_normal((10000, 384)).astype(np.float32)).cuda()
    Q = torch.from_numpy(np.random.default_rng(2).standard_normal((100, 384)).astype(np.float32)).cuda()
    ms = cuda_time(lambda: similarity.cosine_topk(C, Q, 10), args.steps)
    from sklearn.metrics.pairwise import cosine_similarity
    Ch, Qh = C.cpu().numpy(), Q.cpu().numpy()
    t0 = time.perf_counter()
    for b in range(100):
        np.argsort(-cosine_similarity(Qh[b].reshape(1, -1), Ch)[0])[:10]
    cpu_s = time.perf_counter() - t0
    return {"config": "cfg1: 100 queries x 10k chunks x 384 fp32, top-10", "metric": "queries/s", "value": 100 / (ms * 1e-3),
            "ms": ms, "bound": "launch/L2 (15.5 MB working set)", "cpu_baseline": {"value": 100 / cpu_s, "unit": "queries/s",
            "cores": len(os.sched_getaffinity(0)), "kind": "port", "sample": "full config, rank_chunks_optimized.py:215-216,225 per query"}}


def config2(args):
    rng = np.random.default_rng(3)
    if args.realistic:  # the reference corpus's document lengths (median 10, mean 37 sentences): many tiny documents
        D = args.docs or 200_000
        sizes_all = np.clip(np.rint(rng.lognormal(np.log(10.0), 1.618, size=D)), 2, 512).astype(np.int64)
    else:
        D = args.docs or 10000
        sizes_all = rng.integers(16, 513, size=D)
    sizes = my_documents(sizes_all, power=2)
    E = topic_rows(sizes, 768, 4 + RANK, "cuda")
    plan = ragged.make_plan(sizes, "cuda")
    S = torch.empty(plan.total_s, dtype=torch.float32, device="cuda")
    ms_sim = max_over_ranks(cuda_time(lambda: ragged.segmented_simmatrix(E, plan, out=S), args.steps))
    ms_grp = max_over_ranks(cuda_time(lambda: ragged.group_threshold_pass(S, plan, symmetric=True), max(1, args.steps // 2), warmup=1))
    peak, src = hbm_peak()
    alg = 4 * 768 * plan.total_rows + 4 * plan.total_s
    # K4 algorithmic traffic: read S, write sim_sharp, write the neighbour lists (33 x (int32 + fp32) per row) + centrality
    alg_grp = 8 * plan.total_s + plan.total_rows * (33 * 8 + 8)
    tf32_peak = None
    try:
        tf32_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]) / 2.0
    except Exception:
        pass
    # CPU: the reference's arithmetic for the same pass, on a sample of documents
    from oracle import grouping_oracle as go, simmatrix_oracle as so
    sample = list(range(0, len(sizes), max(1, len(sizes) // 40)))[:40]
    Eh = [E[plan.offsets[d]:plan.offsets[d + 1]].cpu().numpy() for d in sample]
    t0 = time.perf_counter()
    for e in Eh:
        go.grouping_pass_ref(so.similarity_matrix_ref(e))
    cpu_s = (time.perf_counter() - t0) / len(sample)
    shape = "clipped log-normal lengths (median 10, mean 37, [2,512])" if args.realistic else "n~U[16,512]"
    return {"config": f"cfg2: {D} docs, {shape}, 768-d fp32: S = En En^T + grouping threshold pass", "metric": "docs/s",
            "value": D / ((ms_sim + ms_grp) * 1e-3), "ms_simmatrix": ms_sim, "ms_group_pass": ms_grp,
            "rows": plan.total_rows, "sum_n2": plan.total_s,
            "roofline": {"bound": "hbm", "kernel": "segmented_simmatrix_tc_kernel (tcgen05 kind::tf32, 3xTF32)",
                         "achieved": alg / (ms_sim * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": alg / (ms_sim * 1e-3) / 1e9 / peak, "peak_source": src,
                         "algorithmic_bytes_per_launch": alg, "fp32_equiv_tflops": 2 * 768 * plan.total_s / (ms_sim * 1e-3) / 1e12,
                         "tf32_mma_tflops_issued": 3 * 2 * 768 * plan.total_s / 2 / (ms_sim * 1e-3) / 1e12,
                         "tf32_peak_tflops_derived": tf32_peak,
                         "note": "fp32-parity products need 3 TF32 MMAs each (upper-triangular tiles only): the kernel is "
                                 "shared-memory-bandwidth bound (operand split + SS-mode MMA reads), not HBM bound"},
            "roofline_group_pass": {"bound": "hbm", "kernel": "group_threshold_kernel", "achieved": alg_grp / (ms_grp * 1e-3) / 1e9,
                                    "peak": peak, "unit": "GB/s", "frac": alg_grp / (ms_grp * 1e-3) / 1e9 / peak,
                                    "algorithmic_bytes_per_launch": alg_grp,
                                    "note": "instruction-issue bound (per-row selection), see profiles/"},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "docs/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                             "sample": f"{len(sample)} documents: numpy S (semantic_common.py:158-191) + sharpen/quantile/kNN "
                                       f"(Semantic_Grouping_Optimized.py:100-115,270-283,343-360)"}}


def config3(args):
    rng = np.random.default_rng(5)
    if args.realistic:
        # document lengths of the reference's corpus (document_length_summary.json: median 10,
        # mean 37 sentences) as a clipped log-normal; the whole 1 M-document set fits one GPU (~54 GB of fp32 rows)
        D = (args.docs or 1_000_000) * WORLD
        sizes_all = np.clip(np.rint(rng.lognormal(np.log(10.0), 1.618, size=D)), 2, 512).astype(np.int64)
    else:
        D = (args.docs or 50000) * WORLD  # weak scaling: 50k documents per GPU
        sizes_all = rng.integers(16, 513, size=D)
    sizes = my_documents(sizes_all, power=1)
    E = topic_rows(sizes, 384, 5 + RANK, "cuda")
    plan = ragged.make_plan(sizes, "cuda")
    adj_holder = {}

    def step():
        adj = ragged.adjacent_cosine(E)
        adj_holder["out"] = ragged.segmented_percentile(adj, plan, 95.0, want_stats=False)

    ms_adj = max_over_ranks(cuda_time(lambda: ragged.adjacent_cosine(E), args.steps))
    ms_all = max_over_ranks(cuda_time(step, args.steps))
    peak, src = hbm_peak()
    alg = 4 * 384 * plan.total_rows + 4 * plan.total_rows + 8 * D
    from oracle import splitter_oracle as spo
    sample = list(range(0, len(sizes), max(1, len(sizes) // 200)))[:200]
    Eh = [E[plan.offsets[d]:plan.offsets[d + 1]].cpu().numpy() for d in sample]
    t0 = time.perf_counter()
    for e in Eh:
        spo.p95_breakpoints_ref(spo.adjacent_sims_ref(e))
    cpu_s = (time.perf_counter() - t0) / len(sample)
    shape = "clipped log-normal lengths (median 10, mean 37, [2,512])" if args.realistic else "n~U[16,512]"
    return {"config": f"cfg3: adjacent-sentence distance + P95 breakpoints, {D}-doc batch of {shape} x 384 fp32 "
                      f"(1M docs = {max(1, 1_000_000 // D)} such batch(es))", "metric": "docs/s", "value": D / (ms_all * 1e-3),
            "ms_adjacent": ms_adj, "ms_total": ms_all, "rows": plan.total_rows,
            "roofline": {"bound": "hbm", "kernel": "adjacent_cosine_kernel", "achieved": alg / (ms_adj * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (ms_adj * 1e-3) / 1e9 / peak, "peak_source": src,
                         "algorithmic_bytes_per_launch": alg},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "docs/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                             "sample": f"{len(sample)} documents: _embed normalise + adjacent dot loop (Semantic_Splitter_Optimized.py:"
                                       f"140-152,412) + np.percentile(1-adj, 95)"}}


def config5(args):
    n, d, b, k = args.rows or 12_500_000, 384, 16, 100
    g = torch.Generator(device="cuda").manual_seed(9)
    C = torch.empty((n, d), dtype=torch.float16, device="cuda")
    for a in range(0, n, 1 << 21):
        e = min(n, a + (1 << 21))
        C[a:e] = torch.randn((e - a, d), generator=g, device="cuda").half()
    Q = torch.randn((b, d), generator=torch.Generator(device="cuda").manual_seed(10), device="cuda").half()
    ms = cuda_time(lambda: similarity.cosine_topk(C, Q, k), args.steps)
    peak, src = hbm_peak()
    alg = 2 * n * d
    return {"config": f"cfg5 (one of 8 shards): top-100 over {n} x 384 fp16, 16-query batch", "metric": "queries/s per shard-GPU",
            "value": b / (ms * 1e-3), "ms": ms,
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "peak_source": src, "algorithmic_bytes_per_launch": alg}}


def config6(args):
    """Not a BASELINE.json config: the per-query body of rank_and_filter (cosine, two rank lookups, RRF, fused order,
    percentile thresholds) for a block of query groups in one launch (K8)."""
    rng = np.random.default_rng(11)
    G = args.docs or 2000
    sizes = rng.integers(50, 1001, size=G)
    rows = int(sizes.sum())
    C = torch.randn((rows, 768), generator=torch.Generator(device="cuda").manual_seed(12), device="cuda")
    Q = torch.randn((G, 768), generator=torch.Generator(device="cuda").manual_seed(13), device="cuda")
    bm = torch.rand(rows, generator=torch.Generator(device="cuda").manual_seed(14), device="cuda")
    off = np.zeros(G + 1, dtype=np.int32)
    off[1:] = np.cumsum(sizes)
    off_d = torch.from_numpy(off).cuda()
    ms = cuda_time(lambda: similarity.segmented_rank_rrf(C, off_d, Q, bm, max_group_rows=int(sizes.max())), args.steps)
    peak, src = hbm_peak()
    alg = 4 * 768 * rows + 4 * rows + rows * (4 + 4 + 4 + 8 + 4)
    from sklearn.metrics.pairwise import cosine_similarity
    sample = list(range(0, G, max(1, G // 50)))[:50]
    Ch = [C[off[g]:off[g + 1]].cpu().numpy() for g in sample]
    Qh = [Q[g:g + 1].cpu().numpy() for g in sample]
    Bh = [bm[off[g]:off[g + 1]].cpu().numpy().astype(np.float64) for g in sample]
    t0 = time.perf_counter()
    for c, q, b in zip(Ch, Qh, Bh):
        cs = cosine_similarity(q, c)[0]
        n = len(cs)
        rc = np.empty(n); rc[np.argsort(-cs)] = np.arange(1, n + 1)
        rb = np.empty(n); rb[np.argsort(-b)] = np.arange(1, n + 1)
        rrf = 1.0 / (60 + rc) + 1.0 / (60 + rb)
        np.argsort(-rrf); np.percentile(rrf, 80); np.percentile(rrf, 20)
    cpu_s = (time.perf_counter() - t0) / len(sample)
    return {"config": f"rank groups (8f-3): {G} query groups, n~U[50,1000] chunks x 768 fp32: cosine + ranks + RRF + order + P80/P20",
            "metric": "groups/s", "value": G / (ms * 1e-3), "ms": ms, "rows": rows,
            "roofline": {"bound": "hbm", "kernel": "segmented_rank_rrf_kernel", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "peak_source": src, "algorithmic_bytes_per_launch": alg},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "groups/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                             "sample": f"{len(sample)} groups: rank_chunks_optimized.py:215-250,518-519 without BM25 scoring"}}


def config7(args):
    """Not a BASELINE.json config: the C99 leg of the splitter — similarity matrix, global
    rank transform and the greedy divisive cut search (Semantic_Splitter_Optimized.py:169-238) for a batch of documents."""
    rng = np.random.default_rng(15)
    D = args.docs or 2000
    sizes = rng.integers(16, 513, size=D)
    E = topic_rows(sizes, 384, 16, "cuda")
    plan = ragged.make_plan(sizes, "cuda")
    mins = np.maximum(3, np.maximum(5, np.rint(sizes / 50.0))).astype(np.int32)  # c99_min_chunk of the reference (:449-453)
    S = torch.empty(plan.total_s, dtype=torch.float32, device="cuda")
    ms_sim = cuda_time(lambda: ragged.segmented_simmatrix(E, plan, out=S), args.steps)
    ms_rank = cuda_time(lambda: ragged.c99_rank_matrix(S, plan, symmetric=True), max(1, args.steps // 2), warmup=1)
    ms_rank_general = cuda_time(lambda: ragged.c99_rank_matrix(S, plan), max(1, args.steps // 2), warmup=1)
    R = ragged.c99_rank_matrix(S, plan, symmetric=True)
    ms_cut = cuda_time(lambda: ragged.c99_divisive_cuts(R, plan, mins), max(1, args.steps // 2), warmup=1)
    cuts, n_cuts, _ = ragged.c99_divisive_cuts(R, plan, mins)
    peak, src = hbm_peak()
    sat = int(((sizes.astype(np.int64) + 1) ** 2).sum())
    alg_cut = 4 * plan.total_s + 8 * sat  # read R once, write the float64 table once
    from oracle import splitter_oracle as spo
    sample = list(range(0, D, max(1, D // 6)))[:6]
    Eh = [E[plan.offsets[d]:plan.offsets[d + 1]].cpu().numpy() for d in sample]
    Eh = [(e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32) for e in Eh]
    t0 = time.perf_counter()
    for e, d in zip(Eh, sample):
        spo.c99_divisive_ref(spo.c99_global_rank_ref(spo.c99_similarity_ref(e)), int(mins[d]))
    cpu_s = (time.perf_counter() - t0) / len(sample)
    total = ms_sim + ms_rank + ms_cut
    return {"config": f"c99 (8f-1): {D} docs, n~U[16,512], 384-d fp32: S, global rank matrix, divisive cut search",
            "metric": "docs/s", "value": D / (total * 1e-3), "ms_simmatrix": ms_sim, "ms_rank": ms_rank, "ms_rank_without_symmetry": ms_rank_general, "ms_cuts": ms_cut,
            "rows": plan.total_rows, "sum_n2": plan.total_s, "mean_cuts_per_doc": float(n_cuts.float().mean().item()),
            "roofline": {"bound": "hbm", "kernel": "c99_divisive_kernel", "achieved": alg_cut / (ms_cut * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg_cut / (ms_cut * 1e-3) / 1e9 / peak, "peak_source": src,
                         "algorithmic_bytes_per_launch": alg_cut,
                         "note": "the search rounds after the table build are latency-bound (one block arg-max per cut)"},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "docs/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                             "sample": f"{len(sample)} documents (mean n = {float(np.mean([len(e) for e in Eh])):.0f}): "
                                       f"Semantic_Splitter_Optimized.py:169-238 as restated in oracle/splitter_oracle.py"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[1, 2, 3, 5, 6, 7])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--docs", type=int, default=0)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--realistic", action="store_true", help="config 3 with the reference corpus's document-length distribution, 1 M documents")
    args = ap.parse_args()
    if WORLD > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        if args.config not in (2, 3):
            raise SystemExit("multi-GPU runs of this script cover the ragged configs 2 and 3 (bench.py covers 4 and 5)")
    res = {1: config1, 2: config2, 3: config3, 5: config5, 6: config6, 7: config7}[args.config](args)
    res["data"] = "synthetic"
    res["n_gpus"] = WORLD
    if WORLD > 1:
        res["partition"] = "static LPT over documents, no data-path collective; time = max over ranks"
        dist.barrier()
    if RANK == 0:
        print(json.dumps(res))
    if WORLD > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
