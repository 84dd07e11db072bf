"""bench.py's `extra` entries: BASELINE.json configs 1, 2, 3, 5 and the two §8f legs (C99, group ranking), each with

  value       whole-job throughput with the inputs resident in HBM (CUDA events on the launching stream),
  roofline    algorithmic bytes / dominant-kernel time against the measured HBM peak,
  e2e         the same metric through the host-buffer operator (pinned host inputs, H2D + kernels + D2H of the results and a
              stream synchronise inside every timed step),
  cpu_baseline the reference's own expression on a bounded sample of the same synthetic workload, host cores stated.

Only bench.py's `cpu_baseline` legs import oracle/ (allowed: it is the checker / baseline, never the measured path)."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from benchmarks.bench_configs import cuda_time, hbm_peak, topic_rows  # noqa: E402
from semanticsearch_b200 import ragged, similarity  # noqa: E402


def _cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _wall(fn, steps, warmup=1):
    """Mean wall time (s) of `fn`, which synchronises its own stream before returning (end-to-end legs)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


def _roof(kernel, alg_bytes, ms, traffic=None, note=None):
    peak, src = hbm_peak()
    r = {"bound": "hbm", "kernel": kernel, "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
         "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak, "peak_source": src, "algorithmic_bytes_per_launch": alg_bytes,
         "kernel_ms": ms, "traffic": traffic}
    if note:
        r["note"] = note
    return r


def _pinned(t: torch.Tensor) -> torch.Tensor:
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t)
    return out


# ------------------------------------------------------------------------------------------------------------------
def cfg1(steps=20, cpu=True):
    C = torch.from_numpy(np.random.default_rng(1).standard_normal((10000, 384)).astype(np.float32))
    Q = torch.from_numpy(np.random.default_rng(2).standard_normal((100, 384)).astype(np.float32))
    Cd, Qd = C.cuda(), Q.cuda()
    ms = cuda_time(lambda: similarity.cosine_topk(Cd, Qd, 10), steps)
    Cp, Qp = _pinned(C), _pinned(Q)
    hs = torch.empty((100, 10), dtype=torch.float32, pin_memory=True)
    hi = torch.empty((100, 10), dtype=torch.int64, pin_memory=True)

    def e2e():   # the reference passes the chunk matrix with every call (rank:199,216): chunks and queries both cross the link
        Cd.copy_(Cp, non_blocking=True)          # into the operator's resident buffers: no allocation inside the timed call
        Qd.copy_(Qp, non_blocking=True)
        s, i = similarity.cosine_topk(Cd, Qd, 10)
        hs.copy_(s, non_blocking=True)
        hi.copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_s = _wall(e2e, steps, warmup=5)
    alg = 4 * (10000 * 384 + 100 * 384) + 100 * 10 * 8
    out = {"workload": "cfg1: 100 queries x 10k chunks x 384 fp32, top-10 cosine (the reference's own scale)",
           "metric": "queries/s", "unit": "queries/s", "value": 100 / (ms * 1e-3), "ms_per_step": ms, "dtype": "f32",
           "e2e": {"value": 100 / e2e_s, "unit": "queries/s", "ms_per_step": e2e_s * 1e3,
                   "h2d_bytes_per_step": 4 * (10000 * 384 + 100 * 384), "d2h_bytes_per_step": 100 * 10 * 12,
                   "api": "similarity.cosine_topk on pinned host chunks + queries"},
           "roofline": _roof("small_scores_kernel + small_select_kernel", alg, ms,
                             note="15.5 MB working set: launch / L2-bound, the fraction is informational (SURVEY.md 8d)"),
           "gpu_launches": steps * 3}
    if cpu:
        from sklearn.metrics.pairwise import cosine_similarity
        Ch, Qh = C.numpy(), Q.numpy()
        t0 = time.perf_counter()
        for b in range(100):
            np.argsort(-cosine_similarity(Qh[b].reshape(1, -1), Ch)[0])[:10]
        cpu_s = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 100 / cpu_s, "unit": "queries/s", "cores": _cores(), "kind": "port",
                               "sample": "the full config, one sklearn cosine_similarity + np.argsort per query (rank_chunks_optimized.py:215-216,225)"}
    return out


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the two config-2 kernels from the final-code ncu captures
CFG2_TRAFFIC = {"k3": (20.030982e9 + 3.766918e9, "profiles/r02_k3_final_ncu_summary.txt"),
                "k4": (8.571646e9 + 4.481500e9 + 0.117e9, "profiles/r02_cfg2_final_ncu_summary.txt")}


def cfg2(steps=3, docs=10000, cpu=True, traffic=None):
    rng = np.random.default_rng(3)
    sizes = rng.integers(16, 513, size=docs)
    E = topic_rows(sizes, 768, 4, "cuda")
    plan = ragged.make_plan(sizes, "cuda")
    S = torch.empty(plan.total_s, dtype=torch.float32, device="cuda")
    ms_sim = cuda_time(lambda: ragged.segmented_simmatrix(E, plan, out=S), steps)
    ms_grp = cuda_time(lambda: ragged.group_threshold_pass(S, plan, symmetric=True), steps, warmup=1)
    alg_sim = 4 * 768 * plan.total_rows + 4 * plan.total_s
    alg_grp = 8 * plan.total_s + plan.total_rows * (33 * 8 + 8)
    out = {"workload": f"cfg2: {docs} docs, n~U[16,512], 768-d fp32: S = En En^T + grouping threshold pass",
           "metric": "docs/s", "unit": "docs/s", "value": docs / ((ms_sim + ms_grp) * 1e-3), "ms_per_step": ms_sim + ms_grp,
           "ms_simmatrix": ms_sim, "ms_group_pass": ms_grp, "rows": plan.total_rows, "sum_n2": plan.total_s, "dtype": "f32 (3xTF32 products)",
           "roofline": _roof("segmented_simmatrix_tc_kernel (tcgen05: kind::tf32 main term + kind::f16 cross terms)", alg_sim, ms_sim,
                             traffic=CFG2_TRAFFIC["k3"][0] if docs == 10000 else None,
                             note="traffic from " + CFG2_TRAFFIC["k3"][1] + "; shared-memory / epilogue bound, not HBM bound"),
           "roofline_group_pass": _roof("group_threshold_kernel", alg_grp, ms_grp, traffic=CFG2_TRAFFIC["k4"][0] if docs == 10000 else None,
                                        note="traffic from " + CFG2_TRAFFIC["k4"][1] + "; instruction-issue bound (per-row selection)"),
           "roofline_whole_pass": _roof("K3 + K4", alg_sim + 4 * plan.total_s, ms_sim + ms_grp,
                                        note="SURVEY.md 8d floor: read E once, write S and sim_sharp once"),
           "gpu_launches": steps * 4}
    # end to end: embeddings in pinned host memory -> S, sim_sharp, centrality, thresholds, neighbour lists in pinned host memory
    Eh = _pinned(E.cpu())
    del E, S
    torch.cuda.empty_cache()
    holder = {}

    def e2e():
        holder["out"] = ragged.grouping_pass_host(Eh, sizes, out=holder.get("out"), plan=plan)

    e2e_s = _wall(e2e, max(2, steps), warmup=1)
    d2h = sum(v.numel() * v.element_size() for v in holder["out"].values())
    out["e2e"] = {"value": docs / e2e_s, "unit": "docs/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": Eh.numel() * 4,
                  "d2h_bytes_per_step": d2h, "api": "ragged.grouping_pass_host (pinned host embeddings in, pinned host results out)",
                  "note": "host-link bound: the step moves %.1f GB in and %.1f GB out; document chunks are pipelined so that both "
                          "directions of the link work at once (%.0f GB/s combined)" % (Eh.numel() * 4 / 1e9, d2h / 1e9,
                                                                                       (Eh.numel() * 4 + d2h) / 1e9 / e2e_s)}
    if cpu:
        from oracle import grouping_oracle as go, simmatrix_oracle as so
        sample = list(range(0, docs, max(1, docs // 40)))[:40]
        Es = [Eh[plan.offsets[d]:plan.offsets[d + 1]].numpy() for d in sample]
        t0 = time.perf_counter()
        for e in Es:
            go.grouping_pass_ref(so.similarity_matrix_ref(e))
        cpu_s = (time.perf_counter() - t0) / len(sample)
        out["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "docs/s", "cores": _cores(), "kind": "port",
                               "sample": f"{len(sample)} documents: numpy S (semantic_common.py:158-191) + sharpen / quantiles / kNN "
                                         f"(Semantic_Grouping_Optimized.py:100-115,270-283,343-360)"}
    del holder, Eh
    torch.cuda.empty_cache()
    return out


def cfg3(steps=3, batch_docs=50000, total_docs=1_000_000, chunk_docs=5000, cpu=True):
    rng = np.random.default_rng(5)
    sizes = rng.integers(16, 513, size=batch_docs)
    E = topic_rows(sizes, 384, 5, "cuda")
    plan = ragged.make_plan(sizes, "cuda")

    def step():
        adj = ragged.adjacent_cosine(E)
        ragged.segmented_percentile(adj, plan, 95.0, want_stats=False)

    ms_adj = cuda_time(lambda: ragged.adjacent_cosine(E), steps)
    ms_all = cuda_time(step, steps)
    alg = 4 * 384 * plan.total_rows + 4 * plan.total_rows + 8 * batch_docs
    out = {"workload": f"cfg3: adjacent-sentence distance + P95 breakpoints, n~U[16,512] x 384 fp32; resident value on a {batch_docs}-document "
                       f"batch, e2e streams {total_docs} documents from pinned host memory",
           "metric": "docs/s", "unit": "docs/s", "value": batch_docs / (ms_all * 1e-3), "ms_per_step": ms_all, "ms_adjacent": ms_adj,
           "rows": plan.total_rows, "dtype": "f32", "roofline": _roof("adjacent_cosine_kernel", alg, ms_adj), "gpu_launches": steps * 2}
    cpu_sample = None
    if cpu:
        take = list(range(0, batch_docs, max(1, batch_docs // 200)))[:200]
        cpu_sample = [E[plan.offsets[d]:plan.offsets[d + 1]].cpu().numpy() for d in take]
    del E
    torch.cuda.empty_cache()
    # end to end at corpus scale: chunk_docs-document batches stream from pinned host buffers through the double-buffered
    # pipeline; the corpus is synthetic, so a ring of 4 distinct pinned chunks is cycled (every byte still crosses the link)
    ring = []
    for c in range(4):
        sz = np.random.default_rng(50 + c).integers(16, 513, size=chunk_docs)
        ring.append((_pinned(topic_rows(sz, 384, 60 + c, "cuda").cpu()), sz, ragged.make_plan(sz, "cuda")))
    torch.cuda.empty_cache()
    max_rows = max(int(s.sum()) for _, s, _ in ring)
    stream = ragged.SplitterStream(max_rows, 384, chunk_docs)
    n_chunks = max(1, total_docs // chunk_docs)
    for i in range(4):                                           # warm-up
        stream.submit(*ring[i % 4])
    stream.drain()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h2d = d2h = 0
    flagged = 0
    for i in range(n_chunks):
        Eh, sz, pl = ring[i % 4]
        res = stream.submit(Eh, sz, pl)
        h2d += pl.total_rows * 384 * 4
        d2h += pl.total_rows + 8 * pl.n_docs
        if res is not None:
            flagged += int(res["flags"][:64].sum())              # the caller reads every batch's result
    for res in stream.drain():
        flagged += int(res["flags"][:64].sum())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    out["e2e"] = {"value": n_chunks * chunk_docs / e2e_s, "unit": "docs/s", "ms_per_step": e2e_s * 1e3 / n_chunks,
                  "documents": n_chunks * chunk_docs, "seconds": e2e_s, "h2d_bytes_per_step": h2d // n_chunks,
                  "d2h_bytes_per_step": d2h // n_chunks, "host_link_GBps": h2d / e2e_s / 1e9,
                  "api": "ragged.SplitterStream (double-buffered pinned-chunk H2D pipeline)",
                  "note": "host-link bound: %.0f GB of embeddings cross PCIe for %d documents" % (h2d / 1e9, n_chunks * chunk_docs)}
    if cpu:
        from oracle import splitter_oracle as spo
        t0 = time.perf_counter()
        for e in cpu_sample:
            spo.p95_breakpoints_ref(spo.adjacent_sims_ref(e))
        cpu_s = (time.perf_counter() - t0) / len(cpu_sample)
        out["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "docs/s", "cores": _cores(), "kind": "port",
                               "sample": f"{len(cpu_sample)} documents: _embed normalise + adjacent dot loop "
                                         f"(Semantic_Splitter_Optimized.py:140-152,412) + np.percentile(1 - adj, 95)"}
    del ring, stream
    torch.cuda.empty_cache()
    return out


def cfg5_shard(steps=20, rows=12_500_000, cpu=True, traffic=None):
    d, b, k = 384, 16, 100
    g = torch.Generator(device="cuda").manual_seed(9)
    C = torch.empty((rows, d), dtype=torch.float16, device="cuda")
    for a in range(0, rows, 1 << 21):
        e = min(rows, a + (1 << 21))
        C[a:e] = torch.randn((e - a, d), generator=g, device="cuda").half()
    Q = torch.randn((b, d), generator=torch.Generator(device="cuda").manual_seed(10), device="cuda").half()
    ms = cuda_time(lambda: similarity.cosine_topk(C, Q, k), steps)
    from semanticsearch_b200.sharded import GraphedSearch, ShardedCorpus
    gs = GraphedSearch(ShardedCorpus(C, 0), b, k)
    Qp = _pinned(Q.cpu())
    hs = torch.empty((b, k), dtype=torch.float32, pin_memory=True)
    hi = torch.empty((b, k), dtype=torch.int64, pin_memory=True)

    def e2e():
        s, i = gs(Qp)
        hs.copy_(s, non_blocking=True)
        hi.copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_s = _wall(e2e, steps, warmup=3)
    alg = 2 * rows * d
    out = {"workload": f"cfg5 (one of 8 shards): top-100 cosine over {rows} x 384 fp16, 16-query batch",
           "metric": "queries/s per shard-GPU", "unit": "queries/s", "value": b / (ms * 1e-3), "ms_per_step": ms, "dtype": "f16 storage, f32 accumulate",
           "e2e": {"value": b / e2e_s, "unit": "queries/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": b * d * 2,
                   "d2h_bytes_per_step": b * k * 12, "api": "sharded.GraphedSearch (one CUDA-graph launch per search)"},
           "roofline": _roof("cosine_topk_tcstream_kernel (tcgen05)", alg, ms, traffic=traffic), "gpu_launches": steps * 2}
    if cpu:
        from sklearn.metrics.pairwise import cosine_similarity
        n_s = 250_000
        Ch = C[:n_s].float().cpu().numpy()
        Qh = Q.float().cpu().numpy()
        t0 = time.perf_counter()
        for q in range(4):
            np.argsort(-cosine_similarity(Qh[q].reshape(1, -1), Ch)[0])[:k]
        cpu_s = (time.perf_counter() - t0) / 4
        out["cpu_baseline"] = {"value": 1.0 / (cpu_s * rows / n_s), "unit": "queries/s", "cores": _cores(), "kind": "port",
                               "extrapolated": True, "sample_factor": n_s / rows,
                               "sample": f"4 queries x {n_s} rows (fp16-rounded values as fp32), sklearn cosine + argsort per query, "
                                         f"scaled x{rows / n_s:.0f} to the shard"}
    del C, gs
    torch.cuda.empty_cache()
    return out


def c99(steps=3, docs=2000, local=True, cpu=True):
    rng = np.random.default_rng(15)
    sizes = rng.integers(16, 513, size=docs)
    E = topic_rows(sizes, 384, 16, "cuda")
    plan = ragged.make_plan(sizes, "cuda")
    mins = np.maximum(3, np.maximum(5, np.rint(sizes / 50.0))).astype(np.int32)
    S = torch.empty(plan.total_s, dtype=torch.float32, device="cuda")
    ms_sim = cuda_time(lambda: ragged.segmented_simmatrix(E, plan, out=S), steps)
    rank = (lambda: ragged.c99_rank_matrix(S, plan, use_local_rank=True, mask_size=11)) if local else \
        (lambda: ragged.c99_rank_matrix(S, plan, symmetric=True))
    ms_rank = cuda_time(rank, steps, warmup=1)
    R = rank()
    ms_cut = cuda_time(lambda: ragged.c99_divisive_cuts(R, plan, mins), steps, warmup=1)
    total = ms_sim + ms_rank + ms_cut
    mode = "local 11x11 rank (the controller's default preset, simple_chunk_controller.py:1451)" if local else "global rank"
    alg_rank = 8 * plan.total_s
    Eh = _pinned(E.cpu())

    def e2e():
        ragged.c99_cuts_host(Eh, sizes, mins, use_local_rank=local, plan=plan)

    e2e_s = _wall(e2e, steps, warmup=1)
    out = {"workload": f"c99 leg (8f-1), {mode}: {docs} docs, n~U[16,512], 384-d fp32: S, rank matrix, divisive cut search",
           "metric": "docs/s", "unit": "docs/s", "value": docs / (total * 1e-3), "ms_per_step": total, "ms_simmatrix": ms_sim,
           "ms_rank": ms_rank, "ms_cuts": ms_cut, "rows": plan.total_rows, "sum_n2": plan.total_s, "dtype": "f32 / f64 block sums",
           "e2e": {"value": docs / e2e_s, "unit": "docs/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": Eh.numel() * 4,
                   "d2h_bytes_per_step": 4 * (plan.total_rows + docs), "api": "ragged.c99_cuts_host"},
           "roofline": _roof("c99 rank kernel (" + ("c99_local_rank_kernel" if local else "c99_rank_rows_kernel + transpose-add") + ")",
                             alg_rank, ms_rank, note="read S once, write R once"),
           "gpu_launches": steps * 4}
    if cpu:
        from oracle import splitter_oracle as spo
        sample = list(range(0, docs, max(1, docs // 4)))[:4]
        Es = [Eh[plan.offsets[d]:plan.offsets[d + 1]].numpy() for d in sample]
        Es = [(e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32) for e in Es]
        rank_ref = (lambda s: spo.c99_local_rank_ref(s, 11)) if local else spo.c99_global_rank_ref
        t0 = time.perf_counter()
        for e, d in zip(Es, sample):
            spo.c99_divisive_ref(rank_ref(spo.c99_similarity_ref(e)), int(mins[d]))
        cpu_s = (time.perf_counter() - t0) / len(sample)
        out["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "docs/s", "cores": _cores(), "kind": "port",
                               "sample": f"{len(sample)} documents (mean n = {float(np.mean([len(e) for e in Es])):.0f}): "
                                         f"Semantic_Splitter_Optimized.py:169-238 as restated in oracle/splitter_oracle.py"}
    del E, S, R, Eh
    torch.cuda.empty_cache()
    return out


def rank_groups(steps=5, groups=2000, cpu=True):
    rng = np.random.default_rng(11)
    sizes = rng.integers(50, 1001, size=groups)
    rows = int(sizes.sum())
    C = torch.randn((rows, 768), generator=torch.Generator(device="cuda").manual_seed(12), device="cuda")
    Q = torch.randn((groups, 768), generator=torch.Generator(device="cuda").manual_seed(13), device="cuda")
    bm = torch.rand(rows, generator=torch.Generator(device="cuda").manual_seed(14), device="cuda")
    off = np.zeros(groups + 1, dtype=np.int32)
    off[1:] = np.cumsum(sizes)
    off_d = torch.from_numpy(off).cuda()
    mx = int(sizes.max())
    ms = cuda_time(lambda: similarity.segmented_rank_rrf(C, off_d, Q, bm, max_group_rows=mx), steps)
    alg = 4 * 768 * rows + 4 * rows + rows * (4 + 4 + 4 + 8 + 4)
    Cp, Qp, Bp = _pinned(C.cpu()), _pinned(Q.cpu()), _pinned(bm.cpu())
    host = {}

    def e2e():
        res = similarity.segmented_rank_rrf(Cp.cuda(non_blocking=True), off_d, Qp.cuda(non_blocking=True), Bp.cuda(non_blocking=True),
                                            max_group_rows=mx)
        for key, v in res.items():
            if v is None:
                continue
            if key not in host:
                host[key] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            host[key].copy_(v, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_s = _wall(e2e, steps, warmup=1)
    d2h = sum(v.numel() * v.element_size() for v in host.values())
    out = {"workload": f"rank groups (8f-3): {groups} query groups, n~U[50,1000] chunks x 768 fp32: cosine + ranks + RRF + order + P80/P20",
           "metric": "groups/s", "unit": "groups/s", "value": groups / (ms * 1e-3), "ms_per_step": ms, "rows": rows, "dtype": "f32 / f64 fusion",
           "e2e": {"value": groups / e2e_s, "unit": "groups/s", "ms_per_step": e2e_s * 1e3,
                   "h2d_bytes_per_step": 4 * (rows * 768 + groups * 768 + rows), "d2h_bytes_per_step": d2h,
                   "api": "similarity.segmented_rank_rrf on pinned host chunks / queries / BM25 scores"},
           "roofline": _roof("segmented_rank_rrf_kernel", alg, ms), "gpu_launches": steps}
    if cpu:
        from sklearn.metrics.pairwise import cosine_similarity
        sample = list(range(0, groups, max(1, groups // 50)))[:50]
        Ch = [Cp[off[g]:off[g + 1]].numpy() for g in sample]
        Qh = [Qp[g:g + 1].numpy() for g in sample]
        Bh = [Bp[off[g]:off[g + 1]].numpy().astype(np.float64) for g in sample]
        t0 = time.perf_counter()
        for c, q, b in zip(Ch, Qh, Bh):
            cs = cosine_similarity(q, c)[0]
            n = len(cs)
            rc = np.empty(n); rc[np.argsort(-cs)] = np.arange(1, n + 1)
            rb = np.empty(n); rb[np.argsort(-b)] = np.arange(1, n + 1)
            rrf = 1.0 / (60 + rc) + 1.0 / (60 + rb)
            np.argsort(-rrf); np.percentile(rrf, 80); np.percentile(rrf, 20)
        cpu_s = (time.perf_counter() - t0) / len(sample)
        out["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "groups/s", "cores": _cores(), "kind": "port",
                               "sample": f"{len(sample)} groups: rank_chunks_optimized.py:215-250,518-519 without BM25 scoring"}
    del C, Q, bm
    torch.cuda.empty_cache()
    return out


ALL = {"cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg5_shard": cfg5_shard, "c99_local": lambda **kw: c99(local=True, **kw),
       "c99_global": lambda **kw: c99(local=False, **kw), "rank_groups": rank_groups}


def run_all(names=None, cpu=True, log=None):
    out = {}
    for name in (names or list(ALL)):
        t0 = time.perf_counter()
        try:
            out[name] = ALL[name](cpu=cpu)
        except Exception as exc:  # noqa: BLE001 — one failing leg must not take the headline line down
            out[name] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()
        out[name]["bench_seconds"] = time.perf_counter() - t0
        if log:
            log(f"[bench] extra {name}: {out[name].get('value')} {out[name].get('unit')} in {out[name]['bench_seconds']:.1f}s")
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    res = run_all([n for n in a.only.split(",") if n] or None, cpu=not a.no_cpu, log=lambda m: print(m, file=sys.stderr))
    print(json.dumps(res))
