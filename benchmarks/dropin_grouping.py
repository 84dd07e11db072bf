#!/usr/bin/env python
"""Wall time per document of the grouping drop-in (semantic_grouping_main semantics) with its breakdown:

    python benchmarks/dropin_grouping.py [--sizes 64,128,256,512] [--dim 768] [--repeat 5]

For each n: synthetic topic-structured embeddings (SURVEY.md 8d generator), then per document
  device pass  = H2D of the embeddings + K3 similarity matrix + K4 threshold pass + D2H of what the host stage reads
  host stage   = kNN graph, spectral embedding (numpy eigh), seeded k-means, split / merge / refine / reassign, whose
                 block means come from ss_group_block_sums launches (count and time reported)
Prints one JSON line.  The reference's own time for the same documents is measured by oracle/time_reference_grouping.py
in the build container (the reference tree does not travel to the GPU box)."""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from semanticsearch_b200 import ragged  # noqa: E402
from semanticsearch_b200.Method import Semantic_Grouping_Optimized as G  # noqa: E402


def topic_doc(rng, n, d, sent_per_topic=12, noise=0.7):
    """Same generator (same random stream) as oracle/gen_golden.py:topic_doc; kept local, the bench never imports oracle/."""
    n_topics = max(1, int(np.ceil(n / sent_per_topic)))
    cent = rng.standard_normal((n_topics, d)).astype(np.float32)
    topic_of = np.minimum(np.arange(n) // sent_per_topic, n_topics - 1)
    return (cent[topic_of] + noise * rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64,128,256,512")
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--repeat", type=int, default=5)
    a = ap.parse_args()
    out = {"what": "grouping drop-in, wall ms per document (median)", "dim": a.dim, "host_cores": len(os.sched_getaffinity(0)), "rows": []}
    rng = np.random.default_rng(1)
    G.device_pass_batch([topic_doc(rng, 32, a.dim)])   # context, module load
    for n in [int(x) for x in a.sizes.split(",")]:
        E = topic_doc(rng, n, a.dim)
        t_dev, t_host, t_sums, calls, n_clusters = [], [], [], 0, 0
        for _ in range(a.repeat):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dp = G.device_pass_batch([E])[0]
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            sums_time = [0.0]
            inner = dp.block_sums

            def timed_sums(groups, _inner=inner, _acc=sums_time):
                s0 = time.perf_counter()
                r = _inner(groups)
                _acc[0] += time.perf_counter() - s0
                return r

            merged, method, _ = G.cluster_from_device_pass(dp, block_sums=timed_sums)
            t2 = time.perf_counter()
            t_dev.append((t1 - t0) * 1e3)
            t_host.append((t2 - t1) * 1e3)
            t_sums.append(sums_time[0] * 1e3)
            calls, n_clusters = inner.calls, len(merged)
        out["rows"].append({"n": n, "total_ms": statistics.median(t_dev) + statistics.median(t_host),
                            "device_pass_ms": statistics.median(t_dev), "host_stage_ms": statistics.median(t_host),
                            "block_sums_ms": statistics.median(t_sums), "block_sums_launches": calls, "method": method,
                            "clusters": n_clusters})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
