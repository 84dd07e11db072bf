#!/usr/bin/env python
"""TEST INFRASTRUCTURE — times the UNMODIFIED reference's semantic_grouping_main (CPU, through oracle/ref_shim.py) on
the synthetic documents benchmarks/dropin_grouping.py uses, so the drop-in's per-document wall time has the reference's
beside it.  Runs in the build container only (needs /root/reference); writes one JSON line.

    python oracle/time_reference_grouping.py [--sizes 64,128,256,512] [--dim 768] [--repeat 3]
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.gen_golden import topic_doc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64,128,256,512")
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--repeat", type=int, default=3)
    a = ap.parse_args()
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(1)
    topic_doc(rng, 32, a.dim)   # same stream position as the drop-in script's warm-up document
    out = {"what": "reference semantic_grouping_main (CPU), wall ms per document (median)", "dim": a.dim,
           "host_cores": len(os.sched_getaffinity(0)), "rows": []}
    for n in [int(x) for x in a.sizes.split(",")]:
        E = topic_doc(rng, n, a.dim, sent_per_topic=12, noise=0.7)
        text, _ = ref_shim.make_doc(E, tag=f"t{n}")
        ts = []
        for _ in range(a.repeat):
            t0 = time.perf_counter()
            res = ref.group.semantic_grouping_main(text, f"doc{n}", "m", device="cpu", silent=True)
            ts.append((time.perf_counter() - t0) * 1e3)
        out["rows"].append({"n": n, "total_ms": statistics.median(ts), "clusters": len(res)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
