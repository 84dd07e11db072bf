#!/usr/bin/env python
"""Interleaved A/B timing of K3 (and K4) variants: every library under benchmarks/_variants named on the command line is
loaded side by side and the launches alternate, so clock drift hits all variants alike.  Prints the median per variant."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticsearch_b200 import ragged, _lib  # noqa: E402
from benchmarks.bench_configs import topic_rows  # noqa: E402

names = sys.argv[1:]
libs = {}
for n in names:
    lib = ctypes.CDLL(os.path.join(ROOT, "benchmarks", "_variants", f"lib_{n.split('@')[0]}.so"))
    for fname, (restype, argtypes) in _lib._SIGNATURES.items():
        fn = getattr(lib, fname, None)
        if fn is not None:
            fn.restype, fn.argtypes = restype, argtypes
    libs[n] = lib
mix = os.environ.get("AB_MIX", "cfg2")
rng = np.random.default_rng(3)
if mix == "cfg2":
    sizes = rng.integers(16, 513, size=10000)
else:
    sizes = np.clip(np.rint(rng.lognormal(np.log(10.0), 1.618, size=200_000)), 2, 512).astype(np.int64)
E = topic_rows(sizes, 768, 4, "cuda")
plan = ragged.make_plan(sizes, "cuda")
units = ragged._units128(plan, E.device)
S = torch.empty(plan.total_s, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def k3(lib):
    rc = lib.ss_segmented_simmatrix_tc(E.data_ptr(), E.shape[0], E.shape[1], plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(),
                                       units.data_ptr(), units.shape[0], S.data_ptr(), None, st)
    assert rc == 0, rc


k3(libs[names[0]])
outs = {k: torch.empty(m, dtype=dt, device="cuda") for k, (m, dt) in dict(sharp=(plan.total_s, torch.float32), cent=(plan.total_rows, torch.float64),
        stats=(plan.n_docs * 8, torch.float64), kidx=(plan.total_rows * 33, torch.int32), kval=(plan.total_rows * 33, torch.float32)).items()}


def k4(lib):
    rc = lib.ss_group_threshold_pass(S.data_ptr(), plan.offsets_d.data_ptr(), plan.s_offsets_d.data_ptr(), plan.n_docs, 0.15, 0, 1,
                                     outs["sharp"].data_ptr(), outs["cent"].data_ptr(), outs["stats"].data_ptr(), outs["kidx"].data_ptr(),
                                     outs["kval"].data_ptr(), st)
    assert rc == 0, rc


if os.environ.get("AB_KERNEL", "k3") == "k4":
    k3 = k4
ref = None
times = {n: [] for n in names}
dbg_of = {}
for n in list(names):
    if "@" in n:
        dbg_of[n] = n.split("@")[1]
for rnd in range(int(os.environ.get("AB_ROUNDS", "12"))):
    for n in names:
        os.environ["SS_K3_DBG"] = dbg_of.get(n, "0")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k3(libs[n])
        e1.record()
        torch.cuda.synchronize()
        if rnd >= 2:
            times[n].append(e0.elapsed_time(e1))
        if rnd == 0:
            h = S[:: 9973].clone()
            if ref is None:
                ref = h
            else:
                print(f"  {n}: max |S - S_first_variant| on a sample = {float((h - ref).abs().max()):.3e}")
for n in names:
    t = np.array(times[n])
    print(f"{n}: K3 median {np.median(t):.3f} ms  min {t.min():.3f}  max {t.max():.3f}", flush=True)
