#!/usr/bin/env python
"""Blackwell evidence from the shipped library: per-kernel counts of the SASS opcodes that only tcgen05 / TMEM / TMA code
produces (profiles/sass_opcodes.txt).

    python profiles/sass_opcodes.py [path/to/libsemsearch_b200.so] > profiles/sass_opcodes.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops; HMMA (legacy mma.sync) must not appear."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "semanticsearch_b200", "libsemsearch_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCCP", "SYNCS", "FFMA2", "HMMA"]
arch = None
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        counts[cur]["arch:" + str(arch)] = 1
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for p in pats:
            if op.startswith(p):
                counts[cur][p + ("." + ".".join(op.split(".")[1:3]) if p in ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM") and "." in op else "")] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  ({len(counts)} kernels); demangle with c++filt")
for fn, c in counts.items():
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
    name = re.sub(r"\(.*", "", name)
    arch = [k for k in c if k.startswith("arch:")][0][5:]
    ops = ", ".join(f"{k} x{v}" for k, v in sorted(c.items()) if not k.startswith(("arch:", "_total")))
    print(f"{name} [{arch}, {c['_total']} instr]: {ops or '-'}")
