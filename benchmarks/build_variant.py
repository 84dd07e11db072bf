#!/usr/bin/env python
"""Build the current csrc/ tree into benchmarks/_variants/lib_<name>.so (git-ignored; travels to the GPU box) for A/B timing
of kernel variants inside one process (`benchmarks/ab_k3.py`)."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semanticsearch_b200 import build as B  # noqa: E402

name = sys.argv[1]
extra = sys.argv[2:]
out_dir = os.path.join(ROOT, "benchmarks", "_variants")
obj_dir = os.path.join(out_dir, "obj_" + name)
os.makedirs(obj_dir, exist_ok=True)


def one(src):
    obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
    subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *extra, "-c", src, "-o", obj], check=True)
    return obj


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(one, B.sources()))
lib = os.path.join(out_dir, f"lib_{name}.so")
subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-cudart", "static"], check=True)
print(lib)
