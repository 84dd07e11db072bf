"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads without a GPU and
exports every symbol that include/semsearch_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "semsearch_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from semanticsearch_b200 import build
    return build.build_library()


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "ss_cosine_topk_stream" in syms and "ss_topk_merge" in syms
    assert len(syms) >= 7


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert missing == []


def test_binding_table_matches_header(lib_path):
    from semanticsearch_b200 import _lib
    assert sorted(_lib.exported_symbols()) == declared_symbols()
    lib = _lib.load()
    assert lib.ss_version() >= 100


def test_no_cpu_fallback():
    import torch
    from semanticsearch_b200 import similarity
    with pytest.raises(RuntimeError, match="CUDA"):
        similarity.cosine_topk(torch.zeros(4, 8), torch.zeros(1, 8), 2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "semanticsearch_b200")
    offenders = []
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert offenders == []
