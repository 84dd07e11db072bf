"""CPU-only checks of host-side planning and dispatch logic (no kernel launches)."""
import ctypes

import numpy as np
import pytest
import torch


def _expected_units(sizes):
    """Host restatement of the 128-tile plan: consecutive documents of <= 64 rows share packed windows of
    <= 128 rows (kind 1: first doc, doc count, rows); everything else is upper-triangular 128 x 128 tiles."""
    want, d = [], 0
    while d < len(sizes):
        if sizes[d] <= 64:
            d1, rows = d, 0
            while d1 < len(sizes) and sizes[d1] <= 64 and rows + sizes[d1] <= 128:
                rows += sizes[d1]
                d1 += 1
            if rows > 0:
                want.append((d, 0, 0, 0) if d1 - d == 1 else (d, d1 - d, rows, 1))
            d = d1
            continue
        T = (sizes[d] + 127) // 128
        want += [(d, i, j, 0) for i in range(T) for j in range(i, T)]
        d += 1
    return want


def test_plan128_unit_table_tiles_and_packed_windows():
    from semanticsearch_b200 import _lib
    lib = _lib.load()
    for sizes in ([0, 1, 128, 129, 300, 512, 513], [10, 20, 64, 64, 1, 0, 0, 5, 65, 3, 120, 7, 7, 7], [0, 0], [64, 65, 64]):
        offsets = np.zeros(len(sizes) + 1, dtype=np.int32)
        offsets[1:] = np.cumsum(sizes)
        total = ctypes.c_int64()
        assert lib.ss_segmented_plan128_host(offsets.ctypes.data, len(sizes), None, 0, ctypes.byref(total)) == 0
        want = _expected_units(sizes)
        assert total.value == len(want)
        units = np.zeros((max(len(want), 1), 4), dtype=np.int32)
        assert lib.ss_segmented_plan128_host(offsets.ctypes.data, len(sizes), units.ctypes.data, len(want), ctypes.byref(total)) == 0
        assert [tuple(u) for u in units[: len(want)].tolist()] == want
    # a table that is too small is refused, decreasing offsets are refused
    sizes = [0, 1, 128, 129, 300]
    offsets = np.zeros(len(sizes) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(sizes)
    units = np.zeros((64, 4), dtype=np.int32)
    total = ctypes.c_int64()
    assert lib.ss_segmented_plan128_host(offsets.ctypes.data, len(sizes), units.ctypes.data, 3, ctypes.byref(total)) != 0
    bad = np.array([0, 5, 3], dtype=np.int32)
    assert lib.ss_segmented_plan128_host(bad.ctypes.data, 2, None, 0, ctypes.byref(total)) != 0


def test_choose_algo_dispatch_rules():
    from semanticsearch_b200 import similarity as sim
    C16 = torch.zeros((1000, 768), dtype=torch.bfloat16)
    C32 = torch.zeros((1000, 768), dtype=torch.float32)
    q = lambda b, dt=torch.bfloat16: torch.zeros((b, 768), dtype=dt)
    assert sim.choose_algo(C16, q(1), 10) == "stream"
    assert sim.choose_algo(C16, q(2), 10) == "tcstream"
    assert sim.choose_algo(C16, q(16), 100) == "tcstream"          # BASELINE config 5
    assert sim.choose_algo(C16, q(sim.GEMM_MIN_BATCH - 1), 10) == "tcstream"
    assert sim.choose_algo(C16, q(4096), 10) == "gemm"             # BASELINE config 4b
    assert sim.choose_algo(C16, q(4096), 100) == "tcstream"        # k > 16: pooled top-k, several corpus passes
    assert sim.choose_algo(C16, q(4096), 2000) == "small"          # beyond K7's k limit, corpus small: score matrix + selection
    huge16 = torch.zeros((1 << 18, 8), dtype=torch.bfloat16)
    assert sim.choose_algo(huge16, torch.zeros((4096, 8), dtype=torch.bfloat16), 2000) == "stream"
    assert sim.choose_algo(C32, q(100, torch.float32), 10) == "small"    # BASELINE config 1 scale: score matrix + selection
    assert sim.choose_algo(C32, q(8, torch.float32), 10) == "stream"     # one K1 query group
    big32 = torch.zeros((1 << 18, 8), dtype=torch.float32)
    assert sim.choose_algo(big32, torch.zeros((100, 8)), 10) == "stream"  # fp32 keeps the 1e-5 bound on CUDA cores
    assert sim.choose_algo(C16, q(8, torch.float16), 10) == "stream"     # mixed dtypes
    odd = torch.zeros((1000, 100), dtype=torch.bfloat16)
    assert sim.choose_algo(odd, torch.zeros((8, 100), dtype=torch.bfloat16), 10) == "stream"  # rows not 16-byte multiples


def test_operators_refuse_cpu_tensors():
    from semanticsearch_b200 import ragged, similarity
    with pytest.raises(RuntimeError, match="CUDA"):
        similarity.segmented_rank_rrf(torch.zeros((4, 8)), torch.zeros(2, dtype=torch.int32), torch.zeros((1, 8)))
    with pytest.raises(RuntimeError, match="CUDA"):
        ragged.adjacent_cosine(torch.zeros((4, 8)))
