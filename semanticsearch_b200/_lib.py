"""ctypes binding of the C ABI declared in include/semsearch_b200.h.

The library is loaded lazily (after a ``spawn``, inside the worker process — SURVEY.md §8b
threading model).  There is no fallback: if the shared object is missing or a call fails, a
``RuntimeError`` carrying ``ss_last_error()`` is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint32, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsemsearch_b200.so")

SS_F32, SS_BF16, SS_F16 = 0, 1, 2

_lock = threading.Lock()
_lib = None

_SIGNATURES = {
    "ss_version": (c_int, []),
    "ss_last_error": (c_char_p, []),
    "ss_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t)]),
    "ss_profile_begin": (c_int, [c_int]),
    "ss_profile_end": (c_int, [POINTER(c_float), c_int, POINTER(c_int), POINTER(c_int)]),
    "ss_cosine_topk_stream_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int, c_int]),
    "ss_cosine_topk_stream": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_int, c_uint32,
                                      c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ss_cosine_topk_gemm_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "ss_cosine_topk_gemm": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_uint32, c_void_p, c_size_t,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "ss_cosine_topk_gemm_resident": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_uint32, c_void_p, c_size_t,
                                             c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ss_cosine_topk_tcstream_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "ss_cosine_topk_tcstream": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_uint32, c_void_p, c_size_t,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "ss_cosine_topk_small_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "ss_cosine_topk_small": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_int, c_uint32, c_void_p, c_size_t,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "ss_cosine_scores": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "ss_rank_order_workspace_bytes": (c_size_t, [c_int, c_int64]),
    "ss_rank_order": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "ss_segmented_rank_rrf": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_double, c_double, c_double,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ss_topk_merge": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p]),
    "ss_peer_buffer_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ss_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "ss_peer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "ss_peer_close": (c_int, [c_void_p]),
    "ss_peer_free": (c_int, [c_void_p]),
    "ss_topk_peer_exchange_merge": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_uint32, c_void_p, c_void_p,
                                            c_void_p, c_void_p]),
    "ss_topk_peer_exchange_merge_auto": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_void_p]),
    "ss_peer_status": (c_int, [c_void_p, c_void_p]),
    "ss_segmented_plan_host": (c_int, [c_void_p, c_int, c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int)]),
    "ss_segmented_simmatrix": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "ss_segmented_plan128_host": (c_int, [c_void_p, c_int, c_void_p, c_int64, POINTER(c_int64)]),
    "ss_segmented_simmatrix_tc": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "ss_group_threshold_pass": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "ss_group_block_sums": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "ss_group_coassociation": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ss_similarity_distribution": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "ss_diameter_split": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ss_c99_rank_matrix": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ss_c99_divisive_cuts": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_double, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ss_segmented_adjacent_cosine": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "ss_segmented_percentile": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "ss_row_inv_norms": (c_int, [c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Return the loaded library handle (thread-safe, cached)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m semanticsearch_b200.build` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


class DeviceError(RuntimeError):
    """A C-ABI entry point failed (bad arguments, CUDA error).  Host layers that isolate per-item failures, as the reference
    does around each query, let this one propagate: there is no CPU path to fall back to."""


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().ss_last_error()
        raise DeviceError(f"{what} failed (status {status}): {msg.decode() if msg else 'unknown error'}")
