"""ctypes binding of libb200hnsw.so (the C-ABI declared in include/b200_hnsw.h).

The library is built in-tree by `build()` (nvcc, sm_100a only). Loading fails loudly when the
shared object is missing — there is no Python or CPU fallback for any entry point.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200hnsw.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200_hnsw.h")


class SearchParams(C.Structure):
    _fields_ = [("efSearch", C.c_int32), ("check_relative_distance", C.c_int32),
                ("warps_per_query", C.c_int32), ("hash_bits", C.c_int32),
                ("stats", C.c_void_p), ("sel_bitmap", C.c_void_p), ("sel_bitmap_bytes", C.c_int64),
                ("visited_policy", C.c_int32), ("reserved_", C.c_int32)]


class BuildParams(C.Structure):
    _fields_ = [("max_batch", C.c_int32), ("batch_divisor", C.c_int32),
                ("warps_per_query", C.c_int32), ("hash_bits", C.c_int32), ("visited_policy", C.c_int32)]


SHARDS_BLOB_BYTES = 160


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libb200hnsw.so (nvcc cross-compiles w/o a GPU)."""
    env = dict(os.environ)
    env["PATH"] = "/usr/bin:/bin:/usr/local/cuda/bin:" + env.get("PATH", "")
    for k in ("CXX", "CC"):
        env.pop(k, None)
    cmd = ["make", "-C", _HERE, "-j8"] + (["-B"] if force else [])
    r = subprocess.run(cmd, env=env, capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libb200hnsw.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None

_P = C.c_void_p
_SIGS = {
    "bh_index_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int]),
    "bh_index_free": (C.c_int, [_P]),
    "bh_index_reset": (C.c_int, [_P]),
    "bh_index_set_vector_storage": (C.c_int, [_P, C.c_int]),
    "bh_index_get_vector_storage": (C.c_int, [_P]),
    "bh_index_train": (C.c_int, [_P, C.c_int64, _P]),
    "bh_index_add": (C.c_int, [_P, C.c_int64, _P]),
    "bh_index_add_ex": (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    "bh_index_search": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, C.POINTER(SearchParams)]),
    "bh_index_search_device": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, C.POINTER(SearchParams)]),
    "bh_index_reconstruct": (C.c_int, [_P, C.c_int64, _P]),
    "bh_index_reconstruct_n": (C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    "bh_index_ntotal": (C.c_int64, [_P]),
    "bh_index_d": (C.c_int, [_P]),
    "bh_index_M": (C.c_int, [_P]),
    "bh_index_metric": (C.c_int, [_P]),
    "bh_index_entry_point": (C.c_int, [_P]),
    "bh_index_max_level": (C.c_int, [_P]),
    "bh_index_get_ef_search": (C.c_int, [_P]),
    "bh_index_set_ef_search": (C.c_int, [_P, C.c_int]),
    "bh_index_get_ef_construction": (C.c_int, [_P]),
    "bh_index_set_ef_construction": (C.c_int, [_P, C.c_int]),
    "bh_index_set_check_relative_distance": (C.c_int, [_P, C.c_int]),
    "bh_index_set_build_params": (C.c_int, [_P, C.POINTER(BuildParams)]),
    "bh_index_neighbors_size": (C.c_int64, [_P]),
    "bh_index_export_graph": (C.c_int, [_P, _P, _P, _P]),
    "bh_index_import_graph": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_int64, C.c_int, C.c_int]),
    "bh_index_stream": (_P, [_P]),
    "bh_index_synchronize": (C.c_int, [_P]),
    "bh_index_last_build_ms": (C.c_float, [_P]),
    "bh_index_last_search_ms": (C.c_float, [_P]),
    "bh_index_last_build_counters": (C.c_int, [_P, _P]),
    "bh_launch_count": (C.c_int64, []),
    "bh_merge_topk_device": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int, _P, _P, _P, _P, _P, _P]),
    "bh_selector_range_to_bitmap": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, _P]),
    "bh_selector_batch_to_bitmap": (C.c_int, [C.c_int64, C.c_int64, _P, _P]),
    "bh_selector_not": (C.c_int, [C.c_int64, _P]),
    "bh_shards_create": (C.c_int, [C.POINTER(_P), _P, C.c_int, C.c_int, C.c_int64, C.c_int64]),
    "bh_shards_free": (C.c_int, [_P]),
    "bh_shards_export": (C.c_int, [_P, _P]),
    "bh_shards_connect": (C.c_int, [_P, _P]),
    "bh_shards_set_ntotals": (C.c_int, [_P, _P]),
    "bh_shards_search_device": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, C.POINTER(SearchParams)]),
    "bh_shards_post": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.POINTER(SearchParams), C.c_int]),
    "bh_shards_gather": (C.c_int, [_P, C.c_int64, C.c_int64, C.POINTER(_P)]),
    "bh_shards_collect": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, C.c_int]),
    "bh_shards_local_lists": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    "bh_shards_status": (C.c_int, [_P]),
    "bh_shards_set_pipelined": (C.c_int, [_P, C.c_int]),
    "bh_shards_join": (C.c_int, [_P, _P]),
    "bh_last_error": (C.c_char_p, []),
    "bh_version": (C.c_char_p, []),
}
EXPORTED = tuple(_SIGS)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and os.environ.get("BH_NO_AUTOBUILD") != "1":
            try:  # the prebuilt library normally travels with the tree; compile it if it did not
                build()
            except Exception:
                pass
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no fallback path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError = the library does not export the header's symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().bh_last_error().decode()


def check(rc: int):
    if rc != 0:
        raise RuntimeError(last_error())
