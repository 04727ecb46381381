"""ctypes wrapper over oracle/liboracle.so — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The CPU restatement of faiss::IndexHNSWFlat semantics used as
the parity checker and as the timed CPU baseline. PARITY UNPINNED: see the header of
hnsw_oracle.cpp. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; hnsw_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(arch: str | None = None, out: str = "liboracle.so", force: bool = False) -> str:
    """Compile the oracle. arch=None -> portable x86-64-v3; 'native' for baseline timing."""
    path = os.path.join(_HERE, out)
    src = os.path.join(_HERE, "hnsw_oracle.cpp")
    if not force and os.path.exists(path) and os.path.getmtime(path) >= os.path.getmtime(src):
        return path
    cmd = ["g++", "-O3", f"-march={arch or 'x86-64-v3'}", "-fopenmp", "-fPIC", "-std=c++17",
           "-shared", "-o", path, src]
    env = dict(os.environ)
    env["PATH"] = "/usr/bin:/bin:" + env.get("PATH", "")
    subprocess.run(cmd, check=True, env=env)
    return path


_libs: dict[str, C.CDLL] = {}


def _load(path: str | None = None) -> C.CDLL:
    path = path or build()
    if path in _libs:
        return _libs[path]
    L = C.CDLL(path)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int]
    L.orc_free.argtypes = [C.c_void_p]
    for name in ("orc_set_ef_construction", "orc_set_ef_search", "orc_set_check_relative_distance"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int]
    L.orc_set_team.argtypes = [C.c_void_p, C.c_int]
    L.orc_set_team.restype = C.c_int
    L.orc_set_half_storage.argtypes = [C.c_void_p, C.c_int]
    L.orc_set_half_storage.restype = C.c_int
    L.orc_round_to_half.argtypes = [C.c_float]
    L.orc_round_to_half.restype = C.c_float
    L.orc_ntotal.argtypes = [C.c_void_p]
    L.orc_ntotal.restype = C.c_int64
    for name in ("orc_entry_point", "orc_max_level", "orc_n_levels_table"):
        getattr(L, name).argtypes = [C.c_void_p]
        getattr(L, name).restype = C.c_int
    L.orc_max_threads.restype = C.c_int
    L.orc_neighbors_size.argtypes = [C.c_void_p]
    L.orc_neighbors_size.restype = C.c_int64
    L.orc_get_assign_probas.argtypes = [C.c_void_p, _f64p]
    L.orc_get_cum_nneighbor.argtypes = [C.c_void_p, _i32p]
    L.orc_add.argtypes = [C.c_void_p, C.c_int64, _f32p, C.c_int, C.c_void_p]
    L.orc_add.restype = C.c_int
    L.orc_add_preset.argtypes = [C.c_void_p, C.c_int64, _f32p, _i32p, C.c_int]
    L.orc_add_preset.restype = C.c_int
    L.orc_peek_levels.argtypes = [C.c_void_p, C.c_int64, _i32p]
    L.orc_search.argtypes = [C.c_void_p, C.c_int64, _f32p, C.c_int64, _f32p, _i64p, C.c_int,
                             C.c_int, C.c_void_p]
    L.orc_search.restype = C.c_int
    L.orc_search_sel.argtypes = [C.c_void_p, C.c_int64, _f32p, C.c_int64, _f32p, _i64p, C.c_int,
                                 C.c_int, C.c_void_p, C.c_void_p]
    L.orc_search_sel.restype = C.c_int
    L.orc_export_graph.argtypes = [C.c_void_p, _i32p, _u64p, _i32p]
    L.orc_import.argtypes = [C.c_void_p, C.c_int64, _f32p, _i32p, _i32p, C.c_int64, C.c_int, C.c_int]
    L.orc_import.restype = C.c_int
    L.orc_distance.argtypes = [C.c_void_p, _f32p, _f32p]
    L.orc_distance.restype = C.c_float
    L.orc_mt19937_first.argtypes = [C.c_uint32]
    L.orc_mt19937_first.restype = C.c_uint32
    L.orc_shrink.argtypes = [C.c_void_p, C.c_int, _i32p, _f32p, C.c_int, _i32p]
    L.orc_shrink.restype = C.c_int
    _libs[path] = L
    return L


class OracleHNSWFlat:
    """Mirrors faiss.IndexHNSWFlat(d, M, metric): add / search, efSearch, efConstruction."""

    def __init__(self, d: int, M: int = 32, metric: int = METRIC_L2, lib_path: str | None = None):
        self._L = _load(lib_path)
        self._h = self._L.orc_create(d, M, metric)
        if not self._h:
            raise ValueError("bad oracle parameters")
        self.d, self.M, self.metric_type = d, M, metric
        self._efS, self._efC = 16, 40
        self.threads = 1

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_free(self._h)
            self._h = None

    # --- parameters
    @property
    def efSearch(self):
        return self._efS

    @efSearch.setter
    def efSearch(self, v):
        self._efS = int(v)
        self._L.orc_set_ef_search(self._h, int(v))

    @property
    def efConstruction(self):
        return self._efC

    @efConstruction.setter
    def efConstruction(self, v):
        self._efC = int(v)
        self._L.orc_set_ef_construction(self._h, int(v))

    def set_team(self, T: int):
        if self._L.orc_set_team(self._h, int(T)) != 0:
            raise ValueError("bad team")

    def set_half_storage(self, on=True):
        """Emulate the engine's opt-in 16-bit vector storage: True / 1 / "fp16" = vectors rounded to binary16 on
        add, 2 / "bf16" = to bfloat16."""
        kind = {"fp16": 1, "bf16": 2}[on] if isinstance(on, str) else int(on)
        if self._L.orc_set_half_storage(self._h, kind) != 0:
            raise ValueError("set_half_storage: index not empty or d % 8 != 0")

    def set_check_relative_distance(self, v: bool):
        self._L.orc_set_check_relative_distance(self._h, int(bool(v)))

    @property
    def ntotal(self):
        return int(self._L.orc_ntotal(self._h))

    @property
    def entry_point(self):
        return int(self._L.orc_entry_point(self._h))

    @property
    def max_level(self):
        return int(self._L.orc_max_level(self._h))

    def max_threads(self):
        return int(self._L.orc_max_threads())

    # --- Index API
    def add(self, x: np.ndarray, return_order: bool = False):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        order = np.empty(x.shape[0], np.int32) if return_order else None
        rc = self._L.orc_add(self._h, x.shape[0], x, int(self.threads),
                             order.ctypes.data if order is not None else None)
        if rc:
            raise RuntimeError(f"orc_add rc={rc}")
        return order

    def add_with_levels(self, x: np.ndarray, levels):
        """add() with preset levels (level + 1 per row), faiss's `hnsw.levels` preset."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        lv = np.ascontiguousarray(levels, np.int32)
        assert x.ndim == 2 and x.shape[1] == self.d and lv.shape == (x.shape[0],)
        rc = self._L.orc_add_preset(self._h, x.shape[0], x, lv, int(self.threads))
        if rc:
            raise RuntimeError(f"orc_add_preset rc={rc}")

    def peek_levels(self, n: int) -> np.ndarray:
        out = np.empty(n, np.int32)
        self._L.orc_peek_levels(self._h, n, out)
        return out

    def search(self, xq: np.ndarray, k: int, efSearch: int | None = None, stats: bool = False,
               sel_bitmap: np.ndarray | None = None):
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        nq = xq.shape[0]
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        st = np.zeros((nq, 4), np.int32) if stats else None
        sel = None if sel_bitmap is None else np.ascontiguousarray(sel_bitmap, np.uint8)
        rc = self._L.orc_search_sel(self._h, nq, xq, k, D, I, int(efSearch or 0), int(self.threads),
                                    st.ctypes.data if st is not None else None,
                                    sel.ctypes.data if sel is not None else None)
        if rc:
            raise RuntimeError(f"orc_search rc={rc}")
        return (D, I, st) if stats else (D, I)

    # --- graph access (faiss layout)
    def export_graph(self):
        n = self.ntotal
        levels = np.empty(n, np.int32)
        offsets = np.empty(n + 1, np.uint64)
        neighbors = np.empty(int(self._L.orc_neighbors_size(self._h)), np.int32)
        self._L.orc_export_graph(self._h, levels, offsets, neighbors)
        return dict(levels=levels, offsets=offsets, neighbors=neighbors,
                    entry_point=self.entry_point, max_level=self.max_level)

    def import_graph(self, x, levels, neighbors, entry_point, max_level):
        x = np.ascontiguousarray(x, np.float32)
        levels = np.ascontiguousarray(levels, np.int32)
        neighbors = np.ascontiguousarray(neighbors, np.int32)
        rc = self._L.orc_import(self._h, x.shape[0], x, levels, neighbors, neighbors.shape[0],
                                int(entry_point), int(max_level))
        if rc:
            raise RuntimeError(f"orc_import rc={rc}")

    def tables(self):
        n = self._L.orc_n_levels_table(self._h)
        p = np.empty(n, np.float64)
        c = np.empty(n + 1, np.int32)
        self._L.orc_get_assign_probas(self._h, p)
        self._L.orc_get_cum_nneighbor(self._h, c)
        return p, c

    def distance(self, a, b) -> float:
        return float(self._L.orc_distance(self._h, np.ascontiguousarray(a, np.float32),
                                          np.ascontiguousarray(b, np.float32)))

    def shrink(self, ids, dq, max_size):
        ids = np.ascontiguousarray(ids, np.int32)
        dq = np.ascontiguousarray(dq, np.float32)
        out = np.empty(max(len(ids), 1), np.int32)
        n = self._L.orc_shrink(self._h, len(ids), ids, dq, int(max_size), out)
        return out[:n].copy()


def mt19937_first(seed: int) -> int:
    return int(_load().orc_mt19937_first(seed))


def brute_force_knn(xb: np.ndarray, xq: np.ndarray, k: int, metric: int = METRIC_L2):
    """Exact top-k in float64 (small cases). Returns (D, I) sorted best-first."""
    xb64, xq64 = xb.astype(np.float64), xq.astype(np.float64)
    if metric == METRIC_L2:
        dm = (xq64 ** 2).sum(1)[:, None] - 2 * xq64 @ xb64.T + (xb64 ** 2).sum(1)[None, :]
        I = np.argsort(dm, axis=1, kind="stable")[:, :k]
    else:
        dm = xq64 @ xb64.T
        I = np.argsort(-dm, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(dm, I, 1), I


def recall_at_k(I: np.ndarray, gt: np.ndarray) -> float:
    k = gt.shape[1]
    hits = sum(len(set(I[i, :k].tolist()) & set(gt[i].tolist())) for i in range(gt.shape[0]))
    return hits / float(gt.size)
