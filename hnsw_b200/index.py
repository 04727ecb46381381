"""IndexHNSWFlat — Python mirror of faiss.IndexHNSWFlat over the B200 engine's C-ABI.

Same constructor, fields and method meanings as faiss:
    index = IndexHNSWFlat(d, M, metric)      faiss.IndexHNSWFlat(d, M, metric)
    index.hnsw.efConstruction / efSearch      index.hnsw.efConstruction / efSearch
    index.train(x); index.add(x)              no-op train; add copies vectors and links them
    D, I = index.search(xq, k)                float32 [nq,k] ascending L2² (descending IP), int64
Errors surface as RuntimeError carrying the library's message (faiss raises on bad input too).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BuildParams, SearchParams
from .selectors import SearchParametersHNSW

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


class _HNSWFields:
    """index.hnsw.* — the faiss `HNSW` struct's tunables."""

    def __init__(self, owner: "IndexHNSWFlat"):
        self._o = owner

    @property
    def efSearch(self):
        return _lib.lib().bh_index_get_ef_search(self._o._h)

    @efSearch.setter
    def efSearch(self, v):
        _lib.check(_lib.lib().bh_index_set_ef_search(self._o._h, int(v)))

    @property
    def efConstruction(self):
        return _lib.lib().bh_index_get_ef_construction(self._o._h)

    @efConstruction.setter
    def efConstruction(self, v):
        _lib.check(_lib.lib().bh_index_set_ef_construction(self._o._h, int(v)))

    @property
    def entry_point(self):
        return _lib.lib().bh_index_entry_point(self._o._h)

    @property
    def max_level(self):
        return _lib.lib().bh_index_max_level(self._o._h)

    @property
    def check_relative_distance(self):
        return self._o._crd

    @check_relative_distance.setter
    def check_relative_distance(self, v):
        self._o._crd = bool(v)
        _lib.check(_lib.lib().bh_index_set_check_relative_distance(self._o._h, int(bool(v))))


STORAGE_F32 = 0
STORAGE_F16 = 1
STORAGE_BF16 = 2


class IndexHNSWFlat:
    def __init__(self, d: int, M: int = 32, metric: int = METRIC_L2, device: int = 0,
                 storage: str | int = "fp32"):
        """`storage="fp16"` / `"bf16"` (opt-in, not faiss-bit-comparable): rows are kept in 16 bits in HBM and
        accumulated in fp32 — half the gather bytes; the API stays fp32."""
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.bh_index_create(C.byref(h), int(d), int(M), int(metric), int(device)))
        self._h = h
        kind = {"fp32": STORAGE_F32, "f32": STORAGE_F32, "fp16": STORAGE_F16, "f16": STORAGE_F16,
                "bf16": STORAGE_BF16, "bfloat16": STORAGE_BF16}.get(storage, storage)
        if kind != STORAGE_F32:
            _lib.check(L.bh_index_set_vector_storage(h, int(kind)))
        self.storage = {STORAGE_F16: "fp16", STORAGE_BF16: "bf16"}.get(kind, "fp32")
        self.d = int(d)
        self.M = int(M)
        self.metric_type = int(metric)
        self.device = int(device)
        self.is_trained = True
        self.verbose = False
        self._crd = True
        self.hnsw = _HNSWFields(self)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().bh_index_free(h)
            except Exception:
                pass
            self._h = None

    # ---- faiss.Index surface
    @property
    def ntotal(self) -> int:
        return int(_lib.lib().bh_index_ntotal(self._h))

    def train(self, x=None):
        _lib.check(_lib.lib().bh_index_train(self._h, 0 if x is None else len(x), None))

    def _as_f32(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"expected a [n, {self.d}] float32 array, got {x.shape}")
        return x

    def add(self, x, levels=None, order=None):
        """faiss Index.add. `levels` (level+1 per row) and `order` preset the faiss-drawn values."""
        x = self._as_f32(x)
        if levels is None and order is None:
            _lib.check(_lib.lib().bh_index_add(self._h, x.shape[0], x.ctypes.data))
            return
        lv = None if levels is None else np.ascontiguousarray(levels, np.int32)
        od = None if order is None else np.ascontiguousarray(order, np.int32)
        _lib.check(_lib.lib().bh_index_add_ex(self._h, x.shape[0], x.ctypes.data,
                                              None if lv is None else lv.ctypes.data,
                                              None if od is None else od.ctypes.data))

    def search(self, x, k: int, params: SearchParams | None = None, efSearch: int | None = None,
               stats: bool = False, out=None, warps_per_query: int = 0, hash_bits: int = 0,
               sel_bitmap=None, visited_policy: int = 0):
        """faiss Index.search → (D, I). `out=(D, I)` reuses caller buffers (e.g. pinned).
        `sel_bitmap`: uint8 array in faiss IDSelectorBitmap layout (np.packbits(mask, bitorder="little"));
        only ids whose bit is set can be returned (SearchParametersHNSW.sel semantics)."""
        x = self._as_f32(x)
        nq = x.shape[0]
        if out is None:
            D = np.empty((nq, k), np.float32)
            I = np.empty((nq, k), np.int64)
        else:
            D, I = out
            for a, dt in ((D, np.float32), (I, np.int64)):  # raw pointers go to C: refuse anything else
                if not (isinstance(a, np.ndarray) and a.dtype == dt and a.shape == (nq, k) and a.flags.c_contiguous):
                    raise ValueError(f"out= needs C-contiguous ({nq}, {k}) float32 / int64 arrays")
        st = np.zeros((nq, 4), np.int32) if stats else None
        if isinstance(params, SearchParametersHNSW):   # faiss-style per-call parameters
            crd = params.check_relative_distance
            if params.sel is not None:
                sel_bitmap = params.sel.to_bitmap(self.ntotal)
            params = SearchParams(int(params.efSearch or efSearch or 0), 0 if crd is None else (1 if crd else 2),
                                  int(warps_per_query), int(hash_bits), None, None, 0, int(visited_policy), 0)
        if params is None:
            params = SearchParams(int(efSearch or 0), 0, int(warps_per_query), int(hash_bits), None, None, 0,
                                  int(visited_policy), 0)
        if st is not None:
            params.stats = st.ctypes.data
        if sel_bitmap is not None:
            sel_bitmap = np.ascontiguousarray(sel_bitmap, np.uint8)
            params.sel_bitmap = sel_bitmap.ctypes.data
            params.sel_bitmap_bytes = sel_bitmap.size
        _lib.check(_lib.lib().bh_index_search(self._h, nq, x.ctypes.data, int(k), D.ctypes.data,
                                              I.ctypes.data, C.byref(params)))
        return (D, I, st) if stats else (D, I)

    def search_device(self, xq_ptr: int, nq: int, k: int, D_ptr: int, I_ptr: int,
                      efSearch: int = 0, stats_ptr: int = 0, warps_per_query: int = 0,
                      hash_bits: int = 0, visited_policy: int = 0):
        """Enqueue a search on device buffers (raw pointers); no host sync."""
        p = SearchParams(int(efSearch), 0, int(warps_per_query), int(hash_bits), stats_ptr or None, None, 0,
                         int(visited_policy), 0)
        _lib.check(_lib.lib().bh_index_search_device(self._h, int(nq), xq_ptr, int(k), D_ptr, I_ptr,
                                                     C.byref(p)))

    def reset(self):
        _lib.check(_lib.lib().bh_index_reset(self._h))

    def reconstruct(self, key: int):
        out = np.empty(self.d, np.float32)
        _lib.check(_lib.lib().bh_index_reconstruct(self._h, int(key), out.ctypes.data))
        return out

    def reconstruct_n(self, i0: int = 0, ni: int | None = None):
        """faiss Index.reconstruct_n: rows [i0, i0+ni) of the stored vectors."""
        ni = self.ntotal - i0 if ni is None else ni
        out = np.empty((ni, self.d), np.float32)
        _lib.check(_lib.lib().bh_index_reconstruct_n(self._h, int(i0), int(ni), out.ctypes.data))
        return out

    # ---- engine extras
    def set_build_params(self, max_batch=0, batch_divisor=0, warps_per_query=0, hash_bits=0, visited_policy=0):
        p = BuildParams(int(max_batch), int(batch_divisor), int(warps_per_query), int(hash_bits),
                        int(visited_policy))
        _lib.check(_lib.lib().bh_index_set_build_params(self._h, C.byref(p)))

    @property
    def stream_ptr(self) -> int:
        """cudaStream_t (as an integer) the index enqueues its kernels on."""
        return int(_lib.lib().bh_index_stream(self._h) or 0)

    def synchronize(self):
        _lib.check(_lib.lib().bh_index_synchronize(self._h))

    @property
    def last_build_ms(self):
        return float(_lib.lib().bh_index_last_build_ms(self._h))

    @property
    def last_build_counters(self):
        """Work counters of the last add(): ndis0, nhops0, ndis_up, nhops_up, sel_rows, bl_rows."""
        out = np.zeros(6, np.uint64)
        _lib.check(_lib.lib().bh_index_last_build_counters(self._h, out.ctypes.data))
        return dict(zip(("ndis0", "nhops0", "ndis_up", "nhops_up", "sel_rows", "bl_rows"), (int(v) for v in out)))

    @property
    def last_search_ms(self):
        return float(_lib.lib().bh_index_last_search_ms(self._h))

    def export_graph(self):
        """Graph in faiss's HNSW layout: levels, offsets, neighbors, entry_point, max_level."""
        L = _lib.lib()
        n = self.ntotal
        levels = np.empty(n, np.int32)
        offsets = np.empty(n + 1, np.uint64)
        neighbors = np.empty(int(L.bh_index_neighbors_size(self._h)), np.int32)
        _lib.check(L.bh_index_export_graph(self._h, levels.ctypes.data, offsets.ctypes.data,
                                           neighbors.ctypes.data))
        return dict(levels=levels, offsets=offsets, neighbors=neighbors,
                    entry_point=self.hnsw.entry_point, max_level=self.hnsw.max_level)

    def import_graph(self, x, levels, neighbors, entry_point, max_level):
        x = self._as_f32(x)
        levels = np.ascontiguousarray(levels, np.int32)
        neighbors = np.ascontiguousarray(neighbors, np.int32)
        _lib.check(_lib.lib().bh_index_import_graph(self._h, x.shape[0], x.ctypes.data,
                                                    levels.ctypes.data, neighbors.ctypes.data,
                                                    neighbors.shape[0], int(entry_point), int(max_level)))


def launch_count() -> int:
    return int(_lib.lib().bh_launch_count())


def merge_topk_device(D_all_ptr: int, I_all_ptr: int, nshard: int, nq: int, k: int, metric: int,
                      id_offsets, D_out_ptr: int, I_out_ptr: int, stream_ptr: int = 0):
    """Warp top-k merge of per-shard sorted lists (device pointers); see bh_merge_topk_device."""
    import numpy as np
    off = np.ascontiguousarray(id_offsets, np.int64)
    _lib.check(_lib.lib().bh_merge_topk_device(int(nshard), int(nq), int(k), int(metric), D_all_ptr,
                                               I_all_ptr, off.ctypes.data, D_out_ptr, I_out_ptr,
                                               stream_ptr or None))
