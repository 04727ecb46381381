"""ShardedIndexHNSWFlat — the database sharded over the ranks of a torch.distributed group.

Semantics of faiss IndexShards(successive_ids=True) + merge_knn_results:
rank r owns the contiguous id range [offset_r, offset_r + ntotal_r) and an independent HNSW graph
over it; a search broadcasts the queries, every rank searches its shard, the per-shard sorted
top-k lists (distance, LOCAL id) are exchanged, and every rank merges them into the global top-k
with ids shifted by the owning shard's offset.

On GPUs this class is a thin binding of the C-ABI's bh_shards_* entry points (include/b200_hnsw.h);
torch.distributed is used only to broadcast the queries and to carry the bootstrap blobs:
  exchange "peer-store"  (default): the traversal kernel's epilogue stores every query's k results as
      8-byte (distance bits, local id) keys straight into every rank's gather buffer over NVLink
      (CUDA-IPC mapped peer memory), a one-warp kernel raises a flag in every peer, the merge kernel
      waits for the flags and merges — no collective library call on the data path;
  exchange "nccl-allgather-packed" (when CUDA IPC is unavailable): the same 8-byte payload, ONE in-place
      all_gather_into_tensor of nq*k*8 bytes per rank, then the merge kernel.
Nothing on either path runs on the CPU. The local index and the merge function can be injected, which
is how the host-side bookkeeping (offsets, gather layout, broadcast) is tested on CPU with the gloo
backend and the CPU oracle (tests/).
"""
from __future__ import annotations

import ctypes as C

import numpy as np


class ShardedIndexHNSWFlat:
    def __init__(self, d: int, M: int = 32, metric: int = 1, group=None, device=None,
                 local_index=None, merge_fn=None, exchange: str = "auto"):
        import torch.distributed as dist

        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.d, self.M, self.metric_type = d, M, metric
        self.device = device
        self._merge_fn = merge_fn
        if local_index is None:
            from .index import IndexHNSWFlat
            import torch
            dev_index = device.index if isinstance(device, torch.device) else int(device or 0)
            local_index = IndexHNSWFlat(d, M, metric, device=dev_index)
        self.local = local_index
        self.offsets = np.zeros(self.world, np.int64)
        self.ntotals = np.zeros(self.world, np.int64)
        self._want_exchange = exchange       # "auto" | "peer-store" | "nccl-allgather-packed"
        self.exchange_kind = None
        self._s = None                        # bh_shards handle
        self._cap = (0, 0)
        self.last_local = None

    def __del__(self):
        self._free()

    def _free(self):
        s = getattr(self, "_s", None)
        if s:
            try:  # (at interpreter shutdown even the import may fail)
                from . import _lib
                _lib.lib().bh_shards_free(s)
            except Exception:
                pass
            self._s = None

    # ---- bookkeeping shared by both paths
    def _refresh_offsets(self):
        import torch
        t = torch.tensor([int(self.local.ntotal)], dtype=torch.int64,
                         device=self.device if self._on_gpu() else "cpu")
        out = [torch.zeros_like(t) for _ in range(self.world)]
        self._dist.all_gather(out, t, group=self.group)
        self.ntotals = np.array([int(x.item()) for x in out], np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(self.ntotals)[:-1]]).astype(np.int64)
        if self._s:
            from . import _lib
            _lib.check(_lib.lib().bh_shards_set_ntotals(self._s, self.ntotals.ctypes.data))

    def _on_gpu(self):
        return self._merge_fn is None

    @property
    def ntotal(self) -> int:
        return int(self.ntotals.sum())

    def add(self, x_local):
        """Add this rank's slice of the database (independent sub-graph; no communication)."""
        self.local.add(x_local)
        self._refresh_offsets()

    # ---- C-ABI handle: created collectively, sized by the first call (re-created if a later call is larger)
    def _ensure_handle(self, nq: int, k: int):
        import torch
        from . import _lib
        if self._s and nq <= self._cap[0] and k <= self._cap[1]:
            return
        self._free()
        L = _lib.lib()
        cap_q = max(4096, 1 << int(np.ceil(np.log2(max(nq, 1)))))
        cap_k = max(16, int(k))
        s = C.c_void_p()
        _lib.check(L.bh_shards_create(C.byref(s), self.local._h, self.rank, self.world, cap_q, cap_k))
        self._s, self._cap = s, (cap_q, cap_k)
        kind = "peer-store"
        if self.world > 1:
            blob = (C.c_ubyte * _lib.SHARDS_BLOB_BYTES)()
            _lib.check(L.bh_shards_export(s, blob))
            mine = torch.frombuffer(bytearray(bytes(blob)), dtype=torch.uint8).to(self.device)
            allb = torch.empty(self.world * _lib.SHARDS_BLOB_BYTES, dtype=torch.uint8, device=self.device)
            self._dist.all_gather_into_tensor(allb, mine, group=self.group)
            ok = 0
            if self._want_exchange != "nccl-allgather-packed":
                raw = allb.cpu().numpy().tobytes()
                ok = 1 if L.bh_shards_connect(s, raw) == 0 else 0
                self._connect_error = None if ok else _lib.last_error()
            t = torch.tensor([ok], dtype=torch.int32, device=self.device)
            self._dist.all_reduce(t, op=self._dist.ReduceOp.MIN, group=self.group)  # all ranks or none
            if int(t.item()) == 0:
                if self._want_exchange == "peer-store":
                    raise RuntimeError(f"peer-store exchange unavailable: {getattr(self, '_connect_error', None)}")
                kind = "nccl-allgather-packed"
        self.exchange_kind = kind
        _lib.check(L.bh_shards_set_ntotals(s, self.ntotals.ctypes.data))

    # ---- pipelined serving loop (GPU path): enqueue batch after batch, join when results are needed
    def set_pipelined(self, on: bool):
        """On: flag + merge kernels run on the engine's exchange stream, so back-to-back `enqueue` calls keep the
        index's stream full of traversal launches (they overlap their drain phases and do not wait for slower
        peers). Results of a call are complete only after `join()`."""
        from . import _lib
        if not self._s:
            raise RuntimeError("call search() once first (it creates and connects the exchange buffers)")
        _lib.check(_lib.lib().bh_shards_set_pipelined(self._s, int(bool(on))))

    def enqueue(self, q_dev, k: int, D_out, I_out, efSearch: int | None = None):
        """Collective. q_dev: float32 [nq, d] CUDA tensor already present (and complete) on every rank; D_out / I_out:
        float32 / int64 [nq, k] CUDA tensors. Enqueues on the index's stream; nothing is synchronised."""
        from . import _lib
        from ._lib import SearchParams
        p = SearchParams(int(efSearch or 0), 0, 0, 0, None, None, 0, 0, 0)
        _lib.check(_lib.lib().bh_shards_search_device(self._s, int(q_dev.shape[0]), q_dev.data_ptr(), int(k),
                                                      D_out.data_ptr(), I_out.data_ptr(), C.byref(p)))

    def join(self, stream=None):
        """Make `stream` (a torch stream; default: the current one) wait for the latest enqueued call's results."""
        import torch
        from . import _lib
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        _lib.check(_lib.lib().bh_shards_join(self._s, st.cuda_stream))

    def search(self, xq, k: int, efSearch: int | None = None, src: int = 0, keep_local: bool = False):
        """xq: [nq, d] float32 on rank `src` (other ranks pass an array of the same shape).
        Returns (D, I) with global ids on every rank (device tensors on the GPU path).
        keep_local: also decode this rank's own lists into self.last_local (for checks)."""
        import torch
        dist = self._dist
        nq = int(xq.shape[0])
        if self._on_gpu():
            from . import _lib
            from ._lib import SearchParams
            L = _lib.lib()
            q = torch.as_tensor(np.ascontiguousarray(xq, np.float32)).to(self.device) \
                if not isinstance(xq, torch.Tensor) else xq.to(self.device, torch.float32).contiguous()
            if self.world > 1:
                dist.broadcast(q, src=src, group=self.group)
            self._ensure_handle(nq, k)
            Dm = torch.empty(nq, k, device=self.device)
            Im = torch.empty(nq, k, dtype=torch.int64, device=self.device)
            p = SearchParams(int(efSearch or 0), 0, 0, 0, None, None, 0, 0, 0)
            cur = torch.cuda.current_stream(self.device)
            ist = torch.cuda.ExternalStream(self.local.stream_ptr, device=self.device)
            ist.wait_stream(cur)
            if self.exchange_kind == "peer-store":
                _lib.check(L.bh_shards_search_device(self._s, nq, q.data_ptr(), k, Dm.data_ptr(), Im.data_ptr(),
                                                     C.byref(p)))
            else:
                if nq > self._cap[0]:
                    raise ValueError("batch larger than the gather buffer")
                _lib.check(L.bh_shards_post(self._s, nq, q.data_ptr(), k, C.byref(p), 0))
                gp = C.c_void_p()
                _lib.check(L.bh_shards_gather(self._s, nq, k, C.byref(gp)))
                G = _device_view_i64(gp.value, self.world * nq * k, self.device)
                cur.wait_stream(ist)
                dist.all_gather_into_tensor(G, G[self.rank * nq * k:(self.rank + 1) * nq * k], group=self.group)
                ist.wait_stream(cur)
                _lib.check(L.bh_shards_collect(self._s, nq, k, Dm.data_ptr(), Im.data_ptr(), 0))
            if keep_local:
                Dl = torch.empty(nq, k, device=self.device)
                Il = torch.empty(nq, k, dtype=torch.int64, device=self.device)
                _lib.check(L.bh_shards_local_lists(self._s, nq, k, Dl.data_ptr(), Il.data_ptr()))
                self.last_local = (Dl, Il)
            cur.wait_stream(ist)
            return Dm, Im
        # injected (CPU / gloo) path — used by tests only
        q = torch.from_numpy(np.ascontiguousarray(xq, np.float32)).clone()
        dist.broadcast(q, src=src, group=self.group)
        Dl, Il = self.local.search(q.numpy(), k, efSearch)
        Dl, Il = torch.from_numpy(np.ascontiguousarray(Dl)), torch.from_numpy(np.ascontiguousarray(Il))
        Dg = [torch.empty_like(Dl) for _ in range(self.world)]
        Ig = [torch.empty_like(Il) for _ in range(self.world)]
        dist.all_gather(Dg, Dl, group=self.group)
        dist.all_gather(Ig, Il, group=self.group)
        return self._merge_fn(np.stack([t.numpy() for t in Dg]), np.stack([t.numpy() for t in Ig]),
                              self.offsets, self.metric_type)


def _device_view_i64(ptr: int, n: int, device):
    """A torch int64 view of `n` 8-byte words of device memory owned by the C library."""
    import torch

    class _Ext:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 3,
                                    "strides": None}
    return torch.as_tensor(_Ext(), device=device)
