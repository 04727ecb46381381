"""ShardedIndexHNSWFlat — the database sharded over the ranks of a torch.distributed group.

Semantics of faiss IndexShards(successive_ids=True) + merge_knn_results:
rank r owns the contiguous id range [offset_r, offset_r + ntotal_r) and an independent HNSW graph
over it; a search broadcasts the queries, every rank searches its shard, the per-shard sorted
top-k lists (distance, LOCAL id) are exchanged with ONE all-gather each, and every rank merges
them into the global top-k with ids shifted by the owning shard's offset.

On GPUs (NCCL backend) the local index is hnsw_b200.IndexHNSWFlat and the merge is the CUDA warp
top-k merge kernel (bh_merge_topk_device); nothing on that path runs on the CPU. The local index
and the merge function can be injected, which is how the host-side bookkeeping (offsets, gather
layout, broadcast) is tested on CPU with the gloo backend and the CPU oracle (tests/).
"""
from __future__ import annotations

import numpy as np


class ShardedIndexHNSWFlat:
    def __init__(self, d: int, M: int = 32, metric: int = 1, group=None, device=None,
                 local_index=None, merge_fn=None):
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

    # ---- bookkeeping shared by both paths
    def _refresh_offsets(self):
        import torch
        t = torch.tensor([int(self.local.ntotal)], dtype=torch.int64,
                         device=self.device if self._on_gpu() else "cpu")
        out = [torch.zeros_like(t) for _ in range(self.world)]
        self._dist.all_gather(out, t, group=self.group)
        self.ntotals = np.array([int(x.item()) for x in out], np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(self.ntotals)[:-1]]).astype(np.int64)

    def _on_gpu(self):
        return self._merge_fn is None

    @property
    def ntotal(self) -> int:
        return int(self.ntotals.sum())

    def add(self, x_local):
        """Add this rank's slice of the database (independent sub-graph; no communication)."""
        self.local.add(x_local)
        self._refresh_offsets()

    def search(self, xq, k: int, efSearch: int | None = None, src: int = 0):
        """xq: [nq, d] float32 on rank `src` (other ranks pass an array of the same shape).
        Returns (D, I) with global ids on every rank."""
        import torch
        dist = self._dist
        nq = int(xq.shape[0])
        if self._on_gpu():
            from .index import merge_topk_device
            q = torch.as_tensor(np.ascontiguousarray(xq, np.float32)).to(self.device) \
                if not isinstance(xq, torch.Tensor) else xq.to(self.device, torch.float32).contiguous()
            dist.broadcast(q, src=src, group=self.group)
            Dl = torch.empty(nq, k, device=self.device)
            Il = torch.empty(nq, k, dtype=torch.int64, device=self.device)
            cur = torch.cuda.current_stream(self.device)
            ist = torch.cuda.ExternalStream(self.local.stream_ptr, device=self.device)
            ist.wait_stream(cur)
            if self.local.ntotal == 0:
                # an empty shard contributes empty lists (faiss pads with +/-FLT_MAX, -1); raising here
                # would leave the peers blocked in the collective below
                Dl.fill_(3.4028234663852886e38 if self.metric_type == 1 else -3.4028234663852886e38)
                Il.fill_(-1)
            else:
                self.local.search_device(q.data_ptr(), nq, k, Dl.data_ptr(), Il.data_ptr(), efSearch=efSearch or 0)
            cur.wait_stream(ist)
            Dg = torch.empty(self.world, nq, k, device=self.device)
            Ig = torch.empty(self.world, nq, k, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(Dg, Dl, group=self.group)
            dist.all_gather_into_tensor(Ig, Il, group=self.group)
            Dm = torch.empty(nq, k, device=self.device)
            Im = torch.empty(nq, k, dtype=torch.int64, device=self.device)
            merge_topk_device(Dg.data_ptr(), Ig.data_ptr(), self.world, nq, k, self.metric_type, self.offsets,
                              Dm.data_ptr(), Im.data_ptr(), cur.cuda_stream)
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
