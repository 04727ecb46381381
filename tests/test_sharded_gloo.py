"""world_size-2 gloo test (CPU) of the sharded search bookkeeping: contiguous id ranges, query
broadcast, all-gather layout, global-id shift — with the CPU oracle as each rank's local index and
a numpy merge standing in for the CUDA merge kernel (which has its own GPU test)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def numpy_merge(D_all, I_all, offsets, metric):
    ns, nq, k = D_all.shape
    allD = D_all.transpose(1, 0, 2).reshape(nq, ns * k)
    allI = np.where(I_all >= 0, I_all + offsets[:, None, None], -1).transpose(1, 0, 2).reshape(nq, ns * k)
    order = np.argsort(allD if metric == 1 else -allD, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(allD, order, 1), np.take_along_axis(allI, order, 1)


class _LocalOracle:
    def __init__(self, d, M, metric):
        from oracle import oracle as om
        self.o = om.OracleHNSWFlat(d, M, metric)

    @property
    def ntotal(self):
        return self.o.ntotal

    def add(self, x):
        self.o.add(x)

    def search(self, xq, k, ef):
        return self.o.search(xq, k, ef)


def _worker(rank, world, port, metric, out_q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from hnsw_b200.datasets import synthetic_dataset
    from hnsw_b200.sharded import ShardedIndexHNSWFlat
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        d, M, k = 32, 8, 10
        xb, xq = synthetic_dataset(d, 3000, 40, normalize=(metric == 0))
        sizes = [1700, 1300]                      # ragged shards
        lo = sum(sizes[:rank])
        sh = ShardedIndexHNSWFlat(d, M, metric, local_index=_LocalOracle(d, M, metric), merge_fn=numpy_merge)
        sh.add(xb[lo:lo + sizes[rank]])
        assert sh.ntotal == 3000 and sh.offsets.tolist() == [0, 1700]
        # only rank 0 holds the real queries; the others pass a placeholder
        q_in = xq if rank == 0 else np.zeros_like(xq)
        D, I = sh.search(q_in, k, efSearch=400)
        out_q.put((rank, D, I))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(metric):
    from hnsw_b200.datasets import synthetic_dataset
    from oracle import oracle as om
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, metric, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict()
    for _ in range(2):
        r, D, I = q.get(timeout=120)
        res[r] = (D, I)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xb, xq = synthetic_dataset(32, 3000, 40, normalize=(metric == 0))
    _, gt = om.brute_force_knn(xb, xq, 10, metric)
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][0], res[1][0])  # same on all ranks
    D, I = res[0]
    assert om.recall_at_k(I, gt) >= 0.99          # ef=400 on 1.5k-point shards is near-exhaustive
    assert I.max() >= 1700                         # ids from the second shard carry its offset
    ref = ((xq[:, None, :] - xb[I]) ** 2).sum(-1) if metric == 1 else (xq[:, None, :] * xb[I]).sum(-1)
    assert np.allclose(D, ref, rtol=1e-4, atol=1e-5)
    assert np.all(np.diff(D, axis=1) >= 0) if metric == 1 else np.all(np.diff(D, axis=1) <= 0)


def test_sharded_search_two_ranks_l2():
    _run(1)


def test_sharded_search_two_ranks_ip():
    _run(0)
