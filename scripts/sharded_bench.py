"""Deep100M-shape sharded config (BASELINE configs[3]): N x (n_shard x d) database, one sub-graph per
GPU, queries broadcast, per-shard top-k exchanged by the C-ABI's bh_shards_* path (peer stores over NVLink +
flag barrier + merge kernel; falls back to one packed NCCL all-gather), and the same shard searched alone as
the efficiency denominator. Launch with torchrun.
Each rank generates its own shard on the device (same recipe, per-rank seed offset in the latent
draw but ONE shared projection, so all shards are samples of the same distribution)."""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hnsw_b200
from hnsw_b200.datasets import exact_knn_torch
from hnsw_b200.sharded import ShardedIndexHNSWFlat

ap = argparse.ArgumentParser()
ap.add_argument("--n_shard", type=int, default=12_500_000)
ap.add_argument("--dim", dest="d", type=int, default=96)
ap.add_argument("--latent", dest="d1", type=int, default=12)
ap.add_argument("--M", type=int, default=32)
ap.add_argument("--efc", type=int, default=200)
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--efs", type=str, default="32,64,128")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--ip", type=int, default=0)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

g = torch.Generator(device=dev); g.manual_seed(1338)            # shared projection / scale / queries
proj = torch.rand(a.d1, a.d, generator=g, device=dev)
scale = torch.rand(a.d, generator=g, device=dev) * 4 + 0.1
xq = torch.sin((torch.randn(a.nq, a.d1, generator=g, device=dev) @ proj) * scale).contiguous()
if a.ip:
    xq = (xq / xq.norm(dim=1, keepdim=True)).contiguous()
g2 = torch.Generator(device=dev); g2.manual_seed(7000 + rank)    # this rank's shard
xb = torch.empty(a.n_shard, a.d, device=dev)
for i0 in range(0, a.n_shard, 1 << 20):
    i1 = min(a.n_shard, i0 + (1 << 20))
    xb[i0:i1] = torch.sin((torch.randn(i1 - i0, a.d1, generator=g2, device=dev) @ proj) * scale)
    if a.ip:
        xb[i0:i1] /= xb[i0:i1].norm(dim=1, keepdim=True)
xb_h = xb.cpu().numpy()

sh = ShardedIndexHNSWFlat(a.d, a.M, 0 if a.ip else 1, device=dev)
sh.local.hnsw.efConstruction = a.efc
torch.cuda.synchronize(); dist.barrier()
t0 = time.time()
sh.add(xb_h)
torch.cuda.synchronize(); dist.barrier()
t_build = time.time() - t0
del xb_h

# global ground truth: per-shard exact top-k -> all-gather -> merge
Dl, Il = exact_knn_torch(xb, xq, a.k, inner_product=bool(a.ip), chunk=1 << 17)
Il = Il + int(sh.offsets[rank])
Dall = torch.empty(world, a.nq, a.k, device=dev); Iall = torch.empty(world, a.nq, a.k, dtype=torch.int64, device=dev)
dist.all_gather_into_tensor(Dall, Dl.contiguous()); dist.all_gather_into_tensor(Iall, Il.contiguous())
sel = torch.topk(Dall.permute(1, 0, 2).reshape(a.nq, -1), a.k, dim=1, largest=bool(a.ip)).indices
gt = torch.gather(Iall.permute(1, 0, 2).reshape(a.nq, -1), 1, sel).cpu().numpy()
del xb, Dall, Iall
torch.cuda.empty_cache()

rows = []
for ef in [int(e) for e in a.efs.split(",")]:
    for _ in range(3):
        D, I = sh.search(xq, a.k, efSearch=ef)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        D, I = sh.search(xq, a.k, efSearch=ef)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    Ih = I.cpu().numpy()
    rec = float(np.mean([len(set(Ih[i].tolist()) & set(gt[i].tolist())) for i in range(a.nq)])) / a.k
    # the same shard alone: no exchange, no merge
    D1 = torch.empty(a.nq, a.k, device=dev); I1 = torch.empty(a.nq, a.k, dtype=torch.int64, device=dev)
    st1 = torch.cuda.ExternalStream(sh.local.stream_ptr, device=dev)
    for _ in range(3):
        sh.local.search_device(xq.data_ptr(), a.nq, a.k, D1.data_ptr(), I1.data_ptr(), efSearch=ef)
    sh.local.synchronize(); dist.barrier()
    iso = 0.0
    for _ in range(a.steps):          # isolated launches: the like-for-like denominator of a synchronous sharded step
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(st1)
        sh.local.search_device(xq.data_ptr(), a.nq, a.k, D1.data_ptr(), I1.data_ptr(), efSearch=ef)
        g1.record(st1); sh.local.synchronize()
        iso += g0.elapsed_time(g1)
    ms1 = torch.tensor([iso / a.steps], device=dev)
    dist.all_reduce(ms1, op=dist.ReduceOp.MAX)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(st1)
    for _ in range(a.steps):          # back-to-back launches overlap their drain phases (pipelined figure)
        sh.local.search_device(xq.data_ptr(), a.nq, a.k, D1.data_ptr(), I1.data_ptr(), efSearch=ef)
    f1.record(st1); sh.local.synchronize()
    ms1p = torch.tensor([f0.elapsed_time(f1) / a.steps], device=dev)
    dist.all_reduce(ms1p, op=dist.ReduceOp.MAX)
    # pipelined serving loop (bh_shards_set_pipelined): enqueue back to back, join once
    Dp = torch.empty(a.nq, a.k, device=dev); Ip = torch.empty(a.nq, a.k, dtype=torch.int64, device=dev)
    sh.set_pipelined(True)
    for _ in range(3):
        sh.enqueue(xq, a.k, Dp, Ip, efSearch=ef)
    sh.join(st1); sh.local.synchronize(); dist.barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(st1)
    for _ in range(a.steps):
        sh.enqueue(xq, a.k, Dp, Ip, efSearch=ef)
    sh.join(st1); p1.record(st1); sh.local.synchronize()
    msp = torch.tensor([p0.elapsed_time(p1) / a.steps], device=dev)
    dist.all_reduce(msp, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(Ip, I)) and bool(torch.equal(Dp, D))
    sh.set_pipelined(False)
    dist.barrier()
    rows.append({"pipelined_ms_per_batch": round(float(msp.item()), 3), "pipelined_qps": round(a.nq / float(msp.item()) * 1e3),
                 "pipelined_equals_synchronous": same,
                 "pipelined_efficiency_vs_pipelined_single_shard": round(float(ms1p.item()) / float(msp.item()), 4),
                 "efSearch": ef, "ms_per_batch": round(float(ms.item()), 3), "qps": round(a.nq / float(ms.item()) * 1e3),
                 "recall_at_10": round(rec, 4), "ms_shard_alone_max_over_ranks": round(float(ms1.item()), 3),
                 "efficiency_vs_single_shard": round(float(ms1.item()) / float(ms.item()), 4),
                 "ms_shard_alone_pipelined": round(float(ms1p.item()), 3),
                 "efficiency_vs_pipelined_single_shard": round(float(ms1p.item()) / float(ms.item()), 4)})
if rank == 0:
    print(json.dumps({"config": f"sharded {world} x {a.n_shard} x {a.d} fp32 {'IP' if a.ip else 'L2'}, M={a.M} efC={a.efc}, {a.nq} queries broadcast, "
                                f"exchange = {sh.exchange_kind} ({a.nq * a.k * 8} B/rank) + merge kernel",
                      "db_vectors": world * a.n_shard, "n_gpus": world,
                      "build_s": round(t_build, 2), "build_vectors_per_s_total": round(world * a.n_shard / t_build),
                      "search": rows}), flush=True)
dist.destroy_process_group()
