"""Under torchrun (N >= 2 GPUs): the two exchange kinds of ShardedIndexHNSWFlat — peer stores (CUDA IPC) and the
packed single NCCL all-gather fallback — must return the same merged (D, I), equal to an exact host-side merge."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hnsw_b200.datasets import synthetic_dataset
from hnsw_b200.sharded import ShardedIndexHNSWFlat
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d, n_sh, nq, k = 64, 30000, 2000, 10
xb, xq = synthetic_dataset(d, n_sh * world, nq)
res = {}
for kind in ("peer-store", "nccl-allgather-packed"):
    sh = ShardedIndexHNSWFlat(d, 16, 1, device=dev, exchange=kind)
    sh.add(xb[rank * n_sh:(rank + 1) * n_sh])
    D, I = sh.search(xq, k, efSearch=48, keep_local=True)
    torch.cuda.synchronize()
    assert sh.exchange_kind == kind, sh.exchange_kind
    Dl, Il = sh.last_local
    gD = [torch.empty_like(Dl) for _ in range(world)]; gI = [torch.empty_like(Il) for _ in range(world)]
    dist.all_gather(gD, Dl); dist.all_gather(gI, Il)
    hD = torch.cat(gD, 1); hI = torch.cat([torch.where(g >= 0, g + r * n_sh, g) for r, g in enumerate(gI)], 1)
    ho = torch.argsort(hD, dim=1, stable=True)[:, :k]
    assert torch.equal(I, torch.gather(hI, 1, ho)) and torch.equal(D, torch.gather(hD, 1, ho)), kind
    res[kind] = (D.cpu().numpy(), I.cpu().numpy())
    del sh
assert np.array_equal(res["peer-store"][1], res["nccl-allgather-packed"][1])
assert np.array_equal(res["peer-store"][0], res["nccl-allgather-packed"][0])
if rank == 0:
    print(f"{world} ranks: peer-store == nccl-allgather-packed == exact host merge ({nq} queries, k={k})")
dist.destroy_process_group()
