"""Does ordering the query batch by graph locality raise L2 reuse? (host-side ordering only)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hnsw_b200
from hnsw_b200.datasets import synthetic_dataset
nq_pool = 80000
xb, xq_all = synthetic_dataset(128, 1000000, nq_pool, d1=12)
idx = hnsw_b200.IndexHNSWFlat(128, 32)
idx.hnsw.efConstruction = 200
idx.add(xb)
lv = idx.export_graph()["levels"]
xb_t = torch.from_numpy(xb).cuda()
def order_by(level_min, xq):
    ids = np.flatnonzero(lv >= level_min + 1)
    c = xb_t[torch.from_numpy(ids).cuda()]
    q = torch.from_numpy(xq).cuda()
    d = (q * q).sum(1, keepdim=True) - 2 * q @ c.T + (c * c).sum(1)[None]
    key = d.argmin(1).cpu().numpy()
    return np.argsort(key, kind="stable"), len(ids)
def timeit(xq, ef):
    ms = []
    for _ in range(4):
        idx.search(xq, 10, efSearch=ef)
        ms.append(idx.last_search_ms)
    return min(ms)
for nq in (10000, 80000):
    xq = xq_all[:nq]
    for ef in (64, 128):
        base = timeit(xq, ef)
        line = f"nq={nq} ef={ef}: unsorted {base:.3f} ms ({nq/base*1e3/1e6:.2f} M QPS)"
        for lm in (3, 2, 1):
            perm, nc = order_by(lm, xq)
            t = timeit(np.ascontiguousarray(xq[perm]), ef)
            line += f" | by level>={lm} ({nc} cells) {t:.3f} ms ({nq/t*1e3/1e6:.2f} M)"
        print(line, flush=True)
