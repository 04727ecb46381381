"""One search configuration on a cached GPU-built graph — the command ncu wraps.
First run builds the graph and caches it in /tmp (within one gpurun call); later runs import it."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hnsw_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--d1", type=int, default=16)
ap.add_argument("--M", type=int, default=32)
ap.add_argument("--efc", type=int, default=200)
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--ef", type=int, default=64)
ap.add_argument("--W", type=int, default=0)
ap.add_argument("--hb", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cache", type=str, default="/tmp/bh_graph")
a = ap.parse_args()

f = f"{a.cache}_{a.n}_{a.d}_{a.d1}_{a.M}_{a.efc}.npz"
idx = hnsw_b200.IndexHNSWFlat(a.d, a.M)
if os.path.exists(f):
    z = np.load(f)
    xb, xq = z["xb"], z["xq"]
    idx.import_graph(xb, z["levels"], z["neighbors"], int(z["entry_point"]), int(z["max_level"]))
else:
    from hnsw_b200.datasets import synthetic_dataset_torch
    xb_t, xq_t = synthetic_dataset_torch(a.d, a.n, a.nq, d1=a.d1)
    xb, xq = xb_t.cpu().numpy(), xq_t.cpu().numpy()
    idx.hnsw.efConstruction = a.efc
    idx.add(xb)
    g = idx.export_graph()
    np.savez(f, xb=xb, xq=xq, levels=g["levels"], neighbors=g["neighbors"], entry_point=g["entry_point"],
             max_level=g["max_level"])
    print("built + cached", f, "build ms", idx.last_build_ms)
xq = xq[:a.nq]
for _ in range(a.reps):
    D, I, S = idx.search(xq, 10, efSearch=a.ef, stats=True, warps_per_query=a.W, hash_bits=a.hb)
    s = S.astype(np.float64).mean(0)
    bq = s[0] * 4 * a.d + s[1] * 8 * a.M + s[2] * 4 * a.d + s[3] * 4 * a.M + 4 * a.d + 120
    print(f"ef={a.ef} W={a.W} hb={a.hb} ms={idx.last_search_ms:.3f} qps={len(xq) / idx.last_search_ms * 1e3:.0f} "
          f"ndis={s[0]:.0f} nhops={s[1]:.0f} gather={bq * len(xq) / idx.last_search_ms / 1e6:.0f} GB/s")
