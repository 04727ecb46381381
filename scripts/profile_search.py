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
ap.add_argument("--bench-data", type=int, default=0,
                help="1: bench.py's data recipe for --config sift/gist/... (numpy, seed 1338, 8 x nq query pool), so "
                     "the ncu capture is of the very launch bench.py times")
a = ap.parse_args()

f = f"{a.cache}_{a.n}_{a.d}_{a.d1}_{a.M}_{a.efc}_{a.bench_data}.npz"
idx = hnsw_b200.IndexHNSWFlat(a.d, a.M)
if os.path.exists(f):
    z = np.load(f)
    xb, xq = z["xb"], z["xq"]
    idx.import_graph(xb, z["levels"], z["neighbors"], int(z["entry_point"]), int(z["max_level"]))
else:
    if a.bench_data:
        from hnsw_b200.datasets import synthetic_dataset
        xb, xq = synthetic_dataset(a.d, a.n, 8 * a.nq, d1=a.d1, seed=1338)
        xq = np.ascontiguousarray(xq[:a.nq])
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
import torch  # noqa: E402
dev = torch.device("cuda", 0)
xq_t = torch.from_numpy(np.ascontiguousarray(xq)).to(dev)
D_d = torch.empty(len(xq), 10, device=dev)
I_d = torch.empty(len(xq), 10, dtype=torch.int64, device=dev)
S_d = torch.zeros(len(xq), 4, dtype=torch.int32, device=dev)
stream = torch.cuda.ExternalStream(idx.stream_ptr, device=dev)
for _ in range(a.reps):  # ONE launch over the whole batch per rep (what bench.py times)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    idx.search_device(xq_t.data_ptr(), len(xq), 10, D_d.data_ptr(), I_d.data_ptr(), efSearch=a.ef,
                      stats_ptr=S_d.data_ptr(), warps_per_query=a.W, hash_bits=a.hb)
    e1.record(stream)
    idx.synchronize()
    ms = e0.elapsed_time(e1)
    s = S_d.cpu().numpy().astype(np.float64).mean(0)
    bq = s[0] * 4 * a.d + s[1] * 8 * a.M + s[2] * 4 * a.d + s[3] * 4 * a.M + 4 * a.d + 120
    print(f"ef={a.ef} W={a.W} hb={a.hb} ms={ms:.3f} qps={len(xq) / ms * 1e3:.0f} "
          f"ndis={s[0]:.0f} nhops={s[1]:.0f} gather={bq * len(xq) / ms / 1e6:.0f} GB/s")
