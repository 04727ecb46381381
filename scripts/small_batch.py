"""Latency regime: few work items on an idle GPU. Search a 1M graph with tiny query batches."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hnsw_b200
f = "/tmp/bh_graph_1000000_128_12_32_200.npz"
idx = hnsw_b200.IndexHNSWFlat(128, 32)
if os.path.exists(f):
    z = np.load(f)
    idx.import_graph(z["xb"], z["levels"], z["neighbors"], int(z["entry_point"]), int(z["max_level"]))
    xq = z["xq"]
else:
    from hnsw_b200.datasets import synthetic_dataset
    xb, xq = synthetic_dataset(128, 200000, 10000, d1=12)
    idx.hnsw.efConstruction = 200
    idx.add(xb)
for nq in (32, 256, 1024):
    for ef in (64, 200):
        for W in (1, 2, 4, 8):
            ms = []
            for _ in range(5):
                D, I, S = idx.search(xq[:nq], 10, efSearch=ef, warps_per_query=W, stats=True)
                ms.append(idx.last_search_ms)
            hops = S[:, 1].mean()
            print(f"nq={nq:5d} ef={ef:3d} W={W} ms={min(ms):7.3f}  hops={hops:.0f}  us/hop(max-ish)={min(ms)*1e3/S[:,1].max():.2f}")
