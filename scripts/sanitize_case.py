"""Small build + search used under compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hnsw_b200
from hnsw_b200.datasets import synthetic_dataset
for d, M, metric, n in ((32, 8, 1, 1500), (96, 16, 0, 1200), (512, 4, 1, 400)):
    xb, xq = synthetic_dataset(d, n, 64, normalize=(metric == 0))
    idx = hnsw_b200.IndexHNSWFlat(d, M, metric)
    idx.hnsw.efConstruction = 32
    idx.set_build_params(max_batch=1)
    idx.add(xb[:200])            # sequential rounds
    idx.set_build_params(max_batch=0)
    idx.add(xb[200:])            # batched rounds (shrinks, chains)
    for ef, W, hb in ((16, 0, 0), (64, 2, 0), (40, 4, 8), (200, 1, 0)):
        D, I = idx.search(xq, 10, efSearch=ef, warps_per_query=W, hash_bits=hb)
        assert (I >= 0).all()
    g = idx.export_graph()
    print(d, M, metric, "ok", int((g["neighbors"] >= 0).sum()))
print("done")
