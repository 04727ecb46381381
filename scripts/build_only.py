"""Build one index on the GPU (the command ncu wraps for the build launch list)."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hnsw_b200
from hnsw_b200.datasets import synthetic_dataset
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--d1", type=int, default=12)
ap.add_argument("--M", type=int, default=32)
ap.add_argument("--efc", type=int, default=200)
ap.add_argument("--ip", type=int, default=0)
a = ap.parse_args()
xb, _ = synthetic_dataset(a.d, a.n, 1, d1=a.d1, normalize=bool(a.ip))
idx = hnsw_b200.IndexHNSWFlat(a.d, a.M, 0 if a.ip else 1)
idx.hnsw.efConstruction = a.efc
t = time.time(); idx.add(xb); t = time.time() - t
print(f"build n={a.n} d={a.d} wall {t:.2f}s device {idx.last_build_ms/1e3:.2f}s = {a.n/(idx.last_build_ms/1e3):.0f} vec/s")
