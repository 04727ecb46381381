"""Debug: per-phase cycles of a hop (needs the library built with EXTRA=-DBH_PHASE_TIMING)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hnsw_b200
from hnsw_b200.datasets import synthetic_dataset
xb, xq = synthetic_dataset(128, 1000000, 10000, d1=12)
idx = hnsw_b200.IndexHNSWFlat(128, 32)
idx.hnsw.efConstruction = 200
idx.add(xb)
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 8
idx.search(xq[:nq], 10, efSearch=64, warps_per_query=1)            # warm
print("=== timed", flush=True)
hb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
idx.search(xq[:nq], 10, efSearch=64, warps_per_query=1, stats=True, hash_bits=hb)
