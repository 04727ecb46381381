"""Generates tests/golden/oracle_l2_d32_n3000_M16.npz from the CPU oracle.

There is no reference implementation to generate goldens from (/root/reference holds only
README.md and LICENSE; faiss is not installed), so this fixture freezes the oracle's own output:
graph + search results + per-query counters for a small seeded case. Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hnsw_b200.datasets import synthetic_dataset  # noqa: E402
from oracle import oracle as om  # noqa: E402

xb, xq = synthetic_dataset(32, 3000, 64)
o = om.OracleHNSWFlat(32, 16)
o.set_team(8)
o.add(xb)
g = o.export_graph()
D, I, st = o.search(xq, 10, 48, stats=True)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_l2_d32_n3000_M16.npz")
np.savez_compressed(out, levels=g["levels"], neighbors=g["neighbors"], entry_point=g["entry_point"],
                    max_level=g["max_level"], D=D, I=I, stats=st)
print("wrote", out, os.path.getsize(out), "bytes")
