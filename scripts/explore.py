"""GPU exploration: build-schedule vs recall, and search QPS / roofline per (ef, W, hash_bits).
Not part of the product or the tests; prints tables to stdout."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hnsw_b200  # noqa: E402
from hnsw_b200.datasets import exact_knn_torch, synthetic_dataset_torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--d1", type=int, default=32)
ap.add_argument("--M", type=int, default=32)
ap.add_argument("--efc", type=int, default=200)
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--divisors", type=str, default="0")
ap.add_argument("--max_batch", type=str, default="8192")
ap.add_argument("--efs", type=str, default="16,32,64,128,256,512")
ap.add_argument("--Ws", type=str, default="0")
ap.add_argument("--hbs", type=str, default="0")
ap.add_argument("--build_W", type=int, default=0)
ap.add_argument("--build_hb", type=int, default=0)
ap.add_argument("--cpu", type=int, default=0, help="also build with the CPU oracle using this many threads")
ap.add_argument("--ip", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--variants", type=str, default="auto")
a = ap.parse_args()

metric = 0 if a.ip else 1
xb_t, xq_t = synthetic_dataset_torch(a.d, a.n, a.nq, d1=a.d1, normalize=bool(a.ip))
_, gt = exact_knn_torch(xb_t, xq_t, 10, inner_product=bool(a.ip))
gt = gt.cpu().numpy()
xb, xq = xb_t.cpu().numpy(), xq_t.cpu().numpy()
del xb_t, xq_t
torch.cuda.empty_cache()


def recall(I):
    return float(np.mean([len(set(I[i].tolist()) & set(gt[i].tolist())) for i in range(len(gt))])) / 10


PEAK = 6530.0
if a.cpu:
    from oracle import oracle as om
    o = om.OracleHNSWFlat(a.d, a.M, metric)
    o.efConstruction = a.efc
    o.threads = a.cpu
    t = time.time()
    o.add(xb)
    tb = time.time() - t
    print(f"CPU oracle build {a.cpu}t: {tb:.1f}s = {a.n / tb:.0f} vec/s")
    for ef in [int(e) for e in a.efs.split(",")]:
        t = time.time()
        D, I = o.search(xq, 10, ef)
        ts = time.time() - t
        print(f"  cpu ef={ef:4d} recall={recall(I):.4f} qps={a.nq / ts:.0f}")

for div in [int(x) for x in a.divisors.split(",")]:
    for mb in [int(x) for x in a.max_batch.split(",")]:
        idx = hnsw_b200.IndexHNSWFlat(a.d, a.M, metric)
        idx.hnsw.efConstruction = a.efc
        idx.set_build_params(max_batch=mb, batch_divisor=div, warps_per_query=a.build_W, hash_bits=a.build_hb)
        l0 = hnsw_b200.launch_count()
        t = time.time()
        idx.add(xb)
        tb = time.time() - t
        print(f"GPU build div={div} max_batch={mb}: wall {tb:.2f}s device {idx.last_build_ms / 1e3:.2f}s = "
              f"{a.n / (idx.last_build_ms / 1e3):.0f} vec/s launches={hnsw_b200.launch_count() - l0}")
        for ef in [int(e) for e in a.efs.split(",")]:
            for W in [int(x) for x in a.Ws.split(",")]:
                for hb, var in [(int(x), v) for x in a.hbs.split(",") for v in a.variants.split(",")]:
                    if var == "auto":
                        os.environ.pop("BH_BEAM_VARIANT", None)
                    else:
                        os.environ["BH_BEAM_VARIANT"] = var
                    try:
                        D, I, S = idx.search(xq, 10, efSearch=ef, stats=True, warps_per_query=W, hash_bits=hb)
                    except RuntimeError as e:
                        print(f"  ef={ef} W={W} hb={hb}: {e}")
                        continue
                    ms = []
                    for _ in range(a.reps):
                        idx.search(xq, 10, efSearch=ef, warps_per_query=W, hash_bits=hb)
                        ms.append(idx.last_search_ms)
                    ms = min(ms)
                    s = S.astype(np.float64).mean(0)
                    bq = s[0] * 4 * a.d + s[1] * 8 * a.M + s[2] * 4 * a.d + s[3] * 4 * a.M + 4 * a.d + 12 * 10
                    gbs = bq * a.nq / (ms * 1e-3) / 1e9
                    print(f"  ef={ef:4d} W={W} hb={hb:2d} v={var} recall={recall(I):.4f} ms={ms:8.3f} qps={a.nq / ms * 1e3:10.0f} "
                          f"ndis={s[0]:.0f} nhops={s[1]:.0f} up=({s[2]:.0f},{s[3]:.0f}) B/q={bq / 1e3:.0f}KB "
                          f"gather={gbs:.0f}GB/s frac={gbs / PEAK:.3f}")
        del idx
