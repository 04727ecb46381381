"""GPU sweep of the visited-table policy / size on the bench workload (1M x 128, M=32, efC=200):
QPS (CUDA events around search_device), mean ndis, re-scored fraction vs the exact table, and the
honest roofline fraction (bytes of the EXACT traversal / time). One JSON line per configuration.
Not part of the product or the tests."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hnsw_b200  # noqa: E402
from hnsw_b200.datasets import exact_knn_torch, synthetic_dataset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--d1", type=int, default=12)
ap.add_argument("--M", type=int, default=32)
ap.add_argument("--efc", type=int, default=200)
ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--efs", type=str, default="32,64,128,256,512")
ap.add_argument("--bits", type=str, default="9,10,11,12,13")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--build-bits", type=str, default="", help="also time builds with these assoc table sizes")
a = ap.parse_args()

PEAK = 6530.0
dev = torch.device("cuda", 0)
xb, xq_all = synthetic_dataset(a.d, a.n, 8 * a.nq, d1=a.d1, seed=1338)
xq = np.ascontiguousarray(xq_all[:a.nq])
xb_t, xq_t = torch.from_numpy(xb).to(dev), torch.from_numpy(xq).to(dev)
_, gt = exact_knn_torch(xb_t, xq_t, 10)
gt = gt.cpu().numpy()
del xb_t
torch.cuda.empty_cache()


def recall(I):
    return float(np.mean([len(set(I[i].tolist()) & set(gt[i].tolist())) for i in range(len(gt))])) / 10


def build(policy=0, bits=0):
    idx = hnsw_b200.IndexHNSWFlat(a.d, a.M)
    idx.hnsw.efConstruction = a.efc
    idx.set_build_params(hash_bits=bits, visited_policy=policy)
    idx.add(xb)
    return idx


idx = build()
print(json.dumps({"build": "default", "device_s": round(idx.last_build_ms / 1e3, 3), "counters": idx.last_build_counters}),
      flush=True)
D_d = torch.empty(a.nq, 10, device=dev)
I_d = torch.empty(a.nq, 10, dtype=torch.int64, device=dev)
S_d = torch.zeros(a.nq, 4, dtype=torch.int32, device=dev)
stream = torch.cuda.ExternalStream(idx.stream_ptr, device=dev)


def run(ef, policy, bits):
    idx.search_device(xq_t.data_ptr(), a.nq, 10, D_d.data_ptr(), I_d.data_ptr(), efSearch=ef,
                      stats_ptr=S_d.data_ptr(), hash_bits=bits, visited_policy=policy)
    idx.synchronize()
    st = S_d.cpu().numpy().astype(np.float64).mean(0)
    best = 1e9
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        idx.search_device(xq_t.data_ptr(), a.nq, 10, D_d.data_ptr(), I_d.data_ptr(), efSearch=ef,
                          hash_bits=bits, visited_policy=policy)
        e1.record(stream)
        idx.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return st, best, I_d.cpu().numpy().copy()


def bytes_q(st):
    return (st[0] + st[2]) * 4 * a.d + st[1] * 8 * a.M + st[3] * 4 * a.M + 4 * a.d + 120


for ef in [int(e) for e in a.efs.split(",")]:
    st_x, ms_x, I_x = run(ef, 1, 15)   # exact: 32768 slots never fill below ndis ~ 24k
    bx = bytes_q(st_x)
    rec = recall(I_x)
    print(json.dumps({"ef": ef, "policy": "exact(2^15)", "ndis": round(st_x[0], 1), "nhops": round(st_x[1], 1),
                      "ms": round(ms_x, 3), "recall": round(rec, 4)}), flush=True)
    cfgs = [("r1-reset(auto)", 1, 0)] + [(f"assoc(bits={b})", 2, int(b)) for b in a.bits.split(",")] + [("auto", 0, 0)]
    for name, pol, bits in cfgs:
        st, ms, I = run(ef, pol, bits)
        assert np.array_equal(I, I_x), (ef, name)
        print(json.dumps({"ef": ef, "policy": name, "ndis": round(st[0], 1), "x_exact": round(st[0] / st_x[0], 4),
                          "ms": round(ms, 3), "qps": round(a.nq / ms * 1e3),
                          "frac_touched": round(bytes_q(st) * a.nq / ms / 1e6 / PEAK, 4),
                          "frac_exact": round(bx * a.nq / ms / 1e6 / PEAK, 4)}), flush=True)

for b in [int(x) for x in a.build_bits.split(",") if x]:
    for pol, bits, name in ((1, 0, "r1-reset(auto)"), (2, b, f"assoc(bits={b})")):
        if pol == 1 and b != int(a.build_bits.split(",")[0]):
            continue
        i2 = build(pol, bits)
        I = i2.search(xq, 10, efSearch=64)[1]
        print(json.dumps({"build": name, "device_s": round(i2.last_build_ms / 1e3, 3), "recall_ef64": round(recall(I), 4),
                          "counters": i2.last_build_counters}), flush=True)
        del i2
