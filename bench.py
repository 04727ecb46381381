#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 HNSW engine (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's engine
  python bench.py --impl reference --gpus N ...            # CPU restatement of faiss IndexHNSWFlat
  python bench.py --config {sift,gist,deep,ip768} ...      # the other BASELINE shapes (default: sift)

Headline workload (BASELINE configs[1], `--config sift`): SIFT1M-shape 1M x 128 fp32 L2, M=32,
efConstruction=200, 10k-query batch, k=10; efSearch swept 16..512; the headline value is QPS at the
smallest swept efSearch whose recall@10 (vs exact brute force) is >= 0.95. A "step" = one pass of the
search path over the 10k-query batch. Build vectors/sec (add() on the same data) is reported beside it.
Data: synthetic (faiss SyntheticDataset recipe, seed 1338; d1 per config — see DESIGN.md §6).

N > 1: one process per GPU (torchrun). Headline = replicas (every GPU holds a 1M-vector index and
answers its own 10k queries; weak scaling, no data-path collective). The same run also measures the
north-star SHARDED path at fixed per-GPU shard size ("sharded_weak": each rank's shard is one slice of
an N x n database, queries broadcast, per-shard top-k written straight into every peer's gather buffer
over NVLink by the traversal kernel's epilogue, one flag barrier, warp top-k merge) against exact
global ground truth.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EF_GRID = [16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512]
SEED = 1338
L2_POLICY = "no flush: the index and the bytes touched per step both exceed the 126 MB L2"

# BASELINE.json configs[1..4] at single-GPU size. d1 sets the difficulty of the synthetic manifold; it is
# chosen so that recall@10 >= 0.95 is reachable inside EF_GRID.
CONFIGS = {
    "sift": dict(shape="SIFT1M-shape", n=1_000_000, d=128, d1=12, ip=False, M=32, efc=200, nq=10_000,
                 metric="QPS at recall@10>=0.95 (1M x128 L2)"),
    "gist": dict(shape="GIST1M-shape", n=1_000_000, d=960, d1=16, ip=False, M=32, efc=200, nq=10_000,
                 metric="QPS at recall@10>=0.95 (1M x960 L2)"),
    "deep": dict(shape="Deep100M-shape shard (1/8 of 100M)", n=12_500_000, d=96, d1=12, ip=False, M=32, efc=200,
                 nq=10_000, metric="QPS at recall@10>=0.95 (12.5M x96 L2 shard)"),
    "ip768": dict(shape="embedding-shape", n=1_000_000, d=768, d1=16, ip=True, M=32, efc=200, nq=10_000,
                  metric="QPS at recall@10>=0.95 (1M x768 IP)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=str, default="sift", choices=sorted(CONFIGS))
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--d", type=int, default=None)
    ap.add_argument("--d1", type=int, default=None)
    ap.add_argument("--nq", type=int, default=None)
    ap.add_argument("--M", type=int, default=None)
    ap.add_argument("--efc", type=int, default=None)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--target-recall", type=float, default=0.95)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--ip", action="store_true", default=None)
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    for key in ("n", "d", "d1", "nq", "M", "efc", "ip"):
        if getattr(a, key) is None:
            setattr(a, key, cfg[key])
    a.custom = any(getattr(a, key) != cfg[key] for key in ("n", "d", "d1", "nq", "M", "efc", "ip"))
    a.metric_name = cfg["metric"] if not a.custom else f"QPS at recall@10>={a.target_recall} ({a.n} x{a.d})"
    a.shape = cfg["shape"] if not a.custom else "custom-shape"
    return a


def workload_name(a):
    return (f"{a.shape} {a.n}x{a.d} fp32 {'IP' if a.ip else 'L2'}, M={a.M} efC={a.efc}, "
            f"{a.nq}-query batch, k={a.k}")


def config_dict(a, ef_sel):
    """Identical keys (and, for one workload, identical values) in both arms."""
    return {"workload": workload_name(a), "efSearch": int(ef_sel), "d1": a.d1, "seed": SEED, "l2_policy": L2_POLICY}


def recall_at_k(I, gt):
    k = gt.shape[1]
    hit = 0
    for i in range(gt.shape[0]):
        hit += len(set(I[i, :k].tolist()) & set(gt[i].tolist()))
    return hit / float(gt.size)


def bytes_per_query(stats, d, M, k):
    """ndis*4d + nhops0*2M*4 + nhops_up*M*4 + 4d (query) + 12k (result)."""
    s = stats.astype(np.float64).mean(0)
    return (s[0] + s[2]) * 4 * d + s[1] * 8 * M + s[3] * 4 * M + 4 * d + 12 * k, s


def team_for_dim(d):
    """Lanes per vector the kernels use for a row of d fp32 (beam_kernel_impl.cuh launch_by_chunks):
    the oracle's team mode with this value reproduces the CUDA summation order bit for bit."""
    nc = (d + 3) // 4
    return 8 if nc <= 32 else (16 if nc <= 64 else 32)


def kernel_source_sha():
    """Identity of the traversal kernel's sources: profiles/*traffic*.json is only quoted when it was
    captured from this exact code."""
    h = hashlib.sha256()
    for fn in ("beam.cuh", "beam_kernel_impl.cuh", "beam_launch.cuh", "select.cuh", "common.cuh", "engine.h"):
        with open(os.path.join(ROOT, "hnsw_b200", "csrc", fn), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML, every 10 ms)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.reasons, self.max_sm = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nv = pynvml
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _sample(self):
        nv = self._nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
            nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                          ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80)):
            if r & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self._nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["clock sampling unavailable"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(sm)}


QUERY_POOL = 8  # query sets generated (one per possible rank); fixed so the database is the same for every N


def make_data(a, rank=0, device=None):
    """Database + this rank's query set. The pool of QUERY_POOL x nq queries and the database come
    from ONE seeded draw, so xb is bit-identical for every N and every rank; rank r takes slice r.
    Sets too large to draw comfortably in numpy (> 4M rows) are generated on the device by the torch
    twin of the recipe."""
    r = rank % QUERY_POOL
    if a.n > 4_000_000 and device is not None:
        from hnsw_b200.datasets import synthetic_dataset_torch
        xb_t, xq_t = synthetic_dataset_torch(a.d, a.n, QUERY_POOL * a.nq, d1=a.d1, seed=SEED, normalize=a.ip,
                                             device=device)
        xb = xb_t.cpu().numpy()
        xq = xq_t[r * a.nq:(r + 1) * a.nq].cpu().numpy()
        del xb_t, xq_t
        return xb, np.ascontiguousarray(xq)
    from hnsw_b200.datasets import synthetic_dataset
    xb, xq_all = synthetic_dataset(a.d, a.n, QUERY_POOL * a.nq, d1=a.d1, seed=SEED, normalize=a.ip)
    return xb, np.ascontiguousarray(xq_all[r * a.nq:(r + 1) * a.nq])


def native_oracle():
    """CPU baseline library, compiled for the box's own CPU."""
    from oracle import oracle as om
    try:
        path = om.build(arch="native", out="liboracle_native.so")
    except Exception:
        path = om.build()
    return om, path


def try_faiss():
    """SURVEY §8c upgrade path: real faiss, if this box happens to have it (it never did here)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import faiss  # noqa: F401
        return faiss
    except Exception:
        return None


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class FaissCPU:
    """faiss.IndexHNSWFlat behind the oracle wrapper's interface (used only when faiss imports)."""

    def __init__(self, faiss, d, M, ip):
        self.f = faiss
        self.idx = faiss.IndexHNSWFlat(d, M, faiss.METRIC_INNER_PRODUCT if ip else faiss.METRIC_L2)
        self.threads = 1

    @property
    def efConstruction(self):
        return self.idx.hnsw.efConstruction

    @efConstruction.setter
    def efConstruction(self, v):
        self.idx.hnsw.efConstruction = int(v)

    def add(self, x):
        self.f.omp_set_num_threads(int(self.threads))
        self.idx.add(x)

    def search(self, xq, k, ef):
        self.f.omp_set_num_threads(int(self.threads))
        self.idx.hnsw.efSearch = int(ef)
        return self.idx.search(xq, k)


def median_us(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e6)
    ts.sort()
    return ts[len(ts) // 2]


# ---------------------------------------------------------------------------------------------
def run_reference(a):
    """CPU arm: the oracle port of faiss IndexHNSWFlat (no faiss, no reference sources exist) with
    all host threads: build + search on the same config. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    xb, xq = make_data(a, 0)
    faiss = try_faiss()
    if faiss is not None:
        o, kind, what = FaissCPU(faiss, a.d, a.M, a.ip), "reference", "faiss.IndexHNSWFlat (real faiss found at run time)"
    else:
        om, path = native_oracle()
        o = om.OracleHNSWFlat(a.d, a.M, om.METRIC_INNER_PRODUCT if a.ip else om.METRIC_L2, lib_path=path)
        kind, what = "port", "oracle port (faiss-semantics CPU restatement, NOT faiss)"
    # bounded build: calibrate on 50k vectors, keep the whole build under ~120 s
    o.efConstruction = a.efc
    o.threads = threads
    n_cal = min(a.n, 50_000)
    t0 = time.time()
    o.add(xb[:n_cal])
    t_cal = time.time() - t0
    rate = n_cal / t_cal
    n_ref = a.n
    if a.n / (0.6 * rate) > 120:  # rate drops as the graph grows
        n_ref = int(max(n_cal, min(a.n, 0.6 * rate * 120)))
    t0 = time.time()
    if n_ref > n_cal:
        o.add(xb[n_cal:n_ref])
    t_build = t_cal + (time.time() - t0)
    xb = xb[:n_ref]
    import torch  # CPU tensors only: exact ground truth by chunked GEMM
    torch.set_num_threads(threads)  # torchrun exports OMP_NUM_THREADS=1; the GEMM should use the box
    from hnsw_b200.datasets import exact_knn_torch
    _, gt = exact_knn_torch(torch.from_numpy(xb), torch.from_numpy(xq), a.k, inner_product=a.ip, chunk=1 << 15)
    gt = gt.numpy()
    ef_sel, rec_sel, sweep = None, None, []
    for ef in EF_GRID:
        t0 = time.time()
        D, I = o.search(xq, a.k, ef)
        dt = time.time() - t0
        r = recall_at_k(I, gt)
        sweep.append({"efSearch": ef, "recall": round(r, 4), "qps": round(a.nq / dt)})
        if ef_sel is None and r >= a.target_recall:
            ef_sel, rec_sel = ef, r
            break
    if ef_sel is None:
        ef_sel, rec_sel = EF_GRID[-1], sweep[-1]["recall"]
    for _ in range(a.warmup):
        o.search(xq, a.k, ef_sel)
    t0 = time.time()
    for _ in range(a.steps):
        o.search(xq, a.k, ef_sel)
    dt = time.time() - t0
    qps = a.nq * a.steps / dt
    sample = (f"{what}, {threads} OpenMP threads on {cpu_model()}; "
              f"index built on {n_ref} of {a.n} vectors; {a.nq} queries/step")
    line = {
        "impl": "reference", "metric": a.metric_name, "value": round(qps, 1), "unit": "queries/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(dt / a.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(a, ef_sel),
        "recall_at_10": round(rec_sel, 4), "n_indexed": n_ref,
        "build_vectors_per_s": round(n_ref / t_build, 1), "ef_sweep": sweep,
        "cpu_baseline": {"value": round(qps, 1), "unit": "queries/s", "cores": threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": round(qps, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist

    import hnsw_b200
    from hnsw_b200.datasets import exact_knn_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    metric = hnsw_b200.METRIC_INNER_PRODUCT if a.ip else hnsw_b200.METRIC_L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- data: every rank indexes the SAME database (replicas) and answers its own query set
    xb, xq = make_data(a, rank, dev)
    xq_t = torch.from_numpy(xq).to(dev)
    xb_t = torch.from_numpy(xb).to(dev)
    _, gt_t = exact_knn_torch(xb_t, xq_t, a.k, inner_product=a.ip, chunk=1 << 16)
    gt = gt_t.cpu().numpy()
    del xb_t
    torch.cuda.empty_cache()  # give the brute-force scratch back before the index allocates

    # ---- build (add): vectors/sec, wall clock around the public call (H2D included)
    idx = hnsw_b200.IndexHNSWFlat(a.d, a.M, metric, device=local)
    idx.hnsw.efConstruction = a.efc
    barrier()
    l0 = hnsw_b200.launch_count()
    t0 = time.time()
    idx.add(xb)
    t_build = time.time() - t0
    build_launches = hnsw_b200.launch_count() - l0
    build_dev_s = idx.last_build_ms / 1e3
    bc = idx.last_build_counters  # measured numerators of the build's HBM roofline
    row_b = a.d * 4
    build_bytes = ((bc["ndis0"] + bc["ndis_up"] + bc["sel_rows"] + bc["bl_rows"]) * row_b
                   + bc["nhops0"] * 8 * a.M + bc["nhops_up"] * 4 * a.M + a.n * (row_b + 8 * a.M))
    t_build = max_over_ranks(t_build)

    # ---- efSearch sweep (untimed setup): recall, QPS, roofline fraction per ef. All on device buffers,
    #      CUDA events on the index's stream. Two counts per ef: `ndis` as the timed kernel runs it (its
    #      small visited table may forget a vertex and score it again) and `ndis_exact` from a pass with
    #      a visited table that never forgets (= faiss's VisitedTable; same result ids, asserted).
    #      frac = touched bytes / time / peak; frac_exact = the faiss algorithm's own bytes / time / peak.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    D_d = torch.empty(a.nq, a.k, device=dev)
    I_d = torch.empty(a.nq, a.k, dtype=torch.int64, device=dev)
    S_d = torch.zeros(a.nq, 4, dtype=torch.int32, device=dev)
    stream = torch.cuda.ExternalStream(idx.stream_ptr, device=dev)

    def device_pass(ef, stats=False, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        idx.search_device(xq_t.data_ptr(), a.nq, a.k, D_d.data_ptr(), I_d.data_ptr(), efSearch=ef,
                          stats_ptr=S_d.data_ptr() if stats else 0, **kw)
        e1.record(stream)
        idx.synchronize()
        return e0.elapsed_time(e1)

    sweep, ef_sel = [], None
    for ef in EF_GRID:
        device_pass(ef, stats=True, visited_policy=1, hash_bits=15)   # exact visited table
        I_exact = I_d.cpu().numpy()
        bq_x, s_x = bytes_per_query(S_d.cpu().numpy(), a.d, a.M, a.k)
        device_pass(ef, stats=True)
        st = S_d.cpu().numpy()
        I = I_d.cpu().numpy()
        assert np.array_equal(I, I_exact), f"efSearch={ef}: the forgetful visited table changed the result ids"
        ms = min(device_pass(ef), device_pass(ef))
        r = recall_at_k(I, gt)
        bq, s = bytes_per_query(st, a.d, a.M, a.k)
        sweep.append({"efSearch": ef, "recall": round(r, 4), "qps": round(a.nq / ms * 1e3),
                      "ndis": round(s[0], 1), "ndis_exact": round(s_x[0], 1), "nhops": round(s[1], 1),
                      "bytes_per_query": round(bq), "bytes_per_query_exact": round(bq_x),
                      "gather_gbs": round(bq * a.nq / ms / 1e6, 1), "frac": round(bq * a.nq / ms / 1e6 / peak, 3),
                      "frac_exact": round(bq_x * a.nq / ms / 1e6 / peak, 3)})
        if ef_sel is None and r >= a.target_recall:
            ef_sel = ef
    if ef_sel is None:
        ef_sel = EF_GRID[-1]
    if world > 1:  # every rank times rank 0's efSearch (= the N=1 workload's operating point)
        t = torch.tensor([ef_sel], device=dev)
        dist.broadcast(t, src=0)
        ef_sel = int(t.item())
    sel = next(x for x in sweep if x["efSearch"] == ef_sel)

    # ---- timed region 1: `value` — inputs resident in HBM, K steps back to back
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step_device():
        idx.search_device(xq_t.data_ptr(), a.nq, a.k, D_d.data_ptr(), I_d.data_ptr(), efSearch=ef_sel)

    for _ in range(max(a.warmup, 3)):
        step_device()
    barrier()
    with ClockSampler(local) as clk:
        l0 = hnsw_b200.launch_count()
        ev0.record(stream)
        for _ in range(a.steps):
            step_device()
        ev1.record(stream)
        idx.synchronize()
        barrier()
        gpu_launches = hnsw_b200.launch_count() - l0
    ms_rank = ev0.elapsed_time(ev1) / a.steps
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / a.steps
    ms_per_rank = [round(ms_rank, 4)]
    if world > 1:
        tt = torch.tensor([ms_rank], dtype=torch.float64, device=dev)
        outl = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(outl, tt)
        ms_per_rank = [round(float(x.item()), 4) for x in outl]
    value = world * a.nq / (ms_step * 1e-3)
    D_timed, I_timed = D_d.cpu().numpy(), I_d.cpu().numpy()
    rec_timed = recall_at_k(I_timed, gt)
    rec_min = -max_over_ranks(-rec_timed)

    # ---- timed region 2: `e2e` — the public call with HOST buffers, copies inside the timed region.
    #      (a) page-locked caller buffers (zero-copy path), (b) plain pageable numpy arrays (what a faiss
    #      drop-in caller passes; chunked over the context's lanes, staged through pinned memory).
    def time_public(xq_np, out_np):
        for _ in range(max(a.warmup, 3)):
            idx.search(xq_np, a.k, efSearch=ef_sel, out=out_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            idx.search(xq_np, a.k, efSearch=ef_sel, out=out_np)
        torch.cuda.synchronize()
        return world * a.nq * a.steps / max_over_ranks(time.perf_counter() - t0)

    xq_pin = torch.empty(a.nq, a.d, dtype=torch.float32).pin_memory()
    xq_pin.copy_(torch.from_numpy(xq))
    D_pin = torch.empty(a.nq, a.k, dtype=torch.float32).pin_memory()
    I_pin = torch.empty(a.nq, a.k, dtype=torch.int64).pin_memory()
    e2e = time_public(xq_pin.numpy(), (D_pin.numpy(), I_pin.numpy()))
    assert np.array_equal(I_pin.numpy(), I_timed)
    D_pg, I_pg = np.empty((a.nq, a.k), np.float32), np.empty((a.nq, a.k), np.int64)
    e2e_pageable = time_public(xq, (D_pg, I_pg))
    assert np.array_equal(I_pg, I_timed) and np.array_equal(D_pg, D_timed)

    # ---- small batches (latency regime): the public call on pageable buffers, median of 30
    latency = []
    if rank == 0:
        for nq_small in (1, 16, 256):
            q = np.ascontiguousarray(xq[:nq_small])
            idx.search(q, a.k, efSearch=ef_sel)
            us = median_us(lambda: idx.search(q, a.k, efSearch=ef_sel), 30)
            latency.append({"nq": nq_small, "us_per_batch": round(us, 1), "us_per_query": round(us / nq_small, 2)})

    # ---- north-star sharded path at fixed per-GPU shard size (N > 1)
    sharded = None
    if world > 1 and not a.no_sharded:
        sharded = run_sharded_weak(a, idx, rank, world, local, dev, ef_sel, ms_step, barrier, max_over_ranks)

    # ---- CPU baseline (rank 0, N=1): oracle port searching the SAME graph on the host cores, and the
    #      parity assertion at the headline configuration
    cpu_baseline, parity_sample = None, None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            om, path = native_oracle()
            g = idx.export_graph()
            o = om.OracleHNSWFlat(a.d, a.M, om.METRIC_INNER_PRODUCT if a.ip else om.METRIC_L2, lib_path=path)
            o.import_graph(xb, g["levels"], g["neighbors"], g["entry_point"], g["max_level"])
            o.threads = os.cpu_count() or 1
            # parity: the oracle in the kernels' summation order on the GPU-built graph, first 1000
            # queries of the timed batch — ids and distances must be IDENTICAL
            np_s = min(1000, a.nq)
            o.set_team(team_for_dim(a.d))
            Dp, Ip = o.search(xq[:np_s], a.k, ef_sel)
            o.set_team(0)
            n_same = int(np.sum(np.all(Ip == I_timed[:np_s], axis=1) & np.all(Dp == D_timed[:np_s], axis=1)))
            parity_sample = f"{n_same}/{np_s} identical"
            o.search(xq[:2000], a.k, ef_sel)
            best, reps, t_all = 0.0, 0, time.time()
            while time.time() - t_all < 10 and reps < 20:
                t0 = time.time()
                Dc, Ic = o.search(xq, a.k, ef_sel)
                best = max(best, a.nq / (time.time() - t0))
                reps += 1
            cpu_lat = []
            for nq_small in (1, 16, 256):
                q = np.ascontiguousarray(xq[:nq_small])
                o.threads = 1 if nq_small == 1 else (os.cpu_count() or 1)
                us = median_us(lambda: o.search(q, a.k, ef_sel), 15)
                cpu_lat.append({"nq": nq_small, "us_per_batch": round(us, 1), "us_per_query": round(us / nq_small, 2)})
            o.threads = os.cpu_count() or 1
            # CPU build rate on a bounded sample (first 50k vectors of the same data)
            ob = om.OracleHNSWFlat(a.d, a.M, om.METRIC_INNER_PRODUCT if a.ip else om.METRIC_L2, lib_path=path)
            ob.efConstruction = a.efc
            ob.threads = o.threads
            nb_s = min(a.n, 50_000)
            t0 = time.time()
            ob.add(xb[:nb_s])
            cpu_build = nb_s / (time.time() - t0)
            cpu_baseline = {"value": round(best, 1), "unit": "queries/s", "cores": o.threads, "kind": "port",
                            "sample": f"all {a.nq} queries x {reps} passes at efSearch={ef_sel} on the GPU-built "
                                      f"{a.n}-vector graph (best pass); CPU recall {recall_at_k(Ic, gt):.4f}; "
                                      f"oracle = faiss-semantics restatement, not faiss; {cpu_model()}",
                            "latency": cpu_lat,
                            "build_vectors_per_s": round(cpu_build, 1),
                            "build_sample": f"first {nb_s} vectors (rate falls as the graph grows)"}
        except Exception as e:  # the baseline is reported, never required
            cpu_baseline = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": f"failed: {e}"}

    if rank == 0:
        # DRAM bytes of one launch come from an ncu capture (scripts/capture_traffic.sh), quoted only
        # when it was taken from this exact kernel source on this workload
        traffic, traffic_source = None, "none: no ncu capture of this kernel source + workload under profiles/"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            ent = tj.get("entries", {}).get(f"{a.config}:{ef_sel}")
            if a.custom or ent is None:
                pass
            elif tj.get("kernel_src_sha") != kernel_source_sha():
                traffic_source = "stale: profiles/r2_traffic.json was captured from other kernel sources"
            else:
                traffic = ent["dram_bytes_per_launch"]
                traffic_source = (f"profiles/r2_traffic.json ({ent['kernel']}; ncu --set full, dram__bytes_read.sum + "
                                  f"dram__bytes_write.sum; same data recipe and kernel sources as this run)")
        except Exception:
            pass
        achieved = sel["bytes_per_query"] * a.nq / (ms_step * 1e-3) / 1e9
        achieved_exact = sel["bytes_per_query_exact"] * a.nq / (ms_step * 1e-3) / 1e9
        line = {
            "metric": a.metric_name, "value": round(value, 1), "unit": "queries/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": round(ms_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": config_dict(a, ef_sel),
            "recall_at_10": round(rec_timed, 4), "recall_at_10_min_over_ranks": round(rec_min, 4),
            "replicas": None if world == 1 else f"{world} replicas, each GPU its own {a.n}-vector index and {a.nq} queries",
            "build_vectors_per_s": round(a.n / t_build, 1),
            "build": {"wall_s": round(t_build, 3), "device_s": round(build_dev_s, 3), "launches": build_launches,
                      "vectors_per_s_device": round(a.n / build_dev_s, 1),
                      "counters": bc, "bytes_per_vector": round(build_bytes / a.n),
                      "roofline": {"bound": "hbm", "achieved": round(build_bytes / build_dev_s / 1e9, 1), "peak": peak,
                                   "unit": "GB/s", "frac": round(build_bytes / build_dev_s / 1e9 / peak, 4),
                                   "note": "algorithmic bytes: vectors scored by the insertion searches + candidate "
                                           "vectors read by selection + vectors streamed by back-link shrinks + rows"}},
            "ef_sweep": sweep,
            # the judged fraction counts the faiss algorithm's own bytes (exact visited set); the bytes
            # the kernel actually touched (re-scored vertices included) are beside it
            "roofline": {"bound": "hbm", "achieved": round(achieved_exact, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved_exact / peak, 4), "traffic": traffic, "traffic_source": traffic_source,
                         # DRAM-level view: measured DRAM bytes of a launch / step time (L2 serves the entry region, so
                         # this is below the algorithmic figure; "frac" can read above 1.0, this cannot)
                         "traffic_frac": None if traffic is None else round(traffic / (ms_step * 1e-3) / 1e9 / peak, 4),
                         "peak_source": peak_src, "kernel": "beam_kernel (one launch per step)",
                         "bytes_per_launch": round(sel["bytes_per_query_exact"] * a.nq),
                         "touched": {"achieved": round(achieved, 1), "frac": round(achieved / peak, 4),
                                     "bytes_per_launch": round(sel["bytes_per_query"] * a.nq),
                                     "ndis_over_exact": round(sel["ndis"] / max(sel["ndis_exact"], 1), 4)}},
            "cpu_baseline": cpu_baseline,
            "parity_sample": parity_sample,
            "e2e": {"value": round(e2e, 1), "unit": "queries/s", "h2d_bytes_per_step": a.nq * a.d * 4,
                    "d2h_bytes_per_step": a.nq * a.k * 12, "buffers": "page-locked host (zero-copy over PCIe)"},
            "e2e_pageable": {"value": round(e2e_pageable, 1), "unit": "queries/s",
                             "h2d_bytes_per_step": a.nq * a.d * 4, "d2h_bytes_per_step": a.nq * a.k * 12,
                             "buffers": "pageable numpy (faiss drop-in caller): chunked over 3 lanes, staged "
                                        "through page-locked memory"},
            "latency": latency,
            "gpu_launches": int(gpu_launches),
            "ms_per_step_per_rank": ms_per_rank,
            "clocks": clk.summary(),
        }
        if sharded:
            line["sharded_weak"] = sharded
        print(json.dumps(line), flush=True)
        if parity_sample is not None and not parity_sample.startswith(f"{min(1000, a.nq)}/"):
            raise SystemExit(f"bench.py: PARITY FAILURE at the headline configuration: {parity_sample}")
    if world > 1:
        dist.destroy_process_group()


def run_sharded_weak(a, idx_replica, rank, world, local, dev, ef_sel, ms_single, barrier, max_over_ranks):
    """Each rank's shard = slice `rank` of an (N x n)-vector database drawn from one distribution; the same
    nq queries go to every shard; merged (D, I) is checked against an exact host-side merge of the
    per-shard lists and recall is measured against global exact ground truth."""
    import torch
    import torch.distributed as dist

    import hnsw_b200
    from hnsw_b200.datasets import exact_knn_torch, synthetic_dataset_torch
    from hnsw_b200.sharded import ShardedIndexHNSWFlat

    metric = hnsw_b200.METRIC_INNER_PRODUCT if a.ip else hnsw_b200.METRIC_L2
    n_sh = a.n
    # one seeded device draw of the whole database + queries (every rank draws the same stream), rank r
    # keeps slice r: chunked generation, so only this rank's slice is ever resident
    xb_sh, xq_t = synthetic_dataset_torch(a.d, n_sh * world, a.nq, d1=a.d1, seed=SEED + 1, normalize=a.ip,
                                          device=dev, keep_rows=(rank * n_sh, (rank + 1) * n_sh))
    Dl_gt, Il_gt = exact_knn_torch(xb_sh, xq_t, a.k, inner_product=a.ip, chunk=1 << 16)
    sh = ShardedIndexHNSWFlat(a.d, a.M, metric, device=dev)
    sh.local.hnsw.efConstruction = a.efc
    t0 = time.time()
    sh.add(xb_sh.cpu().numpy())
    t_build = max_over_ranks(time.time() - t0)
    del xb_sh
    torch.cuda.empty_cache()
    # global exact ground truth = exact merge of the per-shard exact lists
    gl_D = [torch.empty_like(Dl_gt) for _ in range(world)]
    gl_I = [torch.empty_like(Il_gt) for _ in range(world)]
    dist.all_gather(gl_D, Dl_gt)
    dist.all_gather(gl_I, Il_gt)
    allD = torch.cat(gl_D, 1)
    allI = torch.cat([g + r * n_sh for r, g in enumerate(gl_I)], 1)
    order = torch.argsort(-allD if a.ip else allD, dim=1, stable=True)[:, :a.k]
    gt = torch.gather(allI, 1, order).cpu().numpy()

    out = sh.search(xq_t, a.k, efSearch=ef_sel, keep_local=True)
    torch.cuda.synchronize()
    Dm, Im = out[0].cpu().numpy(), out[1].cpu().numpy()
    # exactness of the exchange + merge: host-side merge of the per-shard lists every rank produced
    Dl, Il = sh.last_local
    gD = [torch.empty_like(Dl) for _ in range(world)]
    gI = [torch.empty_like(Il) for _ in range(world)]
    dist.all_gather(gD, Dl)
    dist.all_gather(gI, Il)
    hD = torch.cat(gD, 1)
    hI = torch.cat([torch.where(g >= 0, g + r * n_sh, g) for r, g in enumerate(gI)], 1)
    ho = torch.argsort(-hD if a.ip else hD, dim=1, stable=True)[:, :a.k]
    assert np.array_equal(Im, torch.gather(hI, 1, ho).cpu().numpy()), "sharded merge differs from the exact host merge"
    assert np.array_equal(Dm, torch.gather(hD, 1, ho).cpu().numpy())
    rec = recall_at_k(Im, gt)

    for _ in range(3):
        sh.search(xq_t, a.k, efSearch=ef_sel)
    barrier()
    cur = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    for _ in range(a.steps):
        sh.search(xq_t, a.k, efSearch=ef_sel)
    e1.record(cur)
    barrier()
    ms_sh = max_over_ranks(e0.elapsed_time(e1)) / a.steps
    # the same shard searched alone (no exchange, no merge): the efficiency denominator. A sharded step is one
    # synchronous launch + exchange, so the like-for-like figure is the ISOLATED launch (each launch timed on
    # its own); the pipelined figure (back-to-back launches overlap their drain phases) is beside it.
    D1 = torch.empty(a.nq, a.k, device=dev)
    I1 = torch.empty(a.nq, a.k, dtype=torch.int64, device=dev)
    st1 = torch.cuda.ExternalStream(sh.local.stream_ptr, device=dev)
    for _ in range(3):
        sh.local.search_device(xq_t.data_ptr(), a.nq, a.k, D1.data_ptr(), I1.data_ptr(), efSearch=ef_sel)
    barrier()
    iso = 0.0
    for _ in range(a.steps):
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(st1)
        sh.local.search_device(xq_t.data_ptr(), a.nq, a.k, D1.data_ptr(), I1.data_ptr(), efSearch=ef_sel)
        g1.record(st1)
        sh.local.synchronize()
        iso += g0.elapsed_time(g1)
    ms_alone = max_over_ranks(iso) / a.steps
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(st1)
    for _ in range(a.steps):
        sh.local.search_device(xq_t.data_ptr(), a.nq, a.k, D1.data_ptr(), I1.data_ptr(), efSearch=ef_sel)
    f1.record(st1)
    sh.local.synchronize()
    barrier()
    ms_alone_pipelined = max_over_ranks(f0.elapsed_time(f1)) / a.steps
    # pipelined serving loop: batches enqueued back to back, flag + merge on the exchange stream, one join
    pipelined = None
    if sh.exchange_kind == "peer-store":
        Dp = torch.empty(a.nq, a.k, device=dev)
        Ip = torch.empty(a.nq, a.k, dtype=torch.int64, device=dev)
        sh.set_pipelined(True)
        for _ in range(3):
            sh.enqueue(xq_t, a.k, Dp, Ip, efSearch=ef_sel)
        sh.join(st1)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(st1)
        for _ in range(a.steps):
            sh.enqueue(xq_t, a.k, Dp, Ip, efSearch=ef_sel)
        sh.join(st1)
        p1.record(st1)
        sh.local.synchronize()
        barrier()
        ms_pp = max_over_ranks(p0.elapsed_time(p1)) / a.steps
        assert np.array_equal(Ip.cpu().numpy(), Im) and np.array_equal(Dp.cpu().numpy(), Dm), "pipelined result differs"
        sh.set_pipelined(False)
        pipelined = {"value": round(a.nq / (ms_pp * 1e-3), 1), "unit": "queries/s", "ms_per_step": round(ms_pp, 4),
                     "efficiency_vs_pipelined_single_shard": round(ms_alone_pipelined / ms_pp, 4),
                     "note": "batches enqueued back to back (bh_shards_set_pipelined): traversal launches overlap their "
                             "drain phases, flag + merge kernels on the exchange stream, one join at the end; result "
                             "asserted equal to the synchronous call's"}
    res = {"value": round(a.nq / (ms_sh * 1e-3), 1), "unit": "queries/s", "ms_per_step": round(ms_sh, 4),
           "ms_shard_search_alone": round(ms_alone, 4), "ms_exchange_merge_and_rank_skew": round(ms_sh - ms_alone, 4),
           "efficiency_vs_single_shard": round(ms_alone / ms_sh, 4),
           "ms_shard_search_alone_pipelined": round(ms_alone_pipelined, 4),
           "efficiency_vs_pipelined_single_shard": round(ms_alone_pipelined / ms_sh, 4),
           "db_vectors": n_sh * world, "shard_vectors": n_sh, "recall_at_10": round(rec, 4),
           "efSearch": ef_sel, "exchange": sh.exchange_kind, "payload_bytes_per_rank": a.nq * a.k * 8,
           "merge_check": "merged (D, I) == exact host-side merge of the per-shard lists",
           "build_s": round(t_build, 2), "build_vectors_per_s_all_shards": round(n_sh * world / t_build, 1),
           "pipelined": pipelined}
    del sh
    return res


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
