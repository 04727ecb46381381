"""Synthetic datasets of the shapes BASELINE.json names.

`synthetic_dataset` restates the recipe of faiss `contrib.datasets.SyntheticDataset`
(SURVEY.md §8d): a d1-dimensional Gaussian pushed through a random linear map, a
per-dimension scale and a sine, so the points lie on a low-dimensional non-linear
manifold in R^d. Upstream uses d1=10, seed=1338. The numpy version is bit-reproducible
and is used by tests; the torch version generates the 1M…100M-row shapes directly on
the device for the benchmark (same recipe, torch's generator).
"""
from __future__ import annotations

import numpy as np


def synthetic_dataset(d: int, nb: int, nq: int, d1: int = 10, seed: int = 1338,
                      normalize: bool = False):
    """Returns (xb [nb,d], xq [nq,d]) float32. Rows are database first, then queries."""
    n = nb + nq
    rs = np.random.RandomState(seed)
    x = rs.normal(size=(n, d1))
    x = np.dot(x, rs.rand(d1, d))
    x = x * (rs.rand(d) * 4 + 0.1)
    x = np.sin(x)
    x = x.astype("float32")
    if normalize:
        x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-20)
    return np.ascontiguousarray(x[:nb]), np.ascontiguousarray(x[nb:])


def synthetic_dataset_torch(d: int, nb: int, nq: int, d1: int = 32, seed: int = 1338,
                            normalize: bool = False, device="cuda", chunk: int = 1 << 20,
                            keep_rows: tuple[int, int] | None = None):
    """Same recipe on a torch device, generated in row chunks (fits 100M x 96).
    keep_rows=(lo, hi): every row is still drawn (so the stream is the same on every rank) but only
    database rows [lo, hi) and the queries are kept — a rank's shard of an N-shard database."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    proj = torch.rand(d1, d, generator=g, device=device, dtype=torch.float32)
    scale = torch.rand(d, generator=g, device=device, dtype=torch.float32) * 4 + 0.1
    n = nb + nq
    lo, hi = keep_rows if keep_rows is not None else (0, nb)
    xb = torch.empty(hi - lo, d, device=device, dtype=torch.float32)
    xq = torch.empty(nq, d, device=device, dtype=torch.float32)
    for i0 in range(0, n, chunk):
        i1 = min(n, i0 + chunk)
        z = torch.randn(i1 - i0, d1, generator=g, device=device, dtype=torch.float32)
        x = torch.sin((z @ proj) * scale)
        if normalize:
            x = x / x.norm(dim=1, keepdim=True).clamp_min(1e-20)
        a, b = max(i0, lo), min(i1, hi)          # database rows of this chunk that are kept
        if a < b:
            xb[a - lo:b - lo] = x[a - i0:b - i0]
        a, b = max(i0, nb), i1                     # query rows of this chunk
        if a < b:
            xq[a - nb:b - nb] = x[a - i0:b - i0]
    return xb, xq


def exact_knn_torch(xb, xq, k: int, inner_product: bool = False, chunk: int = 1 << 18):
    """Exact top-k ground truth by chunked fp32 GEMM on the device (NOT on the timed path)."""
    import torch

    nq = xq.shape[0]
    best_d = torch.full((nq, k), float("inf"), device=xq.device)
    best_i = torch.full((nq, k), -1, device=xq.device, dtype=torch.int64)
    qn = (xq * xq).sum(1, keepdim=True)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for i0 in range(0, xb.shape[0], chunk):
            xc = xb[i0:i0 + chunk]
            if inner_product:
                dm = -(xq @ xc.T)
            else:
                dm = qn - 2 * (xq @ xc.T) + (xc * xc).sum(1)[None, :]
            dcat = torch.cat([best_d, dm], 1)
            icat = torch.cat([best_i, torch.arange(i0, i0 + xc.shape[0], device=xq.device)
                              .expand(nq, -1)], 1)
            best_d, sel = torch.topk(dcat, k, dim=1, largest=False)
            best_i = torch.gather(icat, 1, sel)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return (-best_d if inner_product else best_d), best_i
