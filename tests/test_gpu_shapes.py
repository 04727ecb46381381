"""GPU parity at the shapes BASELINE.json names beyond configs[0]/[1] (scaled to sizes the CPU oracle
builds in about a minute): GIST-shape 960-d L2, embedding-shape 768-d inner product, Deep-shape 96-d
with M=32 / efConstruction=200. For each:
  * a graph built by the CPU oracle (all host threads) is imported and searched on the GPU — ids,
    distances and per-query ndis / nhops must be IDENTICAL to the oracle's (kernel summation order);
  * the default (forgetful) visited table returns the same ids / distances;
  * a graph built by the GPU (batched rounds) must reach the oracle-built graph's recall@10 within
    0.5 points at equal M / efConstruction / efSearch, on 5000 queries against exact ground truth."""
import os

import numpy as np
import pytest

from hnsw_b200.datasets import synthetic_dataset

pytestmark = pytest.mark.gpu

SHAPES = [
    # name,            n,       d,  M, efc, metric, team, ef
    ("gist-shape",     100_000, 960, 32, 200, 1, 32, 128),
    ("ip768-shape",    100_000, 768, 32, 200, 0, 32, 128),
    ("deep-shape",     200_000, 96,  32, 200, 1, 8, 64),
]


@pytest.mark.parametrize("name,n,d,M,efc,metric,team,ef", SHAPES, ids=[s[0] for s in SHAPES])
def test_baseline_shapes_bit_exact_search_and_recall_parity(oracle_mod, name, n, d, M, efc, metric, team, ef):
    import torch
    import hnsw_b200
    from hnsw_b200.datasets import exact_knn_torch
    nq, k = 5000, 10
    xb, xq = synthetic_dataset(d, n, nq, d1=12, normalize=(metric == 0))
    o = oracle_mod.OracleHNSWFlat(d, M, metric)
    o.efConstruction = efc
    o.threads = os.cpu_count() or 8
    o.add(xb)
    g = o.export_graph()
    o.set_team(team)
    Do, Io, So = o.search(xq[:1000], k, ef, stats=True)

    idx = hnsw_b200.IndexHNSWFlat(d, M, metric)
    idx.import_graph(xb, g["levels"], g["neighbors"], g["entry_point"], g["max_level"])
    D, I, S = idx.search(xq[:1000], k, efSearch=ef, stats=True, hash_bits=14)
    assert np.array_equal(I, Io) and np.array_equal(D, Do), name
    assert np.array_equal(S, So), name                       # same path: ndis / nhops equal the oracle's
    D2, I2, S2 = idx.search(xq[:1000], k, efSearch=ef, stats=True)          # default visited table
    assert np.array_equal(I2, Io) and np.array_equal(D2, Do)
    assert np.array_equal(S2[:, 1:], So[:, 1:]) and S2[:, 0].sum() <= 1.10 * So[:, 0].sum()

    _, gt = exact_knn_torch(torch.from_numpy(xb).cuda(), torch.from_numpy(xq).cuda(), k, inner_product=(metric == 0))
    gt = gt.cpu().numpy()
    r_cpu_graph = oracle_mod.recall_at_k(idx.search(xq, k, efSearch=ef)[1], gt)
    b = hnsw_b200.IndexHNSWFlat(d, M, metric)
    b.hnsw.efConstruction = efc
    b.add(xb)
    r_gpu_graph = oracle_mod.recall_at_k(b.search(xq, k, efSearch=ef)[1], gt)
    assert r_cpu_graph > 0.8, (name, r_cpu_graph)
    assert r_gpu_graph >= r_cpu_graph - 0.005, (name, r_gpu_graph, r_cpu_graph)
