import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) on a box without a CUDA device."""
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as om
    om.build()
    return om


@pytest.fixture(scope="session")
def small_l2(oracle_mod):
    """d=32, 4000 base, 100 queries, M=16, efC=40 — oracle graph built sequentially, team-8 order."""
    from hnsw_b200.datasets import synthetic_dataset
    xb, xq = synthetic_dataset(32, 4000, 100)
    o = oracle_mod.OracleHNSWFlat(32, 16)
    o.set_team(8)
    o.add(xb)
    return dict(xb=xb, xq=xq, oracle=o, graph=o.export_graph())


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def assert_graph_invariants(g, M, ntotal):
    """rows -1 terminated & packed, no self loops, no duplicates, ids in range."""
    levels, offsets, nb = g["levels"], g["offsets"].astype(np.int64), g["neighbors"]
    assert levels.shape[0] == ntotal and offsets.shape[0] == ntotal + 1
    assert levels.min() >= 1
    cum = [0, 2 * M]
    while len(cum) < levels.max() + 1:
        cum.append(cum[-1] + M)
    assert np.array_equal(np.diff(offsets), np.array([cum[l] for l in levels], dtype=np.int64))
    assert int(levels[g["entry_point"]]) - 1 == g["max_level"] == int(levels.max()) - 1
    assert nb.min() >= -1 and nb.max() < ntotal
    for i in range(ntotal):
        for l in range(levels[i]):
            row = nb[offsets[i] + cum[l]: offsets[i] + cum[l + 1]]
            valid = row[row >= 0]
            nv = len(valid)
            assert np.all(row[:nv] >= 0) and np.all(row[nv:] == -1), f"row {i}/{l} not packed"
            assert i not in valid, f"self loop at {i}/{l}"
            assert len(set(valid.tolist())) == nv, f"duplicate neighbour at {i}/{l}"
            # an edge at level l must point to a vertex that exists at level l
            assert np.all(levels[valid] > l), f"edge to a vertex below level {l} at {i}"
