"""CPU tests of the oracle (faiss-semantics restatement).

PARITY UNPINNED against faiss itself (no faiss, no reference tests/goldens exist — SURVEY §8c).
What pins the oracle: mt19937 known answers, the level law, graph invariants, brute-force recall.
"""
import numpy as np
import pytest

from conftest import assert_graph_invariants
from hnsw_b200.datasets import synthetic_dataset


def test_mt19937_known_answers(oracle_mod):
    # std::mt19937 is standard-defined: 10000th output of default seed 5489 is 4123659995;
    # SURVEY §"What is on this machine" anchors seed 12345 -> 3992670690.
    assert oracle_mod.mt19937_first(12345) == 3992670690
    assert oracle_mod.mt19937_first(5489) == 3499211612


def test_level_tables(oracle_mod):
    for M in (4, 16, 32, 64):
        o = oracle_mod.OracleHNSWFlat(8, M)
        p, cum = o.tables()
        mult = np.float32(1.0 / np.log(M))
        want = [np.exp(-l / mult) * (1 - np.exp(-1 / mult)) for l in range(len(p))]
        assert np.allclose(p, want, rtol=1e-6)
        assert p[-1] >= 1e-9 and abs(p.sum() - 1) < 1e-6
        assert cum[0] == 0 and cum[1] == 2 * M and np.all(np.diff(cum[1:]) == M)


def test_level_histogram_follows_geometric_law(oracle_mod):
    M = 16
    o = oracle_mod.OracleHNSWFlat(4, M)
    lv = o.peek_levels(200000) - 1
    n = len(lv)
    for l in range(3):
        frac = (lv >= l).mean()
        assert abs(frac - M ** (-l)) < 4 * np.sqrt(M ** (-l) / n) + 1e-4
    # peek does not consume the index RNG: levels of a subsequent add are the same
    x = np.zeros((50, 4), np.float32)
    o.add(x)
    assert np.array_equal(o.export_graph()["levels"], lv[:50] + 1)


def test_team_order_distance_matches_float64(oracle_mod):
    rs = np.random.RandomState(0)
    for d, T in ((128, 8), (96, 8), (960, 32), (768, 32), (256, 16), (4, 8)):
        a, b = rs.randn(d).astype(np.float32), rs.randn(d).astype(np.float32)
        for metric in (oracle_mod.METRIC_L2, oracle_mod.METRIC_INNER_PRODUCT):
            o = oracle_mod.OracleHNSWFlat(d, 16, metric)
            ref = ((a.astype(np.float64) - b) ** 2).sum() if metric == 1 else -(a.astype(np.float64) * b).sum()
            nat = o.distance(a, b)
            o.set_team(T)
            tm = o.distance(a, b)
            scale = max(abs(ref), (np.abs(a) * np.abs(b)).sum() if metric == 0 else 1e-30)
            assert abs(nat - ref) <= 1e-5 * scale and abs(tm - ref) <= 1e-5 * scale


def test_graph_invariants_and_recall_l2(oracle_mod, small_l2):
    o, g = small_l2["oracle"], small_l2["graph"]
    assert_graph_invariants(g, 16, 4000)
    _, gt = oracle_mod.brute_force_knn(small_l2["xb"], small_l2["xq"], 10)
    D, I = o.search(small_l2["xq"], 10, 64)
    assert oracle_mod.recall_at_k(I, gt) >= 0.97           # upstream-style recall floor
    assert np.all(np.diff(D, axis=1) >= 0)                  # ascending
    # returned distances are the true squared L2 of the returned ids (1e-4 rel, BASELINE north_star)
    xb, xq = small_l2["xb"].astype(np.float64), small_l2["xq"].astype(np.float64)
    ref = ((xq[:, None, :] - xb[I]) ** 2).sum(-1)
    assert np.allclose(D, ref, rtol=1e-4, atol=1e-6)


def test_native_and_team_order_build_same_graph(oracle_mod):
    xb, _ = synthetic_dataset(32, 1500, 1)
    a = oracle_mod.OracleHNSWFlat(32, 8)
    b = oracle_mod.OracleHNSWFlat(32, 8)
    b.set_team(8)
    a.add(xb)
    b.add(xb)
    ga, gb = a.export_graph(), b.export_graph()
    assert np.array_equal(ga["levels"], gb["levels"])
    # summation order may flip exact near-ties; the graphs must agree almost everywhere
    assert (ga["neighbors"] != gb["neighbors"]).mean() < 0.01


def test_inner_product_metric(oracle_mod):
    xb, xq = synthetic_dataset(48, 3000, 50, normalize=True)
    o = oracle_mod.OracleHNSWFlat(48, 16, oracle_mod.METRIC_INNER_PRODUCT)
    o.add(xb)
    D, I = o.search(xq, 10, 64)
    _, gt = oracle_mod.brute_force_knn(xb, xq, 10, oracle_mod.METRIC_INNER_PRODUCT)
    assert oracle_mod.recall_at_k(I, gt) >= 0.9
    assert np.all(np.diff(D, axis=1) <= 0)                  # similarities, best (largest) first
    ref = (xq[:, None, :].astype(np.float64) * xb[I]).sum(-1)
    assert np.allclose(D, ref, rtol=1e-4, atol=1e-5)


def test_parallel_build_recall_matches_sequential(oracle_mod):
    xb, xq = synthetic_dataset(32, 6000, 100)
    _, gt = oracle_mod.brute_force_knn(xb, xq, 10)
    rec = []
    for thr in (1, 4):
        o = oracle_mod.OracleHNSWFlat(32, 16)
        o.threads = thr
        o.add(xb)
        assert_graph_invariants(o.export_graph(), 16, 6000)
        rec.append(oracle_mod.recall_at_k(o.search(xq, 10, 64)[1], gt))
    assert abs(rec[0] - rec[1]) < 0.01 and min(rec) > 0.95


def test_edge_cases(oracle_mod):
    d = 16
    o = oracle_mod.OracleHNSWFlat(d, 8)
    xq = np.random.RandomState(1).randn(3, d).astype(np.float32)
    D, I = o.search(xq, 5)                                  # empty index
    assert np.all(I == -1) and np.all(D == np.finfo(np.float32).max)
    xb = np.random.RandomState(2).randn(3, d).astype(np.float32)
    o.add(xb)                                               # ntotal < k -> padded with -1
    D, I = o.search(xq, 5)
    assert np.all(I[:, :3] >= 0) and np.all(I[:, 3:] == -1)
    assert np.all(np.sort(I[:, :3], axis=1) == np.arange(3))
    o.add(np.zeros((40, d), np.float32))                    # duplicate (all-zero) vectors
    D, I = o.search(np.zeros((1, d), np.float32), 10, 32)
    assert np.all(D[0] == 0) and len(set(I[0].tolist())) == 10
    # k > efSearch: the list capacity becomes k (App. A.5)
    xb2, xq2 = synthetic_dataset(d, 2000, 20)
    o2 = oracle_mod.OracleHNSWFlat(d, 8)
    o2.add(xb2)
    D, I = o2.search(xq2, 50, 8)
    assert np.all(I >= 0) and np.all(np.diff(D, axis=1) >= 0)
    # incremental add keeps invariants
    o2.add(xb2[:500] + 0.01)
    assert_graph_invariants(o2.export_graph(), 8, 2500)


def test_import_export_roundtrip(oracle_mod, small_l2):
    g = small_l2["graph"]
    o2 = oracle_mod.OracleHNSWFlat(32, 16)
    o2.set_team(8)
    o2.import_graph(small_l2["xb"], g["levels"], g["neighbors"], g["entry_point"], g["max_level"])
    D1, I1 = small_l2["oracle"].search(small_l2["xq"], 10, 32)
    D2, I2 = o2.search(small_l2["xq"], 10, 32)
    assert np.array_equal(I1, I2) and np.array_equal(D1, D2)


def test_shrink_heuristic_small_case(oracle_mod):
    # three collinear points: the middle one shadows the far one (App. A.10)
    xb = np.array([[1, 0], [2, 0], [0, 1.5], [0, 0]], np.float32)
    xb = np.pad(xb, ((0, 0), (0, 2)))
    o = oracle_mod.OracleHNSWFlat(4, 4)
    o.add(xb)
    dq = ((xb[:3] - xb[3]) ** 2).sum(1).astype(np.float32)    # distances to the origin (id 3)
    kept = o.shrink(np.array([0, 1, 2], np.int32), dq, 3)
    assert kept.tolist() == [0, 2]      # id 1 is closer to id 0 than to the base -> pruned
    kept = o.shrink(np.array([0, 1, 2], np.int32), dq, 4)     # fewer than max_size -> unchanged
    assert sorted(kept.tolist()) == [0, 1, 2]


def test_golden_fixture(oracle_mod):
    """The committed fixture was produced by tests/golden/make_golden.py from this oracle; it
    guards the oracle against silent drift (it does not pin it to faiss — nothing can, here)."""
    import os
    f = os.path.join(os.path.dirname(__file__), "golden", "oracle_l2_d32_n3000_M16.npz")
    z = np.load(f)
    xb, xq = synthetic_dataset(32, 3000, 64)
    o = oracle_mod.OracleHNSWFlat(32, 16)
    o.set_team(8)
    o.add(xb)
    g = o.export_graph()
    assert np.array_equal(g["levels"], z["levels"])
    assert np.array_equal(g["neighbors"], z["neighbors"])
    assert g["entry_point"] == int(z["entry_point"]) and g["max_level"] == int(z["max_level"])
    D, I, st = o.search(xq, 10, 48, stats=True)
    assert np.array_equal(I, z["I"]) and np.array_equal(D, z["D"]) and np.array_equal(st, z["stats"])
