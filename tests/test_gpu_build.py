"""GPU tests of batched graph construction (insertion searches + neighbour-selection heuristic +
back-links), through the C-ABI.

* max_batch=1 replays faiss's sequential insertion order one point per round: with the bit-exact
  distance order the resulting graph must EQUAL the oracle's single-threaded graph.
* batched rounds are a different (concurrent) schedule, like faiss's OpenMP build: checked by graph
  invariants and by recall within 0.5 points of the oracle's graph at equal M / ef (north_star).
"""
import numpy as np
import pytest

from conftest import assert_graph_invariants
from hnsw_b200.datasets import synthetic_dataset

pytestmark = pytest.mark.gpu


def _build_gpu(xb, M, efc, metric=1, **bp):
    import hnsw_b200
    idx = hnsw_b200.IndexHNSWFlat(xb.shape[1], M, metric)
    idx.hnsw.efConstruction = efc
    if bp:
        idx.set_build_params(**bp)
    idx.add(xb)
    return idx


@pytest.mark.parametrize("d,M,efc,team,metric,n", [(32, 16, 40, 8, 1, 2500), (128, 8, 24, 8, 1, 1200),
                                                   (64, 4, 16, 8, 0, 1500), (960, 8, 16, 32, 1, 400)])
def test_sequential_build_equals_oracle_graph(oracle_mod, d, M, efc, team, metric, n):
    xb, xq = synthetic_dataset(d, n, 50, normalize=(metric == 0))
    o = oracle_mod.OracleHNSWFlat(d, M, metric)
    o.set_team(team)
    o.efConstruction = efc
    o.add(xb)
    go = o.export_graph()
    idx = _build_gpu(xb, M, efc, metric, max_batch=1)
    gg = idx.export_graph()
    assert np.array_equal(gg["levels"], go["levels"])            # same mt19937(12345) level draw
    assert gg["entry_point"] == go["entry_point"] and gg["max_level"] == go["max_level"]
    mism = np.flatnonzero(gg["neighbors"] != go["neighbors"])
    assert mism.size == 0, f"{mism.size} adjacency slots differ, first at {mism[:5]}"
    Do, Io = o.search(xq, 10, 32)
    D, I = idx.search(xq, 10, efSearch=32)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)


def test_sequential_build_in_two_add_calls(oracle_mod):
    xb, _ = synthetic_dataset(32, 1800, 1)
    o = oracle_mod.OracleHNSWFlat(32, 8)
    o.set_team(8)
    o.add(xb[:1000])
    o.add(xb[1000:])
    import hnsw_b200
    idx = hnsw_b200.IndexHNSWFlat(32, 8)
    idx.set_build_params(max_batch=1)
    idx.add(xb[:1000])
    idx.add(xb[1000:])
    go, gg = o.export_graph(), idx.export_graph()
    assert np.array_equal(gg["levels"], go["levels"]) and np.array_equal(gg["neighbors"], go["neighbors"])
    assert gg["entry_point"] == go["entry_point"]


@pytest.mark.parametrize("metric", [1, 0])
def test_batched_build_invariants_and_recall(oracle_mod, metric):
    """north_star: recall@10 within 0.5 points of the CPU build at equal M / ef. Both builds are
    deterministic here (single-threaded oracle, scheduling-independent GPU rounds)."""
    d, M, efc, n = 64, 16, 64, 20000
    xb, xq = synthetic_dataset(d, n, 5000, d1=12, normalize=(metric == 0))
    _, gt = oracle_mod.brute_force_knn(xb, xq, 10, metric)
    o = oracle_mod.OracleHNSWFlat(d, M, metric)
    o.efConstruction = efc
    o.add(xb)
    o.threads = 8
    idx = _build_gpu(xb, M, efc, metric)
    assert idx.ntotal == n
    assert_graph_invariants(idx.export_graph(), M, n)
    for ef in (32, 64, 128):
        r_cpu = oracle_mod.recall_at_k(o.search(xq, 10, ef)[1], gt)
        r_gpu = oracle_mod.recall_at_k(idx.search(xq, 10, efSearch=ef)[1], gt)
        assert r_gpu >= r_cpu - 0.005, f"ef={ef}: GPU-built recall {r_gpu:.4f} vs CPU-built {r_cpu:.4f}"


def test_batched_build_is_deterministic():
    xb, _ = synthetic_dataset(32, 6000, 1)
    a = _build_gpu(xb, 16, 40).export_graph()
    b = _build_gpu(xb, 16, 40).export_graph()
    assert np.array_equal(a["neighbors"], b["neighbors"])


def test_add_with_preset_levels_and_order(oracle_mod):
    xb, _ = synthetic_dataset(32, 900, 1)
    o = oracle_mod.OracleHNSWFlat(32, 8)
    o.set_team(8)
    lv = o.peek_levels(900)
    order = o.add(xb, return_order=True)
    import hnsw_b200
    idx = hnsw_b200.IndexHNSWFlat(32, 8)
    idx.set_build_params(max_batch=1)
    idx.add(xb, levels=lv, order=order)
    assert np.array_equal(idx.export_graph()["neighbors"], o.export_graph()["neighbors"])
    with pytest.raises(RuntimeError):
        idx.add(xb[:3], order=np.array([0, 1, 2], np.int32))      # not the new ids


def test_add_on_top_of_imported_graph(oracle_mod):
    """Imported rows have no verified prefix (nver = 0), so their first shrinks take the full
    heuristic path; the result must still equal the oracle continuing the same build."""
    import hnsw_b200
    d, M = 32, 8
    xb, xq = synthetic_dataset(d, 2200, 40)
    o = oracle_mod.OracleHNSWFlat(d, M)
    o.set_team(8)
    o.add(xb[:1200])
    g0 = o.export_graph()
    idx = hnsw_b200.IndexHNSWFlat(d, M)
    idx.import_graph(xb[:1200], g0["levels"], g0["neighbors"], g0["entry_point"], g0["max_level"])
    lv = o.peek_levels(1000)
    order = o.add(xb[1200:], return_order=True)
    idx.set_build_params(max_batch=1)
    idx.add(xb[1200:], levels=lv, order=order)
    go, gg = o.export_graph(), idx.export_graph()
    assert np.array_equal(gg["levels"], go["levels"])
    mism = np.flatnonzero(gg["neighbors"] != go["neighbors"])
    assert mism.size == 0, f"{mism.size} adjacency slots differ"
    Do, Io = o.search(xq, 10, 40)
    D, I = idx.search(xq, 10, efSearch=40)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)


@pytest.mark.parametrize("d,M,team,n", [(64, 64, 8, 900), (2048, 4, 32, 250)])
def test_sequential_build_wide_rows_and_wide_vectors(oracle_mod, d, M, team, n):
    xb, xq = synthetic_dataset(d, n, 20)
    o = oracle_mod.OracleHNSWFlat(d, M)
    o.set_team(team)
    o.efConstruction = 48
    o.add(xb)
    idx = _build_gpu(xb, M, 48, 1, max_batch=1)
    go, gg = o.export_graph(), idx.export_graph()
    assert np.array_equal(gg["neighbors"], go["neighbors"])
    Do, Io = o.search(xq, 10, 64)
    D, I = idx.search(xq, 10, efSearch=64)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)


@pytest.mark.parametrize("d,M,team,metric", [(128, 16, 8, 1), (64, 8, 8, 0), (256, 8, 8, 1), (512, 8, 16, 1), (768, 8, 32, 0)])
def test_fp16_storage_bit_exact_vs_oracle_half_mode(oracle_mod, d, M, team, metric):
    """Opt-in fp16 vector storage: same algorithm on rows rounded to binary16 (fp32 accumulate).
    The oracle's half-storage mode emulates it exactly: sequential build and search stay bit-equal,
    reconstruct returns the rounded rows, and recall vs the fp32 ground truth stays high."""
    import hnsw_b200
    n = 1500
    xb, xq = synthetic_dataset(d, n, 60, normalize=(metric == 0))
    o = oracle_mod.OracleHNSWFlat(d, M, metric)
    o.set_half_storage(True)
    o.set_team(team)
    o.efConstruction = 32
    o.add(xb)
    idx = hnsw_b200.IndexHNSWFlat(d, M, metric, storage="fp16")
    idx.hnsw.efConstruction = 32
    idx.set_build_params(max_batch=1)
    idx.add(xb)
    go, gg = o.export_graph(), idx.export_graph()
    assert np.array_equal(gg["levels"], go["levels"])
    mism = np.flatnonzero(gg["neighbors"] != go["neighbors"])
    assert mism.size == 0, f"{mism.size} adjacency slots differ"
    for ef in (16, 64):
        Do, Io, So = o.search(xq, 10, ef, stats=True)
        D, I, S = idx.search(xq, 10, efSearch=ef, stats=True, hash_bits=13)
        assert np.array_equal(I, Io) and np.array_equal(D, Do) and np.array_equal(S, So)
    assert np.array_equal(idx.reconstruct_n(0, 50), xb[:50].astype(np.float16).astype(np.float32))
    _, gt = oracle_mod.brute_force_knn(xb, xq, 10, metric)
    assert oracle_mod.recall_at_k(idx.search(xq, 10, efSearch=128)[1], gt) > 0.93
    # batched build in fp16 mode: invariants hold
    b = hnsw_b200.IndexHNSWFlat(d, M, metric, storage="fp16")
    b.hnsw.efConstruction = 32
    b.add(xb)
    assert_graph_invariants(b.export_graph(), M, n)
    # d % 8 != 0: fp16 rows are zero-padded to whole 16-byte chunks; sequential build == oracle half mode
    # on explicitly padded data
    x12, q12 = synthetic_dataset(12, 600, 20)
    pad = lambda a: np.ascontiguousarray(np.pad(a, ((0, 0), (0, 4))))
    o12 = oracle_mod.OracleHNSWFlat(16, 4, 1)
    o12.set_team(8)
    o12.set_half_storage(True)
    o12.add(pad(x12))
    i12 = hnsw_b200.IndexHNSWFlat(12, 4, 1, storage="fp16")
    i12.set_build_params(max_batch=1)
    i12.add(x12)
    assert np.array_equal(i12.export_graph()["neighbors"], o12.export_graph()["neighbors"])
    Do, Io = o12.search(pad(q12), 5, 32)
    D, I = i12.search(q12, 5, efSearch=32)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    assert np.array_equal(i12.reconstruct(3), x12[3].astype(np.float16).astype(np.float32))


def test_rejected_add_leaves_index_untouched(oracle_mod):
    """A rejected add() (bad order / bad preset level) must not advance the level RNG nor change
    ntotal: the next add() still reproduces the oracle's graph."""
    import hnsw_b200
    xb, _ = synthetic_dataset(32, 700, 1)
    o = oracle_mod.OracleHNSWFlat(32, 8)
    o.set_team(8)
    o.add(xb)
    idx = hnsw_b200.IndexHNSWFlat(32, 8)
    idx.set_build_params(max_batch=1)
    with pytest.raises(RuntimeError):
        idx.add(xb[:5], order=np.array([0, 1, 2, 3, 3], np.int32))
    with pytest.raises(RuntimeError):
        idx.add(xb[:5], levels=np.array([1, 1, 99, 1, 1], np.int32))
    assert idx.ntotal == 0
    idx.add(xb)
    go, gg = o.export_graph(), idx.export_graph()
    assert np.array_equal(gg["levels"], go["levels"]) and np.array_equal(gg["neighbors"], go["neighbors"])


def _round_bf16(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("d,M,team,metric", [(128, 16, 8, 1), (64, 8, 8, 0), (768, 8, 32, 0), (20, 8, 8, 1)])
def test_bf16_storage_bit_exact_vs_oracle_bf16_mode(oracle_mod, d, M, team, metric):
    """Opt-in bfloat16 vector storage (rows rounded to nearest-even bf16 on add, exact widening, fp32
    accumulation in the kernels' fixed order): sequential build and search are bit-equal to the oracle's bf16
    mode, reconstruct returns the rounded rows. d=20 also exercises zero-padding to whole 16-byte chunks."""
    import hnsw_b200
    n = 1200
    dp = (d + 7) // 8 * 8
    xb, xq = synthetic_dataset(d, n, 50, normalize=(metric == 0))
    pad = lambda a: np.ascontiguousarray(np.pad(a, ((0, 0), (0, dp - d))))
    o = oracle_mod.OracleHNSWFlat(dp, M, metric)
    o.set_half_storage("bf16")
    o.set_team(team)
    o.efConstruction = 32
    o.add(pad(xb))
    idx = hnsw_b200.IndexHNSWFlat(d, M, metric, storage="bf16")
    assert idx.storage == "bf16"
    idx.hnsw.efConstruction = 32
    idx.set_build_params(max_batch=1)
    idx.add(xb)
    go, gg = o.export_graph(), idx.export_graph()
    assert np.array_equal(gg["levels"], go["levels"])
    # "identical except for distance ties" (BASELINE north_star): 8 mantissa bits make two neighbours at exactly
    # the same distance from a vertex likelier; faiss's heap and the kernel's (distance, id) order may then list
    # them in either order. Any differing slot must be such a swap: same neighbour SET in that row, both
    # neighbours at the identical distance from the row's owner.
    mism = np.flatnonzero(gg["neighbors"] != go["neighbors"])
    assert mism.size <= 4, f"{mism.size} of {go['neighbors'].size} adjacency slots differ (first at {mism[:5]})"
    if mism.size:
        offs = go["offsets"].astype(np.int64)
        xr = _round_bf16(pad(xb))
        for v in np.unique(np.searchsorted(offs, mism, side="right") - 1):
            a, b2 = gg["neighbors"][offs[v]:offs[v + 1]], go["neighbors"][offs[v]:offs[v + 1]]
            assert np.array_equal(np.sort(a), np.sort(b2)), f"row of vertex {v}: different neighbour sets"
            for sl in np.flatnonzero(a != b2):
                assert o.distance(xr[v], xr[a[sl]]) == o.distance(xr[v], xr[b2[sl]]), "swap without a distance tie"
        idx.import_graph(xb, go["levels"], go["neighbors"], go["entry_point"], go["max_level"])  # search the oracle's graph
    for ef in (16, 64):
        Do, Io, So = o.search(pad(xq), 10, ef, stats=True)
        D, I, S = idx.search(xq, 10, efSearch=ef, stats=True, hash_bits=13)
        assert np.array_equal(I, Io) and np.array_equal(D, Do) and np.array_equal(S, So)
        D, I = idx.search(xq, 10, efSearch=ef)                      # default visited table
        assert np.array_equal(I, Io) and np.array_equal(D, Do)
    assert np.array_equal(idx.reconstruct_n(0, 40), _round_bf16(xb[:40]))
    _, gt = oracle_mod.brute_force_knn(xb, xq, 10, metric)
    assert oracle_mod.recall_at_k(idx.search(xq, 10, efSearch=128)[1], gt) > 0.85
    b = hnsw_b200.IndexHNSWFlat(d, M, metric, storage="bf16")       # batched build: invariants hold
    b.hnsw.efConstruction = 32
    b.add(xb)
    assert_graph_invariants(b.export_graph(), M, n)
