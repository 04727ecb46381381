"""GPU parity tests of the search path (greedy descent + level-0 beam search), through the C-ABI.

The oracle's team mode reproduces the CUDA kernels' summation order bit for bit, so on a graph
imported from the oracle the GPU must return IDENTICAL ids, distances and per-query counters
(stronger than BASELINE.json's "identical except for distance ties" / 1e-4 relative)."""
import numpy as np
import pytest

from hnsw_b200.datasets import synthetic_dataset

pytestmark = pytest.mark.gpu


def _gpu_from_oracle(o, xb, M, metric=1):
    import hnsw_b200
    g = o.export_graph()
    idx = hnsw_b200.IndexHNSWFlat(xb.shape[1], M, metric)
    idx.import_graph(xb, g["levels"], g["neighbors"], g["entry_point"], g["max_level"])
    return idx


def test_import_export_roundtrip(small_l2):
    idx = _gpu_from_oracle(small_l2["oracle"], small_l2["xb"], 16)
    g, g2 = small_l2["graph"], idx.export_graph()
    for key in ("levels", "offsets", "neighbors"):
        assert np.array_equal(g[key], g2[key])
    assert g2["entry_point"] == g["entry_point"] and g2["max_level"] == g["max_level"]
    assert np.array_equal(idx.reconstruct(17), small_l2["xb"][17])


@pytest.mark.parametrize("ef", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("W", [1, 2, 4, 8])
def test_search_bit_exact_vs_oracle(small_l2, ef, W):
    idx = _gpu_from_oracle(small_l2["oracle"], small_l2["xb"], 16)
    Do, Io, So = small_l2["oracle"].search(small_l2["xq"], 10, ef, stats=True)
    # hash_bits=14: the visited table never fills, so the distance-evaluation counts match too
    D, I, S = idx.search(small_l2["xq"], 10, efSearch=ef, stats=True, warps_per_query=W, hash_bits=13)
    assert np.array_equal(I, Io)
    assert np.array_equal(D, Do)
    assert np.array_equal(S, So)          # ndis / nhops at level 0 and above: same path
    # default (small, forgetful) table: identical results, hop count identical, ndis may only grow
    D, I, S = idx.search(small_l2["xq"], 10, efSearch=ef, stats=True, warps_per_query=W)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    assert np.array_equal(S[:, 1:], So[:, 1:]) and np.all(S[:, 0] >= So[:, 0])


def test_distances_within_1e4_of_native_simd_oracle(oracle_mod, small_l2):
    """BASELINE north_star tolerance: returned distances within 1e-4 relative of the CPU path
    that uses its own (SIMD) summation order."""
    g = small_l2["graph"]
    nat = oracle_mod.OracleHNSWFlat(32, 16)
    nat.import_graph(small_l2["xb"], g["levels"], g["neighbors"], g["entry_point"], g["max_level"])
    Dn, In = nat.search(small_l2["xq"], 10, 64)
    idx = _gpu_from_oracle(small_l2["oracle"], small_l2["xb"], 16)
    D, I = idx.search(small_l2["xq"], 10, efSearch=64)
    same = I == In
    assert same.mean() > 0.99                      # only exact-tie flips may differ
    assert np.allclose(D[same], Dn[same], rtol=1e-4, atol=1e-7)


def test_small_hash_forces_resets_but_not_result_changes(small_l2):
    idx = _gpu_from_oracle(small_l2["oracle"], small_l2["xb"], 16)
    Do, Io, So = small_l2["oracle"].search(small_l2["xq"], 10, 128, stats=True)
    for hb in (8, 9, 10):
        D, I, S = idx.search(small_l2["xq"], 10, efSearch=128, stats=True, hash_bits=hb)
        assert np.array_equal(I, Io) and np.array_equal(D, Do)


@pytest.mark.parametrize("W", [1, 4])
def test_set_associative_visited_table_never_changes_results(small_l2, W):
    """Default policy (visited_policy=2): 16-byte buckets with FIFO eviction. Down to a 64-byte table
    (4 buckets: almost every vertex is forgotten and re-scored, list members included, so the merge's
    identical-key rejection is exercised on nearly every hop) the ids and distances stay the oracle's;
    the hop count is the oracle's; only ndis grows."""
    idx = _gpu_from_oracle(small_l2["oracle"], small_l2["xb"], 16)
    for ef in (16, 128):
        Do, Io, So = small_l2["oracle"].search(small_l2["xq"], 10, ef, stats=True)
        prev = None
        for hb in (4, 6, 8, 11):
            D, I, S = idx.search(small_l2["xq"], 10, efSearch=ef, stats=True, hash_bits=hb, visited_policy=2,
                                 warps_per_query=W)
            assert np.array_equal(I, Io) and np.array_equal(D, Do), (ef, hb)
            assert np.array_equal(S[:, 1:], So[:, 1:]) and np.all(S[:, 0] >= So[:, 0])
            if prev is not None:
                assert S[:, 0].sum() <= prev          # a larger table forgets less
            prev = S[:, 0].sum()
        # 2 KB of 16-bit slots remember 1024 vertices: on this 4000-vertex graph that is nearly exact
        assert S[:, 0].sum() <= 1.02 * So[:, 0].sum()


def test_k_larger_than_efsearch_and_unbounded_steps(small_l2):
    o = small_l2["oracle"]
    idx = _gpu_from_oracle(o, small_l2["xb"], 16)
    Do, Io, So = o.search(small_l2["xq"], 40, 8, stats=True)     # ef = k, count_below can fire
    D, I, S = idx.search(small_l2["xq"], 40, efSearch=8, stats=True, hash_bits=13)
    assert np.array_equal(I, Io) and np.array_equal(D, Do) and np.array_equal(S, So)
    o.set_check_relative_distance(False)
    try:
        Do, Io, So = o.search(small_l2["xq"], 10, 24, stats=True)
    finally:
        o.set_check_relative_distance(True)
    idx.hnsw.check_relative_distance = False
    D, I, S = idx.search(small_l2["xq"], 10, efSearch=24, stats=True, hash_bits=13)
    assert np.array_equal(I, Io) and np.array_equal(D, Do) and np.array_equal(S, So)


@pytest.mark.parametrize("d,M,team,metric", [(96, 16, 8, 1), (128, 32, 8, 1), (256, 8, 16, 1),
                                             (512, 8, 32, 1), (960, 8, 32, 1), (768, 16, 32, 0),
                                             (128, 64, 8, 0), (4, 4, 8, 1)])
def test_dims_metrics_and_degrees(oracle_mod, d, M, team, metric):
    xb, xq = synthetic_dataset(d, 1500, 40, normalize=(metric == 0))
    o = oracle_mod.OracleHNSWFlat(d, M, metric)
    o.set_team(team)
    o.efConstruction = 32
    o.add(xb)
    idx = _gpu_from_oracle(o, xb, M, metric)
    for ef in (16, 100):
        Do, Io, So = o.search(xq, 10, ef, stats=True)
        D, I, S = idx.search(xq, 10, efSearch=ef, stats=True, hash_bits=13)
        assert np.array_equal(I, Io) and np.array_equal(D, Do) and np.array_equal(S, So)
        D, I = idx.search(xq, 10, efSearch=ef)
        assert np.array_equal(I, Io) and np.array_equal(D, Do)


def test_edge_cases(oracle_mod):
    import hnsw_b200
    d = 16
    idx = hnsw_b200.IndexHNSWFlat(d, 8)
    xq = np.random.RandomState(1).randn(3, d).astype(np.float32)
    D, I = idx.search(xq, 5)                                   # empty index
    assert np.all(I == -1) and np.all(D == np.finfo(np.float32).max)
    D, I = idx.search(xq[:0], 5)                               # zero queries
    assert D.shape == (0, 5)
    with pytest.raises(RuntimeError):
        idx.search(xq, 0)
    with pytest.raises(ValueError):
        idx.search(np.zeros((2, d + 4), np.float32), 5)
    xb = np.random.RandomState(2).randn(3, d).astype(np.float32)
    o = oracle_mod.OracleHNSWFlat(d, 8)
    o.set_team(8)
    o.add(xb)
    g = o.export_graph()
    idx.import_graph(xb, g["levels"], g["neighbors"], g["entry_point"], g["max_level"])
    Do, Io = o.search(xq, 5)
    D, I = idx.search(xq, 5)                                   # ntotal < k: -1 / FLT_MAX padding
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
    idx.reset()
    assert idx.ntotal == 0


def test_large_query_batch_and_ragged_tail(small_l2):
    o = small_l2["oracle"]
    idx = _gpu_from_oracle(o, small_l2["xb"], 16)
    rs = np.random.RandomState(5)
    xq = small_l2["xb"][rs.randint(0, 4000, 5003)] + 0.01 * rs.randn(5003, 32).astype(np.float32)
    o.threads = 8
    Do, Io = o.search(xq, 7, 40)
    o.threads = 1
    D, I = idx.search(xq, 7, efSearch=40)
    assert np.array_equal(I, Io) and np.array_equal(D, Do)


def test_merge_topk_matches_numpy():
    import ctypes as C
    import torch
    from hnsw_b200 import _lib
    rs = np.random.RandomState(3)
    for metric in (1, 0):
        ns, nq, k = 5, 257, 10
        D = np.sort(rs.rand(ns, nq, k).astype(np.float32), axis=2)
        D[:, :, -2:][rs.rand(ns, nq, 2) < 0.3] = np.finfo(np.float32).max     # some empty slots
        D = np.sort(D, axis=2)
        I = rs.randint(0, 1000, (ns, nq, k)).astype(np.int64)
        I[D == np.finfo(np.float32).max] = -1
        if metric == 0:
            D = -D
        off = np.arange(ns, dtype=np.int64) * 1000
        Dd, Id = torch.from_numpy(D).cuda(), torch.from_numpy(I).cuda()
        Do = torch.empty(nq, k, device="cuda")
        Io = torch.empty(nq, k, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        _lib.check(_lib.lib().bh_merge_topk_device(ns, nq, k, metric, Dd.data_ptr(), Id.data_ptr(),
                                                   off.ctypes.data, Do.data_ptr(), Io.data_ptr(), None))
        allD = D.transpose(1, 0, 2).reshape(nq, ns * k)
        allI = np.where(I >= 0, I + off[:, None, None], -1).transpose(1, 0, 2).reshape(nq, ns * k)
        order = np.argsort(allD if metric == 1 else -allD, axis=1, kind="stable")[:, :k]
        assert np.array_equal(Do.cpu().numpy(), np.take_along_axis(allD, order, 1))
        assert np.array_equal(Io.cpu().numpy(), np.take_along_axis(allI, order, 1))


def test_sharded_index_single_rank_nccl(small_l2):
    """The NCCL code path of ShardedIndexHNSWFlat on a 1-rank group equals the plain index."""
    import socket
    import torch
    import torch.distributed as dist
    import hnsw_b200
    from hnsw_b200.sharded import ShardedIndexHNSWFlat
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=dev)
    try:
        sh = ShardedIndexHNSWFlat(32, 16, 1, device=dev)
        sh.add(small_l2["xb"])
        ref = hnsw_b200.IndexHNSWFlat(32, 16)
        ref.add(small_l2["xb"])
        D, I = sh.search(small_l2["xq"], 10, efSearch=64)
        torch.cuda.synchronize()
        Dr, Ir = ref.search(small_l2["xq"], 10, efSearch=64)
        assert np.array_equal(I.cpu().numpy(), Ir) and np.array_equal(D.cpu().numpy(), Dr)
        assert sh.ntotal == 4000
    finally:
        dist.destroy_process_group()


def test_pinned_host_buffers_zero_copy_path(small_l2):
    """Page-locked caller buffers take the zero-copy path (kernel reads/writes host memory);
    results must equal the staged path."""
    import torch
    idx = _gpu_from_oracle(small_l2["oracle"], small_l2["xb"], 16)
    xq = small_l2["xq"]
    D0, I0 = idx.search(xq, 10, efSearch=48)
    xp = torch.empty(xq.shape, dtype=torch.float32).pin_memory()
    xp.copy_(torch.from_numpy(xq))
    Dp = torch.empty(len(xq), 10, dtype=torch.float32).pin_memory()
    Ip = torch.empty(len(xq), 10, dtype=torch.int64).pin_memory()
    D1, I1 = idx.search(xp.numpy(), 10, efSearch=48, out=(Dp.numpy(), Ip.numpy()))
    assert np.array_equal(I1, I0) and np.array_equal(D1, D0)


def test_baseline_config0_plumbing(oracle_mod):
    """BASELINE.json configs[0]: IndexHNSWFlat M=16 efC=40 on 100k x 128 synthetic fp32 L2, 1k queries,
    k=10, efSearch=64 — the reference's own CPU-runnable case. CPU-built graph searched on the GPU:
    identical ids / distances; GPU-built graph: recall within 0.5 pt of the CPU build."""
    import hnsw_b200
    d, M, n, nq, k, ef = 128, 16, 100_000, 1000, 10, 64
    xb, xq_all = synthetic_dataset(d, n, nq + 5000)      # upstream recipe: d1=10, seed=1338
    xq, xq2 = xq_all[:nq], xq_all[nq:]
    o = oracle_mod.OracleHNSWFlat(d, M)
    o.efConstruction = 40
    o.threads = 8
    o.add(xb)
    o.set_team(8)                                         # search with the CUDA summation order
    Do, Io, So = o.search(xq, k, ef, stats=True)
    idx = _gpu_from_oracle(o, xb, M)
    D, I, S = idx.search(xq, k, efSearch=ef, stats=True, hash_bits=13)
    assert np.array_equal(I, Io) and np.array_equal(D, Do) and np.array_equal(S, So)
    # recall of a GPU-built graph vs the CPU-built one: two different graphs, so compare on a larger
    # query sample (5000) than the config's 1k - with 1k queries the sampling noise alone is ~0.3 pt
    import torch
    from hnsw_b200.datasets import exact_knn_torch
    _, gt = exact_knn_torch(torch.from_numpy(xb).cuda(), torch.from_numpy(xq2).cuda(), k)
    gt = gt.cpu().numpy()
    g = hnsw_b200.IndexHNSWFlat(d, M)
    g.hnsw.efConstruction = 40
    g.add(xb)
    o.set_team(0)
    r_cpu = oracle_mod.recall_at_k(o.search(xq2, k, ef)[1], gt)
    r_gpu = oracle_mod.recall_at_k(g.search(xq2, k, efSearch=ef)[1], gt)
    assert r_cpu > 0.9 and r_gpu >= r_cpu - 0.005, (r_cpu, r_gpu)


@pytest.mark.parametrize("frac", [0.5, 0.1, 0.01])
def test_id_selector_bitmap_matches_oracle(small_l2, frac):
    """SearchParametersHNSW.sel (IDSelectorBitmap): filters the results, not the traversal."""
    o = small_l2["oracle"]
    idx = _gpu_from_oracle(o, small_l2["xb"], 16)
    member = np.random.RandomState(int(frac * 1000)).rand(4000) < frac
    bm = np.packbits(member, bitorder="little")
    for ef, k, hb in ((64, 10, 13), (32, 40, 13), (64, 10, 0), (128, 10, 9)):
        Do, Io, So = o.search(small_l2["xq"], k, ef, stats=True, sel_bitmap=bm)
        D, I, S = idx.search(small_l2["xq"], k, efSearch=ef, stats=True, hash_bits=hb, sel_bitmap=bm)
        assert np.array_equal(I, Io) and np.array_equal(D, Do), (ef, k, hb)
        assert member[I[I >= 0]].all()
        if hb == 13:
            assert np.array_equal(S, So)
    with pytest.raises(RuntimeError):
        idx.search(small_l2["xq"], 10, sel_bitmap=bm[:100])          # bitmap too small


def test_rejected_import_keeps_the_old_index(small_l2):
    idx = _gpu_from_oracle(small_l2["oracle"], small_l2["xb"], 16)
    g = small_l2["graph"]
    D0, I0 = idx.search(small_l2["xq"], 10, efSearch=32)
    bad = g["neighbors"].copy()
    bad[5] = 10 ** 6                                   # neighbour id out of range
    with pytest.raises(RuntimeError):
        idx.import_graph(small_l2["xb"], g["levels"], bad, g["entry_point"], g["max_level"])
    with pytest.raises(RuntimeError):
        idx.import_graph(small_l2["xb"], g["levels"], g["neighbors"][:-3], g["entry_point"], g["max_level"])
    assert idx.ntotal == 4000
    D1, I1 = idx.search(small_l2["xq"], 10, efSearch=32)
    assert np.array_equal(I1, I0) and np.array_equal(D1, D0)


@pytest.mark.parametrize("d,team", [(30, 8), (5, 8), (130, 16)])
def test_dimension_not_a_multiple_of_four(oracle_mod, d, team):
    """d % 4 != 0: rows are stored zero-padded to the next 16-byte chunk, which adds exact zeros to
    L2 / IP — so results equal the oracle's on explicitly padded data, bit for bit, and equal the plain
    (unpadded, native-order) oracle within the 1e-4 tolerance. reconstruct returns the caller's d."""
    import hnsw_b200
    dp = (d + 3) // 4 * 4
    xb, xq = synthetic_dataset(d, 1500, 40)
    pad = lambda a: np.ascontiguousarray(np.pad(a, ((0, 0), (0, dp - d))))
    for metric in (1, 0):
        o = oracle_mod.OracleHNSWFlat(dp, 8, metric)
        o.set_team(team)
        o.add(pad(xb))
        g = o.export_graph()
        idx = hnsw_b200.IndexHNSWFlat(d, 8, metric)
        idx.import_graph(xb, g["levels"], g["neighbors"], g["entry_point"], g["max_level"])
        Do, Io, So = o.search(pad(xq), 10, 48, stats=True)
        D, I, S = idx.search(xq, 10, efSearch=48, stats=True, hash_bits=13)
        assert np.array_equal(I, Io) and np.array_equal(D, Do) and np.array_equal(S, So)
        assert np.array_equal(idx.reconstruct(7), xb[7]) and np.array_equal(idx.reconstruct_n(3, 5), xb[3:8])
        # sequential GPU build on the odd-d data == oracle build on the padded data
        b = hnsw_b200.IndexHNSWFlat(d, 8, metric)
        b.set_build_params(max_batch=1)
        b.add(xb)
        assert np.array_equal(b.export_graph()["neighbors"], g["neighbors"])
        nat = oracle_mod.OracleHNSWFlat(d, 8, metric)      # the caller's own d, CPU summation order
        nat.import_graph(xb, g["levels"], g["neighbors"], g["entry_point"], g["max_level"])
        Dn, In = nat.search(xq, 10, 48)
        same = I == In
        assert same.mean() > 0.99 and np.allclose(D[same], Dn[same], rtol=1e-4, atol=1e-6)


def test_misaligned_device_queries_are_restaged(small_l2):
    """search_device with a query pointer that is not 16-byte aligned (a tensor view at a 4-byte offset):
    the rows are re-laid on the stream instead of being handed to the bulk copy (which would fault)."""
    import torch
    idx = _gpu_from_oracle(small_l2["oracle"], small_l2["xb"], 16)
    xq = small_l2["xq"]
    D0, I0 = idx.search(xq, 10, efSearch=48)
    buf = torch.zeros(xq.size + 3, dtype=torch.float32, device="cuda")
    for off in (1, 2, 3):
        view = buf[off:off + xq.size]
        view.copy_(torch.from_numpy(xq).reshape(-1).cuda())
        assert view.data_ptr() % 16 != 0
        D = torch.empty(len(xq), 10, device="cuda")
        I = torch.empty(len(xq), 10, dtype=torch.int64, device="cuda")
        idx.search_device(view.data_ptr(), len(xq), 10, D.data_ptr(), I.data_ptr(), efSearch=48)
        idx.synchronize()
        assert np.array_equal(I.cpu().numpy(), I0) and np.array_equal(D.cpu().numpy(), D0)
    # host path, odd offset into a numpy buffer (pageable): staged copy re-aligns it
    hb = np.zeros(xq.size + 1, np.float32)
    hb[1:] = xq.reshape(-1)
    D1, I1 = idx.search(hb[1:].reshape(xq.shape), 10, efSearch=48)
    assert np.array_equal(I1, I0) and np.array_equal(D1, D0)


def test_pageable_batch_is_chunked_over_lanes(small_l2):
    """A pageable batch larger than one staging chunk goes round-robin over the context's lanes
    (bh_index_search path 2), including more chunks than lanes and a ragged last chunk, with stats."""
    o = small_l2["oracle"]
    idx = _gpu_from_oracle(o, small_l2["xb"], 16)
    rs = np.random.RandomState(11)
    n = 4 * 32768 + 1234                                    # 5 chunks of 32768 (capped), ragged tail
    xq = small_l2["xb"][rs.randint(0, 4000, n)] + 0.01 * rs.randn(n, 32).astype(np.float32)
    o.threads = 8
    Do, Io, So = o.search(xq, 5, 24, stats=True)
    o.threads = 1
    D, I, S = idx.search(xq, 5, efSearch=24, stats=True, hash_bits=12)
    assert np.array_equal(I, Io) and np.array_equal(D, Do) and np.array_equal(S, So)
    D, I = idx.search(xq[:9000], 5, efSearch=24)            # 4 chunks of 2250
    assert np.array_equal(I, Io[:9000]) and np.array_equal(D, Do[:9000])


def test_concurrent_searches_on_one_handle(small_l2):
    """faiss: search is const and thread-safe. Eight threads search one handle at once, each in its own
    search context; every thread must get the single-threaded answer."""
    import threading
    idx = _gpu_from_oracle(small_l2["oracle"], small_l2["xb"], 16)
    xq = small_l2["xq"]
    want = {ef: idx.search(xq, 10, efSearch=ef) for ef in (16, 32, 64, 128)}
    errs = []

    def work(t):
        try:
            for it in range(20):
                ef = (16, 32, 64, 128)[(t + it) % 4]
                D, I = idx.search(xq, 10, efSearch=ef)
                if not (np.array_equal(I, want[ef][1]) and np.array_equal(D, want[ef][0])):
                    errs.append((t, it, ef))
        except Exception as e:  # noqa: BLE001
            errs.append((t, repr(e)))

    ths = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs[:3]


def _host_merge(Ds, Is, offs, metric):
    """Exact merge of per-shard sorted lists: by distance, ties by shard then position (stable)."""
    allD = np.concatenate(Ds, axis=1)
    allI = np.concatenate([np.where(I >= 0, I + o, -1) for I, o in zip(Is, offs)], axis=1)
    order = np.argsort(allD if metric == 1 else -allD, axis=1, kind="stable")[:, :Ds[0].shape[1]]
    return np.take_along_axis(allD, order, 1), np.take_along_axis(allI, order, 1)


@pytest.mark.parametrize("metric", [1, 0])
def test_shards_cabi_two_ranks_emulated_on_one_gpu(metric):
    """bh_shards_* with two ranks living in ONE process on ONE GPU (the same-process branch of connect):
    each rank's traversal kernel stores its packed lists into both gather buffers, the flags are raised,
    and each rank's merge equals the exact host-side merge of the two per-shard results. The two halves
    (post / collect) are called separately and the streams are synchronised in between, so no kernel ever
    spins on a kernel of the same GPU. Also the caller-moves-the-lists variant (publish_to_peers = 0)."""
    import ctypes as C
    import torch
    import hnsw_b200
    from hnsw_b200 import _lib
    L = _lib.lib()
    d, M, k, nq = 32, 16, 10, 300
    xb, xq = synthetic_dataset(d, 5000, nq, normalize=(metric == 0))
    cut = [0, 2600, 5000]
    idxs = []
    for r in range(2):
        ix = hnsw_b200.IndexHNSWFlat(d, M, metric)
        ix.add(xb[cut[r]:cut[r + 1]])
        idxs.append(ix)
    per = [ix.search(xq, k, efSearch=48) for ix in idxs]
    Dw, Iw = _host_merge([p[0] for p in per], [p[1] for p in per], cut[:2], metric)
    hs = []
    for r in range(2):
        s = C.c_void_p()
        _lib.check(L.bh_shards_create(C.byref(s), idxs[r]._h, r, 2, 512, 16))
        hs.append(s)
    try:
        blobs = (C.c_ubyte * (2 * _lib.SHARDS_BLOB_BYTES))()
        for r in range(2):
            _lib.check(L.bh_shards_export(hs[r], C.byref(blobs, r * _lib.SHARDS_BLOB_BYTES)))
        for r in range(2):
            _lib.check(L.bh_shards_connect(hs[r], blobs))
        q = torch.from_numpy(xq).cuda()
        p = _lib.SearchParams(48, 0, 0, 0, None, None, 0, 0, 0)
        for rep in range(3):                       # three epochs: both parities of the gather buffer
            for r in range(2):
                _lib.check(L.bh_shards_post(hs[r], nq, q.data_ptr(), k, C.byref(p), 1))
            for r in range(2):
                idxs[r].synchronize()              # both ranks have published: collect cannot spin
            outs = []
            for r in range(2):
                D = torch.empty(nq, k, device="cuda")
                I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
                _lib.check(L.bh_shards_collect(hs[r], nq, k, D.data_ptr(), I.data_ptr(), 1))
                Dl = torch.empty(nq, k, device="cuda")
                Il = torch.empty(nq, k, dtype=torch.int64, device="cuda")
                _lib.check(L.bh_shards_local_lists(hs[r], nq, k, Dl.data_ptr(), Il.data_ptr()))
                idxs[r].synchronize()
                assert L.bh_shards_status(hs[r]) == 0
                assert np.array_equal(Dl.cpu().numpy(), per[r][0]) and np.array_equal(Il.cpu().numpy(), per[r][1])
                outs.append((D.cpu().numpy(), I.cpu().numpy()))
            for D, I in outs:
                assert np.array_equal(I, Iw) and np.array_equal(D, Dw)
        # pipelined mode: flag + merge kernels on the exchange stream, results complete after join. (Still no
        # kernel waits on a kernel here: everything queued is drained before the merges are launched.)
        for r in range(2):
            _lib.check(L.bh_shards_set_pipelined(hs[r], 1))
        for rep in range(20):                      # > ring depth: slots are reused, flow-control waits are taken
            for r in range(2):
                _lib.check(L.bh_shards_post(hs[r], nq, q.data_ptr(), k, C.byref(p), 1))
            torch.cuda.synchronize()
            outs = []
            for r in range(2):
                D = torch.empty(nq, k, device="cuda")
                I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
                _lib.check(L.bh_shards_collect(hs[r], nq, k, D.data_ptr(), I.data_ptr(), 1))
                outs.append((D, I))
            for r in range(2):
                _lib.check(L.bh_shards_join(hs[r], None))
                idxs[r].synchronize()
                assert L.bh_shards_status(hs[r]) == 0
                assert np.array_equal(outs[r][1].cpu().numpy(), Iw) and np.array_equal(outs[r][0].cpu().numpy(), Dw)
        for r in range(2):
            _lib.check(L.bh_shards_set_pipelined(hs[r], 0))
        # the caller exchanges the lists itself (what the NCCL fallback does with one all-gather)
        for r in range(2):
            _lib.check(L.bh_shards_post(hs[r], nq, q.data_ptr(), k, C.byref(p), 0))
            idxs[r].synchronize()
        views = []
        for r in range(2):
            gp = C.c_void_p()
            _lib.check(L.bh_shards_gather(hs[r], nq, k, C.byref(gp)))
            from hnsw_b200.sharded import _device_view_i64
            views.append(_device_view_i64(gp.value, 2 * nq * k, torch.device("cuda", 0)))
        views[0][nq * k:].copy_(views[1][nq * k:])
        views[1][:nq * k].copy_(views[0][:nq * k])
        torch.cuda.synchronize()
        for r in range(2):
            D = torch.empty(nq, k, device="cuda")
            I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
            _lib.check(L.bh_shards_collect(hs[r], nq, k, D.data_ptr(), I.data_ptr(), 0))
            idxs[r].synchronize()
            assert np.array_equal(I.cpu().numpy(), Iw) and np.array_equal(D.cpu().numpy(), Dw)
    finally:
        for s in hs:
            L.bh_shards_free(s)


def test_faiss_selector_family_equals_the_bitmap_form(small_l2):
    """IDSelectorRange / IDSelectorBatch / IDSelectorNot and SearchParametersHNSW(sel=...): converted on the
    host into IDSelectorBitmap (bh_selector_*), so the result must equal the oracle's run with that bitmap."""
    import hnsw_b200
    o = small_l2["oracle"]
    idx = _gpu_from_oracle(o, small_l2["xb"], 16)
    n = 4000
    rs = np.random.RandomState(3)
    ids = rs.choice(n, 700, replace=False)
    cases = [
        (hnsw_b200.IDSelectorRange(500, 1700), (np.arange(n) >= 500) & (np.arange(n) < 1700)),
        (hnsw_b200.IDSelectorBatch(ids), np.isin(np.arange(n), ids)),
        (hnsw_b200.IDSelectorNot(hnsw_b200.IDSelectorRange(100, 3900)), (np.arange(n) < 100) | (np.arange(n) >= 3900)),
        (hnsw_b200.IDSelectorBitmap(np.packbits(np.arange(n) % 3 == 0, bitorder="little")), np.arange(n) % 3 == 0),
    ]
    for sel, member in cases:
        assert np.array_equal(np.unpackbits(sel.to_bitmap(n), bitorder="little")[:n].astype(bool), member)
        bm = np.packbits(member, bitorder="little")
        Do, Io = o.search(small_l2["xq"], 10, 64, sel_bitmap=bm)
        D, I = idx.search(small_l2["xq"], 10, params=hnsw_b200.SearchParametersHNSW(efSearch=64, sel=sel))
        assert np.array_equal(I, Io) and np.array_equal(D, Do)
        assert member[I[I >= 0]].all()
    # SearchParametersHNSW without a selector: per-call efSearch / check_relative_distance
    o.set_check_relative_distance(False)
    try:
        Do, Io = o.search(small_l2["xq"], 10, 24)
    finally:
        o.set_check_relative_distance(True)
    D, I = idx.search(small_l2["xq"], 10, params=hnsw_b200.SearchParametersHNSW(efSearch=24, check_relative_distance=False))
    assert np.array_equal(I, Io) and np.array_equal(D, Do)
