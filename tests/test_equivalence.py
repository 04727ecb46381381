"""CPU check of the algorithmic claim in DESIGN.md §3.1: faiss's search_from_candidates
(MinimaxHeap with linear pop_min + k-result heap + exact VisitedTable, as restated by the oracle)
returns the same ids and visits the same number of vertices as the kernel's formulation —
ONE sorted ef-list with an expanded bit, merged once per hop, a *forgetful* visited set that is
cleared and re-seeded from the list when it fills, and an optional selector-filtered result list.
The model below is that formulation in plain Python; it shares no code with the oracle."""
import numpy as np
import pytest

from hnsw_b200.datasets import synthetic_dataset


def model_search(xb, g, cum, q, k, ef_search, check_rel=True, visited_cap=None, sel=None, dist=None):
    levels, offsets, nb = g["levels"], g["offsets"].astype(np.int64), g["neighbors"]
    row = lambda v, l: nb[offsets[v] + cum[l]: offsets[v] + cum[l + 1]]
    # greedy descent (argmin with first-index ties == sequential strict-< scan)
    cur = g["entry_point"]
    dcur = dist(q, xb[cur])
    for level in range(g["max_level"], 0, -1):
        while True:
            r = row(cur, level)
            r = r[: np.argmax(r < 0)] if (r < 0).any() else r
            if len(r) == 0:
                break
            ds = np.array([dist(q, xb[v]) for v in r], np.float32)
            j = int(np.argmin(ds))
            if ds[j] < dcur:
                cur, dcur = int(r[j]), ds[j]
            else:
                break
    ef = max(ef_search, k)
    lst = [(dcur, cur, False)]          # (dist, id, expanded), kept sorted by (dist, id)
    res = [(dcur, cur)] if (sel is None or sel[cur]) else []
    visited = {cur}
    ndis = nhops = nstep = 0
    while True:
        pos = next((i for i, e in enumerate(lst) if not e[2]), -1)
        if pos < 0 or (check_rel and pos >= ef_search) or (not check_rel and nstep > ef_search):
            break
        d0, v0, _ = lst[pos]
        lst[pos] = (d0, v0, True)
        r = row(v0, 0)
        r = r[: np.argmax(r < 0)] if (r < 0).any() else r
        if visited_cap is not None and len(visited) + len(r) > visited_cap:   # forget and re-seed
            visited = {e[1] for e in lst} | {e[1] for e in res}
        new = []
        for v in r:
            if int(v) not in visited:
                visited.add(int(v))
                new.append(int(v))
        nhops += 1
        nstep += 1
        ndis += len(new)
        scored = [(np.float32(dist(q, xb[v])), v) for v in new]
        thr = (lst[-1][0], lst[-1][1]) if len(lst) == ef else (np.inf, 1 << 62)
        acc = [(d, v, False) for d, v in scored if (d, v) < thr]
        lst = sorted(lst + acc, key=lambda e: (e[0], e[1]))[:ef]
        if sel is not None:
            rthr = res[-1] if len(res) == k else (np.inf, 1 << 62)
            res = sorted(res + [(d, v) for d, v in scored if sel[v] and (d, v) < rthr])[:k]
    out = res if sel is not None else [(d, v) for d, v, _ in lst]
    ids = [v for _, v in out[:k]] + [-1] * (k - min(k, len(out)))
    return np.array(ids, np.int64), ndis, nhops


@pytest.mark.parametrize("seed", range(6))
def test_sorted_list_formulation_equals_faiss_heaps(oracle_mod, seed):
    rs = np.random.RandomState(seed)
    d, M = 16, int(rs.choice([4, 8, 16]))
    n = int(rs.choice([300, 900, 2000]))
    xb, xq = synthetic_dataset(d, n, 25, seed=100 + seed)
    o = oracle_mod.OracleHNSWFlat(d, M)
    o.efConstruction = 24
    o.add(xb)
    g = o.export_graph()
    _, cum = o.tables()
    dist = lambda a, b: np.float32(o.distance(a, b))
    for k, ef, crd in ((10, 32, True), (5, 8, True), (20, 6, True), (10, 12, False)):
        o.set_check_relative_distance(crd)
        try:
            Do, Io, So = o.search(xq, k, ef, stats=True)
        finally:
            o.set_check_relative_distance(True)
        for i in range(len(xq)):
            ids, ndis, nhops = model_search(xb, g, cum, xq[i], k, ef, crd, None, None, dist)
            assert np.array_equal(ids, Io[i]), (seed, k, ef, crd, i)
            assert (ndis, nhops) == (So[i, 0], So[i, 1])
            # forgetful visited set: identical ids, never fewer distance evaluations
            ids2, ndis2, _ = model_search(xb, g, cum, xq[i], k, ef, crd, max(ef, k) + 4 * M, None, dist)
            assert np.array_equal(ids2, Io[i]) and ndis2 >= ndis
    # selector: filters results, not traversal
    member = rs.rand(n) < 0.2
    bm = np.packbits(member, bitorder="little")
    Do, Io, So = o.search(xq, 10, 32, stats=True, sel_bitmap=bm)
    for i in range(len(xq)):
        ids, ndis, nhops = model_search(xb, g, cum, xq[i], 10, 32, True, 40 + 4 * M, member, dist)
        assert np.array_equal(ids, Io[i]) and nhops == So[i, 1]
