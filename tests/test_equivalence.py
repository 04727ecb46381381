"""CPU check of the algorithmic claim in DESIGN.md §3.1: faiss's search_from_candidates
(MinimaxHeap with linear pop_min + k-result heap + exact VisitedTable, as restated by the oracle)
returns the same ids and visits the same number of vertices as the kernel's formulation —
ONE sorted ef-list with an expanded bit, merged once per hop, a *forgetful* visited set that is
cleared and re-seeded from the list when it fills, and an optional selector-filtered result list.
The model below is that formulation in plain Python; it shares no code with the oracle."""
import numpy as np
import pytest

from hnsw_b200.datasets import synthetic_dataset


class AssocVisited:
    """The default visited policy of the kernel (beam.cuh assoc_test_and_set): `buckets` FIFO buckets of
    `ways` entries addressed by a multiplicative hash; nothing is ever cleared, old entries fall out."""

    def __init__(self, buckets, ways):
        self.t = [[] for _ in range(buckets)]
        self.bits = int(np.log2(buckets))
        self.ways = ways

    def test_and_set(self, v):
        b = ((v * 2654435761) & 0xFFFFFFFF) >> (32 - self.bits) if self.bits else 0
        t = self.t[b]
        if v in t:
            return False
        if len(t) == self.ways:
            t.pop(0)
        t.append(v)
        return True


def model_search(xb, g, cum, q, k, ef_search, check_rel=True, visited_cap=None, sel=None, dist=None,
                 assoc=None):
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
    if assoc is not None:
        assoc.test_and_set(cur)
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
            if assoc is not None:
                if assoc.test_and_set(int(v)):
                    new.append(int(v))
            elif int(v) not in visited:
                visited.add(int(v))
                new.append(int(v))
        nhops += 1
        nstep += 1
        ndis += len(new)
        scored = [(np.float32(dist(q, xb[v])), v) for v in new]
        thr = (lst[-1][0], lst[-1][1]) if len(lst) == ef else (np.inf, 1 << 62)
        # a re-scored vertex that is still listed arrives with the identical key: the merge drops it
        listed = {(e[0], e[1]) for e in lst}
        acc = [(d, v, False) for d, v in scored if (d, v) < thr and (d, v) not in listed]
        lst = sorted(lst + acc, key=lambda e: (e[0], e[1]))[:ef]
        if sel is not None:
            rthr = res[-1] if len(res) == k else (np.inf, 1 << 62)
            res = sorted(res + [(d, v) for d, v in scored if sel[v] and (d, v) < rthr and (d, v) not in res])[:k]
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
            # set-associative FIFO table, down to one that forgets nearly everything (list members too)
            for buckets, ways in ((1, 2), (4, 4), (64, 8)):
                ids3, ndis3, nhops3 = model_search(xb, g, cum, xq[i], k, ef, crd, None, None, dist,
                                                   assoc=AssocVisited(buckets, ways))
                assert np.array_equal(ids3, Io[i]) and ndis3 >= ndis and nhops3 == nhops
    # selector: filters results, not traversal
    member = rs.rand(n) < 0.2
    bm = np.packbits(member, bitorder="little")
    Do, Io, So = o.search(xq, 10, 32, stats=True, sel_bitmap=bm)
    for i in range(len(xq)):
        ids, ndis, nhops = model_search(xb, g, cum, xq[i], 10, 32, True, 40 + 4 * M, member, dist)
        assert np.array_equal(ids, Io[i]) and nhops == So[i, 1]
        ids, ndis, nhops = model_search(xb, g, cum, xq[i], 10, 32, True, None, member, dist,
                                        assoc=AssocVisited(4, 4))
        assert np.array_equal(ids, Io[i]) and nhops == So[i, 1]


def test_quotient_slots_name_ids_exactly():
    """16-bit visited slots (beam.cuh, kVisitedAssoc16): h = (id * odd) mod 2^(b+16) is a bijection on
    [0, 2^(b+16)), so (bucket = h >> 16, slot = h & 0xFFFF) identifies the id: two different vertices can
    never be mistaken for each other, whatever the table size."""
    for bbits in (2, 6, 8):
        m = bbits + 16
        ids = np.arange(1 << m, dtype=np.uint64)
        h = (ids * np.uint64(2654435761)) & np.uint64((1 << m) - 1)
        assert np.unique(h).size == ids.size
        buckets = (h >> np.uint64(16)).astype(np.int64)
        cnt = np.bincount(buckets, minlength=1 << bbits)
        assert cnt.min() == cnt.max() == 1 << 16        # every bucket serves exactly 2^16 ids


def _incremental_shrink(o, xb, owner, row, nver, src, max_size):
    """DESIGN §3.3 in plain Python: row[:nver] is a verified (self-consistent) prefix; only pairs that
    involve a special candidate (src or row[nver:]) are evaluated."""
    d_to_owner = lambda v: np.float32(o.distance(xb[owner], xb[v]))
    cands = [(d_to_owner(v), int(v)) for v in list(row) + [src]]
    special = set(int(v) for v in row[nver:]) | {int(src)}
    cands.sort()
    kept, evals = [], 0
    for dq, v in cands:
        good = True
        for _, u in kept:
            if v in special or u in special:
                evals += 1
                if np.float32(o.distance(xb[u], xb[v])) < dq:
                    good = False
                    break
        if good:
            kept.append((dq, v))
            if len(kept) >= max_size:
                break
    return [v for _, v in kept], evals


@pytest.mark.parametrize("seed", range(5))
def test_incremental_shrink_equals_full_heuristic(oracle_mod, seed):
    rs = np.random.RandomState(seed)
    d, n, deg = 24, 1500, 12
    ran = with_appended = 0
    xb, _ = synthetic_dataset(d, n, 1, seed=300 + seed)
    o = oracle_mod.OracleHNSWFlat(d, 8)
    o.add(xb[:50])                      # only used for distance() / shrink() on explicit candidate sets
    o2 = oracle_mod.OracleHNSWFlat(d, 8)
    o2.import_graph(xb, np.ones(n, np.int32), np.full(n * 16, -1, np.int32), 0, 0)  # vectors only
    for trial in range(12):
        owner = int(rs.randint(n))
        pool = np.argsort(((xb - xb[owner]) ** 2).sum(1))[1:120]
        # a verified prefix = output of one heuristic run on some candidate set
        first = rs.choice(pool, 40, replace=False).astype(np.int32)
        dq = np.array([o2.distance(xb[owner], xb[v]) for v in first], np.float32)
        core = o2.shrink(first, dq, deg)                       # nearest-first, self-consistent
        room = deg - len(core)
        rest = [v for v in pool if v not in set(core.tolist())]
        appended = rs.choice(rest, room, replace=False).astype(np.int32) if room > 0 else np.zeros(0, np.int32)  # fill the row
        row = np.concatenate([core, appended]).astype(np.int32)
        src = int(rs.choice([v for v in rest if v not in set(appended.tolist())]))
        cand = np.concatenate([row, [src]]).astype(np.int32)
        dqc = np.array([o2.distance(xb[owner], xb[v]) for v in cand], np.float32)
        full = o2.shrink(cand, dqc, deg) if len(cand) >= deg else None
        inc, evals = _incremental_shrink(o2, xb, owner, row, len(core), src, deg)
        assert full is not None and len(cand) == deg + 1
        assert inc == full.tolist(), (seed, trial)
        assert evals <= (len(appended) + 1) * (len(cand))      # far fewer than the ~n^2/2 of a full run
        ran += 1
        with_appended += len(appended) > 0
    assert ran == 12 and with_appended >= 1
