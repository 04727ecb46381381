#!/usr/bin/env python
"""CPU replay of the traversal kernel's formulation (sorted ef-list + per-hop merge) under
different VISITED-SET policies, to size the shared-memory table without GPU time.

Every policy is "forgetful but safe": a forgotten vertex may be scored again (ndis grows) but the
result ids must equal the exact-table run — the script asserts that for every query.

  exact                       faiss VisitedTable (never forgets)
  reset:SLOTS                 round-1 kernel: open addressing, cleared and re-seeded from the list
                              when 3/4 full
  assoc:BUCKETS:WAYS          set-associative, per-bucket FIFO eviction, no global reset; a re-scored
                              vertex that is still in the list is dropped by the merge (same key)

usage: visited_policy_sim.py graph.npz xb.npy [nq] [ef ...]
(graph.npz = oracle export_graph() + xq; see scripts/README.md)
"""
import sys

import numpy as np


class Exact:
    name = "exact"

    def start(self, ids):
        self.s = set(ids)

    def before_hop(self, deg, list_ids):
        pass

    def test_and_set(self, v):
        if v in self.s:
            return False
        self.s.add(v)
        return True


class Reset:
    def __init__(self, slots):
        self.slots, self.name = slots, f"reset:{slots}"
        self.limit = (3 * slots) // 4

    def start(self, ids):
        self.s = set(ids)

    def before_hop(self, deg, list_ids):
        if len(self.s) + deg > self.limit:
            self.s = set(list_ids)

    def test_and_set(self, v):
        if v in self.s:
            return False
        self.s.add(v)
        return True


class Assoc:
    def __init__(self, buckets, ways):
        self.nb, self.ways, self.name = buckets, ways, f"assoc:{buckets}x{ways}"
        self.bits = int(np.log2(buckets))
        assert 1 << self.bits == buckets

    def start(self, ids):
        self.tab = [[] for _ in range(self.nb)]
        for v in ids:
            self.test_and_set(v)

    def before_hop(self, deg, list_ids):
        pass

    def test_and_set(self, v):
        b = ((v * 2654435761) & 0xFFFFFFFF) >> (32 - self.bits)
        t = self.tab[b]
        if v in t:
            return False
        if len(t) == self.ways:
            t.pop(0)  # FIFO
        t.append(v)
        return True


def search(xb, nbr0, q, start, dstart, ef, pol):
    lst = [(dstart, start, False)]
    pol.start([start])
    ndis = nhops = wasted = 0
    while True:
        pos = next((i for i, e in enumerate(lst) if not e[2]), -1)
        if pos < 0 or pos >= ef:
            break
        d0, v0, _ = lst[pos]
        lst[pos] = (d0, v0, True)
        row = nbr0[v0]
        row = row[row >= 0]
        pol.before_hop(len(nbr0[v0]), [e[1] for e in lst])
        new = [int(v) for v in row if pol.test_and_set(int(v))]
        nhops += 1
        ndis += len(new)
        if not new:
            continue
        diff = xb[new] - q
        ds = np.einsum("ij,ij->i", diff, diff)
        thr = (lst[-1][0], lst[-1][1]) if len(lst) == ef else (np.inf, 1 << 62)
        have = {e[1] for e in lst}
        acc = []
        for d, v in zip(ds.tolist(), new):
            if (d, v) < thr:
                if v in have:   # merge finds the identical key already in the list: dropped
                    wasted += 1
                    continue
                acc.append((d, v, False))
        lst = sorted(lst + acc, key=lambda e: (e[0], e[1]))[:ef]
    return [v for _, v, _ in lst[:10]], ndis, nhops, wasted


def main():
    g = np.load(sys.argv[1])
    xb = np.load(sys.argv[2], mmap_mode="r")
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    efs = [int(x) for x in sys.argv[4:]] or [64, 128, 256]
    levels, offsets, nb = g["levels"], g["offsets"].astype(np.int64), g["neighbors"]
    M = 32
    n = levels.shape[0]
    # level-0 rows as a dense matrix (the engine's nbr0 layout)
    nbr0 = np.stack([nb[offsets[:-1] + j] for j in range(2 * M)], axis=1)
    xq = g["xq"][:nq]
    cum = [0, 2 * M] + [2 * M + M * i for i in range(1, 8)]

    def descend(q):
        cur = int(g["entry_point"])
        dcur = float(((xb[cur] - q) ** 2).sum())
        for level in range(int(g["max_level"]), 0, -1):
            while True:
                r = nb[offsets[cur] + cum[level]: offsets[cur] + cum[level + 1]]
                r = r[r >= 0]
                if len(r) == 0:
                    break
                diff = xb[r] - q
                ds = np.einsum("ij,ij->i", diff, diff)
                j = int(np.argmin(ds))
                if ds[j] < dcur:
                    cur, dcur = int(r[j]), float(ds[j])
                else:
                    break
        return cur, dcur

    starts = [descend(q) for q in xq]
    print(f"n={n} nq={nq}")
    for ef in efs:
        pols = [Exact(), Reset(512), Reset(1024), Reset(2048), Assoc(128, 4), Assoc(256, 4), Assoc(128, 8),
                Assoc(256, 8), Assoc(512, 8)]
        base = None
        for pol in pols:
            tot = 0
            was = 0
            ids_all = []
            for q, (s, ds) in zip(xq, starts):
                ids, ndis, nhops, wasted = search(xb, nbr0, q, s, ds, ef, pol)
                tot += ndis
                was += wasted
                ids_all.append(ids)
            if base is None:
                base, base_ids = tot, ids_all
            assert ids_all == base_ids, f"{pol.name}: result ids changed"
            print(f"ef={ef:4d} {pol.name:14s} ndis/query {tot / nq:9.1f}  x{tot / base:.3f}  "
                  f"(re-scored list members {was / nq:.1f})", flush=True)


if __name__ == "__main__":
    main()
