"""A golden that does NOT come from the oracle: a 7-vertex graph whose adjacency rows and two search
traces are worked out BY HAND below from the published faiss algorithm,
then demanded of the CPU oracle (here, no GPU needed) and of the CUDA engine (`-m gpu`).

Setting: points on a line, x-coordinate in component 0 of a 4-d vector, squared L2. M = 2, so a level-0 row
has 2M = 4 slots and a level-1 row has M = 2. efConstruction = 16 (> number of points: every insertion
search sees the whole connected graph). Levels are preset (faiss: fill hnsw.levels before add) and the
points are added ONE PER add() CALL, so neither the level RNG nor the insertion shuffle is involved.

    id   0    1    2    3    4    5    6
    x    0   10    1    2    3   11    4
    lvl  1    0    0    0    0    1    0          (faiss levels[] = lvl + 1)

Notation: d(a,b) = (x_a - x_b)^2. Rows list neighbours in slot order; "-" = -1.
A.9: the kept candidates are linked farthest-first (faiss pops `link_targets`, a max-queue), forward links
first, then the back-links in the same order. A.11: add_link appends while the row has a free slot, else
re-runs the A.10 heuristic on row + newcomer and rewrites the row farthest-first.
A.10 (heuristic, only when #candidates >= row size): nearest first, keep v iff no kept u has d(u,v) < d(v,base).

insert 0: first vertex -> entry_point = 0, max_level = 1, no links.
insert 1 (x=10): search from 0 finds {0:100}. 1 < 4 candidates: keep all. L0[1]=[0]; back: L0[0]=[1].
insert 2 (x=1):  from 0 (d=1): row[0]=[1] -> 1:81. candidates {0:1, 1:81}; 2 < 4: keep all.
                 farthest first: L0[2]=[1,0]; back-links 1<-2, 0<-2: L0[1]=[0,2], L0[0]=[1,2].
insert 3 (x=2):  from 0 (d=4): row[0]=[1,2] -> 1:64, 2:1. {2:1, 0:4, 1:64}; 3 < 4: keep all.
                 L0[3]=[1,0,2]; back: L0[1]=[0,2,3], L0[0]=[1,2,3], L0[2]=[1,0,3].
insert 4 (x=3):  from 0 (d=9): row[0]=[1,2,3] -> 1:49, 2:4, 3:1. 4 candidates >= 4: heuristic, nearest first:
                   3 (1): kept.   2 (4): d(2,3)=1 < 4 rejected.   0 (9): d(0,3)=4 < 9 rejected.
                   1 (49): d(1,3)=64 >= 49 kept.
                 kept {3,1}, farthest first: L0[4]=[1,3]; back: L0[1]=[0,2,3,4] (full), L0[3]=[1,0,2,4] (full).
insert 5 (x=11, level 1 = max_level, entry stays 0): started from (0, d=121) at BOTH levels (A.8).
   level 1: row1[0] empty -> {0:121}; 1 < 2: keep. L1[5]=[0]; back: L1[0]=[5].
   level 0: row[0]=[1,2,3] -> 1:1, 2:100, 3:81; expand 1: row[1]=[0,2,3,4] -> 4:64; expanding 4, 3, 2 finds nothing new.
            5 candidates >= 4: heuristic: 1 (1) kept; 4 (64): d(4,1)=49 < 64 rej; 3 (81): d(3,1)=64 < 81 rej;
            2 (100): d(2,1)=81 < 100 rej; 0 (121): d(0,1)=100 < 121 rej.  L0[5]=[1].
            back-link 1<-5: row[1]=[0,2,3,4] is full -> heuristic on {5:1, 4:49, 3:64, 2:81, 0:100} (distances to 1):
              5 kept; 4: d(4,5)=64 >= 49 kept; 3: d(3,4)=1 < 64 rej; 2: d(2,4)=4 < 81 rej; 0: d(0,4)=9 < 100 rej.
            rewritten farthest first: L0[1]=[4,5,-,-].
insert 6 (x=4):  level 1: row1[0]=[5]: d(6,5)=49 is not < d(6,0)=16 -> stay at 0.
   level 0 from (0,16): row[0]=[1,2,3] -> 1:36, 2:9, 3:4; expand 3: row[3]=[1,0,2,4] -> 4:1; expand 4: nothing;
            expand 2: nothing; expand 1 (36 is not > worst 36): row[1]=[4,5] -> 5:49; expand 5 (49 not > 49): nothing.
            6 candidates: heuristic: 4 (1) kept; 3 (4): d(3,4)=1 < 4 rej; 2 (9): d(2,4)=4 < 9 rej; 0 (16): d(0,4)=9 < 16 rej;
            1 (36): d(1,4)=49 >= 36 kept; 5 (49): d(5,4)=64 >= 49 but d(5,1)=1 < 49 rej.
            L0[6]=[1,4]; back: L0[1]=[4,5,6], L0[4]=[1,3,6].

Final rows:  0: L0 [1,2,3,-] L1 [5,-]     1: [4,5,6,-]     2: [1,0,3,-]     3: [1,0,2,4]
             4: [1,3,6,-]                 5: L0 [1,-,-,-] L1 [0,-]          6: [1,4,-,-]

search A: q = 3.4, k = 3, efSearch = 4 (A.5/A.6). d(q,0)=11.56. level 1: d(q,5)=57.76 not closer -> stay at 0.
   buffer (capacity 4; * = already expanded, it keeps its slot and distance), result heap (k=3):
   expand 0: row [1,2,3] -> 1:43.56, 2:5.76, 3:1.96.                buffer {0*,1,2,3}          results {0,2,3}
   expand 3 (1.96): row [1,0,2,4] -> 4:0.16; buffer full, evict worst (1:43.56). {0*,2,3*,4}    results {2,3,4}
   expand 4 (0.16): row [1,3,6] -> 6:0.36; evict worst (0*:11.56).  buffer {2,3*,4*,6}         results {3,6,4}
   expand 6 (0.36): row [1,4] all visited.   expand 2 (5.76; 3 entries are closer, 3 < efSearch): all visited.
   nothing unexpanded left.  5 hops (0,3,4,6,2), 5 distance evaluations at level 0 (1,2,3,4,6).
   result, ascending: ids [4, 6, 3], distances [0.16, 0.36, 1.96].

search B: q = 10.6, k = 2, efSearch = 2. d(q,0)=112.36. level 1: row1[0]=[5], d(q,5)=0.16 < 112.36 -> move to 5;
   row1[5]=[0] is not closer -> level 0 starts from 5.
   expand 5: row [1] -> 1:0.36.  buffer {5*,1}  results {5,1}
   expand 1 (one closer entry, 1 < efSearch): row [4,5,6] -> 4:57.76 and 6:43.56 are scored, neither beats the
   result heap's worst (0.36) nor the full buffer's worst -> dropped.
   2 hops, 3 distance evaluations.  result ids [5, 1], distances [0.16, 0.36].
"""
import numpy as np
import pytest

X = [0.0, 10.0, 1.0, 2.0, 3.0, 11.0, 4.0]
LEVELS = [2, 1, 1, 1, 1, 2, 1]            # faiss hnsw.levels = level + 1
M, EFC = 2, 16
# faiss `neighbors` array: per vertex its level-0 row (4 slots) then, for vertices 0 and 5, the level-1 row (2)
EXPECTED_NEIGHBORS = [
    1, 2, 3, -1,   5, -1,      # vertex 0
    4, 5, 6, -1,               # vertex 1
    1, 0, 3, -1,               # vertex 2
    1, 0, 2, 4,                # vertex 3
    1, 3, 6, -1,               # vertex 4
    1, -1, -1, -1,   0, -1,    # vertex 5
    1, 4, -1, -1,              # vertex 6
]
EXPECTED_OFFSETS = [0, 6, 10, 14, 18, 22, 28, 32]
QUERIES = [  # (x, k, efSearch, ids, level-0 ndis, level-0 nhops)
    (3.4, 3, 4, [4, 6, 3], 5, 5),
    (10.6, 2, 2, [5, 1], 3, 2),
]


def _vecs(xs):
    v = np.zeros((len(xs), 4), np.float32)
    v[:, 0] = xs
    return v


def _check_graph(g):
    assert g["levels"].tolist() == LEVELS
    assert g["offsets"].tolist() == EXPECTED_OFFSETS
    assert g["neighbors"].tolist() == EXPECTED_NEIGHBORS
    assert g["entry_point"] == 0 and g["max_level"] == 1


def _check_search(search):
    for x, k, ef, ids, ndis, nhops in QUERIES:
        q = _vecs([x])
        D, I, S = search(q, k, ef)
        assert I[0].tolist() == ids
        want = ((np.float32(x) - np.asarray([X[i] for i in ids], np.float32)) ** 2).astype(np.float32)
        assert np.allclose(D[0], want, rtol=1e-6)
        assert (int(S[0, 0]), int(S[0, 1])) == (ndis, nhops)


def test_oracle_reproduces_the_hand_worked_graph_and_traces(oracle_mod):
    o = oracle_mod.OracleHNSWFlat(4, M)
    o.efConstruction = EFC
    xb = _vecs(X)
    for i in range(len(X)):
        o.add_with_levels(xb[i:i + 1], [LEVELS[i]])
    _check_graph(o.export_graph())
    _check_search(lambda q, k, ef: o.search(q, k, ef, stats=True))


@pytest.mark.gpu
def test_gpu_reproduces_the_hand_worked_graph_and_traces():
    import hnsw_b200
    idx = hnsw_b200.IndexHNSWFlat(4, M)
    idx.hnsw.efConstruction = EFC
    xb = _vecs(X)
    for i in range(len(X)):
        idx.add(xb[i:i + 1], levels=[LEVELS[i]])
    _check_graph(idx.export_graph())
    _check_search(lambda q, k, ef: idx.search(q, k, efSearch=ef, stats=True, hash_bits=10))
    _check_search(lambda q, k, ef: idx.search(q, k, efSearch=ef, stats=True))           # default visited table
