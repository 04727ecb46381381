"""Parity against REAL faiss — runs only where `import faiss` works (it does not in the build
container nor on the GPU boxes of this round, so these tests are normally skipped; they are the
north_star checks to run the day a faiss wheel is present: SURVEY.md §8c "run-time upgrade path")."""
import numpy as np
import pytest

faiss = pytest.importorskip("faiss")

from hnsw_b200.datasets import synthetic_dataset  # noqa: E402


def _faiss_graph(index):
    h = index.hnsw
    return dict(levels=faiss.vector_to_array(h.levels).astype(np.int32),
                offsets=faiss.vector_to_array(h.offsets).astype(np.uint64),
                neighbors=faiss.vector_to_array(h.neighbors).astype(np.int32),
                entry_point=int(h.entry_point), max_level=int(h.max_level))


def test_oracle_matches_faiss_single_thread(oracle_mod):
    """Pins the oracle: same levels, same sequential graph, same search results as faiss."""
    xb, xq = synthetic_dataset(32, 3000, 64)
    faiss.omp_set_num_threads(1)
    fi = faiss.IndexHNSWFlat(32, 16)
    fi.add(xb)
    fi.hnsw.efSearch = 48
    Df, If = fi.search(xq, 10)
    o = oracle_mod.OracleHNSWFlat(32, 16)
    o.add(xb)
    g, gf = o.export_graph(), _faiss_graph(fi)
    assert np.array_equal(g["levels"], gf["levels"])
    assert (g["neighbors"] != gf["neighbors"]).mean() < 0.01      # SIMD summation-order ties only
    D, I = o.search(xq, 10, 48)
    assert (I == If).mean() > 0.99 and np.allclose(D, Df, rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
def test_gpu_search_on_faiss_graph_identical_ids():
    """north_star: on a graph imported from faiss with identical entry point, the result ID lists
    must be identical except for distance ties; distances within 1e-4 relative."""
    import hnsw_b200
    xb, xq = synthetic_dataset(128, 50000, 1000)
    fi = faiss.IndexHNSWFlat(128, 32)
    fi.hnsw.efConstruction = 100
    fi.add(xb)
    gf = _faiss_graph(fi)
    gi = hnsw_b200.IndexHNSWFlat(128, 32)
    gi.import_graph(xb, gf["levels"], gf["neighbors"], gf["entry_point"], gf["max_level"])
    for ef in (16, 64, 256):
        fi.hnsw.efSearch = ef
        Df, If = fi.search(xq, 10)
        D, I = gi.search(xq, 10, efSearch=ef)
        same = I == If
        assert same.mean() > 0.995
        assert np.allclose(D[same], Df[same], rtol=1e-4, atol=1e-6)
