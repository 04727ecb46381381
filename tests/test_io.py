"""IHNf (faiss IndexHNSWFlat file layout, as recalled — unverified against faiss) round trips."""
import numpy as np
import pytest

from hnsw_b200 import io as bio


def test_dump_load_roundtrip_cpu(tmp_path, small_l2):
    g = small_l2["graph"]
    p = str(tmp_path / "a.index")
    bio.dump(p, d=32, M=16, metric=1, x=small_l2["xb"], levels=g["levels"], offsets=g["offsets"],
             neighbors=g["neighbors"], entry_point=g["entry_point"], max_level=g["max_level"],
             efConstruction=40, efSearch=33)
    s = bio.load(p)
    assert (s["d"], s["M"], s["metric"], s["efSearch"], s["efConstruction"]) == (32, 16, 1, 33, 40)
    assert np.array_equal(s["x"], small_l2["xb"]) and np.array_equal(s["levels"], g["levels"])
    assert np.array_equal(s["offsets"], g["offsets"]) and np.array_equal(s["neighbors"], g["neighbors"])
    assert s["entry_point"] == g["entry_point"] and s["max_level"] == g["max_level"]
    raw = open(p, "rb").read()
    assert raw[:4] == b"IHNf" and b"IxF2" in raw
    with pytest.raises(ValueError):
        open(p, "wb").write(b"IxF2" + raw[4:])
        bio.load(p)


def test_level_tables_match_oracle(oracle_mod):
    for M in (4, 16, 32):
        p, c = bio._level_tables(M)
        po, co = oracle_mod.OracleHNSWFlat(8, M).tables()
        assert np.array_equal(c, co) and np.allclose(p, po, rtol=0, atol=0)


@pytest.mark.gpu
def test_write_read_index_gpu(tmp_path, small_l2):
    import hnsw_b200
    idx = hnsw_b200.IndexHNSWFlat(32, 16)
    idx.add(small_l2["xb"])
    idx.hnsw.efSearch = 48
    p = str(tmp_path / "b.index")
    bio.write_index(idx, p)
    idx2 = bio.read_index(p)
    assert idx2.ntotal == idx.ntotal and idx2.hnsw.efSearch == 48
    D1, I1 = idx.search(small_l2["xq"], 10)
    D2, I2 = idx2.search(small_l2["xq"], 10)
    assert np.array_equal(I1, I2) and np.array_equal(D1, D2)
