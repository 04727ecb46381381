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


def test_hand_packed_file_bytes(tmp_path):
    """A 3-vertex IndexHNSWFlat(d=2, M=2) file spelled out byte by byte from the published faiss layout
    (write_index_header, write_HNSW, IndexFlat's XB vector), independently of io.dump: the reader must
    parse it and the writer must reproduce it exactly."""
    import struct
    M, d, n = 2, 2, 3
    probas, cum = bio._level_tables(M)           # checked against the oracle's tables above
    # faiss stores levels as "number of levels" (top level + 1); vertex 1 reaches level 1
    levels = [1, 2, 1]
    # per-vertex neighbour storage: cum[levels[i]] slots each -> 4, 6, 4
    offsets = [0, 4, 10, 14]
    nb = [1, 2, -1, -1,   0, 2, -1, -1, -1, -1,   0, 1, -1, -1]
    x = [0.0, 0.0, 1.0, 0.0, 0.0, 2.0]
    hdr = struct.pack("<i", d) + struct.pack("<q", n) + struct.pack("<qq", 1 << 20, 1 << 20) \
        + struct.pack("<B", 1) + struct.pack("<i", 1)
    raw = b"IHNf" + hdr
    raw += struct.pack("<Q", len(probas)) + b"".join(struct.pack("<d", p) for p in probas)
    raw += struct.pack("<Q", len(cum)) + b"".join(struct.pack("<i", int(c)) for c in cum)
    raw += struct.pack("<Q", n) + b"".join(struct.pack("<i", v) for v in levels)
    raw += struct.pack("<Q", n + 1) + b"".join(struct.pack("<Q", v) for v in offsets)
    raw += struct.pack("<Q", len(nb)) + b"".join(struct.pack("<i", v) for v in nb)
    raw += struct.pack("<i", 1) + struct.pack("<i", 1)          # entry_point, max_level
    raw += struct.pack("<i", 40) + struct.pack("<i", 16) + struct.pack("<i", 1)   # efC, efS, upper_beam
    raw += b"IxF2" + hdr + struct.pack("<Q", n * d) + b"".join(struct.pack("<f", v) for v in x)
    assert int(cum[1]) == 2 * M and int(cum[2]) == 3 * M
    p = str(tmp_path / "hand.index")
    open(p, "wb").write(raw)
    s = bio.load(p)
    assert (s["d"], s["M"], s["metric"], s["entry_point"], s["max_level"]) == (2, 2, 1, 1, 1)
    assert s["levels"].tolist() == levels and s["offsets"].tolist() == offsets and s["neighbors"].tolist() == nb
    assert s["x"].tolist() == [[0.0, 0.0], [1.0, 0.0], [0.0, 2.0]]
    p2 = str(tmp_path / "hand2.index")
    bio.dump(p2, d=d, M=M, metric=1, x=np.array(x, np.float32).reshape(n, d), levels=levels, offsets=offsets,
             neighbors=nb, entry_point=1, max_level=1, efConstruction=40, efSearch=16)
    assert open(p2, "rb").read() == raw
    # inner-product files carry metric 0 and the IxFI storage fourcc
    bio.dump(p2, d=d, M=M, metric=0, x=np.array(x, np.float32).reshape(n, d), levels=levels, offsets=offsets,
             neighbors=nb, entry_point=1, max_level=1)
    r = open(p2, "rb").read()
    assert b"IxFI" in r and bio.load(p2)["metric"] == 0
    # truncated file is an error, not garbage
    open(p2, "wb").write(raw[:-5])
    with pytest.raises(ValueError):
        bio.load(p2)
