"""faiss index-file exchange for IndexHNSWFlat ("IHNf") — SURVEY.md §8(f) rank 1.

Layout written/read (faiss `impl/index_write.cpp` / `index_read.cpp`, AS RECALLED in SURVEY.md —
UNVERIFIED against a real faiss file, because neither faiss nor any faiss-written file exists in
this environment; verify against one before trusting it for interchange):

    fourcc "IHNf"
    header : d int32 | ntotal int64 | dummy int64 (1<<20) | dummy int64 (1<<20) |
             is_trained uint8 | metric_type int32 (| metric_arg float32 if metric_type > 1)
    HNSW   : vector<double> assign_probas | vector<int32> cum_nneighbor_per_level |
             vector<int32> levels | vector<uint64> offsets | vector<int32> neighbors |
             entry_point int32 | max_level int32 | efConstruction int32 | efSearch int32 | upper_beam int32
             (every vector = uint64 count followed by the raw elements)
    storage: fourcc "IxF2" (L2) or "IxFI" (inner product) | the same header | uint64 count of floats |
             float32[ntotal * d]

`dump` / `load` work on plain numpy state (CPU-testable); `write_index` / `read_index` wrap them
around a hnsw_b200.IndexHNSWFlat.
"""
from __future__ import annotations

import struct

import numpy as np

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1


def _fourcc(s: str) -> int:
    b = s.encode()
    return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24)


def _level_tables(M: int):
    """assign_probas / cum_nneighbor_per_level of HNSW::set_default_probas(M, 1/ln M)."""
    mult = np.float32(1.0 / np.log(M))
    probas, cum, nn, level = [], [0], 0, 0
    while True:
        p = np.float32(np.exp(-level / np.float64(mult)) * (1 - np.exp(-1 / np.float64(mult))))
        if p < 1e-9:
            break
        probas.append(float(p))
        nn += 2 * M if level == 0 else M
        cum.append(nn)
        level += 1
    return np.array(probas, np.float64), np.array(cum, np.int32)


def _wvec(f, arr, dtype):
    a = np.ascontiguousarray(arr, dtype)
    f.write(struct.pack("<Q", a.size))
    f.write(a.tobytes())


def _rvec(f, dtype):
    (n,) = struct.unpack("<Q", f.read(8))
    a = np.frombuffer(f.read(n * np.dtype(dtype).itemsize), dtype=dtype, count=n)
    if a.size != n:
        raise ValueError("truncated index file")
    return a.copy()


def _wheader(f, d, ntotal, metric):
    f.write(struct.pack("<iqqqBi", d, ntotal, 1 << 20, 1 << 20, 1, metric))


def _rheader(f):
    d, ntotal, _, _, trained, metric = struct.unpack("<iqqqBi", f.read(4 + 8 * 3 + 1 + 4))
    if metric > 1:
        f.read(4)  # metric_arg
    return d, ntotal, metric


def dump(path, *, d, M, metric, x, levels, offsets, neighbors, entry_point, max_level,
         efConstruction=40, efSearch=16):
    x = np.ascontiguousarray(x, np.float32)
    n = x.shape[0]
    probas, cum = _level_tables(M)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", _fourcc("IHNf")))
        _wheader(f, d, n, metric)
        _wvec(f, probas, np.float64)
        _wvec(f, cum, np.int32)
        _wvec(f, levels, np.int32)
        _wvec(f, offsets, np.uint64)
        _wvec(f, neighbors, np.int32)
        f.write(struct.pack("<iiiii", entry_point, max_level, efConstruction, efSearch, 1))
        f.write(struct.pack("<I", _fourcc("IxF2" if metric == METRIC_L2 else "IxFI")))
        _wheader(f, d, n, metric)
        _wvec(f, x.reshape(-1), np.float32)


def load(path):
    with open(path, "rb") as f:
        (cc,) = struct.unpack("<I", f.read(4))
        if cc != _fourcc("IHNf"):
            raise ValueError("not an IndexHNSWFlat file (fourcc %08x)" % cc)
        d, n, metric = _rheader(f)
        probas = _rvec(f, np.float64)
        cum = _rvec(f, np.int32)
        levels = _rvec(f, np.int32)
        offsets = _rvec(f, np.uint64)
        neighbors = _rvec(f, np.int32)
        entry_point, max_level, efc, efs, upper_beam = struct.unpack("<iiiii", f.read(20))
        (cc2,) = struct.unpack("<I", f.read(4))
        if cc2 not in (_fourcc("IxF2"), _fourcc("IxFI"), _fourcc("IxFl")):
            raise ValueError("storage is not an IndexFlat (fourcc %08x)" % cc2)
        d2, n2, _ = _rheader(f)
        x = _rvec(f, np.float32)
        if d2 != d or n2 != n or x.size != n * d:
            raise ValueError("storage does not match the graph")
    if len(cum) < 2 or cum[1] % 2:
        raise ValueError("bad cum_nneighbor_per_level")
    M = int(cum[1]) // 2
    if upper_beam != 1:
        raise ValueError("upper_beam != 1 is not supported")
    return dict(d=d, M=M, metric=metric, x=x.reshape(n, d), levels=levels, offsets=offsets,
                neighbors=neighbors, entry_point=entry_point, max_level=max_level,
                efConstruction=efc, efSearch=efs, assign_probas=probas)


def write_index(index, path):
    """faiss.write_index(index, path) for a hnsw_b200.IndexHNSWFlat."""
    g = index.export_graph()
    x = index.reconstruct_n()
    dump(path, d=index.d, M=index.M, metric=index.metric_type, x=x, levels=g["levels"], offsets=g["offsets"],
         neighbors=g["neighbors"], entry_point=g["entry_point"], max_level=g["max_level"],
         efConstruction=index.hnsw.efConstruction, efSearch=index.hnsw.efSearch)


def read_index(path, device: int = 0):
    """faiss.read_index(path) -> hnsw_b200.IndexHNSWFlat on `device`."""
    from .index import IndexHNSWFlat
    s = load(path)
    idx = IndexHNSWFlat(s["d"], s["M"], s["metric"], device=device)
    idx.hnsw.efConstruction = s["efConstruction"]
    idx.hnsw.efSearch = s["efSearch"]
    if len(s["levels"]):
        idx.import_graph(s["x"], s["levels"], s["neighbors"], s["entry_point"], s["max_level"])
    return idx
