"""faiss IDSelector family → bitmap conversion (bh_selector_* are pure host functions: no GPU needed)."""
import numpy as np
import pytest

import hnsw_b200


def _bits(bm, n):
    return np.unpackbits(bm, bitorder="little")[:n].astype(bool)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 64, 1000, 4097])
def test_range_batch_not_bitmaps(n):
    rs = np.random.RandomState(n)
    ar = np.arange(n)
    for imin, imax in ((0, n), (3, 3), (-5, n // 2), (n // 3, 10 * n), (n, n + 4)):
        bm = hnsw_b200.IDSelectorRange(imin, imax).to_bitmap(n)
        assert bm.size == (n + 7) // 8
        assert np.array_equal(_bits(bm, n), (ar >= imin) & (ar < imax))
        assert not np.unpackbits(bm, bitorder="little")[n:].any()          # no member beyond ntotal
    ids = rs.randint(-3, n + 5, size=max(1, n // 2))
    bm = hnsw_b200.IDSelectorBatch(ids).to_bitmap(n)
    assert np.array_equal(_bits(bm, n), np.isin(ar, ids))
    assert hnsw_b200.IDSelectorArray is hnsw_b200.IDSelectorBatch
    nb = hnsw_b200.IDSelectorNot(hnsw_b200.IDSelectorBatch(ids)).to_bitmap(n)
    assert np.array_equal(_bits(nb, n), ~np.isin(ar, ids))
    assert not np.unpackbits(nb, bitorder="little")[n:].any()              # NOT never selects ids >= ntotal
    member = rs.rand(n) < 0.3
    short = hnsw_b200.IDSelectorBitmap(np.packbits(member[: n // 2], bitorder="little"))   # shorter than ntotal
    got = _bits(short.to_bitmap(n), n)
    assert np.array_equal(got[: n // 2], member[: n // 2]) and not got[(n // 2 + 7) // 8 * 8:].any()


def test_search_parameters_object():
    p = hnsw_b200.SearchParametersHNSW(efSearch=48, check_relative_distance=False,
                                       sel=hnsw_b200.IDSelectorRange(2, 5))
    assert p.efSearch == 48 and p.check_relative_distance is False
    assert p.sel.is_member(2) and p.sel.is_member(4) and not p.sel.is_member(5)
