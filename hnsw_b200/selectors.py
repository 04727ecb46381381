"""faiss IDSelector family and SearchParametersHNSW, mirrored for IndexHNSWFlat.search(params=...).

Every selector is turned into the IDSelectorBitmap form on the host (bh_selector_* in the C-ABI); the
kernel sees one bitmap. As in faiss's HNSW search the selector filters what may be RETURNED; the
traversal itself is unchanged."""
from __future__ import annotations

import numpy as np

from . import _lib


class IDSelector:
    def to_bitmap(self, ntotal: int) -> np.ndarray:
        raise NotImplementedError

    def is_member(self, i: int) -> bool:
        bm = self.to_bitmap(int(i) + 1)
        return bool((bm[i >> 3] >> (i & 7)) & 1)


class IDSelectorBitmap(IDSelector):
    """faiss.IDSelectorBitmap(n, bitmap): id i is a member iff bit (i & 7) of byte (i >> 3) is set."""

    def __init__(self, bitmap):
        self.bitmap = np.ascontiguousarray(bitmap, np.uint8)

    def to_bitmap(self, ntotal):
        need = (ntotal + 7) // 8
        if self.bitmap.size >= need:
            return self.bitmap
        out = np.zeros(need, np.uint8)      # ids beyond the bitmap are non-members, as in faiss
        out[:self.bitmap.size] = self.bitmap
        return out


class IDSelectorRange(IDSelector):
    """faiss.IDSelectorRange(imin, imax): imin <= id < imax."""

    def __init__(self, imin: int, imax: int):
        self.imin, self.imax = int(imin), int(imax)

    def to_bitmap(self, ntotal):
        out = np.empty((ntotal + 7) // 8, np.uint8)
        _lib.check(_lib.lib().bh_selector_range_to_bitmap(ntotal, self.imin, self.imax, out.ctypes.data))
        return out


class IDSelectorBatch(IDSelector):
    """faiss.IDSelectorBatch(ids) / IDSelectorArray(ids)."""

    def __init__(self, ids):
        self.ids = np.ascontiguousarray(ids, np.int64)

    def to_bitmap(self, ntotal):
        out = np.empty((ntotal + 7) // 8, np.uint8)
        _lib.check(_lib.lib().bh_selector_batch_to_bitmap(ntotal, self.ids.size, self.ids.ctypes.data, out.ctypes.data))
        return out


IDSelectorArray = IDSelectorBatch


class IDSelectorNot(IDSelector):
    """faiss.IDSelectorNot(sel)."""

    def __init__(self, sel: IDSelector):
        self.sel = sel

    def to_bitmap(self, ntotal):
        out = self.sel.to_bitmap(ntotal)[:(ntotal + 7) // 8].copy()
        _lib.check(_lib.lib().bh_selector_not(ntotal, out.ctypes.data))
        return out


class SearchParametersHNSW:
    """faiss.SearchParametersHNSW(efSearch=..., check_relative_distance=..., sel=...)."""

    def __init__(self, efSearch: int | None = None, check_relative_distance: bool | None = None,
                 sel: IDSelector | None = None):
        self.efSearch = efSearch
        self.check_relative_distance = check_relative_distance
        self.sel = sel
