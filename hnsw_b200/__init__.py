"""hnsw_b200 — B200-native (sm_100a) HNSW search-and-build engine behind the faiss
IndexHNSWFlat surface. See DESIGN.md; the C-ABI is include/b200_hnsw.h."""
from .index import (METRIC_INNER_PRODUCT, METRIC_L2, IndexHNSWFlat, launch_count,  # noqa: F401
                    merge_topk_device)
from ._lib import BuildParams, SearchParams, build  # noqa: F401
from .selectors import (IDSelector, IDSelectorArray, IDSelectorBatch, IDSelectorBitmap, IDSelectorNot,  # noqa: F401
                        IDSelectorRange, SearchParametersHNSW)
