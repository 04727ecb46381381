/* b200_hnsw.h — C-ABI of the B200-native HNSW engine (libb200hnsw.so).
 *
 * Drop-in boundary for the one hot path this repo accelerates: faiss::IndexHNSWFlat
 * train / add / search, which is what the reference repo is built on
 * (/root/reference/README.md:2 — the mount holds no source, so there is no reference
 * file:line to cite beyond that; the upstream interface each entry point replaces is
 * named instead).
 *
 * Conventions (mirroring faiss's own c_api): every function returns int, 0 = OK,
 * non-zero = error with a thread-local message from bh_last_error(); no C++ exception
 * crosses this boundary. idx_t is int64_t. All `const float* x` arguments are HOST
 * pointers to contiguous row-major fp32 unless the function name ends in `_device`.
 * The caller owns every buffer it passes; add() copies vectors into index-owned HBM.
 * There is NO CPU fallback: creating an index without a usable CUDA device fails.
 */
#ifndef B200_HNSW_H
#define B200_HNSW_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bh_index bh_index; /* opaque */

/* faiss/MetricType.h: METRIC_INNER_PRODUCT = 0, METRIC_L2 = 1 */
#define BH_METRIC_INNER_PRODUCT 0
#define BH_METRIC_L2 1

/* Per-call search parameters — faiss::SearchParametersHNSW {efSearch,
 * check_relative_distance} plus engine knobs. Zero-initialise for defaults. */
typedef struct bh_search_params {
    int32_t efSearch;                /* <=0: use the index's hnsw.efSearch */
    int32_t check_relative_distance; /* 0: index default, 1: on, 2: off */
    int32_t warps_per_query;         /* 0: auto; 1,2,4,8: warps cooperating on one query */
    int32_t hash_bits;               /* 0: auto; else the visited table is 4 << hash_bits bytes per query
                                        (with visited_policy 0 a non-zero value selects the exact table) */
    int32_t* stats;                  /* optional int32[n][4]: {ndis L0, nhops L0, ndis upper,
                                        nhops upper}; host ptr (device ptr for *_device) */
    const uint8_t* sel_bitmap;       /* optional faiss::IDSelectorBitmap (SearchParametersHNSW::sel):
                                        id i may be returned iff bit (i & 7) of byte (i >> 3) is set.
                                        As in faiss it filters results only, not the traversal.
                                        host ptr (device ptr for *_device); NULL = no selector */
    int64_t sel_bitmap_bytes;        /* size of sel_bitmap, must be >= (ntotal + 7) / 8 */
    int32_t visited_policy;          /* visited-set policy (faiss VisitedTable). Results never depend on it,
                                        only the number of distance evaluations does:
                                        0 auto: hash_bits == 0 -> set-associative forgetful table (default);
                                                hash_bits  > 0 -> exact table of 2^hash_bits slots
                                        1 exact until 3/4 full, then cleared and re-seeded from the list
                                        2 set-associative, FIFO eviction per 16-byte bucket */
    int32_t reserved_;               /* keep zero */
} bh_search_params;

/* Build-time knobs (no faiss equivalent: faiss's concurrency is the OpenMP thread count). */
typedef struct bh_build_params {
    int32_t max_batch;       /* points inserted concurrently per round; <=0: auto */
    int32_t batch_divisor;   /* a round inserts at most ntotal_so_far / batch_divisor points
                                (>=1 point); <=0: auto. max_batch=1 => faiss's sequential order */
    int32_t warps_per_query; /* 0: auto */
    int32_t hash_bits;       /* 0: auto */
    int32_t visited_policy;  /* as bh_search_params.visited_policy */
} bh_build_params;

/* vector storage inside the index (the API is fp32 either way) */
#define BH_STORAGE_F32 0 /* faiss IndexFlat codes: the default, bit-comparable to faiss */
#define BH_STORAGE_F16 1 /* opt-in: rows rounded to IEEE fp16 on add, fp32 accumulation; halves the gather
                            bytes; distances differ from fp32 storage by ~1e-3 relative */
#define BH_STORAGE_BF16 2 /* opt-in: rows rounded to bfloat16 on add (fp32's range, 8 bits of mantissa: ~4e-3
                             relative per component), fp32 accumulation */

/* -- lifecycle ------------------------------------------------------------------- */
/* replaces faiss::IndexHNSWFlat::IndexHNSWFlat(int d, int M, MetricType metric) */
int bh_index_create(bh_index** out, int d, int M, int metric, int device);
/* replaces faiss::Index::~Index */
int bh_index_free(bh_index* h);
/* replaces faiss::IndexHNSW::reset */
int bh_index_reset(bh_index* h);
/* extension (no faiss equivalent; closest: IndexHNSWSQ with QT_fp16): choose BH_STORAGE_*.
 * Only on an empty index. */
int bh_index_set_vector_storage(bh_index* h, int storage);
int bh_index_get_vector_storage(const bh_index* h);

/* -- the path: train / add / search ---------------------------------------------- */
/* replaces faiss::IndexHNSW::train (a no-op for Flat storage; is_trained is true) */
int bh_index_train(bh_index* h, int64_t n, const float* x);
/* replaces faiss::IndexHNSW::add → storage->add + hnsw_add_vertices */
int bh_index_add(bh_index* h, int64_t n, const float* x);
/* same, with caller-supplied levels (level+1 per point, faiss `hnsw.levels` preset) and an
 * optional explicit insertion order (permutation of [ntotal, ntotal+n)); either may be NULL */
int bh_index_add_ex(bh_index* h, int64_t n, const float* x, const int32_t* levels,
                    const int32_t* order);
/* replaces faiss::IndexHNSW::search(n, x, k, distances, labels, params) */
int bh_index_search(const bh_index* h, int64_t n, const float* x, int64_t k, float* distances,
                    int64_t* labels, const bh_search_params* params);
/* extension: x, distances, labels (and params->stats) are DEVICE pointers on the index's
 * device; enqueued on the index's stream; returns without synchronising */
int bh_index_search_device(const bh_index* h, int64_t n, const float* x, int64_t k,
                           float* distances, int64_t* labels, const bh_search_params* params);
/* replaces faiss::IndexHNSW::reconstruct */
int bh_index_reconstruct(const bh_index* h, int64_t key, float* out);
/* replaces faiss::Index::reconstruct_n(i0, ni, recons) */
int bh_index_reconstruct_n(const bh_index* h, int64_t i0, int64_t ni, float* out);

/* -- other faiss IDSelectors, converted on the host into the bitmap form above ------------------
 * bitmap: caller-allocated, (ntotal + 7) / 8 bytes; pass it as bh_search_params.sel_bitmap.
 * replaces faiss::IDSelectorRange(imin, imax): member iff imin <= id < imax */
int bh_selector_range_to_bitmap(int64_t ntotal, int64_t imin, int64_t imax, uint8_t* bitmap);
/* replaces faiss::IDSelectorBatch(n, ids) / IDSelectorArray: member iff id is listed (ids outside
 * [0, ntotal) are ignored, as they can never be returned) */
int bh_selector_batch_to_bitmap(int64_t ntotal, int64_t n, const int64_t* ids, uint8_t* bitmap);
/* replaces faiss::IDSelectorNot: flips membership of every id < ntotal, in place */
int bh_selector_not(int64_t ntotal, uint8_t* bitmap);

/* -- fields (faiss: index.d, index.ntotal, index.hnsw.efSearch, …) ---------------- */
int64_t bh_index_ntotal(const bh_index* h);
int bh_index_d(const bh_index* h);
int bh_index_M(const bh_index* h);
int bh_index_metric(const bh_index* h);
int bh_index_entry_point(const bh_index* h);
int bh_index_max_level(const bh_index* h);
int bh_index_get_ef_search(const bh_index* h);
int bh_index_set_ef_search(bh_index* h, int ef);
int bh_index_get_ef_construction(const bh_index* h);
int bh_index_set_ef_construction(bh_index* h, int ef);
int bh_index_set_check_relative_distance(bh_index* h, int on);
int bh_index_set_build_params(bh_index* h, const bh_build_params* p);

/* -- graph exchange in faiss's HNSW layout (hnsw.levels / offsets / neighbors) ----- */
/* number of int32 entries in `neighbors` for the current graph */
int64_t bh_index_neighbors_size(const bh_index* h);
int bh_index_export_graph(const bh_index* h, int32_t* levels, uint64_t* offsets,
                          int32_t* neighbors);
/* replaces read_index for an in-memory graph: vectors + levels + neighbors (+ entry) */
int bh_index_import_graph(bh_index* h, int64_t n, const float* x, const int32_t* levels,
                          const int32_t* neighbors, int64_t nneighbors, int entry_point,
                          int max_level);

/* -- streams & timing -------------------------------------------------------------- */
/* cudaStream_t the index enqueues on (as void*) */
void* bh_index_stream(const bh_index* h);
int bh_index_synchronize(const bh_index* h);
/* device milliseconds of the most recent add() graph-construction phase / search launch */
float bh_index_last_build_ms(const bh_index* h);
float bh_index_last_search_ms(const bh_index* h);
/* work counters of the most recent add(): {distance evaluations level 0, hops level 0, distance
 * evaluations upper levels, hops upper levels, candidate vectors read by the selection heuristic,
 * vectors streamed by back-link shrinks} — the numerators of the build's HBM roofline */
int bh_index_last_build_counters(const bh_index* h, uint64_t out[6]);
/* kernels launched by this library since load (for bench.py's gpu_launches) */
int64_t bh_launch_count(void);

/* -- sharded search: merge of per-shard top-k lists (replaces faiss merge_knn_results
 *    as used by IndexShards(successive_ids=true)) ------------------------------------
 * D_all / I_all: device, [nshard][nq][k], each list sorted best-first, I local to its
 * shard (-1 = empty). Writes device D_out/I_out [nq][k] with I + id_offsets[shard].
 * stream: cudaStream_t as void* (NULL = default stream). Enqueued only: no allocation, no
 * synchronisation (id_offsets is consumed before the call returns). nshard <= 64. */
int bh_merge_topk_device(int nshard, int64_t nq, int64_t k, int metric, const float* D_all,
                         const int64_t* I_all, const int64_t* id_offsets /*host, [nshard]*/,
                         float* D_out, int64_t* I_out, void* stream);

/* -- sharded search over the GPUs of one box (replaces faiss::IndexShards(successive_ids=true) and its
 *    merge_knn_results; faiss/IndexShards.cpp) ------------------------------------------------------
 * One bh_shards per rank (= per GPU / per local index); the ranks may live in one process or in one
 * process each. Rank r owns global ids [sum(ntotal_0..r-1), +ntotal_r). Bootstrap: every rank exports
 * a fixed-size blob, the caller moves the blobs between ranks over any channel it owns (MPI,
 * a socket, a file), every rank connects with all of them (CUDA IPC across processes, plain
 * peer access inside one). A search is collective — every rank calls it with the same queries: the
 * traversal kernel's epilogue stores each query's k results as 8-byte (distance bits, local id) keys
 * straight into every rank's gather buffer over NVLink, a one-warp kernel raises a flag in every
 * peer, and the merge kernel waits for all flags and merges: no library collective, no host round
 * trip. max_queries / max_k size the gather buffers (larger batches are sliced). */
typedef struct bh_shards bh_shards;
#define BH_SHARDS_BLOB_BYTES 160
int bh_shards_create(bh_shards** out, bh_index* local, int rank, int nranks, int64_t max_queries,
                     int64_t max_k);
int bh_shards_free(bh_shards* s);
int bh_shards_export(bh_shards* s, void* blob /* [BH_SHARDS_BLOB_BYTES] */);
int bh_shards_connect(bh_shards* s, const void* blobs /* [nranks][BH_SHARDS_BLOB_BYTES], rank order */);
/* shard sizes for the id offsets, if the shards grew after connect (host array [nranks]) */
int bh_shards_set_ntotals(bh_shards* s, const int64_t* ntotals);
/* replaces IndexShards::search; device pointers on this rank's GPU, enqueued on the local index's
 * stream, no synchronisation; labels are global ids */
int bh_shards_search_device(bh_shards* s, int64_t n, const float* x, int64_t k, float* distances,
                            int64_t* labels, const bh_search_params* params);
/* the two halves of a search, for callers that exchange the lists themselves or emulate ranks:
 * post = traverse + publish (publish_to_peers = 0: only into this rank's own gather buffer);
 * gather = device pointer of this call's [nranks][n][k] packed gather buffer;
 * collect = (wait for every rank's flag if wait_for_peers) + merge;
 * local_lists = this rank's own lists of the current call, local ids */
int bh_shards_post(bh_shards* s, int64_t n, const float* x, int64_t k, const bh_search_params* params,
                   int publish_to_peers);
int bh_shards_gather(bh_shards* s, int64_t n, int64_t k, void** ptr);
int bh_shards_collect(bh_shards* s, int64_t n, int64_t k, float* distances, int64_t* labels,
                      int wait_for_peers);
int bh_shards_local_lists(bh_shards* s, int64_t n, int64_t k, float* distances, int64_t* labels);
/* Pipelined mode (off by default). On: the flag and merge kernels run on a separate exchange stream, so the
 * index's stream carries only traversal launches — consecutive calls overlap their drain phases and do not
 * wait for slower peers; D / I of a call are complete only after bh_shards_join(s, stream) has made `stream`
 * (cudaStream_t as void*, NULL = the local index's stream) wait for the latest merge. Merges run in call
 * order. Join before switching the mode back. */
int bh_shards_set_pipelined(bh_shards* s, int on);
int bh_shards_join(bh_shards* s, void* stream);
/* 0 = fine; 1 = a peer did not publish within the timeout (results of that call are undefined) */
int bh_shards_status(bh_shards* s);

const char* bh_last_error(void);
const char* bh_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_HNSW_H */
