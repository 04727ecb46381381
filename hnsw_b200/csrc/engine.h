// engine.h — host-side declarations shared by the kernel translation units and the
// C-ABI layer (capi.cu). Nothing here is exported; the public surface is include/b200_hnsw.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace bh {

constexpr int kMaxPeers = 16;  // GPUs of one box that exchange results by peer stores

// One batch of traversal work for beam_kernel.
struct BeamTask {
    // search mode (items == nullptr)
    const float* queries;  // device [n_items][d]
    int k;
    float* D;              // device [n_items][k]
    int64_t* I;            // device [n_items][k]
    // construction mode
    const int4* items;     // device [n_items] {point id, search level, stop level (= pt_level), 0}
    unsigned long long* out_lists;  // device [n_items][ef] sorted clean keys
    int32_t* out_counts;            // device [n_items]
    // common
    int n_items;
    int ef;         // list capacity
    int ef_stop;    // count_below threshold (INT_MAX = off)
    int max_steps;  // nstep limit (INT_MAX = off)
    int hash_bits;     // visited table = 4 << hash_bits bytes per query
    int visited_mode;  // kVisitedExact / kVisitedAssoc16 / kVisitedAssoc32 (beam.cuh)
    int32_t* stats;  // device int32[n_items][4] or null
    unsigned long long* build_counters;  // device [6] or null: {ndis0, nhops0, ndis_up, nhops_up, sel_rows, bl_rows}
    const uint8_t* sel;  // device IDSelectorBitmap over the shard's ids, or null (search mode only)
    int* counter;    // device work counter, zeroed before launch
    int drain_prefetch;  // 1: L2-prefetch a hop's rows once the launch is draining (beam.cuh run())
    int pdl;         // 1: launched with programmatic stream serialisation (search mode; see capi.cu search_device)
    // sharded search (search mode only): instead of D / I the epilogue publishes each query's k results as
    // packed (order-preserving distance bits << 32 | local id) keys, ~0 = empty, straight into the gather
    // buffer of every rank of the box — peer memory over NVLink — at shard_out[p][item * k + i]
    int n_shard_out;
    unsigned long long* shard_out[kMaxPeers];
};

size_t beam_group_smem(int d, int ef, int hash_bits, int deg, int rk = 0);

// ---- construction kernels (build_kernels.cu) ------------------------------------
struct BuildBatch {
    const int4* items;                    // [n_items] {pt, level, pt_level, 0}
    const unsigned long long* cand_lists; // [n_items][efc] sorted (dist,id) keys, nearest first
    const int32_t* cand_counts;           // [n_items]
    int n_items;
    int efc;
    // back-edge staging: one slot per (item, kept neighbour)
    int32_t* edge_dst_slot;   // [n_items*kMaxDeg] row slot of the destination (see row_slot), -1 = none
    int32_t* edge_src;        // [..] the new point
    int32_t* edge_dst;        // [..] destination vertex
    int32_t* edge_level;      // [..]
    float* edge_dist;         // [..] d(src, dst)
    int32_t* edge_next;       // [..] intrusive list link
    int32_t* slot_head;       // [n_slots] head of the per-row pending list, -1 = empty (persistent)
    int64_t n_level0;         // number of level-0 rows (= ntotal capacity used for slot numbering)
    // verified prefix per row: members at positions [0, nver) are known to be mutually consistent
    // under the selection heuristic w.r.t. the row owner (see backlink_kernel)
    uint8_t* nver0;           // [ntotal] level-0 rows
    uint8_t* nverU;           // [n_upper_rows]
    int max_special;          // cap on "unverified" candidates handled by the incremental shrink
    unsigned long long* build_counters;  // device [6] or null (see BeamTask)
};

// `fuse` (construction only): run the selection + forward links + back-edge staging of each item in the
// traversal kernel's epilogue instead of a separate launch_select_and_link
cudaError_t launch_beam(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                        cudaStream_t stream, int* grid_out, const BuildBatch* fuse = nullptr);
cudaError_t launch_select_and_link(const GraphView& g, const BuildBatch& b, int num_sms, cudaStream_t stream);
cudaError_t launch_backlinks(const GraphView& g, const BuildBatch& b, int num_sms, cudaStream_t stream);

// ---- sharded top-k merge (merge_kernel.cu) ---------------------------------------
constexpr int kMaxShards = 64;
struct ShardOffsets {  // passed by value in the kernel parameters: no allocation, copy or sync per merge
    int64_t v[kMaxShards];
};
cudaError_t launch_merge_topk(int nshard, int64_t nq, int k, int is_l2, const float* D_all,
                              const int64_t* I_all, const ShardOffsets& id_offsets, float* D_out,
                              int64_t* I_out, cudaStream_t stream);

// Sharded exchange, packed form (merge_kernel.cu). `gather` = this rank's [nshard][nq][k] packed keys (each
// list written by its owner's traversal kernel, see BeamTask::shard_out). signal: store `epoch` into slot
// `my_rank` of every peer's flag array (release, system scope). merge: wait until all nshard local flags
// reach `epoch` (when flags != nullptr), then the same rank-by-counting merge as launch_merge_topk.
// `status` (mapped host memory) is set to 1 if a peer never arrived within timeout_ms.
struct PeerFlags {
    unsigned long long* v[kMaxPeers];
};
constexpr int kFlagStride = 16;  // flags are 128 bytes apart
cudaError_t launch_shard_signal(int nshard, int my_rank, const PeerFlags& peer_flags, unsigned long long epoch,
                                cudaStream_t stream);
cudaError_t launch_merge_packed(int nshard, int64_t nq, int k, int is_l2, const unsigned long long* gather,
                                const ShardOffsets& id_offsets, float* D_out, int64_t* I_out,
                                const unsigned long long* flags, unsigned long long epoch, int* status,
                                int timeout_ms, cudaStream_t stream);
// packed keys [n] -> (D fp32, I int64 local ids) (a rank's own lists, for callers that want them)
cudaError_t launch_unpack(const unsigned long long* keys, int64_t n, int is_l2, float* D, int64_t* I,
                          cudaStream_t stream);

// ---- storage conversion (merge_kernel.cu): fp32 rows <-> 16-bit rows (fmt 1 = IEEE fp16, 2 = bfloat16),
//      element-wise round-to-nearest-even / exact widening
cudaError_t launch_f32_to_f16(const float* src, void* dst, size_t n, cudaStream_t stream, int fmt = 1);
cudaError_t launch_f16_to_f32(const void* src, float* dst, size_t n, cudaStream_t stream, int fmt = 1);

void count_launch(int n = 1);

}  // namespace bh
