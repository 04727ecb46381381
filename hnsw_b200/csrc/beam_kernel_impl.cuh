#pragma once
// beam_kernel_impl.cuh — persistent traversal kernel (search mode and construction-search mode).
// The templated launchers are in beam_launch.cuh; four translation units (fp32 / 16-bit storage x rows up to
// 512 B / wider rows) instantiate disjoint parts of the table so they compile in parallel.
//
// One launch = one batch of queries (IndexHNSW::search's `omp parallel for` over queries,
// SURVEY.md §3.1) or one batch of (point, level) insertion searches
// (add_links_starting_from → search_neighbors_to_add, §3.2). Groups of W warps pull work
// items from an atomic counter until the batch is drained.
#include "beam.cuh"
#include "engine.h"
#include "select.cuh"

#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace bh {

// FUSE (construction only): the warp that finished an insertion search runs shrink_neighbor_list on the
// candidate list it still holds in shared memory, writes the new vertex's row and stages the back-edges —
// what select_and_link_coop_kernel does as a separate launch. Legal inside the round: no vertex of the graph
// links to a point of the round before the back-link kernel runs, so the rows written here are never read by
// the round's other searches; the heuristic is instruction-bound and hides under the other warps' gathers.
// LEAN (search only): the instantiation for the common search launch — no selector, default visited table, no
// construction items — with those switches compile-time constants, so the selector's result list, the exact
// table and the construction epilogue drop out of the hot kernel. Measured on one box, 1M x 128, 10k queries,
// back to back: efSearch 32 / 64 / 128 / 256 -> 1.23 / 2.29 / 4.47 / 9.25 ms against 1.41 / 2.54 / 4.86 / 9.92 ms
// for the general instantiation (-10 %); isolated launches -3..5 %. (Making the row shape and the metric
// compile-time constants as well was measured too: it gives the whole gain back — at the 80-register cap the
// schedule ptxas finds matters more than the instruction count — so those stay run-time switches.)
// LEAN: 0 = general, 1 = lean search, 2 = lean construction search (items present, no selector, default table).
template <int TEAM, int CPL, int W, int R, int G, int MINB, bool HALF, bool FUSE, int LEAN = 0>
__global__ void __launch_bounds__(32 * W * G, MINB) beam_kernel(GraphView g, BeamTask t, BuildBatch b) {
    const int4* const items = LEAN == 1 ? nullptr : t.items;
    const bool is_build = LEAN == 2 ? true : (LEAN == 1 ? false : t.items != nullptr);
    const uint8_t* const sel = LEAN ? nullptr : t.sel;
    const int vmode = LEAN ? kVisitedAssoc16 : t.visited_mode;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // Nothing this grid reads is produced by the launch before it, so the next search launch on the stream
    // may begin as soon as SM slots free up (no-op unless that launch asked for programmatic serialisation).
    if (t.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int grp = warp / W;
    const int wig = warp % W;
    const int hash_slots = 1 << t.hash_bits;
    const int rk = sel ? t.k : 0;
    const size_t gbytes = group_smem_bytes(g.d, t.ef, hash_slots, g.deg0, rk);
    const GroupSmem s = carve_group_smem(smem_raw + grp * gbytes, g.d, t.ef, hash_slots, g.deg0, rk);
    Beam<TEAM, CPL, W, R, HALF> beam(g, s, wig, lane, 1 + grp);

    if (wig == 0 && lane == 0) {
        mbar_init(s.mbar, 1);
        fence_mbar_init();
    }
    beam.group_sync();
    uint32_t phase = 0;
    const uint32_t qbytes_query = (uint32_t)g.d * 4u;             // fp32 query from the caller
    const uint32_t qbytes_row = (uint32_t)g.nchunk * 16u;         // a stored row (fp32 or fp16)

    for (;;) {
        if (wig == 0 && lane == 0) s.ctrl[3] = atomicAdd(t.counter, 1);
        beam.group_sync();
        const int wi = s.ctrl[3];
        if (wi >= t.n_items) break;

        int level = 0, stop_level = 0;
        const void* qsrc;
        uint32_t qbytes = qbytes_query;
        if (is_build) {  // construction: the query is the stored vector of the new point
            const int4 it = __ldg(items + wi);
            qsrc = reinterpret_cast<const char*>(g.vecs) + (size_t)it.x * qbytes_row;
            qbytes = qbytes_row;
            level = it.y;
            stop_level = it.z;
        } else {
            qsrc = t.queries + (size_t)wi * g.d;
        }
        // query -> shared memory by 1-D bulk TMA, completion on the group's mbarrier
        if (wig == 0 && lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(s.mbar, qbytes);
            tma_load_1d(s.qbuf, qsrc, qbytes, s.mbar);
        }
        mbar_wait(s.mbar, phase);
        phase ^= 1;
        beam.load_query_from_smem(is_build);

        BeamStats st;
        uint32_t cur_id = 0;
        float cur_d = 0.f;
        beam.descend(stop_level, cur_id, cur_d, st);
        beam.run(level, t.ef, t.ef_stop, t.max_steps, vmode, t.hash_bits, cur_id, cur_d, st, sel, rk,
                 t.drain_prefetch ? t.counter : nullptr, t.n_items);

        if (wig == 0) {
            const int lsize = sel ? s.ctrl[2] : s.ctrl[1];
            const unsigned long long* L = sel ? s.rlist : s.list;
            if (is_build) {
                if (t.build_counters && lane == 0) {  // totals for the build's roofline (bench.py)
                    atomicAdd(t.build_counters + 0, (unsigned long long)st.ndis0);
                    atomicAdd(t.build_counters + 1, (unsigned long long)st.nhops0);
                    atomicAdd(t.build_counters + 2, (unsigned long long)st.ndis_up);
                    atomicAdd(t.build_counters + 3, (unsigned long long)st.nhops_up);
                }
                if constexpr (FUSE) {
                    const int4 it = __ldg(items + wi);
                    const int pt = it.x;
                    int deg;
                    int32_t* row = row_ptr_rw(g, pt, level, deg);
                    unsigned long long* Lw = s.list;
                    for (int i = lane; i < lsize; i += 32) Lw[i] = key_clean(Lw[i]);
                    __syncwarp();
                    int K = lsize;
                    const unsigned long long* kept = Lw;
                    const bool verified = lsize >= deg;
                    if (verified) {  // (fewer candidates than slots: shrink_neighbor_list keeps everything)
                        K = heuristic<TEAM, CPL, false, HALF>(g, Lw, lsize, deg, s.cand_key, nullptr, nullptr, nullptr,
                                                              nullptr, lane);
                        kept = s.cand_key;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        *nver_ptr(g, b, pt, level) = verified ? (uint8_t)K : (uint8_t)0;
                        if (t.build_counters && verified) atomicAdd(t.build_counters + 4, (unsigned long long)lsize);
                    }
                    // faiss pops link_targets farthest-first: row[i] = kept[K-1-i]
                    for (int i = lane; i < g.deg0; i += 32) {
                        const int e = wi * g.deg0 + i;
                        if (i < K) {
                            const unsigned long long key = kept[K - 1 - i];
                            const int o = (int)key_id(key);
                            row[i] = o;
                            const int slot = row_slot(g, b.n_level0, o, level);
                            b.edge_src[e] = pt;
                            b.edge_dst[e] = o;
                            b.edge_level[e] = level;
                            b.edge_dist[e] = key_dist(key);
                            b.edge_dst_slot[e] = slot;
                            b.edge_next[e] = atomicExch(b.slot_head + slot, e);
                        } else {
                            if (i < deg) row[i] = -1;
                            b.edge_dst_slot[e] = -1;
                        }
                    }
                    __syncwarp();
                } else {
                    unsigned long long* out = t.out_lists + (size_t)wi * t.ef;
                    for (int i = lane; i < lsize; i += 32) out[i] = key_clean(L[i]);
                    if (lane == 0) t.out_counts[wi] = lsize;
                }
            } else if (t.n_shard_out > 0) {
                // sharded search: this shard's list goes straight into every rank's gather buffer (peer
                // stores over NVLink; 8 bytes per result, one contiguous k*8-byte run per rank)
                for (int i = lane; i < t.k; i += 32) {
                    const unsigned long long kk = i < lsize ? key_clean(L[i]) : ~0ull;
                    for (int p = 0; p < t.n_shard_out; p++) t.shard_out[p][(size_t)wi * t.k + i] = kk;
                }
                // the writer itself makes its peer stores visible system-wide; the flag that announces them is
                // raised by a later kernel (release) and read with acquire by the peers' merge kernels
                if (t.n_shard_out > 1) __threadfence_system();
            } else {
                const float pad = g.is_l2 ? FLT_MAX : -FLT_MAX;
                for (int i = lane; i < t.k; i += 32) {
                    float dd = pad;
                    int64_t id = -1;
                    if (i < lsize) {
                        dd = key_dist(L[i]);
                        if (!g.is_l2) dd = -dd;
                        id = (int64_t)key_id(L[i]);
                    }
                    t.D[(size_t)wi * t.k + i] = dd;
                    t.I[(size_t)wi * t.k + i] = id;
                }
            }
            if (t.stats && lane == 0) {
#ifdef BH_PHASE_TIMING  // debug builds only (scripts/phase_timing.py): SM cycles per hop, by phase
                const int h = st.nhops0 > 0 ? st.nhops0 : 1;
                printf("q%d hops=%d ndis=%d resets=%d cycles/hop: pop=%lld row=%lld hash=%lld (reset=%lld probe=%lld) "
                       "gather=%lld merge=%lld\n",
                       wi, st.nhops0, st.ndis0, st.n_reset, st.t_pop / h, st.t_row / h, st.t_hash / h, st.t_reset / h,
                       st.t_probe / h, st.t_gather / h, st.t_merge / h);
#endif
                int4 sv = make_int4(st.ndis0, st.nhops0, st.ndis_up, st.nhops_up);
                reinterpret_cast<int4*>(t.stats)[wi] = sv;
            }
        }
        beam.group_sync();
    }
}

}  // namespace bh
