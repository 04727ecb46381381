// beam.cuh — the traversal core shared by search and construction.
//
// Restates, for a group of W warps working on ONE query:
//   * greedy_update_nearest          -> Beam::descend
//   * search_from_candidates / search_neighbors_to_add     -> Beam::run
//     (App. A.6 / A.9; both reduce to "keep the ef best vertices seen, expand the
//      closest unexpanded one, stop when none is left" — see DESIGN.md §3)
// with the CPU structures replaced by:
//   MinimaxHeap + result heap -> one sorted ef-list of 64-bit (dist,id) keys in shared
//                                memory with an "expanded" bit, merged in place once per hop
//                                (rank + shift); with an IDSelector, a second k-entry list
//                                holds the selector-filtered results;
//   VisitedTable              -> a shared-memory table that may FORGET: a forgotten vertex can be
//                                scored again, but it was rejected against a threshold that only
//                                tightens — or it is still in the list and the merge drops the
//                                identical key — so results are unchanged (only ndis grows).
//                                Default policy (kVisitedAssoc16/32): set-associative, 16-byte
//                                buckets read with one 128-bit load, FIFO eviction inside the
//                                bucket, no global reset, no atomics; with 16-bit slots a bucket
//                                holds eight quotiented ids (the bucket index is a bijective hash's
//                                top bits, the slot keeps the rest), so 4 KB remember 2048 vertices.
//                                kVisitedExact: open addressing over 32-bit slots, exact until 3/4
//                                full (then cleared and re-seeded from the list) — the checker mode;
//   fvec_L2sqr / inner product-> TEAM lanes per vector, 128-bit gathers, R vectors in flight
//                                per team, fixed fmaf order + xor-butterfly (bit-reproducible;
//                                the CPU checker's team mode emulates it exactly). Rows may be
//                                stored as fp16 (HALF): chunks are widened exactly to fp32.
#pragma once
#include "common.cuh"

namespace bh {

struct BeamStats {
    int ndis0 = 0, nhops0 = 0, ndis_up = 0, nhops_up = 0;
#ifdef BH_PHASE_TIMING  // debug builds only: SM cycles per phase, summed over the query's hops
    long long t_pop = 0, t_row = 0, t_hash = 0, t_gather = 0, t_merge = 0, t_reset = 0, t_probe = 0;
    int n_reset = 0;
#endif
};

#ifdef BH_PHASE_TIMING
#define BH_T(var) const long long var = clock64()
#define BH_ACC(field, a, b) st.field += (b) - (a)
#else
#define BH_T(var)
#define BH_ACC(field, a, b)
#endif

// Per-group shared memory carve-up. Host and device must agree: see group_smem_bytes().
struct GroupSmem {
    uint64_t* mbar;
    float4* qbuf;
    unsigned long long* list;      // [ef] sorted keys, merged in place
    int32_t* cand_id;              // [deg] ids to score this hop; reused for merge positions
    unsigned long long* cand_key;  // [deg]
    unsigned long long* acc_key;   // [deg]
    int32_t* ctrl;  // [0]=n_new / stop, [1]=lsize, [2]=rsize, [3]=work index
    uint32_t* hash;
    unsigned long long* rlist;     // [rk] selector-filtered result list (rk = k with an IDSelector, else 0)
};

__host__ __device__ inline size_t round16(size_t x) { return (x + 15) & ~size_t(15); }

// deg = longest adjacency row (2M); rk = result-list entries (k when an IDSelector is set, else 0)
__host__ __device__ inline size_t group_smem_bytes(int d, int ef, int hash_slots, int deg, int rk = 0) {
    return 16 + round16((size_t)d * 4) + round16((size_t)ef * 8) + round16((size_t)deg * 4) +
           2 * round16((size_t)deg * 8) + 32 + (size_t)hash_slots * 4 + round16((size_t)rk * 8);
}

__device__ inline GroupSmem carve_group_smem(unsigned char* p, int d, int ef, int hash_slots, int deg, int rk = 0) {
    GroupSmem s;
    s.mbar = reinterpret_cast<uint64_t*>(p);
    p += 16;
    s.qbuf = reinterpret_cast<float4*>(p);
    p += round16((size_t)d * 4);
    s.list = reinterpret_cast<unsigned long long*>(p);
    p += round16((size_t)ef * 8);
    s.cand_id = reinterpret_cast<int32_t*>(p);
    p += round16((size_t)deg * 4);
    s.cand_key = reinterpret_cast<unsigned long long*>(p);
    p += round16((size_t)deg * 8);
    s.acc_key = reinterpret_cast<unsigned long long*>(p);
    p += round16((size_t)deg * 8);
    s.ctrl = reinterpret_cast<int32_t*>(p);
    p += 32;
    s.hash = reinterpret_cast<uint32_t*>(p);
    p += (size_t)hash_slots * 4;
    s.rlist = rk ? reinterpret_cast<unsigned long long*>(p) : nullptr;
    return s;
}

template <int TEAM, int CPL, int W, int R, bool HALF = false>
struct Beam {
    static constexpr int TPW = 32 / TEAM;  // teams per warp
    static constexpr int NT = TPW * W;     // teams per query group
    static constexpr int ES = HALF ? 2 : 1;  // float4s of query per 16-byte stored chunk

    const GraphView& g;
    const GroupSmem& s;
    const int wig;   // warp index within the group (0 = leader)
    const int lane;
    const int bar_id;
    float4 q[CPL][ES];  // this lane's slice of the query (fp32), laid out like the stored chunks

    __device__ Beam(const GraphView& g_, const GroupSmem& s_, int wig_, int lane_, int bar_id_)
        : g(g_), s(s_), wig(wig_), lane(lane_), bar_id(bar_id_) {}

    __device__ __forceinline__ void group_sync() const {
        if (W == 1)
            __syncwarp();
        else
            bar_sync(bar_id, 32 * W);
    }

    // qbuf holds either an fp32 query (search) or a stored row (construction: fp32 or fp16).
    __device__ __forceinline__ void load_query_from_smem(bool stored_row) {
        const int lit = lane % TEAM;
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const int chunk = c * TEAM + lit;
            const bool ok = chunk < g.nchunk;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (!HALF) {
                q[c][0] = ok ? s.qbuf[chunk] : z;
            } else {
                if (stored_row) {
                    chunk_to_f32<true>(ok ? s.qbuf[chunk] : z, q[c], g.half);
                } else {
                    q[c][0] = ok ? s.qbuf[2 * chunk] : z;
                    q[c][1] = ok ? s.qbuf[2 * chunk + 1] : z;
                }
            }
        }
    }

    // Distances from the query to cand_id[0..n): result keys into cand_key[0..n).
    // Team t scores rows t, t+NT, …; R rows' worth of 128-bit gathers are issued before
    // the first is consumed.
    __device__ __forceinline__ void compute_dists(int n) const {
        if (g.nchunk == TEAM * CPL)
            compute_dists_impl<true>(n);   // every lane owns exactly CPL chunks: no per-chunk predicate
        else
            compute_dists_impl<false>(n);
    }

    template <int C>
    __device__ __forceinline__ void load_chunks(float4 (&x)[CPL], const float4* rowp, bool ok, int lit) const {
        if constexpr (C < CPL) {
            x[C] = (ok && (C * TEAM + lit) < g.nchunk) ? ldg_stream_off<C * TEAM * 16>(rowp)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            load_chunks<C + 1>(x, rowp, ok, lit);
        }
    }
    template <int C>
    __device__ __forceinline__ void load_chunks_full(float4 (&x)[CPL], const float4* rowp) const {
        if constexpr (C < CPL) {
            x[C] = ldg_stream_off<C * TEAM * 16>(rowp);
            load_chunks_full<C + 1>(x, rowp);
        }
    }

    template <bool FULL>
    __device__ __forceinline__ void compute_dists_impl(int n) const {
        const int lit = lane % TEAM;
        const int team = wig * TPW + lane / TEAM;
        const float4* __restrict__ base = reinterpret_cast<const float4*>(g.vecs) + lit;
        const bool l2 = g.is_l2 != 0;
        for (int r0 = 0; r0 < n; r0 += NT * R) {
            float4 x[R][CPL];
            int id[R];
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int r = r0 + k * NT + team;
                id[k] = r < n ? s.cand_id[r] : -1;
            }
#pragma unroll
            for (int k = 0; k < R; k++) {
                // rows past the end re-read row 0 (always valid) instead of branching around the loads
                const float4* rowp = base + (size_t)(id[k] < 0 ? 0 : id[k]) * g.nchunk;
                if (FULL)
                    load_chunks_full<0>(x[k], rowp);
                else
                    load_chunks<0>(x[k], rowp, true, lit);
            }
#pragma unroll
            for (int k = 0; k < R; k++) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < CPL; c++) {
                    float4 xf[ES];
                    chunk_to_f32<HALF>(x[k][c], xf, g.half);
#pragma unroll
                    for (int e = 0; e < ES; e++) acc4(acc, q[c][e], xf[e], l2);
                }
#pragma unroll
                for (int off = TEAM / 2; off >= 1; off >>= 1)
                    acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);
                if (!l2) acc = -acc;
                if (lit == 0 && id[k] >= 0) s.cand_key[r0 + k * NT + team] = pack_key(acc, (uint32_t)id[k]);
            }
        }
    }

    // L2 prefetch of the 128-byte lines of cand_id[0..n)'s rows, spread over the group's lanes.
    __device__ __forceinline__ void prefetch_rows(int n) const {
        const int lines = (g.nchunk + 7) >> 3;  // 128-byte lines per row
        const char* base = reinterpret_cast<const char*>(g.vecs);
        const size_t row_bytes = (size_t)g.nchunk * 16;
        for (int i = wig * 32 + lane; i < n * lines; i += 32 * W) {
            const int r = i / lines, l = i - r * lines;
            const char* p = base + (size_t)s.cand_id[r] * row_bytes + (size_t)l * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
    }

    __device__ __forceinline__ const int32_t* row_ptr(int v, int level, int& deg) const {
        if (level == 0) {
            deg = g.deg0;
            return g.nbr0 + (size_t)v * g.deg0;
        }
        deg = g.degU;
        const int b = __ldg(g.upper_base + v);
        return g.upper_nbr + ((size_t)b + (level - 1)) * g.degU;
    }

    // Leader warp: load a row, cut it at the first -1 (faiss: `if (v < 0) break`).
    // ids[i] holds row[lane + 32 i] or -1.
    __device__ __forceinline__ void load_row(int v, int level, int (&ids)[kMaxIdsPerLane]) const {
        int deg;
        const int32_t* row = row_ptr(v, level, deg);
        int first_neg = kMaxDeg;
#pragma unroll
        for (int i = 0; i < kMaxIdsPerLane; i++) {
            const int idx = lane + 32 * i;
            ids[i] = idx < deg ? __ldcg(row + idx) : -1;
            const unsigned bal = __ballot_sync(0xffffffffu, ids[i] < 0);
            if (bal && first_neg == kMaxDeg) first_neg = 32 * i + __ffs(bal) - 1;
        }
#pragma unroll
        for (int i = 0; i < kMaxIdsPerLane; i++)
            if (lane + 32 * i >= first_neg) ids[i] = -1;
    }

    // ---- visited hash (leader warp) --------------------------------------------
    // The table is read as buckets of four slots (one 128-bit shared load per probe step): a step
    // answers "already visited" for four slots at once and names the first free slot, which is then
    // claimed with one CAS (re-reading the same bucket if another lane took it first). Buckets fill
    // front to back and nothing is ever deleted between clears, so a lookup moves to the next bucket
    // only when its bucket has no free slot.
    __device__ __forceinline__ bool hash_test_and_set(uint32_t id, int bits) const {
        const uint32_t bmask = (1u << (bits - 2)) - 1u;
        uint32_t b = hash_id(id, bits - 2);
        for (;;) {
            const uint4 v = lds128_volatile(s.hash + 4 * b);
            if (v.x == id || v.y == id || v.z == id || v.w == id) return false;  // already visited
            const int j = v.x == kEmpty ? 0 : (v.y == kEmpty ? 1 : (v.z == kEmpty ? 2 : (v.w == kEmpty ? 3 : -1)));
            if (j < 0) {
                b = (b + 1) & bmask;
                continue;
            }
            const uint32_t old = atomicCAS(s.hash + 4 * b + j, kEmpty, id);
            if (old == kEmpty) return true;   // newly inserted
            if (old == id) return false;      // a twin lane inserted the same id first
            // slot taken by another lane meanwhile: look at the same bucket again
        }
    }
    __device__ __forceinline__ void hash_clear(int bits) const {
        const int slots = 1 << bits;
        for (int i = lane; i < slots; i += 32) s.hash[i] = kEmpty;
        __syncwarp();
    }

    // Set-associative policy. `bbits` = log2(buckets). Returns true when `id` was not in its bucket
    // (and records it, pushing the bucket's oldest entry out). No atomics: a lane reads its bucket,
    // and if the id is absent writes the bucket back shifted by one entry with the id on top, then
    // looks again. Two lanes racing on one bucket can only lose an insertion (that id is forgotten —
    // harmless, see the header) — every slot always holds an id that WAS visited, so "visited" is
    // never reported for a fresh vertex. (A row that names one vertex twice — never produced by a build —
    // can make two lanes score it in the same hop; the merge drops the twin key.)
    template <bool SLOT16>
    __device__ __forceinline__ bool assoc_test_and_set(uint32_t id, int bbits) const {
        uint32_t b, rem2 = 0, rem = 0;
        if (SLOT16) {
            // (id * odd) mod 2^(bbits+16) is a bijection on the id range, so (bucket, 16-bit remainder)
            // names the id exactly; 0xFFFF marks an empty slot, and the one id per bucket whose
            // remainder is 0xFFFF is simply never remembered.
            const uint32_t h = (id * 2654435761u) & ((1u << (bbits + 16)) - 1u);
            b = h >> 16;
            rem = h & 0xFFFFu;
            if (rem == 0xFFFFu) return true;
            rem2 = rem | (rem << 16);
        } else {
            b = hash_id(id, bbits);
        }
        bool fresh = false;
        for (;;) {
            const uint4 v = lds128_volatile(s.hash + 4 * b);
            bool found;
            if (SLOT16) {  // "some 16-bit half of v ^ rem2 is zero", the classic has-zero test per word
                const uint32_t x0 = v.x ^ rem2, x1 = v.y ^ rem2, x2 = v.z ^ rem2, x3 = v.w ^ rem2;
                const uint32_t c = 0x00010001u;
                found = ((((x0 - c) & ~x0) | ((x1 - c) & ~x1) | ((x2 - c) & ~x2) | ((x3 - c) & ~x3)) & 0x80008000u) != 0u;
            } else
                found = v.x == id || v.y == id || v.z == id || v.w == id;
            if (found) return fresh;
            uint4 n;
            if (SLOT16) {
                n.x = __funnelshift_r(v.x, v.y, 16);
                n.y = __funnelshift_r(v.y, v.z, 16);
                n.z = __funnelshift_r(v.z, v.w, 16);
                n.w = (v.w >> 16) | (rem << 16);
            } else {
                n = make_uint4(v.y, v.z, v.w, id);
            }
            sts128_volatile(s.hash + 4 * b, n);
            fresh = true;
        }
    }
    __device__ __forceinline__ bool visited_test_and_set(uint32_t id, int mode, int bits) const {
        if (mode == kVisitedAssoc16) return assoc_test_and_set<true>(id, bits - 2);
        if (mode == kVisitedAssoc32) return assoc_test_and_set<false>(id, bits - 2);
        return hash_test_and_set(id, bits);
    }

    // ---- App. A.4: greedy descent from (cur_id) at levels max_level .. stop_level+1 ----
    // All warps of the group call this; on return the leader's (cur_id, cur_d) are valid.
    __device__ void descend(int stop_level, uint32_t& cur_id, float& cur_d, BeamStats& st) const {
        if (wig == 0) {
            if (lane == 0) {
                s.cand_id[0] = g.entry_point;
                s.ctrl[0] = 1;
            }
            __syncwarp();
        }
        group_sync();
        compute_dists(1);
        group_sync();
        int level = g.max_level;
        if (wig == 0) {
            cur_id = (uint32_t)g.entry_point;
            cur_d = key_dist(s.cand_key[0]);
        }
        for (;;) {
            if (wig == 0) {
                int n = -1;
                if (level > stop_level) {
                    int ids[kMaxIdsPerLane];
                    load_row((int)cur_id, level, ids);
                    n = 0;
#pragma unroll
                    for (int i = 0; i < kMaxIdsPerLane; i++) {
                        const unsigned bal = __ballot_sync(0xffffffffu, ids[i] >= 0);
                        if (ids[i] >= 0) s.cand_id[n + __popc(bal & ((1u << lane) - 1u))] = ids[i];
                        n += __popc(bal);
                    }
                    st.ndis_up += n;
                    st.nhops_up += 1;
                }
                if (lane == 0) s.ctrl[0] = n;
                __syncwarp();
            }
            group_sync();
            const int n = s.ctrl[0];
            if (n < 0) break;
            compute_dists(n);
            group_sync();
            if (wig == 0) {
                // sequential "take it if strictly closer" == argmin with first-index ties
                unsigned long long best = ~0ull;
                for (int j = lane; j < n; j += 32) {
                    const unsigned long long kj = (s.cand_key[j] & 0xFFFFFFFF00000000ull) | (unsigned)j;
                    best = kj < best ? kj : best;
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
                    best = o < best ? o : best;
                }
                bool moved = false;
                if (n > 0) {
                    const float bd = ord2f((uint32_t)(best >> 32));
                    if (bd < cur_d) {
                        cur_d = bd;
                        cur_id = (uint32_t)s.cand_id[(int)(best & 0xFFFFFFFFu)];
                        moved = true;
                    }
                }
                if (!moved) level--;
            }
        }
    }

    // ---- App. A.6 / A.9: ef-bounded best-first search at `level` -----------------
    // ef       list capacity (max(efSearch,k) for search, efConstruction for build)
    // ef_stop  stop when the popped entry has >= ef_stop list entries before it
    //          (faiss count_below test; INT_MAX disables)
    // max_steps  faiss `!check_relative_distance && nstep > efSearch` (INT_MAX disables)
    // On return ctrl[1] = list size.
    // sel/rk: faiss IDSelectorBitmap and the k of the selector-filtered result list (nullptr/0: none).
    //          The selector decides only what enters the RESULT list; traversal is unchanged.
    // work_counter / n_items (may be null): the launch's work counter. Once it has run past n_items the launch
    // is DRAINING — resident queries finish one by one and HBM is no longer saturated — and a hop's candidate
    // rows are first prefetched into L2 all at once, so that the register gathers that follow (R rows per team
    // and round, several dependent rounds per hop) wait on L2 instead of DRAM. (In steady state the same
    // prefetch is neutral: the memory system is saturated either way.)
    __device__ void run(int level, int ef, int ef_stop, int max_steps, int vmode, int hash_bits, uint32_t start_id,
                        float start_d, BeamStats& st, const uint8_t* sel = nullptr, int rk = 0,
                        const int* work_counter = nullptr, int n_items = 0) const {
        int lsize = 0, cursor = 0, hcount = 0, nstep = 0, rsize = 0, rcursor = 0;
        const int hlimit = (3 << hash_bits) >> 2;  // kVisitedExact: reset above 75 % load
        if (wig == 0) {
            hash_clear(hash_bits);
            if (lane == 0) {
                s.list[0] = pack_key(start_d, start_id);
                visited_test_and_set(start_id, vmode, hash_bits);
            }
            lsize = 1;
            hcount = 1;
            if (sel && ((sel[start_id >> 3] >> (start_id & 7)) & 1)) {
                if (lane == 0) s.rlist[0] = pack_key(start_d, start_id);
                rsize = 1;
            }
            __syncwarp();
        }
        for (;;) {
            if (wig == 0) {
                // -- pop_min: first unexpanded entry of the sorted list
                BH_T(t0);
                int pos = -1;
                const unsigned long long* L = s.list;
                for (int b = cursor & ~31; b < lsize; b += 32) {
                    const int i = b + lane;
                    const bool un = i >= cursor && i < lsize && !(L[i] & kExpanded);
                    const unsigned bal = __ballot_sync(0xffffffffu, un);
                    if (bal) {
                        pos = b + __ffs(bal) - 1;
                        break;
                    }
                }
                int n_new = -1;
                if (pos >= 0 && pos < ef_stop && nstep <= max_steps) {
                    const uint32_t v0 = key_id(L[pos]);
                    __syncwarp();
                    if (lane == 0) s.list[pos] = L[pos] | kExpanded;
                    cursor = pos + 1;
                    int ids[kMaxIdsPerLane];
                    BH_T(t1);
                    BH_ACC(t_pop, t0, t1);
                    load_row((int)v0, level, ids);
                    BH_T(t2);
                    BH_ACC(t_row, t1, t2);
                    BH_T(t2a);
                    if (vmode == kVisitedExact && hcount + (level == 0 ? g.deg0 : g.degU) > hlimit) {  // forget-and-reseed (see header)
                        hash_clear(hash_bits);
#ifdef BH_PHASE_TIMING
                        st.n_reset++;
#endif
                        for (int i = lane; i < lsize; i += 32) hash_test_and_set(key_id(L[i]), hash_bits);
                        hcount = lsize;
                        if (sel) {  // results outside the list must stay "visited" too, or they could re-enter twice
                            int extra = 0;
                            for (int i = lane; i < rsize; i += 32)
                                extra += hash_test_and_set(key_id(s.rlist[i]), hash_bits) ? 1 : 0;
#pragma unroll
                            for (int off = 16; off >= 1; off >>= 1) extra += __shfl_xor_sync(0xffffffffu, extra, off);
                            hcount += extra;
                        }
                        __syncwarp();
                    }
                    BH_T(t2b);
                    BH_ACC(t_reset, t2a, t2b);
                    n_new = 0;
#pragma unroll
                    for (int i = 0; i < kMaxIdsPerLane; i++) {
                        const bool isnew = ids[i] >= 0 && visited_test_and_set((uint32_t)ids[i], vmode, hash_bits);
                        const unsigned bal = __ballot_sync(0xffffffffu, isnew);
                        if (isnew) s.cand_id[n_new + __popc(bal & ((1u << lane) - 1u))] = ids[i];
                        n_new += __popc(bal);
                    }
                    hcount += n_new;
                    st.ndis0 += n_new;
                    st.nhops0 += 1;
                    nstep++;
                    BH_T(t3);
                    BH_ACC(t_hash, t2, t3);
                    BH_ACC(t_probe, t2b, t3);
                }
                if (lane == 0) s.ctrl[0] = n_new;
                __syncwarp();
            }
            group_sync();
            const int n_new = s.ctrl[0];
            if (n_new < 0) break;
            BH_T(t4);
            if (work_counter && n_new > NT * R && __ldcg(work_counter) >= n_items) prefetch_rows(n_new);
            compute_dists(n_new);
            group_sync();
            BH_T(t5);
            BH_ACC(t_gather, t4, t5);
            if (wig == 0 && n_new > 0) {
                merge(s.list, n_new, ef, lsize, cursor, nullptr);
                if (sel) merge(s.rlist, n_new, rk, rsize, rcursor, sel);
            }
            BH_T(t6);
            BH_ACC(t_merge, t5, t6);
        }
        if (wig == 0) {
            if (lane == 0) {
                s.ctrl[1] = lsize;
                s.ctrl[2] = rsize;
            }
            __syncwarp();
        }
    }

    // Merge the hop's scored candidates into the sorted list, in place (leader warp).
    // Equivalent to pushing them one by one into faiss's bounded MinimaxHeap: the list ends up
    // holding the ef smallest keys of (list ∪ candidates); on exact distance ties the id breaks it.
    // Every old entry moves right by the number of accepted keys below it, so the list is walked
    // from its tail in 32-entry chunks (read chunk, sync, write chunk: a chunk's writes land at or
    // above its own base, i.e. only on slots already vacated); entries below the smallest
    // insertion point are not touched at all; accepted keys drop into the holes at the end.
    // `L` is the candidate list (capacity ef) or, with `sel`, the selector-filtered result list.
    __device__ __forceinline__ void merge(unsigned long long* L, int n_new, int ef, int& lsize, int& cursor,
                                          const uint8_t* sel) const {
        const bool full = lsize == ef;
        const unsigned long long thr = full ? key_clean(L[ef - 1]) : ~0ull;
        int n_acc = 0;
        for (int b = 0; b < n_new; b += 32) {
            const int j = b + lane;
            unsigned long long kj = ~0ull;
            bool ok = false;
            if (j < n_new) {
                kj = s.cand_key[j];
                ok = kj < thr;
                if (ok && sel) {
                    const uint32_t id = key_id(kj);
                    ok = (__ldg(sel + (id >> 3)) >> (id & 7)) & 1;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (ok) s.acc_key[n_acc + __popc(bal & ((1u << lane) - 1u))] = kj;
            n_acc += __popc(bal);
        }
        if (n_acc == 0) return;
        __syncwarp();
        int* acc_pos = s.cand_id;  // free again: the hop's ids have been scored
        int minpos;
        for (;;) {
            minpos = ef;
            bool dup = false;
            for (int a0 = 0; a0 < n_acc; a0 += 32) {
                const int a = a0 + lane;
                if (a < n_acc) {
                    const unsigned long long ka = s.acc_key[a];
                    int ra = 0;
                    bool twin = false;  // the same key earlier in this batch (a row that names a vertex twice)
                    for (int b2 = 0; b2 < n_acc; b2++) {
                        const unsigned long long kb = s.acc_key[b2];
                        ra += kb < ka;
                        twin |= kb == ka && b2 < a;
                    }
                    int lo = 0, hi = lsize;  // lower_bound on clean keys
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (key_clean(L[mid]) < ka) lo = mid + 1; else hi = mid;
                    }
                    // the identical (distance, id) key is already listed: a vertex the visited table
                    // forgot was scored again — it must not enter twice
                    if (twin || (lo < lsize && key_clean(L[lo]) == ka)) dup = true, s.acc_key[a] = ~0ull;
                    const int pos = ra + lo;
                    acc_pos[a] = pos;
                    minpos = pos < minpos ? pos : minpos;
                }
            }
            if (!__any_sync(0xffffffffu, dup)) break;
            // rare: squeeze the dropped keys out of acc_key (in place, front to back) and rank again
            __syncwarp();
            int m = 0;
            for (int a0 = 0; a0 < n_acc; a0 += 32) {
                const int a = a0 + lane;
                const unsigned long long ka = a < n_acc ? s.acc_key[a] : ~0ull;
                const bool keep = ka != ~0ull;
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                __syncwarp();
                if (keep) s.acc_key[m + __popc(bal & ((1u << lane) - 1u))] = ka;
                m += __popc(bal);
                __syncwarp();
            }
            n_acc = m;
            if (n_acc == 0) return;
        }
        minpos = __reduce_min_sync(0xffffffffu, minpos);  // REDUX: one instruction
        __syncwarp();
        // entries before the first insertion point stay where they are (nothing accepted is below them)
        if (lsize > 0) {
            for (int b = (lsize - 1) & ~31; b >= (minpos & ~31); b -= 32) {
                const int i = b + lane;
                unsigned long long e = 0;
                int np = ef;
                if (i < lsize) {
                    e = L[i];
                    int cnt = 0;
                    if (i >= minpos) {
                        const unsigned long long ec = key_clean(e);
                        for (int b2 = 0; b2 < n_acc; b2++) cnt += s.acc_key[b2] < ec;
                    }
                    np = i + cnt;
                }
                __syncwarp();
                if (np < ef) L[np] = e;
                __syncwarp();
            }
        }
        for (int a = lane; a < n_acc; a += 32) {
            const int pos = acc_pos[a];
            if (pos < ef) L[pos] = s.acc_key[a];
        }
        lsize = lsize + n_acc < ef ? lsize + n_acc : ef;
        cursor = minpos < cursor ? minpos : cursor;
        __syncwarp();
    }
};

}  // namespace bh
