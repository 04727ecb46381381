// build_kernels.cu — batched graph construction after the insertion searches.
//
// Restates (SURVEY.md App. A.9–A.11):
//   shrink_neighbor_list  -> heuristic()             (diversity heuristic, warp-parallel)
//   add_links_starting_from: forward links pt→o      -> select_and_link_kernel
//   add_link(o→pt) under omp locks                   -> backlink_kernel
// faiss serialises concurrent back-links to one vertex with a per-vertex omp_lock_t. Here every
// back-edge of the batch is pushed (atomicExch) onto an intrusive per-row pending list; the
// warp that owns the list's tail then applies that row's edges one after the other in edge-index
// order (= insertion order, farthest neighbour first, as faiss does) — no locks, no spinning,
// and the result does not depend on scheduling.
#include <cstdlib>

#include "beam.cuh"
#include "engine.h"
#include "select.cuh"

namespace bh {

namespace {

constexpr int kChainCap = 64;

struct WarpSmem {
    unsigned long long* kept_key;  // [deg0]
    unsigned long long* cand_a;    // [deg0 + 8]
    unsigned long long* cand_b;    // [deg0 + 8]
    int32_t* slot_b;               // [deg0 + 8] staging slot of the sorted candidates
    int32_t* kept_slot;            // [deg0]
    int32_t* rowid;                // [deg0]
    int32_t* chain;                // [kChainCap]
    float4* vbuf;                  // [(deg0 + 1)][nchunk] vector cache (kept / staged candidates) or nullptr
    float* dsx;                    // [max_special][deg0 + 8] d(special_i, candidate_a)   (backlink only)
    float4* spec_vec;              // [max_special][nchunk] vectors of the special candidates (backlink only)
};

__host__ __device__ inline size_t warp_smem_bytes(int d, int deg0, bool vbuf, int max_special = 0) {
    const size_t c8 = (size_t)(deg0 + 8);
    size_t b = (size_t)deg0 * 8 + 2 * c8 * 8 + c8 * 4 + 2 * (size_t)deg0 * 4 + (size_t)kChainCap * 4;
    b = (b + 15) & ~size_t(15);
    b += (vbuf ? (size_t)(deg0 + 1) * d * 4 : 0);
    b += (size_t)max_special * c8 * 4;
    b = (b + 15) & ~size_t(15);
    b += (size_t)max_special * d * 4;
    return b;
}

__device__ inline WarpSmem carve_warp_smem(unsigned char* p, int d, int deg0, bool vbuf, int max_special = 0) {
    WarpSmem w;
    unsigned char* p0 = p;
    const size_t c8 = (size_t)(deg0 + 8);
    w.kept_key = reinterpret_cast<unsigned long long*>(p);
    p += (size_t)deg0 * 8;
    w.cand_a = reinterpret_cast<unsigned long long*>(p);
    p += c8 * 8;
    w.cand_b = reinterpret_cast<unsigned long long*>(p);
    p += c8 * 8;
    w.slot_b = reinterpret_cast<int32_t*>(p);
    p += c8 * 4;
    w.kept_slot = reinterpret_cast<int32_t*>(p);
    p += (size_t)deg0 * 4;
    w.rowid = reinterpret_cast<int32_t*>(p);
    p += (size_t)deg0 * 4;
    w.chain = reinterpret_cast<int32_t*>(p);
    p += (size_t)kChainCap * 4;
    size_t used = ((size_t)(p - p0) + 15) & ~size_t(15);
    w.vbuf = vbuf ? reinterpret_cast<float4*>(p0 + used) : nullptr;
    used += vbuf ? (size_t)(deg0 + 1) * d * 4 : 0;
    w.dsx = reinterpret_cast<float*>(p0 + used);
    used += (size_t)max_special * c8 * 4;
    used = (used + 15) & ~size_t(15);
    w.spec_vec = reinterpret_cast<float4*>(p0 + used);
    return w;
}

// ---- forward links + back-edge staging, CTA-cooperative: NW warps share one (point, level) item ------
//
// The selection heuristic is ~8 500 pairwise distances per item (200 candidates x the ~40 vertices kept so
// far) and strictly sequential in the candidates, so one warp per item is latency-bound: at 33 KB of kept-
// vector cache per warp only ~6 warps fit on an SM and each stalls on its own fmaf chains (round 1:
// issue-active 39 %). Here the NW warps of a CTA hold the SAME 32/TEAM candidates (one per team) and split
// the KEPT set: warp w scores kept blocks w, w+NW, ... (4 kept per block), the per-team "rejected" bits are
// OR-ed through shared memory with one CTA barrier per step, and the short acceptance phase (appending
// to the kept set in candidate order, testing later candidates of the same step against a newly kept one)
// is replayed identically by every warp. One kept-vector cache per CTA, NW x the warps per SM, 1/NW of the
// scan latency per step. Same comparisons on the same values as the one-warp version => same graph.
template <int TEAM, int CPL, int NW, bool KV, bool HALF>
__global__ void __launch_bounds__(32 * NW) select_and_link_coop_kernel(GraphView g, BuildBatch b) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TPW = 32 / TEAM;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int lit = lane % TEAM, team = lane / TEAM;
    unsigned long long* kept_key = reinterpret_cast<unsigned long long*>(smem_raw);     // [deg0]
    unsigned* flags = reinterpret_cast<unsigned*>(smem_raw + (size_t)g.deg0 * 8);       // [2][NW]
    float4* kvec = reinterpret_cast<float4*>(smem_raw + (size_t)g.deg0 * 8 + 64);       // [deg0][nchunk] (KV)
    const float4* __restrict__ vecs = reinterpret_cast<const float4*>(g.vecs);
    const bool is_l2 = g.is_l2 != 0;
    const int nchunk = g.nchunk;
    for (int item = blockIdx.x; item < b.n_items; item += gridDim.x) {
        const int4 it = __ldg(b.items + item);
        const int pt = it.x, level = it.y;
        int deg;
        int32_t* row = row_ptr_rw(g, pt, level, deg);
        const unsigned long long* cand = b.cand_lists + (size_t)item * b.efc;
        const int n = b.cand_counts[item];
        int K = 0;
        const unsigned long long* kept = kept_key;
        const bool verified = n >= deg;
        if (!verified) {  // shrink_neighbor_list returns early: keep everything
            K = n;
            kept = cand;
        } else {
            TeamVec<TEAM, CPL, HALF> nxt;
            unsigned long long nxt_key = ~0ull;
            auto fetch = [&](int c0) {
                const int c = c0 + team;
                const bool valid = c < n;
                nxt_key = valid ? cand[c] : ~0ull;
                nxt.load(vecs + (size_t)(valid ? key_id(nxt_key) : 0) * nchunk, nchunk, lit, valid, g.half);
            };
            auto kept_row = [&](int j) -> const float4* {
                return KV ? kvec + (size_t)j * nchunk : vecs + (size_t)key_id(kept_key[j]) * nchunk;
            };
            fetch(0);
            int parity = 0;
            for (int c0 = 0; c0 < n && K < deg; c0 += TPW, parity ^= 1) {
                const TeamVec<TEAM, CPL, HALF> v = nxt;
                const unsigned long long key = nxt_key;
                const bool valid = c0 + team < n;
                if (c0 + TPW < n) fetch(c0 + TPW);  // in flight while this group is tested
                const uint32_t id = key_id(key);
                const float dq = key_dist(key);
                bool bad = !valid;
                // this warp's share of the kept set: blocks of four, round-robin over the warps
                const int nfull = K >> 2;
                for (int blk = wid; blk < nfull; blk += NW) {
                    if (__all_sync(0xffffffffu, bad)) break;
                    const int j = blk << 2;
                    float duv[4];
                    v.dist4(kept_row(j), kept_row(j + 1), kept_row(j + 2), kept_row(j + 3), nchunk, lit, is_l2, duv, g.half);
                    if (duv[0] < dq || duv[1] < dq || duv[2] < dq || duv[3] < dq) bad = true;
                }
                if (wid == nfull % NW) {  // the partial last block
                    for (int j = nfull << 2; j < K; j++) {
                        if (__all_sync(0xffffffffu, bad)) break;
                        const float duv = v.dist(kept_row(j), nchunk, lit, is_l2, g.half);
                        if (duv < dq) bad = true;
                    }
                }
                const unsigned mine = __ballot_sync(0xffffffffu, bad);
                if (lane == 0) flags[parity * NW + wid] = mine;
                __syncthreads();
                unsigned all = 0;
#pragma unroll
                for (int w2 = 0; w2 < NW; w2++) all |= flags[parity * NW + w2];
                bad = (all >> (team * TEAM)) & 1u;
                // acceptance in candidate order — every warp replays it identically (same values written)
                for (int t = 0; t < TPW; t++) {
                    const int bad_t = __shfl_sync(0xffffffffu, (int)bad, t * TEAM);
                    if (bad_t) continue;
                    if (team == t) {
                        if (lit == 0) kept_key[K] = key;
                        if (KV) {
#pragma unroll
                            for (int cc = 0; cc < CPL; cc++) {
                                const int chunk = cc * TEAM + lit;
                                if (chunk < nchunk) kvec[(size_t)K * nchunk + chunk] = v.raw[cc];
                            }
                        }
                    }
                    __syncwarp();
                    K++;
                    if (K >= deg) break;
                    if (t + 1 < TPW) {
                        const uint32_t id_t = __shfl_sync(0xffffffffu, id, t * TEAM);
                        const float4* u = KV ? kvec + (size_t)(K - 1) * nchunk : vecs + (size_t)id_t * nchunk;
                        const float duv = v.dist(u, nchunk, lit, is_l2, g.half);
                        if (team > t && duv < dq) bad = true;
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            *nver_ptr(g, b, pt, level) = verified ? (uint8_t)K : (uint8_t)0;
            if (b.build_counters && verified) atomicAdd(b.build_counters + 4, (unsigned long long)n);  // candidate rows read
        }
        // faiss pops link_targets farthest-first: row[i] = kept[K-1-i]
        for (int i = threadIdx.x; i < g.deg0; i += blockDim.x) {
            const int e = item * g.deg0 + i;
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
        __syncthreads();  // kept_key / kvec are reused by the next item
    }
}

// ---- App. A.11 add_link(dst ← src) for every staged back-edge ------------------------
//
// Full rows re-run shrink_neighbor_list on the 2M+1 candidates. Done naively that is ~2000
// pairwise distances per back-link (what faiss does). Here each row carries a *verified prefix*
// `nver`: members at positions [0, nver) came out of one heuristic run, so every pair u ≺ v among
// them already satisfies d(u,v) >= d(v,owner) — a static fact about those vectors. Re-running the
// heuristic on (row ∪ {src}) can therefore only be decided by pairs that involve a *special*
// candidate: src itself or a member appended since the last shrink. The incremental path streams
// the 2M+1 vectors once, scoring each against the owner and against the (few) specials, and then
// replays the heuristic on that small table — the same comparisons on the same values as the full
// run, hence the identical result, at ~1/10 of the arithmetic and no large shared-memory stage.
template <int TEAM, int CPL, bool HALF, int kGR /* candidate rows in flight per team */, int MINB>
__global__ void __launch_bounds__(64, MINB) backlink_kernel(GraphView g, BuildBatch b, int use_kvec) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TPW = 32 / TEAM;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int lit = lane % TEAM, team = lane / TEAM;
    const WarpSmem w = carve_warp_smem(smem_raw + wib * warp_smem_bytes(g.nchunk * 4, g.deg0, false, b.max_special),
                                       g.nchunk * 4, g.deg0, false, b.max_special);
    const float4* __restrict__ vecs = reinterpret_cast<const float4*>(g.vecs);
    const bool is_l2 = g.is_l2 != 0;
    const int stride = g.deg0 + 8;
    int32_t* kspec = w.chain + kChainCap - 32;  // last 32 ints of the chain buffer: kept specials
    const int chain_cap = kChainCap - 32;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const int n_edges = b.n_items * g.deg0;
    (void)use_kvec;
    // one edge slot per warp step (measured: examining 32 slots per step and working through the owners among
    // them serially is slower — the chains are better spread over the warps one by one)
    for (int e0 = blockIdx.x * (blockDim.x >> 5) + wib; e0 < n_edges; e0 += nwarps) {
        const int slot = b.edge_dst_slot[e0];
        if (slot < 0 || b.edge_next[e0] != -1) continue;  // only the tail of a row's list owns it
        // collect the row's pending edges (pushed in arbitrary order); applied in edge-index order
        int c = 0, head = -1;
        if (lane == 0) {
            head = b.slot_head[slot];
            for (int e = head; e >= 0; e = b.edge_next[e]) {
                if (c < chain_cap) w.chain[c] = e;
                c++;
            }
            b.slot_head[slot] = -1;
        }
        c = __shfl_sync(0xffffffffu, c, 0);
        head = __shfl_sync(0xffffffffu, head, 0);
        __syncwarp();
        const bool overflow = c > chain_cap;
        int last_done = -1;
        const int dst = b.edge_dst[e0], level = b.edge_level[e0];
        int deg;
        int32_t* row = row_ptr_rw(g, dst, level, deg);
        uint8_t* nvp = nver_ptr(g, b, dst, level);
        int nv = *nvp;
        const int nv0 = nv;
        TeamVec<TEAM, CPL, HALF> q;  // the destination vertex's own vector (base of the distances)
        bool q_loaded = false;

        for (int done = 0; done < c; done++) {
            // next edge = smallest edge index > last_done
            int e = 0x7fffffff;
            if (!overflow) {
                for (int i = lane; i < c; i += 32) {
                    const int v = w.chain[i];
                    if (v > last_done && v < e) e = v;
                }
            } else if (lane == 0) {  // rare (more pending edges than the buffer holds): re-walk the list
                for (int x = head; x >= 0; x = b.edge_next[x])
                    if (x > last_done && x < e) e = x;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const int o = __shfl_xor_sync(0xffffffffu, e, off);
                e = o < e ? o : e;
            }
            if (e == 0x7fffffff) break;
            last_done = e;
            const int src = b.edge_src[e];
            const float d_src = b.edge_dist[e];

            // current row state
            int ids[kMaxIdsPerLane];
            int last_valid = -1;
#pragma unroll
            for (int i = 0; i < kMaxIdsPerLane; i++) {
                const int idx = lane + 32 * i;
                ids[i] = idx < deg ? __ldcg(row + idx) : -1;
                const unsigned bal = __ballot_sync(0xffffffffu, ids[i] >= 0);
                if (bal) last_valid = 32 * i + 31 - __clz(bal);
            }
            if (last_valid < deg - 1) {  // room: first free slot after the last valid entry
                if (lane == 0) row[last_valid + 1] = src;
                __syncwarp();
                continue;
            }
            // full row: the deg+1 candidates fight it out (shrink_neighbor_list).
            // candidate index a: 0 = src, 1 + p = row member at position p
            if (!q_loaded) {
                q.load(vecs + (size_t)dst * g.nchunk, g.nchunk, lit, true, g.half);
                q_loaded = true;
            }
            const int n = deg + 1;
            if (lane == 0 && b.build_counters) atomicAdd(b.build_counters + 5, (unsigned long long)(n + 1));  // rows streamed (+ owner)
#pragma unroll
            for (int i = 0; i < kMaxIdsPerLane; i++) {
                const int idx = lane + 32 * i;
                if (idx < deg) w.rowid[idx] = ids[i];
            }
            if (nv > deg) nv = deg;
            const int ns = 1 + (deg - nv);  // specials: src + members appended since the last shrink
            const bool incremental = ns <= b.max_special;
            __syncwarp();
            if (incremental) {  // specials' vectors -> shared memory
                for (int s0 = 0; s0 < ns; s0 += TPW) {
                    const int si = s0 + team;
                    if (si < ns) {
                        const int id = si == 0 ? src : w.rowid[nv + si - 1];
                        TeamVec<TEAM, CPL, HALF> tv;
                        tv.load(vecs + (size_t)id * g.nchunk, g.nchunk, lit, true, g.half);
#pragma unroll
                        for (int c2 = 0; c2 < CPL; c2++) {
                            const int chunk = c2 * TEAM + lit;
                            if (chunk < g.nchunk) w.spec_vec[(size_t)si * g.nchunk + chunk] = tv.raw[c2];
                        }
                    }
                }
                __syncwarp();
            }
            // stream every candidate once: distance to the owner (+ to each special)
            for (int r0 = 0; r0 < n; r0 += TPW * kGR) {
                TeamVec<TEAM, CPL, HALF> x[kGR];
                int cid[kGR];
#pragma unroll
                for (int k2 = 0; k2 < kGR; k2++) {
                    const int a = r0 + k2 * TPW + team;
                    cid[k2] = a >= n ? -1 : (a == 0 ? src : w.rowid[a - 1]);
                    const float4* rowv = vecs + (size_t)(cid[k2] < 0 ? 0 : cid[k2]) * g.nchunk;
#pragma unroll
                    for (int c2 = 0; c2 < CPL; c2++) {
                        const int chunk = c2 * TEAM + lit;
                        x[k2].set_raw(c2, (cid[k2] >= 0 && chunk < g.nchunk) ? ldg_stream(rowv + chunk)
                                                                            : make_float4(0.f, 0.f, 0.f, 0.f), g.half);
                    }
                }
#pragma unroll
                for (int k2 = 0; k2 < kGR; k2++) {
                    const int a = r0 + k2 * TPW + team;
                    const float dd = x[k2].dist(q, is_l2);
                    if (cid[k2] >= 0 && lit == 0)
                        w.cand_a[a] = a == 0 ? pack_key(d_src, (uint32_t)src) : pack_key(dd, (uint32_t)cid[k2]);
                    if (incremental) {
                        for (int si = 0; si < ns; si++) {
                            const float ds = x[k2].dist(w.spec_vec + (size_t)si * g.nchunk, g.nchunk, lit, is_l2, g.half);
                            if (cid[k2] >= 0 && lit == 0) w.dsx[si * stride + a] = ds;
                        }
                    }
                }
            }
            __syncwarp();
            // rank sort cand_a[0..n) -> cand_b[0..n), remembering each entry's candidate index
            for (int a = lane; a < n; a += 32) {
                const unsigned long long ka = w.cand_a[a];
                int rk = 0;
                for (int j = 0; j < n; j++) {
                    const unsigned long long kj = w.cand_a[j];
                    rk += (kj < ka) || (kj == ka && j < a);
                }
                w.cand_b[rk] = ka;
                w.slot_b[rk] = a;
            }
            __syncwarp();
            int K = 0;
            if (incremental) {
                int nks = 0;  // kept specials so far
                for (int cpos = 0; cpos < n; cpos++) {
                    const unsigned long long key = w.cand_b[cpos];
                    const int a = w.slot_b[cpos];
                    const float dq = key_dist(key);
                    const int si = a == 0 ? 0 : ((a - 1) >= nv ? 1 + (a - 1 - nv) : -1);
                    bool bad = false;
                    if (si < 0) {  // verified member: only a kept special can prune it
                        if (lane < nks) bad = w.dsx[kspec[lane] * stride + a] < dq;
                    } else {       // special: any kept candidate can prune it
                        for (int j = lane; j < K; j += 32) bad |= w.dsx[si * stride + w.kept_slot[j]] < dq;
                    }
                    if (__any_sync(0xffffffffu, bad)) continue;
                    if (lane == 0) {
                        w.kept_key[K] = key;
                        w.kept_slot[K] = a;
                        if (si >= 0) kspec[nks] = si;
                    }
                    K++;
                    nks += si >= 0;
                    __syncwarp();
                    if (K >= deg) break;
                }
            } else {
                K = heuristic<TEAM, CPL, false, HALF>(g, w.cand_b, n, deg, w.kept_key, nullptr, nullptr, nullptr,
                                                nullptr, lane);
            }
            __syncwarp();
            for (int i = lane; i < deg; i += 32) row[i] = i < K ? (int)key_id(w.kept_key[K - 1 - i]) : -1;
            nv = K;  // the whole row is now one heuristic output
            __syncwarp();
        }
        if (lane == 0 && nv != nv0) *nvp = (uint8_t)nv;
    }
}

template <int TEAM, int CPL, bool HALF>
cudaError_t launch_build_pair(bool backlinks, const GraphView& g, const BuildBatch& b, int num_sms,
                              cudaStream_t stream) {
    const int wpb = 2;
    if (!backlinks) {
        constexpr int NW = 4;
        const size_t kv_bytes = (size_t)g.deg0 * g.nchunk * 16;
        const bool kvec = kv_bytes <= 64 * 1024;  // kept-vector cache per CTA; wider rows re-read kept vectors (L2)
        const size_t smem = (size_t)g.deg0 * 8 + 64 + (kvec ? kv_bytes : 0);
        auto launch = [&](auto ks) -> cudaError_t {
            cudaError_t e2 = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e2 != cudaSuccess) return e2;
            int occ = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ks, 32 * NW, smem);
            long long grid = (long long)num_sms * (occ < 1 ? 1 : occ);
            if (grid > b.n_items) grid = b.n_items;
            if (grid < 1) grid = 1;
            ks<<<(unsigned)grid, 32 * NW, smem, stream>>>(g, b);
            return cudaGetLastError();
        };
        return kvec ? launch(select_and_link_coop_kernel<TEAM, CPL, NW, true, HALF>)
                    : launch(select_and_link_coop_kernel<TEAM, CPL, NW, false, HALF>);
    } else {
        const size_t smem = wpb * warp_smem_bytes(g.nchunk * 4, g.deg0, false, b.max_special);
        auto launch = [&](auto kb) -> cudaError_t {
            cudaError_t e2 = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e2 != cudaSuccess) return e2;
            int occ = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kb, 32 * wpb, smem);
            long long grid = (long long)num_sms * (occ < 1 ? 1 : occ);
            const long long need = ((long long)b.n_items * g.deg0 + wpb - 1) / wpb;
            if (grid > need) grid = need;
            if (grid < 1) grid = 1;
            kb<<<(unsigned)grid, 32 * wpb, smem, stream>>>(g, b, 0);
            return cudaGetLastError();
        };
        // two candidate rows in flight per team, registers capped at 128 (8 blocks/SM): measured 5 % faster
        // over a 1M x 128 build than four rows at 188 registers (5 blocks/SM) — occupancy hides more latency
        // than the deeper prefetch
        return launch(backlink_kernel<TEAM, CPL, HALF, 2, 8>);
    }
    return cudaGetLastError();
}

template <bool HALF>
cudaError_t dispatch_build_t(bool backlinks, const GraphView& g, const BuildBatch& b, int num_sms,
                             cudaStream_t stream) {
    const int nc = g.nchunk;  // same (TEAM, CPL) table as the traversal kernel
    if (nc <= 16) return launch_build_pair<8, 2, HALF>(backlinks, g, b, num_sms, stream);
    if (nc == 24) return launch_build_pair<8, 3, HALF>(backlinks, g, b, num_sms, stream);
    if (nc <= 32) return launch_build_pair<8, 4, HALF>(backlinks, g, b, num_sms, stream);
    if (nc <= 64) return launch_build_pair<16, 4, HALF>(backlinks, g, b, num_sms, stream);
    if (nc <= 128) return launch_build_pair<32, 4, HALF>(backlinks, g, b, num_sms, stream);
    if (nc <= 256) return launch_build_pair<32, 8, HALF>(backlinks, g, b, num_sms, stream);
    if (nc <= 512) return launch_build_pair<32, 16, HALF>(backlinks, g, b, num_sms, stream);
    return cudaErrorInvalidValue;
}

cudaError_t dispatch_build(bool backlinks, const GraphView& g, const BuildBatch& b, int num_sms,
                           cudaStream_t stream) {
    return g.half ? dispatch_build_t<true>(backlinks, g, b, num_sms, stream)
                  : dispatch_build_t<false>(backlinks, g, b, num_sms, stream);
}

}  // namespace

cudaError_t launch_select_and_link(const GraphView& g, const BuildBatch& b, int num_sms, cudaStream_t stream) {
    return dispatch_build(false, g, b, num_sms, stream);
}
cudaError_t launch_backlinks(const GraphView& g, const BuildBatch& b, int num_sms, cudaStream_t stream) {
    return dispatch_build(true, g, b, num_sms, stream);
}

}  // namespace bh
