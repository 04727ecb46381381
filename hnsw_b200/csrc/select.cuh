// select.cuh — device pieces shared by the construction kernels (build_kernels.cu) and by the traversal
// kernel's fused selection epilogue (beam_kernel_impl.cuh): a stored vector sliced over a team of lanes, the
// shrink_neighbor_list heuristic and the row / slot addressing helpers.
#pragma once
#include "beam.cuh"
#include "engine.h"

namespace bh {

// One stored vector, sliced over the TEAM lanes of a team: lane `lit` owns the 16-byte chunks
// lit, lit+TEAM, ... `raw` keeps them as stored (fp32 x4 or fp16 x8; used to stage vectors in
// shared memory), `f` is their exact fp32 widening used for arithmetic.
template <int TEAM, int CPL, bool HALF>
struct TeamVec {
    static constexpr int ES = HALF ? 2 : 1;
    float4 raw[CPL];
    float4 f[CPL][ES];
    __device__ __forceinline__ void set_raw(int c, const float4& v, int fmt) {
        raw[c] = v;
        chunk_to_f32<HALF>(v, f[c], fmt);
    }
    // Load this lane's slice of a stored vector (generic pointer: global or shared).
    __device__ __forceinline__ void load(const float4* row, int nchunk, int lit, bool valid, int fmt) {
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const int chunk = c * TEAM + lit;
            set_raw(c, (valid && chunk < nchunk) ? row[chunk] : make_float4(0.f, 0.f, 0.f, 0.f), fmt);
        }
    }
    __device__ __forceinline__ float reduce(float acc, bool is_l2) const {
#pragma unroll
        for (int off = TEAM / 2; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);
        return is_l2 ? acc : -acc;
    }
    // Same arithmetic as Beam::compute_dists: one fmaf chain per lane, xor-butterfly over the team.
    // (rows of exactly TEAM * CPL chunks — d = 128, 96, ... — take the predicate-free path)
    __device__ __forceinline__ float dist(const float4* row, int nchunk, int lit, bool is_l2, int fmt) const {
        return nchunk == TEAM * CPL ? dist_t<true>(row, nchunk, lit, is_l2, fmt) : dist_t<false>(row, nchunk, lit, is_l2, fmt);
    }
    template <bool FULL>
    __device__ __forceinline__ float dist_t(const float4* row, int nchunk, int lit, bool is_l2, int fmt) const {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const int chunk = c * TEAM + lit;
            const float4 u = (FULL || chunk < nchunk) ? row[chunk] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 uf[ES];
            chunk_to_f32<HALF>(u, uf, fmt);
#pragma unroll
            for (int e = 0; e < ES; e++) acc4(acc, uf[e], f[c][e], is_l2);
        }
        return reduce(acc, is_l2);
    }
    // Distance to a vector held in another TeamVec (same arithmetic).
    __device__ __forceinline__ float dist(const TeamVec& o, bool is_l2) const {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; c++)
#pragma unroll
            for (int e = 0; e < ES; e++) acc4(acc, o.f[c][e], f[c][e], is_l2);
        return reduce(acc, is_l2);
    }
    // Four independent pairs at once (same per-pair arithmetic; the four fmaf chains interleave).
    __device__ __forceinline__ void dist4(const float4* r0, const float4* r1, const float4* r2,
                                          const float4* r3, int nchunk, int lit, bool is_l2,
                                          float (&out)[4], int fmt) const {
        if (nchunk == TEAM * CPL)
            dist4_t<true>(r0, r1, r2, r3, nchunk, lit, is_l2, out, fmt);
        else
            dist4_t<false>(r0, r1, r2, r3, nchunk, lit, is_l2, out, fmt);
    }
    template <bool FULL>
    __device__ __forceinline__ void dist4_t(const float4* r0, const float4* r1, const float4* r2,
                                            const float4* r3, int nchunk, int lit, bool is_l2,
                                            float (&out)[4], int fmt) const {
        const float4* rows[4] = {r0, r1, r2, r3};
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const int chunk = c * TEAM + lit;
            float4 u[4];
#pragma unroll
            for (int p = 0; p < 4; p++)
                u[p] = (FULL || chunk < nchunk) ? rows[p][chunk] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int p = 0; p < 4; p++) {
                float4 uf[ES];
                chunk_to_f32<HALF>(u[p], uf, fmt);
#pragma unroll
                for (int e = 0; e < ES; e++) acc4(acc[p], uf[e], f[c][e], is_l2);
            }
        }
#pragma unroll
        for (int off = TEAM / 2; off >= 1; off >>= 1) {
#pragma unroll
            for (int p = 0; p < 4; p++) acc[p] = acc[p] + __shfl_xor_sync(0xffffffffu, acc[p], off);
        }
#pragma unroll
        for (int p = 0; p < 4; p++) out[p] = is_l2 ? acc[p] : -acc[p];
    }
};

// App. A.10 — keep candidate v (nearest first) iff no already-kept u has d(u,v) < d(v,base).
// cand: sorted clean keys (generic pointer), n >= 1. kept_key: shared, capacity >= max_size.
// Two vector sources:
//   STAGED = false: candidate vectors come from HBM/L2 (the next group of 32/TEAM candidates is
//                   prefetched into registers while the current group is tested); kept vectors are
//                   cached in shared memory `kvec` ([max_size][nchunk] float4) when it is non-null.
//   STAGED = true : every candidate vector already sits in shared memory `stage` at slot
//                   cand_slot[c]; kept vectors are read back from their slots (kept_slot[]).
// 32/TEAM candidates are examined per step, one per team; dependencies inside a step are
// resolved in candidate order, so the outcome equals the sequential scan.
template <int TEAM, int CPL, bool STAGED, bool HALF>
__device__ int heuristic(const GraphView& g, const unsigned long long* cand, int n, int max_size,
                         unsigned long long* kept_key, float4* kvec, const float4* stage,
                         const int32_t* cand_slot, int32_t* kept_slot, int lane) {
    constexpr int TPW = 32 / TEAM;
    const int lit = lane % TEAM, team = lane / TEAM;
    const float4* __restrict__ vecs = reinterpret_cast<const float4*>(g.vecs);
    const bool is_l2 = g.is_l2 != 0;
    int K = 0;
    TeamVec<TEAM, CPL, HALF> nxt;
    unsigned long long nxt_key = ~0ull;
    int nxt_slot = 0;
    auto fetch = [&](int c0) {
        const int c = c0 + team;
        const bool valid = c < n;
        nxt_key = valid ? cand[c] : ~0ull;
        if (STAGED) {
            nxt_slot = valid ? cand_slot[c] : 0;
            nxt.load(stage + (size_t)nxt_slot * g.nchunk, g.nchunk, lit, valid, g.half);
        } else {
            nxt.load(vecs + (size_t)(valid ? key_id(nxt_key) : 0) * g.nchunk, g.nchunk, lit, valid, g.half);
        }
    };
    fetch(0);
    for (int c0 = 0; c0 < n && K < max_size; c0 += TPW) {
        const TeamVec<TEAM, CPL, HALF> v = nxt;
        const unsigned long long key = nxt_key;
        const int slot = nxt_slot;
        const bool valid = c0 + team < n;
        if (c0 + TPW < n) fetch(c0 + TPW);  // in flight while this group is tested
        const uint32_t id = key_id(key);
        const float dq = key_dist(key);
        bool bad = !valid;
        auto kept_row = [&](int j) -> const float4* {
            return STAGED ? stage + (size_t)kept_slot[j] * g.nchunk
                          : (kvec ? kvec + (size_t)j * g.nchunk
                                  : vecs + (size_t)key_id(kept_key[j]) * g.nchunk);
        };
        int j = 0;
        for (; j + 4 <= K; j += 4) {  // four kept vectors per iteration: independent fmaf chains
            if (__all_sync(0xffffffffu, bad)) break;
            float duv[4];
            v.dist4(kept_row(j), kept_row(j + 1), kept_row(j + 2), kept_row(j + 3), g.nchunk, lit, is_l2, duv, g.half);
            if (duv[0] < dq || duv[1] < dq || duv[2] < dq || duv[3] < dq) bad = true;
        }
        for (; j < K; j++) {
            if (__all_sync(0xffffffffu, bad)) break;
            const float duv = v.dist(kept_row(j), g.nchunk, lit, is_l2, g.half);
            if (duv < dq) bad = true;
        }
        for (int t = 0; t < TPW; t++) {
            const int bad_t = __shfl_sync(0xffffffffu, (int)bad, t * TEAM);
            if (bad_t) continue;
            if (team == t) {
                if (lit == 0) {
                    kept_key[K] = key;
                    if (STAGED) kept_slot[K] = slot;
                }
                if (!STAGED && kvec) {
#pragma unroll
                    for (int cc = 0; cc < CPL; cc++) {
                        const int chunk = cc * TEAM + lit;
                        if (chunk < g.nchunk) kvec[(size_t)K * g.nchunk + chunk] = v.raw[cc];
                    }
                }
            }
            __syncwarp();
            K++;
            if (K >= max_size) break;
            if (t + 1 < TPW) {
                const uint32_t id_t = __shfl_sync(0xffffffffu, id, t * TEAM);
                const int slot_t = __shfl_sync(0xffffffffu, slot, t * TEAM);
                const float4* u = STAGED ? stage + (size_t)slot_t * g.nchunk
                                         : (kvec ? kvec + (size_t)(K - 1) * g.nchunk
                                                 : vecs + (size_t)id_t * g.nchunk);
                const float duv = v.dist(u, g.nchunk, lit, is_l2, g.half);
                if (team > t && duv < dq) bad = true;
            }
        }
    }
    return K;
}

__device__ __forceinline__ int32_t* row_ptr_rw(const GraphView& g, int v, int level, int& deg) {
    if (level == 0) {
        deg = g.deg0;
        return g.nbr0 + (size_t)v * g.deg0;
    }
    deg = g.degU;
    const int b = __ldg(g.upper_base + v);
    return g.upper_nbr + ((size_t)b + (level - 1)) * g.degU;
}

__device__ __forceinline__ uint8_t* nver_ptr(const GraphView& g, const BuildBatch& b, int v, int level) {
    return level == 0 ? b.nver0 + v : b.nverU + (__ldg(g.upper_base + v) + (level - 1));
}

__device__ __forceinline__ int row_slot(const GraphView& g, int64_t n_level0, int v, int level) {
    return level == 0 ? v : (int)(n_level0 + __ldg(g.upper_base + v) + (level - 1));
}

}  // namespace bh
