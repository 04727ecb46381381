// common.cuh — shared device/host helpers for the sm_100a HNSW kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bh {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;          // empty visited-hash slot
constexpr unsigned long long kExpanded = 0x80000000ull;  // "already expanded" bit of a list key
constexpr int kMaxDeg = 128;                       // 2*M <= 128
constexpr int kMaxIdsPerLane = kMaxDeg / 32;

// Visited-set policies (BeamTask::visited_mode). The table is `4 << hash_bits` bytes in every mode.
constexpr int kVisitedExact = 0;    // open addressing, 32-bit slots, exact until 3/4 full, then clear + re-seed
constexpr int kVisitedAssoc16 = 1;  // 8-way buckets of 16-bit quotients, FIFO eviction (ntotal <= 2^(hash_bits+14))
constexpr int kVisitedAssoc32 = 2;  // 4-way buckets of 32-bit ids, FIFO eviction

// Device-side view of one index shard (SURVEY §8a1/a2 re-laid-out for HBM):
//   vecs       [ntotal][d] row-major (faiss IndexFlat codes): fp32, or fp16 / bf16 when `half` is set
//              (opt-in storage mode). Either way a row is `nchunk` 16-byte chunks, so row addressing
//              is the same; only a chunk's interpretation differs (4 floats vs 8 halfs).
//   nbr0       int32 [ntotal][deg0]   level-0 rows, deg0 = 2M, -1 terminated
//   upper_base int32 [ntotal]         first upper row of vertex i (in rows of degU), -1 if level 0
//   upper_nbr  int32 [nupper][degU]   rows for levels 1..L of a vertex are consecutive
struct GraphView {
    const float* vecs;
    int32_t* nbr0;
    const int32_t* upper_base;
    int32_t* upper_nbr;
    int d;
    int nchunk;  // 16-byte chunks per stored row: d / 4 (fp32) or d / 8 (fp16)
    int half;    // 16-bit storage: 1 = IEEE fp16, 2 = bfloat16 (0 = fp32)
    int deg0;
    int degU;
    int entry_point;
    int max_level;
    int is_l2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// Order-preserving map float -> uint32 (works for negative values: IP search runs on -dot).
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t b = __float_as_uint(f);
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    uint32_t b = u ^ (((u >> 31) - 1u) | 0x80000000u);
    return __uint_as_float(b);
}
__device__ __forceinline__ unsigned long long pack_key(float dist, uint32_t id) {
    return ((unsigned long long)f2ord(dist) << 32) | id;
}
__device__ __forceinline__ float key_dist(unsigned long long k) { return ord2f((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_id(unsigned long long k) { return (uint32_t)k & 0x7FFFFFFFu; }
__device__ __forceinline__ unsigned long long key_clean(unsigned long long k) { return k & ~kExpanded; }

__device__ __forceinline__ uint32_t hash_id(uint32_t id, int bits) {
    return (id * 2654435761u) >> (32 - bits);
}

// ---- stored chunk -> fp32 --------------------------------------------------------------
// ES = float4s per 16-byte stored chunk (1: fp32 storage, 2: 16-bit storage, exact widening).
// fmt (16-bit storage only; uniform per index): 1 = IEEE fp16, 2 = bfloat16 (GraphView::half).
template <bool HALF>
__device__ __forceinline__ void chunk_to_f32(const float4& raw, float4 (&out)[HALF ? 2 : 1], int fmt = 1) {
    if constexpr (!HALF) {
        out[0] = raw;
    } else {
        if (fmt == 2) {  // bf16 -> fp32 is the top half of the fp32 pattern
            const uint32_t w0 = __float_as_uint(raw.x), w1 = __float_as_uint(raw.y);
            const uint32_t w2 = __float_as_uint(raw.z), w3 = __float_as_uint(raw.w);
            out[0] = make_float4(__uint_as_float(w0 << 16), __uint_as_float(w0 & 0xFFFF0000u),
                                 __uint_as_float(w1 << 16), __uint_as_float(w1 & 0xFFFF0000u));
            out[1] = make_float4(__uint_as_float(w2 << 16), __uint_as_float(w2 & 0xFFFF0000u),
                                 __uint_as_float(w3 << 16), __uint_as_float(w3 & 0xFFFF0000u));
        } else {
            const __half2* h = reinterpret_cast<const __half2*>(&raw);
            const float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
            const float2 c = __half22float2(h[2]), d = __half22float2(h[3]);
            out[0] = make_float4(a.x, a.y, b.x, b.y);
            out[1] = make_float4(c.x, c.y, d.x, d.y);
        }
    }
}
// acc += sum over the 4 lanes of (a-b)^2 (L2) or a*b (IP), one fmaf chain, fixed order x,y,z,w.
__device__ __forceinline__ void acc4(float& acc, const float4& a, const float4& b, bool l2) {
    if (l2) {
        float t;
        t = a.x - b.x; acc = fmaf(t, t, acc);
        t = a.y - b.y; acc = fmaf(t, t, acc);
        t = a.z - b.z; acc = fmaf(t, t, acc);
        t = a.w - b.w; acc = fmaf(t, t, acc);
    } else {
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc);
        acc = fmaf(a.w, b.w, acc);
    }
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk → UBLKCP in SASS) -------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!ok);
}

// Named barrier over the `nthreads` threads of one query group (ids 1..15; 0 = __syncthreads).
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One volatile 128-bit shared-memory load (a 4-slot hash bucket).
__device__ __forceinline__ uint4 lds128_volatile(const void* smem_ptr) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(smem_u32(smem_ptr))
                 : "memory");
    return v;
}

__device__ __forceinline__ void sts128_volatile(void* smem_ptr, const uint4& v) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(smem_ptr)), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// 128-bit read-only gather load that does not allocate in L1 (rows are not reused by the SM).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// Same, with a compile-time byte offset folded into the address (saves the 64-bit add per load).
template <int OFF>
__device__ __forceinline__ float4 ldg_stream_off(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4+%5];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "n"(OFF));
    return r;
}

}  // namespace bh
