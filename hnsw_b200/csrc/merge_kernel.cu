// merge_kernel.cu — warp-per-query merge of per-shard sorted top-k lists.
//
// Replaces faiss merge_knn_results as used by IndexShards(successive_ids=true)
// (SURVEY.md §8e): after an all-gather every rank holds [nshard][nq][k] (distance, local id)
// lists, each sorted best-first. Each element's final rank is the number of elements, over all
// shard lists, that precede it (binary search per list; ties broken by shard then position), so
// the merge is a scatter with no serial heap.
#include <algorithm>
#include <cfloat>

#include "engine.h"

namespace bh {

namespace {

__global__ void __launch_bounds__(128) merge_topk_kernel(int nshard, int64_t nq, int k, int is_l2,
                                                         const float* __restrict__ D_all,
                                                         const int64_t* __restrict__ I_all,
                                                         const ShardOffsets id_off,
                                                         float* __restrict__ D_out,
                                                         int64_t* __restrict__ I_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q = warp; q < nq; q += nwarps) {
        const int total = nshard * k;
        for (int e = lane; e < total; e += 32) {
            const int s = e / k, i = e - s * k;
            const float v = D_all[((size_t)s * nq + q) * k + i];
            int rank = i;
            for (int s2 = 0; s2 < nshard; s2++) {
                if (s2 == s) continue;
                const float* L = D_all + ((size_t)s2 * nq + q) * k;
                // number of entries of list s2 that come before (v, s): strictly better, or equal
                // and from a lower shard
                int lo = 0, hi = k;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const float u = L[mid];
                    const bool before = is_l2 ? (u < v || (u == v && s2 < s)) : (u > v || (u == v && s2 < s));
                    if (before) lo = mid + 1; else hi = mid;
                }
                rank += lo;
            }
            if (rank < k) {
                const int64_t id = I_all[((size_t)s * nq + q) * k + i];
                D_out[(size_t)q * k + rank] = v;
                I_out[(size_t)q * k + rank] = id >= 0 ? id + id_off.v[s] : -1;
            }
        }
    }
}

// ---- packed exchange ------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Runs after the traversal kernel in stream order, so that kernel's peer stores are complete: tell every
// rank that this rank's lists for `epoch` are in its gather buffer.
__global__ void shard_signal_kernel(int nshard, int my_rank, PeerFlags pf, unsigned long long epoch) {
    const int p = threadIdx.x;
    if (p < nshard) {
        __threadfence_system();
        st_release_sys(pf.v[p] + (size_t)my_rank * kFlagStride, epoch);
    }
}

__global__ void __launch_bounds__(128) merge_packed_kernel(int nshard, int64_t nq, int k, int is_l2,
                                                           const unsigned long long* __restrict__ G,
                                                           const ShardOffsets id_off, float* __restrict__ D_out,
                                                           int64_t* __restrict__ I_out,
                                                           const unsigned long long* flags,
                                                           unsigned long long epoch, int* status, int timeout_ms) {
    if (flags) {  // every block waits for every rank's lists (flags arrive over NVLink)
        __shared__ int timed_out;
        if (threadIdx.x == 0) timed_out = 0;
        __syncthreads();
        if ((int)threadIdx.x < nshard) {
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_sys(flags + (size_t)threadIdx.x * kFlagStride) < epoch) {
                __nanosleep(200);
                if (global_timer_ns() - t0 > (unsigned long long)timeout_ms * 1000000ull) {
                    timed_out = 1;  // a peer never published: do not hang the GPU, report instead
                    break;
                }
            }
        }
        __syncthreads();
        if (timed_out) {
            if (threadIdx.x == 0 && status) *status = 1;
            return;
        }
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q = warp; q < nq; q += nwarps) {
        const int total = nshard * k;
        for (int e = lane; e < total; e += 32) {
            const int s = e / k, i = e - s * k;
            const unsigned long long key = G[((size_t)s * nq + q) * k + i];
            const uint32_t v = (uint32_t)(key >> 32);  // order-preserving bits of the search-space distance
            int rank = i;
            for (int s2 = 0; s2 < nshard; s2++) {
                if (s2 == s) continue;
                const unsigned long long* L = G + ((size_t)s2 * nq + q) * k;
                int lo = 0, hi = k;  // entries of list s2 before (v, s): smaller, or equal from a lower shard
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const uint32_t u = (uint32_t)(L[mid] >> 32);
                    if (u < v || (u == v && s2 < s)) lo = mid + 1; else hi = mid;
                }
                rank += lo;
            }
            if (rank < k) {
                const bool empty = key == ~0ull;
                float dd = is_l2 ? FLT_MAX : -FLT_MAX;
                int64_t id = -1;
                if (!empty) {
                    dd = ord2f(v);
                    if (!is_l2) dd = -dd;
                    id = (int64_t)(uint32_t)key + id_off.v[s];
                }
                D_out[(size_t)q * k + rank] = dd;
                I_out[(size_t)q * k + rank] = id;
            }
        }
    }
}

__global__ void unpack_kernel(const unsigned long long* __restrict__ keys, int64_t n, int is_l2,
                              float* __restrict__ D, int64_t* __restrict__ I) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (key == ~0ull) {
            D[i] = is_l2 ? FLT_MAX : -FLT_MAX;
            I[i] = -1;
        } else {
            const float dd = ord2f((uint32_t)(key >> 32));
            D[i] = is_l2 ? dd : -dd;
            I[i] = (int64_t)(uint32_t)key;
        }
    }
}

__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __float2half_rn(src[i]);
}
// fp32 -> bfloat16, round to nearest even (NaN kept quiet); bfloat16 -> fp32 is exact
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t u = __float_as_uint(src[i]);
        uint16_t r;
        if ((u & 0x7FFFFFFFu) > 0x7F800000u)
            r = (uint16_t)((u >> 16) | 0x0040u);
        else
            r = (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
        dst[i] = r;
    }
}
__global__ void bf16_to_f32_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __uint_as_float((uint32_t)src[i] << 16);
}
__global__ void f16_to_f32_kernel(const __half* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __half2float(src[i]);
}

}  // namespace

cudaError_t launch_f32_to_f16(const float* src, void* dst, size_t n, cudaStream_t stream, int fmt) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 32);
    if (fmt == 2)
        f32_to_bf16_kernel<<<grid, 256, 0, stream>>>(src, static_cast<uint16_t*>(dst), n);
    else
        f32_to_f16_kernel<<<grid, 256, 0, stream>>>(src, static_cast<__half*>(dst), n);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_f16_to_f32(const void* src, float* dst, size_t n, cudaStream_t stream, int fmt) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 32);
    if (fmt == 2)
        bf16_to_f32_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(src), dst, n);
    else
        f16_to_f32_kernel<<<grid, 256, 0, stream>>>(static_cast<const __half*>(src), dst, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_shard_signal(int nshard, int my_rank, const PeerFlags& peer_flags, unsigned long long epoch,
                                cudaStream_t stream) {
    shard_signal_kernel<<<1, 32, 0, stream>>>(nshard, my_rank, peer_flags, epoch);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_merge_packed(int nshard, int64_t nq, int k, int is_l2, const unsigned long long* gather,
                                const ShardOffsets& id_offsets, float* D_out, int64_t* I_out,
                                const unsigned long long* flags, unsigned long long epoch, int* status,
                                int timeout_ms, cudaStream_t stream) {
    if (nq == 0) return cudaSuccess;
    const int wpb = 4;
    long long grid = (nq + wpb - 1) / wpb;
    if (grid > 148 * 8) grid = 148 * 8;
    merge_packed_kernel<<<(unsigned)grid, 32 * wpb, 0, stream>>>(nshard, nq, k, is_l2, gather, id_offsets, D_out,
                                                                I_out, flags, epoch, status, timeout_ms);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_unpack(const unsigned long long* keys, int64_t n, int is_l2, float* D, int64_t* I,
                          cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    unpack_kernel<<<grid, 256, 0, stream>>>(keys, n, is_l2, D, I);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_merge_topk(int nshard, int64_t nq, int k, int is_l2, const float* D_all,
                              const int64_t* I_all, const ShardOffsets& id_offsets, float* D_out,
                              int64_t* I_out, cudaStream_t stream) {
    if (nq == 0) return cudaSuccess;
    const int wpb = 4;
    long long grid = (nq + wpb - 1) / wpb;
    if (grid > 148 * 16) grid = 148 * 16;
    merge_topk_kernel<<<(unsigned)grid, 32 * wpb, 0, stream>>>(nshard, nq, k, is_l2, D_all, I_all,
                                                              id_offsets, D_out, I_out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace bh
