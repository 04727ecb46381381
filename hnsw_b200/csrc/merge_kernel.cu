// merge_kernel.cu — warp-per-query merge of per-shard sorted top-k lists.
//
// Replaces faiss merge_knn_results as used by IndexShards(successive_ids=true)
// (SURVEY.md §8e): after an all-gather every rank holds [nshard][nq][k] (distance, local id)
// lists, each sorted best-first. Each element's final rank is the number of elements, over all
// shard lists, that precede it (binary search per list; ties broken by shard then position), so
// the merge is a scatter with no serial heap.
#include <algorithm>

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

__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __float2half_rn(src[i]);
}
__global__ void f16_to_f32_kernel(const __half* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __half2float(src[i]);
}

}  // namespace

cudaError_t launch_f32_to_f16(const float* src, void* dst, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 32);
    f32_to_f16_kernel<<<grid, 256, 0, stream>>>(src, static_cast<__half*>(dst), n);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_f16_to_f32(const void* src, float* dst, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 32);
    f16_to_f32_kernel<<<grid, 256, 0, stream>>>(static_cast<const __half*>(src), dst, n);
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
