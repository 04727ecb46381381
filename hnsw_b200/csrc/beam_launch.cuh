#pragma once
// beam_launch.cuh — host-side launchers of beam_kernel (beam_kernel_impl.cuh): occupancy-sized persistent grid,
// register / residency variants, and the (TEAM, CPL) table by row width. Included by the four beam_kernel_*.cu
// translation units; each instantiates only its share of the table.
#include "beam_kernel_impl.cuh"

namespace bh {

namespace {

template <int TEAM, int CPL, int W, int R, int G, int MINB, bool HALF, bool FUSE, int LEAN = 0>
cudaError_t launch_one_t(const GraphView& g, const BeamTask& t, int num_sms, cudaStream_t stream,
                         int* grid_out, const BuildBatch& b) {
    auto kern = beam_kernel<TEAM, CPL, W, R, G, MINB, HALF, FUSE, LEAN>;
    const size_t smem = (size_t)G * group_smem_bytes(g.d, t.ef, 1 << t.hash_bits, g.deg0, t.sel ? t.k : 0);
    // the dynamic-shared-memory limit is an attribute of the FUNCTION: concurrent searches with different
    // efSearch would race between setting it and launching, so the pair is one critical section
    static std::mutex launch_mu;
    std::lock_guard<std::mutex> lk(launch_mu);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W * G, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorInvalidConfiguration;
    long long groups_needed = ((long long)t.n_items + G - 1) / G;
    long long grid = (long long)num_sms * occ;
    if (grid > groups_needed) grid = groups_needed;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    if (t.pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(32 * W * G);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, g, t, b);
    }
    kern<<<(unsigned)grid, 32 * W * G, smem, stream>>>(g, t, b);
    return cudaGetLastError();
}

template <int TEAM, int CPL, int W, int R, int G, int MINB, bool HALF>
cudaError_t launch_one(const GraphView& g, const BeamTask& t, int num_sms, cudaStream_t stream,
                       int* grid_out, const BuildBatch* fuse) {
    // The fused selection is instantiated for wide rows only (TEAM >= 16, i.e. rows above 512 B): measured at
    // 300k x 768 it makes the build 9 % faster (the selection of a 3 KB-row item is long enough to be worth
    // hiding under the gathers), at 1M x 128 it makes it 7 % slower (registers, longer drain of every round).
    if constexpr (TEAM >= 16) {
        if (fuse) return launch_one_t<TEAM, CPL, W, R, G, MINB, HALF, true>(g, t, num_sms, stream, grid_out, *fuse);
    }
    if (fuse) return cudaErrorInvalidValue;
    if constexpr (W == 1 && MINB == 6) {  // the throughput variant of the search path gets the lean instantiation
        if (!t.items && !t.sel && t.visited_mode == kVisitedAssoc16)
            return launch_one_t<TEAM, CPL, W, R, G, MINB, HALF, false, 1>(g, t, num_sms, stream, grid_out, BuildBatch{});
        if (t.items && !t.sel && t.visited_mode == kVisitedAssoc16)  // (construction: 1.42 -> 1.37 s per 1M x 128 build)
            return launch_one_t<TEAM, CPL, W, R, G, MINB, HALF, false, 2>(g, t, num_sms, stream, grid_out, BuildBatch{});
    }
    return launch_one_t<TEAM, CPL, W, R, G, MINB, HALF, false>(g, t, num_sms, stream, grid_out, BuildBatch{});
}

// variant (W == 1 only): 0 = R rows in flight per team, 4 blocks/SM (<=128 regs);
// 1 = R/2 rows, 6 blocks/SM (<=80 regs); 2 = R/2 rows, 8 blocks/SM (<=64 regs);
// 3 = R/2 rows, 5 blocks/SM (<=96 regs).
template <int TEAM, int CPL, int R, bool HALF>
cudaError_t launch_w(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                     cudaStream_t stream, int* grid_out, const BuildBatch* fuse) {
    constexpr int RH = R >= 2 ? R / 2 : 1;
    switch (W) {
        case 1:
            if (variant == 1) return launch_one<TEAM, CPL, 1, RH, 4, 6, HALF>(g, t, num_sms, stream, grid_out, fuse);
            if (variant == 2) return launch_one<TEAM, CPL, 1, RH, 4, 8, HALF>(g, t, num_sms, stream, grid_out, fuse);
            if (variant == 3) return launch_one<TEAM, CPL, 1, RH, 4, 5, HALF>(g, t, num_sms, stream, grid_out, fuse);
            return launch_one<TEAM, CPL, 1, R, 4, 4, HALF>(g, t, num_sms, stream, grid_out, fuse);
        case 2: return launch_one<TEAM, CPL, 2, R, 2, 1, HALF>(g, t, num_sms, stream, grid_out, fuse);
        case 4: return launch_one<TEAM, CPL, 4, R, 1, 1, HALF>(g, t, num_sms, stream, grid_out, fuse);
        case 8: return launch_one<TEAM, CPL, 8, R, 1, 1, HALF>(g, t, num_sms, stream, grid_out, fuse);
        default: return cudaErrorInvalidValue;
    }
}

// (TEAM, CPL) by the number of 16-byte chunks per stored row (fp32: d/4, 16-bit storage: d/8).
// Rows up to 32 chunks (512 B): 8 lanes per row.
template <bool HALF>
cudaError_t launch_narrow(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                          cudaStream_t stream, int* grid_out, const BuildBatch* fuse) {
    const int nc = g.nchunk;
    if (nc <= 16) return launch_w<8, 2, 8, HALF>(g, t, W, variant, num_sms, stream, grid_out, fuse);  // 2 chunks/lane: 8 rows in flight
    if (nc == 24) return launch_w<8, 3, 4, HALF>(g, t, W, variant, num_sms, stream, grid_out, fuse);  // d=96 fp32: exact fit
    if (nc <= 32) return launch_w<8, 4, 4, HALF>(g, t, W, variant, num_sms, stream, grid_out, fuse);
    return cudaErrorInvalidValue;
}
// Wider rows: 16 lanes up to 64 chunks, the whole warp above.
template <bool HALF>
cudaError_t launch_wide(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                        cudaStream_t stream, int* grid_out, const BuildBatch* fuse) {
    const int nc = g.nchunk;
    if (nc <= 32) return cudaErrorInvalidValue;
    if (nc <= 64) return launch_w<16, 4, 4, HALF>(g, t, W, variant, num_sms, stream, grid_out, fuse);
    if (nc <= 128) return launch_w<32, 4, 4, HALF>(g, t, W, variant, num_sms, stream, grid_out, fuse);
    if (nc <= 256) return launch_w<32, 8, 2, HALF>(g, t, W, variant, num_sms, stream, grid_out, fuse);
    if (nc <= 512) return launch_w<32, 16, 1, HALF>(g, t, W, variant, num_sms, stream, grid_out, fuse);
    return cudaErrorInvalidValue;
}

}  // namespace

}  // namespace bh
