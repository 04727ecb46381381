// beam_kernel.cu — dispatcher of the traversal kernel (see beam_kernel_impl.cuh).
#include "beam.cuh"
#include "engine.h"

namespace bh {

#define BH_DECL(name)                                                                                   \
    cudaError_t name(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms, cudaStream_t stream, \
                     int* grid_out, const BuildBatch* fuse)
BH_DECL(launch_beam_f32);   // fp32 rows up to 512 B
BH_DECL(launch_beam_f32w);  // fp32 rows above 512 B
BH_DECL(launch_beam_f16);   // 16-bit rows up to 512 B
BH_DECL(launch_beam_f16w);  // 16-bit rows above 512 B
#undef BH_DECL

size_t beam_group_smem(int d, int ef, int hash_bits, int deg, int rk) {
    return group_smem_bytes(d, ef, 1 << hash_bits, deg, rk);
}

cudaError_t launch_beam(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                        cudaStream_t stream, int* grid_out, const BuildBatch* fuse) {
    const bool wide = g.nchunk > 32;
    if (g.half)
        return wide ? launch_beam_f16w(g, t, W, variant, num_sms, stream, grid_out, fuse)
                    : launch_beam_f16(g, t, W, variant, num_sms, stream, grid_out, fuse);
    return wide ? launch_beam_f32w(g, t, W, variant, num_sms, stream, grid_out, fuse)
                : launch_beam_f32(g, t, W, variant, num_sms, stream, grid_out, fuse);
}

}  // namespace bh
