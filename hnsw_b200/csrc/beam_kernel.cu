// beam_kernel.cu — dispatcher of the traversal kernel (see beam_kernel_impl.cuh).
#include "beam.cuh"
#include "engine.h"

namespace bh {

cudaError_t launch_beam_f32(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                            cudaStream_t stream, int* grid_out, const BuildBatch* fuse);
cudaError_t launch_beam_f16(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                            cudaStream_t stream, int* grid_out, const BuildBatch* fuse);

size_t beam_group_smem(int d, int ef, int hash_bits, int deg, int rk) {
    return group_smem_bytes(d, ef, 1 << hash_bits, deg, rk);
}

cudaError_t launch_beam(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                        cudaStream_t stream, int* grid_out, const BuildBatch* fuse) {
    return g.half ? launch_beam_f16(g, t, W, variant, num_sms, stream, grid_out, fuse)
                  : launch_beam_f32(g, t, W, variant, num_sms, stream, grid_out, fuse);
}

}  // namespace bh
