// beam_kernel_f32.cu — instantiates the traversal kernel for fp32 vector storage, rows up to 512 B.
#include "beam_launch.cuh"

namespace bh {
cudaError_t launch_beam_f32(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                 cudaStream_t stream, int* grid_out, const BuildBatch* fuse) {
    return launch_narrow<false>(g, t, W, variant, num_sms, stream, grid_out, fuse);
}
}  // namespace bh
