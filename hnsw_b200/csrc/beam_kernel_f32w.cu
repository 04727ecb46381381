// beam_kernel_f32w.cu — instantiates the traversal kernel for fp32 vector storage, rows wider than 512 B.
#include "beam_launch.cuh"

namespace bh {
cudaError_t launch_beam_f32w(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                 cudaStream_t stream, int* grid_out, const BuildBatch* fuse) {
    return launch_wide<false>(g, t, W, variant, num_sms, stream, grid_out, fuse);
}
}  // namespace bh
