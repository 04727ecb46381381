// beam_kernel_f16.cu — instantiates the traversal kernel for 16-bit (fp16 / bf16) vector storage, rows up to 512 B.
#include "beam_launch.cuh"

namespace bh {
cudaError_t launch_beam_f16(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                 cudaStream_t stream, int* grid_out, const BuildBatch* fuse) {
    return launch_narrow<true>(g, t, W, variant, num_sms, stream, grid_out, fuse);
}
}  // namespace bh
