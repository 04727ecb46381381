// beam_kernel_f16.cu — instantiates the traversal kernel for IEEE fp16 vector storage.
#include "beam_kernel_impl.cuh"

namespace bh {
cudaError_t launch_beam_f16(const GraphView& g, const BeamTask& t, int W, int variant, int num_sms,
                              cudaStream_t stream, int* grid_out, const BuildBatch* fuse) {
    return launch_by_chunks<true>(g, t, W, variant, num_sms, stream, grid_out, fuse);
}
}  // namespace bh
