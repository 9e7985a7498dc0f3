// asg_kernels_block_r2.cu -- the BLOCK walk with 2 points per thread, unmasked (see asg_kernels_block.cuh)
#include "asg_kernels_block.cuh"

namespace asg {
template cudaError_t blk_launch<2, false, kBlkThreads, false, false>(const DeviceGrid &, const EvalArgs &, int, size_t, cudaStream_t);
}  // namespace asg
