// asg_kernels_block_r2pm.cu -- the BLOCK walk with 2 points per thread, depth-masked, PAIR mode (neighbours of the cell-key order share their coefficient loads) (see asg_kernels_block.cuh)
#include "asg_kernels_block.cuh"

namespace asg {
template cudaError_t blk_launch<2, true, kBlkThreads, true, false>(const DeviceGrid &, const EvalArgs &, int, size_t, cudaStream_t);
}  // namespace asg
