// asg_kernels_block_r2mw.cu -- the BLOCK walk with 2 points per thread, depth-masked, WIDE (suffix budgets 8 and 9 compiled in) (see asg_kernels_block.cuh)
#include "asg_kernels_block.cuh"

namespace asg {
template cudaError_t blk_launch<2, true, kBlkThreads, false, true>(const DeviceGrid &, const EvalArgs &, int, size_t, cudaStream_t);
}  // namespace asg
