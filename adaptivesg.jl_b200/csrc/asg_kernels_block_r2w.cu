// asg_kernels_block_r2w.cu -- the BLOCK walk with 2 points per thread, unmasked, WIDE (suffix budgets 8 and 9 compiled in) (see asg_kernels_block.cuh)
#include "asg_kernels_block.cuh"

namespace asg {
template cudaError_t blk_launch<2, false, kBlkThreads, false, true>(const DeviceGrid &, const EvalArgs &, int, size_t, cudaStream_t);
}  // namespace asg
