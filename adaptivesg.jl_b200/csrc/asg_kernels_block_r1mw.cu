// asg_kernels_block_r1mw.cu -- the BLOCK walk with 1 point per thread, depth-masked, WIDE (suffix budgets 8 and 9 compiled in) (see asg_kernels_block.cuh)
#include "asg_kernels_block.cuh"

namespace asg {
template cudaError_t blk_launch<1, true, kBlkThreads, false, true>(const DeviceGrid &, const EvalArgs &, int, size_t, cudaStream_t);
}  // namespace asg
