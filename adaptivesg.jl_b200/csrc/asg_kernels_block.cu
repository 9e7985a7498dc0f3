// asg_kernels_block.cu -- host side of the BLOCK walk (kernel: asg_kernels_block.cuh): shared-memory sizing, choice of
// the variant (one / two points per thread, depth mask), launch.
#include "asg_kernels_block.cuh"

namespace asg {

static int blk_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

size_t block_smem_bytes(int D, int slots, int tile_coefs, int R, int threads) {
  return sizeof(BlkSmem) + 2 * (size_t)tile_coefs * 8 + ((size_t)D * 8 + (size_t)slots * 12) * R * threads;
}
static size_t blk_smem_bytes(const DeviceGrid &g, int R, int threads) {
  return block_smem_bytes(g.D, g.blk_slots, g.blk_tile_coefs, R, threads);
}

static int blk_threads() { return kBlkThreads; }

bool block_kernel_fits(const DeviceGrid &g) { return g.n_btiles > 0 && blk_smem_bytes(g, 1, blk_threads()) <= 227 * 1024; }

extern template cudaError_t blk_launch<1, false, kBlkThreads>(const DeviceGrid &, const EvalArgs &, int, size_t, cudaStream_t);
extern template cudaError_t blk_launch<1, true, kBlkThreads>(const DeviceGrid &, const EvalArgs &, int, size_t, cudaStream_t);
extern template cudaError_t blk_launch<2, false, kBlkThreads>(const DeviceGrid &, const EvalArgs &, int, size_t, cudaStream_t);
extern template cudaError_t blk_launch<2, true, kBlkThreads>(const DeviceGrid &, const EvalArgs &, int, size_t, cudaStream_t);

template <int NT>
static cudaError_t blk_launch_nt(const DeviceGrid &g, const EvalArgs &a, bool two, bool masked, int blocks, size_t smem,
                                 cudaStream_t st) {
  if (two) return masked ? blk_launch<2, true, NT>(g, a, blocks, smem, st) : blk_launch<2, false, NT>(g, a, blocks, smem, st);
  return masked ? blk_launch<1, true, NT>(g, a, blocks, smem, st) : blk_launch<1, false, NT>(g, a, blocks, smem, st);
}

cudaError_t launch_eval_block(const DeviceGrid &g, const EvalArgs &a, cudaStream_t st, int64_t *launches) {
  if (a.M <= 0) return cudaSuccess;
  const int sms = blk_sm_count();
  const int nt = blk_threads();
  // two points per thread as soon as one-point CTAs would need a second wave (and the shared memory fits)
  bool two = a.M > (int64_t)sms * nt && blk_smem_bytes(g, 2, nt) <= 227 * 1024;
  if (const char *e = getenv("ASG_BLOCK_R")) two = (e[0] == '2') && blk_smem_bytes(g, 2, nt) <= 227 * 1024;  // experiment knob
  const int R = two ? 2 : 1;
  const size_t smem = blk_smem_bytes(g, R, nt);
  const int64_t per_block = (int64_t)nt * R;
  const int64_t need = (a.M + per_block - 1) / per_block;
  const int blocks = (int)(need < sms ? need : sms);
  const bool masked = a.rem < INT_MAX / 4;
  const cudaError_t e = blk_launch_nt<kBlkThreads>(g, a, two, masked, blocks, smem, st);
  if (e != cudaSuccess) return e;  // the shared-memory limit of the kernel could not be raised
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace asg
