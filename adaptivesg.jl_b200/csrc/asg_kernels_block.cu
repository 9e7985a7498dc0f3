// asg_kernels_block.cu -- host side of the BLOCK walk (kernel: asg_kernels_block.cuh): shared-memory sizing, choice of
// the variant (one / two points per thread, depth mask), launch.
#include <cmath>

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

#define ASG_BLK_EXTERN(R, M, P, W) \
  extern template cudaError_t blk_launch<R, M, kBlkThreads, P, W>(const DeviceGrid &, const EvalArgs &, int, size_t, cudaStream_t);
ASG_BLK_EXTERN(1, false, false, false) ASG_BLK_EXTERN(1, true, false, false) ASG_BLK_EXTERN(2, false, false, false)
ASG_BLK_EXTERN(2, true, false, false) ASG_BLK_EXTERN(2, false, true, false) ASG_BLK_EXTERN(2, true, true, false)
ASG_BLK_EXTERN(1, false, false, true) ASG_BLK_EXTERN(1, true, false, true) ASG_BLK_EXTERN(2, false, false, true)
ASG_BLK_EXTERN(2, true, false, true)
#undef ASG_BLK_EXTERN

template <int NT>
static cudaError_t blk_launch_nt(const DeviceGrid &g, const EvalArgs &a, bool two, bool masked, bool pair, int blocks, size_t smem,
                                 cudaStream_t st) {
  if (g.blk_wide) {  // suffix budgets 8 / 9: the WIDE instantiations (never paired)
    if (two) return masked ? blk_launch<2, true, NT, false, true>(g, a, blocks, smem, st) : blk_launch<2, false, NT, false, true>(g, a, blocks, smem, st);
    return masked ? blk_launch<1, true, NT, false, true>(g, a, blocks, smem, st) : blk_launch<1, false, NT, false, true>(g, a, blocks, smem, st);
  }
  if (two && pair) return masked ? blk_launch<2, true, NT, true, false>(g, a, blocks, smem, st) : blk_launch<2, false, NT, true, false>(g, a, blocks, smem, st);
  if (two) return masked ? blk_launch<2, true, NT, false, false>(g, a, blocks, smem, st) : blk_launch<2, false, NT, false, false>(g, a, blocks, smem, st);
  return masked ? blk_launch<1, true, NT, false, false>(g, a, blocks, smem, st) : blk_launch<1, false, NT, false, false>(g, a, blocks, smem, st);
}

// PAIR mode pays when most neighbours of the cell-key order share their 2-bit cells. Measured on the D = 10
// benchmark grid (benchmarks/pair_threshold.py, 4^10 cells): break-even at 4 points per cell, 1.09x at 8, 1.13x at 16,
// 1.18x at 95 (1e8 points). ASG_BLOCK_PAIR=0 disables it, =<n> sets the minimum number of points per cell.
static double pair_min_per_cell() {  // read per call: the tests switch it between evaluations
  const char *e = getenv("ASG_BLOCK_PAIR");
  return e ? atof(e) : 5.0;
}
bool block_pair_mode(const DeviceGrid &g, int64_t M, bool sorted) {
  if (!sorted || !g.blk_pair_ok || g.blk_wide || g.D > 15 || pair_min_per_cell() <= 0.0) return false;
  if (M <= (int64_t)blk_sm_count() * blk_threads() || blk_smem_bytes(g, 2, blk_threads()) > 227 * 1024) return false;
  return (double)M >= pair_min_per_cell() * std::ldexp(1.0, 2 * g.D);
}
// (Measured and rejected: groups of FOUR neighbours per thread at 256 threads x 255 registers -- 24 % fewer warp
// instructions per point, but 8 instead of 16 warps per SM: 433 ms instead of 407 ms per 1e8 points. The kernel template
// still takes any R >= 2 in PAIR mode.)

cudaError_t launch_eval_block(const DeviceGrid &g, const EvalArgs &a_in, cudaStream_t st, int64_t *launches,
                              uint32_t *left_list, uint32_t *left_count) {
  EvalArgs a = a_in;
  if (a.M <= 0) return cudaSuccess;
  const int sms = blk_sm_count();
  const int nt = blk_threads();
  // two points per thread as soon as one-point CTAs would need a second wave (and the shared memory fits)
  bool two = a.M > (int64_t)sms * nt && blk_smem_bytes(g, 2, nt) <= 227 * 1024;
  if (const char *e = getenv("ASG_BLOCK_R")) two = (e[0] == '2') && blk_smem_bytes(g, 2, nt) <= 227 * 1024;  // experiment knob
  const bool pair = two && left_list && left_count && block_pair_mode(g, a.M, a.perm != nullptr);
  const int R = two ? 2 : 1;
  const size_t smem = blk_smem_bytes(g, R, nt);
  const int64_t per_block = (int64_t)nt * R;
  const int64_t need = (a.M + per_block - 1) / per_block;
  const int blocks = (int)(need < sms ? need : sms);
  const bool masked = a.rem < INT_MAX / 4;
  if (pair) {
    cudaError_t e = cudaMemsetAsync(left_count, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    a.left_list = left_list;
    a.left_count = left_count;
  }
  cudaError_t e = blk_launch_nt<kBlkThreads>(g, a, two, masked, pair, blocks, smem, st);
  if (e != cudaSuccess) return e;  // the shared-memory limit of the kernel could not be raised
  if (launches) ++*launches;
  e = cudaGetLastError();
  if (e != cudaSuccess || !pair) return e;
  // follow-up: the second points of the pairs that straddle a cell boundary (about half a point per occupied cell),
  // two unpaired points per thread; their number is only known on the device, so the launch covers the worst case and
  // surplus CTAs leave at once
  EvalArgs b = a_in;
  b.perm = left_list;
  b.M_dev = left_count;
  b.M = a.M / 2 + 1;  // worst case of the list
  const int64_t need1 = (b.M + per_block - 1) / per_block;
  e = blk_launch_nt<kBlkThreads>(g, b, true, masked, false, (int)(need1 < sms ? need1 : sms), smem, st);
  if (e != cudaSuccess) return e;
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace asg
