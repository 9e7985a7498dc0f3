// asg_sort.cu -- optional pre-pass of the BLOCK walk for large batches: order the query points by a
// bit-interleaved cell key so that the 32 lanes of a warp fall into the same coarse cell of [0,1]^D.
// Lanes of one cell read the SAME coefficient in every subspace whose levels the key resolves, so the
// shared-memory gathers of the walk turn from 32 scattered 8-byte reads (3-6 bank-conflict wavefronts)
// into broadcasts. The points are not moved: the walk reads point perm[j] and writes out[perm[j]].
// The sort itself is the CUDA toolkit's cub::DeviceRadixSort (plumbing, 1-2 % of the walk's time).
#include <cub/device/device_radix_sort.cuh>

#include "asg_internal.h"

namespace asg {

// key = bit planes of the unit-cube coordinates, most significant plane first: [bit 1 of every dimension]
// [bit 2 of every dimension] ...  (`planes` planes of min(D, 32/planes...) bits, 32 bits in total at most).
__global__ void k_point_keys(const DeviceGrid g, const double *X, int64_t M, int64_t sp, int64_t sd, int planes,
                             uint32_t *keys, uint32_t *idx) {
  const int D = g.D;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    uint32_t key = 0u;
    bool inside = true;
    for (int d = 0; d < D; ++d) {
      const double u = __dadd_rn(0.0, __ddiv_rn(__dsub_rn(X[m * sp + d * sd], g.lb[d]), g.width[d]));  // scale()
      inside = inside && (u >= 0.0) && (u <= 1.0);
      const uint32_t q = u < 1.0 ? (uint32_t)(u * 256.0) : 255u;  // top 8 bits of the coordinate
      if (d < 32)
        for (int p = 0; p < planes; ++p) {
          const int pos = (planes - 1 - p) * D + (D - 1 - d);  // plane p occupies bits [(planes-1-p)*D, +D)
          if (pos < 32) key |= ((q >> (7 - p)) & 1u) << pos;
        }
    }
    keys[m] = inside ? key : 0xffffffffu;  // out-of-domain / NaN points: all at the end
    idx[m] = (uint32_t)m;
  }
}

size_t sort_points_temp_bytes(int64_t M) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)M, 0, 32);
  return bytes;
}

cudaError_t sort_points(const DeviceGrid &g, const double *X, int64_t M, int64_t sp, int64_t sd, uint32_t *keys,
                        uint32_t *keys_alt, uint32_t *idx, uint32_t *perm_out, void *temp, size_t temp_bytes,
                        cudaStream_t st, int64_t *launches) {
  if (M <= 0) return cudaSuccess;
  int planes = 32 / (g.D < 1 ? 1 : g.D);
  if (planes < 1) planes = 1;
  if (planes > 8) planes = 8;
  int64_t blocks = (M + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_point_keys<<<(int)blocks, 256, 0, st>>>(g, X, M, sp, sd, planes, keys, idx);
  if (launches) ++*launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  int end_bit = planes * g.D;
  if (end_bit > 32) end_bit = 32;
  // out-of-domain keys are 0xffffffff: sort on all 32 bits when any plane is cut so that they stay last
  e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const uint32_t *)keys, keys_alt, (const uint32_t *)idx, perm_out,
                                      (int)M, 0, 32, st);
  // (the histogram / onesweep passes of the library sort are not counted as launches of this library's kernels)
  (void)end_bit;
  return e;
}

}  // namespace asg
