// asg_sort.cu -- optional pre-pass of the BLOCK walk for large batches: order the query points by a
// bit-interleaved cell key so that the 32 lanes of a warp fall into the same coarse cell of [0,1]^D.
// Lanes of one cell read the SAME coefficient in every subspace whose levels the key resolves, so the
// shared-memory gathers of the walk turn from 32 scattered 8-byte reads (3-6 bank-conflict wavefronts)
// into broadcasts. The points are not moved: the walk reads point perm[j] and writes out[perm[j]].
// The sort itself is the CUDA toolkit's cub::DeviceRadixSort (plumbing, 1-2 % of the walk's time).
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

#include "asg_internal.h"

namespace asg {

// key = bit planes of the unit-cube coordinates, most significant plane first: [bit 1 of every dimension]
// [bit 2 of every dimension] ...  (`planes` planes of min(D, 32/planes...) bits, 32 bits in total at most).
__global__ void k_point_keys(const DeviceGrid g, const double *X, int64_t M, int64_t sp, int64_t sd, int planes,
                             uint32_t *keys, uint32_t *idx, int suffix_first) {
  const int D = g.D;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    uint32_t key = 0u;
    bool inside = true;
    for (int d = 0; d < D; ++d) {
      const double u = __dadd_rn(0.0, __ddiv_rn(__dsub_rn(X[m * sp + d * sd], g.lb[d]), g.width[d]));  // scale()
      inside = inside && (u >= 0.0) && (u <= 1.0);
      const uint32_t q = u < 1.0 ? (uint32_t)(u * 256.0) : 255u;  // top 8 bits of the coordinate
      if (suffix_first && D >= 4) {
        // experiment (ASG_SORT_KEY=suffix): the three compiled dimensions first, `planes` bits each, then the bit
        // planes of the prefix dimensions
        const int K = D - 3;
        if (d >= K) {
          const int pos = 32 - planes * (D - d);  // C on top, then B, then A
          if (pos >= 0) key |= (q >> (8 - planes)) << pos;
        } else {
          for (int p = 0; p < planes; ++p) {
            const int pos = 32 - 3 * planes - (p + 1) * K + (K - 1 - d);
            if (pos >= 0 && pos < 32) key |= ((q >> (7 - p)) & 1u) << pos;
          }
        }
      } else if (d < 32)
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
  const char *kenv = getenv("ASG_SORT_KEY");
  k_point_keys<<<(int)blocks, 256, 0, st>>>(g, X, M, sp, sd, planes, keys, idx, (kenv && kenv[0] == 's') ? 1 : 0);
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

// ---- CSR -> CSC on the device (basis matrices: SparseMatrixCSC is what the reference returns, src/class.jl:483-503) ----
namespace asg {

__global__ void k_csr_expand(const long long *row_ptr, int64_t M, const long long *colidx, uint32_t *keys, uint32_t *ent,
                             uint32_t *rows) {
  // one warp per row: entry e of row m gets key = its column, payload = e, and remembers m
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp; m < M; m += nwarps)
    for (long long e = row_ptr[m] + lane; e < row_ptr[m + 1]; e += 32) {
      keys[e] = (uint32_t)colidx[e];
      ent[e] = (uint32_t)e;
      rows[e] = (uint32_t)m;
    }
}
__global__ void k_csc_gather(int64_t nnz, const uint32_t *perm, const uint32_t *rows, const double *vals, long long *rowidx,
                             double *vout) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t e = perm[k];
    rowidx[k] = rows[e];
    vout[k] = vals[e];
  }
}
__global__ void k_csc_colptr(int64_t nnz, int64_t N, const uint32_t *sorted_cols, long long *colptr) {
  // colptr[c] = first position whose column is >= c (binary search in the sorted column keys)
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= N; c += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)sorted_cols[mid] < c) lo = mid + 1; else hi = mid;
    }
    colptr[c] = lo;
  }
}

size_t csc_temp_bytes(int64_t nnz) { return sort_points_temp_bytes(nnz); }

// CSR (row_ptr, colidx, vals; rows in ascending order) -> CSC (colptr, rowidx, vout): a STABLE sort of the entries by
// column keeps the rows ascending inside every column. keys / keys_alt / ent / perm / rows: nnz uint32 each.
cudaError_t csr_to_csc(const long long *row_ptr, const long long *colidx, const double *vals, int64_t M, int64_t N,
                       int64_t nnz, uint32_t *keys, uint32_t *keys_alt, uint32_t *ent, uint32_t *perm, uint32_t *rows,
                       void *temp, size_t temp_bytes, long long *colptr, long long *rowidx, double *vout, cudaStream_t st,
                       int64_t *launches) {
  if (nnz <= 0) {
    k_csc_colptr<<<(int)((N + 256) / 256), 256, 0, st>>>(0, N, keys, colptr);
    if (launches) ++*launches;
    return cudaGetLastError();
  }
  int64_t blocks = (M * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_csr_expand<<<(int)blocks, 256, 0, st>>>(row_ptr, M, colidx, keys, ent, rows);
  int bits = 1;
  while (((int64_t)1 << bits) < N) ++bits;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const uint32_t *)keys, keys_alt, (const uint32_t *)ent, perm,
                                                  (int)nnz, 0, bits, st);
  if (e != cudaSuccess) return e;
  blocks = (nnz + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_csc_gather<<<(int)blocks, 256, 0, st>>>(nnz, perm, rows, vals, rowidx, vout);
  k_csc_colptr<<<(int)((N + 256) / 256), 256, 0, st>>>(nnz, N, keys_alt, colptr);
  if (launches) *launches += 3;
  return cudaGetLastError();
}

}  // namespace asg

namespace asg {
// 64-bit -> 32-bit indices (asg_basis_fill_csc32: half the bytes to bring home for callers whose sparse type holds Int32)
__global__ void k_narrow_i64(const long long *in, int *out, int64_t n) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) out[k] = (int)in[k];
}
cudaError_t launch_narrow_i64(const long long *in, int *out, int64_t n, cudaStream_t st, int64_t *launches) {
  if (n <= 0) return cudaSuccess;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_narrow_i64<<<(int)blocks, 256, 0, st>>>(in, out, n);
  if (launches) ++*launches;
  return cudaGetLastError();
}
}  // namespace asg
