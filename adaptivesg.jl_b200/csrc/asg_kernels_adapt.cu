// asg_kernels_adapt.cu -- adapt! candidate generation on the device (sm_100a):
// the children of all nodes of one depth (src/train.jl:337-398: for j in 1:D, for parent, left then right;
// child: src/math.jl:907-959; isvalid: src/math.jl:418-448), with the duplicates removed BEFORE anything is
// evaluated. The reference evaluates f and G at every child and lets a Dict drop the repeats afterwards
// (src/train.jl:365-371, 412-413); a child of depth d+1 is reached from up to D parents, so removing the
// repeats first divides the calls of the user's f and the surplus batch by up to D with the same node set.
#include "asg_internal.h"

namespace asg {

// candidate k = (j * n_p + p) * 2 + side: row p of the parents with dimension j replaced by its left / right child
__global__ void k_children(int D, int64_t n_p, const long long *PL, const long long *PI, long long *CL, long long *CI,
                           unsigned char *valid) {
  const int64_t n_cand = n_p * D * 2;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_cand; k += (int64_t)gridDim.x * blockDim.x) {
    const int side = (int)(k & 1);
    const int64_t jp = k >> 1;
    const int j = (int)(jp / n_p);
    const int64_t p = jp - (int64_t)j * n_p;
    bool ok = true;
    long long dsum = 0;
    for (int d = 0; d < D; ++d) {
      long long l = PL[p * D + d], i = PI[p * D + d];
      if (d == j) {  // _child_left / _child_right, src/math.jl:907-959
        if (l > 2) i = side ? 2 * i + 1 : 2 * i - 1;
        else if (l == 1) i = side ? 2 : 0;
        else i = side ? (i == 0 ? 1 : -1) : (i == 0 ? -1 : 3);  // (2,0): only a right child; (2,2): only a left child
        l = l + 1;
      }
      CL[k * D + d] = l;
      CI[k * D + d] = i;
      // isvalid, src/math.jl:418-448: l = 1 -> i = 1; l = 2 -> i in {0, 2}; l > 2 -> i odd, >= 1
      ok = ok && ((l == 1 && i == 1) || (l == 2 && (i == 0 || i == 2)) || (l > 2 && i >= 1 && (i & 1)));
      dsum += l;
    }
    valid[k] = (ok && dsum - D + 1 >= 1) ? 1 : 0;
  }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v) {
  h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  h *= 0xFF51AFD7ED558CCDull;
  return h ^ (h >> 32);
}

__device__ __forceinline__ bool same_row(int D, const long long *CL, const long long *CI, int64_t a, int64_t b) {
  for (int d = 0; d < D; ++d)
    if (CL[a * D + d] != CL[b * D + d] || CI[a * D + d] != CI[b * D + d]) return false;
  return true;
}

// open-addressing set of candidate numbers keyed by the full (levels, indices) row; of equal rows the smallest
// candidate number stays (= first generated, the reference's `!haskey` rule in its visiting order)
__global__ void k_children_dedupe(int D, int64_t n_cand, const long long *CL, const long long *CI,
                                  const unsigned char *valid, int *table, unsigned int mask) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_cand; k += (int64_t)gridDim.x * blockDim.x) {
    if (!valid[k]) continue;
    unsigned long long h = 0x243F6A8885A308D3ull;
    for (int d = 0; d < D; ++d) h = mix64(mix64(h, (unsigned long long)CL[k * D + d]), (unsigned long long)CI[k * D + d]);
    unsigned int s = (unsigned int)h & mask;
    for (;;) {
      int cur = table[s];
      if (cur < 0) {
        cur = atomicCAS(&table[s], -1, (int)k);
        if (cur < 0) break;  // inserted
      }
      if (same_row(D, CL, CI, (int64_t)cur, k)) {
        atomicMin(&table[s], (int)k);
        break;
      }
      s = (s + 1u) & mask;
    }
  }
}

__global__ void k_children_flags(const int *table, int64_t size, unsigned int *flags) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < size; s += (int64_t)gridDim.x * blockDim.x)
    if (table[s] >= 0) flags[table[s]] = 1u;
}

// exclusive scan of 0/1 flags in three small steps (the candidate lists are at most a few million long)
constexpr int kScanBlock = 1024;
__global__ void k_scan_block_sums(const unsigned int *flags, int64_t n, unsigned int *sums) {
  __shared__ unsigned int sh[32];
  const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  unsigned int v = i < n ? flags[i] : 0u;
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = sh[threadIdx.x];
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) sums[blockIdx.x] = v;
  }
}
__global__ void k_scan_offsets(unsigned int *sums, int nblocks, unsigned int *total) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned int run = 0;
    for (int b = 0; b < nblocks; ++b) { const unsigned int v = sums[b]; sums[b] = run; run += v; }
    *total = run;
  }
}
// unique candidates, in candidate order, to the output rows
__global__ void k_children_compact(int D, int64_t n, const unsigned int *flags, const unsigned int *offsets,
                                   const long long *CL, const long long *CI, long long *OL, long long *OI) {
  __shared__ unsigned int sh[kScanBlock];
  const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const unsigned int f = i < n ? flags[i] : 0u;
  sh[threadIdx.x] = f;
  __syncthreads();
  for (int o = 1; o < kScanBlock; o <<= 1) {  // inclusive Hillis-Steele scan of the block
    const unsigned int v = threadIdx.x >= (unsigned)o ? sh[threadIdx.x - o] : 0u;
    __syncthreads();
    sh[threadIdx.x] += v;
    __syncthreads();
  }
  if (f) {
    const int64_t dst = (int64_t)offsets[blockIdx.x] + sh[threadIdx.x] - 1;
    for (int d = 0; d < D; ++d) {
      OL[dst * D + d] = CL[i * D + d];
      OI[dst * D + d] = CI[i * D + d];
    }
  }
}

static int adapt_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

// All buffers are device memory sized by the caller: CL, CI [n_cand][D]; valid [n_cand]; table [size] (power of
// two >= 2 * n_cand); flags [n_cand]; sums [ceil(n_cand / 1024) + 1]; OL, OI [n_cand][D] (upper bound).
cudaError_t launch_children(int D, int64_t n_p, const long long *PL, const long long *PI, long long *CL, long long *CI,
                            unsigned char *valid, int *table, int64_t table_size, unsigned int *flags,
                            unsigned int *sums, unsigned int *total, long long *OL, long long *OI, cudaStream_t st,
                            int64_t *launches) {
  const int64_t n_cand = n_p * D * 2;
  if (n_cand <= 0) return cudaSuccess;
  cudaError_t e;
  if ((e = cudaMemsetAsync(table, 0xff, (size_t)table_size * sizeof(int), st)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(flags, 0, (size_t)n_cand * sizeof(unsigned int), st)) != cudaSuccess) return e;
  k_children<<<adapt_blocks(n_cand), 256, 0, st>>>(D, n_p, PL, PI, CL, CI, valid);
  k_children_dedupe<<<adapt_blocks(n_cand), 256, 0, st>>>(D, n_cand, CL, CI, valid, table, (unsigned int)(table_size - 1));
  k_children_flags<<<adapt_blocks(table_size), 256, 0, st>>>(table, table_size, flags);
  const int nblocks = (int)((n_cand + kScanBlock - 1) / kScanBlock);
  k_scan_block_sums<<<nblocks, kScanBlock, 0, st>>>(flags, n_cand, sums);
  k_scan_offsets<<<1, 32, 0, st>>>(sums, nblocks, total);
  k_children_compact<<<nblocks, kScanBlock, 0, st>>>(D, n_cand, flags, sums, CL, CI, OL, OI);
  if (launches) *launches += 6;
  return cudaGetLastError();
}

}  // namespace asg
