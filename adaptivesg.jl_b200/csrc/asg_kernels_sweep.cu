// asg_kernels_sweep.cu -- SWEEP engine (sm_100a): the reference's loop nest itself,
//   vals = copy(coefs); for d: vals .*= phi.(x01_d, L[:,d], I[:,d]); sum(vals)   (src/class.jl:635-642)
// turned inside out: one thread owns one query point, the CTA streams tiles of nodes through shared
// memory, and per node the product ((c*phi_1)*phi_2)...*phi_D is formed in the reference's order
// (src/class.jl:638-641) with a warp-ballot early exit once the product is 0 for all 32 points.
// It accepts every node set phi() accepts (src/math.jl:345-357: any i for l > 2, duplicates,
// non-finite coefficients, levels up to 62) and is the engine for grids the subspace trie rejects.
//
// Per (node, dim) the packer stored (s, c) with phi = hat(x*s - c): s = 2^(l-1), c = i for l >= 2,
// s = c = 0 for l == 1. For l <= 2 the reference additionally returns 0 outside [0,1]
// (src/math.jl:349-353); that is the `low` mask tested against the point's out-of-range mask.
#include <climits>

#include "asg_internal.h"

namespace asg {

constexpr int kSweepTile = 64;      // nodes per shared-memory tile
constexpr int kSweepThreads = 256;

struct SweepTile {
  SweepDim *sd;     // [tile][D]
  double *coef;     // [tile]
  uint64_t *low;    // [tile]
  int32_t *depth;   // [tile]
};

__device__ __forceinline__ SweepTile carve(unsigned char *smem, int D) {
  SweepTile t;
  t.sd = reinterpret_cast<SweepDim *>(smem);
  t.coef = reinterpret_cast<double *>(t.sd + (size_t)kSweepTile * D);
  t.low = reinterpret_cast<uint64_t *>(t.coef + kSweepTile);
  t.depth = reinterpret_cast<int32_t *>(t.low + kSweepTile);
  return t;
}
static size_t sweep_smem(int D) { return (size_t)kSweepTile * ((size_t)D * sizeof(SweepDim) + 8 + 8 + 4); }

__device__ __forceinline__ void load_tile(const DeviceGrid &g, const SweepTile &t, int64_t n0, int cnt, int D) {
  // vectorised copy: the (s,c) pairs of `cnt` consecutive nodes are contiguous in global memory
  const double2 *src = reinterpret_cast<const double2 *>(g.sw + n0 * D);
  double2 *dst = reinterpret_cast<double2 *>(t.sd);
  for (int k = threadIdx.x; k < cnt * D; k += blockDim.x) dst[k] = __ldg(src + k);
  for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
    t.coef[k] = g.sw_coef ? __ldg(g.sw_coef + n0 + k) : 1.0;
    t.low[k] = __ldg(g.sw_low + n0 + k);
    t.depth[k] = __ldg(g.sw_depth + n0 + k);
  }
}

// DT > 0: compile-time dimension (coordinates in registers); DT == 0: run-time dimension
template <int DT>
struct PointRegs {
  double x[DT > 0 ? DT : kMaxDim];
  uint64_t out;  // bit d: coordinate d is outside [0,1] (or NaN)
};

template <int DT>
__device__ __forceinline__ void load_point_sweep(const DeviceGrid &g, const double *X, int64_t m, bool active,
                                                 int64_t sp, int64_t sd, int D, PointRegs<DT> &p) {
  p.out = 0;
#pragma unroll
  for (int d = 0; d < (DT > 0 ? DT : D); ++d) {
    double u = __longlong_as_double(0x7ff8000000000000ll);  // idle lanes carry NaN: every phi is 0
    if (active) {
      const double xd = X[m * sp + d * sd];
      u = __dadd_rn(0.0, __ddiv_rn(__dsub_rn(xd, g.lb[d]), g.width[d]));  // src/math.jl:308
    }
    p.x[d] = u;
    if (!(u >= 0.0 && u <= 1.0)) p.out |= (1ull << d);
  }
}

// product of one node at one point, reference order; early exit when all 32 lanes are at 0
template <int DT>
__device__ __forceinline__ double node_product(const SweepDim *nd, uint64_t low, double start,
                                               const PointRegs<DT> &p, int D) {
  double v = start;
  const uint64_t dead = low & p.out;
#pragma unroll
  for (int d = 0; d < (DT > 0 ? DT : D); ++d) {
    const SweepDim sc = nd[d];
    const double t = fma(p.x[d], sc.s, -sc.c);  // x*2^(l-1) - i: exact product, one rounding
    double ph = (t >= -1.0 && t <= 1.0) ? 1.0 - fabs(t) : 0.0;  // src/math.jl:333
    if ((dead >> d) & 1ull) ph = 0.0;                           // src/math.jl:349-353
    v *= ph;
    if (__all_sync(0xffffffffu, v == 0.0)) break;  // 0 * finite stays 0: nothing left to add
  }
  return v;
}

template <int DT>
__global__ void __launch_bounds__(kSweepThreads) k_eval_sweep(const DeviceGrid g, const EvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int D = DT > 0 ? DT : g.D;
  const SweepTile t = carve(smem, D);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < a.M; base += stride) {
    const int64_t m = base + threadIdx.x;
    const bool active = m < a.M;
    PointRegs<DT> p;
    load_point_sweep<DT>(g, a.X, m, active, a.sp, a.sd, D, p);
    double acc = 0.0;
    for (int64_t n0 = a.n_first; n0 < g.N; n0 += kSweepTile) {
      const int cnt = (int)((g.N - n0) < kSweepTile ? (g.N - n0) : kSweepTile);
      __syncthreads();
      load_tile(g, t, n0, cnt, D);
      __syncthreads();
      for (int k = 0; k < cnt; ++k) {
        if (t.depth[k] - 1 > a.rem) continue;  // mask of _interpolate_until_level, src/class.jl:674-679
        acc += node_product<DT>(t.sd + (size_t)k * D, t.low[k], t.coef[k], p, D);
      }
    }
    if (active) {
      if (a.accumulate) a.out[m] = a.fvals ? (a.out[m] - acc) : (a.out[m] + acc);  // on top of the packed nodes' result
      else a.out[m] = a.fvals ? (a.fvals[m] - acc) : acc;
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(kSweepThreads) k_basis_sweep(const DeviceGrid g, const BasisArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int D = DT > 0 ? DT : g.D;
  const SweepTile t = carve(smem, D);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < a.M; base += stride) {
    const int64_t m = base + threadIdx.x;
    const bool active = m < a.M;
    PointRegs<DT> p;
    load_point_sweep<DT>(g, a.X, m, active, a.sp, a.sd, D, p);
    long long cnt_out = 0;
    const long long p0 = (a.mode == 1 && active) ? a.row_ptr[m] : 0;
    for (int64_t n0 = 0; n0 < g.N; n0 += kSweepTile) {
      const int cnt = (int)((g.N - n0) < kSweepTile ? (g.N - n0) : kSweepTile);
      __syncthreads();
      load_tile(g, t, n0, cnt, D);
      __syncthreads();
      for (int k = 0; k < cnt; ++k) {
        double v = node_product<DT>(t.sd + (size_t)k * D, t.low[k], 1.0, p, D);  // starts from 1, src/class.jl:445
        const bool keep = active && v != 0.0 && !(a.dropzeros && v < a.atol);      // src/class.jl:450-453
        if (a.mode == 2) {
          if (active) a.dense[m * a.ldm + (n0 + k) * a.ldn] = keep ? v : 0.0;
        } else if (keep) {
          if (a.mode == 1) {
            a.colidx[p0 + cnt_out] = n0 + k;
            a.vals[p0 + cnt_out] = v;
          }
          ++cnt_out;
        }
      }
    }
    if (a.mode == 0 && active) a.row_nnz[m] = (unsigned long long)cnt_out;
  }
}

static int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}
static int grid_for(int64_t M) {
  const int64_t need = (M + kSweepThreads - 1) / kSweepThreads;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

#define ASG_DISPATCH_DT(D_, CALL)                                                         \
  switch (D_) {                                                                           \
    case 1: CALL(1); break;   case 2: CALL(2); break;   case 3: CALL(3); break;           \
    case 4: CALL(4); break;   case 5: CALL(5); break;   case 6: CALL(6); break;           \
    case 7: CALL(7); break;   case 8: CALL(8); break;   case 9: CALL(9); break;           \
    case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break;         \
    default: CALL(0); break;                                                              \
  }

cudaError_t launch_eval_sweep(const DeviceGrid &g, const EvalArgs &a, cudaStream_t st, int64_t *launches) {
  if (a.M <= 0) return cudaSuccess;
  const size_t sm = sweep_smem(g.D);
  const int blocks = grid_for(a.M);
  cudaError_t lim = cudaSuccess;
  // limit raised once per device and instantiation to the largest tile the instantiation can need (D > 12 share one)
#define CALL(DD)                                                                                        \
  do {                                                                                                  \
    static std::atomic<unsigned> limit_set{0};                                                          \
    lim = smem_limit_once(k_eval_sweep<DD>, (int)sweep_smem(DD ? DD : kMaxDim), limit_set);            \
    if (lim == cudaSuccess) k_eval_sweep<DD><<<blocks, kSweepThreads, sm, st>>>(g, a);                  \
  } while (0)
  ASG_DISPATCH_DT(g.D, CALL)
#undef CALL
  if (lim != cudaSuccess) return lim;
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_basis_sweep(const DeviceGrid &g, const BasisArgs &a, cudaStream_t st, int64_t *launches) {
  if (a.M <= 0) return cudaSuccess;
  const size_t sm = sweep_smem(g.D);
  const int blocks = grid_for(a.M);
  cudaError_t lim = cudaSuccess;
#define CALL(DD)                                                                                        \
  do {                                                                                                  \
    static std::atomic<unsigned> limit_set{0};                                                          \
    lim = smem_limit_once(k_basis_sweep<DD>, (int)sweep_smem(DD ? DD : kMaxDim), limit_set);           \
    if (lim == cudaSuccess) k_basis_sweep<DD><<<blocks, kSweepThreads, sm, st>>>(g, a);                 \
  } while (0)
  ASG_DISPATCH_DT(g.D, CALL)
#undef CALL
  if (lim != cudaSuccess) return lim;
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace asg
