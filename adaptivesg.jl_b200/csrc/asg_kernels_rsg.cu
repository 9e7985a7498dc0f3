// asg_kernels_rsg.cu -- REGULAR-GRID fast path of the subspace engine (sm_100a).
//
// When the node set is exactly a regular sparse grid -- what rsg!/train! build (src/train.jl:40-78):
// all level vectors with l_d <= ml_d and sum(l_d - 1) <= budget, every subspace complete -- the
// level-vector trie is a function of (dimension, remaining budget) alone. The walk then needs no
// trie records: child ranges are loop bounds and coefficient offsets are prefix sums of the tiny
// table S[k][rem] = number of coefficient slots below dimension k with `rem` budget left (passed
// as a kernel parameter, i.e. read through the constant bank on the uniform datapath).
// The coefficient array itself is the one asg_pack.cpp builds (subspaces in lexicographic order,
// position = (cell_last << pb) | O), so the generic trie kernels and this one are interchangeable.
//
// One thread owns R points. Dimensions 0..D-3 are walked with run-time loops and the on-the-fly hat;
// the LAST TWO dimensions are fully unrolled over their levels against per-point register tables
// (hat value + cell of every level <= 8), so that the inner work per (point, subspace) is
//        IMAD.WIDE (address), LDG (coefficient gather), DFMA.
// Arithmetic per 1-D hat is the reference's phi on its support (see asg_kernels_subspace.cu).
#include <climits>
#include <cstdlib>

#include "asg_internal.h"

namespace asg {

__device__ __forceinline__ double r_u32_to_f64(uint32_t i) {
  return __hiloint2double(0x43300000, (int)i) - 4503599627370496.0;  // exact for i < 2^32
}
__device__ __forceinline__ double r_pow2_lm1(int l) { return __hiloint2double((1023 + l - 1) << 20, 0); }

// hat + cell of level l >= 2 (b = cell bits, odd = [l > 2], s = 2^(l-1))
__device__ __forceinline__ void r_hat_cell(double x, uint32_t q, int b, uint32_t odd, double s, double &ph,
                                           uint32_t &cell) {
  cell = q >> (kKeyBits - b);
  const uint32_t i = (cell << 1) + odd;
  const double t = fma(x, s, -r_u32_to_f64(i));  // x*2^(l-1) - i: exact product, the reference's one rounding
  ph = 1.0 - fabs(t);
}

template <int D, int R>
struct RsgPoints {
  static constexpr int DW = D > 2 ? D - 2 : 1;  // dimensions walked with run-time loops
  double x[R][DW];
  uint32_t q[R][DW];
  // register tables of the last two dimensions, levels 2..kLeafTable (level 1: phi = 1, cell = 0)
  double phA[R][kLeafTable + 1];    // dimension D-2
  uint32_t ceA[R][kLeafTable + 1];  //   cell
  double phB[R][kLeafTable + 1];    // dimension D-1
  uint32_t wB[R][kLeafTable + 1];   //   off(l) + cell, off(l) = 0,1,3,5,9,... cells of the lower levels
};

template <int R>
struct RsgAcc {
  double acc[R];
};

// levels of dimension D-2 (unrolled) x levels of dimension D-1 (unrolled)
template <int D, int R>
__device__ __forceinline__ void rsg_group(const DeviceGrid &g, const RsgParams &rp, const RsgPoints<D, R> &pt,
                                          const double (&P)[R], const uint32_t (&O)[R], uint32_t B, int pb, int rem,
                                          RsgAcc<R> &out) {
  const int lmaxA = min(rp.ml[D - 2], rem - rp.cut + 1);
  double inner2[R];
#pragma unroll
  for (int k = 0; k < R; ++k) inner2[k] = 0.0;
  uint32_t Bc = B;
#pragma unroll
  for (int lA = 1; lA <= kLeafTable; ++lA) {
    if (lA > lmaxA) break;
    const int bA = lA == 1 ? 0 : (lA > 3 ? lA - 2 : 1);
    const int pbA = pb + bA;
    const int remA = rem - (lA - 1);
    const int lmaxB = min(rp.ml[D - 1], remA - rp.cut + 1);
    const uint32_t npre8 = 8u << pbA;  // byte stride between consecutive cells of the last dimension
    const double *A0[R];
    double inner1[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const uint32_t OA = lA == 1 ? O[k] : ((O[k] << bA) | pt.ceA[k][lA]);
      A0[k] = g.coef + (Bc + OA);
      inner1[k] = __ldg(A0[k]);  // level 1 of the last dimension: phi = 1, cell = 0
    }
#pragma unroll
    for (int lB = 2; lB <= kLeafTable; ++lB) {
      if (lB > lmaxB) break;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const double *p = reinterpret_cast<const double *>(reinterpret_cast<const char *>(A0[k]) +
                                                           (uint64_t)pt.wB[k][lB] * npre8);
        inner1[k] = fma(__ldg(p), pt.phB[k][lB], inner1[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) inner2[k] = lA == 1 ? inner1[k] : fma(pt.phA[k][lA], inner1[k], inner2[k]);
    Bc += rp.S[D - 1][remA] << pbA;
  }
#pragma unroll
  for (int k = 0; k < R; ++k) out.acc[k] = fma(P[k], inner2[k], out.acc[k]);
}

template <int D, int R, int K>
struct RsgWalk {
  static __device__ __forceinline__ void run(const DeviceGrid &g, const RsgParams &rp, const RsgPoints<D, R> &pt,
                                             const double (&P)[R], const uint32_t (&O)[R], uint32_t B, int pb, int rem,
                                             RsgAcc<R> &out) {
    if constexpr (K >= D - 2) {
      rsg_group<D, R>(g, rp, pt, P, O, B, pb, rem, out);
    } else {
      const int lmax = min(rp.ml[K], rem - rp.cut + 1);
      double nP[R];
      uint32_t nO[R];
#pragma unroll
      for (int k = 0; k < R; ++k) { nP[k] = P[k]; nO[k] = O[k]; }
      uint32_t Bc = B;
      for (int l = 1; l <= lmax; ++l) {
        int b = 0;
        if (l > 1) {
          b = l > 3 ? l - 2 : 1;
          const uint32_t odd = l > 2 ? 1u : 0u;
          const uint32_t twob = 1u << b;
          const double s = r_pow2_lm1(l);
#pragma unroll
          for (int k = 0; k < R; ++k) {
            double ph;
            uint32_t cell;
            r_hat_cell(pt.x[k][K], pt.q[k][K], b, odd, s, ph, cell);
            nP[k] = P[k] * ph;
            nO[k] = O[k] * twob + cell;
          }
        }
        const int rem2 = rem - (l - 1);
        RsgWalk<D, R, K + 1>::run(g, rp, pt, nP, nO, Bc, pb + b, rem2, out);
        Bc += rp.S[K + 1][rem2] << (pb + b);
      }
    }
  }
};

template <int D, int R>
__device__ __forceinline__ bool rsg_load_point(const DeviceGrid &g, const double *X, int64_t m, int64_t sp, int64_t sd,
                                               RsgPoints<D, R> &pt, int k) {
  bool inside = true;
  double u[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double xd = X[m * sp + d * sd];
    u[d] = __dadd_rn(0.0, __ddiv_rn(__dsub_rn(xd, g.lb[d]), g.width[d]));  // scale(), src/math.jl:308
    inside = inside && (u[d] >= 0.0) && (u[d] <= 1.0);                      // NaN fails both
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double v = inside ? u[d] : 0.5;  // walked at the centre, result replaced by the exact 0.0
    const uint32_t q = v < 1.0 ? (uint32_t)(v * 1073741824.0) : ((1u << kKeyBits) - 1u);
    if (d < D - 2) {
      pt.x[k][d] = v;
      pt.q[k][d] = q;
    } else {
      const bool last = (d == D - 1);
#pragma unroll
      for (int l = 2; l <= kLeafTable; ++l) {
        const int b = l > 3 ? l - 2 : 1;
        double ph;
        uint32_t cell;
        r_hat_cell(v, q, b, l > 2 ? 1u : 0u, r_pow2_lm1(l), ph, cell);
        if (last) {
          pt.phB[k][l] = ph;
          pt.wB[k][l] = (l == 2 ? 1u : ((1u << (l - 2)) + 1u)) + cell;
        } else {
          pt.phA[k][l] = ph;
          pt.ceA[k][l] = cell;
        }
      }
    }
  }
  return inside;
}

template <int D, int R>
__global__ void __launch_bounds__(128, (R == 1 ? 4 : 2)) k_eval_rsg(const DeviceGrid g, const RsgParams rp,
                                                                    const EvalArgs a) {
  const int64_t tile = (int64_t)blockDim.x * R;
  const int64_t stride = (int64_t)gridDim.x * tile;
  for (int64_t base = (int64_t)blockIdx.x * tile; base < a.M; base += stride) {
    RsgPoints<D, R> pt;
    bool inside[R];
    int64_t m[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      m[k] = base + (int64_t)k * blockDim.x + threadIdx.x;
      const int64_t mm = m[k] < a.M ? m[k] : a.M - 1;  // tail lanes redo the last point, never store
      inside[k] = rsg_load_point<D, R>(g, a.X, mm, a.sp, a.sd, pt, k);
    }
    RsgAcc<R> out;
    double P[R];
    uint32_t O[R];
#pragma unroll
    for (int k = 0; k < R; ++k) { out.acc[k] = 0.0; P[k] = 1.0; O[k] = 0u; }
    RsgWalk<D, R, 0>::run(g, rp, pt, P, O, 0u, 0, rp.budget, out);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (m[k] < a.M) {
        const double v = inside[k] ? out.acc[k] : 0.0;
        a.out[m[k]] = a.fvals ? (a.fvals[m[k]] - v) : v;
      }
    }
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

#define ASG_DISPATCH_D2(D_, CALL)                                                         \
  switch (D_) {                                                                           \
    case 2: CALL(2); break;   case 3: CALL(3); break;                                     \
    case 4: CALL(4); break;   case 5: CALL(5); break;   case 6: CALL(6); break;           \
    case 7: CALL(7); break;   case 8: CALL(8); break;   case 9: CALL(9); break;           \
    case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break;         \
    default: return cudaErrorInvalidValue;                                                \
  }

cudaError_t launch_eval_rsg(const DeviceGrid &g, const RsgParams &rp_in, const EvalArgs &a, cudaStream_t st,
                            int64_t *launches) {
  if (a.M <= 0) return cudaSuccess;
  RsgParams rp = rp_in;
  // depth mask of _interpolate_until_level: only subspaces with sum(l-1) <= a.rem are visited
  rp.cut = a.rem >= rp.budget ? 0 : rp.budget - (a.rem < 0 ? 0 : a.rem);
  if (a.rem < 0) rp.cut = rp.budget + 1;  // nothing qualifies: every loop bound becomes 0
  const int threads = 128;
  const int sms = sm_count();
  bool two = a.M >= (int64_t)sms * threads * 8;
  if (const char *e = getenv("ASG_RSG_R")) two = (e[0] == '2');  // experiment knob
  const int per_block = threads * (two ? 2 : 1);
  int64_t need = (a.M + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sms * 32;
  const int blocks = (int)(need < 1 ? 1 : (need < cap ? need : cap));
  if (two) {
#define CALL(DD) k_eval_rsg<DD, 2><<<blocks, threads, 0, st>>>(g, rp, a)
    ASG_DISPATCH_D2(g.D, CALL)
#undef CALL
  } else {
#define CALL(DD) k_eval_rsg<DD, 1><<<blocks, threads, 0, st>>>(g, rp, a)
    ASG_DISPATCH_D2(g.D, CALL)
#undef CALL
  }
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace asg
