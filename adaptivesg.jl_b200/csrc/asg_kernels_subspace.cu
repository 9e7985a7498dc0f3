// asg_kernels_subspace.cu -- SUBSPACE engine (sm_100a): batched evaluation of
//   G(x) = sum_n c_n prod_d phi(x01_d; l_nd, i_nd)                 (src/class.jl:629-643)
// and of the basis rows prod_d phi(...) (src/class.jl:434-454) by walking the level-vector trie
// built in asg_pack.cpp.
//
// One thread owns R query points (registers: D unit-cube coordinates + D integer cell keys each); a
// warp owns 32*R points, so every trie record it reads is a warp-uniform broadcast, amortised over R
// points, and the only per-lane memory traffic is one 8-byte coefficient gather per (point, subspace).
// Per trie node the thread computes the 1-D hat of the single supporting node of that level,
//      cell = floor(x01 * 2^b(l)),  i = 2*cell + [l > 2],  phi = 1 - |x01 * 2^(l-1) - i|,
// which is the reference's phi (src/math.jl:345-357) evaluated only where it is non-zero:
// x01*2^(l-1) is exact, the subtraction is the reference's single rounding (an FMA cannot change
// it), and the l == 2 forms 1-2x / 2x-1 are bit-identical to 1-|2x-i| on [0,1].
// For the LAST dimension the hats and cells of levels 1..8 are tabulated once per point in registers,
// and a FULL prefix group (children = levels 1..lm, dense, back to back) is evaluated as
//      acc += P * sum_l coef[z + ((off(l)+cell_l) << pb) + O] * phi_l
// i.e. three instructions (IMAD.WIDE, LDG, DFMA) per (point, subspace).
// Points outside [0,1]^D (or NaN) evaluate to exactly 0.0 as in the reference (src/math.jl:349-353).
#include "asg_device.cuh"

namespace asg {

// ---- the walk: compile-time recursion over dimensions, so x[K], q[K], P, O all live in registers -
template <int D, int R, int K>
struct Walk {
  template <class Sink>
  static __device__ __forceinline__ void run(const DeviceGrid &g, const Points<D, R> &pt, const Prefix<R> &pf,
                                             uint32_t jb, uint32_t je, int rem, Sink &sink) {
    if constexpr (K == D - 1) {
      leaves_generic<D, R>(g, pt, pf, jb, je, rem, 0, sink);  // D == 1: the root's children are the leaves
    } else {
      const TrieRec *rec = g.rec[K];
      TrieRec r = ld_rec(rec + jb);
      for (uint32_t j = jb; j < je; ++j) {
        const TrieRec rn = ld_rec(rec + j + 1);  // next sibling (or the sentinel): gives this node's child end
        const int l = (int)(r.y & 0xffu);
        if (l - 1 > rem) break;  // depth mask of _interpolate_until_level; siblings are sorted by level
        Prefix<R> nx;
        if (l > 1) {
          const int b = (int)((r.y >> 24) & 0x3fu);  // packer: cell bits of the level
          const uint32_t odd = (r.y >> 30) & 1u;     // packer: l > 2
          const uint32_t twob = 1u << b;
          const double s = pow2_lm1(l);
#pragma unroll
          for (int k = 0; k < R; ++k) {
            double ph;
            uint32_t cell;
            hat_cell(pt.x[k][K], pt.q[k][K], b, odd, s, ph, cell);
            nx.P[k] = pf.P[k] * ph;
            nx.O[k] = pf.O[k] * twob + cell;  // (O << b) | cell as one IMAD
          }
        } else {
#pragma unroll
          for (int k = 0; k < R; ++k) { nx.P[k] = pf.P[k]; nx.O[k] = pf.O[k]; }
        }
        const int rem2 = rem - (l - 1);
        if constexpr (K < D - 2) {
          Walk<D, R, K + 1>::run(g, pt, nx, r.x, rn.x, rem2, sink);
        } else {  // K == D-2: this node is a prefix group, its children are leaves
          const int pb = (int)(r.w & 0xffu);
          bool fast = false;
          if constexpr (Sink::kFast) {
            if (r.w & kRecFull) {
              const int lm = min((int)((r.w >> 8) & 0xffu), rem2 + 1);
              leaves_full<D, R>(g, pt, nx, r.z, pb, lm, sink);
              fast = true;
            }
          }
          if (!fast) leaves_generic<D, R>(g, pt, nx, r.x, rn.x, rem2, pb, sink);
        }
        r = rn;
      }
    }
  }
};

// scale() of src/math.jl:301-309 with to = [0,1]:  0.0 + (1.0 * (x - lb)) / (ub - lb)
// Returns false for a point outside the domain / NaN; such a point is walked at the centre of the
// cube (valid addresses) and its result is replaced by the reference's exact 0.0 afterwards.
template <int D, int R>
__device__ __forceinline__ bool load_point(const DeviceGrid &g, const double *X, int64_t m, int64_t sp, int64_t sd,
                                           Points<D, R> &pt, int k) {
  bool inside = true;
  double u[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double xd = X[m * sp + d * sd];
    u[d] = __dadd_rn(0.0, __ddiv_rn(__dsub_rn(xd, g.lb[d]), g.width[d]));
    inside = inside && (u[d] >= 0.0) && (u[d] <= 1.0);  // NaN fails both
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double v = inside ? u[d] : 0.5;
    pt.x[k][d] = v;
    pt.q[k][d] = v < 1.0 ? (uint32_t)(v * 1073741824.0) : ((1u << kKeyBits) - 1u);  // floor(v * 2^30), exact
  }
  // last-dimension tables
  const double xl = pt.x[k][D - 1];
  const uint32_t ql = pt.q[k][D - 1];
  pt.ph[k][0] = pt.ph[k][1] = 1.0;
  pt.w[k][0] = pt.w[k][1] = 0u;
#pragma unroll
  for (int l = 2; l <= kLeafTable; ++l) {
    const int b = l > 3 ? l - 2 : 1;
    uint32_t cell;
    hat_cell(xl, ql, b, l > 2 ? 1u : 0u, pow2_lm1(l), pt.ph[k][l], cell);
    // off(l) = number of cells of levels 1..l-1 = 0,1,3,5,9,17,33,65 = 2^(l-2) + 1 for l >= 3
    const uint32_t off = l == 2 ? 1u : ((1u << (l - 2)) + 1u);
    pt.w[k][l] = off + cell;
  }
  return inside;
}

template <int D, int R>
__global__ void __launch_bounds__(256, (R == 1 ? 2 : 1)) k_eval_subspace(const DeviceGrid g, const EvalArgs a) {
  const int64_t tile = (int64_t)blockDim.x * R;
  const int64_t stride = (int64_t)gridDim.x * tile;
  for (int64_t base = (int64_t)blockIdx.x * tile; base < a.M; base += stride) {
    Points<D, R> pt;
    bool inside[R];
    int64_t m[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      m[k] = base + (int64_t)k * blockDim.x + threadIdx.x;
      const int64_t mm = m[k] < a.M ? m[k] : a.M - 1;  // tail lanes redo the last point, never store
      inside[k] = load_point<D, R>(g, a.X, mm, a.sp, a.sd, pt, k);
    }
    EvalSink<R> sink;
    Prefix<R> pf;
#pragma unroll
    for (int k = 0; k < R; ++k) { sink.acc[k] = 0.0; pf.P[k] = 1.0; pf.O[k] = 0u; }
    Walk<D, R, 0>::run(g, pt, pf, 0u, g.n_root, a.rem, sink);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (m[k] < a.M) {
        const double v = inside[k] ? sink.acc[k] : 0.0;
        a.out[m[k]] = a.fvals ? (a.fvals[m[k]] - v) : v;
      }
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256) k_basis_subspace(const DeviceGrid g, const BasisArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < a.M; m += stride) {
    Points<D, 1> pt;
    const bool inside = load_point<D, 1>(g, a.X, m, a.sp, a.sd, pt, 0);
    BasisSink<1> sink;
    sink.mode = a.mode;
    sink.atol = a.atol;
    sink.dropzeros = a.dropzeros;
    sink.cnt[0] = 0;
    sink.colidx[0] = nullptr;
    sink.vals[0] = nullptr;
    sink.row[0] = nullptr;
    sink.ldn = a.ldn;
    if (a.mode == 1) {
      const long long p0 = a.row_ptr[m];
      sink.colidx[0] = a.colidx + p0;
      sink.vals[0] = a.vals + p0;
    } else if (a.mode == 2) {
      sink.row[0] = a.dense + m * a.ldm;
    }
    if (inside) {  // outside the domain every basis value is 0: the row is empty
      Prefix<1> pf;
      pf.P[0] = 1.0;
      pf.O[0] = 0u;
      Walk<D, 1, 0>::run(g, pt, pf, 0u, g.n_root, INT_MAX / 2, sink);
    }
    if (a.mode == 0) a.row_nnz[m] = (unsigned long long)sink.cnt[0];
  }
}

// ---- dense basis rows: the scatter pass split over the subtrees of the level-vector trie --------------------------
// asg_basis_dense clears the M x N block and scatters one entry per subspace and row into it. With one thread per
// row the scatter is latency bound -- a thread visits its ~715 subspaces (BASELINE configs[4]) one dependent trie
// record after the other, 1.6 ms for 50 000 rows -- although it moves almost no data. Here a thread owns one row AND
// one node of trie level 1 (a pair of levels of dimensions 0 and 1): ~50 times more threads, each with a ~50 times
// shorter chain. What is left of the scatter (1.4 ms) is the memory system's handling of 3.6e7 scattered 8-byte writes
// into freshly cleared lines. (Measured and rejected on configs[4], 18.7 GB, against 3.95 ms for memset + this scatter:
// rows built in CSR form first and written once by a row-streaming kernel -- count + fill + writer 4.7 ms; the same
// fused into one kernel 5.3 ms; per-row strips filled by this split walk with atomic appends (0.92 ms) + a writer that
// looks every column up in a shared-memory bitmap (4.0 ms) or assembles 32 KB segments in shared memory (3.3 ms).
// cudaMemset2D itself runs at 7.4 TB/s; none of the hand-written row writers came close to it.)
template <int D>
__global__ void __launch_bounds__(256) k_basis_dense_split(const DeviceGrid g, const BasisArgs a, uint32_t n1) {
  static_assert(D >= 3, "two trie levels above the prefix groups");
  const int64_t total = a.M * (int64_t)n1;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / n1;
    const uint32_t j1 = (uint32_t)(t - m * n1);
    Points<D, 1> pt;
    if (!load_point<D, 1>(g, a.X, m, a.sp, a.sd, pt, 0)) continue;  // outside the domain the row stays zero
    // the level-0 node that owns child j1 (children ranges are consecutive; a handful of nodes)
    uint32_t j0 = 0;
    while (j0 + 1 < g.n_root && ld_rec(g.rec[0] + j0 + 1).x <= j1) ++j0;
    const TrieRec r = ld_rec(g.rec[0] + j0);
    const int l = (int)(r.y & 0xffu);
    Prefix<1> pf;
    pf.P[0] = 1.0;
    pf.O[0] = 0u;
    if (l > 1) {
      double ph;
      uint32_t cell;
      hat_cell(pt.x[0][0], pt.q[0][0], (int)((r.y >> 24) & 0x3fu), (r.y >> 30) & 1u, pow2_lm1(l), ph, cell);
      pf.P[0] = ph;
      pf.O[0] = cell;
    }
    BasisSink<1> sink;
    sink.mode = 2;
    sink.atol = a.atol;
    sink.dropzeros = a.dropzeros;
    sink.cnt[0] = 0;
    sink.colidx[0] = nullptr;
    sink.vals[0] = nullptr;
    sink.row[0] = a.dense + m * a.ldm;
    sink.ldn = a.ldn;
    Walk<D, 1, 1>::run(g, pt, pf, j1, j1 + 1, INT_MAX / 2, sink);
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

static int grid_for(int64_t M, int per_block, int sms, int waves) {
  const int64_t need = (M + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sms * waves;  // resident CTAs x a few waves, grid-stride beyond that
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

#define ASG_DISPATCH_D(D_, CALL)                                                          \
  switch (D_) {                                                                           \
    case 1: CALL(1); break;   case 2: CALL(2); break;   case 3: CALL(3); break;           \
    case 4: CALL(4); break;   case 5: CALL(5); break;   case 6: CALL(6); break;           \
    case 7: CALL(7); break;   case 8: CALL(8); break;   case 9: CALL(9); break;           \
    case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break;         \
    default: return cudaErrorInvalidValue;                                                \
  }

cudaError_t launch_eval_subspace(const DeviceGrid &g, const EvalArgs &a, cudaStream_t st, int64_t *launches) {
  if (a.M <= 0) return cudaSuccess;
  const int threads = 256;
  const int sms = sm_count();
  // two points per thread once there is enough work to fill the machine with half the threads
  const bool two = a.M >= (int64_t)sms * threads * 4;
  if (two) {
    const int blocks = grid_for(a.M, threads * 2, sms, 16);
#define CALL(DD) k_eval_subspace<DD, 2><<<blocks, threads, 0, st>>>(g, a)
    ASG_DISPATCH_D(g.D, CALL)
#undef CALL
  } else {
    const int blocks = grid_for(a.M, threads, sms, 16);
#define CALL(DD) k_eval_subspace<DD, 1><<<blocks, threads, 0, st>>>(g, a)
    ASG_DISPATCH_D(g.D, CALL)
#undef CALL
  }
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_basis_subspace(const DeviceGrid &g, const BasisArgs &a, cudaStream_t st, int64_t *launches) {
  if (a.M <= 0) return cudaSuccess;
  const int threads = 256;
  const int blocks = grid_for(a.M, threads, sm_count(), 16);
#define CALL(DD) k_basis_subspace<DD><<<blocks, threads, 0, st>>>(g, a)
  ASG_DISPATCH_D(g.D, CALL)
#undef CALL
  if (launches) ++*launches;
  return cudaGetLastError();
}

// dense scatter (a.mode == 2, output already cleared), split over the nodes of trie level 1 where the trie has them
cudaError_t launch_basis_dense_split(const DeviceGrid &g, const BasisArgs &a, uint32_t n1, cudaStream_t st, int64_t *launches) {
  if (a.M <= 0) return cudaSuccess;
  if (g.D < 3 || n1 == 0) return launch_basis_subspace(g, a, st, launches);
  const int64_t need = (a.M * (int64_t)n1 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 64;
  const int blocks = (int)(need < cap ? need : cap);
  switch (g.D) {
#define CASE(DD) case DD: k_basis_dense_split<DD><<<blocks, 256, 0, st>>>(g, a, n1); break;
    CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
#undef CASE
    default: return cudaErrorInvalidValue;
  }
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace asg
