// asg_kernels_subspace.cu -- SUBSPACE engine (sm_100a): batched evaluation of
//   G(x) = sum_n c_n prod_d phi(x01_d; l_nd, i_nd)                 (src/class.jl:629-643)
// and of the basis rows prod_d phi(...) (src/class.jl:434-454) by walking the level-vector trie
// built in asg_pack.cpp.
//
// One thread owns one query point (registers: D unit-cube coordinates + D integer cell keys); a warp
// owns 32 consecutive points, so all trie records it reads are warp-uniform broadcasts and the only
// per-lane memory traffic is one 8-byte coefficient gather per subspace. Per trie node the thread
// computes the 1-D hat of the single supporting node of that level,
//      cell = floor(x01 * 2^b(l)),  i = 2*cell + [l > 2],  phi = 1 - |x01 * 2^(l-1) - i|,
// which is the reference's phi (src/math.jl:345-357) evaluated only where it is non-zero:
// x01*2^(l-1) is exact, the subtraction is the reference's single rounding (an FMA cannot change
// it), and the l == 2 forms 1-2x / 2x-1 are bit-identical to 1-|2x-i| on [0,1].
// Points outside [0,1]^D (or NaN) evaluate to exactly 0.0 as in the reference (src/math.jl:349-353).
#include <climits>

#include "asg_internal.h"

namespace asg {

__device__ __forceinline__ uint32_t hash_slot(uint32_t pos, int lg) {
  return lg == 0 ? 0u : (uint32_t)((pos * 0x9E3779B1u) >> (32 - lg));  // == asg_pack.cpp
}

__device__ __forceinline__ TrieRec ld_rec(const TrieRec *p) {
  const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
  TrieRec r;
  r.x = v.x;
  r.y = v.y;
  return r;
}

// 1-D hat of the supporting node of level l >= 2 at unit-cube coordinate x (q = its 30-bit cell key).
__device__ __forceinline__ void hat_li(double x, uint32_t q, int l, double &ph, uint32_t &cell, int &b) {
  b = l > 3 ? l - 2 : 1;
  cell = q >> (kKeyBits - b);
  const uint32_t i = 2u * cell + (l > 2 ? 1u : 0u);
  // (double) i without the slow I2F.F64: 2^52 + i has i in its low mantissa bits
  const double di = __hiloint2double(0x43300000, (int)i) - 4503599627370496.0;
  const double s = __hiloint2double((1023 + l - 1) << 20, 0);  // 2^(l-1) = power2(l-1), src/math.jl:59-61
  const double t = fma(x, s, -di);                             // x*2^(l-1) - i, one rounding as in the reference
  ph = 1.0 - fabs(t);                                          // |t| <= 1 by construction of the cell
}

// ---- sinks: what happens at a leaf (= one subspace) --------------------------------------------
struct EvalSink {
  double acc;
  __device__ __forceinline__ void dense(const DeviceGrid &g, uint32_t slot, double P) {
    acc = fma(__ldg(g.coef + slot), P, acc);
  }
  __device__ __forceinline__ void hashed(const DeviceGrid &, const HashEnt &e, uint32_t, double P) {
    acc = fma(e.coef, P, acc);
  }
};

struct BasisSink {
  int mode;  // 0 count, 1 fill, 2 dense
  double atol;
  int dropzeros;
  long long cnt;      // entries emitted so far in this row
  long long *colidx;  // row start (mode 1)
  double *vals;
  double *row;  // dense row base (mode 2)
  int64_t ldn;
  __device__ __forceinline__ void emit(int32_t col, double v) {
    if (col < 0 || v == 0.0 || (dropzeros && v < atol)) return;  // src/class.jl:450-453
    if (mode == 1) {
      colidx[cnt] = col;
      vals[cnt] = v;
    } else if (mode == 2) {
      row[(int64_t)col * ldn] = v;
    }
    ++cnt;
  }
  __device__ __forceinline__ void dense(const DeviceGrid &g, uint32_t slot, double P) { emit(__ldg(g.orig + slot), P); }
  __device__ __forceinline__ void hashed(const DeviceGrid &g, const HashEnt &, uint32_t hslot, double P) {
    emit(__ldg(g.horig + hslot), P);
  }
};

template <class Sink>
__device__ __forceinline__ void leaf(const DeviceGrid &g, const TrieRec &r, double P, uint32_t O, Sink &sink) {
  if (!(r.y & kRecHashed)) {
    sink.dense(g, r.x + O, P);
  } else {
    const int lg = (int)((r.y >> 16) & 0xffu);
    const uint32_t mask = (1u << lg) - 1u;
    const unsigned long long key = (unsigned long long)O + 1ull;
    uint32_t s = hash_slot(O, lg);
    for (;;) {  // table is at most half full: an empty slot always ends the probe
      const uint32_t hs = r.x + s;
      const double2 raw = __ldg(reinterpret_cast<const double2 *>(g.hent + hs));
      HashEnt e;
      e.key = (unsigned long long)__double_as_longlong(raw.x);
      e.coef = raw.y;
      if (e.key == key) { sink.hashed(g, e, hs, P); break; }
      if (e.key == 0ull) break;
      s = (s + 1u) & mask;
    }
  }
}

// ---- the walk: compile-time recursion over dimensions, so x[K], q[K], P, O all live in registers -
template <int D, int K>
struct Walk {
  template <class Sink>
  static __device__ __forceinline__ void run(const DeviceGrid &g, const double (&x)[D], const uint32_t (&q)[D],
                                             double P, uint32_t O, uint32_t jb, uint32_t je, int rem, Sink &sink) {
    const TrieRec *rec = g.rec[K];
    TrieRec r = ld_rec(rec + jb);
    for (uint32_t j = jb; j < je; ++j) {
      const TrieRec rn = ld_rec(rec + j + 1);  // next sibling (or the sentinel): gives this node's child end
      const int l = (int)(r.y & 0xffu);
      if (l - 1 > rem) break;  // depth mask of _interpolate_until_level; siblings are sorted by level
      double Pn = P;
      uint32_t On = O;
      if (l > 1) {
        double ph;
        uint32_t cell;
        int b;
        hat_li(x[K], q[K], l, ph, cell, b);
        Pn = P * ph;
        On = (O << b) | cell;
      }
      if constexpr (K < D - 1) {
        Walk<D, K + 1>::run(g, x, q, Pn, On, r.x, rn.x, rem - (l - 1), sink);
      } else {
        leaf(g, r, Pn, On, sink);
      }
      r = rn;
    }
  }
};

// scale() of src/math.jl:301-309 with to = [0,1]:  0.0 + (1.0 * (x - lb)) / (ub - lb)
template <int D>
__device__ __forceinline__ bool load_point(const DeviceGrid &g, const double *X, int64_t m, int64_t sp, int64_t sd,
                                           double (&x)[D], uint32_t (&q)[D]) {
  bool inside = true;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double xd = X[m * sp + d * sd];
    const double u = __dadd_rn(0.0, __ddiv_rn(__dsub_rn(xd, g.lb[d]), g.width[d]));
    x[d] = u;
    inside = inside && (u >= 0.0) && (u <= 1.0);  // NaN fails both
    const double k = u * 1073741824.0;              // exact
    q[d] = (u >= 0.0 && u < 1.0) ? (uint32_t)k : ((1u << kKeyBits) - 1u);
  }
  return inside;
}

template <int D>
__global__ void __launch_bounds__(256) k_eval_subspace(const DeviceGrid g, const EvalArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < a.M; m += stride) {
    double x[D];
    uint32_t q[D];
    const bool inside = load_point<D>(g, a.X, m, a.sp, a.sd, x, q);
    EvalSink sink;
    sink.acc = 0.0;
    if (inside) Walk<D, 0>::run(g, x, q, 1.0, 0u, 0u, g.n_root, a.rem, sink);
    a.out[m] = a.fvals ? (a.fvals[m] - sink.acc) : sink.acc;
  }
}

template <int D>
__global__ void __launch_bounds__(256) k_basis_subspace(const DeviceGrid g, const BasisArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < a.M; m += stride) {
    double x[D];
    uint32_t q[D];
    const bool inside = load_point<D>(g, a.X, m, a.sp, a.sd, x, q);
    BasisSink sink;
    sink.mode = a.mode;
    sink.atol = a.atol;
    sink.dropzeros = a.dropzeros;
    sink.cnt = 0;
    sink.colidx = nullptr;
    sink.vals = nullptr;
    sink.row = nullptr;
    sink.ldn = a.ldn;
    if (a.mode == 1) {
      const long long p0 = a.row_ptr[m];
      sink.colidx = a.colidx + p0;
      sink.vals = a.vals + p0;
    } else if (a.mode == 2) {
      sink.row = a.dense + m * a.ldm;
    }
    if (inside) Walk<D, 0>::run(g, x, q, 1.0, 0u, 0u, g.n_root, INT_MAX / 2, sink);
    if (a.mode == 0) a.row_nnz[m] = (unsigned long long)sink.cnt;
  }
}

static int grid_for(int64_t M, int threads, int sms) {
  const int64_t need = (M + threads - 1) / threads;
  const int64_t cap = (int64_t)sms * 16;  // a few waves of resident CTAs, grid-stride beyond that
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
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
  const int blocks = grid_for(a.M, threads, sm_count());
#define CALL(DD) k_eval_subspace<DD><<<blocks, threads, 0, st>>>(g, a)
  ASG_DISPATCH_D(g.D, CALL)
#undef CALL
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_basis_subspace(const DeviceGrid &g, const BasisArgs &a, cudaStream_t st, int64_t *launches) {
  if (a.M <= 0) return cudaSuccess;
  const int threads = 256;
  const int blocks = grid_for(a.M, threads, sm_count());
#define CALL(DD) k_basis_subspace<DD><<<blocks, threads, 0, st>>>(g, a)
  ASG_DISPATCH_D(g.D, CALL)
#undef CALL
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace asg
