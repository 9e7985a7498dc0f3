// asg_kernels_stream.cu -- STREAM walk of the subspace engine (sm_100a): the main evaluation kernel.
//
//   G(x) = sum_n c_n prod_d phi(x01_d; l_nd, i_nd)                       (src/class.jl:629-643)
//
// The level-vector trie of asg_pack.cpp is flattened on the host into one 32-byte record per prefix
// group (all subspaces that share the levels of dimensions 0..D-2), in depth-first order, so the
// device runs ONE flat loop with no tree bookkeeping:
//   * the CTA streams tiles of {records, the coefficient slices they reference} from L2 into shared
//     memory with 1-D TMA bulk copies (cp.async.bulk) completed on mbarriers, double buffered;
//   * a thread owns R query points; a record is a warp-uniform broadcast read from shared memory;
//   * per record the prefix products P[k] / prefix positions O[k] are re-derived only from the first
//     dimension k0 whose level changed (a fall-through switch, so P[k], O[k] stay in registers);
//   * the leaves of the group are the levels 1..lm of the last dimension, evaluated against
//     per-point register tables (hat value, cell) as  IMAD (address), LDS.64 (gather), DFMA.
// The 1-D hat is the reference's phi (src/math.jl:345-357) evaluated on its support only, with the
// reference's single rounding (see asg_device.cuh). Out-of-domain / NaN points give exactly 0.0.
// Groups that are not FULL (adapt! output: missing levels, hashed subspaces) or whose slice exceeds
// a tile take the same loop but read leaf records / coefficients from global memory.
#include <cstdlib>

#include "asg_device.cuh"

namespace asg {

// threads per CTA: one point per thread -> 2 CTAs of 256 per SM; two points per thread -> 1 CTA of 384
// Two configurations. V = 0 (one point per thread, small batches): saved prefix products in registers,
// cell keys in shared memory, 2 CTAs of 256 threads per SM. V = 1 (two points per thread, large batches):
// the saved prefix products live in shared memory too and cells are re-derived from the coordinates,
// which keeps the kernel at 128 registers so that one CTA of 512 threads (16 warps, 1024 points) owns an SM
// and the warp-uniform bookkeeping of a record is shared by 64 points.
template <int R, int V = 0>
struct StreamCfg {
  static constexpr int kThreads = V == 1 ? 512 : (R == 1 ? 256 : 384);
  static constexpr int kCtasPerSm = V == 1 ? 1 : (R == 1 ? 2 : 1);
};

struct __align__(16) StreamSmem {
  double coef[2][kTileCoefs];
  GroupRec rec[2][kTileRecs];
  unsigned long long bar[2];
  // followed by the unit-cube coordinates / cell keys of the prefix dimensions of every point of the
  // CTA: double xs[D][R][threads]; uint32_t qs[D][R][threads]  (read only at non-trivial stages)
};

// producer side (one thread): start the copies of tile t into buffer `buf`
__device__ __forceinline__ void issue_tile(const DeviceGrid &g, StreamSmem &sm, uint32_t t, int buf) {
  const StreamTile tl = g.tiles[t];
  const uint32_t rbytes = (tl.r1 - tl.r0) * (uint32_t)sizeof(GroupRec);
  const uint32_t cbytes = (tl.c1 - tl.c0) * 8u;
  // the buffer was last read through the generic proxy: order those reads before the async-proxy writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(&sm.bar[buf], rbytes + cbytes);
  tma_load_1d(sm.rec[buf], g.grp + tl.r0, rbytes, &sm.bar[buf]);
  if (cbytes) tma_load_1d(sm.coef[buf], g.coef + tl.c0, cbytes, &sm.bar[buf]);
}

template <int D, int R, int V = 0>
struct StreamPoints {
  static constexpr int NT = StreamCfg<R, V>::kThreads;
  const double *xs;                 // coordinates of all D dimensions: xs[(K*R + k)*NT + tid]
  const uint32_t *qs;               // 30-bit cell keys, same layout (V == 0 only)
  double *Ps;                       // V == 1: saved prefix products  Ps[(K*R + k)*NT + tid], K = 0..D-4
  uint32_t *Os;                     // V == 1: saved prefix positions
  // unit-cube coordinate and cell (top b bits of the key) of point k in dimension K
  __device__ __forceinline__ double x(int K, int k) const { return xs[(K * R + k) * NT + threadIdx.x]; }
  __device__ __forceinline__ uint32_t cell(int K, int k, int b, double xv) const {
    if constexpr (V == 0) {
      return qs[(K * R + k) * NT + threadIdx.x] >> (kKeyBits - b);
    } else {  // floor(x * 2^b), clamped for x == 1 (exact: scaling by a power of two, truncation of a value >= 0)
      const uint32_t c = (uint32_t)(xv * __hiloint2double((1023 + b) << 20, 0));
      return min(c, (1u << b) - 1u);
    }
  }
  double ph[R][kLeafTable + 1];    // last dimension: hat of level l (levels 2..8)
  uint32_t w[R][kLeafTable + 1];   // last dimension: off(l) + cell_l
  double phA[R][kLeafTable + 1];   // dimension D-2: hat of level l   (SUPER records)
  uint32_t ceA[R][kLeafTable + 1]; // dimension D-2: cell of level l
};

template <int D, int R, int V>
__device__ __forceinline__ bool stream_load_point(const DeviceGrid &g, const double *X, int64_t m, int64_t sp,
                                                  int64_t sd, StreamPoints<D, R, V> &pt, int k, double *xs, uint32_t *qs) {
  constexpr int NT = StreamCfg<R, V>::kThreads;
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
    const double v = inside ? u[d] : 0.5;  // walked at the centre; the result is replaced by the exact 0.0
    const uint32_t q = v < 1.0 ? (uint32_t)(v * 1073741824.0) : ((1u << kKeyBits) - 1u);  // floor(v*2^30), exact
    xs[(d * R + k) * NT + threadIdx.x] = v;  // all D dimensions: the last one is read by generic groups only
    if constexpr (V == 0) qs[(d * R + k) * NT + threadIdx.x] = q;
    if (d == D - 2) {
#pragma unroll
      for (int l = 2; l <= kLeafTable; ++l) {
        hat_cell(v, q, l > 3 ? l - 2 : 1, l > 2 ? 1u : 0u, pow2_lm1(l), pt.phA[k][l], pt.ceA[k][l]);
        pin_f64(pt.phA[k][l]);
        pin_u32(pt.ceA[k][l]);
      }
    }
    if (d == D - 1) {
      pt.ph[k][0] = pt.ph[k][1] = 1.0;
      pt.w[k][0] = pt.w[k][1] = 0u;
#pragma unroll
      for (int l = 2; l <= kLeafTable; ++l) {
        uint32_t cell;
        hat_cell(v, q, l > 3 ? l - 2 : 1, l > 2 ? 1u : 0u, pow2_lm1(l), pt.ph[k][l], cell);
        pt.w[k][l] = (l == 2 ? 1u : ((1u << (l - 2)) + 1u)) + cell;  // off(l) = 0,1,3,5,9,17,33,65
        pin_f64(pt.ph[k][l]);
        pin_u32(pt.w[k][l]);
      }
    }
  }
  return inside;
}

// Saved prefix products / positions of R points for dimensions 0..D-4. The running value `cur`
// (product over the dimensions walked so far) lives in its own registers: a record restarts it from
// the saved value of dimension k0-1, walks dimensions k0..D-3 and saves every dimension but the last,
// whose value is consumed at once by the group evaluation. V == 0: saved values in registers;
// V == 1: in shared memory (pt.Ps / pt.Os).
template <int D, int R, int V>
struct PrefixRegs {
  static constexpr int NS = (V == 0 && D > 3) ? D - 3 : 1;
  double P[R][NS];
  uint32_t O[R][NS];
};

// hat of the supporting node of level l >= 2 in dimension K for point k
template <int D, int R, int V>
__device__ __forceinline__ void hat_of(const StreamPoints<D, R, V> &pt, int K, int k, int b, uint32_t odd, double s,
                                       double &ph, uint32_t &cell) {
  const double xv = pt.x(K, k);
  cell = pt.cell(K, k, b, xv);
  const uint32_t i = (cell << 1) + odd;
  const double t = fma(xv, s, -u32_to_f64(i));  // x*2^(l-1) - i, one rounding as in the reference
  ph = 1.0 - fabs(t);
}

// stage K: multiply the running prefix by the hat of the level-l node of dimension K that supports
// the point (nothing to do for l == 1), then save it for later restarts unless K is the last prefix dim
template <int D, int R, int V, int K>
__device__ __forceinline__ void stage(const StreamPoints<D, R, V> &pt, PrefixRegs<D, R, V> &pr, double (&cur)[R],
                                      uint32_t (&cO)[R], int l, int b, uint32_t odd, double s) {
  constexpr int NT = StreamCfg<R, V>::kThreads;
  if (l > 1) {
    const uint32_t twob = 1u << b;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      double ph;
      uint32_t cell;
      hat_of<D, R, V>(pt, K, k, b, odd, s, ph, cell);
      cur[k] *= ph;
      cO[k] = cO[k] * twob + cell;  // (O << b) | cell as one IMAD
    }
  }
  if constexpr (K <= D - 4) {
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if constexpr (V == 0) {
        pr.P[k][K] = cur[k];
        pr.O[k][K] = cO[k];
      } else if constexpr (K == D - 4) {  // the deepest saved level stays in registers: half of all restarts use it
        pr.P[k][0] = cur[k];
        pr.O[k][0] = cO[k];
      } else {
        pt.Ps[(K * R + k) * NT + threadIdx.x] = cur[k];
        pt.Os[(K * R + k) * NT + threadIdx.x] = cO[k];
      }
    }
  }
}

template <int D, int R, int V, int K>
__device__ __forceinline__ void stage_from_levels(const StreamPoints<D, R, V> &pt, PrefixRegs<D, R, V> &pr,
                                                  double (&cur)[R], uint32_t (&cO)[R], uint32_t lev_lo,
                                                  uint32_t lev_hi) {
  const int l = (int)(((K < 6 ? lev_lo : lev_hi) >> (5 * (K % 6))) & 31u);
  stage<D, R, V, K>(pt, pr, cur, cO, l, bits_of(l), l > 2 ? 1u : 0u, pow2_lm1(l));
}

template <int D, int R, int V, int K>
__device__ __forceinline__ void restart(const StreamPoints<D, R, V> &pt, const PrefixRegs<D, R, V> &pr,
                                        double (&cur)[R], uint32_t (&cO)[R]) {
  constexpr int NT = StreamCfg<R, V>::kThreads;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    if constexpr (K == 0) { cur[k] = 1.0; cO[k] = 0u; }
    else if constexpr (V == 0) { cur[k] = pr.P[k][K - 1]; cO[k] = pr.O[k][K - 1]; }
    else if constexpr (K - 1 == D - 4) { cur[k] = pr.P[k][0]; cO[k] = pr.O[k][0]; }
    else {
      cur[k] = pt.Ps[((K - 1) * R + k) * NT + threadIdx.x];
      cO[k] = pt.Os[((K - 1) * R + k) * NT + threadIdx.x];
    }
  }
}

// FULL group from the shared-memory tile: 32-bit shared addresses, LDS gathers
template <int D, int R, int V>
__device__ __forceinline__ void group_full_smem(const StreamPoints<D, R, V> &pt, const double (&P)[R],
                                                const uint32_t (&O)[R], uint32_t tile_addr, uint32_t slot0,
                                                uint32_t npre8, int lm, double (&acc)[R]) {
  uint32_t a0[R];
  double inner[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    a0[k] = tile_addr + (slot0 + O[k]) * 8u;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(inner[k]) : "r"(a0[k]));  // level 1: phi = 1, cell = 0
  }
#pragma unroll
  for (int l = 2; l <= kLeafTable; ++l) {
    if (l > lm) break;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      double c;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(c) : "r"(a0[k] + pt.w[k][l] * npre8));
      inner[k] = fma(c, pt.ph[k][l], inner[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < R; ++k) acc[k] = fma(P[k], inner[k], acc[k]);
}

// FULL group whose slice is too large for a tile: same arithmetic, coefficients from global memory
template <int D, int R, int V>
__device__ __forceinline__ void group_full_global(const DeviceGrid &g, const StreamPoints<D, R, V> &pt,
                                                  const double (&P)[R], const uint32_t (&O)[R], uint32_t base1,
                                                  uint32_t npre8, int lm, double (&acc)[R]) {
  const double *A0[R];
  double inner[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    A0[k] = g.coef + (base1 + O[k]);
    inner[k] = __ldg(A0[k]);
  }
#pragma unroll
  for (int l = 2; l <= kLeafTable; ++l) {
    if (l > lm) break;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const double *p = reinterpret_cast<const double *>(reinterpret_cast<const char *>(A0[k]) +
                                                         (uint64_t)pt.w[k][l] * npre8);
      inner[k] = fma(__ldg(p), pt.ph[k][l], inner[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < R; ++k) acc[k] = fma(P[k], inner[k], acc[k]);
}

// SUPER record: levels 1..lmA of dimension D-2 x levels 1..lmB(lA) of dimension D-1, both unrolled
// against the register tables. P, O are the prefix product / position over dimensions 0..D-3.
// mask_rem < 0: no depth mask; otherwise budget left for the last two dimensions.
template <int D, int R, int V, bool SMEM>
__device__ __forceinline__ void group_super(const DeviceGrid &g, const StreamPoints<D, R, V> &pt, const double (&P)[R],
                                            const uint32_t (&O)[R], uint32_t sbase0, uint32_t base1, uint32_t npre8,
                                            int lmA, uint32_t lmBs, const uint4 &offs, int mask_rem,
                                            double (&acc)[R]) {
  // sbase0: SMEM: shared byte address of the first slice
  double inner2[R];
#pragma unroll
  for (int k = 0; k < R; ++k) inner2[k] = 0.0;
  const char *gbase0 = reinterpret_cast<const char *>(g.coef + base1);  // !SMEM
  if (mask_rem >= 0) lmA = min(lmA, mask_rem + 1);
#pragma unroll
  for (int lA = 1; lA <= kLeafTable; ++lA) {
    if (lA > lmA) break;
    constexpr int kDummy = 0;
    (void)kDummy;
    const int bA = lA == 1 ? 0 : (lA > 3 ? lA - 2 : 1);
    const uint32_t np = npre8 << bA;  // byte stride between consecutive last-dimension cells
    const int lmB_all = (int)((lmBs >> (4 * (lA - 1))) & 15u);
    const int lmB = mask_rem >= 0 ? min(lmB_all, mask_rem - (lA - 1) + 1) : lmB_all;
    // slice of group lA: precomputed by the packer as a 16-bit count of npre8-byte cells
    const uint32_t ow = lA <= 2 ? offs.x : (lA <= 4 ? offs.y : (lA <= 6 ? offs.z : offs.w));
    const uint32_t ocells = (lA & 1) ? (ow & 0xffffu) : (ow >> 16);
    const uint32_t sbase = sbase0 + ocells * npre8;
    const char *gbase = gbase0 + (uint64_t)ocells * npre8;
    uint32_t off[R];  // SMEM: shared byte address of the level-1 leaf; !SMEM: byte offset from gbase
    double inner1[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const uint32_t OA = lA == 1 ? O[k] : ((O[k] << bA) | pt.ceA[k][lA]);
      if constexpr (SMEM) {
        off[k] = OA * 8u + sbase;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(inner1[k]) : "r"(off[k]));
      } else {
        off[k] = OA * 8u;
        inner1[k] = __ldg(reinterpret_cast<const double *>(gbase + off[k]));
      }
    }
#pragma unroll
    for (int lB = 2; lB <= kLeafTable; ++lB) {
      if (lB > lmB) break;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        double c;
        if constexpr (SMEM) {
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(c) : "r"(pt.w[k][lB] * np + off[k]));  // one IMAD
        } else {
          c = __ldg(reinterpret_cast<const double *>(gbase + off[k] + (uint64_t)pt.w[k][lB] * np));
        }
        inner1[k] = fma(c, pt.ph[k][lB], inner1[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) inner2[k] = lA == 1 ? inner1[k] : fma(pt.phA[k][lA], inner1[k], inner2[k]);
  }
#pragma unroll
  for (int k = 0; k < R; ++k) acc[k] = fma(P[k], inner2[k], acc[k]);
}

// GENERIC group: explicit leaf records (any levels, dense or hashed), global memory
template <int D, int R, int V>
__device__ __forceinline__ void group_generic(const DeviceGrid &g, const StreamPoints<D, R, V> &pt, const double (&P)[R],
                                              const uint32_t (&O)[R], uint32_t jb, uint32_t n, int rem, int pb,
                                              double (&acc)[R]) {
  EvalSink<R> sink;
  Prefix<R> pf;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    sink.acc[k] = acc[k];
    pf.P[k] = P[k];
    pf.O[k] = O[k];
  }
  const TrieRec *rec = g.rec[D - 1];
  for (uint32_t j = jb; j < jb + n; ++j) {
    const TrieRec r = ld_rec(rec + j);
    const int l = (int)(r.y & 0xffu);
    if (l - 1 > rem) break;
    uint32_t cell[R];
    double ph[R];
    if (l > 1) {
      const int b = bits_of(l);
      const double s = pow2_lm1(l);
      const uint32_t odd = l > 2 ? 1u : 0u;
#pragma unroll
      for (int k = 0; k < R; ++k) hat_of<D, R, V>(pt, D - 1, k, b, odd, s, ph[k], cell[k]);
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) { cell[k] = 0u; ph[k] = 1.0; }
    }
    leaf_generic<R>(g, r, pf, cell, ph, pb, sink);
  }
#pragma unroll
  for (int k = 0; k < R; ++k) acc[k] = sink.acc[k];
}

template <int D, int R, bool MASKED, int V>
__global__ void __launch_bounds__(StreamCfg<R, V>::kThreads, StreamCfg<R, V>::kCtasPerSm)
    k_eval_stream(const DeviceGrid g, const EvalArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StreamSmem &sm = *reinterpret_cast<StreamSmem *>(smem_raw);
  constexpr int NT = StreamCfg<R, V>::kThreads;
  constexpr int NSV = D > 3 ? D - 3 : 1;  // saved prefix dimensions (V == 1: in shared memory)
  double *xs = reinterpret_cast<double *>(smem_raw + sizeof(StreamSmem));
  double *Ps = xs + (size_t)D * R * NT;                                                // V == 1
  uint32_t *Os = reinterpret_cast<uint32_t *>(Ps + (size_t)NSV * R * NT);             // V == 1
  uint32_t *qs = reinterpret_cast<uint32_t *>(xs + (size_t)D * R * NT);               // V == 0
  const int64_t tile_pts = (int64_t)blockDim.x * R;
  const int64_t stride = (int64_t)gridDim.x * tile_pts;
  const uint32_t ntiles = g.n_tiles;
  if (threadIdx.x == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phases = 0u;  // bit b: parity the next wait on buffer b expects

  for (int64_t base = (int64_t)blockIdx.x * tile_pts; base < a.M; base += stride) {
    StreamPoints<D, R, V> pt;
    pt.xs = xs;
    pt.qs = qs;
    pt.Ps = Ps;
    pt.Os = Os;
    bool inside[R];
    int64_t m[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      m[k] = base + (int64_t)k * blockDim.x + threadIdx.x;
      const int64_t mm = m[k] < a.M ? m[k] : a.M - 1;  // tail lanes redo the last point, never store
      inside[k] = stream_load_point<D, R, V>(g, a.X, mm, a.sp, a.sd, pt, k, xs, qs);
    }
    if (threadIdx.x == 0) {
      issue_tile(g, sm, 0, 0);
      if (ntiles > 1) issue_tile(g, sm, 1, 1);
    }
    PrefixRegs<D, R, V> pr;
    double acc[R];
#pragma unroll
    for (int k = 0; k < R; ++k) acc[k] = 0.0;
    int kmin = 0;  // saved prefix registers are valid for dimensions 0 .. kmin-1 (nothing at the start)

    for (uint32_t t = 0; t < ntiles; ++t) {
      const int buf = (int)(t & 1u);
      const StreamTile tl = g.tiles[t];
      mbar_wait(&sm.bar[buf], (phases >> buf) & 1u);
      phases ^= 1u << buf;
      // shared addresses of this tile, pinned so that they are not re-derived per record
      uint32_t rbase = smem_u32(sm.rec[buf]);
      uint32_t tbias = smem_u32(sm.coef[buf]) - tl.c0 * 8u;  // shared byte address of coefficient slot 0
      pin_u32(rbase);
      pin_u32(tbias);
      const uint32_t nrec = tl.r1 - tl.r0;
      for (uint32_t j = 0; j < nrec; ++j) {
        uint4 ra, rb, rc;  // {base1, npre8, meta, nchild} {lev_lo, lev_hi, s_hi, bod} {slice offsets}
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(ra.x), "=r"(ra.y), "=r"(ra.z), "=r"(ra.w) : "r"(rbase));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(rb.x), "=r"(rb.y), "=r"(rb.z), "=r"(rb.w) : "r"(rbase));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+32];" : "=r"(rc.x), "=r"(rc.y), "=r"(rc.z), "=r"(rc.w) : "r"(rbase));
        rbase += (uint32_t)sizeof(GroupRec);
        const int used = (int)((ra.z >> 16) & 0xffu);
        const int k0 = (int)((ra.z >> 8) & 0xffu);
        int kfrom = k0;
        if constexpr (MASKED) {
          if (used > a.rem) {  // depth mask of _interpolate_until_level: the whole group lies deeper
            kmin = min(kmin, k0);
            continue;
          }
          kfrom = min(k0, kmin);
          kmin = D;  // after this record every saved prefix register is valid
        }
        // running prefix: restart from the saved value below kfrom, walk dimensions kfrom .. D-3
        double cur[R];
        uint32_t cO[R];
        if constexpr (D >= 3) {
          kfrom = min(kfrom, D - 3);  // the value of dimension D-3 is never saved: always redo that stage
          switch (kfrom) {
            case 0: restart<D, R, V, 0>(pt, pr, cur, cO); break;
            case 1: if constexpr (D - 3 >= 1) restart<D, R, V, (D - 3 >= 1 ? 1 : 0)>(pt, pr, cur, cO); break;
            case 2: if constexpr (D - 3 >= 2) restart<D, R, V, (D - 3 >= 2 ? 2 : 0)>(pt, pr, cur, cO); break;
            case 3: if constexpr (D - 3 >= 3) restart<D, R, V, (D - 3 >= 3 ? 3 : 0)>(pt, pr, cur, cO); break;
            case 4: if constexpr (D - 3 >= 4) restart<D, R, V, (D - 3 >= 4 ? 4 : 0)>(pt, pr, cur, cO); break;
            case 5: if constexpr (D - 3 >= 5) restart<D, R, V, (D - 3 >= 5 ? 5 : 0)>(pt, pr, cur, cO); break;
            case 6: if constexpr (D - 3 >= 6) restart<D, R, V, (D - 3 >= 6 ? 6 : 0)>(pt, pr, cur, cO); break;
            case 7: if constexpr (D - 3 >= 7) restart<D, R, V, (D - 3 >= 7 ? 7 : 0)>(pt, pr, cur, cO); break;
            case 8: if constexpr (D - 3 >= 8) restart<D, R, V, (D - 3 >= 8 ? 8 : 0)>(pt, pr, cur, cO); break;
            default: if constexpr (D - 3 >= 9) restart<D, R, V, (D - 3 >= 9 ? 9 : 0)>(pt, pr, cur, cO); break;
          }
          switch (kfrom) {  // fall through: stages kfrom .. D-4 from the packed levels
            case 0: if constexpr (D - 4 >= 0) stage_from_levels<D, R, V, 0>(pt, pr, cur, cO, rb.x, rb.y);
            case 1: if constexpr (D - 4 >= 1) stage_from_levels<D, R, V, (D - 4 >= 1 ? 1 : 0)>(pt, pr, cur, cO, rb.x, rb.y);
            case 2: if constexpr (D - 4 >= 2) stage_from_levels<D, R, V, (D - 4 >= 2 ? 2 : 0)>(pt, pr, cur, cO, rb.x, rb.y);
            case 3: if constexpr (D - 4 >= 3) stage_from_levels<D, R, V, (D - 4 >= 3 ? 3 : 0)>(pt, pr, cur, cO, rb.x, rb.y);
            case 4: if constexpr (D - 4 >= 4) stage_from_levels<D, R, V, (D - 4 >= 4 ? 4 : 0)>(pt, pr, cur, cO, rb.x, rb.y);
            case 5: if constexpr (D - 4 >= 5) stage_from_levels<D, R, V, (D - 4 >= 5 ? 5 : 0)>(pt, pr, cur, cO, rb.x, rb.y);
            case 6: if constexpr (D - 4 >= 6) stage_from_levels<D, R, V, (D - 4 >= 6 ? 6 : 0)>(pt, pr, cur, cO, rb.x, rb.y);
            case 7: if constexpr (D - 4 >= 7) stage_from_levels<D, R, V, (D - 4 >= 7 ? 7 : 0)>(pt, pr, cur, cO, rb.x, rb.y);
            case 8: if constexpr (D - 4 >= 8) stage_from_levels<D, R, V, (D - 4 >= 8 ? 8 : 0)>(pt, pr, cur, cO, rb.x, rb.y);
            default: break;
          }
        } else {
#pragma unroll
          for (int k = 0; k < R; ++k) { cur[k] = 1.0; cO[k] = 0u; }
        }
        const int rem = MASKED ? a.rem - used : INT_MAX / 2;  // budget left below the prefix
        if (ra.z & kGrpSuper) {
          // stage D-3 (every record) from the fields the packer precomputed: s_hi, bod describe dimension D-3
          if constexpr (D >= 3) {
            const int l3 = (int)(((D - 3 < 6 ? rb.x : rb.y) >> (5 * ((D - 3) % 6))) & 31u);
            stage<D, R, V, D - 3>(pt, pr, cur, cO, l3, (int)(rb.w & 0xffu), (rb.w >> 8) & 1u, __hiloint2double((int)rb.z, 0));
          }
          const int lmA = (int)(ra.z & 0xffu);
          if (!(ra.z & kGrpGlobal))
            group_super<D, R, V, true>(g, pt, cur, cO, ra.x * 8u + tbias, ra.x, ra.y, lmA, ra.w, rc, MASKED ? rem : -1, acc);
          else
            group_super<D, R, V, false>(g, pt, cur, cO, 0u, ra.x, ra.y, lmA, ra.w, rc, MASKED ? rem : -1, acc);
          continue;
        }
        // per-group record (irregular parts): stage D-3 from the packed levels, stage D-2 from s_hi / bod
        if constexpr (D >= 3) stage_from_levels<D, R, V, (D >= 3 ? D - 3 : 0)>(pt, pr, cur, cO, rb.x, rb.y);
        {
          const int lA = (int)(((D - 2 < 6 ? rb.x : rb.y) >> (5 * ((D - 2) % 6))) & 31u);
          if (lA > 1) {
            const int b = (int)(rb.w & 0xffu);
            const uint32_t twob = 1u << b;
#pragma unroll
            for (int k = 0; k < R; ++k) {
              double ph;
              uint32_t cell;
              hat_of<D, R, V>(pt, D - 2, k, b, (rb.w >> 8) & 1u, __hiloint2double((int)rb.z, 0), ph, cell);
              cur[k] *= ph;
              cO[k] = cO[k] * twob + cell;
            }
          }
        }
        if (ra.z & kGrpFull) {
          const int lm = min((int)(ra.z & 0xffu), rem + 1);
          if (!(ra.z & kGrpGlobal)) group_full_smem<D, R, V>(pt, cur, cO, tbias, ra.x, ra.y, lm, acc);
          else group_full_global<D, R, V>(g, pt, cur, cO, ra.x, ra.y, lm, acc);
        } else {
          group_generic<D, R, V>(g, pt, cur, cO, ra.x, ra.w, rem, (int)((rb.w >> 16) & 0xffu), acc);
        }
      }
      __syncthreads();  // every warp is done with buffer `buf`
      if (threadIdx.x == 0 && t + 2 < ntiles) issue_tile(g, sm, t + 2, buf);
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (m[k] < a.M) {
        const double v = inside[k] ? acc[k] : 0.0;
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

cudaError_t launch_eval_stream(const DeviceGrid &g, const EvalArgs &a, cudaStream_t st, int64_t *launches) {
  if (a.M <= 0) return cudaSuccess;
  const int sms = sm_count();
  // two configurations: small batches -> one point per thread, saved prefixes in registers, 2 CTAs x 256
  // threads per SM; large batches -> two points per thread, saved prefixes in shared memory, 1 CTA x 512
  const int nsv = g.D > 3 ? g.D - 3 : 1;
  const size_t smem2 = sizeof(StreamSmem) + ((size_t)g.D * 8 + (size_t)nsv * 12) * 2 * StreamCfg<2, 1>::kThreads;
  bool two = a.M >= (int64_t)sms * StreamCfg<2, 1>::kThreads * 4 && smem2 <= 227 * 1024;
  if (const char *e = getenv("ASG_STREAM_R")) two = (e[0] == '2') && smem2 <= 227 * 1024;  // experiment knob
  const int R = two ? 2 : 1;
  const int threads = two ? StreamCfg<2, 1>::kThreads : StreamCfg<1, 0>::kThreads;
  const size_t smem = two ? smem2 : sizeof(StreamSmem) + (size_t)g.D * 12 * threads;
  const int per_block = threads * R;
  // persistent-style grid: every CTA re-streams the whole grid per batch of points, so use as few
  // batches as possible: one CTA per resident slot, grid-stride over the points
  const int resident = two ? StreamCfg<2, 1>::kCtasPerSm : StreamCfg<1, 0>::kCtasPerSm;
  int64_t need = (a.M + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sms * resident;
  const int blocks = (int)(need < 1 ? 1 : (need < cap ? need : cap));
  const bool masked = a.rem < INT_MAX / 4;
  cudaError_t lim = cudaSuccess;
  // (the shared memory of an instantiation depends on D, R only: its limit is set once per device)
#define LAUNCH(DD, RR, MM, VV)                                                                                    \
  do {                                                                                                            \
    static std::atomic<unsigned> limit_set{0};                                                                    \
    lim = smem_limit_once(k_eval_stream<DD, RR, MM, VV>, (int)smem, limit_set);                                   \
    if (lim == cudaSuccess) k_eval_stream<DD, RR, MM, VV><<<blocks, threads, smem, st>>>(g, a);                   \
  } while (0)
#define CALL(DD)                                                                                          \
  do {                                                                                                    \
    if (two) { if (masked) LAUNCH(DD, 2, true, 1); else LAUNCH(DD, 2, false, 1); }                        \
    else     { if (masked) LAUNCH(DD, 1, true, 0); else LAUNCH(DD, 1, false, 0); }                        \
  } while (0)
  ASG_DISPATCH_D2(g.D, CALL)
#undef CALL
#undef LAUNCH
  if (lim != cudaSuccess) return lim;
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace asg
