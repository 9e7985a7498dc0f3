// asg_device.cuh -- device-side building blocks shared by the subspace-engine kernels (sm_100a):
// the 1-D hat of the supporting node, hash probing of sparsely filled subspaces, the leaf sinks.
#pragma once
#include <climits>

#include "asg_internal.h"

namespace asg {


// ---- shared-memory addresses, mbarriers, 1-D TMA bulk copies --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// keep a table entry in its register: without this ptxas re-derives cells from the coordinate
// (DMUL + F2I + SEL per use) to save registers, which costs far more issue slots than it saves
__device__ __forceinline__ void pin_u32(uint32_t &v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void pin_f64(double &v) { asm volatile("" : "+d"(v)); }

__device__ __forceinline__ uint32_t hash_slot(uint32_t pos, int lg) {
  return lg == 0 ? 0u : (uint32_t)((pos * 0x9E3779B1u) >> (32 - lg));  // == asg_pack.cpp
}

__device__ __forceinline__ TrieRec ld_rec(const TrieRec *p) {
  const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
  TrieRec r;
  r.x = v.x;
  r.y = v.y;
  r.z = v.z;
  r.w = v.w;
  return r;
}

__device__ __forceinline__ double u32_to_f64(uint32_t i) {
  // (double) i without the slow I2F.F64: 2^52 + i has i in its low mantissa bits
  return __hiloint2double(0x43300000, (int)i) - 4503599627370496.0;
}

// 1-D hat of the supporting node of level l >= 2 at unit-cube coordinate x (q = its 30-bit cell key).
// b = cell bits of the level (uniform), s = 2^(l-1) (uniform).
__device__ __forceinline__ void hat_cell(double x, uint32_t q, int b, uint32_t odd, double s, double &ph,
                                         uint32_t &cell) {
  cell = q >> (kKeyBits - b);
  const uint32_t i = (cell << 1) + odd;
  const double t = fma(x, s, -u32_to_f64(i));  // x*2^(l-1) - i, one rounding as in the reference
  ph = 1.0 - fabs(t);                          // |t| <= 1 by construction of the cell
}

__device__ __forceinline__ int bits_of(int l) { return l > 3 ? l - 2 : 1; }                     // l >= 2
__device__ __forceinline__ double pow2_lm1(int l) { return __hiloint2double((1023 + l - 1) << 20, 0); }  // 2^(l-1)

// ---- sinks: what happens at a leaf (= one subspace) --------------------------------------------
template <int R>
struct EvalSink {
  double acc[R];
  __device__ __forceinline__ void dense(const DeviceGrid &g, int r, uint32_t slot, double P) {
    acc[r] = fma(__ldg(g.coef + slot), P, acc[r]);
  }
  __device__ __forceinline__ void hashed(const DeviceGrid &, int r, const HashEnt &e, uint32_t, double P) {
    acc[r] = fma(e.coef, P, acc[r]);
  }
  static constexpr bool kFast = true;  // may use the FULL-group fast path (plain sum of products)
};

template <int R>
struct BasisSink {
  int mode;  // 0 count, 1 fill, 2 dense
  double atol;
  int dropzeros;
  long long cnt[R];      // entries emitted so far in this row
  long long *colidx[R];  // row start (mode 1)
  double *vals[R];
  double *row[R];  // dense row base (mode 2)
  int64_t ldn;
  __device__ __forceinline__ void emit(int r, int32_t col, double v) {
    if (col < 0 || v == 0.0 || (dropzeros && v < atol)) return;  // src/class.jl:450-453
    if (mode == 1) {
      colidx[r][cnt[r]] = col;
      vals[r][cnt[r]] = v;
    } else if (mode == 2) {
      row[r][(int64_t)col * ldn] = v;
    }
    ++cnt[r];
  }
  __device__ __forceinline__ void dense(const DeviceGrid &g, int r, uint32_t slot, double P) {
    emit(r, __ldg(g.orig + slot), P);
  }
  __device__ __forceinline__ void hashed(const DeviceGrid &g, int r, const HashEnt &, uint32_t hslot, double P) {
    emit(r, __ldg(g.horig + hslot), P);
  }
  static constexpr bool kFast = false;
};

// per-thread state of R points
template <int D, int R>
struct Points {
  double x[R][D];
  uint32_t q[R][D];
  // last-dimension tables, levels 2..kLeafTable (level 1: phi = 1, w = 0)
  double ph[R][kLeafTable + 1];
  uint32_t w[R][kLeafTable + 1];  // off(l) + cell_l
};

template <int R>
struct Prefix {  // running product / position over the dimensions fixed so far
  double P[R];
  uint32_t O[R];
};

template <int R, class Sink>
__device__ __forceinline__ void leaf_generic(const DeviceGrid &g, const TrieRec &r, const Prefix<R> &pf,
                                             const uint32_t (&cell)[R], const double (&ph)[R], int pb, Sink &sink) {
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const uint32_t pos = (cell[k] << pb) | pf.O[k];
    const double P = pf.P[k] * ph[k];
    if (!(r.y & kRecHashed)) {
      sink.dense(g, k, r.x + pos, P);
    } else {
      const int lg = (int)((r.y >> 16) & 0xffu);
      const uint32_t mask = (1u << lg) - 1u;
      const unsigned long long key = (unsigned long long)pos + 1ull;
      uint32_t s = hash_slot(pos, lg);
      for (;;) {  // table is at most half full: an empty slot always ends the probe
        const uint32_t hs = r.x + s;
        const double2 raw = __ldg(reinterpret_cast<const double2 *>(g.hent + hs));
        HashEnt e;
        e.key = (unsigned long long)__double_as_longlong(raw.x);
        e.coef = raw.y;
        if (e.key == key) { sink.hashed(g, k, e, hs, P); break; }
        if (e.key == 0ull) break;
        s = (s + 1u) & mask;
      }
    }
  }
}

// leaves of one prefix group [jb, je) of trie level D-1, generic form (any levels, dense or hashed)
template <int D, int R, class Sink>
__device__ __forceinline__ void leaves_generic(const DeviceGrid &g, const Points<D, R> &pt, const Prefix<R> &pf,
                                               uint32_t jb, uint32_t je, int rem, int pb, Sink &sink) {
  const TrieRec *rec = g.rec[D - 1];
  for (uint32_t j = jb; j < je; ++j) {
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
      for (int k = 0; k < R; ++k) hat_cell(pt.x[k][D - 1], pt.q[k][D - 1], b, odd, s, ph[k], cell[k]);
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) { cell[k] = 0u; ph[k] = 1.0; }
    }
    leaf_generic<R>(g, r, pf, cell, ph, pb, sink);
  }
}

// FULL prefix group: levels 1..lm of the last dimension, dense, back to back (see asg_internal.h)
template <int D, int R>
__device__ __forceinline__ void leaves_full(const DeviceGrid &g, const Points<D, R> &pt, const Prefix<R> &pf,
                                            uint32_t base1, int pb, int lm, EvalSink<R> &sink) {
  const double *A0[R];
  double inner[R];
  const uint32_t npre8 = 8u << pb;  // byte stride between consecutive last-dimension cells
#pragma unroll
  for (int k = 0; k < R; ++k) {
    A0[k] = g.coef + (base1 + pf.O[k]);
    inner[k] = __ldg(A0[k]);  // level 1: phi = 1, cell = 0
  }
#pragma unroll
  for (int l = 2; l <= kLeafTable; ++l) {
    if (l > lm) break;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const double *p;  // A0 + w * npre8 bytes in ONE instruction (IMAD.WIDE.U32)
      asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(pt.w[k][l]), "r"(npre8), "l"(A0[k]));
      inner[k] = fma(__ldg(p), pt.ph[k][l], inner[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < R; ++k) sink.acc[k] = fma(pf.P[k], inner[k], sink.acc[k]);
}


}  // namespace asg
