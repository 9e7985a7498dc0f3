// asg_kernels_block.cuh -- BLOCK walk of the subspace engine (sm_100a): the main evaluation kernel for
// grids that are (nearly) regular in their last three dimensions (asg_block.h).
//
//   G(x) = sum_n c_n prod_d phi(x01_d; l_nd, i_nd)                       (src/class.jl:629-643)
//
// * One thread owns R = 2 query points, a CTA of 512 threads owns 1024 points and walks the whole record
//   stream for them; one CTA per SM, grid-stride over the points.
// * Records (32 B, pre-decoded, one per prefix level vector) and the coefficient blocks they reference are streamed
//   from L2 into shared memory by 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) completed on
//   mbarriers, double buffered. A tile starts with a header record (its extent and the extent of the tile that takes
//   the buffer next), so the walk reads no table from global memory; the warps are not synchronised per tile: each
//   counts itself out of a buffer and the last one out requests the refill.
// * The points of a batch arrive as asynchronous 8-byte copies (cp.async, SASS LDGSTS), D lanes per row.
// * A record applies ONE data-driven stage (the level of its last non-trivial prefix dimension) to the
//   prefix product / position saved by its parent record (a small stack in shared memory), then runs
//   the COMPILED walk of the last three dimensions: the kernel is templated on the suffix budget, all
//   levels and slot offsets are compile-time constants, so a term costs an address add, one LDS.64 and
//   one DFMA, a pole (fixed levels of all dimensions but the last) one IMAD and one DFMA more.
// * The hats of levels 2..4 of the last two dimensions live in per-point register tables; levels 5..8
//   (286 of the 19 448 terms per point on the D = 10 benchmark grid) are formed on the fly.
// The 1-D hat is the reference's phi (src/math.jl:345-357) evaluated on its support only, with the
// reference's single rounding (asg_device.cuh). Out-of-domain / NaN points give exactly 0.0.
#pragma once
#include <cstdlib>

#include "asg_device.cuh"

namespace asg {

#ifndef ASG_BLK_THREADS
#define ASG_BLK_THREADS 512  // build knob: 384 gives 168 registers per thread (3 warps per sub-partition); see DESIGN.md section 8
#endif
constexpr int kBlkThreads = ASG_BLK_THREADS;  // CTA size of the walk
// threads of the CTA as a function of the points per thread: a CTA always owns 1 024 points at R >= 2
__host__ __device__ constexpr int blk_nt(int R) { return R == 4 ? 256 : kBlkThreads; }
constexpr int kBlkTab = 4;  // levels 2..kBlkTab of dimensions B, C are tabulated in registers
// PAIR mode (two points per thread, large ordered batches): the two points of a thread are NEIGHBOURS in the cell-key
// order and share the cell of every level <= kBlkTab in EVERY dimension (checked per thread; the second point of a pair
// that does not is evaluated by a follow-up launch). Wherever all levels on the path to a coefficient are <= kBlkTab
// the two points read the SAME coefficient: one address, one LDS.64, two DFMA ("SH" below). 16 588 of the 19 448
// terms of the D = 10 benchmark grid are of that kind.
static_assert(kBlkMaxBudget == 9, "the dispatch lists the suffix budgets 0..9");
static_assert(kBlkTab == kBlkPairLevel, "the packer flags shared records with the level the tables cover");

static_assert(kBlkTileBufs == 2, "BlkSmem, the phase bits and the buffer index are written for two buffers");
struct __align__(16) BlkSmem {
  BlockRec rec[2][kBlkTileRecs + 1];  // a tile's header record (blk_tile_header, asg_block.h) + its records
  unsigned long long bar[2];
  unsigned done[2];  // warps that have left the tile in buffer b (wraps at the warp count: the last one out refills it)
  unsigned pad_[2];
  BlockTile first[2];  // tiles 0 and 1, which open every batch
  // followed by  double coef[2][tile_coefs]  (the two coefficient tile buffers; size chosen by the packer),
  //              double xs[D][R][threads]  (unit-cube coordinates of every dimension),
  //              double Ps[slots][R][threads], uint32_t Os[slots][R][threads]  (saved prefixes)
};

template <int R, bool PAIR>
struct BlkPoint {
  static constexpr int RC = PAIR ? 1 : R;  // copies of the cell tables: the points of a pair share them
  double phC[R][kBlkTab + 1];     // last dimension: hat of level l
  uint32_t wC8[RC][kBlkTab + 1];  // last dimension: (off1(l) + cell_l) * 8
  double phB[R][kBlkTab + 1];     // dimension D-2: hat of level l
  uint32_t cB[RC][kBlkTab + 1];   // dimension D-2: cell of level l
  double xA[R];                   // unit-cube coordinate of dimension D-3 (its hats are formed on the fly)
  uint32_t qA[R];                 // and its 30-bit cell key
  uint32_t xsBC;                  // shared byte address of this thread's coordinate of dimension D-2 (point 0)
};

// coordinate of dimension D-2+which (which = 0, 1: B, C at rare levels; which < 0: dimensions above A of the
// deeper compiled suffixes) of point k, from shared memory
template <int R, bool PAIR>
__device__ __forceinline__ double blk_x_rare(const BlkPoint<R, PAIR> &pt, int which, int k) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(pt.xsBC + (uint32_t)((which * R + k) * (blk_nt(R) * 8))));
  return v;
}
__device__ __forceinline__ uint32_t blk_key(double v) {  // floor(v * 2^30), clamped for v == 1 (exact)
  return v < 1.0 ? (uint32_t)(v * 1073741824.0) : ((1u << kKeyBits) - 1u);
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
template <int IMM>
__device__ __forceinline__ double lds_f64_imm(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(IMM));
  return v;
}

// hat and cell of the supporting node of compile-time level L >= 2
template <int L>
__device__ __forceinline__ void hat_ct(double x, uint32_t q, double &ph, uint32_t &cell) {
  constexpr int b = blk_bits(L);
  cell = q >> (kKeyBits - b);
  const uint32_t i = (cell << 1) + (L > 2 ? 1u : 0u);
  const double t = fma(x, (double)(1u << (L - 1)), -u32_to_f64(i));  // x*2^(l-1) - i, one rounding (math.jl:347)
  ph = 1.0 - fabs(t);
}

// the same from the coordinate alone (dimensions whose key is not kept in a register)
template <int L>
__device__ __forceinline__ uint32_t cell_ct_x(double x) {
  constexpr int b = blk_bits(L);
  return min((uint32_t)(x * (double)(1u << b)), (1u << b) - 1u);  // floor(x*2^b), clamped for x == 1 (exact)
}
template <int L>
__device__ __forceinline__ void hat_ct_x(double x, double &ph, uint32_t &cell) {
  cell = cell_ct_x<L>(x);
  const uint32_t i = (cell << 1) + (L > 2 ? 1u : 0u);
  const double t = fma(x, (double)(1u << (L - 1)), -u32_to_f64(i));
  ph = 1.0 - fabs(t);
}
// hat of level L at x when the cell is known (a pair's common cell): fi = (double)(2*cell + [L > 2])
template <int L>
__device__ __forceinline__ double hat_ct_cell(double x, double fi) {
  return 1.0 - fabs(fma(x, (double)(1u << (L - 1)), -fi));
}

template <int L, int R, bool PAIR>
__device__ __forceinline__ void get_C(const BlkPoint<R, PAIR> &pt, int k, double &ph, uint32_t &w8) {
  if constexpr (L <= kBlkTab) {
    ph = pt.phC[k][L];
    w8 = pt.wC8[PAIR ? 0 : k][L];
  } else {
    uint32_t cell;
    hat_ct_x<L>(blk_x_rare<R, PAIR>(pt, 1, k), ph, cell);
    w8 = (blk_off1(L) + cell) * 8u;
  }
}
template <int L, int R, bool PAIR>
__device__ __forceinline__ void get_B(const BlkPoint<R, PAIR> &pt, int k, double &ph, uint32_t &cell) {
  if constexpr (L <= kBlkTab) {
    ph = pt.phB[k][L];
    cell = pt.cB[PAIR ? 0 : k][L];
  } else {
    hat_ct_x<L>(blk_x_rare<R, PAIR>(pt, 0, k), ph, cell);
  }
}

// ---- compiled suffix walk ---------------------------------------------------------------------------
// SH (PAIR mode only): every level on the path so far is <= kBlkTab, so the R points share the address: only
// element [0] of an address array is meaningful and one load serves all points.
// pole: levels 1..RP+1 of the last dimension at shared byte address tp[k] + IMM
template <int RP, int LC, int IMM, int R, bool MASKED, bool PAIR, bool SH>
__device__ __forceinline__ void blk_pole(const BlkPoint<R, PAIR> &pt, const uint32_t (&tp)[R], int reff, double (&inner)[R]) {
  if constexpr (LC <= RP + 1) {
    if (MASKED && LC - 1 > reff) return;
    if constexpr (SH && LC <= kBlkTab) {
      const double c = lds_f64_imm<IMM>(tp[0] + pt.wC8[0][LC]);
#pragma unroll
      for (int k = 0; k < R; ++k) inner[k] = fma(c, pt.phC[k][LC], inner[k]);
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        double ph;
        uint32_t w8;
        get_C<LC, R, PAIR>(pt, k, ph, w8);
        const double c = lds_f64_imm<IMM>(tp[SH ? 0 : k] + w8);
        inner[k] = fma(c, ph, inner[k]);
      }
    }
    blk_pole<RP, LC + 1, IMM, R, MASKED, PAIR, SH>(pt, tp, reff, inner);
  }
}

// 2-D block of storage budget RB at shared byte address t2[k] + IMM: levels LB.. of dimension B (Horner form:
// one DFMA per pole, none for level 1)
template <int RB, int LB, int IMM, int R, bool MASKED, bool PAIR, bool SH>
__device__ __forceinline__ void blk_suffix2(const BlkPoint<R, PAIR> &pt, const uint32_t (&t2)[R], int reff, double (&accA)[R]) {
  if constexpr (LB <= RB + 1) {
    if (MASKED && LB - 1 > reff) return;
    constexpr int RP = RB - (LB - 1);
    constexpr int IMM2 = IMM + (int)blk_offB(LB, RB) * 8;
    constexpr bool SHN = SH && LB <= kBlkTab;  // the pole's address is still common to the points
    uint32_t tp[R];
    double phB[R], inner[R];
    if constexpr (SHN) {
      if constexpr (LB == 1) {
        tp[0] = t2[0];
      } else {
        tp[0] = t2[0] + pt.cB[0][LB] * (blk_n1(RP) * 8u);
#pragma unroll
        for (int k = 0; k < R; ++k) phB[k] = pt.phB[k][LB];
      }
      const double c = lds_f64_imm<IMM2>(tp[0]);  // level 1 of the last dimension: phi = 1, cell = 0
#pragma unroll
      for (int k = 0; k < R; ++k) inner[k] = c;
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if constexpr (LB == 1) {
          tp[k] = t2[k];
        } else {
          uint32_t cell;
          get_B<(LB > 1 ? LB : 2), R, PAIR>(pt, k, phB[k], cell);
          tp[k] = t2[SH ? 0 : k] + cell * (blk_n1(RP) * 8u);
        }
        inner[k] = lds_f64_imm<IMM2>(tp[k]);  // level 1 of the last dimension: phi = 1, cell = 0
      }
    }
    blk_pole<RP, 2, IMM2, R, MASKED, PAIR, SHN>(pt, tp, reff - (LB - 1), inner);
#pragma unroll
    for (int k = 0; k < R; ++k) accA[k] = LB == 1 ? inner[k] : fma(phB[k], inner[k], accA[k]);
    blk_suffix2<RB, LB + 1, IMM, R, MASKED, PAIR, SH>(pt, t2, reff, accA);
  }
}

// compiled suffix of S >= 3 dimensions (D-S .. D-1) of storage budget RS at shared byte address t[k] + IMM:
// levels L.. of dimension D-S. S = 3: that dimension is A (coordinate and key in registers); S > 3: its
// coordinate comes from shared memory (those records are few and small, see blk_max_r).
template <int S, int RS, int L, int IMM, int R, bool MASKED, bool PAIR, bool SH>
__device__ __forceinline__ void blk_suffixN(const BlkPoint<R, PAIR> &pt, const uint32_t (&t)[R], int reff, double (&out)[R]) {
  if constexpr (L <= RS + 1) {
    if (MASKED && L - 1 > reff) return;
    constexpr int RB = RS - (L - 1);
    constexpr int IMM2 = IMM + (int)blk_off(S, L, RS) * 8;
    constexpr bool SHN = SH && L <= kBlkTab;
    constexpr int LL = L > 1 ? L : 2;
    uint32_t t2[R];
    double ph[R], sub[R];
    if constexpr (L == 1) {
#pragma unroll
      for (int k = 0; k < R; ++k) t2[k] = t[k];
    } else if constexpr (SHN) {
      // the points' common cell comes from point 0; each point forms its own hat on it
      uint32_t cell;
      double x[R];
#pragma unroll
      for (int k = 0; k < R; ++k) x[k] = S == 3 ? pt.xA[k] : blk_x_rare<R, PAIR>(pt, 2 - S, k);
      if constexpr (S == 3) cell = pt.qA[0] >> (kKeyBits - blk_bits(LL));
      else cell = cell_ct_x<LL>(x[0]);
      const double fi = u32_to_f64((cell << 1) + (LL > 2 ? 1u : 0u));
#pragma unroll
      for (int k = 0; k < R; ++k) ph[k] = hat_ct_cell<LL>(x[k], fi);
      t2[0] = t[0] + cell * (blk_size(S - 1, RB) * 8u);
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        uint32_t cell;
        if constexpr (S == 3) {
          hat_ct<LL>(pt.xA[k], pt.qA[k], ph[k], cell);
        } else {
          // dimension D-S is S-2 rows above dimension D-2 in xs
          hat_ct_x<LL>(blk_x_rare<R, PAIR>(pt, 2 - S, k), ph[k], cell);
        }
        t2[k] = t[SH ? 0 : k] + cell * (blk_size(S - 1, RB) * 8u);
      }
    }
    if constexpr (S == 3) blk_suffix2<RB, 1, IMM2, R, MASKED, PAIR, SHN>(pt, t2, reff - (L - 1), sub);
    else blk_suffixN<S - 1, RB, 1, IMM2, R, MASKED, PAIR, SHN>(pt, t2, reff - (L - 1), sub);
#pragma unroll
    for (int k = 0; k < R; ++k) out[k] = L == 1 ? sub[k] : fma(ph[k], sub[k], out[k]);
    blk_suffixN<S, RS, L + 1, IMM, R, MASKED, PAIR, SH>(pt, t, reff, out);
  }
}

// ---- the data-driven stage of a record ---------------------------------------------------------------
template <int R>
struct BlkStaged {     // a record after its stage: everything its compiled suffix needs
  double cur[R];       // prefix product over dimensions 0..D-4 (S2: and A)
  uint32_t t[R];       // cO * chunk bytes is formed by the suffix; here: prefix position
  uint32_t sw;         // suffix code (blk_code(S, r), kBrNone), flags, used
  uint32_t cb;         // shared byte address of the record's chunk 0
};

// Fetch the record at shared address rbase and apply its stage: hat of the level-l node of dimension k0
// that supports the point, on top of the prefix saved by the parent record (stack slot ra.z).
template <int R, bool MASKED, int NT>
__device__ __forceinline__ void blk_stage(uint32_t rbase, uint32_t ps_t, uint32_t os_t, uint32_t xs_t, uint32_t tbias,
                                          int rem, BlkStaged<R> &out, double (&xpre)[R], uint32_t kmax) {
  uint4 ra, rb;  // {cbase, sw, slot, k0} {b, s_hi, lim, odd}
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(ra.x), "=r"(ra.y), "=r"(ra.z), "=r"(ra.w) : "r"(rbase));
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(rb.x), "=r"(rb.y), "=r"(rb.z), "=r"(rb.w) : "r"(rbase));
  const uint32_t pa = ps_t + ra.z * (R * NT * 8u), oa = os_t + ra.z * (R * NT * 4u);
  uint32_t Oin[R];
#pragma unroll
  for (int k = 0; k < R; ++k) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Oin[k]) : "r"(oa + k * (NT * 4u)));
  // (keeping the inputs of sibling records -- same parent, same stage dimension -- in registers instead of
  // re-loading them was measured 4 % SLOWER: the eight registers are worth more to the compiled suffix)
  double Pin[R], xin[R];
#ifdef ASG_BLK_PREFETCH_X
  // the coordinate of the stage dimension was fetched while the previous record's suffix ran (coordinates never change
  // during a walk): one shared-memory round trip less on the record's critical path. Now fetch the next record's.
  (void)kmax;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    Pin[k] = lds_f64(pa + k * (NT * 8u));
    xin[k] = xpre[k];
  }
  {
    uint32_t k0n;
    asm volatile("ld.shared.u32 %0, [%1+44];" : "=r"(k0n) : "r"(rbase));  // k0 of the next record (32 + 12)
    k0n = min(k0n, kmax);  // past the tile's last record: any valid row
    const uint32_t xn = xs_t + k0n * (R * NT * 8u);
#pragma unroll
    for (int k = 0; k < R; ++k) xpre[k] = lds_f64(xn + k * (NT * 8u));
  }
#else
  (void)xpre; (void)kmax;
  const uint32_t xa = xs_t + ra.w * (R * NT * 8u);
#pragma unroll
  for (int k = 0; k < R; ++k) {
    Pin[k] = lds_f64(pa + k * (NT * 8u));
    xin[k] = lds_f64(xa + k * (NT * 8u));
  }
#endif
  const double s = __hiloint2double((int)rb.y, 0);                            // 2^(l-1), or 0 for the root record
  const double s2b = __hiloint2double((int)((rb.x << 20) + 0x3ff00000u), 0);  // 2^b
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const uint32_t cell = min((uint32_t)(xin[k] * s2b), rb.z);  // floor(x*2^b), clamped for x == 1 (exact)
    const double tt = fma(xin[k], s, -u32_to_f64((cell << 1) + rb.w));  // x*2^(l-1) - i, one rounding (math.jl:347)
    out.cur[k] = Pin[k] * (1.0 - fabs(tt));
    out.t[k] = (Oin[k] << rb.x) | cell;
  }
  if (ra.y & kBrSave) {
#pragma unroll
    for (int k = 0; k < R; ++k) {
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(pa + (R + k) * (NT * 8u)), "d"(out.cur[k]));
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(oa + (R + k) * (NT * 4u)), "r"(out.t[k]));
    }
  }
  out.sw = ra.y;
  if (MASKED && (int)(ra.y >> kBrUsedShift) > rem) out.sw = (ra.y & ~0xffu) | kBrNone;  // depth mask: lies deeper
  out.cb = ra.x * 8u + tbias;
}

// One record: the compiled suffix of the staged record. (Staging the NEXT record inside the same basic block,
// so that ptxas interleaves the two dependency chains, was measured 5 % slower: registers.)
template <int S, int RS, int R, bool MASKED, bool PAIR, bool SH>
__device__ __forceinline__ void blk_record(const BlkPoint<R, PAIR> &pt, const BlkStaged<R> &st, int rem, double (&acc)[R]) {
  uint32_t t[R];
#pragma unroll
  for (int k = 0; k < R; ++k) t[k] = st.t[k] * (blk_size(S, RS) * 8u) + st.cb;
  const int reff = MASKED ? rem - (int)(st.sw >> kBrUsedShift) : 0;
  double v[R];
  if constexpr (S == 2) blk_suffix2<RS, 1, 0, R, MASKED, PAIR, SH>(pt, t, reff, v);
  else blk_suffixN<S, RS, 1, 0, R, MASKED, PAIR, SH>(pt, t, reff, v);
#pragma unroll
  for (int k = 0; k < R; ++k) acc[k] = fma(st.cur[k], v[k], acc[k]);
}

// the compiled suffix of a staged record, chosen by its code. PAIR kernels compile the unshared form (a prefix level
// above kBlkTab) only for the small budgets such a prefix can leave on the grids that take PAIR mode (kBlkPairMaxR).
// WIDE kernels additionally compile the suffix budgets 8 and 9 (levels up to 10, e.g. BASELINE configs[4]: D = 4,
// depth 10); they are separate instantiations because the larger bodies cost the common ones registers (the D = 10
// benchmark grid walked 4.5 % slower with them in one kernel).
template <int R, bool MASKED, bool PAIR, bool SH, bool WIDE>
__device__ __forceinline__ void blk_dispatch(const BlkPoint<R, PAIR> &pt, const BlkStaged<R> &st, int rem, double (&acc)[R]) {
  // (pairs this build does not compile -- RS above blk_max_r(S), 3-D chunks no tile can hold -- get labels outside
  // the code range)
#define BLK_CASE(S, RS)                                                                                      \
  case ((S <= 3 || RS <= blk_max_r(S)) ? blk_code(S, RS) : 160u + 10u * S + RS):                             \
    if constexpr ((S <= 3 || RS <= blk_max_r(S)) && (!PAIR || SH || RS <= kBlkPairMaxR) &&                   \
                  (S != 3 || blk_size(3, RS) + 2u <= (uint32_t)kBlkTileCoefsMax) && (WIDE || RS <= 7))       \
      blk_record<S, RS, R, MASKED, PAIR, SH>(pt, st, rem, acc);                                              \
    break;
  // a quarter of the records of a regular grid are single terms (3-D suffix, r = 0): test that first; every other
  // compiled suffix goes through ONE indexed branch (the codes are dense, asg_block.h)
  const uint32_t code = st.sw & 0xffu;
  if (code == blk_code(3, 0)) blk_record<3, 0, R, MASKED, PAIR, SH>(pt, st, rem, acc);
  else switch (code) {
    BLK_CASE(3, 1) BLK_CASE(3, 2) BLK_CASE(3, 3) BLK_CASE(3, 4) BLK_CASE(3, 5) BLK_CASE(3, 6) BLK_CASE(3, 7) BLK_CASE(3, 8)
    BLK_CASE(3, 9)
    BLK_CASE(2, 0) BLK_CASE(2, 1) BLK_CASE(2, 2) BLK_CASE(2, 3) BLK_CASE(2, 4) BLK_CASE(2, 5) BLK_CASE(2, 6) BLK_CASE(2, 7)
    BLK_CASE(2, 8) BLK_CASE(2, 9)
    BLK_CASE(4, 1) BLK_CASE(4, 2) BLK_CASE(4, 3) BLK_CASE(4, 4) BLK_CASE(4, 5)
    BLK_CASE(5, 1) BLK_CASE(5, 2) BLK_CASE(5, 3) BLK_CASE(5, 4)
    BLK_CASE(6, 1) BLK_CASE(6, 2) BLK_CASE(6, 3) BLK_CASE(6, 4)
    BLK_CASE(7, 1) BLK_CASE(7, 2) BLK_CASE(7, 3)
    BLK_CASE(8, 1) BLK_CASE(8, 2)
    default: break;  // kBrNone: a stage-only record
  }
#undef BLK_CASE
}

// producer side (one thread): start the copies of a tile -- rbytes of records from device record r0 on (header first),
// cc coefficient slots from c0 on -- into buffer `buf`
__device__ __forceinline__ void blk_issue_tile(const DeviceGrid &g, BlkSmem &sm, uint32_t r0, uint32_t rbytes, uint32_t c0,
                                               uint32_t cc, int buf) {
  const uint32_t cbytes = cc * 8u;
  // the buffer was last read through the generic proxy: order those reads before the async-proxy writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(&sm.bar[buf], rbytes + cbytes);
  tma_load_1d(sm.rec[buf], g.brec + r0, rbytes, &sm.bar[buf]);
  if (cbytes)
    tma_load_1d(reinterpret_cast<double *>(&sm + 1) + (size_t)buf * g.blk_tile_coefs, g.bcoef + c0, cbytes, &sm.bar[buf]);
}
__device__ __forceinline__ void blk_issue_tile(const DeviceGrid &g, BlkSmem &sm, const BlockTile tl, int buf) {
  blk_issue_tile(g, sm, tl.r0, (tl.r1 - tl.r0) * (uint32_t)sizeof(BlockRec), tl.c0, tl.c1 - tl.c0, buf);
}

// registers per thread: the whole file for one CTA per SM (65536 / threads, in units of 8)
template <int R, bool MASKED, int NT, bool PAIR, bool WIDE>
__global__ void __maxnreg__(NT == 512 ? 128 : (NT == 256 ? 255 : 168)) k_eval_block(const DeviceGrid g, const EvalArgs a) {
  static_assert(!(PAIR && WIDE), "grids with suffix budgets above 7 do not take PAIR mode");
  static_assert(!PAIR || R >= 2, "PAIR mode shares loads between the points of a thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BlkSmem &sm = *reinterpret_cast<BlkSmem *>(smem_raw);
  const int D = g.D;
  const int K = D - 3;
  double *coefbuf = reinterpret_cast<double *>(smem_raw + sizeof(BlkSmem));
  double *xs = coefbuf + 2 * (size_t)g.blk_tile_coefs;
  double *Ps = xs + (size_t)D * R * NT;
  uint32_t *Os = reinterpret_cast<uint32_t *>(Ps + (size_t)g.blk_slots * R * NT);
  const uint32_t tid = threadIdx.x;
  // shared byte addresses of this thread's column in xs / Ps / Os
  uint32_t xs_t = smem_u32(xs) + tid * 8u;
  uint32_t ps_t = smem_u32(Ps) + tid * 8u;
  uint32_t os_t = smem_u32(Os) + tid * 4u;
  const int64_t M = a.M_dev ? (int64_t)*a.M_dev : a.M;  // follow-up launch of PAIR mode: the count lives on the device
  const int64_t tile_pts = (int64_t)NT * R;
  const int64_t stride = (int64_t)gridDim.x * tile_pts;
  const uint32_t ntiles = g.n_btiles;
  if ((int64_t)blockIdx.x * tile_pts >= M) return;
  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    sm.done[0] = sm.done[1] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {  // slot 0: the empty prefix (never overwritten)
    Ps[k * NT + tid] = 1.0;
    Os[k * NT + tid] = 0u;
  }
  if (tid < 2u && tid < ntiles) sm.first[tid] = g.btiles[tid];
  __syncthreads();
  uint32_t phases = 0u;  // bit b: parity the next wait on buffer b expects
  // row-wise fetch of the points (see the loader below): lane = row_sub * D + row_d fetches coordinate row_d of the
  // row_sub-th of the 32 / D points one instruction covers
  const uint32_t row_ppi = 32u / (uint32_t)D;
  const uint32_t row_sub = (tid & 31u) / (uint32_t)D, row_d = (tid & 31u) - row_sub * (uint32_t)D;
  const uint32_t xs_w = xs_t - (tid & 31u) * 8u;  // column of lane 0 of this warp

  for (int64_t base = (int64_t)blockIdx.x * tile_pts; base < M; base += stride) {
    BlkPoint<R, PAIR> pt;
    pt.xsBC = xs_t + (uint32_t)((K + 1) * R * NT * 8);
    static_assert(NT == blk_nt(R), "blk_x_rare derives the CTA size from R");
    // flags: bit k: point k lies in the domain; bit 8 + k: point k's result is stored by this launch
    uint32_t flags = ((1u << R) - 1u) << 8;
    uint32_t code2[R];     // PAIR: the cells of level <= kBlkTab of every dimension (2 bits each)
    // The first two tiles are requested before anything else of the batch: they arrive while the points are fetched.
    if (tid == 0) {
      blk_issue_tile(g, sm, sm.first[0], 0);
      if (ntiles > 1) blk_issue_tile(g, sm, sm.first[1], 1);
    }
    // The raw coordinates travel global -> shared as asynchronous 8-byte copies, ALL of them in flight at once (the
    // rows of an ordered batch are a random gather; loaded one by one in front of scale()'s divide they cost one DRAM
    // latency per coordinate). Row-major points: D lanes fetch one row, so that a warp instruction touches 32 / D rows
    // (3 rows of 80 bytes for D = 10) instead of 32 different ones, each lane copying into the column of the thread that
    // owns the point. Each thread then transforms its own column in place.
    int64_t mm[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      // PAIR: a thread owns R NEIGHBOURS of the cell-key order; otherwise its points are a CTA width apart
      const int64_t j = PAIR ? base + R * (int64_t)tid + k : base + (int64_t)k * NT + tid;
      const int64_t jj = j < M ? j : M - 1;  // tail lanes redo the last point, never store
      if (j >= M) flags &= ~(0x100u << k);
      mm[k] = a.perm ? (int64_t)a.perm[jj] : jj;
    }
    if (a.sd == 1) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        for (uint32_t i = 0; i < 32u; i += row_ppi) {
          const uint32_t src = i + row_sub;  // lane that owns the point this lane helps to fetch
          const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)mm[k], (int)(src & 31u));
          const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)((uint64_t)mm[k] >> 32), (int)(src & 31u));
          if (row_sub < row_ppi && src < 32u) {
            const double *p = a.X + (int64_t)(((uint64_t)hi << 32) | lo) * a.sp + row_d;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(xs_w + src * 8u + (uint32_t)((row_d * R + k) * NT) * 8u), "l"(p) : "memory");
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const double *row = a.X + mm[k] * a.sp;
        for (int d = 0; d < D; ++d)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(xs_t + (uint32_t)((d * R + k) * NT) * 8u), "l"(row + d * a.sd) : "memory");
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();  // the copies into a thread's column were made by other lanes of its warp
#pragma unroll
    for (int k = 0; k < R; ++k) {
      bool in = true;
      uint32_t c2 = 0u;
      for (int d = 0; d < D; ++d) {
        const double xd = xs[(d * R + k) * NT + tid];
        const double u = __dadd_rn(0.0, __ddiv_rn(__dsub_rn(xd, g.lb[d]), g.width[d]));  // scale(), src/math.jl:308
        in = in && (u >= 0.0) && (u <= 1.0);                                             // NaN fails both
        xs[(d * R + k) * NT + tid] = u;
        if (PAIR) c2 = (c2 << 2) | min((uint32_t)(u * 4.0), 3u);
      }
      if (in) flags |= 1u << k;
      else {  // walked at the centre of the cube; the result is replaced by the exact 0.0
        for (int d = 0; d < D; ++d) xs[(d * R + k) * NT + tid] = 0.5;
        c2 = 0xaaaaaaaau & ((D >= 16) ? 0xffffffffu : ((1u << (2 * D)) - 1u));
      }
      code2[k] = c2;
    }
    if constexpr (PAIR) {
#pragma unroll
      for (int k = 1; k < R; ++k) {
        if (code2[k] != code2[0]) {
          // not in point 0's cell: this launch evaluates point 0 in its place, the follow-up launch takes point k
          if (flags & (0x100u << k)) {
            const int64_t jk = base + R * (int64_t)tid + k;
            a.left_list[atomicAdd(a.left_count, 1u)] = a.perm ? a.perm[jk] : (uint32_t)jk;
          }
          flags = (flags & ~((0x100u | 1u) << k)) | ((flags & 1u) << k);
          for (int d = 0; d < D; ++d) xs[(d * R + k) * NT + tid] = xs[(d * R + 0) * NT + tid];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const double uA = xs[(K * R + k) * NT + tid], uB = xs[((K + 1) * R + k) * NT + tid], uC = xs[((K + 2) * R + k) * NT + tid];
      const uint32_t qB = blk_key(uB), qC = blk_key(uC);
      pt.xA[k] = uA;
      pt.qA[k] = blk_key(uA);
#pragma unroll
      for (int l = 2; l <= kBlkTab; ++l) {
        uint32_t cellC, cellB;
        hat_cell(uC, qC, l > 3 ? l - 2 : 1, l > 2 ? 1u : 0u, pow2_lm1(l), pt.phC[k][l], cellC);
        hat_cell(uB, qB, l > 3 ? l - 2 : 1, l > 2 ? 1u : 0u, pow2_lm1(l), pt.phB[k][l], cellB);
        if (!PAIR || k == 0) {
          pt.wC8[PAIR ? 0 : k][l] = (blk_off1(l) + cellC) * 8u;
          pt.cB[PAIR ? 0 : k][l] = cellB;
          pin_u32(pt.wC8[PAIR ? 0 : k][l]);
          pin_u32(pt.cB[PAIR ? 0 : k][l]);
        }
        pin_f64(pt.phC[k][l]);
        pin_f64(pt.phB[k][l]);
      }
    }
    double acc[R];
#pragma unroll
    for (int k = 0; k < R; ++k) acc[k] = 0.0;

    for (uint32_t t = 0; t < ntiles; ++t) {
      const int buf = (int)(t & 1u);
      mbar_wait(&sm.bar[buf], (phases >> buf) & 1u);
      phases ^= 1u << buf;
      // the tile's header record came in with the tile: first coefficient slot and number of records (the tile table in
      // global memory cost an L2 round trip in front of each of the 182 tiles of the benchmark grid: 3.7 % of the
      // stall samples of the round-2 profile)
      const uint32_t hbase = smem_u32(sm.rec[buf]);
      uint32_t tl_c0, tl_nrec;
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(tl_c0), "=r"(tl_nrec) : "r"(hbase));
      uint32_t rbase = hbase + (uint32_t)sizeof(BlockRec);
      uint32_t tbias = smem_u32(coefbuf + (size_t)buf * g.blk_tile_coefs) - tl_c0 * 8u;  // shared byte address of coefficient slot 0
      pin_u32(rbase);
      pin_u32(tbias);
      pin_u32(xs_t);  // keep the three column addresses in registers: ptxas would re-derive them per record
      pin_u32(ps_t);
      pin_u32(os_t);
      const uint32_t rend = rbase + tl_nrec * (uint32_t)sizeof(BlockRec);
      double xpre[R];
#ifdef ASG_BLK_PREFETCH_X
      {
        uint32_t k0f;
        asm volatile("ld.shared.u32 %0, [%1+12];" : "=r"(k0f) : "r"(rbase));  // k0 of the tile's first record
        const uint32_t xf = xs_t + min(k0f, (uint32_t)(D - 1)) * (R * NT * 8u);
#pragma unroll
        for (int k = 0; k < R; ++k) xpre[k] = lds_f64(xf + k * (NT * 8u));
      }
#endif
      for (; rbase < rend; rbase += (uint32_t)sizeof(BlockRec)) {
        BlkStaged<R> st;
        blk_stage<R, MASKED, NT>(rbase, ps_t, os_t, xs_t, tbias, a.rem, st, xpre, (uint32_t)(D - 1));
        if constexpr (PAIR) {
          // warp-uniform: every level of the record's prefix is <= kBlkTab, the pair shares the prefix position
          if (st.sw & kBrShared) blk_dispatch<R, MASKED, true, true, false>(pt, st, a.rem, acc);
          else blk_dispatch<R, MASKED, true, false, false>(pt, st, a.rem, acc);
        } else {
          blk_dispatch<R, MASKED, false, false, WIDE>(pt, st, a.rem, acc);
        }
      }
      // No CTA-wide barrier per tile (182 per batch on the benchmark grid, 2.4 % of the stall samples): every warp
      // counts itself out of the buffer and goes on to the next tile; the LAST one out starts the refill. The warps
      // drift at most two tiles apart (a refill needs all of them).
      __syncwarp();
      if ((tid & 31u) == 0u) {
        __threadfence_block();  // the warp's reads of the buffer are done before it is counted out
        if (atomicInc(&sm.done[buf], (unsigned)(NT / 32 - 1)) == (unsigned)(NT / 32 - 1) && t + kBlkTileBufs < ntiles) {
          __threadfence_block();
          uint32_t nr0, nrb, nc0, ncc;  // the header also names the tile after next
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(nr0), "=r"(nrb), "=r"(nc0), "=r"(ncc) : "r"(hbase));
          blk_issue_tile(g, sm, nr0, nrb, nc0, ncc, buf);
        }
      }
    }
    __syncthreads();  // both buffers are free again before the next batch's first tiles are issued
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if ((flags >> (8 + k)) & 1u) {
        const int64_t j = PAIR ? base + R * (int64_t)tid + k : base + (int64_t)k * NT + tid;
        const int64_t mm = a.perm ? (int64_t)a.perm[j] : j;
        const double v = ((flags >> k) & 1u) ? acc[k] : 0.0;
        a.out[mm] = a.fvals ? (a.fvals[mm] - v) : v;
      }
    }
  }
}

// One launch function per (points per thread, masked, paired) variant; each is instantiated in its own translation
// unit (asg_kernels_block_r{1,2,2p}{,m}.cu) so that the large kernels compile in parallel.
template <int R, bool MASKED, int NT, bool PAIR, bool WIDE>
cudaError_t blk_launch(const DeviceGrid &g, const EvalArgs &a, int blocks, size_t smem, cudaStream_t st) {
  // one instantiation serves every grid: raise its dynamic shared-memory limit once per device to the device maximum
  static std::atomic<unsigned> limit_set{0};
  const cudaError_t e = smem_limit_once(k_eval_block<R, MASKED, NT, PAIR, WIDE>, 227 * 1024, limit_set);
  if (e != cudaSuccess) return e;
  k_eval_block<R, MASKED, NT, PAIR, WIDE><<<blocks, NT, smem, st>>>(g, a);
  return cudaSuccess;
}

}  // namespace asg
