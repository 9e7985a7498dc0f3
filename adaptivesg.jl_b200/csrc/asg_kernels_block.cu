// asg_kernels_block.cu -- BLOCK walk of the subspace engine (sm_100a): the main evaluation kernel for
// grids that are (nearly) regular in their last three dimensions (asg_block.h).
//
//   G(x) = sum_n c_n prod_d phi(x01_d; l_nd, i_nd)                       (src/class.jl:629-643)
//
// * One thread owns R = 2 query points, a CTA of 512 threads owns 1024 points and walks the whole record
//   stream for them; one CTA per SM, grid-stride over the points.
// * Records (16 B, one per prefix level vector) and the coefficient blocks they reference are streamed
//   from L2 into shared memory by 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) completed on
//   mbarriers, double buffered.
// * A record applies ONE data-driven stage (the level of its last non-trivial prefix dimension) to the
//   prefix product / position saved by its parent record (a small stack in shared memory), then runs
//   the COMPILED walk of the last three dimensions: the kernel is templated on the suffix budget, all
//   levels and slot offsets are compile-time constants, so a term costs an address add, one LDS.64 and
//   one DFMA, a pole (fixed levels of all dimensions but the last) one IMAD and one DFMA more.
// * The hats of levels 2..5 of the last two dimensions live in per-point register tables; levels 6..8
//   (66 of the 19 448 terms per point on the D = 10 benchmark grid) are formed on the fly.
// The 1-D hat is the reference's phi (src/math.jl:345-357) evaluated on its support only, with the
// reference's single rounding (asg_device.cuh). Out-of-domain / NaN points give exactly 0.0.
#include <cstdlib>

#include "asg_device.cuh"

namespace asg {

constexpr int kBlkThreads = 512;
constexpr int kBlkTab = 4;  // levels 2..kBlkTab of dimensions B, C are tabulated in registers

struct __align__(16) BlkSmem {
  double coef[2][kBlkTileCoefs];
  BlockRec rec[2][kBlkTileRecs];
  unsigned long long bar[2];
  // followed by  double xs[D][R][threads]  (unit-cube coordinates of every dimension),
  //              double Ps[slots][R][threads], uint32_t Os[slots][R][threads]  (saved prefixes)
};

template <int R>
struct BlkPoint {
  double phC[R][kBlkTab + 1];    // last dimension: hat of level l
  uint32_t wC8[R][kBlkTab + 1];  // last dimension: (off1(l) + cell_l) * 8
  double phB[R][kBlkTab + 1];    // dimension D-2: hat of level l
  uint32_t cB[R][kBlkTab + 1];   // dimension D-2: cell of level l
  double xA[R];                  // unit-cube coordinate of dimension D-3 (its hats are formed on the fly)
  uint32_t qA[R];                // and its 30-bit cell key
  uint32_t xsBC;                 // shared byte address of this thread's coordinate of dimension D-2 (point 0)
};

// coordinate of dimension D-2 (which = 0) or D-1 (which = 1) of point k, from shared memory (rare levels only)
template <int R>
__device__ __forceinline__ double blk_x_rare(const BlkPoint<R> &pt, int which, int k) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(pt.xsBC + (uint32_t)((which * R + k) * kBlkThreads * 8)));
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

template <int L, int R>
__device__ __forceinline__ void get_C(const BlkPoint<R> &pt, int k, double &ph, uint32_t &w8) {
  if constexpr (L <= kBlkTab) {
    ph = pt.phC[k][L];
    w8 = pt.wC8[k][L];
  } else {
    uint32_t cell;
    const double x = blk_x_rare<R>(pt, 1, k);
    hat_ct<L>(x, blk_key(x), ph, cell);
    w8 = (blk_off1(L) + cell) * 8u;
  }
}
template <int L, int R>
__device__ __forceinline__ void get_B(const BlkPoint<R> &pt, int k, double &ph, uint32_t &cell) {
  if constexpr (L <= kBlkTab) {
    ph = pt.phB[k][L];
    cell = pt.cB[k][L];
  } else {
    const double x = blk_x_rare<R>(pt, 0, k);
    hat_ct<L>(x, blk_key(x), ph, cell);
  }
}

// ---- compiled suffix walk ---------------------------------------------------------------------------
// pole: levels 1..RP+1 of the last dimension at shared byte address tp[k] + IMM
template <int RP, int LC, int IMM, int R, bool MASKED>
__device__ __forceinline__ void blk_pole(const BlkPoint<R> &pt, const uint32_t (&tp)[R], int reff, double (&inner)[R]) {
  if constexpr (LC <= RP + 1) {
    if (MASKED && LC - 1 > reff) return;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      double ph;
      uint32_t w8;
      get_C<LC, R>(pt, k, ph, w8);
      const double c = lds_f64_imm<IMM>(tp[k] + w8);
      inner[k] = fma(c, ph, inner[k]);
    }
    blk_pole<RP, LC + 1, IMM, R, MASKED>(pt, tp, reff, inner);
  }
}

// 2-D block of storage budget RB at shared byte address t2[k] + IMM: levels LB.. of dimension B
template <int RB, int LB, int IMM, int R, bool MASKED>
__device__ __forceinline__ void blk_suffix2(const BlkPoint<R> &pt, const uint32_t (&t2)[R], int reff, double (&accA)[R]) {
  if constexpr (LB <= RB + 1) {
    if (MASKED && LB - 1 > reff) return;
    constexpr int RP = RB - (LB - 1);
    constexpr int IMM2 = IMM + (int)blk_offB(LB, RB) * 8;
    uint32_t tp[R];
    double phB[R], inner[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if constexpr (LB == 1) {
        tp[k] = t2[k];
      } else {
        uint32_t cell;
        get_B<(LB > 1 ? LB : 2), R>(pt, k, phB[k], cell);
        tp[k] = t2[k] + cell * (blk_n1(RP) * 8u);
      }
      inner[k] = lds_f64_imm<IMM2>(tp[k]);  // level 1 of the last dimension: phi = 1, cell = 0
    }
    blk_pole<RP, 2, IMM2, R, MASKED>(pt, tp, reff - (LB - 1), inner);
#pragma unroll
    for (int k = 0; k < R; ++k) accA[k] = LB == 1 ? inner[k] : fma(phB[k], inner[k], accA[k]);
    blk_suffix2<RB, LB + 1, IMM, R, MASKED>(pt, t2, reff, accA);
  }
}

// chunk of storage budget RS at shared byte address t3[k]: levels LA.. of dimension A
template <int RS, int LA, int R, bool MASKED>
__device__ __forceinline__ void blk_suffix3(const BlkPoint<R> &pt, const uint32_t (&t3)[R], int reff, double (&inner3)[R]) {
  if constexpr (LA <= RS + 1) {
    if (MASKED && LA - 1 > reff) return;
    constexpr int RB = RS - (LA - 1);
    constexpr int IMM = (int)blk_offA(LA, RS) * 8;
    uint32_t t2[R];
    double phA[R], accA[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if constexpr (LA == 1) {
        t2[k] = t3[k];
      } else {
        uint32_t cell;
        hat_ct<(LA > 1 ? LA : 2)>(pt.xA[k], pt.qA[k], phA[k], cell);
        t2[k] = t3[k] + cell * (blk_size2(RB) * 8u);
      }
    }
    blk_suffix2<RB, 1, IMM, R, MASKED>(pt, t2, reff - (LA - 1), accA);
#pragma unroll
    for (int k = 0; k < R; ++k) inner3[k] = LA == 1 ? accA[k] : fma(phA[k], accA[k], inner3[k]);
    blk_suffix3<RS, LA + 1, R, MASKED>(pt, t3, reff, inner3);
  }
}

template <int RS, bool S2, int R, bool MASKED>
__device__ __forceinline__ void blk_record(const BlkPoint<R> &pt, const double (&cur)[R], const uint32_t (&cO)[R],
                                           uint32_t cb, int reff, double (&acc)[R]) {
  uint32_t t[R];
  double v[R];
#pragma unroll
  for (int k = 0; k < R; ++k) t[k] = cO[k] * ((S2 ? blk_size2(RS) : blk_size3(RS)) * 8u) + cb;
  if constexpr (S2) blk_suffix2<RS, 1, 0, R, MASKED>(pt, t, reff, v);
  else blk_suffix3<RS, 1, R, MASKED>(pt, t, reff, v);
#pragma unroll
  for (int k = 0; k < R; ++k) acc[k] = fma(cur[k], v[k], acc[k]);
}

// producer side (one thread): start the copies of tile t into buffer `buf`
__device__ __forceinline__ void blk_issue_tile(const DeviceGrid &g, BlkSmem &sm, uint32_t t, int buf) {
  const BlockTile tl = g.btiles[t];
  const uint32_t rbytes = (tl.r1 - tl.r0) * (uint32_t)sizeof(BlockRec);
  const uint32_t cbytes = (tl.c1 - tl.c0) * 8u;
  // the buffer was last read through the generic proxy: order those reads before the async-proxy writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(&sm.bar[buf], rbytes + cbytes);
  tma_load_1d(sm.rec[buf], g.brec + tl.r0, rbytes, &sm.bar[buf]);
  if (cbytes) tma_load_1d(sm.coef[buf], g.bcoef + tl.c0, cbytes, &sm.bar[buf]);
}

template <int R, bool MASKED>
__global__ void __launch_bounds__(kBlkThreads, 1) k_eval_block(const DeviceGrid g, const EvalArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BlkSmem &sm = *reinterpret_cast<BlkSmem *>(smem_raw);
  constexpr int NT = kBlkThreads;
  const int D = g.D;
  const int K = D - 3;
  double *xs = reinterpret_cast<double *>(smem_raw + sizeof(BlkSmem));
  double *Ps = xs + (size_t)D * R * NT;
  uint32_t *Os = reinterpret_cast<uint32_t *>(Ps + (size_t)g.blk_slots * R * NT);
  const uint32_t tid = threadIdx.x;
  // shared byte addresses of this thread's column in xs / Ps / Os
  const uint32_t xs_t = smem_u32(xs) + tid * 8u;
  const uint32_t ps_t = smem_u32(Ps) + tid * 8u;
  const uint32_t os_t = smem_u32(Os) + tid * 4u;
  const int64_t tile_pts = (int64_t)NT * R;
  const int64_t stride = (int64_t)gridDim.x * tile_pts;
  const uint32_t ntiles = g.n_btiles;
  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {  // slot 0: the empty prefix (never overwritten)
    Ps[k * NT + tid] = 1.0;
    Os[k * NT + tid] = 0u;
  }
  __syncthreads();
  uint32_t phases = 0u;  // bit b: parity the next wait on buffer b expects

  for (int64_t base = (int64_t)blockIdx.x * tile_pts; base < a.M; base += stride) {
    BlkPoint<R> pt;
    pt.xsBC = xs_t + (uint32_t)((K + 1) * R * NT * 8);
    uint32_t inside = 0u;  // bit k: point k lies in the domain
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t j = base + (int64_t)k * NT + tid;
      const int64_t jj = j < a.M ? j : a.M - 1;  // tail lanes redo the last point, never store
      const int64_t mm = a.perm ? (int64_t)a.perm[jj] : jj;
      bool in = true;
      for (int d = 0; d < D; ++d) {
        const double xd = a.X[mm * a.sp + d * a.sd];
        const double u = __dadd_rn(0.0, __ddiv_rn(__dsub_rn(xd, g.lb[d]), g.width[d]));  // scale(), src/math.jl:308
        in = in && (u >= 0.0) && (u <= 1.0);                                             // NaN fails both
        xs[(d * R + k) * NT + tid] = u;
      }
      if (in) inside |= 1u << k;
      else  // walked at the centre of the cube; the result is replaced by the exact 0.0
        for (int d = 0; d < D; ++d) xs[(d * R + k) * NT + tid] = 0.5;
      const double uA = xs[(K * R + k) * NT + tid], uB = xs[((K + 1) * R + k) * NT + tid], uC = xs[((K + 2) * R + k) * NT + tid];
      const uint32_t qB = blk_key(uB), qC = blk_key(uC);
      pt.xA[k] = uA;
      pt.qA[k] = blk_key(uA);
#pragma unroll
      for (int l = 2; l <= kBlkTab; ++l) {
        uint32_t cell;
        hat_cell(uC, qC, l > 3 ? l - 2 : 1, l > 2 ? 1u : 0u, pow2_lm1(l), pt.phC[k][l], cell);
        pt.wC8[k][l] = (blk_off1(l) + cell) * 8u;
        hat_cell(uB, qB, l > 3 ? l - 2 : 1, l > 2 ? 1u : 0u, pow2_lm1(l), pt.phB[k][l], pt.cB[k][l]);
        pin_f64(pt.phC[k][l]); pin_u32(pt.wC8[k][l]);
        pin_f64(pt.phB[k][l]); pin_u32(pt.cB[k][l]);
      }
    }
    if (tid == 0) {
      blk_issue_tile(g, sm, 0, 0);
      if (ntiles > 1) blk_issue_tile(g, sm, 1, 1);
    }
    double acc[R];
#pragma unroll
    for (int k = 0; k < R; ++k) acc[k] = 0.0;

    for (uint32_t t = 0; t < ntiles; ++t) {
      const int buf = (int)(t & 1u);
      const BlockTile tl = g.btiles[t];
      mbar_wait(&sm.bar[buf], (phases >> buf) & 1u);
      phases ^= 1u << buf;
      uint32_t rbase = smem_u32(sm.rec[buf]);
      uint32_t tbias = smem_u32(sm.coef[buf]) - tl.c0 * 8u;  // shared byte address of coefficient slot 0
      pin_u32(rbase);
      pin_u32(tbias);
      const uint32_t nrec = tl.r1 - tl.r0;
      for (uint32_t j = 0; j < nrec; ++j) {
        uint4 rc;  // {cbase, meta, s_hi, used}
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rc.x), "=r"(rc.y), "=r"(rc.z), "=r"(rc.w) : "r"(rbase));
        rbase += (uint32_t)sizeof(BlockRec);
        if (MASKED && (int)rc.w > a.rem) continue;  // depth mask of _interpolate_until_level: the whole record lies deeper
        const uint32_t meta = rc.y;
        const uint32_t slot = (meta >> kBrSlotShift) & 15u;
        const uint32_t k0 = (meta >> kBrDimShift) & 15u;
        const uint32_t b = (meta >> kBrBitsShift) & 63u;
        const uint32_t odd = (meta >> 24) & 1u;
        const double s = __hiloint2double((int)rc.z, 0);                    // 2^(l-1), or 0 for the root record
        const double s2b = __hiloint2double((int)((1023u + b) << 20), 0);  // 2^b
        const uint32_t lim = (1u << b) - 1u;
        // the stage: hat of the level-l node of dimension k0 that supports the point, on top of the saved prefix
        double cur[R];
        uint32_t cO[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const double Pin = lds_f64(ps_t + (slot * R + k) * (NT * 8u));
          uint32_t Oin;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Oin) : "r"(os_t + (slot * R + k) * (NT * 4u)));
          const double xv = lds_f64(xs_t + (k0 * R + k) * (NT * 8u));
          const uint32_t cell = min((uint32_t)(xv * s2b), lim);  // floor(x*2^b), clamped for x == 1 (exact)
          const double tt = fma(xv, s, -u32_to_f64((cell << 1) + odd));
          cur[k] = Pin * (1.0 - fabs(tt));
          cO[k] = (Oin << b) | cell;
        }
        if (meta & kBrSave) {
#pragma unroll
          for (int k = 0; k < R; ++k) {
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(ps_t + ((slot + 1u) * R + k) * (NT * 8u)), "d"(cur[k]));
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(os_t + ((slot + 1u) * R + k) * (NT * 4u)), "r"(cO[k]));
          }
        }
        const uint32_t cb = rc.x * 8u + tbias;
        const int reff = MASKED ? a.rem - (int)rc.w : 0;
        switch (meta & (15u | kBrS2)) {
          case 0: blk_record<0, false, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case 1: blk_record<1, false, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case 2: blk_record<2, false, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case 3: blk_record<3, false, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case 4: blk_record<4, false, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case 5: blk_record<5, false, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case 6: blk_record<6, false, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case 7: blk_record<7, false, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case kBrS2 | 0: blk_record<0, true, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case kBrS2 | 1: blk_record<1, true, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case kBrS2 | 2: blk_record<2, true, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case kBrS2 | 3: blk_record<3, true, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case kBrS2 | 4: blk_record<4, true, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case kBrS2 | 5: blk_record<5, true, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case kBrS2 | 6: blk_record<6, true, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          case kBrS2 | 7: blk_record<7, true, R, MASKED>(pt, cur, cO, cb, reff, acc); break;
          default: break;  // kBrNone: a stage-only record
        }
      }
      __syncthreads();  // every warp is done with buffer `buf`
      if (tid == 0 && t + 2 < ntiles) blk_issue_tile(g, sm, t + 2, buf);
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int64_t j = base + (int64_t)k * NT + tid;
      if (j < a.M) {
        const int64_t mm = a.perm ? (int64_t)a.perm[j] : j;
        const double v = ((inside >> k) & 1u) ? acc[k] : 0.0;
        a.out[mm] = a.fvals ? (a.fvals[mm] - v) : v;
      }
    }
  }
}

static int blk_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static size_t blk_smem_bytes(const DeviceGrid &g, int R) {
  return sizeof(BlkSmem) + ((size_t)g.D * 8 + (size_t)g.blk_slots * 12) * R * kBlkThreads;
}

bool block_kernel_fits(const DeviceGrid &g) { return g.n_btiles > 0 && blk_smem_bytes(g, 1) <= 227 * 1024; }

cudaError_t launch_eval_block(const DeviceGrid &g, const EvalArgs &a, cudaStream_t st, int64_t *launches) {
  if (a.M <= 0) return cudaSuccess;
  const int sms = blk_sm_count();
  // two points per thread when the batch fills the machine and the shared memory fits, else one
  bool two = a.M >= (int64_t)sms * kBlkThreads * 2 && blk_smem_bytes(g, 2) <= 227 * 1024;
  if (const char *e = getenv("ASG_BLOCK_R")) two = (e[0] == '2') && blk_smem_bytes(g, 2) <= 227 * 1024;  // experiment knob
  const int R = two ? 2 : 1;
  const size_t smem = blk_smem_bytes(g, R);
  const int64_t per_block = (int64_t)kBlkThreads * R;
  const int64_t need = (a.M + per_block - 1) / per_block;
  const int blocks = (int)(need < sms ? need : sms);
  const bool masked = a.rem < INT_MAX / 4;
#define LAUNCH(RR, MM)                                                                                        \
  do {                                                                                                        \
    cudaFuncSetAttribute(k_eval_block<RR, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    k_eval_block<RR, MM><<<blocks, kBlkThreads, smem, st>>>(g, a);                                            \
  } while (0)
  if (two) { if (masked) LAUNCH(2, true); else LAUNCH(2, false); }
  else     { if (masked) LAUNCH(1, true); else LAUNCH(1, false); }
#undef LAUNCH
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace asg
