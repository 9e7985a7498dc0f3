// asg_api.cu -- the C ABI of libasg_cuda.so (include/asg_cuda.h): context, node store mirror,
// host-buffer pipelines around the kernels. No torch types, no CPU compute fallback.
#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges are no-ops unless a profiler is attached

#include "asg_internal.h"

using namespace asg;

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
namespace {

thread_local std::string t_err;
int set_err(int code, const std::string &msg) {
  t_err = msg;
  return code;
}

constexpr int kMaxDev = 16;   // devices one process may drive (asg_init_devices)
constexpr int kPipeSlots = 3;  // host-buffer pipeline depth per device

struct DevCtx {  // per device: streams and events of this library
  int device = -1;
  cudaStream_t stream = nullptr;                 // library stream (device-pointer entry points, uploads)
  cudaStream_t pipe[kPipeSlots] = {};            // host-buffer pipeline: H2D k+1 / walk k / D2H k-1 overlap
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;      // around the last *_device call
  cudaEvent_t evw0 = nullptr, evw1 = nullptr;    // around its walk kernel alone (BLOCK walk: after the point ordering)
  bool walk_timed = false;
  bool last_timed = false;
  // pinned staging of large device -> PAGEABLE host copies (d2h_fast)
  static constexpr int kStage = 4;
  void *stage[kStage] = {};
  cudaEvent_t stage_ev[kStage] = {};
  std::mutex stage_mu;
  // pinned staging of large PAGEABLE host -> device copies (the feeder of the evaluation pipeline)
  void *stage_in[kStage] = {};
  cudaEvent_t stage_in_ev[kStage] = {};
  std::mutex stage_in_mu;
};
constexpr size_t kStagePiece = (size_t)16 << 20;  // bytes per staging buffer
// ASG_STAGE_HOST=0: leave pageable host buffers to the driver's own staging (A/B knob)
bool stage_host() {
  static const bool v = [] { const char *e = getenv("ASG_STAGE_HOST"); return !(e && e[0] == '0'); }();
  return v;
}
bool is_pageable(const void *p) {
  cudaPointerAttributes at;
  const bool pinned_or_dev = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered;
  cudaGetLastError();
  return !pinned_or_dev;
}

struct Context {
  std::mutex mu;
  bool ready = false;
  int ndev = 0;
  DevCtx dev[kMaxDev];
  int last_dev = 0;  // replica of the last timed *_device call
  std::atomic<long long> launches{0};
} ctx;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512];                                                                                \
      snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return set_err(e_ == cudaErrorMemoryAllocation ? ASG_ENOMEM : ASG_ECUDA, b_);               \
    }                                                                                              \
  } while (0)

int use_dev(int r) {  // make replica r's device current for the calling thread
  cudaError_t e = cudaSetDevice(ctx.dev[r].device);
  if (e != cudaSuccess) return set_err(ASG_ECUDA, cudaGetErrorString(e));
  return ASG_OK;
}

int require_ctx() {
  if (!ctx.ready) return set_err(ASG_ENODEVICE, "asg_init has not succeeded: no sm_100 device selected (no CPU fallback)");
  return use_dev(0);
}

template <class T>
struct DevBuf {  // grow-only device buffer
  T *p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return ASG_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return set_err(ASG_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cap = n;
    return ASG_OK;
  }
  // grow to at least n elements KEEPING the first `keep` (device-to-device copy on `st`, then synchronised)
  int grow_keep(size_t n, size_t keep, cudaStream_t st) {
    if (n <= cap) return ASG_OK;
    T *q = nullptr;
    const size_t ncap = std::max<size_t>(n, cap + cap / 2 + 64);
    cudaError_t e = cudaMalloc(&q, ncap * sizeof(T));
    if (e != cudaSuccess) return set_err(ASG_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    if (p && keep) e = cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(q); return set_err(ASG_ECUDA, std::string("device copy: ") + cudaGetErrorString(e)); }
    if (p) cudaFree(p);
    p = q;
    cap = ncap;
    return ASG_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

void bump(int64_t n) { ctx.launches += n; }

// NVTX range around a C-ABI call (tracing), visible in nsys / ncu --nvtx
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define ASG_TRACE() NvtxRange nvtx_range_(__func__)

}  // namespace

// device storage of one grid on one device (replica 0 is the primary: basis, children and node coordinates run there)
struct Replica {
  DeviceGrid dg{};
  DevBuf<TrieRec> d_rec;
  DevBuf<GroupRec> d_grp;
  DevBuf<StreamTile> d_tiles;
  DevBuf<double> d_bcoef;
  DevBuf<BlockRec> d_brec;
  DevBuf<BlockTile> d_btiles;
  DevBuf<double> d_coef, d_sw_coef;
  DevBuf<int32_t> d_orig, d_horig, d_sw_depth;
  DevBuf<HashEnt> d_hent;
  DevBuf<SweepDim> d_sw;
  DevBuf<uint64_t> d_sw_low;
  // scratch for host-buffer calls (one set per pipeline slot)
  DevBuf<double> s_X[kPipeSlots], s_out[kPipeSlots], s_f[kPipeSlots];
  DevBuf<long long> s_i64a, s_i64b;
  DevBuf<double> s_tmp, s_tmp2;
  DevBuf<int> s_flag;
  // adapt! candidate generation (asg_kernels_adapt.cu): the unique children of the last asg_children call
  DevBuf<long long> c_PL, c_PI, c_CL, c_CI, c_OL, c_OI;
  DevBuf<unsigned char> c_valid;
  DevBuf<int> c_table;
  DevBuf<unsigned int> c_flags, c_sums;
  // point-order scratch of the BLOCK walk (asg_sort.cu), one set per pipeline slot
  DevBuf<uint32_t> s_key[kPipeSlots], s_key2[kPipeSlots], s_idx[kPipeSlots], s_perm[kPipeSlots];
  DevBuf<unsigned char> s_sorttmp[kPipeSlots];
  DevBuf<uint32_t> s_left[kPipeSlots], s_leftcnt[kPipeSlots];  // PAIR mode of the BLOCK walk: points left to the follow-up launch
  cudaEvent_t sort_ev[kPipeSlots] = {};  // last use of the slot's scratch
  std::vector<cudaEvent_t> piece_ev;     // host-buffer pipeline: arrival of each piece of points
  cudaEvent_t walk_ev[2] = {};           // and the end of the last two windows' walks
  void release_all() {
    d_rec.release(); d_grp.release(); d_tiles.release(); d_bcoef.release(); d_brec.release(); d_btiles.release();
    d_coef.release(); d_sw_coef.release(); d_orig.release(); d_horig.release(); d_sw_depth.release(); d_hent.release();
    d_sw.release(); d_sw_low.release();
    for (int s = 0; s < kPipeSlots; ++s) {
      s_X[s].release(); s_out[s].release(); s_f[s].release();
      s_key[s].release(); s_key2[s].release(); s_idx[s].release(); s_perm[s].release(); s_sorttmp[s].release();
      s_left[s].release(); s_leftcnt[s].release();
      if (sort_ev[s]) cudaEventDestroy(sort_ev[s]);
      sort_ev[s] = nullptr;
    }
    for (cudaEvent_t e : piece_ev) cudaEventDestroy(e);
    piece_ev.clear();
    for (int k = 0; k < 2; ++k) { if (walk_ev[k]) cudaEventDestroy(walk_ev[k]); walk_ev[k] = nullptr; }
    c_PL.release(); c_PI.release(); c_CL.release(); c_CI.release(); c_OL.release(); c_OI.release();
    c_valid.release(); c_table.release(); c_flags.release(); c_sums.release();
    s_i64a.release(); s_i64b.release(); s_tmp.release(); s_tmp2.release(); s_flag.release();
  }
};

struct asg_grid {
  std::recursive_mutex mu;
  int D = 0;
  double lb[kMaxDim], ub[kMaxDim];
  // host copy of the caller's arrays (row-major N x D): the caller's struct stays the source of truth,
  // this copy exists so append / set_coefs / surplus_nodes need no second pass over caller memory
  int64_t N = 0;
  // nodes 0 .. N_packed-1 are in the packed layouts; nodes N_packed .. N-1 (appended since, asg_grid_append) only in the
  // sweep arrays: every evaluation adds their sum on top (run_eval) until the next full pack folds them in
  int64_t N_packed = 0;
  int64_t packs = 0;
  std::vector<int64_t> levels, indices;
  std::vector<double> coefs;
  std::vector<int64_t> slot_of_node, bslot_of_node;
  bool blk_ok = false;
  std::string blk_why;
  bool uploaded = false, canonical = false, finite = true, regular = false;
  RsgParams rp{};
  std::string why;
  int forced_path = ASG_PATH_AUTO;
  asg_grid_info_t info{};
  Replica rep[kMaxDev];  // one per device of the context; all hold the same packed layouts
  int64_t ch_n = -1, ch_generation = -1;
  std::vector<unsigned char> blob;  // asg_grid_export: between the size query and the copy
  int64_t blob_generation = -1;
  int path() const {
    if (forced_path == ASG_PATH_SWEEP) return ASG_PATH_SWEEP;
    return (canonical && finite) ? ASG_PATH_SUBSPACE : ASG_PATH_SWEEP;
  }
  // which walk of the subspace engine: 3 = BLOCK walk (default where the layout exists), 0 = flat group
  // stream (other canonical grids, ASG_EVAL_KERNEL=stream), 1 = record-free regular walk (=rsg),
  // 2 = recursive trie walk (ASG_PATH_TRIE, D == 1, or ASG_EVAL_KERNEL=trie)
  int walk() const {
    static const char *env = getenv("ASG_EVAL_KERNEL");
    const DeviceGrid &dg = rep[0].dg;
    if (forced_path == ASG_PATH_TRIE || D < 2 || dg.n_tiles == 0) return 2;
    if (forced_path == ASG_PATH_STREAM) return 0;
    if (forced_path == ASG_PATH_BLOCK) return (blk_ok && block_kernel_fits(dg)) ? 3 : 0;
    if (env && env[0] == 't') return 2;
    if (env && env[0] == 'r' && regular) return 1;
    if (env && env[0] == 's') return 0;
    if (blk_ok && block_kernel_fits(dg)) return 3;
    return 0;
  }
};

// bring replica r of g into scope of a function body
#define USE_REPLICA(g, r)        \
  Replica &R = (g)->rep[(r)];    \
  DevCtx &dc = ctx.dev[(r)];     \
  (void)R;                       \
  (void)dc

struct asg_event {
  std::thread th;
  int rc = ASG_OK;
  std::string err;
};

namespace {

int device_upload(asg_grid *g, int r, const PackResult &pk) {
  if (int rc = use_dev(r)) return rc;
  USE_REPLICA(g, r);
  const int D = g->D;
  const int64_t N = g->N;
  DeviceGrid &dg = R.dg;
  dg = DeviceGrid{};
  dg.D = D;
  std::vector<BlockRec> blk_drec;    // live until the stream is drained at the end of the function
  std::vector<BlockTile> blk_dtiles;
  for (int d = 0; d < D; ++d) {
    dg.lb[d] = g->lb[d];
    volatile double w = g->ub[d] - g->lb[d];  // (from_ub .- from_lb), src/math.jl:308
    dg.width[d] = w;
  }
  dg.N = N;
  cudaStream_t st = dc.stream;
  // sweep engine arrays (always present: they serve every node set)
  if (int rc = R.d_sw.reserve((size_t)N * D)) return rc;
  if (int rc = R.d_sw_low.reserve((size_t)N)) return rc;
  if (int rc = R.d_sw_coef.reserve((size_t)N)) return rc;
  if (int rc = R.d_sw_depth.reserve((size_t)N)) return rc;
  if (N) {
    CK(cudaMemcpyAsync(R.d_sw.p, pk.sw.data(), (size_t)N * D * sizeof(SweepDim), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(R.d_sw_low.p, pk.sw_low.data(), (size_t)N * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(R.d_sw_coef.p, g->coefs.data(), (size_t)N * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(R.d_sw_depth.p, pk.depth.data(), (size_t)N * 4, cudaMemcpyHostToDevice, st));
  }
  dg.sw = R.d_sw.p;
  dg.sw_low = R.d_sw_low.p;
  dg.sw_coef = R.d_sw_coef.p;
  dg.sw_depth = R.d_sw_depth.p;
  // subspace engine arrays
  if (pk.canonical) {
    size_t total = 0;
    for (int k = 0; k < D; ++k) total += pk.rec[(size_t)k].size();
    if (int rc = R.d_rec.reserve(total)) return rc;
    size_t off = 0;
    for (int k = 0; k < D; ++k) {
      const auto &r = pk.rec[(size_t)k];
      CK(cudaMemcpyAsync(R.d_rec.p + off, r.data(), r.size() * sizeof(TrieRec), cudaMemcpyHostToDevice, st));
      dg.rec[k] = R.d_rec.p + off;
      off += r.size();
    }
    dg.n_root = (uint32_t)(pk.rec[0].size() - 1);
    dg.n_level1 = D >= 2 ? (uint32_t)(pk.rec[1].size() - 1) : 0u;
    if (int rc = R.d_coef.reserve(pk.coef.size())) return rc;
    if (int rc = R.d_orig.reserve(pk.orig.size())) return rc;
    if (int rc = R.d_hent.reserve(pk.hent.size())) return rc;
    if (int rc = R.d_horig.reserve(pk.horig.size())) return rc;
    if (!pk.coef.empty()) {
      CK(cudaMemcpyAsync(R.d_coef.p, pk.coef.data(), pk.coef.size() * 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(R.d_orig.p, pk.orig.data(), pk.orig.size() * 4, cudaMemcpyHostToDevice, st));
    }
    if (!pk.hent.empty()) {
      CK(cudaMemcpyAsync(R.d_hent.p, pk.hent.data(), pk.hent.size() * sizeof(HashEnt), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(R.d_horig.p, pk.horig.data(), pk.horig.size() * 4, cudaMemcpyHostToDevice, st));
    }
    if (int rc = R.d_grp.reserve(pk.grp.size())) return rc;
    if (int rc = R.d_tiles.reserve(pk.tiles.size())) return rc;
    if (!pk.grp.empty()) {
      CK(cudaMemcpyAsync(R.d_grp.p, pk.grp.data(), pk.grp.size() * sizeof(GroupRec), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(R.d_tiles.p, pk.tiles.data(), pk.tiles.size() * sizeof(StreamTile), cudaMemcpyHostToDevice, st));
    }
    dg.grp = R.d_grp.p;
    dg.tiles = R.d_tiles.p;
    dg.n_grp = (uint32_t)pk.grp.size();
    dg.n_tiles = (uint32_t)pk.tiles.size();
    dg.coef = R.d_coef.p;
    dg.orig = R.d_orig.p;
    dg.hent = R.d_hent.p;
    dg.horig = R.d_horig.p;
    if (pk.blk_ok) {
      if (int rc = R.d_bcoef.reserve(pk.bcoef.size())) return rc;
      // device record stream: a header record in front of every tile (blk_tile_header, asg_block.h)
      block_device_stream(pk, blk_drec, blk_dtiles);
      if (int rc = R.d_brec.reserve(blk_drec.size())) return rc;
      if (int rc = R.d_btiles.reserve(blk_dtiles.size())) return rc;
      CK(cudaMemcpyAsync(R.d_bcoef.p, pk.bcoef.data(), pk.bcoef.size() * 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(R.d_brec.p, blk_drec.data(), blk_drec.size() * sizeof(BlockRec), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(R.d_btiles.p, blk_dtiles.data(), blk_dtiles.size() * sizeof(BlockTile), cudaMemcpyHostToDevice, st));
      dg.bcoef = R.d_bcoef.p;
      dg.brec = R.d_brec.p;
      dg.btiles = R.d_btiles.p;
      dg.n_brec = (uint32_t)blk_drec.size();
      dg.n_btiles = (uint32_t)pk.btiles.size();
      dg.blk_slots = pk.blk_slots;
      dg.blk_tile_coefs = pk.blk_tile_coefs;
      dg.blk_pair_ok = pk.blk_pair_ok ? 1 : 0;
      dg.blk_wide = pk.blk_wide ? 1 : 0;
    }
  }
  CK(cudaStreamSynchronize(st));
  return ASG_OK;
}

int install_pack(asg_grid *g, PackResult &pk);

int repack(asg_grid *g) {
  PackResult pk;
  std::string err;
  int rc = pack_grid(g->D, g->N, g->levels.data(), g->indices.data(), g->coefs.data(), pk, err);
  if (rc != ASG_OK) return set_err(rc, err);
  return install_pack(g, pk);
}

// make a packed grid (fresh from pack_grid, or read from a blob) the state of the handle: device upload, slot maps, info
int install_pack(asg_grid *g, PackResult &pk) {
  int rc;
  g->canonical = pk.canonical;
  g->regular = pk.regular;
  g->rp = pk.rp;
  g->why = pk.why;
  g->finite = true;
  for (double c : g->coefs)
    if (!std::isfinite(c)) { g->finite = false; break; }
  // every device of the context gets the same packed layouts (packed once on the host)
  for (int r = ctx.ndev - 1; r >= 0; --r) {
    PackTimer t("device upload (one replica)");
    rc = device_upload(g, r, pk);
    if (rc != ASG_OK) return rc;
  }
  USE_REPLICA(g, 0);
  g->slot_of_node.swap(pk.slot_of_node);
  g->bslot_of_node.swap(pk.bslot_of_node);
  g->blk_ok = pk.canonical && pk.blk_ok;
  g->blk_why = pk.blk_why;
  g->uploaded = true;
  g->N_packed = g->N;
  g->packs++;
  asg_grid_info_t &I = g->info;
  const int64_t gen = I.generation + 1;
  I = asg_grid_info_t{};
  I.D = g->D;
  I.canonical = pk.canonical ? 1 : 0;
  I.max_level = pk.max_level;
  I.N = g->N;
  I.max_depth = pk.max_depth;
  I.subspaces = pk.subspaces;
  I.dense_subspaces = pk.dense_subspaces;
  I.trie_nodes = pk.trie_nodes;
  I.coef_slots = (int64_t)pk.coef.size() - pk.coef_pad;
  I.hash_slots = (int64_t)pk.hent.size();
  I.device_bytes = (int64_t)(R.d_rec.bytes() + R.d_grp.bytes() + R.d_tiles.bytes() + R.d_coef.bytes() + R.d_orig.bytes() + R.d_hent.bytes() +
                             R.d_horig.bytes() + R.d_sw.bytes() + R.d_sw_low.bytes() + R.d_sw_coef.bytes() +
                             R.d_sw_depth.bytes());
  I.generation = gen;
  I.regular = pk.regular ? 1 : 0;
  I.block = g->blk_ok ? 1 : 0;
  I.block_slots = g->blk_ok ? (int64_t)pk.bcoef.size() - 2 : 0;
  I.block_records = g->blk_ok ? (int64_t)pk.brec.size() : 0;
  I.block_tiles = g->blk_ok ? (int64_t)pk.btiles.size() : 0;
  I.block_stack = g->blk_ok ? pk.blk_slots : 0;
  I.block_pair = (g->blk_ok && pk.blk_pair_ok && !pk.blk_wide) ? 1 : 0;
  I.packs = g->packs;
  I.delta_nodes = 0;
  if (g->blk_ok) I.device_bytes += (int64_t)(R.d_bcoef.bytes() + R.d_brec.bytes() + R.d_btiles.bytes());
  return ASG_OK;
}

int copy_nodes(asg_grid *g, int64_t at, int64_t n, const int64_t *levels, const int64_t *indices, int64_t sn,
               int64_t sd, const double *coefs) {
  const int D = g->D;
  if (sn == D && sd == 1 && n > 0) {
    // row-major, contiguous: appended in one pass each (no zero fill first), the two large arrays side by side
    g->levels.resize((size_t)at * D);
    g->indices.resize((size_t)at * D);
    g->coefs.resize((size_t)at);
    auto app = [&](std::vector<int64_t> &v, const int64_t *p) { v.insert(v.end(), p, p + (size_t)n * D); };
    if (n * D >= (1 << 20)) {
      std::thread t([&]() { app(g->levels, levels); });
      app(g->indices, indices);
      t.join();
    } else {
      app(g->levels, levels);
      app(g->indices, indices);
    }
    if (coefs) g->coefs.insert(g->coefs.end(), coefs, coefs + n);
    else g->coefs.resize((size_t)(at + n), 0.0);
    g->N = at + n;
    return ASG_OK;
  }
  g->levels.resize((size_t)(at + n) * D);
  g->indices.resize((size_t)(at + n) * D);
  g->coefs.resize((size_t)(at + n));
  pack_parallel(n, pack_threads(n), [&](int, int64_t k0, int64_t k1) {
    for (int64_t k = k0; k < k1; ++k) {
      for (int d = 0; d < D; ++d) {
        g->levels[(size_t)(at + k) * D + d] = levels[k * sn + d * sd];
        g->indices[(size_t)(at + k) * D + d] = indices[k * sn + d * sd];
      }
      g->coefs[(size_t)(at + k)] = coefs ? coefs[k] : 0.0;
    }
  });
  g->N = at + n;
  return ASG_OK;
}

int rem_of(int64_t max_depth) {
  if (max_depth < 0) return INT_MAX / 2;
  return (int)std::min<int64_t>(max_depth - 1, INT_MAX / 2);  // depth <= L  <=>  sum(l-1) <= L-1
}

// minimum batch for which the BLOCK walk orders its points by cell key first (ASG_SORT_POINTS=0 disables,
// =<n> sets the threshold). Measured on the D = 10 benchmark grid: ordered 0.61 ms vs 0.88 ms at 5 000 points,
// 1.59 vs 3.28 ms at 300 000 -- the ordering pays from the smallest batches on.
int64_t sort_threshold() {
  static const int64_t thr = [] {
    const char *e = getenv("ASG_SORT_POINTS");
    if (!e) return (int64_t)1 << 10;
    const long long v = atoll(e);
    return v <= 0 ? INT64_MAX : (int64_t)v;
  }();
  return thr;
}

int run_eval(asg_grid *g, int r, const EvalArgs &a_in, cudaStream_t st, int slot = 0, bool timed = false) {
  USE_REPLICA(g, r);
  int64_t l = 0;
  if (timed) dc.walk_timed = false;
  cudaError_t e;
  EvalArgs a = a_in;
  if (g->path() != ASG_PATH_SUBSPACE) e = launch_eval_sweep(R.dg, a, st, &l);
  else if (g->walk() == 3) {
    bool sorted = false;
    // (grids of a few hundred subspaces walk in less time than the ordering's five launches take)
    if (a.M >= sort_threshold() && g->info.subspaces >= 512) {
      if (a.M >= ((int64_t)1 << 31)) {
        // the permutation is 32-bit: callers with device-resident batches this large should split them (the host
        // pipeline always does). Say so once instead of silently walking ~1.6x slower.
        static std::atomic<bool> warned{false};
        if (!warned.exchange(true))
          fprintf(stderr, "libasg_cuda: batch of %lld points >= 2^31: cell-key ordering skipped (walk ~1.6x slower); "
                          "split the batch\n", (long long)a.M);
      } else {
        const size_t tb = sort_points_temp_bytes(a.M);
        if (int rc = R.s_key[slot].reserve((size_t)a.M)) return rc;
        if (int rc = R.s_key2[slot].reserve((size_t)a.M)) return rc;
        if (int rc = R.s_idx[slot].reserve((size_t)a.M)) return rc;
        if (int rc = R.s_perm[slot].reserve((size_t)a.M)) return rc;
        if (int rc = R.s_sorttmp[slot].reserve(tb)) return rc;
        if (!R.sort_ev[slot]) CK(cudaEventCreateWithFlags(&R.sort_ev[slot], cudaEventDisableTiming));
        else CK(cudaStreamWaitEvent(st, R.sort_ev[slot], 0));  // an earlier launch on another stream may still read the scratch
        CK(sort_points(R.dg, a.X, a.M, a.sp, a.sd, R.s_key[slot].p, R.s_key2[slot].p, R.s_idx[slot].p, R.s_perm[slot].p,
                       R.s_sorttmp[slot].p, tb, st, &l));
        a.perm = R.s_perm[slot].p;
        sorted = true;
      }
    }
    uint32_t *left = nullptr, *leftcnt = nullptr;
    if (sorted && block_pair_mode(R.dg, a.M, true)) {
      if (int rc = R.s_left[slot].reserve((size_t)a.M + 2)) return rc;  // (groups of four: up to 3/4 of the points)
      if (int rc = R.s_leftcnt[slot].reserve(4)) return rc;
      left = R.s_left[slot].p;
      leftcnt = R.s_leftcnt[slot].p;
    }
    if (timed) CK(cudaEventRecord(dc.evw0, st));
    e = launch_eval_block(R.dg, a, st, &l, left, leftcnt);
    if (timed) { CK(cudaEventRecord(dc.evw1, st)); dc.walk_timed = true; }
    if (sorted && e == cudaSuccess) CK(cudaEventRecord(R.sort_ev[slot], st));
  }
  else if (g->walk() == 0) e = launch_eval_stream(R.dg, a, st, &l);
  else if (g->walk() == 1) e = launch_eval_rsg(R.dg, g->rp, a, st, &l);
  else e = launch_eval_subspace(R.dg, a, st, &l);
  if (e == cudaSuccess && g->path() == ASG_PATH_SUBSPACE && g->N_packed < g->N) {
    // nodes appended since the last pack: their sum goes on top of the packed layouts' result
    EvalArgs b = a_in;
    b.n_first = g->N_packed;
    b.accumulate = 1;
    e = launch_eval_sweep(R.dg, b, st, &l);
  }
  bump(l);
  CK(e);
  return ASG_OK;
}

// Fold appended nodes into the packed layouts when the extra sweep over them would cost more than it saves:
// M points x delta nodes x D hat evaluations against a host pack of tens to hundreds of milliseconds.
int consolidate(asg_grid *g) {
  if (g->uploaded && g->N_packed < g->N) return repack(g);
  return ASG_OK;
}
int maybe_consolidate(asg_grid *g, int64_t M) {
  const double work = (double)M * (double)(g->N - g->N_packed) * g->D;
  return work > 4e10 ? consolidate(g) : ASG_OK;
}

// Stage rows [m0, m0+cnt) of a strided host matrix into a dense device block and return the device
// strides. Fast paths: row-major rows (sdim == 1) and column-major columns (sp == 1).
int stage_points(const double *X, int64_t m0, int64_t cnt, int D, int64_t sp, int64_t sdim, double *dX,
                 int64_t &dsp, int64_t &dsd, cudaStream_t st, std::vector<double> &repack_buf) {
  if ((sdim == 1 || D == 1) && sp == D) {  // dense row-major block: one flat copy
    dsp = D;
    dsd = 1;
    CK(cudaMemcpyAsync(dX, X + m0 * sp, (size_t)cnt * D * 8, cudaMemcpyHostToDevice, st));
  } else if (sdim == 1 || D == 1) {
    dsp = D;
    dsd = 1;
    CK(cudaMemcpy2DAsync(dX, (size_t)D * 8, X + m0 * sp, (size_t)sp * 8, (size_t)D * 8, (size_t)cnt,
                         cudaMemcpyHostToDevice, st));
  } else if (sp == 1) {
    dsp = 1;
    dsd = cnt;
    for (int d = 0; d < D; ++d)
      CK(cudaMemcpyAsync(dX + (size_t)d * cnt, X + m0 + d * sdim, (size_t)cnt * 8, cudaMemcpyHostToDevice, st));
  } else {
    dsp = D;
    dsd = 1;
    repack_buf.resize((size_t)cnt * D);
    for (int64_t k = 0; k < cnt; ++k)
      for (int d = 0; d < D; ++d) repack_buf[(size_t)k * D + d] = X[(m0 + k) * sp + d * sdim];
    CK(cudaMemcpyAsync(dX, repack_buf.data(), (size_t)cnt * D * 8, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // repack_buf is reused by the next chunk
  }
  return ASG_OK;
}

// ---- host-buffer evaluation pipeline --------------------------------------------------------------------------------
// The points of a call go to the device in PIECES of pipe_piece() points, issued back to back by a feeder thread on
// the copy-in stream (so that a pageable source, whose cudaMemcpyAsync blocks the issuing thread, still overlaps with
// the walks). The walk side is GREEDY: whenever the GPU is free it takes EVERY piece that has arrived since the last
// window as one batch (point ordering + walk), then looks again. The window sizes therefore adapt to the machine by
// themselves: with one GPU a copy runs ~3x faster than the walk of the same points and the windows grow 0.3 M,
// 0.9 M, 2.7 M, ... until everything has arrived and the rest goes in one large, sharply ordered batch (PAIR mode);
// with eight GPUs sharing the host's PCIe fabric (23-36 GB/s each, profiles/r02_pcie_bandwidth_8gpu.json) copy and
// walk run at about the same rate and the windows stay moderate. Only the first piece's copy, and the walk + D2H copy
// of the last window -- a short tail split off on purpose -- are not hidden behind other work. Results return on a
// third stream, one window behind. Batches above pipe_super() points run as consecutive super-batches (device buffers).
// ASG_PIPE_PIECE / ASG_PIPE_SUPER (points) override the two sizes.
int64_t pipe_piece() {
  // two full waves of the walk's CTAs (one CTA of 1 024 points per SM): a window of k pieces is 2k full waves, where
  // 2^18 points were 1.73 waves on 148 SMs -- 14 % of the small windows' time went to the half-empty last wave
  static const int64_t v = [] {
    const char *e = getenv("ASG_PIPE_PIECE");
    const long long x = e ? atoll(e) : 0;
    if (x > 0) return (int64_t)x;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int64_t)2 * (sms > 0 ? sms : 148) * 1024;
  }();
  return v;
}
int64_t pipe_super() {
  static const int64_t v = [] { const char *e = getenv("ASG_PIPE_SUPER"); const long long x = e ? atoll(e) : 0; return x > 0 ? (int64_t)x : (int64_t)1 << 27; }();
  return v;
}

int d2h_fast(DevCtx &dc, void *dst, const void *dsrc, size_t bytes, cudaStream_t st);

struct Feeder {  // copy-in side of one super-batch, run by its own host thread
  std::atomic<int64_t> issued{0};  // pieces whose completion event has been recorded
  std::atomic<bool> abort{false};
  int rc = ASG_OK;
  std::string err;
};

int eval_super(asg_grid *g, int r, const double *X, int64_t cnt, int64_t sp, int64_t sdim, const double *fvals,
               int64_t max_depth, double *out) {
  USE_REPLICA(g, r);
  const int D = g->D;
  const int64_t P = pipe_piece();
  const int64_t np = (cnt + P - 1) / P;
  // device layout of the points: row-major unless the host matrix is column-major (Julia's M x D)
  const bool colmajor = (sp == 1 && D > 1 && sdim != 1);
  const bool rowlike = (sdim == 1 || D == 1);
  const int64_t dsp = colmajor ? 1 : D, dsd = colmajor ? cnt : 1;
  if (int rc = R.s_X[0].reserve((size_t)cnt * D)) return rc;
  if (int rc = R.s_out[0].reserve((size_t)cnt)) return rc;
  if (fvals)
    if (int rc = R.s_f[0].reserve((size_t)cnt)) return rc;
  // the ordering scratch of the largest possible window, so that no allocation happens while the pipeline runs
  if (g->path() == ASG_PATH_SUBSPACE && g->walk() == 3 && cnt >= sort_threshold() && cnt < ((int64_t)1 << 31) && g->info.subspaces >= 512) {
    if (int rc = R.s_key[0].reserve((size_t)cnt)) return rc;
    if (int rc = R.s_key2[0].reserve((size_t)cnt)) return rc;
    if (int rc = R.s_idx[0].reserve((size_t)cnt)) return rc;
    if (int rc = R.s_perm[0].reserve((size_t)cnt)) return rc;
    if (int rc = R.s_sorttmp[0].reserve(sort_points_temp_bytes(cnt))) return rc;
    if (block_pair_mode(R.dg, cnt, true)) {
      if (int rc = R.s_left[0].reserve((size_t)cnt + 2)) return rc;
      if (int rc = R.s_leftcnt[0].reserve(4)) return rc;
    }
  }
  while ((int64_t)R.piece_ev.size() < np) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    R.piece_ev.push_back(e);
  }
  for (int k = 0; k < 2; ++k)
    if (!R.walk_ev[k]) CK(cudaEventCreateWithFlags(&R.walk_ev[k], cudaEventDisableTiming));
  cudaStream_t s_in = dc.pipe[0], s_walk = dc.pipe[1], s_out = dc.pipe[2];
  double *dX = R.s_X[0].p, *dOut = R.s_out[0].p, *dF = fvals ? R.s_f[0].p : nullptr;

  // PAGEABLE row-major points (a numpy array, a Julia Matrix): cudaMemcpyAsync would go through the driver's own staging
  // at a fraction of the PCIe rate. Instead kStage host threads copy 16 MiB chunks into pinned buffers and the feeder
  // sends them on in order; the piece events are recorded as the chunks pass the piece boundaries.
  const size_t xbytes = (size_t)cnt * D * 8;
  std::unique_lock<std::mutex> stage_lk(dc.stage_in_mu, std::defer_lock);
  // (column-major points -- Julia's M x D Matrix -- are D contiguous segments per piece, staged the same way)
  bool staged = stage_host() && ((rowlike && sp == D) || colmajor) && xbytes >= 4 * kStagePiece && is_pageable(X) && stage_lk.try_lock();
  if (staged)
    for (int k = 0; k < DevCtx::kStage; ++k) {
      if (!dc.stage_in[k]) CK(cudaHostAlloc(&dc.stage_in[k], kStagePiece, cudaHostAllocDefault));
      if (!dc.stage_in_ev[k]) CK(cudaEventCreateWithFlags(&dc.stage_in_ev[k], cudaEventDisableTiming));
    }

  Feeder fd;
  std::thread feeder([&]() {
    if (cudaSetDevice(dc.device) != cudaSuccess) { fd.rc = ASG_ECUDA; fd.err = "cudaSetDevice failed in the copy thread"; fd.issued = np; return; }
    std::vector<double> repack;
    auto fail = [&](cudaError_t e) {
      fd.rc = e == cudaErrorMemoryAllocation ? ASG_ENOMEM : ASG_ECUDA;
      fd.err = std::string("host -> device copy: ") + cudaGetErrorString(e);
    };
    if (staged) {
      constexpr int K = DevCtx::kStage;
      // the copy as a list of contiguous segments of at most one staging buffer each; `pieces` = pipeline pieces that
      // are complete once the segment is on the stream
      struct Seg { const char *src; char *dst; size_t n; int64_t pieces; };
      std::vector<Seg> segs;
      if (colmajor) {
        for (int64_t k = 0; k < np; ++k) {
          const int64_t p0 = k * P, pn = std::min(P, cnt - p0);
          for (int d = 0; d < D; ++d)
            for (size_t o = 0; o < (size_t)pn * 8; o += kStagePiece)
              segs.push_back(Seg{(const char *)(X + p0 + (int64_t)d * sdim) + o, (char *)(dX + (size_t)d * cnt + p0) + o,
                                 std::min(kStagePiece, (size_t)pn * 8 - o), k});
          segs.back().pieces = k + 1;
        }
      } else {
        int64_t piece = 0;
        for (size_t off = 0; off < xbytes; off += kStagePiece) {
          const size_t n = std::min(kStagePiece, xbytes - off);
          while (piece < np && (size_t)std::min(cnt, (piece + 1) * P) * D * 8 <= off + n) ++piece;
          segs.push_back(Seg{(const char *)X + off, (char *)dX + off, n, piece});
        }
      }
      const int64_t nc = (int64_t)segs.size();
      std::atomic<int64_t> sent{0};       // segments whose H2D copy has been issued (in order)
      std::atomic<int64_t> filled[K];     // segments of buffer k that have been copied into it
      for (int k = 0; k < K; ++k) filled[k] = 0;
      std::atomic<int> bad{0};
      std::vector<std::thread> th;
      for (int k = 0; k < K; ++k)
        th.emplace_back([&, k]() {
          cudaSetDevice(dc.device);
          for (int64_t c = k; c < nc; c += K) {
            if (c >= K) {  // the buffer's previous segment must have been sent on ...
              while (sent.load(std::memory_order_acquire) <= c - K && !bad.load() && !fd.abort.load()) std::this_thread::yield();
              if (bad.load() || fd.abort.load()) return;
            }
            // ... and have left the buffer (for the first segments: whatever an earlier, aborted call left in flight)
            if (cudaEventSynchronize(dc.stage_in_ev[k]) != cudaSuccess) { bad = 1; return; }
            std::memcpy(dc.stage_in[k], segs[(size_t)c].src, segs[(size_t)c].n);
            filled[k].fetch_add(1, std::memory_order_release);
          }
        });
      cudaError_t e = cudaSuccess;
      int64_t piece = 0;  // next piece whose event is due
      for (int64_t c = 0; c < nc && e == cudaSuccess && !bad.load() && !fd.abort.load(); ++c) {
        const int k = (int)(c % K);
        while (filled[k].load(std::memory_order_acquire) <= c / K && !bad.load() && !fd.abort.load()) std::this_thread::yield();
        if (bad.load() || fd.abort.load()) break;
        e = cudaMemcpyAsync(segs[(size_t)c].dst, dc.stage_in[k], segs[(size_t)c].n, cudaMemcpyHostToDevice, s_in);
        if (e == cudaSuccess) e = cudaEventRecord(dc.stage_in_ev[k], s_in);
        if (e == cudaSuccess) sent.store(c + 1, std::memory_order_release);
        while (e == cudaSuccess && piece < segs[(size_t)c].pieces) {  // pieces that are complete on the stream now
          const int64_t p0 = piece * P, pn = std::min(P, cnt - p0);
          if (fvals) e = cudaMemcpyAsync(dF + p0, fvals + p0, (size_t)pn * 8, cudaMemcpyHostToDevice, s_in);
          if (e == cudaSuccess) e = cudaEventRecord(R.piece_ev[(size_t)piece], s_in);
          if (e == cudaSuccess) fd.issued.store(++piece, std::memory_order_release);
        }
      }
      if (e != cudaSuccess || bad.load()) {
        bad = 1;
        fail(e != cudaSuccess ? e : cudaErrorUnknown);
      }
      for (auto &t : th) t.join();
      fd.issued = np;
      return;
    }
    for (int64_t k = 0; k < np && !fd.abort.load(); ++k) {
      const int64_t p0 = k * P, n = std::min(P, cnt - p0);
      cudaError_t e = cudaSuccess;
      if (rowlike && sp == D) {
        e = cudaMemcpyAsync(dX + p0 * D, X + p0 * sp, (size_t)n * D * 8, cudaMemcpyHostToDevice, s_in);
      } else if (rowlike) {
        e = cudaMemcpy2DAsync(dX + p0 * D, (size_t)D * 8, X + p0 * sp, (size_t)sp * 8, (size_t)D * 8, (size_t)n, cudaMemcpyHostToDevice, s_in);
      } else if (colmajor) {
        for (int d = 0; d < D && e == cudaSuccess; ++d)
          e = cudaMemcpyAsync(dX + (size_t)d * cnt + p0, X + p0 + d * sdim, (size_t)n * 8, cudaMemcpyHostToDevice, s_in);
      } else {  // generic strides: repack on the host
        repack.resize((size_t)n * D);
        for (int64_t q = 0; q < n; ++q)
          for (int d = 0; d < D; ++d) repack[(size_t)q * D + d] = X[(p0 + q) * sp + d * sdim];
        e = cudaMemcpyAsync(dX + p0 * D, repack.data(), (size_t)n * D * 8, cudaMemcpyHostToDevice, s_in);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s_in);  // the staging vector is reused by the next piece
      }
      if (e == cudaSuccess && fvals) e = cudaMemcpyAsync(dF + p0, fvals + p0, (size_t)n * 8, cudaMemcpyHostToDevice, s_in);
      if (e == cudaSuccess) e = cudaEventRecord(R.piece_ev[(size_t)k], s_in);
      if (e != cudaSuccess) { fail(e); fd.issued = np; return; }
      fd.issued.store(k + 1, std::memory_order_release);
    }
    fd.issued = np;
  });
  auto finish = [&](int rc) {  // every exit joins the copy thread first
    if (rc != ASG_OK) fd.abort = true;
    feeder.join();
    return rc;
  };

  int64_t done = 0, prev_p0 = -1, prev_n = 0;
  int wk = 0;
  const bool out_pageable = stage_host() && is_pageable(out);
  while (done < np) {
    // at least one more piece: wait until its copy has been issued, then until it has arrived
    while (fd.issued.load(std::memory_order_acquire) <= done) std::this_thread::yield();
    if (fd.rc != ASG_OK) { const int rc = fd.rc; const std::string m = fd.err; finish(rc); return set_err(rc, m); }
    cudaError_t e = cudaEventSynchronize(R.piece_ev[(size_t)done]);
    if (e != cudaSuccess) { finish(ASG_ECUDA); return set_err(ASG_ECUDA, std::string("cudaEventSynchronize: ") + cudaGetErrorString(e)); }
    int64_t arrived = done + 1;
    const int64_t iss = fd.issued.load(std::memory_order_acquire);
    while (arrived < np && arrived < iss && cudaEventQuery(R.piece_ev[(size_t)arrived]) == cudaSuccess) ++arrived;
    cudaGetLastError();  // cudaErrorNotReady of the query is not an error
    // everything is here: split a short tail off the last window, so that little is left un-overlapped at the end
    if (arrived == np && np - done > 8) arrived = np - std::max<int64_t>(1, std::min<int64_t>(4, (np - done) / 8));
    const int64_t p0 = done * P, n = std::min(cnt, arrived * P) - p0;
    EvalArgs a{};
    a.X = colmajor ? dX + p0 : dX + p0 * D;
    a.sp = dsp;
    a.sd = dsd;
    a.M = n;
    a.rem = rem_of(max_depth);
    a.fvals = dF ? dF + p0 : nullptr;
    a.out = dOut + p0;
    if (int rc = run_eval(g, r, a, s_walk, 0)) return finish(rc);
    e = cudaEventRecord(R.walk_ev[wk], s_walk);
    // results of the PREVIOUS window go home while this one walks (its walk has finished: we waited for it below);
    // a pageable destination through the pinned staging of d2h_fast (returns when the copy is done)
    if (e == cudaSuccess && prev_p0 >= 0) {
      if (out_pageable) {
        if (int rc = d2h_fast(dc, out + prev_p0, dOut + prev_p0, (size_t)prev_n * 8, s_out)) return finish(rc);
      } else {
        e = cudaMemcpyAsync(out + prev_p0, dOut + prev_p0, (size_t)prev_n * 8, cudaMemcpyDeviceToHost, s_out);
      }
    }
    if (e == cudaSuccess) e = cudaEventSynchronize(R.walk_ev[wk]);
    if (e != cudaSuccess) { finish(ASG_ECUDA); return set_err(ASG_ECUDA, std::string("evaluation pipeline: ") + cudaGetErrorString(e)); }
    prev_p0 = p0;
    prev_n = n;
    wk ^= 1;
    done = arrived;
  }
  feeder.join();
  if (fd.rc != ASG_OK) return set_err(fd.rc, fd.err);
  if (prev_p0 >= 0) {
    if (out_pageable) {
      if (int rc = d2h_fast(dc, out + prev_p0, dOut + prev_p0, (size_t)prev_n * 8, s_out)) return rc;
    } else {
      CK(cudaMemcpyAsync(out + prev_p0, dOut + prev_p0, (size_t)prev_n * 8, cudaMemcpyDeviceToHost, s_out));
    }
  }
  CK(cudaStreamSynchronize(s_out));
  return ASG_OK;
}

int eval_host_dev(asg_grid *g, int r, const double *X, int64_t M, int64_t sp, int64_t sdim, const double *fvals,
                  int64_t max_depth, double *out) {
  if (int rc = use_dev(r)) return rc;
  // a super-batch keeps its points on the device: at most pipe_super() points and at most 8 GiB of coordinates
  const int64_t S = std::max<int64_t>(pipe_piece(), std::min<int64_t>(pipe_super(), ((int64_t)8 << 30) / (8 * (int64_t)g->D)));
  for (int64_t m0 = 0; m0 < M; m0 += S) {
    const int64_t cnt = std::min(S, M - m0);
    if (int rc = eval_super(g, r, X + m0 * sp, cnt, sp, sdim, fvals ? fvals + m0 : nullptr, max_depth, out + m0)) return rc;
  }
  return ASG_OK;
}

// minimum points per device before a host-buffer call is split over the devices of the context
int64_t multi_dev_min() {
  static const int64_t v = [] { const char *e = getenv("ASG_MULTI_MIN"); const long long x = e ? atoll(e) : 0; return x > 0 ? (int64_t)x : (int64_t)1 << 16; }();
  return v;
}

// Host-buffer evaluation over every device of the context (grid replicated, points split into
// contiguous ranges, results written straight into the caller's buffer; no collective). One host thread per device
// drives that device's pipeline. Small batches stay on the primary device.
int eval_host(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, const double *fvals,
              int64_t max_depth, double *out) {
  const int nd = (int)std::min<int64_t>(ctx.ndev, std::max<int64_t>(1, M / multi_dev_min()));
  if (nd <= 1) return eval_host_dev(g, 0, X, M, sp, sdim, fvals, max_depth, out);
  std::vector<int> rcs((size_t)nd, ASG_OK);
  std::vector<std::string> errs((size_t)nd);
  std::vector<std::thread> th;
  const int64_t per = (M + nd - 1) / nd;
  for (int r = 0; r < nd; ++r) {
    const int64_t lo = std::min<int64_t>(M, (int64_t)r * per), hi = std::min<int64_t>(M, lo + per);
    th.emplace_back([=, &rcs, &errs]() {
      if (hi > lo) rcs[(size_t)r] = eval_host_dev(g, r, X + lo * sp, hi - lo, sp, sdim, fvals ? fvals + lo : nullptr, max_depth, out + lo);
      if (rcs[(size_t)r] != ASG_OK) errs[(size_t)r] = t_err;  // thread-local message of the worker
    });
  }
  for (auto &t : th) t.join();
  for (int r = 0; r < nd; ++r)
    if (rcs[(size_t)r] != ASG_OK) return set_err(rcs[(size_t)r], "device " + std::to_string(ctx.dev[r].device) + ": " + errs[(size_t)r]);
  return use_dev(0);
}

// Device -> host copy of a LARGE result (basis matrices: hundreds of MB) that also runs at PCIe speed when the
// destination is ordinary pageable memory (a numpy array, a Julia Vector). cudaMemcpy to pageable memory goes through
// the driver's own staging at ~5 GB/s on this platform; here the data comes home in 16 MiB pieces through four pinned
// buffers and four host threads move the pieces to their destination in parallel (~25 GB/s). Blocks until done; work
// queued on `st` before the call is waited for like cudaMemcpy would. Pinned destinations take the direct path.
int d2h_fast(DevCtx &dc, void *dst, const void *dsrc, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return ASG_OK;
  // pieces of 16 MiB for large copies, 2 MiB for medium ones (the result windows of the evaluation pipeline)
  const size_t kPiece = bytes >= 4 * kStagePiece ? kStagePiece : (size_t)2 << 20;
  if (!stage_host() || !is_pageable(dst) || bytes < 4 * kPiece) {
    CK(cudaMemcpyAsync(dst, dsrc, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ASG_OK;
  }
  std::lock_guard<std::mutex> lk(dc.stage_mu);
  constexpr int K = DevCtx::kStage;
  for (int k = 0; k < K; ++k) {
    if (!dc.stage[k]) CK(cudaHostAlloc(&dc.stage[k], kStagePiece, cudaHostAllocDefault));
    if (!dc.stage_ev[k]) CK(cudaEventCreateWithFlags(&dc.stage_ev[k], cudaEventDisableTiming));
  }
  const int64_t np = (int64_t)((bytes + kPiece - 1) / kPiece);
  std::atomic<int64_t> issued{0};
  std::atomic<int64_t> done[K];
  for (int k = 0; k < K; ++k) done[k] = 0;  // pieces of buffer k that have reached their destination
  std::atomic<int> failed{0};
  const int device = dc.device;
  std::vector<std::thread> th;
  for (int k = 0; k < K; ++k)
    th.emplace_back([&, k]() {
      cudaSetDevice(device);
      for (int64_t c = k; c < np; c += K) {
        while (issued.load(std::memory_order_acquire) <= c && !failed.load()) std::this_thread::yield();
        if (failed.load()) return;
        if (cudaEventSynchronize(dc.stage_ev[k]) != cudaSuccess) { failed = 1; return; }
        const size_t off = (size_t)c * kPiece, n = std::min(kPiece, bytes - off);
        std::memcpy((char *)dst + off, dc.stage[k], n);
        done[k].fetch_add(1, std::memory_order_release);
      }
    });
  cudaError_t e = cudaSuccess;
  for (int64_t c = 0; c < np && e == cudaSuccess && !failed.load(); ++c) {
    const int k = (int)(c % K);
    while (done[k].load(std::memory_order_acquire) < c / K && !failed.load()) std::this_thread::yield();  // buffer k is free again
    const size_t off = (size_t)c * kPiece, n = std::min(kPiece, bytes - off);
    e = cudaMemcpyAsync(dc.stage[k], (const char *)dsrc + off, n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(dc.stage_ev[k], st);
    if (e == cudaSuccess) issued.store(c + 1, std::memory_order_release);
  }
  if (e != cudaSuccess) failed = 1;
  for (auto &t : th) t.join();
  if (e != cudaSuccess || failed.load()) return set_err(ASG_ECUDA, std::string("device -> host copy: ") + cudaGetErrorString(e));
  return ASG_OK;
}

int check_eval_args(asg_grid *g, const void *X, int64_t M, const void *out) {
  if (!g) return set_err(ASG_EINVAL, "null grid handle");
  if (M < 0) return set_err(ASG_EINVAL, "M < 0");
  if (M > 0 && (!X || !out)) return set_err(ASG_EINVAL, "null buffer");
  if (!g->uploaded) return set_err(ASG_ESTATE, "no nodes uploaded: call asg_grid_upload first");
  return ASG_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// exported functions
// ---------------------------------------------------------------------------------------------
extern "C" {

const char *asg_last_error(void) { return t_err.c_str(); }
const char *asg_version(void) { return "asg-b200 0.1.0 (sm_100a)"; }

int32_t asg_device_count(int32_t *count) {
  if (!count) return set_err(ASG_EINVAL, "null pointer");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return set_err(ASG_ENODEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *count = n;
  return ASG_OK;
}

static int init_locked(const int32_t *devices, int32_t ndev) {
  if (!devices || ndev <= 0) return set_err(ASG_EINVAL, "empty device list");
  if (ndev > kMaxDev) return set_err(ASG_EUNSUPPORTED, "more than 16 devices");
  if (ctx.ready) {
    bool same = ctx.ndev == ndev;
    for (int r = 0; same && r < ndev; ++r) same = ctx.dev[r].device == devices[r];
    if (same) return ASG_OK;
    return set_err(ASG_ESTATE, "already initialised on another device list (asg_shutdown first)");
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return set_err(ASG_ENODEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0") +
                                      " (libasg_cuda has no CPU fallback)");
  for (int r = 0; r < ndev; ++r) {
    const int device = devices[r];
    if (device < 0 || device >= n) return set_err(ASG_ENODEVICE, "device index out of range");
    for (int q = 0; q < r; ++q)
      if (devices[q] == device) return set_err(ASG_EINVAL, "device listed twice");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      return set_err(ASG_ENODEVICE, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                                        ", this library is built for sm_100a only");
  }
  for (int r = 0; r < ndev; ++r) {
    DevCtx &dc = ctx.dev[r];
    dc.device = devices[r];
    CK(cudaSetDevice(dc.device));
    CK(cudaStreamCreateWithFlags(&dc.stream, cudaStreamNonBlocking));
    for (int s = 0; s < kPipeSlots; ++s) CK(cudaStreamCreateWithFlags(&dc.pipe[s], cudaStreamNonBlocking));
    CK(cudaEventCreate(&dc.ev0));
    CK(cudaEventCreate(&dc.ev1));
    CK(cudaEventCreate(&dc.evw0));
    CK(cudaEventCreate(&dc.evw1));
  }
  CK(cudaSetDevice(ctx.dev[0].device));
  ctx.ndev = ndev;
  ctx.last_dev = 0;
  ctx.ready = true;
  return ASG_OK;
}

int32_t asg_init(int32_t device) {
  std::lock_guard<std::mutex> lk(ctx.mu);
  return init_locked(&device, 1);
}

int32_t asg_init_devices(const int32_t *devices, int32_t ndev) {
  std::lock_guard<std::mutex> lk(ctx.mu);
  return init_locked(devices, ndev);
}

int32_t asg_context_devices(int32_t *devices, int32_t capacity, int32_t *ndev) {
  if (!ndev) return set_err(ASG_EINVAL, "null pointer");
  std::lock_guard<std::mutex> lk(ctx.mu);
  *ndev = ctx.ready ? ctx.ndev : 0;
  for (int r = 0; devices && r < *ndev && r < capacity; ++r) devices[r] = ctx.dev[r].device;
  return ASG_OK;
}

int32_t asg_shutdown(void) {
  std::lock_guard<std::mutex> lk(ctx.mu);
  if (!ctx.ready) return ASG_OK;
  for (int r = 0; r < ctx.ndev; ++r) {
    DevCtx &dc = ctx.dev[r];
    cudaSetDevice(dc.device);
    cudaDeviceSynchronize();
    cudaStreamDestroy(dc.stream);
    for (int s = 0; s < kPipeSlots; ++s) cudaStreamDestroy(dc.pipe[s]);
    cudaEventDestroy(dc.ev0);
    cudaEventDestroy(dc.ev1);
    cudaEventDestroy(dc.evw0);
    cudaEventDestroy(dc.evw1);
    for (int k = 0; k < DevCtx::kStage; ++k) {
      if (dc.stage[k]) cudaFreeHost(dc.stage[k]);
      if (dc.stage_ev[k]) cudaEventDestroy(dc.stage_ev[k]);
      dc.stage[k] = nullptr;
      dc.stage_ev[k] = nullptr;
      if (dc.stage_in[k]) cudaFreeHost(dc.stage_in[k]);
      if (dc.stage_in_ev[k]) cudaEventDestroy(dc.stage_in_ev[k]);
      dc.stage_in[k] = nullptr;
      dc.stage_in_ev[k] = nullptr;
    }
    dc.device = -1;
    dc.stream = nullptr;
    dc.walk_timed = dc.last_timed = false;
  }
  ctx.ready = false;
  ctx.ndev = 0;
  return ASG_OK;
}

int32_t asg_launch_count(int64_t *count) {
  if (!count) return set_err(ASG_EINVAL, "null pointer");
  *count = ctx.launches.load();
  return ASG_OK;
}

int32_t asg_host_register(void *ptr, size_t bytes) {
  if (int rc = require_ctx()) return rc;
  if (!ptr || !bytes) return set_err(ASG_EINVAL, "null buffer");
  CK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return ASG_OK;
}
int32_t asg_host_unregister(void *ptr) {
  if (int rc = require_ctx()) return rc;
  if (!ptr) return set_err(ASG_EINVAL, "null buffer");
  CK(cudaHostUnregister(ptr));
  return ASG_OK;
}

int32_t asg_grid_create(int32_t D, const double *lb, const double *ub, asg_grid **out) {
  if (!out) return set_err(ASG_EINVAL, "null out pointer");
  *out = nullptr;
  if (D <= 0) return set_err(ASG_EINVAL, "Dimension must be positive.");  // src/class.jl:301
  if (D > kMaxDim) return set_err(ASG_EUNSUPPORTED, "D > 64");
  std::unique_ptr<asg_grid> g(new (std::nothrow) asg_grid());
  if (!g) return set_err(ASG_ENOMEM, "out of host memory");
  g->D = D;
  for (int d = 0; d < D; ++d) {
    g->lb[d] = lb ? lb[d] : 0.0;
    g->ub[d] = ub ? ub[d] : 1.0;
    if (!std::isfinite(g->lb[d])) return set_err(ASG_EINVAL, "Lower bound must be finite.");  // class.jl:305
    if (!std::isfinite(g->ub[d])) return set_err(ASG_EINVAL, "Upper bound must be finite.");  // class.jl:306
    if (!(g->ub[d] > g->lb[d])) return set_err(ASG_EINVAL, "Upper bound must be greater than lower bound.");
  }
  *out = g.release();
  return ASG_OK;
}

int32_t asg_grid_destroy(asg_grid *g) {
  if (!g) return ASG_OK;
  {
    std::lock_guard<std::recursive_mutex> lk(g->mu);
    if (ctx.ready) {
      for (int r = 0; r < ctx.ndev; ++r) {
        cudaSetDevice(ctx.dev[r].device);
        g->rep[r].release_all();
      }
      cudaSetDevice(ctx.dev[0].device);
    }
  }
  delete g;
  return ASG_OK;
}

int32_t asg_grid_upload(asg_grid *g, int64_t N, const int64_t *levels, const int64_t *indices, int64_t sn,
                        int64_t sd, const double *coefs) {
  ASG_TRACE();
  if (!g) return set_err(ASG_EINVAL, "null grid handle");
  if (N < 0 || (N > 0 && (!levels || !indices))) return set_err(ASG_EINVAL, "bad node arrays");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  g->uploaded = false;
  g->levels.clear(); g->indices.clear(); g->coefs.clear();
  {
    PackTimer t("copy of the caller's node arrays");
    copy_nodes(g, 0, N, levels, indices, sn, sd, coefs);
  }
  return repack(g);
}

// adapt! appends children of the deepest nodes (src/train.jl:283, 434-438). Such rows cannot repeat a node the grid
// holds (depth is a function of the level vector), so if they are canonical, finite and distinct they can join the
// sweep arrays without touching the packed layouts. Returns false when the full pack is needed.
static bool delta_rows_ok(const asg_grid *g, int64_t N0, int64_t n) {
  if (!g->uploaded || !g->canonical || !g->finite || g->N_packed <= 0 || g->forced_path == ASG_PATH_SWEEP) return false;
  static const bool off = getenv("ASG_APPEND_FULL_PACK") != nullptr;  // experiment knob: the round-1 behaviour
  if (off) return false;
  const int D = g->D;
  const int64_t delta = N0 + n - g->N_packed;
  if (delta > std::max<int64_t>(8192, g->N_packed / 2)) return false;
  std::vector<int64_t> order((size_t)n);
  for (int64_t k = 0; k < n; ++k) {
    order[(size_t)k] = k;
    int64_t dsum = 0;
    for (int d = 0; d < D; ++d) {
      const int64_t l = g->levels[(size_t)(N0 + k) * D + d], i = g->indices[(size_t)(N0 + k) * D + d];
      if (l < 1 || l > kMaxLevelFast) return false;
      if (l == 1 ? i != 1 : (l == 2 ? (i != 0 && i != 2) : (i < 1 || !(i & 1) || i > ((int64_t)1 << (l - 1)) - 1))) return false;
      dsum += l;
    }
    if (dsum - D + 1 <= g->info.max_depth) return false;  // could repeat an existing node
    if (!std::isfinite(g->coefs[(size_t)(N0 + k)])) return false;
  }
  const int64_t *L = g->levels.data() + (size_t)N0 * D, *I = g->indices.data() + (size_t)N0 * D;
  auto cmp = [&](int64_t a, int64_t b) {
    for (int d = 0; d < D; ++d) {
      if (L[a * D + d] != L[b * D + d]) return L[a * D + d] < L[b * D + d];
      if (I[a * D + d] != I[b * D + d]) return I[a * D + d] < I[b * D + d];
    }
    return false;
  };
  std::sort(order.begin(), order.end(), cmp);
  for (int64_t k = 1; k < n; ++k)
    if (!cmp(order[(size_t)k - 1], order[(size_t)k]) && !cmp(order[(size_t)k], order[(size_t)k - 1])) return false;  // repeated row
  return true;
}

static int append_delta(asg_grid *g, int64_t N0, int64_t n) {
  const int D = g->D;
  std::vector<SweepDim> sw((size_t)n * D);
  std::vector<uint64_t> low((size_t)n, 0);
  std::vector<int32_t> depth((size_t)n);
  int max_level = g->info.max_level;
  int64_t max_depth = g->info.max_depth;
  for (int64_t k = 0; k < n; ++k) {
    int64_t dsum = 0;
    for (int d = 0; d < D; ++d) {  // the (s, c) pairs of asg_pack.cpp: phi = hat(x * s - c)
      const int64_t l = g->levels[(size_t)(N0 + k) * D + d], i = g->indices[(size_t)(N0 + k) * D + d];
      SweepDim &q = sw[(size_t)k * D + d];
      if (l == 1) { q.s = 0.0; q.c = 0.0; low[(size_t)k] |= 1ull << d; }
      else { q.s = std::ldexp(1.0, (int)(l - 1)); q.c = (double)i; if (l == 2) low[(size_t)k] |= 1ull << d; }
      dsum += l;
      max_level = std::max(max_level, (int)l);
    }
    depth[(size_t)k] = (int32_t)(dsum - D + 1);
    max_depth = std::max<int64_t>(max_depth, dsum - D + 1);
  }
  const int64_t N1 = N0 + n;
  for (int r = 0; r < ctx.ndev; ++r) {
    if (int rc = use_dev(r)) return rc;
    USE_REPLICA(g, r);
    cudaStream_t st = dc.stream;
    if (int rc = R.d_sw.grow_keep((size_t)N1 * D, (size_t)N0 * D, st)) return rc;
    if (int rc = R.d_sw_low.grow_keep((size_t)N1, (size_t)N0, st)) return rc;
    if (int rc = R.d_sw_coef.grow_keep((size_t)N1, (size_t)N0, st)) return rc;
    if (int rc = R.d_sw_depth.grow_keep((size_t)N1, (size_t)N0, st)) return rc;
    CK(cudaMemcpyAsync(R.d_sw.p + (size_t)N0 * D, sw.data(), sw.size() * sizeof(SweepDim), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(R.d_sw_low.p + N0, low.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(R.d_sw_coef.p + N0, g->coefs.data() + N0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(R.d_sw_depth.p + N0, depth.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    R.dg.sw = R.d_sw.p;
    R.dg.sw_low = R.d_sw_low.p;
    R.dg.sw_coef = R.d_sw_coef.p;
    R.dg.sw_depth = R.d_sw_depth.p;
    R.dg.N = N1;
  }
  if (int rc = use_dev(0)) return rc;
  g->info.N = N1;
  g->info.max_level = max_level;
  g->info.max_depth = max_depth;
  g->info.delta_nodes = N1 - g->N_packed;
  g->info.generation++;
  return ASG_OK;
}

int32_t asg_grid_append(asg_grid *g, int64_t n, const int64_t *levels, const int64_t *indices, int64_t sn,
                        int64_t sd, const double *coefs) {
  ASG_TRACE();
  if (!g) return set_err(ASG_EINVAL, "null grid handle");
  if (n < 0 || (n > 0 && (!levels || !indices))) return set_err(ASG_EINVAL, "bad node arrays");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (n == 0) return ASG_OK;
  const int64_t N0 = g->N;
  copy_nodes(g, N0, n, levels, indices, sn, sd, coefs);
  int rc;
  if (delta_rows_ok(g, N0, n)) {
    rc = append_delta(g, N0, n);
    if (rc != ASG_OK) {  // device trouble: the host arrays shrink back; the packed layouts were never touched
      g->N = N0;
      g->levels.resize((size_t)N0 * g->D);
      g->indices.resize((size_t)N0 * g->D);
      g->coefs.resize((size_t)N0);
      for (int r = 0; r < ctx.ndev; ++r) g->rep[r].dg.N = N0;
    }
    return rc;
  }
  g->uploaded = false;
  rc = repack(g);
  if (rc != ASG_OK) {  // roll back: the reference would not have grown the grid either
    g->N = N0;
    g->levels.resize((size_t)N0 * g->D);
    g->indices.resize((size_t)N0 * g->D);
    g->coefs.resize((size_t)N0);
    std::string keep = t_err;
    if (N0 >= 0 && repack(g) == ASG_OK) t_err = keep;
  }
  return rc;
}

static int scatter_coefs_locked(asg_grid *g, int64_t n, const int64_t *idx, int64_t first, const double *coefs,
                                const double *dcoefs_or_null) {
  // idx == nullptr: nodes first .. first+n-1. dcoefs_or_null: the same values already on the PRIMARY device.
  if (n == 0) return ASG_OK;
  std::vector<long long> slots((size_t)n), bslots, nodes((size_t)n);
  bool any_hash = false, any_dense = false;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t node = idx ? idx[k] : first + k;
    if (node < 0 || node >= g->N) return set_err(ASG_EINVAL, "node number out of range");
    nodes[(size_t)k] = node;
    if (coefs) {
      g->coefs[(size_t)node] = coefs[k];
      if (!std::isfinite(coefs[k])) g->finite = false;
    }
    if (g->canonical) {
      if (node >= g->N_packed) { slots[(size_t)k] = LLONG_MIN; continue; }  // appended, not packed yet: sweep arrays only
      slots[(size_t)k] = g->slot_of_node[(size_t)node];
      (slots[(size_t)k] < 0 ? any_hash : any_dense) = true;
    }
  }
  if (g->canonical && g->blk_ok) {  // the BLOCK layout holds its own copy of the coefficients
    bslots.resize((size_t)n);
    for (int64_t k = 0; k < n; ++k)
      bslots[(size_t)k] = nodes[(size_t)k] >= g->N_packed ? -1 : g->bslot_of_node[(size_t)nodes[(size_t)k]];
  }
  int64_t l = 0;
  for (int r = 0; r < ctx.ndev; ++r) {
    if (int rc = use_dev(r)) return rc;
    USE_REPLICA(g, r);
    cudaStream_t st = dc.stream;
    if (int rc = R.s_i64a.reserve((size_t)n)) return rc;
    const double *dsrc = r == 0 ? dcoefs_or_null : nullptr;
    if (!dsrc) {
      if (!coefs) return set_err(ASG_EINVAL, "coefficient values are needed on the host to reach every device");
      if (int rc = R.s_tmp2.reserve((size_t)n)) return rc;
      CK(cudaMemcpyAsync(R.s_tmp2.p, coefs, (size_t)n * 8, cudaMemcpyHostToDevice, st));
      dsrc = R.s_tmp2.p;
    }
    CK(cudaMemcpyAsync(R.s_i64a.p, nodes.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(launch_scatter_f64(R.d_sw_coef.p, R.s_i64a.p, dsrc, n, st, &l));
    if (g->canonical) {
      if (int rc = R.s_i64b.reserve((size_t)n)) return rc;
      CK(cudaMemcpyAsync(R.s_i64b.p, slots.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
      if (any_dense) CK(launch_scatter_f64(R.d_coef.p, R.s_i64b.p, dsrc, n, st, &l));
      if (any_hash) CK(launch_scatter_hash(R.d_hent.p, R.s_i64b.p, dsrc, n, st, &l));
      if (g->blk_ok) {
        CK(cudaMemcpyAsync(R.s_i64a.p, bslots.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));  // stream order: after the sweep scatter
        CK(launch_scatter_f64(R.d_bcoef.p, R.s_i64a.p, dsrc, n, st, &l));
      }
    }
  }
  bump(l);
  for (int r = 0; r < ctx.ndev; ++r) {
    if (int rc = use_dev(r)) return rc;
    CK(cudaStreamSynchronize(ctx.dev[r].stream));
  }
  if (int rc = use_dev(0)) return rc;
  g->info.generation++;
  return ASG_OK;
}

int32_t asg_grid_set_coefs(asg_grid *g, int64_t first, int64_t n, const double *coefs) {
  ASG_TRACE();
  if (!g || (n > 0 && !coefs)) return set_err(ASG_EINVAL, "null argument");
  if (n < 0 || first < 0) return set_err(ASG_EINVAL, "negative range");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (!g->uploaded) return set_err(ASG_ESTATE, "no nodes uploaded");
  if (first + n > g->N) return set_err(ASG_EINVAL, "range exceeds node count");
  int rc = scatter_coefs_locked(g, n, nullptr, first, coefs, nullptr);
  if (rc == ASG_OK && !g->finite) {  // re-derive: a full overwrite may have removed the non-finite values
    g->finite = true;
    for (double c : g->coefs)
      if (!std::isfinite(c)) { g->finite = false; break; }
  }
  return rc;
}

int32_t asg_grid_scatter_coefs(asg_grid *g, int64_t n, const int64_t *idx, const double *coefs) {
  ASG_TRACE();
  if (!g || (n > 0 && (!coefs || !idx))) return set_err(ASG_EINVAL, "null argument");
  if (n < 0) return set_err(ASG_EINVAL, "negative count");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (!g->uploaded) return set_err(ASG_ESTATE, "no nodes uploaded");
  return scatter_coefs_locked(g, n, idx, 0, coefs, nullptr);
}

int32_t asg_grid_get_coefs(asg_grid *g, int64_t first, int64_t n, double *coefs) {
  if (!g || (n > 0 && !coefs)) return set_err(ASG_EINVAL, "null argument");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  USE_REPLICA(g, 0);
  if (!g->uploaded) return set_err(ASG_ESTATE, "no nodes uploaded");
  if (first < 0 || n < 0 || first + n > g->N) return set_err(ASG_EINVAL, "range exceeds node count");
  // read the DEVICE copy (caller's node order) so that this is a true read-back of what the kernels see
  if (n) CK(cudaMemcpy(coefs, R.d_sw_coef.p + first, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return ASG_OK;
}

int32_t asg_grid_info(asg_grid *g, asg_grid_info_t *info) {
  if (!g || !info) return set_err(ASG_EINVAL, "null argument");
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  *info = g->info;
  info->D = g->D;
  info->N = g->N;
  info->packs = g->packs;
  info->delta_nodes = g->uploaded ? g->N - g->N_packed : 0;
  info->path = g->uploaded ? g->path() : 0;
  return ASG_OK;
}

int32_t asg_grid_set_path(asg_grid *g, int32_t path) {
  if (!g) return set_err(ASG_EINVAL, "null grid handle");
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  USE_REPLICA(g, 0);
  if (path < ASG_PATH_AUTO || path > ASG_PATH_BLOCK) return set_err(ASG_EINVAL, "bad path");
  if (path >= ASG_PATH_SUBSPACE && g->uploaded && g->N > 0 && !(g->canonical && g->finite))
    return set_err(ASG_EUNSUPPORTED, "grid is not canonical (" + g->why + "): only the sweep engine applies");
  if (path == ASG_PATH_BLOCK && g->uploaded && g->N > 0 && !(g->blk_ok && block_kernel_fits(R.dg)))
    return set_err(ASG_EUNSUPPORTED, "grid has no BLOCK layout (" + g->blk_why + ")");
  g->forced_path = path;
  return ASG_OK;
}

int32_t asg_eval(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, int64_t max_depth, double *out) {
  ASG_TRACE();
  if (int rc = check_eval_args(g, X, M, out)) return rc;
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (int rc = maybe_consolidate(g, M)) return rc;
  return eval_host(g, X, M, sp, sdim, nullptr, max_depth, out);
}

int32_t asg_surplus(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, const double *fvals,
                    int64_t max_depth, double *alpha) {
  ASG_TRACE();
  if (int rc = check_eval_args(g, X, M, alpha)) return rc;
  if (M > 0 && !fvals) return set_err(ASG_EINVAL, "null fvals");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (int rc = maybe_consolidate(g, M)) return rc;
  return eval_host(g, X, M, sp, sdim, fvals, max_depth, alpha);
}

// replica whose device owns a device pointer (device-pointer entry points work on any device of the context)
static int replica_of_pointer(const void *p) {
  if (ctx.ndev <= 1) return 0;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (at.type == cudaMemoryTypeDevice)
    for (int r = 0; r < ctx.ndev; ++r)
      if (ctx.dev[r].device == at.device) return r;
  return at.type == cudaMemoryTypeDevice ? -1 : 0;
}

static int eval_device_common(asg_grid *g, const double *dX, int64_t M, int64_t sp, int64_t sdim, const double *dfvals,
                              int64_t max_depth, double *dout, void *stream) {
  const int r = replica_of_pointer(dX);
  if (r < 0) return set_err(ASG_EINVAL, "device pointer belongs to a device outside the context (asg_init_devices)");
  if (int rc = use_dev(r)) return rc;
  USE_REPLICA(g, r);
  cudaStream_t st = (cudaStream_t)stream;  // NULL is the CUDA default stream, as everywhere in CUDA
  EvalArgs a{dX, M, sp, sdim, rem_of(max_depth), dfvals, dout};
  CK(cudaEventRecord(dc.ev0, st));
  if (int rc = run_eval(g, r, a, st, 0, true)) return rc;
  CK(cudaEventRecord(dc.ev1, st));
  dc.last_timed = true;
  ctx.last_dev = r;
  return ASG_OK;
}

int32_t asg_eval_device(asg_grid *g, const double *dX, int64_t M, int64_t sp, int64_t sdim, int64_t max_depth,
                        double *dout, void *stream) {
  ASG_TRACE();
  if (int rc = check_eval_args(g, dX, M, dout)) return rc;
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (int rc = maybe_consolidate(g, M)) return rc;
  return eval_device_common(g, dX, M, sp, sdim, nullptr, max_depth, dout, stream);
}

int32_t asg_surplus_device(asg_grid *g, const double *dX, int64_t M, int64_t sp, int64_t sdim, const double *dfvals,
                           int64_t max_depth, double *dalpha, void *stream) {
  ASG_TRACE();
  if (int rc = check_eval_args(g, dX, M, dalpha)) return rc;
  if (M > 0 && !dfvals) return set_err(ASG_EINVAL, "null fvals");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (int rc = maybe_consolidate(g, M)) return rc;
  return eval_device_common(g, dX, M, sp, sdim, dfvals, max_depth, dalpha, stream);
}

int32_t asg_last_kernel_ms(double *ms) {
  if (!ms) return set_err(ASG_EINVAL, "null pointer");
  if (int rc = require_ctx()) return rc;
  DevCtx &dc = ctx.dev[ctx.last_dev];
  if (!dc.last_timed) return set_err(ASG_ESTATE, "no timed launch yet");
  if (int rc = use_dev(ctx.last_dev)) return rc;
  CK(cudaEventSynchronize(dc.ev1));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, dc.ev0, dc.ev1));
  *ms = (double)f;
  return use_dev(0);
}

int32_t asg_last_walk_ms(double *ms) {
  if (!ms) return set_err(ASG_EINVAL, "null pointer");
  if (int rc = require_ctx()) return rc;
  DevCtx &dc = ctx.dev[ctx.last_dev];
  if (!dc.last_timed) return set_err(ASG_ESTATE, "no timed launch yet");
  if (int rc = use_dev(ctx.last_dev)) return rc;
  cudaEvent_t a = dc.walk_timed ? dc.evw0 : dc.ev0, b = dc.walk_timed ? dc.evw1 : dc.ev1;
  CK(cudaEventSynchronize(b));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, a, b));
  *ms = (double)f;
  return use_dev(0);
}

int32_t asg_eval_async(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, int64_t max_depth,
                       double *out, asg_event **ev) {
  if (!ev) return set_err(ASG_EINVAL, "null event pointer");
  *ev = nullptr;
  if (int rc = check_eval_args(g, X, M, out)) return rc;
  if (int rc = require_ctx()) return rc;
  asg_event *e = new (std::nothrow) asg_event();
  if (!e) return set_err(ASG_ENOMEM, "out of host memory");
  e->th = std::thread([=]() {
    e->rc = asg_eval(g, X, M, sp, sdim, max_depth, out);
    if (e->rc != ASG_OK) e->err = asg_last_error();
  });
  *ev = e;
  return ASG_OK;
}

int32_t asg_wait(asg_event *ev) {
  if (!ev) return set_err(ASG_EINVAL, "null event");
  if (ev->th.joinable()) ev->th.join();
  const int rc = ev->rc;
  if (rc != ASG_OK) t_err = ev->err;
  delete ev;
  return rc;
}

static int nodes_x_device(asg_grid *g, int r, int64_t n, const long long *h_lev, const long long *h_idx, double *dX,
                          int64_t sp, int64_t sd, cudaStream_t st) {
  USE_REPLICA(g, r);
  const int D = g->D;
  if (int rc = R.s_i64a.reserve((size_t)n * D)) return rc;
  if (int rc = R.s_i64b.reserve((size_t)n * D)) return rc;
  if (int rc = R.s_flag.reserve(1)) return rc;
  CK(cudaMemsetAsync(R.s_flag.p, 0, sizeof(int), st));
  CK(cudaMemcpyAsync(R.s_i64a.p, h_lev, (size_t)n * D * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(R.s_i64b.p, h_idx, (size_t)n * D * 8, cudaMemcpyHostToDevice, st));
  int64_t l = 0;
  // the kernel only reads lb / width of the DeviceGrid, valid even before any upload
  DeviceGrid dg = R.dg;
  dg.D = D;
  for (int d = 0; d < D; ++d) {
    dg.lb[d] = g->lb[d];
    volatile double w = g->ub[d] - g->lb[d];
    dg.width[d] = w;
  }
  CK(launch_nodes_x(D, nullptr, dg, R.s_i64a.p, R.s_i64b.p, n, dX, sp, sd, R.s_flag.p, st, &l));
  bump(l);
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, R.s_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (bad) return set_err(ASG_EINVALID_NODE, "invalid level-index pair (get_x, src/math.jl:474)");
  return ASG_OK;
}

static int copy_out_matrix(const double *dX, int64_t n, int D, double *Xout, int64_t sp, int64_t sdim, cudaStream_t st) {
  // device block is row-major n x D
  if (sdim == 1 || D == 1) {
    CK(cudaMemcpy2DAsync(Xout, (size_t)sp * 8, dX, (size_t)D * 8, (size_t)D * 8, (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  } else {
    std::vector<double> tmp((size_t)n * D);
    CK(cudaMemcpyAsync(tmp.data(), dX, (size_t)n * D * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int64_t k = 0; k < n; ++k)
      for (int d = 0; d < D; ++d) Xout[k * sp + d * sdim] = tmp[(size_t)k * D + d];
  }
  return ASG_OK;
}

int32_t asg_nodes_x(asg_grid *g, int64_t first, int64_t n, double *Xout, int64_t sp, int64_t sdim) {
  ASG_TRACE();
  if (!g || (n > 0 && !Xout)) return set_err(ASG_EINVAL, "null argument");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  USE_REPLICA(g, 0);
  if (first < 0 || n < 0 || first + n > g->N) return set_err(ASG_EINVAL, "range exceeds node count");
  if (n == 0) return ASG_OK;
  const int D = g->D;
  if (int rc = R.s_tmp.reserve((size_t)n * D)) return rc;
  if (int rc = nodes_x_device(g, 0, n, (const long long *)g->levels.data() + first * D,
                              (const long long *)g->indices.data() + first * D, R.s_tmp.p, D, 1, dc.stream))
    return rc;
  return copy_out_matrix(R.s_tmp.p, n, D, Xout, sp, sdim, dc.stream);
}

int32_t asg_nodes_x_li(asg_grid *g, int64_t n, const int64_t *levels, const int64_t *indices, int64_t sn, int64_t sd,
                       double *Xout, int64_t sp, int64_t sdim) {
  ASG_TRACE();
  if (!g || (n > 0 && (!Xout || !levels || !indices))) return set_err(ASG_EINVAL, "null argument");
  if (n < 0) return set_err(ASG_EINVAL, "negative count");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  USE_REPLICA(g, 0);
  if (n == 0) return ASG_OK;
  const int D = g->D;
  std::vector<long long> L((size_t)n * D), I((size_t)n * D);
  for (int64_t k = 0; k < n; ++k)
    for (int d = 0; d < D; ++d) {
      L[(size_t)k * D + d] = levels[k * sn + d * sd];
      I[(size_t)k * D + d] = indices[k * sn + d * sd];
    }
  if (int rc = R.s_tmp.reserve((size_t)n * D)) return rc;
  if (int rc = nodes_x_device(g, 0, n, L.data(), I.data(), R.s_tmp.p, D, 1, dc.stream)) return rc;
  return copy_out_matrix(R.s_tmp.p, n, D, Xout, sp, sdim, dc.stream);
}

// surplus at n of the grid's own nodes on replica r: get_x on the device, fused alpha = f - G epilogue, D2H
static int surplus_nodes_dev(asg_grid *g, int r, int64_t n, const long long *L, const long long *I, const double *fvals,
                             int64_t max_depth, double *alpha) {
  if (int rc = use_dev(r)) return rc;
  USE_REPLICA(g, r);
  const int D = g->D;
  cudaStream_t st = dc.stream;
  if (int rc = R.s_tmp.reserve((size_t)n * D)) return rc;
  if (int rc = nodes_x_device(g, r, n, L, I, R.s_tmp.p, D, 1, st)) return rc;  // get_x on device
  if (int rc = R.s_f[0].reserve((size_t)n)) return rc;
  if (int rc = R.s_out[0].reserve((size_t)n)) return rc;
  CK(cudaMemcpyAsync(R.s_f[0].p, fvals, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  EvalArgs a{R.s_tmp.p, n, D, 1, rem_of(max_depth), R.s_f[0].p, R.s_out[0].p};
  if (int rc = run_eval(g, r, a, st)) return rc;
  CK(cudaMemcpyAsync(alpha, R.s_out[0].p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return ASG_OK;
}

int32_t asg_surplus_nodes(asg_grid *g, int64_t n, const int64_t *idx, const double *fvals, int64_t max_depth,
                          int32_t store, double *alpha) {
  ASG_TRACE();
  if (!g || (n > 0 && (!idx || !fvals || !alpha))) return set_err(ASG_EINVAL, "null argument");
  if (n < 0) return set_err(ASG_EINVAL, "negative count");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (!g->uploaded) return set_err(ASG_ESTATE, "no nodes uploaded");
  if (n == 0) return ASG_OK;
  const int D = g->D;
  std::vector<long long> L((size_t)n * D), I((size_t)n * D);
  for (int64_t k = 0; k < n; ++k) {
    if (idx[k] < 0 || idx[k] >= g->N) return set_err(ASG_EINVAL, "node number out of range");
    std::memcpy(&L[(size_t)k * D], &g->levels[(size_t)idx[k] * D], (size_t)D * 8);
    std::memcpy(&I[(size_t)k * D], &g->indices[(size_t)idx[k] * D], (size_t)D * 8);
  }
  if (int rc = maybe_consolidate(g, n)) return rc;
  // the nodes of one depth are independent query points: split them over the devices like asg_eval does
  const int nd = (int)std::min<int64_t>(ctx.ndev, std::max<int64_t>(1, n / multi_dev_min()));
  if (nd <= 1) {
    if (int rc = surplus_nodes_dev(g, 0, n, L.data(), I.data(), fvals, max_depth, alpha)) return rc;
  } else {
    std::vector<int> rcs((size_t)nd, ASG_OK);
    std::vector<std::string> errs((size_t)nd);
    std::vector<std::thread> th;
    const int64_t per = (n + nd - 1) / nd;
    for (int r = 0; r < nd; ++r) {
      const int64_t lo = std::min<int64_t>(n, (int64_t)r * per), hi = std::min<int64_t>(n, lo + per);
      th.emplace_back([=, &rcs, &errs, &L, &I]() {
        if (hi > lo)
          rcs[(size_t)r] = surplus_nodes_dev(g, r, hi - lo, L.data() + lo * D, I.data() + lo * D, fvals + lo, max_depth, alpha + lo);
        if (rcs[(size_t)r] != ASG_OK) errs[(size_t)r] = t_err;
      });
    }
    for (auto &t : th) t.join();
    if (int rc = use_dev(0)) return rc;
    for (int r = 0; r < nd; ++r)
      if (rcs[(size_t)r] != ASG_OK) return set_err(rcs[(size_t)r], "device " + std::to_string(ctx.dev[r].device) + ": " + errs[(size_t)r]);
  }
  if (store) {
    // alpha is on the host: keep the host mirror, the finite flag and the slots of every replica in step
    return scatter_coefs_locked(g, n, idx, 0, alpha, (nd <= 1) ? g->rep[0].s_out[0].p : nullptr);
  }
  return ASG_OK;
}

// ---- phi: the exported basis function itself (src/math.jl:332-391) -------------------------------------------
int32_t asg_phi(int64_t n, int32_t D, const double *x01, const int64_t *levels, const int64_t *indices, double *out) {
  ASG_TRACE();
  if (n < 0 || D <= 0 || D > kMaxDim) return set_err(ASG_EINVAL, "bad size");
  if (n > 0 && (!x01 || !out)) return set_err(ASG_EINVAL, "null buffer");
  if ((levels == nullptr) != (indices == nullptr)) return set_err(ASG_EINVAL, "levels and indices go together");
  if (int rc = require_ctx()) return rc;
  if (n == 0) return ASG_OK;
  DevCtx &dc = ctx.dev[0];
  cudaStream_t st = dc.stream;
  const size_t e = (size_t)n * D;
  double *dx = nullptr, *dout = nullptr;
  long long *dl = nullptr, *di = nullptr;
  int *dbad = nullptr;
  int bad = 0, rc = ASG_OK;
  cudaError_t ce = cudaSuccess;
  auto ok = [&](cudaError_t c) { if (ce == cudaSuccess) ce = c; return ce == cudaSuccess; };
  if (ok(cudaMalloc(&dx, e * 8)) && ok(cudaMalloc(&dout, (size_t)n * 8)) && ok(cudaMalloc(&dbad, sizeof(int))) &&
      (!levels || (ok(cudaMalloc(&dl, e * 8)) && ok(cudaMalloc(&di, e * 8))))) {
    ok(cudaMemcpyAsync(dx, x01, e * 8, cudaMemcpyHostToDevice, st));
    if (levels) {
      ok(cudaMemcpyAsync(dl, levels, e * 8, cudaMemcpyHostToDevice, st));
      ok(cudaMemcpyAsync(di, indices, e * 8, cudaMemcpyHostToDevice, st));
    }
    ok(cudaMemsetAsync(dbad, 0, sizeof(int), st));
    int64_t l = 0;
    ok(launch_phi(D, n, dx, dl, di, dout, dbad, st, &l));
    bump(l);
    ok(cudaMemcpyAsync(out, dout, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    ok(cudaMemcpyAsync(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost, st));
    ok(cudaStreamSynchronize(st));
  }
  cudaFree(dx); cudaFree(dout); cudaFree(dl); cudaFree(di); cudaFree(dbad);
  if (ce != cudaSuccess) rc = set_err(ce == cudaErrorMemoryAllocation ? ASG_ENOMEM : ASG_ECUDA, std::string("asg_phi: ") + cudaGetErrorString(ce));
  else if (bad == 1) rc = set_err(ASG_EINVALID_NODE, "invalid level-index pair (phi, src/math.jl:355)");
  else if (bad == 2) rc = set_err(ASG_EUNSUPPORTED, "level > 62: power2(l - 1) leaves the Int64 range");
  return rc;
}

// ---- adapt! candidate generation ----------------------------------------------------------------
int32_t asg_children(asg_grid *g, int64_t depth, int64_t *n_parents, int64_t *n_candidates, int64_t *n_unique) {
  ASG_TRACE();
  if (!g || !n_unique) return set_err(ASG_EINVAL, "null argument");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  USE_REPLICA(g, 0);
  if (!g->uploaded) return set_err(ASG_ESTATE, "no nodes uploaded");
  const int D = g->D;
  std::vector<long long> PL, PI;
  for (int64_t n = 0; n < g->N; ++n) {  // parents = nodes of that depth (src/train.jl:316-318), in node order
    int64_t dsum = 0;
    for (int d = 0; d < D; ++d) dsum += g->levels[(size_t)n * D + d];
    if (dsum - D + 1 != depth) continue;
    for (int d = 0; d < D; ++d) {
      const int64_t l = g->levels[(size_t)n * D + d], i = g->indices[(size_t)n * D + d];
      if (l < 1 || (l == 2 && i != 0 && i != 2))  // child() throws on these, src/math.jl:919, 949
        return set_err(ASG_EINVALID_NODE, "invalid level-index pair among the parents");
      PL.push_back(l);
      PI.push_back(i);
    }
  }
  const int64_t n_p = (int64_t)PL.size() / D, n_cand = n_p * 2 * D;
  if (n_parents) *n_parents = n_p;
  if (n_candidates) *n_candidates = n_cand;
  *n_unique = 0;
  g->ch_n = 0;
  g->ch_generation = g->info.generation;
  if (n_cand == 0) return ASG_OK;
  if (n_cand >= ((int64_t)1 << 30)) return set_err(ASG_EUNSUPPORTED, "more than 2^30 candidates in one level");
  int64_t tsize = 1024;
  while (tsize < 2 * n_cand) tsize <<= 1;
  const int64_t nblk = (n_cand + 1023) / 1024;
  if (int rc = R.c_PL.reserve((size_t)n_p * D)) return rc;
  if (int rc = R.c_PI.reserve((size_t)n_p * D)) return rc;
  if (int rc = R.c_CL.reserve((size_t)n_cand * D)) return rc;
  if (int rc = R.c_CI.reserve((size_t)n_cand * D)) return rc;
  if (int rc = R.c_OL.reserve((size_t)n_cand * D)) return rc;
  if (int rc = R.c_OI.reserve((size_t)n_cand * D)) return rc;
  if (int rc = R.c_valid.reserve((size_t)n_cand)) return rc;
  if (int rc = R.c_table.reserve((size_t)tsize)) return rc;
  if (int rc = R.c_flags.reserve((size_t)n_cand)) return rc;
  if (int rc = R.c_sums.reserve((size_t)nblk + 1)) return rc;
  cudaStream_t st = dc.stream;
  CK(cudaMemcpyAsync(R.c_PL.p, PL.data(), PL.size() * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(R.c_PI.p, PI.data(), PI.size() * 8, cudaMemcpyHostToDevice, st));
  int64_t l = 0;
  unsigned int *total = R.c_sums.p + nblk;
  CK(launch_children(D, n_p, R.c_PL.p, R.c_PI.p, R.c_CL.p, R.c_CI.p, R.c_valid.p, R.c_table.p, tsize, R.c_flags.p,
                     R.c_sums.p, total, R.c_OL.p, R.c_OI.p, st, &l));
  bump(l);
  unsigned int nu = 0;
  CK(cudaMemcpyAsync(&nu, total, sizeof nu, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  g->ch_n = (int64_t)nu;
  *n_unique = g->ch_n;
  return ASG_OK;
}

int32_t asg_children_fetch(asg_grid *g, int64_t n, int64_t *levels, int64_t *indices, int64_t sn, int64_t sd, double *X,
                           int64_t sp, int64_t sdim) {
  ASG_TRACE();
  if (!g || (n > 0 && (!levels || !indices))) return set_err(ASG_EINVAL, "null argument");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  USE_REPLICA(g, 0);
  if (g->ch_n < 0 || g->ch_generation != g->info.generation)
    return set_err(ASG_ESTATE, "asg_children has not been called on this state of the grid");
  if (n != g->ch_n) return set_err(ASG_EINVAL, "n differs from the count asg_children returned");
  if (n == 0) return ASG_OK;
  const int D = g->D;
  cudaStream_t st = dc.stream;
  std::vector<long long> L((size_t)n * D), I((size_t)n * D);
  CK(cudaMemcpyAsync(L.data(), R.c_OL.p, (size_t)n * D * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(I.data(), R.c_OI.p, (size_t)n * D * 8, cudaMemcpyDeviceToHost, st));
  if (X) {  // coordinates of the children: the bit-exact get_x of the evaluation path (mul + add, never fused)
    if (int rc = R.s_tmp.reserve((size_t)n * D)) return rc;
    if (int rc = R.s_flag.reserve(1)) return rc;
    CK(cudaMemsetAsync(R.s_flag.p, 0, sizeof(int), st));
    DeviceGrid dg = R.dg;
    dg.D = D;
    for (int d = 0; d < D; ++d) {
      dg.lb[d] = g->lb[d];
      volatile double w = g->ub[d] - g->lb[d];
      dg.width[d] = w;
    }
    int64_t l = 0;
    CK(launch_nodes_x(D, nullptr, dg, R.c_OL.p, R.c_OI.p, n, R.s_tmp.p, D, 1, R.s_flag.p, st, &l));
    bump(l);
    if (int rc = copy_out_matrix(R.s_tmp.p, n, D, X, sp, sdim, st)) return rc;
  }
  CK(cudaStreamSynchronize(st));
  for (int64_t k = 0; k < n; ++k)
    for (int d = 0; d < D; ++d) {
      levels[k * sn + d * sd] = L[(size_t)k * D + d];
      indices[k * sn + d * sd] = I[(size_t)k * D + d];
    }
  return ASG_OK;
}

// ---- basis ------------------------------------------------------------------------------------
static int upload_points(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, int64_t &dsp, int64_t &dsd,
                         cudaStream_t st) {
  USE_REPLICA(g, 0);
  std::vector<double> rb;
  if (int rc = R.s_X[0].reserve((size_t)M * g->D)) return rc;
  if (int rc = stage_points(X, 0, M, g->D, sp, sdim, R.s_X[0].p, dsp, dsd, st, rb)) return rc;
  CK(cudaStreamSynchronize(st));
  return ASG_OK;
}

static cudaError_t run_basis(asg_grid *g, const BasisArgs &a, cudaStream_t st) {
  USE_REPLICA(g, 0);
  int64_t l = 0;
  cudaError_t e = (g->path() == ASG_PATH_SUBSPACE) ? launch_basis_subspace(R.dg, a, st, &l)
                                                   : launch_basis_sweep(R.dg, a, st, &l);
  bump(l);
  return e;
}

int32_t asg_basis_count(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                        int32_t dropzeros, int64_t *row_nnz) {
  ASG_TRACE();
  if (int rc = check_eval_args(g, X, M, row_nnz)) return rc;
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (int rc = consolidate(g)) return rc;  // appended nodes join the packed layouts first
  USE_REPLICA(g, 0);
  if (M == 0) return ASG_OK;
  cudaStream_t st = dc.stream;
  BasisArgs a{};
  if (int rc = upload_points(g, X, M, sp, sdim, a.sp, a.sd, st)) return rc;
  if (int rc = R.s_i64a.reserve((size_t)M)) return rc;
  a.X = R.s_X[0].p; a.M = M; a.atol = atol; a.dropzeros = dropzeros; a.mode = 0;
  a.row_nnz = (unsigned long long *)R.s_i64a.p;
  CK(run_basis(g, a, st));
  CK(cudaMemcpyAsync(row_nnz, R.s_i64a.p, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return ASG_OK;
}

int32_t asg_basis_fill(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                       int32_t dropzeros, const int64_t *row_ptr, int64_t *colidx, double *vals) {
  ASG_TRACE();
  if (int rc = check_eval_args(g, X, M, row_ptr)) return rc;
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (int rc = consolidate(g)) return rc;  // appended nodes join the packed layouts first
  USE_REPLICA(g, 0);
  if (M == 0) return ASG_OK;
  const int64_t nnz = row_ptr[M];
  if (nnz < 0 || (nnz > 0 && (!colidx || !vals))) return set_err(ASG_EINVAL, "bad CSR buffers");
  cudaStream_t st = dc.stream;
  BasisArgs a{};
  if (int rc = upload_points(g, X, M, sp, sdim, a.sp, a.sd, st)) return rc;
  if (int rc = R.s_i64a.reserve((size_t)M + 1)) return rc;
  if (int rc = R.s_i64b.reserve((size_t)nnz)) return rc;
  if (int rc = R.s_tmp.reserve((size_t)nnz)) return rc;
  CK(cudaMemcpyAsync(R.s_i64a.p, row_ptr, (size_t)(M + 1) * 8, cudaMemcpyHostToDevice, st));
  a.X = R.s_X[0].p; a.M = M; a.atol = atol; a.dropzeros = dropzeros; a.mode = 1;
  a.row_ptr = R.s_i64a.p; a.colidx = R.s_i64b.p; a.vals = R.s_tmp.p;
  CK(run_basis(g, a, st));
  if (g->path() == ASG_PATH_SUBSPACE) {  // trie order -> ascending node numbers inside each row
    int64_t l = 0;
    CK(launch_sort_rows(R.s_i64a.p, R.s_i64b.p, R.s_tmp.p, M, st, &l));
    bump(l);
  }
  if (nnz) {
    if (int rc = d2h_fast(dc, colidx, R.s_i64b.p, (size_t)nnz * 8, st)) return rc;
    if (int rc = d2h_fast(dc, vals, R.s_tmp.p, (size_t)nnz * 8, st)) return rc;
  }
  CK(cudaStreamSynchronize(st));
  return ASG_OK;
}

static int basis_fill_csc_common(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                                 int32_t dropzeros, const int64_t *row_ptr, void *colptr, void *rowidx, double *vals, bool narrow) {
  if (int rc = check_eval_args(g, X, M, row_ptr)) return rc;
  if (!colptr) return set_err(ASG_EINVAL, "null colptr");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (int rc = consolidate(g)) return rc;  // appended nodes join the packed layouts first
  USE_REPLICA(g, 0);
  const int64_t N = g->N;
  if (M == 0) {
    for (int64_t c = 0; c <= N; ++c)
      if (narrow) ((int32_t *)colptr)[c] = 0; else ((int64_t *)colptr)[c] = 0;
    return ASG_OK;
  }
  const int64_t nnz = row_ptr[M];
  if (nnz < 0 || nnz >= ((int64_t)1 << 31) || (nnz > 0 && (!rowidx || !vals))) return set_err(ASG_EINVAL, "bad CSC buffers");
  cudaStream_t st = dc.stream;
  BasisArgs a{};
  if (int rc = upload_points(g, X, M, sp, sdim, a.sp, a.sd, st)) return rc;
  if (int rc = R.s_i64a.reserve((size_t)M + 1)) return rc;
  if (int rc = R.s_i64b.reserve((size_t)std::max<int64_t>(nnz, 1))) return rc;
  if (int rc = R.s_tmp.reserve((size_t)std::max<int64_t>(nnz, 1))) return rc;
  CK(cudaMemcpyAsync(R.s_i64a.p, row_ptr, (size_t)(M + 1) * 8, cudaMemcpyHostToDevice, st));
  a.X = R.s_X[0].p; a.M = M; a.atol = atol; a.dropzeros = dropzeros; a.mode = 1;
  a.row_ptr = R.s_i64a.p; a.colidx = R.s_i64b.p; a.vals = R.s_tmp.p;
  CK(run_basis(g, a, st));
  // CSR -> CSC: stable sort of the entries by column (rows stay ascending inside a column); scratch shared with the
  // point ordering of the BLOCK walk (slot 1 holds the outputs)
  const size_t n1 = (size_t)std::max<int64_t>(nnz, 1);
  const size_t tb = csc_temp_bytes(std::max<int64_t>(nnz, 1));
  if (int rc = R.s_key[0].reserve(n1)) return rc;
  if (int rc = R.s_key2[0].reserve(n1)) return rc;
  if (int rc = R.s_idx[0].reserve(n1)) return rc;
  if (int rc = R.s_perm[0].reserve(n1)) return rc;
  if (int rc = R.s_key[1].reserve(n1)) return rc;
  if (int rc = R.s_sorttmp[0].reserve(tb)) return rc;
  DevBuf<long long> &d_colptr = R.c_PL, &d_rowidx = R.c_PI;  // adapt! scratch, free between calls
  if (int rc = d_colptr.reserve((size_t)N + 1)) return rc;
  if (int rc = d_rowidx.reserve(n1)) return rc;
  if (int rc = R.s_tmp2.reserve(n1)) return rc;
  if (R.sort_ev[0]) CK(cudaStreamWaitEvent(st, R.sort_ev[0], 0));
  int64_t l = 0;
  CK(csr_to_csc(R.s_i64a.p, R.s_i64b.p, R.s_tmp.p, M, N, nnz, R.s_key[0].p, R.s_key2[0].p, R.s_idx[0].p,
                R.s_perm[0].p, R.s_key[1].p, R.s_sorttmp[0].p, tb, d_colptr.p, d_rowidx.p, R.s_tmp2.p, st, &l));
  g->ch_n = -1;  // the adapt! scratch was reused
  if (narrow) {  // 32-bit indices: the CSR column scratch (free now) takes the narrowed copies
    if (int rc = R.s_idx[1].reserve((size_t)N + 1)) return rc;
    int *n_colptr = reinterpret_cast<int *>(R.s_idx[1].p);
    int *n_rowidx = reinterpret_cast<int *>(R.s_i64b.p);
    CK(launch_narrow_i64(d_colptr.p, n_colptr, N + 1, st, &l));
    CK(launch_narrow_i64(d_rowidx.p, n_rowidx, nnz, st, &l));
    bump(l);
    if (int rc = d2h_fast(dc, colptr, n_colptr, (size_t)(N + 1) * 4, st)) return rc;
    if (nnz)
      if (int rc = d2h_fast(dc, rowidx, n_rowidx, (size_t)nnz * 4, st)) return rc;
  } else {
    bump(l);
    if (int rc = d2h_fast(dc, colptr, d_colptr.p, (size_t)(N + 1) * 8, st)) return rc;
    if (nnz)
      if (int rc = d2h_fast(dc, rowidx, d_rowidx.p, (size_t)nnz * 8, st)) return rc;
  }
  if (nnz)
    if (int rc = d2h_fast(dc, vals, R.s_tmp2.p, (size_t)nnz * 8, st)) return rc;
  CK(cudaStreamSynchronize(st));
  return ASG_OK;
}

int32_t asg_basis_fill_csc(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                           int32_t dropzeros, const int64_t *row_ptr, int64_t *colptr, int64_t *rowidx, double *vals) {
  ASG_TRACE();
  return basis_fill_csc_common(g, X, M, sp, sdim, atol, dropzeros, row_ptr, colptr, rowidx, vals, false);
}

int32_t asg_basis_fill_csc32(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                             int32_t dropzeros, const int64_t *row_ptr, int32_t *colptr, int32_t *rowidx, double *vals) {
  ASG_TRACE();
  if (g && g->N >= ((int64_t)1 << 31) - 1) return set_err(ASG_EUNSUPPORTED, "N >= 2^31: use asg_basis_fill_csc");
  return basis_fill_csc_common(g, X, M, sp, sdim, atol, dropzeros, row_ptr, colptr, rowidx, vals, true);
}

int32_t asg_basis_dense_device(asg_grid *g, const double *dX, int64_t M, int64_t sp, int64_t sdim, double atol,
                               int32_t dropzeros, double *dout, int64_t ldm, int64_t ldn, void *stream) {
  ASG_TRACE();
  if (int rc = check_eval_args(g, dX, M, dout)) return rc;
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (int rc = consolidate(g)) return rc;  // appended nodes join the packed layouts first
  if (M == 0 || g->N == 0) return ASG_OK;
  if (replica_of_pointer(dX) != 0) return set_err(ASG_EINVAL, "basis entry points run on the primary device of the context");
  USE_REPLICA(g, 0);
  cudaStream_t st = (cudaStream_t)stream;  // NULL is the CUDA default stream, as everywhere in CUDA
  BasisArgs a{};
  a.X = dX; a.M = M; a.sp = sp; a.sd = sdim; a.atol = atol; a.dropzeros = dropzeros; a.mode = 2;
  a.dense = dout; a.ldm = ldm; a.ldn = ldn;
  static const bool no_split = getenv("ASG_DENSE_NO_SPLIT") != nullptr;  // experiment knob: one thread per row (round 1)
  CK(cudaEventRecord(dc.ev0, st));
  if (g->path() == ASG_PATH_SUBSPACE) {
    // the walk only visits supporting nodes: clear first (needs a dense M x N block), then scatter
    const bool rowmajor = (ldn == 1 && ldm >= g->N), colmajor = (ldm == 1 && ldn >= M);
    if (rowmajor) CK(cudaMemset2DAsync(dout, (size_t)ldm * 8, 0, (size_t)g->N * 8, (size_t)M, st));
    else if (colmajor) CK(cudaMemset2DAsync(dout, (size_t)ldn * 8, 0, (size_t)M * 8, (size_t)g->N, st));
    else return set_err(ASG_EUNSUPPORTED, "dense basis needs ldn == 1 (row-major) or ldm == 1 (column-major)");
    if (no_split) {
      CK(run_basis(g, a, st));
    } else {
      int64_t l = 0;
      CK(launch_basis_dense_split(R.dg, a, R.dg.n_level1, st, &l));
      bump(l);
    }
  } else {
    CK(run_basis(g, a, st));
  }
  CK(cudaEventRecord(dc.ev1, st));
  dc.last_timed = true;
  dc.walk_timed = false;
  ctx.last_dev = 0;
  return ASG_OK;
}

int32_t asg_basis_dense(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                        int32_t dropzeros, double *out, int64_t ldm, int64_t ldn) {
  ASG_TRACE();
  if (int rc = check_eval_args(g, X, M, out)) return rc;
  if (int rc = require_ctx()) return rc;
  const int64_t N = g->N;
  if (M == 0 || N == 0) return ASG_OK;
  const bool rowmajor = (ldn == 1 && ldm >= N), colmajor = (ldm == 1 && ldn >= M);
  if (!rowmajor && !colmajor)
    return set_err(ASG_EUNSUPPORTED, "dense basis needs ldn == 1 (row-major) or ldm == 1 (column-major)");
  // row chunks of at most ~1 GiB of device output
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(M, ((int64_t)1 << 27) / std::max<int64_t>(N, 1)));
  std::vector<double> rb;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  USE_REPLICA(g, 0);
  for (int64_t m0 = 0; m0 < M; m0 += chunk) {
    const int64_t cnt = std::min(chunk, M - m0);
    int64_t dsp, dsd;
    if (int rc = R.s_X[1].reserve((size_t)cnt * g->D)) return rc;
    if (int rc = R.s_tmp2.reserve((size_t)cnt * N)) return rc;
    if (int rc = stage_points(X, m0, cnt, g->D, sp, sdim, R.s_X[1].p, dsp, dsd, dc.stream, rb)) return rc;
    const int64_t dldm = rowmajor ? N : 1, dldn = rowmajor ? 1 : cnt;
    if (int rc = asg_basis_dense_device(g, R.s_X[1].p, cnt, dsp, dsd, atol, dropzeros, R.s_tmp2.p, dldm, dldn, dc.stream))
      return rc;
    if (rowmajor)
      CK(cudaMemcpy2DAsync(out + m0 * ldm, (size_t)ldm * 8, R.s_tmp2.p, (size_t)N * 8, (size_t)N * 8, (size_t)cnt,
                           cudaMemcpyDeviceToHost, dc.stream));
    else
      CK(cudaMemcpy2DAsync(out + m0, (size_t)ldn * 8, R.s_tmp2.p, (size_t)cnt * 8, (size_t)cnt * 8, (size_t)N,
                           cudaMemcpyDeviceToHost, dc.stream));
    CK(cudaStreamSynchronize(dc.stream));
  }
  return ASG_OK;
}

// Host-only: run the packer on a node set and report what it would build (no device needed).
// Used by the CPU test-suite to check the trie construction; performs no evaluation.
int32_t asg_pack_info(int32_t D, int64_t N, const int64_t *levels, const int64_t *indices, int64_t sn, int64_t sd,
                      asg_grid_info_t *info) {
  if (!info || D <= 0 || N < 0 || (N > 0 && (!levels || !indices))) return set_err(ASG_EINVAL, "bad argument");
  if (D > kMaxDim) return set_err(ASG_EUNSUPPORTED, "D > 64");
  std::vector<int64_t> L((size_t)N * D), I((size_t)N * D);
  for (int64_t k = 0; k < N; ++k)
    for (int d = 0; d < D; ++d) {
      L[(size_t)k * D + d] = levels[k * sn + d * sd];
      I[(size_t)k * D + d] = indices[k * sn + d * sd];
    }
  PackResult pk;
  std::string err;
  int rc = pack_grid(D, N, L.data(), I.data(), nullptr, pk, err);
  if (rc != ASG_OK) return set_err(rc, err);
  *info = asg_grid_info_t{};
  info->D = D;
  info->canonical = pk.canonical ? 1 : 0;
  info->path = pk.canonical ? ASG_PATH_SUBSPACE : ASG_PATH_SWEEP;
  info->max_level = pk.max_level;
  info->N = N;
  info->max_depth = pk.max_depth;
  info->subspaces = pk.subspaces;
  info->dense_subspaces = pk.dense_subspaces;
  info->trie_nodes = pk.trie_nodes;
  info->coef_slots = (int64_t)pk.coef.size() - pk.coef_pad;
  info->hash_slots = (int64_t)pk.hent.size();
  info->regular = pk.regular ? 1 : 0;
  info->block = (pk.canonical && pk.blk_ok) ? 1 : 0;
  info->block_slots = info->block ? (int64_t)pk.bcoef.size() - 2 : 0;
  info->block_records = info->block ? (int64_t)pk.brec.size() : 0;
  info->block_tiles = info->block ? (int64_t)pk.btiles.size() : 0;
  info->block_stack = info->block ? pk.blk_slots : 0;
  info->block_pair = (info->block && pk.blk_pair_ok && !pk.blk_wide) ? 1 : 0;
  if (!pk.canonical) t_err = pk.why;
  else if (!pk.blk_ok) t_err = "no BLOCK layout: " + pk.blk_why;
  return ASG_OK;
}

// ---- device-format persistence -------------------------------------------------
int32_t asg_grid_export(asg_grid *g, void *buf, int64_t capacity, int64_t *bytes) {
  ASG_TRACE();
  if (!g || !bytes) return set_err(ASG_EINVAL, "null argument");
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (!g->uploaded) return set_err(ASG_ESTATE, "no nodes uploaded");
  if (ctx.ready && g->N_packed < g->N) {
    if (int rc = require_ctx()) return rc;
    if (int rc = consolidate(g)) return rc;
  }
  if (g->blob_generation != g->info.generation || g->blob.empty()) {
    PackResult pk;
    std::string err;
    int rc = pack_grid(g->D, g->N, g->levels.data(), g->indices.data(), g->coefs.data(), pk, err);
    if (rc != ASG_OK) return set_err(rc, err);
    pack_serialize(g->D, g->lb, g->ub, g->levels, g->indices, g->coefs, pk, g->blob);
    g->blob_generation = g->info.generation;
  }
  *bytes = (int64_t)g->blob.size();
  if (!buf) return ASG_OK;  // size query
  if (capacity < *bytes) return set_err(ASG_EINVAL, "buffer too small for the packed grid");
  std::memcpy(buf, g->blob.data(), g->blob.size());
  std::vector<unsigned char>().swap(g->blob);  // the blob can be large: do not keep a second copy around
  return ASG_OK;
}

static int import_common(asg_grid *g, const void *buf, int64_t bytes, int64_t N, const int64_t *levels,
                         const int64_t *indices, int64_t sn, int64_t sd, const double *coefs, bool check) {
  PackResult pk;
  std::vector<int64_t> L, I;
  std::vector<double> c;
  std::string err;
  int rc = pack_deserialize((const unsigned char *)buf, (size_t)bytes, g->D, g->lb, g->ub, L, I, c, pk, err);
  if (rc != ASG_OK) return set_err(rc, err);
  if (check) {
    // the blob is a cache of THESE arrays: a blob written before an in-place coefficient update, or one of another
    // grid with the same D / domain / node count, must not be installed
    const int D = g->D;
    bool same = (int64_t)c.size() == N;
    for (int64_t k = 0; same && k < N; ++k) {
      for (int d = 0; same && d < D; ++d)
        same = L[(size_t)k * D + d] == levels[k * sn + d * sd] && I[(size_t)k * D + d] == indices[k * sn + d * sd];
      if (same && coefs) same = std::memcmp(&c[(size_t)k], &coefs[k], 8) == 0;
    }
    if (!same) return set_err(ASG_ESTATE, "packed blob does not describe these node arrays (stale or foreign blob)");
  }
  g->uploaded = false;
  g->levels.swap(L);
  g->indices.swap(I);
  g->coefs.swap(c);
  g->N = (int64_t)g->coefs.size();
  return install_pack(g, pk);
}

int32_t asg_grid_import(asg_grid *g, const void *buf, int64_t bytes) {
  ASG_TRACE();
  if (!g || !buf || bytes <= 0) return set_err(ASG_EINVAL, "null argument");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  return import_common(g, buf, bytes, 0, nullptr, nullptr, 0, 0, nullptr, false);
}

int32_t asg_grid_import_checked(asg_grid *g, const void *buf, int64_t bytes, int64_t N, const int64_t *levels,
                                const int64_t *indices, int64_t sn, int64_t sd, const double *coefs) {
  ASG_TRACE();
  if (!g || !buf || bytes <= 0) return set_err(ASG_EINVAL, "null argument");
  if (N < 0 || (N > 0 && (!levels || !indices))) return set_err(ASG_EINVAL, "bad node arrays");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  return import_common(g, buf, bytes, N, levels, indices, sn, sd, coefs, true);
}

int32_t asg_rsg_count(int32_t D, int64_t maxdepth, const int64_t *maxlevels, int64_t *N) {
  if (D <= 0 || D > kMaxDim || !maxlevels || !N) return set_err(ASG_EINVAL, "bad argument");
  std::string err;
  const int rc = rsg_count(D, maxdepth, maxlevels, N, err);
  return rc == ASG_OK ? ASG_OK : set_err(rc, err);
}

int32_t asg_rsg_fill(int32_t D, int64_t maxdepth, const int64_t *maxlevels, int64_t N, int64_t *levels, int64_t *indices,
                     int64_t sn, int64_t sd) {
  if (D <= 0 || D > kMaxDim || !maxlevels || N < 0 || (N > 0 && (!levels || !indices))) return set_err(ASG_EINVAL, "bad argument");
  std::string err;
  const int rc = rsg_fill(D, maxdepth, maxlevels, N, levels, indices, sn, sd, err);
  return rc == ASG_OK ? ASG_OK : set_err(rc, err);
}

int32_t asg_measure_fp64_peak(double *fma_per_second) {
  DevCtx &dc = ctx.dev[0];
  if (!fma_per_second) return set_err(ASG_EINVAL, "null pointer");
  if (int rc = require_ctx()) return rc;
  double *sink = nullptr;
  CK(cudaMalloc(&sink, 8));
  const int iters = 1 << 16;
  int blocks = 0, threads = 0;
  int64_t l = 0;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(dc.ev0, dc.stream));
    CK(launch_fp64_peak(iters, sink, dc.stream, &l, &blocks, &threads));
    CK(cudaEventRecord(dc.ev1, dc.stream));
    CK(cudaEventSynchronize(dc.ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, dc.ev0, dc.ev1));
    const double rate = (double)blocks * threads * 8.0 * iters / (ms * 1e-3);
    if (rep > 0) best = std::max(best, rate);
  }
  bump(l);
  cudaFree(sink);
  *fma_per_second = best;
  return ASG_OK;
}

}  // extern "C"
