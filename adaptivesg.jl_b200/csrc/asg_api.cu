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

struct Context {
  std::mutex mu;
  bool ready = false;
  int device = -1;
  cudaStream_t stream = nullptr;      // library stream (device-pointer entry points)
  cudaStream_t pipe[2] = {nullptr, nullptr};  // host-buffer pipelines (double buffering)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // around the last *_device call
  cudaEvent_t evw0 = nullptr, evw1 = nullptr;  // around its walk kernel alone (BLOCK walk: after the point ordering)
  bool walk_timed = false;
  std::atomic<long long> launches{0};
  double last_ms = 0.0;
  bool last_timed = false;
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

int require_ctx() {
  if (!ctx.ready) return set_err(ASG_ENODEVICE, "asg_init has not succeeded: no sm_100 device selected (no CPU fallback)");
  cudaError_t e = cudaSetDevice(ctx.device);
  if (e != cudaSuccess) return set_err(ASG_ECUDA, cudaGetErrorString(e));
  return ASG_OK;
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
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

void bump(int64_t n) { ctx.launches += n; }

}  // namespace

struct asg_grid {
  std::recursive_mutex mu;
  int D = 0;
  double lb[kMaxDim], ub[kMaxDim];
  // host copy of the caller's arrays (row-major N x D): the caller's struct stays the source of truth,
  // this copy exists so append / set_coefs / surplus_nodes need no second pass over caller memory
  int64_t N = 0;
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
  DeviceGrid dg{};
  // device storage
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
  // scratch for host-buffer calls (two pipeline slots)
  DevBuf<double> s_X[2], s_out[2], s_f[2];
  DevBuf<long long> s_i64a, s_i64b;
  DevBuf<double> s_tmp, s_tmp2;
  DevBuf<int> s_flag;
  // adapt! candidate generation (asg_kernels_adapt.cu): the unique children of the last asg_children call
  DevBuf<long long> c_PL, c_PI, c_CL, c_CI, c_OL, c_OI;
  DevBuf<unsigned char> c_valid;
  DevBuf<int> c_table;
  DevBuf<unsigned int> c_flags, c_sums;
  int64_t ch_n = -1, ch_generation = -1;
  std::vector<unsigned char> blob;  // asg_grid_export: between the size query and the copy
  int64_t blob_generation = -1;
  // point-order scratch of the BLOCK walk (asg_sort.cu), one set per pipeline slot
  DevBuf<uint32_t> s_key[2], s_key2[2], s_idx[2], s_perm[2];
  DevBuf<unsigned char> s_sorttmp[2];
  cudaEvent_t sort_ev[2] = {nullptr, nullptr};  // last use of the slot's scratch
  int path() const {
    if (forced_path == ASG_PATH_SWEEP) return ASG_PATH_SWEEP;
    return (canonical && finite) ? ASG_PATH_SUBSPACE : ASG_PATH_SWEEP;
  }
  // which walk of the subspace engine: 3 = BLOCK walk (default where the layout exists), 0 = flat group
  // stream (other canonical grids, ASG_EVAL_KERNEL=stream), 1 = record-free regular walk (=rsg),
  // 2 = recursive trie walk (ASG_PATH_TRIE, D == 1, or ASG_EVAL_KERNEL=trie)
  int walk() const {
    static const char *env = getenv("ASG_EVAL_KERNEL");
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

struct asg_event {
  std::thread th;
  int rc = ASG_OK;
  std::string err;
};

namespace {

int device_upload(asg_grid *g, const PackResult &pk) {
  const int D = g->D;
  const int64_t N = g->N;
  DeviceGrid &dg = g->dg;
  dg = DeviceGrid{};
  dg.D = D;
  for (int d = 0; d < D; ++d) {
    dg.lb[d] = g->lb[d];
    volatile double w = g->ub[d] - g->lb[d];  // (from_ub .- from_lb), src/math.jl:308
    dg.width[d] = w;
  }
  dg.N = N;
  cudaStream_t st = ctx.stream;
  // sweep engine arrays (always present: they serve every node set)
  if (int rc = g->d_sw.reserve((size_t)N * D)) return rc;
  if (int rc = g->d_sw_low.reserve((size_t)N)) return rc;
  if (int rc = g->d_sw_coef.reserve((size_t)N)) return rc;
  if (int rc = g->d_sw_depth.reserve((size_t)N)) return rc;
  if (N) {
    CK(cudaMemcpyAsync(g->d_sw.p, pk.sw.data(), (size_t)N * D * sizeof(SweepDim), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g->d_sw_low.p, pk.sw_low.data(), (size_t)N * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g->d_sw_coef.p, g->coefs.data(), (size_t)N * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g->d_sw_depth.p, pk.depth.data(), (size_t)N * 4, cudaMemcpyHostToDevice, st));
  }
  dg.sw = g->d_sw.p;
  dg.sw_low = g->d_sw_low.p;
  dg.sw_coef = g->d_sw_coef.p;
  dg.sw_depth = g->d_sw_depth.p;
  // subspace engine arrays
  if (pk.canonical) {
    size_t total = 0;
    for (int k = 0; k < D; ++k) total += pk.rec[(size_t)k].size();
    if (int rc = g->d_rec.reserve(total)) return rc;
    size_t off = 0;
    for (int k = 0; k < D; ++k) {
      const auto &r = pk.rec[(size_t)k];
      CK(cudaMemcpyAsync(g->d_rec.p + off, r.data(), r.size() * sizeof(TrieRec), cudaMemcpyHostToDevice, st));
      dg.rec[k] = g->d_rec.p + off;
      off += r.size();
    }
    dg.n_root = (uint32_t)(pk.rec[0].size() - 1);
    if (int rc = g->d_coef.reserve(pk.coef.size())) return rc;
    if (int rc = g->d_orig.reserve(pk.orig.size())) return rc;
    if (int rc = g->d_hent.reserve(pk.hent.size())) return rc;
    if (int rc = g->d_horig.reserve(pk.horig.size())) return rc;
    if (!pk.coef.empty()) {
      CK(cudaMemcpyAsync(g->d_coef.p, pk.coef.data(), pk.coef.size() * 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(g->d_orig.p, pk.orig.data(), pk.orig.size() * 4, cudaMemcpyHostToDevice, st));
    }
    if (!pk.hent.empty()) {
      CK(cudaMemcpyAsync(g->d_hent.p, pk.hent.data(), pk.hent.size() * sizeof(HashEnt), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(g->d_horig.p, pk.horig.data(), pk.horig.size() * 4, cudaMemcpyHostToDevice, st));
    }
    if (int rc = g->d_grp.reserve(pk.grp.size())) return rc;
    if (int rc = g->d_tiles.reserve(pk.tiles.size())) return rc;
    if (!pk.grp.empty()) {
      CK(cudaMemcpyAsync(g->d_grp.p, pk.grp.data(), pk.grp.size() * sizeof(GroupRec), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(g->d_tiles.p, pk.tiles.data(), pk.tiles.size() * sizeof(StreamTile), cudaMemcpyHostToDevice, st));
    }
    dg.grp = g->d_grp.p;
    dg.tiles = g->d_tiles.p;
    dg.n_grp = (uint32_t)pk.grp.size();
    dg.n_tiles = (uint32_t)pk.tiles.size();
    dg.coef = g->d_coef.p;
    dg.orig = g->d_orig.p;
    dg.hent = g->d_hent.p;
    dg.horig = g->d_horig.p;
    if (pk.blk_ok) {
      if (int rc = g->d_bcoef.reserve(pk.bcoef.size())) return rc;
      if (int rc = g->d_brec.reserve(pk.brec.size())) return rc;
      if (int rc = g->d_btiles.reserve(pk.btiles.size())) return rc;
      CK(cudaMemcpyAsync(g->d_bcoef.p, pk.bcoef.data(), pk.bcoef.size() * 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(g->d_brec.p, pk.brec.data(), pk.brec.size() * sizeof(BlockRec), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(g->d_btiles.p, pk.btiles.data(), pk.btiles.size() * sizeof(BlockTile), cudaMemcpyHostToDevice, st));
      dg.bcoef = g->d_bcoef.p;
      dg.brec = g->d_brec.p;
      dg.btiles = g->d_btiles.p;
      dg.n_brec = (uint32_t)pk.brec.size();
      dg.n_btiles = (uint32_t)pk.btiles.size();
      dg.blk_slots = pk.blk_slots;
      dg.blk_tile_coefs = pk.blk_tile_coefs;
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
  rc = device_upload(g, pk);
  if (rc != ASG_OK) return rc;
  g->slot_of_node.swap(pk.slot_of_node);
  g->bslot_of_node.swap(pk.bslot_of_node);
  g->blk_ok = pk.canonical && pk.blk_ok;
  g->blk_why = pk.blk_why;
  g->uploaded = true;
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
  I.device_bytes = (int64_t)(g->d_rec.bytes() + g->d_grp.bytes() + g->d_tiles.bytes() + g->d_coef.bytes() + g->d_orig.bytes() + g->d_hent.bytes() +
                             g->d_horig.bytes() + g->d_sw.bytes() + g->d_sw_low.bytes() + g->d_sw_coef.bytes() +
                             g->d_sw_depth.bytes());
  I.generation = gen;
  I.regular = pk.regular ? 1 : 0;
  I.block = g->blk_ok ? 1 : 0;
  I.block_slots = g->blk_ok ? (int64_t)pk.bcoef.size() - 2 : 0;
  I.block_records = g->blk_ok ? (int64_t)pk.brec.size() : 0;
  I.block_tiles = g->blk_ok ? (int64_t)pk.btiles.size() : 0;
  I.block_stack = g->blk_ok ? pk.blk_slots : 0;
  if (g->blk_ok) I.device_bytes += (int64_t)(g->d_bcoef.bytes() + g->d_brec.bytes() + g->d_btiles.bytes());
  return ASG_OK;
}

int copy_nodes(asg_grid *g, int64_t at, int64_t n, const int64_t *levels, const int64_t *indices, int64_t sn,
               int64_t sd, const double *coefs) {
  const int D = g->D;
  g->levels.resize((size_t)(at + n) * D);
  g->indices.resize((size_t)(at + n) * D);
  g->coefs.resize((size_t)(at + n));
  for (int64_t k = 0; k < n; ++k) {
    for (int d = 0; d < D; ++d) {
      g->levels[(size_t)(at + k) * D + d] = levels[k * sn + d * sd];
      g->indices[(size_t)(at + k) * D + d] = indices[k * sn + d * sd];
    }
    g->coefs[(size_t)(at + k)] = coefs ? coefs[k] : 0.0;
  }
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

int run_eval(asg_grid *g, const EvalArgs &a_in, cudaStream_t st, int slot = 0, bool timed = false) {
  int64_t l = 0;
  if (timed) ctx.walk_timed = false;
  cudaError_t e;
  EvalArgs a = a_in;
  if (g->path() != ASG_PATH_SUBSPACE) e = launch_eval_sweep(g->dg, a, st, &l);
  else if (g->walk() == 3) {
    bool sorted = false;
    // (grids of a few hundred subspaces walk in less time than the ordering's five launches take)
    if (a.M >= sort_threshold() && a.M < ((int64_t)1 << 31) && g->info.subspaces >= 512) {
      const size_t tb = sort_points_temp_bytes(a.M);
      if (int rc = g->s_key[slot].reserve((size_t)a.M)) return rc;
      if (int rc = g->s_key2[slot].reserve((size_t)a.M)) return rc;
      if (int rc = g->s_idx[slot].reserve((size_t)a.M)) return rc;
      if (int rc = g->s_perm[slot].reserve((size_t)a.M)) return rc;
      if (int rc = g->s_sorttmp[slot].reserve(tb)) return rc;
      if (!g->sort_ev[slot]) CK(cudaEventCreateWithFlags(&g->sort_ev[slot], cudaEventDisableTiming));
      else CK(cudaStreamWaitEvent(st, g->sort_ev[slot], 0));  // an earlier launch on another stream may still read the scratch
      CK(sort_points(g->dg, a.X, a.M, a.sp, a.sd, g->s_key[slot].p, g->s_key2[slot].p, g->s_idx[slot].p, g->s_perm[slot].p,
                     g->s_sorttmp[slot].p, tb, st, &l));
      a.perm = g->s_perm[slot].p;
      sorted = true;
    }
    if (timed) CK(cudaEventRecord(ctx.evw0, st));
    e = launch_eval_block(g->dg, a, st, &l);
    if (timed) { CK(cudaEventRecord(ctx.evw1, st)); ctx.walk_timed = true; }
    if (sorted && e == cudaSuccess) CK(cudaEventRecord(g->sort_ev[slot], st));
  }
  else if (g->walk() == 0) e = launch_eval_stream(g->dg, a, st, &l);
  else if (g->walk() == 1) e = launch_eval_rsg(g->dg, g->rp, a, st, &l);
  else e = launch_eval_subspace(g->dg, a, st, &l);
  bump(l);
  CK(e);
  return ASG_OK;
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

// host-buffer evaluation pipeline: chunks alternate between two streams so that the H2D copy of
// chunk k+1 and the D2H copy of chunk k-1 overlap the kernel of chunk k
int eval_host(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, const double *fvals,
              int64_t max_depth, double *out) {
  const int D = g->D;
  // chunk sizes ramp up 1 Mi, 2 Mi, 4 Mi, 8 Mi, 8 Mi, ...: the first copy (the only one no kernel hides) is short,
  // the later chunks are large enough for the cell-key ordering of the BLOCK walk to bite
  const int64_t first = std::min<int64_t>(M, (int64_t)1 << 20);
  const int64_t chunk = std::min<int64_t>(M, (int64_t)1 << 23);  // buffer size per pipeline slot
  std::vector<double> repack_buf;
  const int nslot = M > first ? 2 : 1;
  for (int s = 0; s < nslot; ++s) {
    if (int rc = g->s_X[s].reserve((size_t)chunk * D)) return rc;
    if (int rc = g->s_out[s].reserve((size_t)chunk)) return rc;
    if (fvals)
      if (int rc = g->s_f[s].reserve((size_t)chunk)) return rc;
  }
  int slot = 0;
  int64_t next = first;
  for (int64_t m0 = 0, cnt = 0; m0 < M; m0 += cnt, slot ^= (nslot - 1)) {
    cnt = std::min(next, M - m0);
    next = std::min(2 * next, chunk);
    cudaStream_t st = ctx.pipe[slot];
    EvalArgs a{};
    if (int rc = stage_points(X, m0, cnt, D, sp, sdim, g->s_X[slot].p, a.sp, a.sd, st, repack_buf)) return rc;
    a.X = g->s_X[slot].p;
    a.M = cnt;
    a.rem = rem_of(max_depth);
    a.fvals = nullptr;
    if (fvals) {
      CK(cudaMemcpyAsync(g->s_f[slot].p, fvals + m0, (size_t)cnt * 8, cudaMemcpyHostToDevice, st));
      a.fvals = g->s_f[slot].p;
    }
    a.out = g->s_out[slot].p;
    if (int rc = run_eval(g, a, st, slot)) return rc;
    CK(cudaMemcpyAsync(out + m0, g->s_out[slot].p, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(ctx.pipe[0]));
  if (nslot > 1) CK(cudaStreamSynchronize(ctx.pipe[1]));
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

int32_t asg_init(int32_t device) {
  std::lock_guard<std::mutex> lk(ctx.mu);
  if (ctx.ready && ctx.device == device) return ASG_OK;
  if (ctx.ready) return set_err(ASG_ESTATE, "already initialised on another device (one process per GPU)");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return set_err(ASG_ENODEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0") +
                                      " (libasg_cuda has no CPU fallback)");
  if (device < 0 || device >= n) return set_err(ASG_ENODEVICE, "device index out of range");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_err(ASG_ENODEVICE, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                                      ", this library is built for sm_100a only");
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&ctx.pipe[0], cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&ctx.pipe[1], cudaStreamNonBlocking));
  CK(cudaEventCreate(&ctx.ev0));
  CK(cudaEventCreate(&ctx.ev1));
  CK(cudaEventCreate(&ctx.evw0));
  CK(cudaEventCreate(&ctx.evw1));
  ctx.device = device;
  ctx.ready = true;
  return ASG_OK;
}

int32_t asg_shutdown(void) {
  std::lock_guard<std::mutex> lk(ctx.mu);
  if (!ctx.ready) return ASG_OK;
  cudaSetDevice(ctx.device);
  cudaDeviceSynchronize();
  cudaStreamDestroy(ctx.stream);
  cudaStreamDestroy(ctx.pipe[0]);
  cudaStreamDestroy(ctx.pipe[1]);
  cudaEventDestroy(ctx.ev0);
  cudaEventDestroy(ctx.ev1);
  cudaEventDestroy(ctx.evw0);
  cudaEventDestroy(ctx.evw1);
  ctx.ready = false;
  ctx.device = -1;
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
  CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
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
      cudaSetDevice(ctx.device);
      g->d_rec.release(); g->d_grp.release(); g->d_tiles.release(); g->d_bcoef.release(); g->d_brec.release(); g->d_btiles.release(); g->d_coef.release(); g->d_sw_coef.release(); g->d_orig.release(); g->d_horig.release();
      g->d_sw_depth.release(); g->d_hent.release(); g->d_sw.release(); g->d_sw_low.release();
      for (int s = 0; s < 2; ++s) { g->s_X[s].release(); g->s_out[s].release(); g->s_f[s].release(); }
      for (int s = 0; s < 2; ++s) {
        g->s_key[s].release(); g->s_key2[s].release(); g->s_idx[s].release(); g->s_perm[s].release(); g->s_sorttmp[s].release();
        if (g->sort_ev[s]) cudaEventDestroy(g->sort_ev[s]);
      }
      g->c_PL.release(); g->c_PI.release(); g->c_CL.release(); g->c_CI.release(); g->c_OL.release(); g->c_OI.release();
      g->c_valid.release(); g->c_table.release(); g->c_flags.release(); g->c_sums.release();
      g->s_i64a.release(); g->s_i64b.release(); g->s_tmp.release(); g->s_tmp2.release(); g->s_flag.release();
    }
  }
  delete g;
  return ASG_OK;
}

int32_t asg_grid_upload(asg_grid *g, int64_t N, const int64_t *levels, const int64_t *indices, int64_t sn,
                        int64_t sd, const double *coefs) {
  if (!g) return set_err(ASG_EINVAL, "null grid handle");
  if (N < 0 || (N > 0 && (!levels || !indices))) return set_err(ASG_EINVAL, "bad node arrays");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  g->uploaded = false;
  g->levels.clear(); g->indices.clear(); g->coefs.clear();
  copy_nodes(g, 0, N, levels, indices, sn, sd, coefs);
  return repack(g);
}

int32_t asg_grid_append(asg_grid *g, int64_t n, const int64_t *levels, const int64_t *indices, int64_t sn,
                        int64_t sd, const double *coefs) {
  if (!g) return set_err(ASG_EINVAL, "null grid handle");
  if (n < 0 || (n > 0 && (!levels || !indices))) return set_err(ASG_EINVAL, "bad node arrays");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  const int64_t N0 = g->N;
  copy_nodes(g, N0, n, levels, indices, sn, sd, coefs);
  g->uploaded = false;
  int rc = repack(g);
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
  // idx == nullptr: nodes first .. first+n-1
  if (n == 0) return ASG_OK;
  cudaStream_t st = ctx.stream;
  std::vector<long long> slots((size_t)n), nodes((size_t)n);
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
      slots[(size_t)k] = g->slot_of_node[(size_t)node];
      (slots[(size_t)k] < 0 ? any_hash : any_dense) = true;
    }
  }
  if (int rc = g->s_i64a.reserve((size_t)n)) return rc;
  const double *dsrc = dcoefs_or_null;
  if (!dsrc) {
    if (int rc = g->s_tmp2.reserve((size_t)n)) return rc;
    CK(cudaMemcpyAsync(g->s_tmp2.p, coefs, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    dsrc = g->s_tmp2.p;
  }
  int64_t l = 0;
  CK(cudaMemcpyAsync(g->s_i64a.p, nodes.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CK(launch_scatter_f64(g->d_sw_coef.p, g->s_i64a.p, dsrc, n, st, &l));
  if (g->canonical) {
    if (int rc = g->s_i64b.reserve((size_t)n)) return rc;
    CK(cudaMemcpyAsync(g->s_i64b.p, slots.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    if (any_dense) CK(launch_scatter_f64(g->d_coef.p, g->s_i64b.p, dsrc, n, st, &l));
    if (any_hash) CK(launch_scatter_hash(g->d_hent.p, g->s_i64b.p, dsrc, n, st, &l));
    if (g->blk_ok) {  // the BLOCK layout holds its own copy of the coefficients
      for (int64_t k = 0; k < n; ++k) slots[(size_t)k] = g->bslot_of_node[(size_t)nodes[(size_t)k]];
      CK(cudaStreamSynchronize(st));  // s_i64b is about to be overwritten
      CK(cudaMemcpyAsync(g->s_i64b.p, slots.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
      CK(launch_scatter_f64(g->d_bcoef.p, g->s_i64b.p, dsrc, n, st, &l));
    }
  }
  bump(l);
  CK(cudaStreamSynchronize(st));
  g->info.generation++;
  return ASG_OK;
}

int32_t asg_grid_set_coefs(asg_grid *g, int64_t first, int64_t n, const double *coefs) {
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
  if (!g->uploaded) return set_err(ASG_ESTATE, "no nodes uploaded");
  if (first < 0 || n < 0 || first + n > g->N) return set_err(ASG_EINVAL, "range exceeds node count");
  // read the DEVICE copy (caller's node order) so that this is a true read-back of what the kernels see
  if (n) CK(cudaMemcpy(coefs, g->d_sw_coef.p + first, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return ASG_OK;
}

int32_t asg_grid_info(asg_grid *g, asg_grid_info_t *info) {
  if (!g || !info) return set_err(ASG_EINVAL, "null argument");
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  *info = g->info;
  info->D = g->D;
  info->N = g->N;
  info->path = g->uploaded ? g->path() : 0;
  return ASG_OK;
}

int32_t asg_grid_set_path(asg_grid *g, int32_t path) {
  if (!g) return set_err(ASG_EINVAL, "null grid handle");
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (path < ASG_PATH_AUTO || path > ASG_PATH_BLOCK) return set_err(ASG_EINVAL, "bad path");
  if (path >= ASG_PATH_SUBSPACE && g->uploaded && g->N > 0 && !(g->canonical && g->finite))
    return set_err(ASG_EUNSUPPORTED, "grid is not canonical (" + g->why + "): only the sweep engine applies");
  if (path == ASG_PATH_BLOCK && g->uploaded && g->N > 0 && !(g->blk_ok && block_kernel_fits(g->dg)))
    return set_err(ASG_EUNSUPPORTED, "grid has no BLOCK layout (" + g->blk_why + ")");
  g->forced_path = path;
  return ASG_OK;
}

int32_t asg_eval(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, int64_t max_depth, double *out) {
  if (int rc = check_eval_args(g, X, M, out)) return rc;
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  return eval_host(g, X, M, sp, sdim, nullptr, max_depth, out);
}

int32_t asg_surplus(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, const double *fvals,
                    int64_t max_depth, double *alpha) {
  if (int rc = check_eval_args(g, X, M, alpha)) return rc;
  if (M > 0 && !fvals) return set_err(ASG_EINVAL, "null fvals");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  return eval_host(g, X, M, sp, sdim, fvals, max_depth, alpha);
}

int32_t asg_eval_device(asg_grid *g, const double *dX, int64_t M, int64_t sp, int64_t sdim, int64_t max_depth,
                        double *dout, void *stream) {
  if (int rc = check_eval_args(g, dX, M, dout)) return rc;
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  cudaStream_t st = (cudaStream_t)stream;  // NULL is the CUDA default stream, as everywhere in CUDA
  EvalArgs a{dX, M, sp, sdim, rem_of(max_depth), nullptr, dout};
  CK(cudaEventRecord(ctx.ev0, st));
  if (int rc = run_eval(g, a, st, 0, true)) return rc;
  CK(cudaEventRecord(ctx.ev1, st));
  ctx.last_timed = true;
  return ASG_OK;
}

int32_t asg_surplus_device(asg_grid *g, const double *dX, int64_t M, int64_t sp, int64_t sdim, const double *dfvals,
                           int64_t max_depth, double *dalpha, void *stream) {
  if (int rc = check_eval_args(g, dX, M, dalpha)) return rc;
  if (M > 0 && !dfvals) return set_err(ASG_EINVAL, "null fvals");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  cudaStream_t st = (cudaStream_t)stream;  // NULL is the CUDA default stream, as everywhere in CUDA
  EvalArgs a{dX, M, sp, sdim, rem_of(max_depth), dfvals, dalpha};
  CK(cudaEventRecord(ctx.ev0, st));
  if (int rc = run_eval(g, a, st, 0, true)) return rc;
  CK(cudaEventRecord(ctx.ev1, st));
  ctx.last_timed = true;
  return ASG_OK;
}

int32_t asg_last_kernel_ms(double *ms) {
  if (!ms) return set_err(ASG_EINVAL, "null pointer");
  if (int rc = require_ctx()) return rc;
  if (!ctx.last_timed) return set_err(ASG_ESTATE, "no timed launch yet");
  CK(cudaEventSynchronize(ctx.ev1));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, ctx.ev0, ctx.ev1));
  *ms = (double)f;
  return ASG_OK;
}

int32_t asg_last_walk_ms(double *ms) {
  if (!ms) return set_err(ASG_EINVAL, "null pointer");
  if (int rc = require_ctx()) return rc;
  if (!ctx.last_timed) return set_err(ASG_ESTATE, "no timed launch yet");
  cudaEvent_t a = ctx.walk_timed ? ctx.evw0 : ctx.ev0, b = ctx.walk_timed ? ctx.evw1 : ctx.ev1;
  CK(cudaEventSynchronize(b));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, a, b));
  *ms = (double)f;
  return ASG_OK;
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

static int nodes_x_device(asg_grid *g, int64_t n, const long long *h_lev, const long long *h_idx, double *dX,
                          int64_t sp, int64_t sd, cudaStream_t st) {
  const int D = g->D;
  if (int rc = g->s_i64a.reserve((size_t)n * D)) return rc;
  if (int rc = g->s_i64b.reserve((size_t)n * D)) return rc;
  if (int rc = g->s_flag.reserve(1)) return rc;
  CK(cudaMemsetAsync(g->s_flag.p, 0, sizeof(int), st));
  CK(cudaMemcpyAsync(g->s_i64a.p, h_lev, (size_t)n * D * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g->s_i64b.p, h_idx, (size_t)n * D * 8, cudaMemcpyHostToDevice, st));
  int64_t l = 0;
  // the kernel only reads lb / width of the DeviceGrid, valid even before any upload
  DeviceGrid dg = g->dg;
  dg.D = D;
  for (int d = 0; d < D; ++d) {
    dg.lb[d] = g->lb[d];
    volatile double w = g->ub[d] - g->lb[d];
    dg.width[d] = w;
  }
  CK(launch_nodes_x(D, nullptr, dg, g->s_i64a.p, g->s_i64b.p, n, dX, sp, sd, g->s_flag.p, st, &l));
  bump(l);
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, g->s_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
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
  if (!g || (n > 0 && !Xout)) return set_err(ASG_EINVAL, "null argument");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (first < 0 || n < 0 || first + n > g->N) return set_err(ASG_EINVAL, "range exceeds node count");
  if (n == 0) return ASG_OK;
  const int D = g->D;
  if (int rc = g->s_tmp.reserve((size_t)n * D)) return rc;
  if (int rc = nodes_x_device(g, n, (const long long *)g->levels.data() + first * D,
                              (const long long *)g->indices.data() + first * D, g->s_tmp.p, D, 1, ctx.stream))
    return rc;
  return copy_out_matrix(g->s_tmp.p, n, D, Xout, sp, sdim, ctx.stream);
}

int32_t asg_nodes_x_li(asg_grid *g, int64_t n, const int64_t *levels, const int64_t *indices, int64_t sn, int64_t sd,
                       double *Xout, int64_t sp, int64_t sdim) {
  if (!g || (n > 0 && (!Xout || !levels || !indices))) return set_err(ASG_EINVAL, "null argument");
  if (n < 0) return set_err(ASG_EINVAL, "negative count");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (n == 0) return ASG_OK;
  const int D = g->D;
  std::vector<long long> L((size_t)n * D), I((size_t)n * D);
  for (int64_t k = 0; k < n; ++k)
    for (int d = 0; d < D; ++d) {
      L[(size_t)k * D + d] = levels[k * sn + d * sd];
      I[(size_t)k * D + d] = indices[k * sn + d * sd];
    }
  if (int rc = g->s_tmp.reserve((size_t)n * D)) return rc;
  if (int rc = nodes_x_device(g, n, L.data(), I.data(), g->s_tmp.p, D, 1, ctx.stream)) return rc;
  return copy_out_matrix(g->s_tmp.p, n, D, Xout, sp, sdim, ctx.stream);
}

int32_t asg_surplus_nodes(asg_grid *g, int64_t n, const int64_t *idx, const double *fvals, int64_t max_depth,
                          int32_t store, double *alpha) {
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
  cudaStream_t st = ctx.stream;
  if (int rc = g->s_tmp.reserve((size_t)n * D)) return rc;
  if (int rc = nodes_x_device(g, n, L.data(), I.data(), g->s_tmp.p, D, 1, st)) return rc;  // get_x on device
  if (int rc = g->s_f[0].reserve((size_t)n)) return rc;
  if (int rc = g->s_out[0].reserve((size_t)n)) return rc;
  CK(cudaMemcpyAsync(g->s_f[0].p, fvals, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  EvalArgs a{g->s_tmp.p, n, D, 1, rem_of(max_depth), g->s_f[0].p, g->s_out[0].p};
  if (int rc = run_eval(g, a, st)) return rc;
  CK(cudaMemcpyAsync(alpha, g->s_out[0].p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (store) {
    // alpha is now on the host too: keep the host mirror, the finite flag and the device slots in step
    return scatter_coefs_locked(g, n, idx, 0, alpha, g->s_out[0].p);
  }
  return ASG_OK;
}

// ---- adapt! candidate generation ----------------------------------------------------------------
int32_t asg_children(asg_grid *g, int64_t depth, int64_t *n_parents, int64_t *n_candidates, int64_t *n_unique) {
  if (!g || !n_unique) return set_err(ASG_EINVAL, "null argument");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
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
  if (int rc = g->c_PL.reserve((size_t)n_p * D)) return rc;
  if (int rc = g->c_PI.reserve((size_t)n_p * D)) return rc;
  if (int rc = g->c_CL.reserve((size_t)n_cand * D)) return rc;
  if (int rc = g->c_CI.reserve((size_t)n_cand * D)) return rc;
  if (int rc = g->c_OL.reserve((size_t)n_cand * D)) return rc;
  if (int rc = g->c_OI.reserve((size_t)n_cand * D)) return rc;
  if (int rc = g->c_valid.reserve((size_t)n_cand)) return rc;
  if (int rc = g->c_table.reserve((size_t)tsize)) return rc;
  if (int rc = g->c_flags.reserve((size_t)n_cand)) return rc;
  if (int rc = g->c_sums.reserve((size_t)nblk + 1)) return rc;
  cudaStream_t st = ctx.stream;
  CK(cudaMemcpyAsync(g->c_PL.p, PL.data(), PL.size() * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g->c_PI.p, PI.data(), PI.size() * 8, cudaMemcpyHostToDevice, st));
  int64_t l = 0;
  unsigned int *total = g->c_sums.p + nblk;
  CK(launch_children(D, n_p, g->c_PL.p, g->c_PI.p, g->c_CL.p, g->c_CI.p, g->c_valid.p, g->c_table.p, tsize, g->c_flags.p,
                     g->c_sums.p, total, g->c_OL.p, g->c_OI.p, st, &l));
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
  if (!g || (n > 0 && (!levels || !indices))) return set_err(ASG_EINVAL, "null argument");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (g->ch_n < 0 || g->ch_generation != g->info.generation)
    return set_err(ASG_ESTATE, "asg_children has not been called on this state of the grid");
  if (n != g->ch_n) return set_err(ASG_EINVAL, "n differs from the count asg_children returned");
  if (n == 0) return ASG_OK;
  const int D = g->D;
  cudaStream_t st = ctx.stream;
  std::vector<long long> L((size_t)n * D), I((size_t)n * D);
  CK(cudaMemcpyAsync(L.data(), g->c_OL.p, (size_t)n * D * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(I.data(), g->c_OI.p, (size_t)n * D * 8, cudaMemcpyDeviceToHost, st));
  if (X) {  // coordinates of the children: the bit-exact get_x of the evaluation path (mul + add, never fused)
    if (int rc = g->s_tmp.reserve((size_t)n * D)) return rc;
    if (int rc = g->s_flag.reserve(1)) return rc;
    CK(cudaMemsetAsync(g->s_flag.p, 0, sizeof(int), st));
    DeviceGrid dg = g->dg;
    dg.D = D;
    for (int d = 0; d < D; ++d) {
      dg.lb[d] = g->lb[d];
      volatile double w = g->ub[d] - g->lb[d];
      dg.width[d] = w;
    }
    int64_t l = 0;
    CK(launch_nodes_x(D, nullptr, dg, g->c_OL.p, g->c_OI.p, n, g->s_tmp.p, D, 1, g->s_flag.p, st, &l));
    bump(l);
    if (int rc = copy_out_matrix(g->s_tmp.p, n, D, X, sp, sdim, st)) return rc;
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
  std::vector<double> rb;
  if (int rc = g->s_X[0].reserve((size_t)M * g->D)) return rc;
  if (int rc = stage_points(X, 0, M, g->D, sp, sdim, g->s_X[0].p, dsp, dsd, st, rb)) return rc;
  CK(cudaStreamSynchronize(st));
  return ASG_OK;
}

static cudaError_t run_basis(asg_grid *g, const BasisArgs &a, cudaStream_t st) {
  int64_t l = 0;
  cudaError_t e = (g->path() == ASG_PATH_SUBSPACE) ? launch_basis_subspace(g->dg, a, st, &l)
                                                   : launch_basis_sweep(g->dg, a, st, &l);
  bump(l);
  return e;
}

int32_t asg_basis_count(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                        int32_t dropzeros, int64_t *row_nnz) {
  if (int rc = check_eval_args(g, X, M, row_nnz)) return rc;
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (M == 0) return ASG_OK;
  cudaStream_t st = ctx.stream;
  BasisArgs a{};
  if (int rc = upload_points(g, X, M, sp, sdim, a.sp, a.sd, st)) return rc;
  if (int rc = g->s_i64a.reserve((size_t)M)) return rc;
  a.X = g->s_X[0].p; a.M = M; a.atol = atol; a.dropzeros = dropzeros; a.mode = 0;
  a.row_nnz = (unsigned long long *)g->s_i64a.p;
  CK(run_basis(g, a, st));
  CK(cudaMemcpyAsync(row_nnz, g->s_i64a.p, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return ASG_OK;
}

int32_t asg_basis_fill(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                       int32_t dropzeros, const int64_t *row_ptr, int64_t *colidx, double *vals) {
  if (int rc = check_eval_args(g, X, M, row_ptr)) return rc;
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (M == 0) return ASG_OK;
  const int64_t nnz = row_ptr[M];
  if (nnz < 0 || (nnz > 0 && (!colidx || !vals))) return set_err(ASG_EINVAL, "bad CSR buffers");
  cudaStream_t st = ctx.stream;
  BasisArgs a{};
  if (int rc = upload_points(g, X, M, sp, sdim, a.sp, a.sd, st)) return rc;
  if (int rc = g->s_i64a.reserve((size_t)M + 1)) return rc;
  if (int rc = g->s_i64b.reserve((size_t)nnz)) return rc;
  if (int rc = g->s_tmp.reserve((size_t)nnz)) return rc;
  CK(cudaMemcpyAsync(g->s_i64a.p, row_ptr, (size_t)(M + 1) * 8, cudaMemcpyHostToDevice, st));
  a.X = g->s_X[0].p; a.M = M; a.atol = atol; a.dropzeros = dropzeros; a.mode = 1;
  a.row_ptr = g->s_i64a.p; a.colidx = g->s_i64b.p; a.vals = g->s_tmp.p;
  CK(run_basis(g, a, st));
  if (g->path() == ASG_PATH_SUBSPACE) {  // trie order -> ascending node numbers inside each row
    int64_t l = 0;
    CK(launch_sort_rows(g->s_i64a.p, g->s_i64b.p, g->s_tmp.p, M, st, &l));
    bump(l);
  }
  if (nnz) {
    CK(cudaMemcpyAsync(colidx, g->s_i64b.p, (size_t)nnz * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(vals, g->s_tmp.p, (size_t)nnz * 8, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  return ASG_OK;
}

int32_t asg_basis_fill_csc(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                           int32_t dropzeros, const int64_t *row_ptr, int64_t *colptr, int64_t *rowidx, double *vals) {
  if (int rc = check_eval_args(g, X, M, row_ptr)) return rc;
  if (!colptr) return set_err(ASG_EINVAL, "null colptr");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  const int64_t N = g->N;
  if (M == 0) {
    for (int64_t c = 0; c <= N; ++c) colptr[c] = 0;
    return ASG_OK;
  }
  const int64_t nnz = row_ptr[M];
  if (nnz < 0 || nnz >= ((int64_t)1 << 31) || (nnz > 0 && (!rowidx || !vals))) return set_err(ASG_EINVAL, "bad CSC buffers");
  cudaStream_t st = ctx.stream;
  BasisArgs a{};
  if (int rc = upload_points(g, X, M, sp, sdim, a.sp, a.sd, st)) return rc;
  if (int rc = g->s_i64a.reserve((size_t)M + 1)) return rc;
  if (int rc = g->s_i64b.reserve((size_t)std::max<int64_t>(nnz, 1))) return rc;
  if (int rc = g->s_tmp.reserve((size_t)std::max<int64_t>(nnz, 1))) return rc;
  CK(cudaMemcpyAsync(g->s_i64a.p, row_ptr, (size_t)(M + 1) * 8, cudaMemcpyHostToDevice, st));
  a.X = g->s_X[0].p; a.M = M; a.atol = atol; a.dropzeros = dropzeros; a.mode = 1;
  a.row_ptr = g->s_i64a.p; a.colidx = g->s_i64b.p; a.vals = g->s_tmp.p;
  CK(run_basis(g, a, st));
  // CSR -> CSC: stable sort of the entries by column (rows stay ascending inside a column); scratch shared with the
  // point ordering of the BLOCK walk (slot 1 holds the outputs)
  const size_t n1 = (size_t)std::max<int64_t>(nnz, 1);
  const size_t tb = csc_temp_bytes(std::max<int64_t>(nnz, 1));
  if (int rc = g->s_key[0].reserve(n1)) return rc;
  if (int rc = g->s_key2[0].reserve(n1)) return rc;
  if (int rc = g->s_idx[0].reserve(n1)) return rc;
  if (int rc = g->s_perm[0].reserve(n1)) return rc;
  if (int rc = g->s_key[1].reserve(n1)) return rc;
  if (int rc = g->s_sorttmp[0].reserve(tb)) return rc;
  DevBuf<long long> &d_colptr = g->c_PL, &d_rowidx = g->c_PI;  // adapt! scratch, free between calls
  if (int rc = d_colptr.reserve((size_t)N + 1)) return rc;
  if (int rc = d_rowidx.reserve(n1)) return rc;
  if (int rc = g->s_tmp2.reserve(n1)) return rc;
  if (g->sort_ev[0]) CK(cudaStreamWaitEvent(st, g->sort_ev[0], 0));
  int64_t l = 0;
  CK(csr_to_csc(g->s_i64a.p, g->s_i64b.p, g->s_tmp.p, M, N, nnz, g->s_key[0].p, g->s_key2[0].p, g->s_idx[0].p,
                g->s_perm[0].p, g->s_key[1].p, g->s_sorttmp[0].p, tb, d_colptr.p, d_rowidx.p, g->s_tmp2.p, st, &l));
  bump(l);
  CK(cudaMemcpyAsync(colptr, d_colptr.p, (size_t)(N + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (nnz) {
    CK(cudaMemcpyAsync(rowidx, d_rowidx.p, (size_t)nnz * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(vals, g->s_tmp2.p, (size_t)nnz * 8, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  g->ch_n = -1;  // the adapt! scratch was reused
  return ASG_OK;
}

int32_t asg_basis_dense_device(asg_grid *g, const double *dX, int64_t M, int64_t sp, int64_t sdim, double atol,
                               int32_t dropzeros, double *dout, int64_t ldm, int64_t ldn, void *stream) {
  if (int rc = check_eval_args(g, dX, M, dout)) return rc;
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (M == 0 || g->N == 0) return ASG_OK;
  cudaStream_t st = (cudaStream_t)stream;  // NULL is the CUDA default stream, as everywhere in CUDA
  BasisArgs a{};
  a.X = dX; a.M = M; a.sp = sp; a.sd = sdim; a.atol = atol; a.dropzeros = dropzeros; a.mode = 2;
  a.dense = dout; a.ldm = ldm; a.ldn = ldn;
  CK(cudaEventRecord(ctx.ev0, st));
  if (g->path() == ASG_PATH_SUBSPACE) {
    // the walk only visits supporting nodes: clear first (needs a dense M x N block)
    const bool rowmajor = (ldn == 1 && ldm >= g->N), colmajor = (ldm == 1 && ldn >= M);
    if (rowmajor) CK(cudaMemset2DAsync(dout, (size_t)ldm * 8, 0, (size_t)g->N * 8, (size_t)M, st));
    else if (colmajor) CK(cudaMemset2DAsync(dout, (size_t)ldn * 8, 0, (size_t)M * 8, (size_t)g->N, st));
    else return set_err(ASG_EUNSUPPORTED, "dense basis needs ldn == 1 (row-major) or ldm == 1 (column-major)");
  }
  CK(run_basis(g, a, st));
  CK(cudaEventRecord(ctx.ev1, st));
  ctx.last_timed = true;
  return ASG_OK;
}

int32_t asg_basis_dense(asg_grid *g, const double *X, int64_t M, int64_t sp, int64_t sdim, double atol,
                        int32_t dropzeros, double *out, int64_t ldm, int64_t ldn) {
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
  for (int64_t m0 = 0; m0 < M; m0 += chunk) {
    const int64_t cnt = std::min(chunk, M - m0);
    int64_t dsp, dsd;
    if (int rc = g->s_X[1].reserve((size_t)cnt * g->D)) return rc;
    if (int rc = g->s_tmp2.reserve((size_t)cnt * N)) return rc;
    if (int rc = stage_points(X, m0, cnt, g->D, sp, sdim, g->s_X[1].p, dsp, dsd, ctx.stream, rb)) return rc;
    const int64_t dldm = rowmajor ? N : 1, dldn = rowmajor ? 1 : cnt;
    if (int rc = asg_basis_dense_device(g, g->s_X[1].p, cnt, dsp, dsd, atol, dropzeros, g->s_tmp2.p, dldm, dldn, ctx.stream))
      return rc;
    if (rowmajor)
      CK(cudaMemcpy2DAsync(out + m0 * ldm, (size_t)ldm * 8, g->s_tmp2.p, (size_t)N * 8, (size_t)N * 8, (size_t)cnt,
                           cudaMemcpyDeviceToHost, ctx.stream));
    else
      CK(cudaMemcpy2DAsync(out + m0, (size_t)ldn * 8, g->s_tmp2.p, (size_t)cnt * 8, (size_t)cnt * 8, (size_t)N,
                           cudaMemcpyDeviceToHost, ctx.stream));
    CK(cudaStreamSynchronize(ctx.stream));
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
  if (!pk.canonical) t_err = pk.why;
  else if (!pk.blk_ok) t_err = "no BLOCK layout: " + pk.blk_why;
  return ASG_OK;
}

// ---- device-format persistence -------------------------------------------------
int32_t asg_grid_export(asg_grid *g, void *buf, int64_t capacity, int64_t *bytes) {
  if (!g || !bytes) return set_err(ASG_EINVAL, "null argument");
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  if (!g->uploaded) return set_err(ASG_ESTATE, "no nodes uploaded");
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

int32_t asg_grid_import(asg_grid *g, const void *buf, int64_t bytes) {
  if (!g || !buf || bytes <= 0) return set_err(ASG_EINVAL, "null argument");
  if (int rc = require_ctx()) return rc;
  std::lock_guard<std::recursive_mutex> lk(g->mu);
  PackResult pk;
  std::vector<int64_t> L, I;
  std::vector<double> c;
  std::string err;
  int rc = pack_deserialize((const unsigned char *)buf, (size_t)bytes, g->D, g->lb, g->ub, L, I, c, pk, err);
  if (rc != ASG_OK) return set_err(rc, err);
  g->uploaded = false;
  g->levels.swap(L);
  g->indices.swap(I);
  g->coefs.swap(c);
  g->N = (int64_t)g->coefs.size();
  return install_pack(g, pk);
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
  if (!fma_per_second) return set_err(ASG_EINVAL, "null pointer");
  if (int rc = require_ctx()) return rc;
  double *sink = nullptr;
  CK(cudaMalloc(&sink, 8));
  const int iters = 1 << 16;
  int blocks = 0, threads = 0;
  int64_t l = 0;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(ctx.ev0, ctx.stream));
    CK(launch_fp64_peak(iters, sink, ctx.stream, &l, &blocks, &threads));
    CK(cudaEventRecord(ctx.ev1, ctx.stream));
    CK(cudaEventSynchronize(ctx.ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx.ev0, ctx.ev1));
    const double rate = (double)blocks * threads * 8.0 * iters / (ms * 1e-3);
    if (rep > 0) best = std::max(best, rate);
  }
  bump(l);
  cudaFree(sink);
  *fma_per_second = best;
  return ASG_OK;
}

}  // extern "C"
