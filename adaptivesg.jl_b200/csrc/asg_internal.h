// asg_internal.h -- shared declarations of libasg_cuda.so (host packer, kernels, C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <algorithm>
#include <string>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <utility>
#include <vector>

#include "asg_cuda.h"
#include "asg_block.h"

namespace asg {

constexpr int kMaxDim = ASG_MAX_DIM;
constexpr int kMaxDimFast = ASG_MAX_DIM_FAST;
constexpr int kKeyBits = 30;        // per-dimension cell key q = floor(x01 * 2^30), clamped to 2^30-1
constexpr int kMaxLevelFast = 32;   // cell bits of a level: b(l) = 0, 1, l-2  ->  l <= 32
constexpr int kMaxPosBits = 32;     // position of a node inside its subspace must fit a uint32

// ---- level-vector trie (device layout) -------------------------------------------------------
// One 16-byte record per trie node. Level k of the trie fixes the hierarchical level of dimension k.
//   inner node (k < D-1): x = index of its first child in level k+1 (children are contiguous and
//                         sorted by level; the end is the next record's x, a sentinel closes the array)
//   leaf       (k = D-1): x = first coefficient slot of the subspace (dense) or first hash slot
//   y: bits 0..7 level, bit 8 = hashed subspace, bits 16..23 = log2(hash table size),
//      bits 24..29 = cell bits b(level), bit 30 = (level > 2)
//   nodes of level D-2 ("prefix groups", the parents of the leaves) additionally carry
//   w: bits 0..7 pb = position bits of dimensions 0..D-2, bits 8..15 lm, bit 16 FULL
//   z: FULL only: coefficient slot of the first child. FULL means the children are the levels
//      1..lm (lm <= kLeafTable), all dense and stored back to back, so the slot of the supporting
//      node of level l is  z + ((off(l) + cell_l) << pb) + O  with off(l) = 0,1,3,5,9,17,...
// Position of a node inside its subspace: the cells of dimensions 0..D-2 concatenated (dimension 0
// most significant) form O (pb bits); the cell of the LAST dimension sits above them.
struct __align__(16) TrieRec {
  uint32_t x;
  uint32_t y;
  uint32_t z;
  uint32_t w;
};
constexpr uint32_t kRecHashed = 1u << 8;
constexpr uint32_t kRecFull = 1u << 16;
constexpr int kLeafTable = 8;  // last-dimension levels 1..8 are served from per-thread register tables

struct HashEnt {  // open addressing, linear probing; key = pos + 1 (0 = empty)
  unsigned long long key;
  double coef;
};

// allocator whose vectors do not zero-fill on resize(n): for the large host arrays of a pack that are written in full by
// the (threaded) passes right after -- the serial zero fill of 100+ MB was a fifth of the pack of the benchmark grid
template <class T>
struct NoInitAlloc {
  using value_type = T;
  NoInitAlloc() = default;
  template <class U> NoInitAlloc(const NoInitAlloc<U> &) {}
  T *allocate(size_t n) { return static_cast<T *>(::operator new(n * sizeof(T))); }
  void deallocate(T *p, size_t) { ::operator delete(p); }
  template <class U> void construct(U *p) { ::new ((void *)p) U; }  // default-initialisation: no fill
  template <class U, class... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
  template <class U> bool operator==(const NoInitAlloc<U> &) const { return true; }
  template <class U> bool operator!=(const NoInitAlloc<U> &) const { return false; }
};

struct SweepDim {  // one (node, dimension) of the sweep engine: phi = hat(x*s - c)
  double s;        // 2^(l-1), or 0 for l == 1
  double c;        // (double) i, or 0 for l == 1
};

// ---- flat group stream (asg_kernels_stream.cu) -----------------------------------------------------
// One 48-byte record per prefix group (= trie node of level D-2) in depth-first order, carrying the
// whole path so that no tree bookkeeping is left on the device: the kernel runs ONE flat loop over
// the records, re-derives the prefix products only from dimension k0 on, and evaluates the group's
// leaves. Records and the coefficient slices they reference are streamed through shared memory in
// tiles (cp.async.bulk + mbarrier, double buffered).
struct __align__(16) GroupRec {
  uint32_t base1;   // FULL: coefficient slot of the level-1 leaf; GENERIC: first leaf record (rec[D-1])
  uint32_t npre8;   // 8 << pb: byte stride between consecutive last-dimension cells
  uint32_t meta;    // bits 0..7 lm | 8..15 k0 | 16..23 used = sum(l-1) of the prefix | 24 FULL | 25 GLOBAL
  uint32_t nchild;  // GENERIC: number of leaf records; SUPER: lmB of lA = 1..8, 4 bits each
  uint32_t lev_lo;  // levels of dimensions 0..D-2, 5 bits each (dims 0..5 here,
  uint32_t lev_hi;  //  dims 6..11 here)
  uint32_t s_hi;    // dimension D-2 (SUPER: D-3): high word of the double 2^(l-1) (unused when l == 1)
  uint32_t bod;     // same dimension: bits 0..7 cell bits b | bit 8 (l > 2); bits 16..23 pb
  uint16_t offs[8]; // SUPER: start of the slice of group lA = 1..8, in cells of npre8 bytes, relative to base1
};
static_assert(sizeof(GroupRec) == 48, "GroupRec is streamed as three 16-byte words");
constexpr uint32_t kGrpFull = 1u << 24;
// SUPER record: all groups of one level-(D-3) node merged. They are the levels 1..lmA (meta bits 0..7)
// of dimension D-2, each FULL with lmB(lA) levels of the last dimension (nchild: 4 bits per lA), stored
// back to back starting at base1; npre8 = 8 << (position bits of dimensions 0..D-3).
constexpr uint32_t kGrpSuper = 1u << 26;
constexpr uint32_t kGrpGlobal = 1u << 25;  // coefficients are read from global memory, not from the tile
struct StreamTile {  // records [r0, r1) reference FULL-group coefficient slots inside [c0, c1)
  uint32_t r0, r1, c0, c1;
};
constexpr int kTileCoefs = 2048;  // doubles per shared-memory coefficient tile (16 KB)
constexpr int kTileRecs = 128;    // records per tile (6 KB)

// Everything the kernels need, passed by value as a kernel parameter.
struct DeviceGrid {
  int D;
  double lb[kMaxDim];
  double width[kMaxDim];  // ub - lb, rounded once like (from_ub .- from_lb), src/math.jl:308
  // subspace engine
  const TrieRec *rec[kMaxDimFast];  // per trie level, each with a sentinel record
  uint32_t n_root;                  // records in level 0
  uint32_t n_level1;                // records in level 1 (0 for D == 1)
  const double *coef;               // dense subspace blocks (zero padded)
  const int32_t *orig;              // slot -> caller's node number, -1 = padding
  const HashEnt *hent;              // hashed subspaces
  const int32_t *horig;
  const GroupRec *grp;              // flat group stream
  const StreamTile *tiles;
  uint32_t n_grp, n_tiles;
  // BLOCK layout (asg_block.h): present when the grid is (nearly) regular in its last three dimensions
  const double *bcoef;
  const BlockRec *brec;      // device record stream: a header record in front of every tile (blk_tile_header)
  const BlockTile *btiles;   // [r0, r1) counts the header
  uint32_t n_brec, n_btiles;
  int blk_slots;                    // saved prefix slots the walk needs (<= kBlkMaxSlots)
  int blk_tile_coefs;               // doubles per coefficient tile buffer (chosen by the packer)
  int blk_pair_ok;                  // the record set admits PAIR mode (asg_block.h)
  int blk_wide;                     // some record has a suffix budget above 7: the WIDE kernels walk this grid
  // sweep engine (caller's node order)
  int64_t N;
  const SweepDim *sw;       // [N][D]
  const uint64_t *sw_low;   // bit d set: l_nd <= 2 (phi is 0 outside [0,1] whatever the formula says)
  const double *sw_coef;    // [N]
  const int32_t *sw_depth;  // [N]
};

// ---- regular-sparse-grid fast path (asg_kernels_rsg.cu) ------------------------------------------
constexpr int kRsgMaxBudget = 31;
struct RsgParams {
  int budget;                 // max over nodes of sum(l_d - 1)  (= max depth - 1)
  int cut;                    // set per launch: budget - mask budget (0 = no depth mask)
  int ml[kMaxDimFast];        // effective max level per dimension: min(max level, budget + 1)
  // S[k][rem]: coefficient slots (per unit of prefix positions) of all level vectors over dimensions
  // k..D-1 with sum(l-1) <= rem; S[D][*] = 1
  uint32_t S[kMaxDimFast + 1][kRsgMaxBudget + 1];
};

// ---- host-side packed grid -------------------------------------------------------------------
struct PackResult {
  bool canonical = false;
  std::string why;  // why the grid is not canonical (empty when it is)
  int max_level = 0;
  int64_t max_depth = 0;
  int64_t subspaces = 0, dense_subspaces = 0, trie_nodes = 0, full_groups = 0;
  std::vector<std::vector<TrieRec>> rec;  // [D] levels, sentinel included
  std::vector<double> coef;           // + coef_pad zero slots at the end (tile copies are 16-byte granular)
  int64_t coef_pad = 0;
  std::vector<int32_t> orig;
  std::vector<HashEnt> hent;
  std::vector<int32_t> horig;
  std::vector<int64_t> slot_of_node;  // >= 0: dense slot; < 0: -(hash slot) - 1
  std::vector<GroupRec> grp;          // flat group stream (D >= 2)
  std::vector<StreamTile> tiles;
  bool regular = false;               // node set is exactly a regular sparse grid: RsgParams valid
  RsgParams rp{};
  // BLOCK layout (asg_block.h)
  bool blk_ok = false;
  std::string blk_why;                // why the grid does not take the BLOCK layout
  std::vector<double> bcoef;          // + 2 zero slots at the end (tile copies are 16-byte granular)
  std::vector<BlockRec> brec;
  std::vector<BlockTile> btiles;
  std::vector<int64_t> bslot_of_node;
  int blk_slots = 0;
  int blk_tile_coefs = 0;             // coefficient tile size the tiles were cut for
  bool blk_pair_ok = false;           // every record without kBrShared has a budget <= kBlkPairMaxR
  bool blk_wide = false;              // some record has a suffix budget above 7
  // sweep engine
  std::vector<SweepDim, NoInitAlloc<SweepDim>> sw;  // written in full by pack_grid's per-node pass
  std::vector<uint64_t> sw_low;
  std::vector<int32_t> depth;
};

// Validate + pack. levels/indices: row-major N x D copies owned by the caller. Returns ASG_OK,
// ASG_EINVALID_NODE (and err) or ASG_EUNSUPPORTED.
// ASG_PACK_TIMING=1: phases of the host packer on stderr (diagnostic)
struct PackTimer {
  const char *name;
  std::chrono::steady_clock::time_point t0;
  bool on;
  explicit PackTimer(const char *n) : name(n), t0(std::chrono::steady_clock::now()), on(getenv("ASG_PACK_TIMING") != nullptr) {}
  ~PackTimer() {
    if (on) fprintf(stderr, "[asg pack] %-28s %8.2f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
};
// host threads of the packer's data-parallel passes (ASG_PACK_THREADS overrides; small grids stay serial)
int pack_threads(int64_t N);
// fn(t, n0, n1) on nth contiguous ranges of [0, N), one host thread each
template <class F>
inline void pack_parallel(int64_t N, int nth, F fn) {
  if (nth <= 1) { fn(0, (int64_t)0, N); return; }
  std::vector<std::thread> th;
  const int64_t per = (N + nth - 1) / nth;
  for (int t = 0; t < nth; ++t) {
    const int64_t n0 = std::min<int64_t>(N, (int64_t)t * per), n1 = std::min<int64_t>(N, n0 + per);
    th.emplace_back([=]() { fn(t, n0, n1); });
  }
  for (auto &x : th) x.join();
}
int pack_grid(int D, int64_t N, const int64_t *levels, const int64_t *indices, const double *coefs,
              PackResult &out, std::string &err);

// BLOCK layout of a canonical grid (asg_pack_block.cpp). lv: N x D level bytes, cells: N x D cell numbers,
// perm: nodes sorted by level vector. Fills the blk_* members; blk_ok = false (and blk_why) if the grid
// does not qualify. selfcheck: walk the records on the host at a few points and compare with the direct sum.
int pack_blocks(int D, int64_t N, const uint8_t *lv, const uint32_t *cells, const int64_t *perm, const double *coefs,
                bool selfcheck, PackResult &out, std::string &err);
// host model of the BLOCK walk of asg_kernels_block.cu for one unit-cube point (packer self-check, tests)
double block_walk_host(int D, const PackResult &pk, const double *x01, int rem);
// device form of the BLOCK record stream: a header record in front of every tile (blk_tile_header, asg_block.h);
// dtiles[t] = tile t with [r0, r1) in device record numbers, header included
void block_device_stream(const PackResult &pk, std::vector<BlockRec> &drec, std::vector<BlockTile> &dtiles);
// follows the stream the way the kernel does (first tiles from the table, every later one from the header of the tile
// that held its buffer) and compares what it meets with the packed records
bool block_device_stream_ok(const PackResult &pk, std::string &why);

// device-format persistence (asg_persist.cpp): node arrays + every packed layout as one blob
void pack_serialize(int D, const double *lb, const double *ub, const std::vector<int64_t> &levels,
                    const std::vector<int64_t> &indices, const std::vector<double> &coefs, const PackResult &pk,
                    std::vector<unsigned char> &out);
int pack_deserialize(const unsigned char *buf, size_t bytes, int D, const double *lb, const double *ub,
                     std::vector<int64_t> &levels, std::vector<int64_t> &indices, std::vector<double> &coefs, PackResult &pk,
                     std::string &err);

// rsg! node enumeration on the host (asg_rsg.cpp, src/train.jl:40-78)
int rsg_count(int D, int64_t maxdepth, const int64_t *maxlevels, int64_t *N, std::string &err);
int rsg_fill(int D, int64_t maxdepth, const int64_t *maxlevels, int64_t N, int64_t *levels, int64_t *indices, int64_t sn,
             int64_t sd, std::string &err);

// Raise a kernel's dynamic shared-memory limit ONCE PER DEVICE to `bytes` (its maximum over all launches): function
// attributes are per device, and setting the limit per launch to that launch's own size races between host threads
// that launch the same instantiation with different sizes. `mask` is one static per instantiation (bit = device).
template <class K>
inline cudaError_t smem_limit_once(K kernel, int bytes, std::atomic<unsigned> &mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned bit = 1u << (dev & 31);
  if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) mask.fetch_or(bit, std::memory_order_release);
  return e;
}

// ---- kernel launchers (asg_kernels_*.cu) -------------------------------------------------------
struct EvalArgs {
  const double *X;   // device
  int64_t M, sp, sd;
  int rem;             // depth budget: sum(l-1) <= rem  (INT_MAX/2: no mask)
  const double *fvals; // optional: out = fvals - G
  double *out;
  const uint32_t *perm = nullptr;  // optional: thread j handles point perm[j] (points pre-sorted by cell key)
  // PAIR mode of the BLOCK walk: second points of pairs that do not share a cell go to this list (original point
  // numbers); a follow-up launch evaluates them with perm = left_list and its point count read from M_dev
  uint32_t *left_list = nullptr;
  uint32_t *left_count = nullptr;
  const uint32_t *M_dev = nullptr;
  // sweep engine only: nodes n_first .. N-1, and accumulate != 0 ADDS their sum to what `out` already holds
  // (out += G_delta, or out -= G_delta for a surplus): the appended-but-not-yet-packed nodes of adapt! (asg_grid_append)
  int64_t n_first = 0;
  int accumulate = 0;
};
struct BasisArgs {
  const double *X;
  int64_t M, sp, sd;
  double atol;
  int dropzeros;
  int mode;                 // 0 count, 1 fill CSR, 2 dense
  unsigned long long *row_nnz;     // mode 0 (device)
  const long long *row_ptr;        // mode 1
  long long *colidx;               // mode 1
  double *vals;                    // mode 1
  double *dense;                   // mode 2
  int64_t ldm, ldn;
};

cudaError_t launch_eval_subspace(const DeviceGrid &g, const EvalArgs &a, cudaStream_t st, int64_t *launches);
cudaError_t launch_eval_stream(const DeviceGrid &g, const EvalArgs &a, cudaStream_t st, int64_t *launches);
// left_list (capacity M / 2 + 1) / left_count: scratch of PAIR mode, may be null (then PAIR mode is not used)
cudaError_t launch_eval_block(const DeviceGrid &g, const EvalArgs &a, cudaStream_t st, int64_t *launches,
                              uint32_t *left_list = nullptr, uint32_t *left_count = nullptr);
bool block_pair_mode(const DeviceGrid &g, int64_t M, bool sorted);  // would launch_eval_block pair the points?
bool block_kernel_fits(const DeviceGrid &g);  // shared memory of the BLOCK walk fits an SM for this grid
// shared memory (bytes) of the BLOCK walk: R points per thread, `threads` threads, D dimensions, stack slots, tile size
size_t block_smem_bytes(int D, int slots, int tile_coefs, int R, int threads);
// bit-interleaved cell-key ordering of the query points (asg_sort.cu): perm_out[j] = j-th point in key order
size_t sort_points_temp_bytes(int64_t M);
cudaError_t sort_points(const DeviceGrid &g, const double *X, int64_t M, int64_t sp, int64_t sd, uint32_t *keys,
                        uint32_t *keys_alt, uint32_t *idx, uint32_t *perm_out, void *temp, size_t temp_bytes,
                        cudaStream_t st, int64_t *launches);
// CSR -> CSC of a basis matrix on the device (asg_sort.cu)
size_t csc_temp_bytes(int64_t nnz);
cudaError_t csr_to_csc(const long long *row_ptr, const long long *colidx, const double *vals, int64_t M, int64_t N,
                       int64_t nnz, uint32_t *keys, uint32_t *keys_alt, uint32_t *ent, uint32_t *perm, uint32_t *rows,
                       void *temp, size_t temp_bytes, long long *colptr, long long *rowidx, double *vout, cudaStream_t st,
                       int64_t *launches);
cudaError_t launch_narrow_i64(const long long *in, int *out, int64_t n, cudaStream_t st, int64_t *launches);
cudaError_t launch_eval_rsg(const DeviceGrid &g, const RsgParams &rp, const EvalArgs &a, cudaStream_t st,
                            int64_t *launches);
cudaError_t launch_eval_sweep(const DeviceGrid &g, const EvalArgs &a, cudaStream_t st, int64_t *launches);
cudaError_t launch_basis_subspace(const DeviceGrid &g, const BasisArgs &a, cudaStream_t st, int64_t *launches);
// dense basis scatter split over the nodes of trie level 1 (n1 of them) instead of one thread per row
cudaError_t launch_basis_dense_split(const DeviceGrid &g, const BasisArgs &a, uint32_t n1, cudaStream_t st, int64_t *launches);
cudaError_t launch_basis_sweep(const DeviceGrid &g, const BasisArgs &a, cudaStream_t st, int64_t *launches);
// node coordinates: li = [n][D] pairs (level, index) as int64x2, Xout element (k,d) at k*sp + d*sd
cudaError_t launch_nodes_x(int D, const double *lb_dev_or_null, const DeviceGrid &g, const long long *lev,
                           const long long *idx, int64_t n, double *Xout, int64_t sp, int64_t sd, int *bad,
                           cudaStream_t st, int64_t *launches);
// adapt! candidate generation (asg_kernels_adapt.cu): unique valid children of n_p parent rows, in the reference's
// visiting order; *total (device) receives their number
cudaError_t launch_children(int D, int64_t n_p, const long long *PL, const long long *PI, long long *CL, long long *CI,
                            unsigned char *valid, int *table, int64_t table_size, unsigned int *flags,
                            unsigned int *sums, unsigned int *total, long long *OL, long long *OI, cudaStream_t st,
                            int64_t *launches);
// elementwise phi (src/math.jl:332-391): out[k] = prod_d phi(x[k,d]; lev[k,d], idx[k,d]); lev == nullptr: phi(x[k,d])
cudaError_t launch_phi(int D, int64_t n, const double *x, const long long *lev, const long long *idx, double *out, int *bad,
                       cudaStream_t st, int64_t *launches);
cudaError_t launch_scatter_f64(double *dst, const long long *slots, const double *src, int64_t n, cudaStream_t st,
                               int64_t *launches);
cudaError_t launch_scatter_hash(HashEnt *dst, const long long *slots, const double *src, int64_t n, cudaStream_t st,
                                int64_t *launches);
cudaError_t launch_fp64_peak(int iters, double *sink, cudaStream_t st, int64_t *launches, int *blocks, int *threads);
cudaError_t launch_sort_rows(const long long *row_ptr, long long *colidx, double *vals, int64_t M, cudaStream_t st,
                             int64_t *launches);

}  // namespace asg
