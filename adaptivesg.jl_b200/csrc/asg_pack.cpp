// asg_pack.cpp -- host-side packing of the reference node store (src/class.jl:281-293: levels,
// indices :: Matrix{Int} N x D, coefs :: Vector{Float64}) into the two device layouts:
//
//  (1) SWEEP layout, caller's node order: per (node, dim) the pair (s, c) with phi = hat(x*s - c),
//      s = 2^(l-1) and c = i for l >= 2 (src/math.jl:347-351; the l == 2 forms are the same function
//      on [0,1]), s = c = 0 for l == 1 (src/math.jl:352-353). Works for every node set phi accepts.
//
//  (2) SUBSPACE layout, for "canonical" grids (every node passes isvalid, src/math.jl:418-448, has
//      an in-range index, appears once, and all coefficients are finite): nodes are grouped by level
//      vector ("subspace"). In one subspace the supports of the nodes tile [0,1]^D, so a point has
//      exactly one supporting node there and its position inside the subspace follows from the bits
//      of the point's coordinates. Subspaces are sorted lexicographically and organised as a trie
//      over dimensions so that prefix products are shared. A well filled subspace is stored as a
//      full block of 2^bits coefficients (missing nodes are 0.0), a sparsely filled one (adapt!
//      output) as a small open-addressing hash table keyed by the position.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

#include "asg_internal.h"

namespace asg {

static inline int cell_bits(int l) { return l == 1 ? 0 : (l == 2 ? 1 : l - 2); }

static inline uint32_t hash_slot(uint32_t pos, int lg) {
  // multiplicative hash; must match hash_slot() in asg_kernels_subspace.cu
  return lg == 0 ? 0u : (uint32_t)((pos * 0x9E3779B1u) >> (32 - lg));
}

int pack_threads(int64_t N) {
  if (N < (1 << 15)) return 1;
  static const int v = [] {
    const char *e = getenv("ASG_PACK_THREADS");
    int t = e ? atoi(e) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    return std::max(1, std::min(t, 64));
  }();
  return v;
}

int pack_grid(int D, int64_t N, const int64_t *levels, const int64_t *indices, const double *coefs,
              PackResult &out, std::string &err) {
  PackTimer t_all("pack_grid total");
  out = PackResult();
  out.sw.resize((size_t)N * D);
  out.sw_low.assign((size_t)N, 0);
  out.depth.resize((size_t)N);
  out.slot_of_node.assign((size_t)N, 0);

  bool canonical = (D <= kMaxDimFast);
  std::string why = canonical ? "" : "D > ASG_MAX_DIM_FAST";
  auto demote = [&](const char *w) {
    if (canonical) { canonical = false; why = w; }
  };

  std::vector<uint8_t> lv;   // N x D level bytes (canonical path only)
  std::vector<uint32_t> pos; // position inside the subspace
  std::vector<uint32_t, NoInitAlloc<uint32_t>> cells;  // N x D cell numbers (BLOCK layout); written by the per-node pass
  if (canonical) { lv.resize((size_t)N * D); pos.resize((size_t)N); cells.resize((size_t)N * D); }

  // per-node pass (validation, sweep words, level bytes, cells, subspace position): independent per node, so large
  // grids are scanned by a few host threads; errors and demotions are merged to what the serial scan reports (the
  // first offending node in input order)
  struct Scan {
    int err_code = ASG_OK;
    std::string err_msg;
    int64_t err_n = -1;
    bool canon = true;
    const char *why = "";
    int64_t why_n = -1;
    int max_level = 0;
    int64_t max_depth = 0;
  };
  const bool canon0 = canonical;
  auto scan = [&](int64_t n0, int64_t n1, Scan &sc) {
    sc.canon = canon0;
    auto demote_l = [&](const char *w, int64_t n) {
      if (sc.canon) { sc.canon = false; sc.why = w; sc.why_n = n; }
    };
    for (int64_t n = n0; n < n1; ++n) {
      int64_t dsum = 0;
      uint64_t p = 0;
      int tb = 0;
      for (int d = 0; d < D; ++d) {
        const int64_t l = levels[n * D + d], i = indices[n * D + d];
        // the rule of phi(x, l, i), src/math.jl:345-357: these are the pairs it throws on
        if (l < 1 || (l == 2 && i != 0 && i != 2)) {
          sc.err_msg = "invalid level-index pair (" + std::to_string(l) + ", " + std::to_string(i) + ") at node " +
                       std::to_string(n) + ", dim " + std::to_string(d);
          sc.err_code = ASG_EINVALID_NODE;
          sc.err_n = n;
          return;
        }
        if (l > 62) { sc.err_msg = "level > 62 is not supported"; sc.err_code = ASG_EUNSUPPORTED; sc.err_n = n; return; }
        dsum += l;
        sc.max_level = std::max<int>(sc.max_level, (int)l);
        SweepDim &s = out.sw[(size_t)n * D + d];
        if (l == 1) {
          s.s = 0.0; s.c = 0.0;
          out.sw_low[n] |= (1ull << d);
          if (i != 1) demote_l("level-1 index != 1", n);
        } else {
          s.s = std::ldexp(1.0, (int)(l - 1));  // power2(l-1), src/math.jl:59-61
          s.c = (double)i;                       // Int -> Float64 promotion in x*2^(l-1) - i
          if (l == 2) out.sw_low[n] |= (1ull << d);
        }
        if (sc.canon) {
          if (l > kMaxLevelFast) { demote_l("level > 32", n); continue; }
          if (l > 2 && (i < 1 || !(i & 1) || i > ((int64_t)1 << (l - 1)) - 1)) { demote_l("index out of range", n); continue; }
          const int b = cell_bits((int)l);
          const uint64_t cell = l == 1 ? 0 : (l == 2 ? (uint64_t)(i >> 1) : (uint64_t)((i - 1) >> 1));
          if (tb + b > kMaxPosBits) { demote_l("subspace position needs more than 32 bits", n); continue; }
          // dims 0..D-2: concatenate (dim 0 most significant); the last dim goes on top of them
          if (d < D - 1 || D == 1) p = (b ? (p << b) : p) | cell;
          else p |= cell << tb;
          tb += b;
          lv[(size_t)n * D + d] = (uint8_t)l;
          cells[(size_t)n * D + d] = (uint32_t)cell;
        }
      }
      const int64_t depth = dsum - D + 1;
      if (depth > 0x3fffffff) { sc.err_msg = "depth overflow"; sc.err_code = ASG_EUNSUPPORTED; sc.err_n = n; return; }
      out.depth[n] = (int32_t)depth;
      sc.max_depth = std::max(sc.max_depth, depth);
      if (sc.canon) pos[n] = (uint32_t)p;
    }
  };
  const int nth = pack_threads(N);
  std::vector<Scan> scans((size_t)nth);
  {
    PackTimer t_scan("per-node scan");
    pack_parallel(N, nth, [&](int t, int64_t n0, int64_t n1) { scan(n0, n1, scans[(size_t)t]); });
  }
  int max_level = 0;
  int64_t max_depth = 0;
  {
    const Scan *first_err = nullptr, *first_dem = nullptr;
    for (const Scan &sc : scans) {
      max_level = std::max(max_level, sc.max_level);
      max_depth = std::max(max_depth, sc.max_depth);
      if (sc.err_code != ASG_OK && (!first_err || sc.err_n < first_err->err_n)) first_err = &sc;
      if (canon0 && !sc.canon && (!first_dem || sc.why_n < first_dem->why_n)) first_dem = &sc;
    }
    if (first_err) {  // the serial scan stops at the first invalid node
      err = first_err->err_msg;
      return first_err->err_code;
    }
    if (first_dem) demote(first_dem->why);
  }
  out.max_level = max_level;
  out.max_depth = max_depth;
  if (N >= 0x7fffffff) demote("N >= 2^31");

  if (canonical && N > 0) {
    std::vector<int64_t> perm((size_t)N);
    std::iota(perm.begin(), perm.end(), (int64_t)0);
    const uint8_t *L = lv.data();
    {
      // order: level vector (bytes, lexicographic), position inside the subspace, node number. The D <= 12 level bytes
      // are packed big-endian into integers, so the order of the keys is the order memcmp gives; chunks are sorted by a
      // few threads and merged pairwise.
      PackTimer t_sort("sort nodes by level vector");
      static_assert(kMaxDimFast <= 12, "the sort key holds 12 level bytes");
      struct SortKey {
        uint64_t hi, lo;  // level bytes 0..7 | level bytes 8..11, position
        int64_t n;
        bool operator<(const SortKey &o) const { return hi != o.hi ? hi < o.hi : (lo != o.lo ? lo < o.lo : n < o.n); }
      };
      std::vector<SortKey, NoInitAlloc<SortKey>> keys((size_t)N);
      const int nth = pack_threads(N);
      const int64_t per = (N + nth - 1) / nth;
      pack_parallel(N, nth, [&](int, int64_t n0, int64_t n1) {
        for (int64_t n = n0; n < n1; ++n) {
          uint64_t hi = 0, lo = 0;
          for (int d = 0; d < 8; ++d) hi = (hi << 8) | (d < D ? L[(size_t)n * D + d] : 0u);
          for (int d = 8; d < 12; ++d) lo = (lo << 8) | (d < D ? L[(size_t)n * D + d] : 0u);
          keys[(size_t)n] = SortKey{hi, (lo << 32) | pos[n], n};
        }
        std::sort(keys.begin() + n0, keys.begin() + n1);
      });
      for (int64_t width = per; width < N && nth > 1; width *= 2) {  // merge tree over the sorted chunks
        std::vector<std::thread> th;
        for (int64_t lo0 = 0; lo0 + width < N; lo0 += 2 * width) {
          const int64_t mid = lo0 + width, hi0 = std::min<int64_t>(N, lo0 + 2 * width);
          th.emplace_back([&keys, lo0, mid, hi0]() { std::inplace_merge(keys.begin() + lo0, keys.begin() + mid, keys.begin() + hi0); });
        }
        for (auto &x : th) x.join();
      }
      for (int64_t k = 0; k < N; ++k) perm[(size_t)k] = keys[(size_t)k].n;
    }
    PackTimer t_rest("subspaces, trie, stream layout");
    out.rec.assign((size_t)D, {});
    std::vector<uint64_t> grp_levels;  // per prefix group: levels of dims 0..D-2, 5 bits each
    std::vector<uint8_t> grp_k0, grp_used;
    const uint8_t *prev = nullptr;
    int64_t g0 = 0;
    while (g0 < N && canonical) {
      const uint8_t *cur = L + (size_t)perm[g0] * D;
      int64_t g1 = g0 + 1;
      while (g1 < N && std::memcmp(L + (size_t)perm[g1] * D, cur, (size_t)D) == 0) {
        if (pos[perm[g1]] == pos[perm[g1 - 1]]) { demote("duplicate node"); break; }
        ++g1;
      }
      if (!canonical) break;
      const int64_t cnt = g1 - g0;
      int tb = 0;
      for (int d = 0; d < D; ++d) tb += cell_bits(cur[d]);
      const uint64_t full = 1ull << tb;
      uint32_t base, flags = 0;
      if (full <= (uint64_t)std::max<int64_t>(4, 2 * cnt)) {
        if (out.coef.size() + full >= 0xffffffffull) { demote("coefficient slots >= 2^32"); break; }
        base = (uint32_t)out.coef.size();
        out.coef.resize(out.coef.size() + full, 0.0);
        out.orig.resize(out.orig.size() + full, -1);
        for (int64_t k = g0; k < g1; ++k) {
          const int64_t n = perm[k];
          const size_t slot = (size_t)base + pos[n];
          out.coef[slot] = coefs ? coefs[n] : 0.0;
          out.orig[slot] = (int32_t)n;
          out.slot_of_node[n] = (int64_t)slot;
        }
        ++out.dense_subspaces;
      } else {
        int lg = 1;
        while ((1ll << lg) < 2 * cnt) ++lg;
        const uint64_t size = 1ull << lg;
        if (out.hent.size() + size >= 0xffffffffull) { demote("hash slots >= 2^32"); break; }
        base = (uint32_t)out.hent.size();
        out.hent.resize(out.hent.size() + size, HashEnt{0ull, 0.0});
        out.horig.resize(out.horig.size() + size, -1);
        for (int64_t k = g0; k < g1; ++k) {
          const int64_t n = perm[k];
          uint32_t s = hash_slot(pos[n], lg);
          while (out.hent[(size_t)base + s].key != 0ull) s = (s + 1) & (uint32_t)(size - 1);
          out.hent[(size_t)base + s] = HashEnt{(unsigned long long)pos[n] + 1ull, coefs ? coefs[n] : 0.0};
          out.horig[(size_t)base + s] = (int32_t)n;
          out.slot_of_node[n] = -(int64_t)((size_t)base + s) - 1;
        }
        flags = kRecHashed | ((uint32_t)lg << 16);
      }
      // trie: open new records from the first dimension whose level differs from the previous subspace
      int k0 = 0;
      if (prev) while (k0 < D - 1 && prev[k0] == cur[k0]) ++k0;
      for (int k = k0; k < D; ++k) {
        TrieRec r;
        r.y = cur[k] | ((uint32_t)cell_bits(cur[k]) << 24) | (cur[k] > 2 ? (1u << 30) : 0u);
        r.z = 0;
        r.w = 0;
        if (k < D - 1) {
          r.x = (uint32_t)out.rec[(size_t)k + 1].size();
        } else {
          r.x = base;
          r.y |= flags;
        }
        if (k == D - 2) {
          int pb = 0, used = 0;
          uint64_t lvp = 0;
          for (int d = 0; d < D - 1; ++d) {
            pb += cell_bits(cur[d]);
            used += cur[d] - 1;
            lvp |= (uint64_t)(cur[d] & 31) << (5 * d);
          }
          r.w = (uint32_t)pb;
          grp_levels.push_back(lvp);
          grp_k0.push_back((uint8_t)k0);  // first dimension whose level differs from the previous subspace
          grp_used.push_back((uint8_t)std::min(used, 255));
        }
        out.rec[(size_t)k].push_back(r);
      }
      prev = cur;
      ++out.subspaces;
      g0 = g1;
    }
    if (canonical && D >= 2) {
      // mark FULL prefix groups: children = levels 1..lm, all dense, back to back
      auto &par = out.rec[(size_t)D - 2];
      const auto &lf = out.rec[(size_t)D - 1];
      for (size_t j = 0; j < par.size(); ++j) {
        const uint32_t cb = par[j].x, ce = (j + 1 < par.size()) ? par[j + 1].x : (uint32_t)lf.size();
        const int pb = (int)(par[j].w & 0xffu);
        const int lm = (int)(ce - cb);
        bool full = lm >= 1 && lm <= kLeafTable;
        uint64_t expect = lf[cb].x;
        for (int c = 0; full && c < lm; ++c) {
          const TrieRec &r = lf[cb + c];
          if ((int)(r.y & 0xffu) != c + 1 || (r.y & kRecHashed) || r.x != expect) full = false;
          expect += 1ull << (pb + cell_bits(c + 1));
        }
        if (full) {
          par[j].z = lf[cb].x;
          par[j].w |= ((uint32_t)lm << 8) | kRecFull;
          ++out.full_groups;
        }
      }
    }
    if (canonical && D >= 2 && out.hent.empty() && (int64_t)out.coef.size() == N && max_depth - 1 <= kRsgMaxBudget) {
      // Is the node set exactly a regular sparse grid (src/train.jl:56-63)? Every subspace is complete
      // (coef slots == N, no padding, no hash) -- it remains to check that the set of level vectors is
      // {l : l_d <= ml_d, sum(l_d - 1) <= budget}: all present vectors satisfy that by construction of
      // ml and budget, so it suffices to compare counts.
      RsgParams &rp = out.rp;
      rp.budget = (int)(max_depth - 1);
      rp.cut = 0;
      std::vector<int> mlv((size_t)D, 1);
      for (int64_t n = 0; n < N; ++n)
        for (int d = 0; d < D; ++d) mlv[(size_t)d] = std::max<int>(mlv[(size_t)d], lv[(size_t)n * D + d]);
      for (int d = 0; d < D; ++d) rp.ml[d] = std::min(mlv[(size_t)d], rp.budget + 1);
      // C[k][rem]: number of level vectors over dims k..D-1; S[k][rem]: coefficient slots
      std::vector<std::vector<uint64_t>> Cn((size_t)D + 1, std::vector<uint64_t>((size_t)rp.budget + 1, 1));
      std::vector<std::vector<uint64_t>> Sn = Cn;
      bool fits = true;
      for (int k = D - 1; k >= 0; --k)
        for (int rem = 0; rem <= rp.budget; ++rem) {
          uint64_t c = 0, sz = 0;
          for (int l = 1; l <= std::min(rp.ml[k], rem + 1); ++l) {
            c += Cn[(size_t)k + 1][(size_t)(rem - (l - 1))];
            sz += Sn[(size_t)k + 1][(size_t)(rem - (l - 1))] << cell_bits(l);
          }
          Cn[(size_t)k][(size_t)rem] = c;
          Sn[(size_t)k][(size_t)rem] = sz;
          if (sz > 0xffffffffull) fits = false;
        }
      if (fits && Cn[0][(size_t)rp.budget] == (uint64_t)out.subspaces && Sn[0][(size_t)rp.budget] == (uint64_t)N &&
          rp.ml[D - 1] <= kLeafTable && rp.ml[D - 2] <= kLeafTable) {
        for (int k = 0; k <= D; ++k)
          for (int rem = 0; rem <= rp.budget; ++rem) rp.S[k][rem] = (uint32_t)Sn[(size_t)k][(size_t)rem];
        out.regular = true;
      }
    }
    if (canonical && D >= 2) {
      // flat group stream + shared-memory tiles
      const auto &par = out.rec[(size_t)D - 2];
      const auto &lf = out.rec[(size_t)D - 1];
      out.grp.resize(par.size());
      std::vector<uint32_t> span_lo(par.size(), 0), span_hi(par.size(), 0);
      for (size_t j = 0; j < par.size(); ++j) {
        const TrieRec &r = par[j];
        const uint32_t cb = r.x, ce = (j + 1 < par.size()) ? par[j + 1].x : (uint32_t)lf.size();
        const int pb = (int)(r.w & 0xffu);
        const int lA = (int)(r.y & 0xffu);
        GroupRec &q = out.grp[j];
        std::memset(&q, 0, sizeof(q));
        q.npre8 = 8u << pb;
        q.lev_lo = (uint32_t)(grp_levels[j] & 0x3fffffffull);
        q.lev_hi = (uint32_t)(grp_levels[j] >> 30);
        q.s_hi = (uint32_t)(1023 + lA - 1) << 20;
        q.bod = (uint32_t)cell_bits(lA) | (lA > 2 ? 0x100u : 0u) | ((uint32_t)pb << 16);
        q.meta = ((uint32_t)grp_k0[j] << 8) | ((uint32_t)grp_used[j] << 16);
        if (r.w & kRecFull) {
          const int lm = (int)((r.w >> 8) & 0xffu);
          q.base1 = r.z;
          q.nchild = (uint32_t)lm;
          q.meta |= (uint32_t)lm | kGrpFull;
          span_lo[j] = r.z;
          uint64_t cells = 0;
          for (int l = 1; l <= lm; ++l) cells += 1ull << cell_bits(l);
          span_hi[j] = (uint32_t)(r.z + (cells << pb));
        } else {
          q.base1 = cb;
          q.nchild = ce - cb;
          q.meta |= kGrpGlobal;
        }
      }
      // merge the groups of one level-(D-3) node into a single SUPER record when they are the levels
      // 1..lmA of dimension D-2, all FULL and stored back to back (always true on regular grids)
      {
        std::vector<GroupRec> merged;
        std::vector<uint32_t> mlo, mhi;
        const size_t ngrp = par.size();
        std::vector<uint32_t> pbeg;  // child ranges of the level-(D-3) nodes inside rec[D-2]
        if (D >= 3) {
          for (const TrieRec &r : out.rec[(size_t)D - 3]) pbeg.push_back(r.x);
        } else {
          pbeg.push_back(0u);
        }
        pbeg.push_back((uint32_t)ngrp);
        for (size_t pi = 0; pi + 1 < pbeg.size(); ++pi) {
          const uint32_t cb = pbeg[pi], ce = pbeg[pi + 1];
          const int lmA = (int)(ce - cb);
          bool super = lmA >= 1 && lmA <= kLeafTable;
          uint32_t lmBs = 0;
          for (uint32_t j = cb; super && j < ce; ++j) {
            const GroupRec &q = out.grp[j];
            if (!(q.meta & kGrpFull) || (int)(par[j].y & 0xffu) != (int)(j - cb) + 1) super = false;
            else if (j > cb && span_lo[j] != span_hi[j - 1]) super = false;
            else lmBs |= (q.meta & 0xfu) << (4 * (j - cb));
          }
          // slice starts are stored as 16-bit cell counts (cells of group lA = 1, npre8 bytes each)
          if (super && (span_lo[ce - 1] - span_lo[cb]) / (out.grp[cb].npre8 / 8u) > 0xffffu) super = false;
          if (super) {
            GroupRec q = out.grp[cb];
            const uint32_t np_cells = q.npre8 / 8u;
            for (int a = 0; a < 8; ++a) q.offs[a] = 0;
            for (uint32_t j = cb; j < ce; ++j) q.offs[j - cb] = (uint16_t)((span_lo[j] - span_lo[cb]) / np_cells);
            q.meta = (uint32_t)lmA | (q.meta & 0x00ffff00u) | kGrpSuper;  // k0 and used of the first child (lA = 1)
            q.nchild = lmBs;
            if (D >= 3) {  // s_hi / bod of a SUPER record describe the level of dimension D-3 (walked by every record)
              const int l3 = (int)((grp_levels[cb] >> (5 * (D - 3))) & 31u);
              q.s_hi = (uint32_t)(1023 + l3 - 1) << 20;
              q.bod = (uint32_t)cell_bits(l3) | (l3 > 2 ? 0x100u : 0u) | (q.bod & 0x00ff0000u);
            }
            merged.push_back(q);
            mlo.push_back(span_lo[cb]);
            mhi.push_back(span_hi[ce - 1]);
          } else {
            for (uint32_t j = cb; j < ce; ++j) {
              merged.push_back(out.grp[j]);
              mlo.push_back(span_lo[j]);
              mhi.push_back(span_hi[j]);
            }
          }
        }
        out.grp.swap(merged);
        span_lo.swap(mlo);
        span_hi.swap(mhi);
      }
      // greedy tiling: a tile covers consecutive records whose dense slices fit kTileCoefs doubles
      size_t j = 0;
      const size_t nrec = out.grp.size();
      while (j < nrec) {
        StreamTile t{(uint32_t)j, (uint32_t)j, 0u, 0u};
        bool have = false;
        while (j < nrec && (int)(j - t.r0) < kTileRecs) {
          GroupRec &q = out.grp[j];
          if (q.meta & (kGrpFull | kGrpSuper)) {
            const uint32_t lo = span_lo[j] & ~1u, hi = (span_hi[j] + 1u) & ~1u;  // 16-byte aligned bulk copies
            if (hi - lo > (uint32_t)kTileCoefs) {
              q.meta |= kGrpGlobal;  // a single slice larger than a tile: read it from global memory
            } else if (!have) {
              t.c0 = lo; t.c1 = hi; have = true;
            } else {
              const uint32_t nlo = std::min(t.c0, lo), nhi = std::max(t.c1, hi);
              if (nhi - nlo > (uint32_t)kTileCoefs) break;  // does not fit any more: close the tile
              t.c0 = nlo; t.c1 = nhi;
            }
          }
          ++j;
        }
        t.r1 = (uint32_t)j;
        out.tiles.push_back(t);
      }
      // self-check of the stream the kernels will trust blindly: every SUPER / FULL record that reads from a
      // tile must have its whole slice inside that tile's [c0, c1) window, windows must fit the buffers
      for (const StreamTile &t : out.tiles) {
        if (t.c1 - t.c0 > (uint32_t)kTileCoefs || t.r1 - t.r0 > (uint32_t)kTileRecs || (t.c0 & 1u) || (t.c1 & 1u) ||
            (uint64_t)t.c1 > out.coef.size() + 2) {
          err = "internal error: stream tile out of bounds";
          return ASG_ESTATE;
        }
        for (uint32_t r = t.r0; r < t.r1; ++r) {
          const GroupRec &q = out.grp[r];
          if ((q.meta & (kGrpFull | kGrpSuper)) && !(q.meta & kGrpGlobal) && (span_lo[r] < t.c0 || span_hi[r] > t.c1)) {
            err = "internal error: record slice outside its tile";
            return ASG_ESTATE;
          }
        }
      }
      out.coef_pad = 2;
      out.coef.resize(out.coef.size() + 2, 0.0);  // tiles round their end up to an even slot
      out.orig.resize(out.orig.size() + 2, -1);
    }
    if (canonical) {
      static const bool force_check = getenv("ASG_PACK_SELFCHECK") != nullptr;
      // a grid with a NaN / Inf coefficient evaluates through the sweep engine (Inf * 0 = NaN like the reference,
      // install_pack): its BLOCK layout is built but the self-check's comparison of sums would be NaN vs NaN
      bool finite_c = true;
      for (int64_t n = 0; coefs && n < N && finite_c; ++n) finite_c = std::isfinite(coefs[n]);
      PackTimer t_blk("BLOCK layout (all tile sizes)");
      const int rc = pack_blocks(D, N, lv.data(), cells.data(), perm.data(), coefs, (force_check || N <= 50000) && finite_c, out, err);
      if (rc != ASG_OK) return rc;
    }
    if (canonical) {
      for (int k = 0; k < D; ++k) {
        out.trie_nodes += (int64_t)out.rec[(size_t)k].size();
        TrieRec s;
        s.x = (k < D - 1) ? (uint32_t)out.rec[(size_t)k + 1].size() : 0u;
        s.y = 0xffu;  // level 255: terminates every masked walk
        s.z = 0;
        s.w = 0;
        // the sentinel of level k+1 is pushed after this one, so .size() above is the real count
        out.rec[(size_t)k].push_back(s);
      }
    }
  }
  if (!canonical) {
    out.rec.clear(); out.coef.clear(); out.orig.clear(); out.hent.clear(); out.horig.clear();
    out.subspaces = out.dense_subspaces = out.trie_nodes = 0;
  }
  out.canonical = canonical && N > 0;
  out.why = why;
  return ASG_OK;
}

}  // namespace asg
