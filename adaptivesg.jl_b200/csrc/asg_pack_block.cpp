// asg_pack_block.cpp -- host-side construction of the BLOCK layout (asg_block.h) of a canonical grid,
// and a host model of the walk the kernel performs over it (used as the packer's self-check).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

#include "asg_internal.h"

namespace asg {

namespace {

using Key = std::vector<uint8_t>;  // levels of the prefix dimensions 0..D-4 (always full length)

struct Rec {             // one record before S2 splitting
  int depth = 0;         // data-driven prefix: dimensions 0..depth-1; compiled suffix: S = D - depth dimensions
  int S = 3;
  int r = -1;            // budget of the compiled suffix: max sum(l-1) over its dimensions; -1: stage only
  int pb = 0;            // position bits of the prefix
  int used = 0;          // sum(l-1) of the prefix
  int nt = 0;            // non-trivial prefix dimensions (level > 1)
  int k0 = 0;            // last non-trivial prefix dimension
  bool parent_of = false;  // some other record restarts from this one
  bool split = false;    // S = 3 only: S2 records, one per level of A
  uint32_t cbase = 0;
  uint32_t cbaseA[kBlkMaxBudget + 1] = {0};  // S2: first slot of level lA = 1..r+1
};

Key parent_of(const Key &p) {
  Key q = p;
  for (int k = (int)q.size() - 1; k >= 0; --k)
    if (q[(size_t)k] > 1) { q[(size_t)k] = 1; break; }
  return q;
}

inline double hat_host(double x, int l, uint32_t cell) {  // phi of the supporting node, as the kernel forms it
  if (l == 1) return 1.0;
  const uint32_t i = 2u * cell + (l > 2 ? 1u : 0u);
  const double t = std::fma(x, std::ldexp(1.0, l - 1), -(double)i);
  return 1.0 - std::fabs(t);
}
inline uint32_t cell_host(double x, int l) {
  const int b = blk_bits(l);
  const uint32_t c = (uint32_t)(x * std::ldexp(1.0, b));
  return std::min(c, (1u << b) - 1u);
}

// the layout functions of asg_block.h are recursive (made for compile-time use): tabulate them once for the host
struct BlkTables {
  uint32_t size[kBlkMaxS + 1][kBlkMaxBudget + 1];
  uint32_t off[kBlkMaxS + 1][kBlkMaxBudget + 3][kBlkMaxBudget + 1];
  BlkTables() {
    for (int S = 1; S <= kBlkMaxS; ++S)
      for (int r = 0; r <= kBlkMaxBudget; ++r) {
        size[S][r] = blk_size(S, r);
        for (int l = 1; l <= kBlkMaxBudget + 2; ++l) off[S][l][r] = S >= 2 ? blk_off(S, l, r) : 0u;
      }
  }
};
const BlkTables &tabs() {
  static const BlkTables t;
  return t;
}
inline uint32_t t_size(int S, int r) { return r < 0 ? 0u : tabs().size[S][r]; }
inline uint32_t t_off(int S, int l, int r) { return tabs().off[S][l][r]; }

int max_compiled_dims() {  // experiment knob: ASG_BLOCK_MAXS=3 keeps one record per prefix level vector
  static const int v = [] {
    const char *e = getenv("ASG_BLOCK_MAXS");
    const int s = e ? atoi(e) : kBlkMaxS;
    return std::min(std::max(s, 3), kBlkMaxS);
  }();
  return v;
}

// compiled suffix of S dimensions (D-S .. D-1), storage budget rs, evaluated up to budget reff, Horner form.
// Every read must stay inside the block of size(S, rs) doubles it belongs to: NaN (a failed self-check) otherwise.
double eval_suffix_host(int D, int S, int rs, int reff, const double *blk, const double *x01) {
  if (S == 1) {  // pole
    const double xC = x01[D - 1];
    double inner = blk[0];
    for (int lC = 2; lC <= reff + 1; ++lC) {
      const uint32_t cC = cell_host(xC, lC);
      if (blk_off1(lC) + cC >= blk_n1(rs)) return std::nan("");
      inner = std::fma(blk[blk_off1(lC) + cC], hat_host(xC, lC, cC), inner);
    }
    return inner;
  }
  const double x = x01[D - S];
  double out = 0.0;
  for (int l = 1; l <= reff + 1; ++l) {
    const int rb = rs - (l - 1);
    const uint32_t c = cell_host(x, l);
    if (t_off(S, l, rs) + ((size_t)c + 1) * t_size(S - 1, rb) > t_size(S, rs)) return std::nan("");
    const double sub = eval_suffix_host(D, S - 1, rb, reff - (l - 1), blk + t_off(S, l, rs) + (size_t)c * t_size(S - 1, rb), x01);
    out = l == 1 ? sub : std::fma(hat_host(x, l, c), sub, out);
  }
  return out;
}

}  // namespace

static int pack_blocks_tile(int D, int64_t N, const uint8_t *lv, const uint32_t *cells, const int64_t *perm,
                            const double *coefs, bool selfcheck, int tile_coefs, PackResult &out, std::string &err);

// Try the tile sizes from the largest down: a larger tile lets more prefixes take a deep compiled suffix (fewer
// records, a shallower stack), but two buffers of it must fit the SM next to the per-point state of the walk.
int pack_blocks(int D, int64_t N, const uint8_t *lv, const uint32_t *cells, const int64_t *perm, const double *coefs,
                bool selfcheck, PackResult &out, std::string &err) {
  static const int forced = [] { const char *e = getenv("ASG_BLOCK_TILE"); return e ? atoi(e) : 0; }();  // experiment knob
  std::string first_why;
  for (int tc : kBlkTileCandidates) {
    if (forced && tc != forced) continue;
    const int rc = pack_blocks_tile(D, N, lv, cells, perm, coefs, selfcheck, tc, out, err);
    if (rc != ASG_OK) return rc;
    if (!out.blk_ok) {  // limits that do not depend on the tile size end the search
      if (first_why.empty()) first_why = out.blk_why;
      if (out.blk_why.find("tile") == std::string::npos && out.blk_why.find("padding") == std::string::npos) return ASG_OK;
      continue;
    }
    if (block_smem_bytes(D, out.blk_slots, tc, 2, kBlkThreadsWalk) <= 227u * 1024u) return ASG_OK;
    if (tc == kBlkTileCandidates[sizeof(kBlkTileCandidates) / sizeof(int) - 1] &&
        block_smem_bytes(D, out.blk_slots, tc, 1, kBlkThreadsWalk) <= 227u * 1024u)
      return ASG_OK;  // smallest tile, one point per thread
    out.blk_ok = false;
    out.blk_why = "the walk's shared memory does not fit an SM";
  }
  if (!out.blk_ok && !first_why.empty() && out.blk_why.empty()) out.blk_why = first_why;
  return ASG_OK;
}

static int pack_blocks_tile(int D, int64_t N, const uint8_t *lv, const uint32_t *cells, const int64_t *perm,
                            const double *coefs, bool selfcheck, int tile_coefs, PackResult &out, std::string &err) {
  PackTimer t_tile("  BLOCK layout, one tile size");
  out.blk_ok = false;
  out.blk_tile_coefs = tile_coefs;
  out.blk_why.clear();
  out.bcoef.clear(); out.brec.clear(); out.btiles.clear(); out.bslot_of_node.clear();
  auto no = [&](const char *w) { out.blk_why = w; return ASG_OK; };
  if (D < 3) return no("D < 3");
  if (D - 3 > 15) return no("D > 18");
  if (N <= 0) return no("empty grid");
  const int K = D - 3;  // longest prefix: dimensions 0..K-1; A = K, B = K+1, C = K+2
  const uint64_t cap = (uint64_t)tile_coefs - 2;  // a tile window is aligned down / up to even slots

  // 1. full-length prefix level vectors and the budgets of their 3-D suffixes
  std::map<Key, int> pre;  // -> max sum(l-1) over dimensions A, B, C
  {
    Key cur;
    int *pr = nullptr;
    for (int64_t k = 0; k < N; ++k) {
      const uint8_t *L = lv + (size_t)perm[k] * D;
      if (!pr || std::memcmp(cur.data(), L, (size_t)K) != 0) {
        cur.assign(L, L + K);
        pr = &pre.emplace(cur, -1).first->second;
      }
      const int eS = (L[K] - 1) + (L[K + 1] - 1) + (L[K + 2] - 1);
      if (eS > kBlkMaxBudget) return no("suffix budget > 9");
      *pr = std::max(*pr, eS);
    }
  }
  // 2. records. A group of prefix vectors that share their first d components and leaves a small budget
  //    below them becomes ONE record with a compiled suffix of S = D - d dimensions (shallowest d first);
  //    every other prefix vector is a record of its own with the 3-D suffix.
  std::map<Key, Rec> recs;         // keyed by the prefix padded with ones: lexicographic = depth-first order
  std::map<Key, Key> cover;        // full prefix vector -> key of the record that holds its nodes
  {
    std::vector<const Key *> keys;
    std::vector<int> rr;
    for (auto &kv : pre) { keys.push_back(&kv.first); rr.push_back(kv.second); }
    std::vector<char> done(keys.size(), 0);
    const int maxS = max_compiled_dims();
    for (int d = std::max(0, D - maxS); d < K; ++d) {
      const int S = D - d;
      size_t g0 = 0;
      while (g0 < keys.size()) {
        size_t g1 = g0 + 1;
        while (g1 < keys.size() && std::memcmp(keys[g1]->data(), keys[g0]->data(), (size_t)d) == 0) ++g1;
        if (!done[g0]) {  // groups nest: either every member is free or the whole group is already held
          int rsub = 0, pb = 0;
          for (size_t m = g0; m < g1; ++m) {
            int e = rr[m];
            for (int j = d; j < K; ++j) e += (*keys[m])[(size_t)j] - 1;
            rsub = std::max(rsub, e);
          }
          for (int j = 0; j < d; ++j) pb += blk_bits((*keys[g0])[(size_t)j]);
          if (rsub >= 1 && rsub <= blk_max_r(S) && pb <= 24 && (((uint64_t)t_size(S, rsub)) << pb) <= cap) {
            Key key(keys[g0]->begin(), keys[g0]->begin() + d);
            key.resize((size_t)K, 1);
            Rec R;
            R.depth = d;
            R.S = S;
            R.r = rsub;
            recs[key] = R;
            for (size_t m = g0; m < g1; ++m) { done[m] = 1; cover[*keys[m]] = key; }
          }
        }
        g0 = g1;
      }
    }
    for (size_t m = 0; m < keys.size(); ++m)
      if (!done[m]) {
        Rec R;
        R.depth = K;
        R.S = 3;
        R.r = rr[m];
        recs[*keys[m]] = R;
        cover[*keys[m]] = *keys[m];
      }
  }
  // 3. closure under "reset the last non-trivial dimension": the records a walk restarts from must exist
  {
    std::vector<Key> todo;
    for (auto &kv : recs) todo.push_back(kv.first);
    while (!todo.empty()) {
      Key p = todo.back();
      todo.pop_back();
      bool triv = true;
      for (uint8_t l : p) triv = triv && l == 1;
      if (triv) continue;
      Key q = parent_of(p);
      auto it = recs.find(q);
      if (it == recs.end()) {
        Rec R;  // stage-only record
        R.depth = K;
        it = recs.emplace(q, R).first;
        todo.push_back(q);
      }
      it->second.parent_of = true;
    }
    if (recs.find(Key((size_t)K, 1)) == recs.end()) { Rec R; R.depth = K; recs[Key((size_t)K, 1)] = R; }
  }
  // 4. geometry, storage
  uint64_t total = 0;
  int max_slot = 0;
  for (auto &kv : recs) {
    const Key &p = kv.first;
    Rec &I = kv.second;
    for (int d = 0; d < K; ++d) {
      const int l = p[(size_t)d];
      if (l > 1) { I.nt++; I.k0 = d; }
      I.pb += blk_bits(l);
      I.used += l - 1;
    }
    if (I.pb > 24) return no("prefix position needs more than 24 bits");
    if (I.r >= 0) {
      const uint64_t sz = ((uint64_t)t_size(I.S, I.r)) << I.pb;
      if (sz <= cap) {
        I.cbase = (uint32_t)total;
        total += sz;
      } else {  // only S = 3 records can be too large (deeper suffixes were admitted by size)
        I.split = true;
        for (int lA = 1; lA <= I.r + 1; ++lA) {
          const uint64_t s2 = ((uint64_t)t_size(2, I.r - (lA - 1))) << (I.pb + blk_bits(lA));
          if (s2 > cap) return no("a suffix block exceeds the shared-memory tile");
          I.cbaseA[lA - 1] = (uint32_t)total;
          total += s2;
        }
        if (I.r >= 1) I.parent_of = true;  // its own lA >= 2 records restart from it
      }
      if (total > 0x7fffffffull) return no("too many block slots");
    }
    if (I.parent_of) max_slot = std::max(max_slot, I.nt);
  }
  if (max_slot >= kBlkMaxSlots) return no("more than 7 non-trivial prefix dimensions on a path");
  if (total > 2ull * (uint64_t)N + 4096ull) return no("padding exceeds the number of nodes");
  out.blk_slots = max_slot + 1;
  {  // the stack depth is known: a tile size whose walk cannot fit an SM is given up before its storage is laid out
    const bool last = tile_coefs == kBlkTileCandidates[sizeof(kBlkTileCandidates) / sizeof(int) - 1];
    if (block_smem_bytes(D, out.blk_slots, tile_coefs, 2, kBlkThreadsWalk) > 227u * 1024u &&
        !(last && block_smem_bytes(D, out.blk_slots, tile_coefs, 1, kBlkThreadsWalk) <= 227u * 1024u))
      return no("the walk's shared memory does not fit an SM (tile size)");
  }
  out.bcoef.assign((size_t)total + 2, 0.0);
  out.bslot_of_node.assign((size_t)N, -1);
  std::vector<double> synth;  // self-check of a coefficient-free pack (asg_pack_info): use synthetic values
  if (selfcheck && !coefs) {
    synth.resize((size_t)N);
    for (int64_t n = 0; n < N; ++n) synth[(size_t)n] = std::sin(1.0 + 0.37 * (double)n) + 0.25;
    coefs = synth.data();
  }
  // 5. node slots (independent per node: a few host threads on large grids)
  {
    const int nth = pack_threads(N);
    std::vector<int> bad((size_t)nth, 0);
    pack_parallel(N, nth, [&](int th, int64_t k0, int64_t k1) {
      Key cur;
      const Rec *pi = nullptr;
      for (int64_t k = k0; k < k1; ++k) {
        const int64_t n = perm[k];
        const uint8_t *L = lv + (size_t)n * D;
        const uint32_t *C = cells + (size_t)n * D;
        if (!pi || std::memcmp(cur.data(), L, (size_t)K) != 0) {
          cur.assign(L, L + K);
          const auto it = cover.find(cur);  // (find: no insertion from concurrent readers)
          if (it == cover.end()) { bad[(size_t)th] = 1; return; }
          const auto ir = recs.find(it->second);
          if (ir == recs.end()) { bad[(size_t)th] = 1; return; }
          pi = &ir->second;
        }
        uint64_t O = 0;
        for (int d = 0; d < pi->depth; ++d) O = (O << blk_bits(L[d])) | C[d];
        uint64_t slot;
        int rr = pi->r;
        if (!pi->split) {
          slot = pi->cbase + O * t_size(pi->S, pi->r);
          for (int j = pi->S; j >= 3; --j) {  // compiled dimensions above the 2-D block, outermost first
            const int l = L[D - j];
            slot += t_off(j, l, rr) + (uint64_t)C[D - j] * t_size(j - 1, rr - (l - 1));
            rr -= l - 1;
          }
        } else {
          const int lA = L[K];
          rr -= lA - 1;
          slot = pi->cbaseA[lA - 1] + ((O << blk_bits(lA)) | C[K]) * t_size(2, rr);
        }
        const int lB = L[K + 1], lC = L[K + 2];
        slot += (uint64_t)t_off(2, lB, rr) + (uint64_t)C[K + 1] * blk_n1(rr - (lB - 1)) + blk_off1(lC) + C[K + 2];
        if (slot >= total || out.bslot_of_node[(size_t)n] != -1) { bad[(size_t)th] = 1; return; }  // duplicates were rejected before: packer bug
        out.bslot_of_node[(size_t)n] = (int64_t)slot;
        out.bcoef[(size_t)slot] = coefs ? coefs[n] : 0.0;
      }
    });
    for (int b : bad)
      if (b) {
        err = "internal error: block slot out of range or assigned twice";
        return ASG_ESTATE;
      }
  }
  // 6. records in lexicographic order + the slot span each one reads
  std::vector<uint32_t> lo, hi;
  auto stage_of = [](BlockRec &q, int slot, int k0, int l, bool save, int used) {
    q.slot = (uint32_t)slot;
    q.k0 = (uint32_t)k0;
    q.b = (uint32_t)blk_bits(l);
    q.s_hi = l > 1 ? (uint32_t)(1023 + l - 1) << 20 : 0u;  // l == 1 only for the root: s = 0 gives phi = 1
    q.lim = (1u << blk_bits(l)) - 1u;
    q.odd = l > 2 ? 1u : 0u;
    q.sw = (save ? kBrSave : 0u) | ((uint32_t)used << kBrUsedShift);
  };
  out.blk_pair_ok = true;
  out.blk_wide = false;
  for (auto &kv : recs) {
    const Key &p = kv.first;
    const Rec &I = kv.second;
    BlockRec q{};
    const int l0 = I.nt ? p[(size_t)I.k0] : 1;
    stage_of(q, I.nt ? I.nt - 1 : 0, I.k0, l0, I.parent_of && I.nt, I.used);
    // PAIR mode: neighbours of the cell-key order share every cell of level <= kBlkPairLevel, so they share the
    // prefix position of this record iff every prefix level is that low
    bool shared = true;
    for (uint8_t l : p) shared = shared && l <= kBlkPairLevel;
    if (shared) q.sw |= kBrShared;
    if (I.r < 0) {
      q.sw |= kBrNone;
      out.brec.push_back(q);
      lo.push_back(0); hi.push_back(0);
    } else if (!I.split) {
      q.sw |= blk_code(I.S, I.r);
      if (!shared && I.r > kBlkPairMaxR) out.blk_pair_ok = false;
      if (I.r > 7) out.blk_wide = true;
      q.cbase = I.cbase;
      out.brec.push_back(q);
      lo.push_back(I.cbase);
      hi.push_back(I.cbase + (t_size(I.S, I.r) << I.pb));
    } else {
      for (int lA = 1; lA <= I.r + 1; ++lA) {
        BlockRec s = q;
        const int rA = I.r - (lA - 1);
        if (lA > 1) {  // stage = level lA of dimension A, restart from the prefix product this prefix saved
          stage_of(s, I.nt, K, lA, false, I.used + lA - 1);
        }
        if (shared && lA <= kBlkPairLevel) s.sw |= kBrShared;
        else if (rA > kBlkPairMaxR) out.blk_pair_ok = false;
        if (rA > 7) out.blk_wide = true;
        s.sw |= blk_code(2, rA);
        s.cbase = I.cbaseA[lA - 1];
        out.brec.push_back(s);
        lo.push_back(s.cbase);
        hi.push_back(s.cbase + (t_size(2, rA) << (I.pb + blk_bits(lA))));
      }
    }
  }
  // 7. tiles: consecutive records whose spans fit one coefficient buffer
  {
    size_t j = 0;
    const size_t nrec = out.brec.size();
    while (j < nrec) {
      BlockTile t{(uint32_t)j, (uint32_t)j, 0u, 0u};
      bool have = false;
      while (j < nrec && (int)(j - t.r0) < kBlkTileRecs) {
        if (hi[j] > lo[j]) {
          const uint32_t a = lo[j] & ~1u, b = (hi[j] + 1u) & ~1u;
          if (!have) {
            t.c0 = a; t.c1 = b; have = true;
          } else {
            const uint32_t na = std::min(t.c0, a), nb = std::max(t.c1, b);
            if (nb - na > (uint32_t)tile_coefs) break;
            t.c0 = na; t.c1 = nb;
          }
        }
        ++j;
      }
      t.r1 = (uint32_t)j;
      out.btiles.push_back(t);
    }
    for (const BlockTile &t : out.btiles) {
      if (t.c1 - t.c0 > (uint32_t)tile_coefs || t.r1 - t.r0 > (uint32_t)kBlkTileRecs || (t.c0 & 1u) || (t.c1 & 1u) ||
          (uint64_t)t.c1 > out.bcoef.size()) {
        err = "internal error: block tile out of bounds";
        return ASG_ESTATE;
      }
      for (uint32_t r = t.r0; r < t.r1; ++r)
        if (hi[r] > lo[r] && (lo[r] < t.c0 || hi[r] > t.c1)) {
          err = "internal error: block record outside its tile";
          return ASG_ESTATE;
        }
    }
  }
  out.blk_ok = true;
  // 8. self-check: the walk over the records must reproduce the direct sum over the nodes
  if (selfcheck) {
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      return (double)(s >> 11) * (1.0 / 9007199254740992.0);
    };
    std::vector<double> x((size_t)D);
    for (int trial = 0; trial < 6; ++trial) {
      for (int d = 0; d < D; ++d) x[(size_t)d] = trial == 0 ? 1.0 : (trial == 1 ? 0.0 : (trial == 2 ? 0.5 : rnd()));
      for (int rem : {INT32_MAX / 2, 3}) {
        long double want = 0.0L, scale = 0.0L;
        for (int64_t n = 0; n < N; ++n) {
          const uint8_t *L = lv + (size_t)n * D;
          int e = 0;
          for (int d = 0; d < D; ++d) e += L[d] - 1;
          if (e > rem) continue;
          double v = coefs ? coefs[n] : 0.0;
          for (int d = 0; d < D && v != 0.0; ++d) {
            const uint32_t c = cell_host(x[(size_t)d], L[d]);
            v *= (c == cells[(size_t)n * D + d]) ? hat_host(x[(size_t)d], L[d], c) : 0.0;
          }
          want += v;
          scale += std::fabs(v);
        }
        const double got = block_walk_host(D, out, x.data(), rem);
        if (!(std::fabs((double)(got - want)) <= 1e-13 * std::max<double>(1.0, (double)scale))) {
          err = "internal error: BLOCK walk self-check failed (got " + std::to_string(got) + ", want " +
                std::to_string((double)want) + ")";
          out.blk_ok = false;
          return ASG_ESTATE;
        }
      }
    }
  }
  if (!synth.empty()) std::fill(out.bcoef.begin(), out.bcoef.end(), 0.0);
  {  // the device form of the record stream must lead the walk through exactly these records (cheap: every pack)
    std::string why;
    if (!block_device_stream_ok(out, why)) {
      err = "internal error: BLOCK device record stream: " + why;
      out.blk_ok = false;
      return ASG_ESTATE;
    }
  }
  return ASG_OK;
}

void block_device_stream(const PackResult &pk, std::vector<BlockRec> &drec, std::vector<BlockTile> &dtiles) {
  const size_t ntl = pk.btiles.size();
  std::vector<uint32_t> at(ntl);
  uint32_t pos = 0;
  for (size_t t = 0; t < ntl; ++t) {
    at[t] = pos;
    pos += 1u + (pk.btiles[t].r1 - pk.btiles[t].r0);
  }
  drec.assign(pos, BlockRec{});
  dtiles.resize(ntl);
  for (size_t t = 0; t < ntl; ++t) {
    const BlockTile &tl = pk.btiles[t];
    const size_t n2 = t + (size_t)kBlkTileBufs;
    drec[at[t]] = blk_tile_header(tl, n2 < ntl ? &pk.btiles[n2] : nullptr, n2 < ntl ? at[n2] : 0u);
    std::copy(pk.brec.begin() + tl.r0, pk.brec.begin() + tl.r1, drec.begin() + at[t] + 1);
    dtiles[t] = BlockTile{at[t], at[t] + 1u + (tl.r1 - tl.r0), tl.c0, tl.c1};
  }
}

bool block_device_stream_ok(const PackResult &pk, std::string &why) {
  std::vector<BlockRec> drec;
  std::vector<BlockTile> dtiles;
  block_device_stream(pk, drec, dtiles);
  const size_t ntl = pk.btiles.size();
  struct Req { uint32_t r0, rbytes, c0, cc; };  // what blk_issue_tile is given
  Req buf[kBlkTileBufs] = {};
  for (int b = 0; b < kBlkTileBufs && (size_t)b < ntl; ++b)
    buf[b] = Req{dtiles[(size_t)b].r0, (dtiles[(size_t)b].r1 - dtiles[(size_t)b].r0) * (uint32_t)sizeof(BlockRec), dtiles[(size_t)b].c0,
                 dtiles[(size_t)b].c1 - dtiles[(size_t)b].c0};
  size_t next_rec = 0;  // packed records met so far
  for (size_t t = 0; t < ntl; ++t) {
    const Req q = buf[t % kBlkTileBufs];
    const BlockTile &tl = pk.btiles[t];
    if (q.rbytes % sizeof(BlockRec) || q.rbytes > (uint32_t)((kBlkTileRecs + 1) * sizeof(BlockRec)) ||
        (size_t)q.r0 + q.rbytes / sizeof(BlockRec) > drec.size() || q.cc > (uint32_t)pk.blk_tile_coefs ||
        (size_t)q.c0 + q.cc > pk.bcoef.size() || ((q.c0 | q.cc) & 1u)) {  // 16-byte granules for the bulk copies
      why = "tile " + std::to_string(t) + ": request outside the stream or larger than a buffer";
      return false;
    }
    const BlockRec &h = drec[q.r0];
    const uint32_t nrec = h.sw;
    if (h.cbase != tl.c0 || q.c0 != tl.c0 || q.cc != tl.c1 - tl.c0 || nrec != tl.r1 - tl.r0 || nrec + 1u != q.rbytes / sizeof(BlockRec)) {
      why = "tile " + std::to_string(t) + ": header does not describe the tile";
      return false;
    }
    for (uint32_t k = 0; k < nrec; ++k, ++next_rec)
      if (next_rec != (size_t)tl.r0 + k || std::memcmp(&drec[q.r0 + 1u + k], &pk.brec[next_rec], sizeof(BlockRec)) != 0) {
        why = "tile " + std::to_string(t) + ": records out of order";
        return false;
      }
    if (t + (size_t)kBlkTileBufs < ntl) buf[t % kBlkTileBufs] = Req{h.b, h.s_hi, h.lim, h.odd};
  }
  if (next_rec != pk.brec.size()) {
    why = "the tiles do not cover the record stream";
    return false;
  }
  return true;
}

// The walk of asg_kernels_block.cu, one point, plain loops: records in order, one stage per record from
// the saved slot, then the compiled suffix of budget min(r, rem - used) with the layout of asg_block.h.
double block_walk_host(int D, const PackResult &pk, const double *x01, int rem) {
  double P[kBlkMaxSlots + 2];
  uint64_t O[kBlkMaxSlots + 2];
  P[0] = 1.0;
  O[0] = 0;
  double acc = 0.0;
  for (const BlockRec &q : pk.brec) {
    const int used = (int)(q.sw >> kBrUsedShift);
    if (used > rem) continue;
    const double Pin = P[q.slot], xin = x01[q.k0];
    const uint64_t Oin = O[q.slot];
    double s;
    {
      const uint64_t bits = (uint64_t)q.s_hi << 32;
      std::memcpy(&s, &bits, 8);
    }
    const uint32_t cell = std::min((uint32_t)(xin * std::ldexp(1.0, (int)q.b)), q.lim);
    const double t = std::fma(xin, s, -(double)(2u * cell + q.odd));
    const double cur = Pin * (1.0 - std::fabs(t));
    const uint64_t cO = (Oin << q.b) | cell;
    if (q.sw & kBrSave) { P[q.slot + 1] = cur; O[q.slot + 1] = cO; }
    const uint32_t code = q.sw & 0xffu;
    if (code == kBrNone) continue;
    int S, rs;
    blk_decode(code, S, rs);
    if (S < 2) return std::nan("");
    const int reff = std::min(rs, rem - used);
    if ((uint64_t)q.cbase + (cO + 1) * t_size(S, rs) > pk.bcoef.size()) return std::nan("");  // chunk outside the storage
    const double *chunk = pk.bcoef.data() + q.cbase + cO * t_size(S, rs);
    acc = std::fma(cur, eval_suffix_host(D, S, rs, reff, chunk, x01), acc);
  }
  return acc;
}

}  // namespace asg
