// asg_persist.cpp -- device-format persistence of a packed grid: the caller's node arrays
// (src/io.jl:30-38 schema: levels, indices, coefs) together with every packed layout the kernels read (sweep pairs,
// subspace trie / blocks / hash tables, flat group stream, BLOCK records and tiles, slot maps) as one relocatable
// blob, so that loading a trained grid skips validation, sorting and packing (0.5 s for the 652 065-node benchmark
// grid) and goes straight to the device upload. The blob is tied to the layout constants of this build (fingerprint)
// and protected against accidental corruption by a 64-bit checksum; it is trusted like any model file otherwise.
#include <cstring>

#include "asg_internal.h"

namespace asg {

namespace {

constexpr char kMagic[8] = {'A', 'S', 'G', 'P', 'A', 'C', 'K', 1};

uint64_t fnv1a(const unsigned char *p, size_t n, uint64_t h = 0xcbf29ce484222325ull) {
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
  return h;
}
// checksum of the (hundreds of MB) payload: four interleaved multiply-xor lanes over 8-byte words, FNV for the tail
uint64_t checksum(const unsigned char *p, size_t n) {
  uint64_t h[4] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0x27D4EB2F165667C5ull};
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    uint64_t w[4];
    std::memcpy(w, p + i, 32);
    for (int k = 0; k < 4; ++k) { h[k] = (h[k] ^ w[k]) * 0x100000001b3ull; h[k] ^= h[k] >> 29; }
  }
  uint64_t r = fnv1a(p + i, n - i);
  for (int k = 0; k < 4; ++k) r = (r ^ h[k]) * 0xFF51AFD7ED558CCDull;
  return r ^ (r >> 32);
}

uint64_t layout_fingerprint() {  // everything a blob's bytes depend on besides the grid itself
  const uint32_t c[] = {5u /* format version */, kBlkCodes, (uint32_t)kBlkMaxBudget, (uint32_t)kBlkPairLevel, (uint32_t)kBlkPairMaxR, (uint32_t)sizeof(TrieRec), (uint32_t)sizeof(GroupRec), (uint32_t)sizeof(BlockRec),
                        (uint32_t)sizeof(StreamTile), (uint32_t)sizeof(BlockTile), (uint32_t)sizeof(HashEnt),
                        (uint32_t)sizeof(SweepDim), (uint32_t)sizeof(RsgParams), (uint32_t)kLeafTable, (uint32_t)kTileCoefs,
                        (uint32_t)kTileRecs, (uint32_t)kBlkTileCoefsMax, (uint32_t)kBlkTileRecs, (uint32_t)kBlkMaxS,
                        (uint32_t)blk_max_r(4), (uint32_t)blk_max_r(5), (uint32_t)blk_max_r(6), (uint32_t)blk_max_r(7),
                        (uint32_t)blk_max_r(8), (uint32_t)kKeyBits, (uint32_t)kMaxDimFast, (uint32_t)kBlkMaxSlots};
  return fnv1a(reinterpret_cast<const unsigned char *>(c), sizeof c);
}

struct Writer {
  std::vector<unsigned char> &out;
  template <class T> void pod(const T &v) {
    const unsigned char *p = reinterpret_cast<const unsigned char *>(&v);
    out.insert(out.end(), p, p + sizeof(T));
  }
  template <class T, class A> void vec(const std::vector<T, A> &v) {
    pod<uint64_t>((uint64_t)v.size());
    const unsigned char *p = reinterpret_cast<const unsigned char *>(v.data());
    out.insert(out.end(), p, p + v.size() * sizeof(T));
  }
  void str(const std::string &s) {
    pod<uint64_t>((uint64_t)s.size());
    out.insert(out.end(), s.begin(), s.end());
  }
};

struct Reader {
  const unsigned char *p, *end;
  bool ok = true;
  template <class T> void pod(T &v) {
    if (!ok || (size_t)(end - p) < sizeof(T)) { ok = false; return; }
    std::memcpy(&v, p, sizeof(T));
    p += sizeof(T);
  }
  template <class T, class A> void vec(std::vector<T, A> &v) {
    uint64_t n = 0;
    pod(n);
    if (!ok || n > (uint64_t)(end - p) / sizeof(T)) { ok = false; return; }
    v.resize((size_t)n);
    std::memcpy(v.data(), p, (size_t)n * sizeof(T));
    p += (size_t)n * sizeof(T);
  }
  void str(std::string &s) {
    uint64_t n = 0;
    pod(n);
    if (!ok || n > (uint64_t)(end - p)) { ok = false; return; }
    s.assign(reinterpret_cast<const char *>(p), (size_t)n);
    p += (size_t)n;
  }
};

template <class IO, class PK>
void fields(IO &io, PK &pk) {  // one list for both directions
  io.pod(pk.canonical); io.str(pk.why); io.pod(pk.max_level); io.pod(pk.max_depth); io.pod(pk.subspaces);
  io.pod(pk.dense_subspaces); io.pod(pk.trie_nodes); io.pod(pk.full_groups);
  io.vec(pk.coef); io.pod(pk.coef_pad); io.vec(pk.orig); io.vec(pk.hent); io.vec(pk.horig); io.vec(pk.slot_of_node);
  io.vec(pk.grp); io.vec(pk.tiles); io.pod(pk.regular); io.pod(pk.rp);
  io.pod(pk.blk_ok); io.str(pk.blk_why); io.vec(pk.bcoef); io.vec(pk.brec); io.vec(pk.btiles); io.vec(pk.bslot_of_node);
  io.pod(pk.blk_slots); io.pod(pk.blk_tile_coefs); io.pod(pk.blk_pair_ok); io.pod(pk.blk_wide); io.vec(pk.sw); io.vec(pk.sw_low); io.vec(pk.depth);
}

}  // namespace

void pack_serialize(int D, const double *lb, const double *ub, const std::vector<int64_t> &levels,
                    const std::vector<int64_t> &indices, const std::vector<double> &coefs, const PackResult &pk,
                    std::vector<unsigned char> &out) {
  out.clear();
  Writer w{out};
  out.insert(out.end(), kMagic, kMagic + 8);
  w.pod<uint64_t>(layout_fingerprint());
  w.pod<int32_t>(D);
  for (int d = 0; d < D; ++d) { w.pod(lb[d]); w.pod(ub[d]); }
  w.vec(levels); w.vec(indices); w.vec(coefs);
  w.pod<uint64_t>((uint64_t)pk.rec.size());
  for (const auto &r : pk.rec) w.vec(r);
  fields(w, pk);
  w.pod<uint64_t>(checksum(out.data(), out.size()));
}

int pack_deserialize(const unsigned char *buf, size_t bytes, int D, const double *lb, const double *ub,
                     std::vector<int64_t> &levels, std::vector<int64_t> &indices, std::vector<double> &coefs, PackResult &pk,
                     std::string &err) {
  if (bytes < 8 + 8 + 4 + 8 || std::memcmp(buf, kMagic, 8) != 0) { err = "not a packed-grid blob"; return ASG_EINVAL; }
  uint64_t sum = 0;
  std::memcpy(&sum, buf + bytes - 8, 8);
  if (sum != checksum(buf, bytes - 8)) { err = "packed-grid blob is corrupted (checksum)"; return ASG_EINVAL; }
  Reader r{buf + 8, buf + bytes - 8};
  uint64_t fp = 0;
  int32_t d32 = 0;
  r.pod(fp);
  if (fp != layout_fingerprint()) { err = "packed-grid blob was written by a build with other layout constants"; return ASG_EUNSUPPORTED; }
  r.pod(d32);
  if (!r.ok || d32 != D) { err = "packed-grid blob has another dimension than the handle"; return ASG_EINVAL; }
  for (int d = 0; d < D; ++d) {
    double a = 0, b = 0;
    r.pod(a); r.pod(b);
    if (!r.ok || a != lb[d] || b != ub[d]) { err = "packed-grid blob has another domain than the handle"; return ASG_EINVAL; }
  }
  pk = PackResult();
  r.vec(levels); r.vec(indices); r.vec(coefs);
  uint64_t nlev = 0;
  r.pod(nlev);
  if (!r.ok || nlev > (uint64_t)kMaxDimFast) { err = "packed-grid blob is malformed"; return ASG_EINVAL; }
  pk.rec.resize((size_t)nlev);
  for (auto &v : pk.rec) r.vec(v);
  fields(r, pk);
  const size_t N = coefs.size();
  if (!r.ok || r.p != r.end || levels.size() != N * (size_t)D || indices.size() != N * (size_t)D || pk.sw.size() != N * (size_t)D ||
      pk.sw_low.size() != N || pk.depth.size() != N || pk.slot_of_node.size() != N ||
      (pk.canonical && pk.rec.size() != (size_t)D) || (pk.blk_ok && pk.bslot_of_node.size() != N)) {
    err = "packed-grid blob is malformed";
    return ASG_EINVAL;
  }
  // The slot maps, tiles and records become device indices (coefficient scatters, tile copies, record fetches): a
  // damaged or crafted file must fail here, not as an out-of-bounds access on the device. (The checksum only guards
  // against accidents.)
  auto bad = [&](const char *what) { err = std::string("packed-grid blob is inconsistent: ") + what; return ASG_EINVAL; };
  if (pk.canonical) {
    for (int64_t s : pk.slot_of_node)
      if (s >= 0 ? (uint64_t)s >= pk.coef.size() : (uint64_t)(-s - 1) >= pk.hent.size()) return bad("node slot out of range");
    if (pk.orig.size() != pk.coef.size() || pk.horig.size() != pk.hent.size()) return bad("slot maps differ in length");
    for (int32_t o : pk.orig) if (o < -1 || (uint64_t)(o + 1) > N) return bad("slot owner out of range");
    for (int32_t o : pk.horig) if (o < -1 || (uint64_t)(o + 1) > N) return bad("hash slot owner out of range");
    for (const StreamTile &t : pk.tiles)
      if (t.r0 > t.r1 || t.r1 > pk.grp.size() || t.c0 > t.c1 || t.c1 > pk.coef.size() || t.c1 - t.c0 > (uint32_t)kTileCoefs + 2u ||
          t.r1 - t.r0 > (uint32_t)kTileRecs)
        return bad("stream tile out of range");
  }
  if (pk.blk_ok) {
    bool tile_ok = false;
    for (int tc : kBlkTileCandidates) tile_ok = tile_ok || tc == pk.blk_tile_coefs;
    if (!tile_ok || pk.blk_slots < 1 || pk.blk_slots > kBlkMaxSlots) return bad("BLOCK walk parameters");
    for (int64_t s : pk.bslot_of_node)
      if (s < 0 || (uint64_t)s >= pk.bcoef.size()) return bad("BLOCK node slot out of range");
    for (const BlockTile &t : pk.btiles)
      if (t.r0 > t.r1 || t.r1 > pk.brec.size() || t.c0 > t.c1 || t.c1 > pk.bcoef.size() || ((t.c0 | t.c1) & 1u) ||
          t.c1 - t.c0 > (uint32_t)pk.blk_tile_coefs || t.r1 - t.r0 > (uint32_t)kBlkTileRecs)
        return bad("BLOCK tile out of range");
    for (size_t ti = 0; ti < pk.btiles.size(); ++ti) {
      const BlockTile &t = pk.btiles[ti];
      for (uint32_t q = t.r0; q < t.r1; ++q) {
        const BlockRec &b = pk.brec[q];
        const uint32_t code = b.sw & 0xffu;
        const bool save = (b.sw & kBrSave) != 0;
        if (b.slot + (save ? 1u : 0u) >= (uint32_t)pk.blk_slots || b.k0 >= (uint32_t)D || b.b > 24u || b.lim != (1u << b.b) - 1u)
          return bad("BLOCK record stage");
        if (code != kBrNone) {
          int S, r;
          blk_decode(code, S, r);
          if (S < 2 || S > kBlkMaxS || (S > 3 && r > blk_max_r(S)) || b.cbase < t.c0 || b.cbase >= t.c1) return bad("BLOCK record suffix");
        }
      }
    }
  }
  return ASG_OK;
}

}  // namespace asg
