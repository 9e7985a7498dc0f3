// asg_block.h -- BLOCK layout of the subspace engine (asg_kernels_block.cu, asg_pack.cpp).
//
// The evaluation  G(x) = sum_n c_n prod_d phi(x01_d; l_nd, i_nd)  (src/class.jl:629-643) visits one
// supporting node per subspace (level vector). The level vectors form a trie over the dimensions; the
// cost of a walk is (nodes of the trie at depth k) x (instructions per node at depth k), and on
// regular grids the three deepest dimensions hold 95 % of the trie. The BLOCK layout therefore cuts the
// dimensions into a PREFIX (dimensions 0..D-4, data driven, one 16-byte record per prefix level
// vector) and a SUFFIX (dimensions A = D-3, B = D-2, C = D-1) that is a complete regular sparse grid
// of budget r = max sum(l-1) under its prefix. The suffix walk is compiled: the kernel is templated on
// r, every level of A, B, C is a compile-time constant, all slot offsets below are immediates and the
// only warp-uniform bookkeeping left is one record fetch and one switch per prefix vector.
//
// Storage of one record (prefix level vector p with pb position bits, suffix budget r):
//   2^pb chunks of size3(r) doubles, chunk O = prefix position (cells of dimensions 0..D-4
//   concatenated, dimension 0 most significant). Inside a chunk, recursively:
//     for lA = 1..r+1:  2^b(lA) cells of A, each a 2-D block of size2(r - eA) doubles;     e = l - 1
//       for lB = 1..r-eA+1:  2^b(lB) cells of B, each a POLE of n1(r - eA - eB) doubles;
//         the pole holds the 1-D hierarchy of dimension C: level 1 at 0, level l >= 2 at
//         off1(l) + cell  (off1 = 1, 3, 5, 9, 17, 33, 65)
//   slot = cbase + O*size3(r) + offA(lA, r) + cA*size2(r-eA) + offB(lB, r-eA) + cB*n1(r-eA-eB) + off1(lC) + cC
// (S = 4, 5, 6 records, see blk_max_r below, continue the same recursion upwards: dimension D-S outermost.)
// Nodes that the grid does not hold (level caps, adapt! output) are stored as 0.0; a grid takes the
// BLOCK layout only while that padding stays below the number of real nodes.
//
// A record whose chunks exceed a shared-memory tile is split along A into S2 records, one per level lA
// (lA is applied as a data-driven stage, the 2-D suffix is compiled). Each S2 record has its own storage:
// 2^(pb + b(lA)) 2-D blocks of size2(r - eA) doubles, block number = (O << b(lA)) | cA.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ASG_HD __host__ __device__
#else
#define ASG_HD
#endif

namespace asg {

constexpr int kBlkMaxBudget = 9;     // suffix budget r <= 9  <=>  suffix levels <= 10 (levels above 4 are formed on the fly)
// doubles per shared-memory coefficient tile: the packer takes the largest candidate whose walk (two buffers of it,
// the records, the coordinates and the prefix stack of 1024 points) still fits the 227 KB of an SM
constexpr int kBlkTileCandidates[] = {6144, 4608, 3072, 2048, 1024};
constexpr int kBlkTileCoefsMax = 6144;
constexpr int kBlkThreadsWalk = 512;  // threads of the walk's CTA (two points each at full size)
constexpr int kBlkTileRecs = 128;    // records per tile (4 KB)
constexpr int kBlkMaxSlots = 8;      // saved prefix products: slots 0..7 (slot s = s non-trivial prefix dims)

ASG_HD constexpr int blk_bits(int l) { return l == 1 ? 0 : (l == 2 ? 1 : l - 2); }
// nodes of the 1-D hierarchy with levels 1..r+1: 1, 3, 5, 9, 17, 33, 65, 129
ASG_HD constexpr uint32_t blk_n1(int r) { return r < 0 ? 0u : (r == 0 ? 1u : (1u << r) + 1u); }
// offset of level l inside a pole: 0, 1, 3, 5, 9, ...
ASG_HD constexpr uint32_t blk_off1(int l) { return l == 1 ? 0u : blk_n1(l - 2); }
// Compiled suffix of S dimensions and budget r (S = 1: a pole). size(S, r) doubles; level l of the suffix's
// first dimension starts at off(S, l, r) and holds 2^b(l) sub-blocks of size(S-1, r-(l-1)).
ASG_HD constexpr uint32_t blk_size(int S, int r);
ASG_HD constexpr uint32_t blk_off(int S, int l, int r) {
  uint32_t s = 0;
  for (int k = 1; k < l; ++k) s += (1u << blk_bits(k)) * blk_size(S - 1, r - (k - 1));
  return s;
}
ASG_HD constexpr uint32_t blk_size(int S, int r) { return r < 0 ? 0u : (S <= 1 ? blk_n1(r) : blk_off(S, r + 2, r)); }
ASG_HD constexpr uint32_t blk_offB(int lB, int rA) { return blk_off(2, lB, rA); }  // 2-D block: start of level lB of B
ASG_HD constexpr uint32_t blk_size2(int r) { return blk_size(2, r); }              // 1, 5, 13, 29, 65, 145, 321, 705
ASG_HD constexpr uint32_t blk_offA(int lA, int r) { return blk_off(3, lA, r); }    // 3-D chunk: start of level lA of A
ASG_HD constexpr uint32_t blk_size3(int r) { return blk_size(3, r); }              // 1, 7, 25, 69, 177, 441, 1073, 2561
static_assert(blk_size2(7) == 705 && blk_size3(7) == 2561 && blk_n1(7) == 129, "suffix block sizes");
static_assert(blk_n1(9) == 513 && blk_size2(9) <= (uint32_t)kBlkTileCoefsMax, "a 2-D block of the largest budget fits a tile");
// Deeper compiled suffixes for prefixes that leave little budget: half of the prefix vectors of a regular grid
// leave r = 0 (one term), a quarter r = 1. Where a SHORTER prefix (S = 4, 5, 6 compiled dimensions) still has a
// small budget, ONE record with a compiled S-dimensional suffix replaces all the records below it.
#ifndef ASG_BLK_R4
#define ASG_BLK_R4 5  // largest budget compiled for a 4-D suffix (126 terms)
#define ASG_BLK_R5 4  // 5-D (126 terms)
#define ASG_BLK_R6 3  // 6-D (84 terms)
#define ASG_BLK_R7 2  // 7-D (36 terms)
#endif
#ifndef ASG_BLK_R8
#define ASG_BLK_R8 0
#endif
constexpr int kBlkMaxS = ASG_BLK_R8 > 0 ? 8 : 7;
ASG_HD constexpr int blk_max_r(int S) {
  return S == 3 ? kBlkMaxBudget
                : (S == 4 ? ASG_BLK_R4 : (S == 5 ? ASG_BLK_R5 : (S == 6 ? ASG_BLK_R6 : (S == 7 ? ASG_BLK_R7 : (S == 8 ? ASG_BLK_R8 : -1)))));
}

// One 32-byte record per prefix level vector, in lexicographic (depth-first) order, stored pre-decoded
// (one field per word: the kernel does no shifting or masking on its critical path). Every record but
// the first applies exactly ONE stage: the level of its last non-trivial prefix dimension k0; its input
// is the prefix product / position saved by the record that has level 1 there (stack slot = number of
// non-trivial dimensions before k0).
struct __align__(16) BlockRec {
  uint32_t cbase;  // first coefficient slot of the record (chunk 0)
  uint32_t sw;     // bits 0..7: blk_code(S, r) | kBrNone (compiled suffix to run); 8..15: flags; 16..31: used
  uint32_t slot;   // stack slot to restart from
  uint32_t k0;     // stage dimension
  uint32_t b;      // stage: cell bits of the level
  uint32_t s_hi;   // stage: high word of the double 2^(l-1); 0 for the root record (phi = 1)
  uint32_t lim;    // stage: 2^b - 1
  uint32_t odd;    // stage: 1 if level > 2 (index = 2*cell + 1)
};
static_assert(sizeof(BlockRec) == 32, "BlockRec is fetched as two 16-byte words");
// suffix code of a record: a DENSE numbering of the compiled (S, r) pairs. S = 3 the 3-D suffix (r = 0..kBlkMaxBudget);
// S = 2 the 2-D suffix of a split record (dimension A is the record's stage or at level 1), numbered after them;
// S = 4.. deeper suffixes with r = 1..blk_max_r(S), numbered on from 2 * (kBlkMaxBudget + 1).
ASG_HD constexpr uint32_t blk_code_base(int S) {
  uint32_t b = 2u * (kBlkMaxBudget + 1);
  for (int s = 4; s < S; ++s) b += (uint32_t)(blk_max_r(s) > 0 ? blk_max_r(s) : 0);
  return b;
}
ASG_HD constexpr uint32_t blk_code(int S, int r) {
  return S == 3 ? (uint32_t)r : (S == 2 ? (uint32_t)(kBlkMaxBudget + 1 + r) : blk_code_base(S) + (uint32_t)(r - 1));
}
constexpr uint32_t kBlkCodes = blk_code_base(kBlkMaxS + 1);  // number of codes
ASG_HD inline void blk_decode(uint32_t code, int &S, int &r) {  // host side: packer self-check, persistence
  if (code <= (uint32_t)kBlkMaxBudget) { S = 3; r = (int)code; return; }
  if (code < 2u * (kBlkMaxBudget + 1)) { S = 2; r = (int)code - (kBlkMaxBudget + 1); return; }
  for (int s = 4; s <= kBlkMaxS; ++s)
    if (code < blk_code_base(s + 1)) { S = s; r = (int)(code - blk_code_base(s)) + 1; return; }
  S = -1;
  r = -1;
}
constexpr uint32_t kBrNone = 255u;  // stage only (no nodes under this prefix)
constexpr uint32_t kBrSave = 1u << 8;    // save the result in slot + 1 (a later record restarts from it)
// PAIR mode of the walk (asg_kernels_block.cuh): two neighbours of the cell-key order share the cell of every level
// <= kBlkPairLevel in every dimension. A record whose prefix stages all have such levels is flagged kBrShared: the
// pair has ONE prefix position, so its chunk is read through one address. A grid takes PAIR mode when every other
// record leaves a suffix budget <= kBlkPairMaxR (always true while total budgets are <= 7: a level above 4 spends >= 4).
constexpr int kBlkPairLevel = 4;
constexpr int kBlkPairMaxR = 3;
constexpr uint32_t kBrShared = 1u << 9;
constexpr int kBrUsedShift = 16;         // used = sum(l-1) over the prefix (S2: including dimension A): depth mask

struct BlockTile {  // records [r0, r1) read coefficient slots inside [c0, c1)
  uint32_t r0, r1, c0, c1;
};
// DEVICE copy of the record stream: every tile's records are preceded by a HEADER record, so that the kernel learns
// a tile's extent -- and the extent of the tile after next, which it has to request -- from the shared-memory copy of
// the tile itself instead of a table in global memory. `at[t]` is the device index of tile t's header; the device tile
// table counts the header in [r0, r1).
constexpr int kBlkTileBufs = 2;  // tile buffers of the walk: tile t + kBlkTileBufs takes the buffer of tile t
inline BlockRec blk_tile_header(const BlockTile &t, const BlockTile *next2, uint32_t next2_at) {
  BlockRec h{};
  h.cbase = t.c0;        // first coefficient slot of the tile
  h.sw = t.r1 - t.r0;    // records behind the header
  if (next2) {
    h.b = next2_at;                                                  // device index of the header of tile t + 2
    h.s_hi = (1u + next2->r1 - next2->r0) * (uint32_t)sizeof(BlockRec);  // its bytes, header included
    h.lim = next2->c0;                                               // its coefficient slots
    h.odd = next2->c1 - next2->c0;
  }
  return h;
}

}  // namespace asg
