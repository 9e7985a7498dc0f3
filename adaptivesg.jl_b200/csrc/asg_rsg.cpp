// asg_rsg.cpp -- rsg! node enumeration (src/train.jl:40-78) as host C++: the regular sparse grid node set in the
// reference's order. Level vectors in Iterators.product(1:maxlevels_1, ..., 1:maxlevels_D) order (dimension 1
// fastest) filtered by sum(ks) <= maxdepth + D - 1 (src/train.jl:56-63); inside a level vector the index tuples in
// Iterators.product(allindices_of_level(k_1), ...) order (src/math.jl:651-661), dimension 1 fastest. Integer host
// logic that the reference keeps on the host; no device is needed.
#include <functional>
#include <vector>

#include "asg_internal.h"

namespace asg {

namespace {

inline int64_t nodes_of_level(int64_t l) { return l == 1 ? 1 : (l == 2 ? 2 : (int64_t)1 << (l - 2)); }
inline int64_t index_of(int64_t l, int64_t k) { return l == 1 ? 1 : (l == 2 ? 2 * k : 2 * k + 1); }  // allindices_of_level(l)[k]

// calls emit(levels) for every admissible level vector, dimension 0 fastest, with budget pruning
void for_each_level_vector(int D, int64_t budget, const int64_t *ml, const std::function<void(const std::vector<int64_t> &)> &emit) {
  std::vector<int64_t> l((size_t)D, 1);
  std::function<void(int, int64_t)> rec = [&](int d, int64_t rem) {
    if (d < 0) { emit(l); return; }
    const int64_t top = std::min<int64_t>(ml[d], rem + 1);
    for (int64_t lv = 1; lv <= top; ++lv) {
      l[(size_t)d] = lv;
      rec(d - 1, rem - (lv - 1));
    }
    l[(size_t)d] = 1;
  };
  rec(D - 1, budget);
}

}  // namespace

int rsg_count(int D, int64_t maxdepth, const int64_t *maxlevels, int64_t *N, std::string &err) {
  for (int d = 0; d < D; ++d)
    if (maxlevels[d] < 1 || maxlevels[d] > 62) { err = "maxlevels must be in 1..62"; return ASG_EINVAL; }
  if (maxdepth < 1) { err = "maxdepth must be >= 1"; return ASG_EINVAL; }
  int64_t n = 0;
  bool overflow = false;
  for_each_level_vector(D, maxdepth - 1, maxlevels, [&](const std::vector<int64_t> &l) {
    int64_t c = 1;
    for (int d = 0; d < D; ++d) {
      const int64_t m = nodes_of_level(l[(size_t)d]);
      if (c > ((int64_t)1 << 40) / m) overflow = true;
      c *= m;
    }
    n += c;
    if (n > ((int64_t)1 << 40)) overflow = true;
  });
  if (overflow) { err = "regular sparse grid has more than 2^40 nodes"; return ASG_EUNSUPPORTED; }
  *N = n;
  return ASG_OK;
}

int rsg_fill(int D, int64_t maxdepth, const int64_t *maxlevels, int64_t N, int64_t *levels, int64_t *indices, int64_t sn,
             int64_t sd, std::string &err) {
  int64_t row = 0;
  bool overrun = false;
  std::vector<int64_t> k((size_t)D), cnt((size_t)D);
  for_each_level_vector(D, maxdepth - 1, maxlevels, [&](const std::vector<int64_t> &l) {
    int64_t total = 1;
    for (int d = 0; d < D; ++d) {
      cnt[(size_t)d] = nodes_of_level(l[(size_t)d]);
      total *= cnt[(size_t)d];
      k[(size_t)d] = 0;
    }
    if (row + total > N) { overrun = true; return; }
    for (int64_t t = 0; t < total; ++t, ++row) {
      for (int d = 0; d < D; ++d) {
        levels[row * sn + d * sd] = l[(size_t)d];
        indices[row * sn + d * sd] = index_of(l[(size_t)d], k[(size_t)d]);
      }
      for (int d = 0; d < D; ++d) {  // odometer, dimension 0 fastest
        if (++k[(size_t)d] < cnt[(size_t)d]) break;
        k[(size_t)d] = 0;
      }
    }
  });
  if (overrun || row != N) { err = "N differs from the node count of this regular sparse grid"; return ASG_EINVAL; }
  return ASG_OK;
}

}  // namespace asg
