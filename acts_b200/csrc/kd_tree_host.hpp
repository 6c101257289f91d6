// kd_tree_host.hpp -- host-side twin of the orthogonal seeder's k-d tree construction (the default path builds the
// trees on the device: k_kd_select / k_kd_split / k_kd_small in orthogonal_kernels.cuh; B200SEED_KD_HOST=1 selects
// this one, which the tests use as a cross-check).  Also defines the node record both builders produce.
//
// The reference builds one Acts::KDTree<3, SpacePointIndex, float, std::array, 4> per event over
// (phi, r, z) of the selected space points (OrthogonalTripletSeedingAlgorithm.cpp:113-150,
// Core/include/Acts/Utilities/KDTree.hpp:66-80,271-350) and then
//   * iterates the middles in the tree's ELEMENT ORDER (.cpp:247) and
//   * receives the candidates of every range search in that order (KDTree.hpp:352-396),
// so the element order after construction is part of the result (seed order, tie order of the
// unstable cotTheta sort).  That order is the outcome of libstdc++'s std::partition (nodes with
// more than 128 elements: split at the middle of the bounding box) and std::sort (smaller nodes:
// split at the median) applied node by node; it is produced here with the very same library
// calls, on the host, in the plugin's host layer -- construction is a sequential recursion over
// in-place permutations; the per-middle range searches and everything after them run on the device.
//
// Output layout: nodes in PRE-ORDER with a skip index (first node after the subtree) and the left
// child's index, local to the event; seeding_plugin.cu re-bases them into the batch-wide node array
// of the device (node e = root of event e, ropes instead of skip indices), where the range search is
// the stack-free scan `id = descend ? lhs[id] : rope[id]`.  This builder is the host-side twin of the
// device construction (k_kd_split / k_kd_small) and is what B200SEED_KD_HOST=1 selects.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <utility>
#include <vector>

#include "seed_math.h"

namespace B200SEED_NS {

constexpr uint32_t kKdEnd = 0xFFFFFFFFu;  // rope of the last subtree: the search is over

struct KdNodeDev {
  float mn[3], mx[3];    // bounding box (phi, r, z): [min, nextafter(max)) like KDTree::boundingBox
  uint32_t begin, end;   // element positions (host builder: inside the event; on the device: packed, batch-wide)
  uint32_t skip;         // rope: the node to continue with when this subtree is done or skipped (kKdEnd: none)
  uint32_t internal;     // 1: internal node, 0: leaf
  uint32_t lhs;          // internal nodes: left child; the right child is the left child's rope
  uint32_t pad;
};
static_assert(sizeof(KdNodeDev) == 48, "three 16-byte words per node");

struct KdEventTree {
  std::vector<uint32_t> posOrig;  // element position -> index of the space point in the caller's event
  std::vector<float> posPhi;      // element position -> phi
  std::vector<KdNodeDev> nodes;
  float rMiddleMin = 0.f, rMiddleMax = 0.f;  // useVariableMiddleSPRange (.cpp:227-232)
  bool runaway = false;  // construction does not terminate (> 128 identical points: the reference recurses without end)
};

inline void build_kd_event(const DeviceConfig& cfg, uint32_t n, const float* x, const float* y, const float* z,
                           const float* r, KdEventTree& out) {
  using Coord = std::array<float, 3>;
  using Elem = std::pair<Coord, uint32_t>;
  constexpr std::size_t kLeaf = 4, kExactMedian = 128;
  std::vector<Elem> elems;
  elems.reserve(n);
  double rLo = 0., rHi = 0.;
  bool any = false;
  for (uint32_t i = 0; i < n; ++i) {
    if (cfg.useExtraCuts && !itk_sp_select(r[i], z[i])) continue;  // .cpp:123-127
    const float phi = std::atan2(y[i], x[i]);                       // .cpp:136: std::atan2(float, float)
    elems.push_back({Coord{phi, r[i], z[i]}, i});
    const double xd = x[i], yd = y[i];
    const double perp = std::sqrt(xd * xd + yd * yd);  // Extent::extend, AxisR = VectorHelpers::perp
    if (!any) { rLo = rHi = perp; any = true; }
    rLo = std::min(rLo, perp);
    rHi = std::max(rHi, perp);
  }
  out.rMiddleMin = static_cast<float>(std::floor(rLo / 2) * 2 + cfg.deltaRMiddleMinSPRange);
  out.rMiddleMax = static_cast<float>(std::floor(rHi / 2) * 2 - cfg.deltaRMiddleMaxSPRange);
  out.nodes.clear();
  if (!elems.empty()) {
    struct Frame {
      std::size_t b, e;
      bool internal;
      uint32_t dim;
    };
    std::vector<Frame> todo;
    todo.push_back({0, elems.size(), elems.size() > kLeaf, 0});
    while (!todo.empty()) {
      if (out.nodes.size() > 4 * elems.size() + 64) {  // a binary tree over n elements with non-empty leaves has < 2 n nodes
        out.runaway = true;
        break;
      }
      const Frame f = todo.back();
      todo.pop_back();
      KdNodeDev node{};
      for (int j = 0; j < 3; ++j) {
        node.mn[j] = std::numeric_limits<float>::max();
        node.mx[j] = std::numeric_limits<float>::lowest();
      }
      for (std::size_t i = f.b; i != f.e; ++i) {
        for (int j = 0; j < 3; ++j) {
          node.mn[j] = std::min(node.mn[j], elems[i].first[j]);
          node.mx[j] = std::max(node.mx[j], elems[i].first[j]);
        }
      }
      for (int j = 0; j < 3; ++j) node.mx[j] = std::nextafter(node.mx[j], std::numeric_limits<float>::max());
      node.begin = static_cast<uint32_t>(f.b);
      node.end = static_cast<uint32_t>(f.e);
      node.internal = f.internal ? 1u : 0u;
      out.nodes.push_back(node);
      if (!f.internal) continue;
      const auto first = elems.begin() + static_cast<std::ptrdiff_t>(f.b);
      const auto last = elems.begin() + static_cast<std::ptrdiff_t>(f.e);
      const uint32_t d = f.dim;
      auto pivot = first;
      if (f.e - f.b > kExactMedian) {
        const float mid = 0.5f * (node.mx[d] + node.mn[d]);
        pivot = std::partition(first, last, [=](const Elem& v) { return v.first[d] < mid; });
      } else {
        std::sort(first, last, [d](const Elem& a, const Elem& b) { return a.first[d] < b.first[d]; });
        pivot = first + (last - first) / 2;
      }
      if (pivot == first || pivot == std::prev(last)) pivot = first + static_cast<std::ptrdiff_t>(kLeaf);
      const std::size_t p = static_cast<std::size_t>(pivot - elems.begin());
      // pre-order: the left child is the next node, the right child follows the left subtree
      todo.push_back({p, f.e, f.e - p > kLeaf, (d + 1) % 3});
      todo.push_back({f.b, p, p - f.b > kLeaf, (d + 1) % 3});
    }
    if (out.runaway) return;
    // subtree sizes by a reverse sweep (children have larger indices than their parent)
    const std::size_t nn = out.nodes.size();
    std::vector<uint32_t> size(nn, 1);
    for (std::size_t k = nn; k-- > 0;) {
      if (out.nodes[k].internal != 0u) {
        const std::size_t lhs = k + 1, rhs = lhs + size[lhs];
        size[k] = 1 + size[lhs] + size[rhs];
      }
      out.nodes[k].skip = static_cast<uint32_t>(k) + size[k];  // (== nodes.size(): end of the event's list)
      out.nodes[k].lhs = static_cast<uint32_t>(k) + 1;
    }
  }
  out.posOrig.resize(elems.size());
  out.posPhi.resize(elems.size());
  for (std::size_t i = 0; i < elems.size(); ++i) {
    out.posOrig[i] = elems[i].second;
    out.posPhi[i] = elems[i].first[0];
  }
}

}  // namespace B200SEED_NS
