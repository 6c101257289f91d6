// gpu_model.cpp -- sequential CPU model of the GPU algorithm (TEST INFRASTRUCTURE).
//
// The CUDA kernels do not run the reference's loops literally: r windows are
// found by binary search, doublets are ordered by (cotTheta, emission index),
// every bottom doublet derives its top window independently (prefix-max
// formulation) and heap pushes are replayed from a
// compacted candidate list.  This file executes exactly that restructured
// algorithm, single threaded, with the SAME arithmetic (acts_b200/csrc/seed_math.h)
// so that its equivalence with the oracle can be tested on a machine without a
// GPU.  It is never part of the product library.
#include "../../acts_b200/csrc/seed_math.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <utility>
#include <vector>

using namespace b200seed;

namespace {

struct Packed {
  const uint32_t* copiedFrom;
  const float *x, *y, *z, *r, *varZ, *varR;
};

struct SortItem {
  float key;
  uint32_t val;
};
inline bool sortItemLess(const SortItem& a, const SortItem& b) { return a.key < b.key; }

struct SeedOut {
  uint32_t b, m, t;
  float q, z;
};

// first position in [lo, hi) where pred is true (pred monotone false..true)
template <typename Pred>
uint32_t firstTrue(uint32_t lo, uint32_t hi, Pred pred) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (pred(mid)) hi = mid; else lo = mid + 1;
  }
  return lo;
}

struct Window {
  uint32_t begin, end;
};

struct Doublet {
  DoubletRec rec;
  uint32_t pos;  // packed position of the other space point
  uint32_t seq;  // emission index in the reference's order
};

// seedConfirmation: what k_seed_middles<kConf> leaves in HBM for one middle
struct ConfRecord {
  uint32_t b, t;  // packed positions
  float weight, zOrigin;
  uint32_t group;      // sorted rank of the bottom
  uint32_t groupSize;  // min(#candidates of the group, 3)
  bool needsTwoTops, quality;
};
struct ConfMiddle {
  uint32_t m;
  std::vector<ConfRecord> rec;
};

// Strip triplet path (k_seed_middles<..., kStrip>): set by model_set_strips, NULL = pixel path
static const float* gStrip = nullptr;  // 12 floats per ORIGINAL space point
static float gCotThetaDiffMax2 = 0.f, gToleranceParam = 1.1f;

// One middle space point, restructured algorithm.  tieMode: 0 canonical
// (key, seq) order, 1 replay libstdc++ std::sort from the emission order.
void processMiddle(const DeviceConfig& c, const Packed& p, uint32_t m, uint32_t firstMiddleInBin,
                   const std::vector<Window>& bottomBins, const std::vector<Window>& topBins,
                   const float* zLo, const float* zHi, int nZWin, int tieMode,
                   std::vector<SeedOut>& out, uint64_t* stats, std::vector<ConfMiddle>* conf = nullptr) {
  MiddleSp mid{p.x[m], p.y[m], p.z[m], p.r[m], p.varZ[m], p.varR[m], 0, 0, 0, 0};
  middle_info(mid);
  const float firstMiddleR = p.r[firstMiddleInBin];

  auto collect = [&](bool bottom, const std::vector<Window>& bins, std::vector<Doublet>& dst) {
    uint32_t seq = 0;
    for (const Window& bin : bins) {
      uint32_t start, end;
      if (bottom) {
        // TripletSeeder.cpp:157-170 pre-trim + DoubletSeedFinder.cpp:73-93,105-113
        const float trimValue = fsub(firstMiddleR, c.dRMaxB);
        const uint32_t trim = firstTrue(bin.begin, bin.end, [&](uint32_t i) { return !(p.r[i] < trimValue); });
        start = firstTrue(trim, bin.end, [&](uint32_t i) { return fsub(mid.r, p.r[i]) <= c.dRMaxB; });
        end = firstTrue(start, bin.end, [&](uint32_t i) { return fsub(mid.r, p.r[i]) < c.dRMinB; });
      } else {
        const float trimValue = fadd(firstMiddleR, c.dRMinT);
        const uint32_t trim = firstTrue(bin.begin, bin.end, [&](uint32_t i) { return !(p.r[i] < trimValue); });
        start = firstTrue(trim, bin.end, [&](uint32_t i) { return fsub(p.r[i], mid.r) >= c.dRMinT; });
        end = firstTrue(start, bin.end, [&](uint32_t i) { return fsub(p.r[i], mid.r) > c.dRMaxT; });
      }
      for (uint32_t o = start; o < end; ++o, ++seq) {
        float dR, dZ;
        const bool ok = bottom ? doublet_zr_cuts<true>(c, mid, p.z[o], p.r[o], dR, dZ)
                               : doublet_zr_cuts<false>(c, mid, p.z[o], p.r[o], dR, dZ);
        if (!ok) continue;
        Doublet d;
        const bool ok2 = bottom ? doublet_finish<true>(c, mid, dR, dZ, p.x[o], p.y[o], p.r[o], p.varZ[o], p.varR[o], zLo, zHi, nZWin, d.rec)
                                : doublet_finish<false>(c, mid, dR, dZ, p.x[o], p.y[o], p.r[o], p.varZ[o], p.varR[o], zLo, zHi, nZWin, d.rec);
        if (!ok2) continue;
        d.pos = o;
        d.seq = seq;
        dst.push_back(d);
      }
    }
  };

  std::vector<Doublet> tops, bottoms;
  collect(false, topBins, tops);
  if (tops.empty()) return;
  if (c.seedConfirmation && tops.size() < conf_n_top(conf_range(c, mid.z), mid.r)) return;  // sufficientTopDoublets
  collect(true, bottomBins, bottoms);
  if (bottoms.empty()) return;
  stats[0] += bottoms.size();
  stats[1] += tops.size();

  auto sortList = [&](std::vector<Doublet>& v) {
    if (tieMode == 0) {
      // GPU canonical order: (cotTheta, emission index); the kernel gets there
      // with a bucket sort, any correct sort gives the same unique order
      std::sort(v.begin(), v.end(), [](const Doublet& a, const Doublet& b) {
        if (a.rec.cotTheta != b.rec.cotTheta) return a.rec.cotTheta < b.rec.cotTheta;
        return a.seq < b.seq;
      });
    } else if (tieMode == 1) {
      // exact replay: v is in emission order, sort {index, cotTheta} like
      // DoubletSeedFinder.hpp:94-104 with the libstdc++ algorithm
      std::vector<SortItem> items(v.size());
      for (uint32_t i = 0; i < v.size(); ++i) items[i] = {v[i].rec.cotTheta, i};
      std_sort(items.data(), (int)items.size(), sortItemLess);
      std::vector<Doublet> sorted(v.size());
      for (uint32_t i = 0; i < v.size(); ++i) sorted[i] = v[items[i].val];
      v.swap(sorted);
    } else {
      // what the kernel does: canonical order first, then the pruned replay of
      // std::sort decides the order inside every group of equal cotTheta
      const uint32_t n = (uint32_t)v.size();
      std::vector<uint32_t> canon(n);
      for (uint32_t i = 0; i < n; ++i) canon[i] = i;  // v is in emission order: index == seq rank
      std::sort(canon.begin(), canon.end(), [&](uint32_t a, uint32_t b) {
        if (v[a].rec.cotTheta != v[b].rec.cotTheta) return v[a].rec.cotTheta < v[b].rec.cotTheta;
        return a < b;
      });
      std::vector<uint32_t> group(n, 0xFFFFu);
      bool any = false;
      for (uint32_t i = 0; i < n; ++i) {
        const bool tp = i > 0 && v[canon[i]].rec.cotTheta == v[canon[i - 1]].rec.cotTheta;
        const bool tn = i + 1 < n && v[canon[i]].rec.cotTheta == v[canon[i + 1]].rec.cotTheta;
        if (tp || tn) {
          uint32_t q = i;
          while (q > 0 && v[canon[q - 1]].rec.cotTheta == v[canon[i]].rec.cotTheta) --q;
          group[canon[i]] = q;
          any = true;
        }
      }
      if (any && n > 16) {
        std::vector<SortItem> w(n);
        for (uint32_t i = 0; i < n; ++i) w[i] = {v[i].rec.cotTheta, i | (group[i] << 16)};
        std_sort_replay_ties(w.data(), (int)n, sortItemLess, [](const SortItem& e) { return (e.val >> 16) != 0xFFFFu; });
        for (uint32_t i = 0; i < n; ++i) if (group[canon[i]] != 0xFFFFu) canon[i] = 0xFFFFFFFFu;
        for (uint32_t i = 0; i < n; ++i) {
          const uint32_t g = w[i].val >> 16;
          if (g == 0xFFFFu) continue;
          uint32_t s2 = g;
          while (canon[s2] != 0xFFFFFFFFu) ++s2;
          canon[s2] = w[i].val & 0xFFFFu;
        }
      }
      std::vector<Doublet> sorted(n);
      for (uint32_t i = 0; i < n; ++i) sorted[i] = v[canon[i]];
      v.swap(sorted);
    }
  };
  sortList(bottoms);
  sortList(tops);

  const int nT = (int)tops.size(), nB = (int)bottoms.size();

  // --- per bottom, independent of the others: H_j, brk_j --------------------
  std::vector<int> H(nB), brk(nB), ub(nB);
  std::vector<BottomCtx> ctx(nB);
  for (int j = 0; j < nB; ++j) {
    BottomCtx& b = ctx[j];
    b.cotThetaB = bottoms[j].rec.cotTheta; b.erB = bottoms[j].rec.er;
    b.iDeltaRB = bottoms[j].rec.iDeltaR; b.Ub = bottoms[j].rec.u; b.Vb = bottoms[j].rec.v;
    bottom_ctx(c, b);
    // |P_j| = number of tops with cotT <= cotB  (i.e. !(cotB < cotT))
    int lo = 0, hi = nT;
    while (lo < hi) {
      const int mid2 = (lo + hi) >> 1;
      if (b.cotThetaB < tops[mid2].rec.cotTheta) hi = mid2; else lo = mid2 + 1;
    }
    ub[j] = lo;
    float cu, im;
    int h = 0;
    for (int t = lo - 1; t >= 0; --t) {  // last failing top of the prefix
      const DoubletRec& T = tops[t].rec;
      const int cls = eval_pair(c, mid.r, mid.varZ, mid.varR, b, T.cotTheta, T.er, T.iDeltaR, T.u, T.v, cu, im);
      if (cls == kPairFailA) { h = t + 1; break; }
      if (cls == kPairFailB) { h = t; break; }
    }
    H[j] = h;
    int k = lo;
    for (; k < nT; ++k) {  // first failing top beyond the prefix
      const DoubletRec& T = tops[k].rec;
      const int cls = eval_pair(c, mid.r, mid.varZ, mid.varR, b, T.cotTheta, T.er, T.iDeltaR, T.u, T.v, cu, im);
      if (cls == kPairFailA || cls == kPairFailB) break;
    }
    brk[j] = k;
  }

  // --- window starts: exclusive prefix max of H ------------------------------
  std::vector<int> start(nB);
  {
    int run = 0;
    for (int j = 0; j < nB; ++j) {
      start[j] = run;
      run = std::max(run, H[j]);
    }
  }
  // --- strip path: the window of a bottom is the run of tops inside the cot(theta) pre-filter (device: kStrip scan)
  StripDerived calM{};
  float cosPhiM = 0.f, sinPhiM = 0.f;
  if (gStrip != nullptr) {
    strip_derive(gStrip + 12ull * p.copiedFrom[m], calM);
    cosPhiM = fdiv(mid.x, mid.r);
    sinPhiM = fdiv(mid.y, mid.r);
    for (int j = 0; j < nB; ++j) {
      const float cB = bottoms[j].rec.cotTheta;
      int lo = 0, hi = nT;
      while (lo < hi) {
        const int md = (lo + hi) >> 1;
        if (strip_outside_window(cB, tops[md].rec.cotTheta, gCotThetaDiffMax2) && !(cB < tops[md].rec.cotTheta)) lo = md + 1; else hi = md;
      }
      int k = lo;
      while (k < nT && !(strip_outside_window(cB, tops[k].rec.cotTheta, gCotThetaDiffMax2) && cB < tops[k].rec.cotTheta)) ++k;
      start[j] = lo;
      brk[j] = k;
    }
  }

  // --- candidates, filter, bounded heap -------------------------------------
  if (c.seedConfirmation) conf->push_back({m, {}});
  struct Cand { float curv, impact, topR; uint32_t topPos; };
  const int nLow = (int)c.maxSeedsPerSpMConf;
  std::vector<WeightIndex> heap;  // size <= nLow
  struct Stored { uint32_t b, t; float w, z; };
  std::vector<Stored> storage;
  std::vector<Cand> cands;
  std::vector<SortItem> order;
  for (int j = 0; j < nB; ++j) {
    cands.clear();
    const BottomCtx& b = ctx[j];
    for (int t = start[j]; t < brk[j]; ++t) {
      const DoubletRec& T = tops[t].rec;
      float cu, im;
      if (gStrip != nullptr) {
        if (strip_outside_window(bottoms[j].rec.cotTheta, T.cotTheta, gCotThetaDiffMax2)) continue;  // (cannot happen inside the window)
        stats[2]++;
        StripBottomCtx sb;
        const DoubletRec& Br = bottoms[j].rec;
        sb.cotThetaB0 = Br.cotTheta; sb.iDeltaRB = Br.iDeltaR; sb.erB = Br.er; sb.Ub0 = Br.u; sb.Vb0 = Br.v; sb.xB = Br.xNew; sb.yB = Br.yNew;
        strip_bottom_ctx(c, cosPhiM, sinPhiM, sb);
        StripDerived calB, calT;
        strip_derive(gStrip + 12ull * p.copiedFrom[bottoms[j].pos], calB);
        strip_derive(gStrip + 12ull * p.copiedFrom[tops[t].pos], calT);
        if (!eval_strip_pair(c, gToleranceParam, mid.varZ, mid.varR, sb, calM, calB, calT, T.er, T.iDeltaR, T.u, T.v, T.xNew, T.yNew, cu, im)) continue;
      } else {
      stats[2]++;
      const int cls = eval_pair(c, mid.r, mid.varZ, mid.varR, b, T.cotTheta, T.er, T.iDeltaR, T.u, T.v, cu, im);
      if (cls != kPairEmit) continue;
      }
      const uint32_t tp = tops[t].pos;
      float topR = p.r[tp];
      if (c.useDeltaRinsteadOfTopRadius) {
        const float dr = fsub(p.r[tp], mid.r), dz = fsub(p.z[tp], mid.z);
        topR = fsqrt(fadd(fmul(dr, dr), fmul(dz, dz)));
      }
      cands.push_back({cu, im, topR, tp});
    }
    const int n = (int)cands.size();
    if (n < 1) continue;
    stats[3] += n;
    const float zOrigin = fsub(mid.z, fmul(mid.r, b.cotThetaB));
    order.resize(n);
    for (int i = 0; i < n; ++i) order[i] = {cands[i].curv, (uint32_t)i};
    std_sort(order.data(), n, sortItemLess);  // BroadTripletSeedFilter.cpp:143-148
    if (c.seedConfirmation) {
      // map- and collector-independent part of the filter; the rest is confReplay()
      const uint32_t bp = bottoms[j].pos;
      const float rMaxSeedConfMid = conf_range(c, mid.z).rMaxSeedConf;
      for (int k = 0; k < n; ++k) {
        const Cand& cd = cands[order[k].val];
        uint32_t nCompat;
        float w = filter_weight(
            c, n, k, cd.impact, [&](int i) { return cands[order[i].val].curv; },
            [&](int i) { return cands[order[i].val].topR; }, nCompat);
        int dsc;
        if (!conf_candidate(c, conf_range(c, p.z[bp]), p.r[bp], zOrigin, cd.impact, nCompat, w, dsc)) continue;
        conf->back().rec.push_back({bp, cd.topPos, w, zOrigin, (uint32_t)j, (uint32_t)std::min(n, 3),
                                    !(p.r[bp] > rMaxSeedConfMid), dsc > 0});
      }
      continue;
    }
    for (int k = 0; k < n; ++k) {
      const Cand& cd = cands[order[k].val];
      const float w = filter_weight(
          c, n, k, cd.impact, [&](int i) { return cands[order[i].val].curv; },
          [&](int i) { return cands[order[i].val].topR; });
      // CandidatesForMiddleSp::push, detail/CandidatesForMiddleSp.cpp:44-75
      if (nLow == 0) continue;
      if ((int)heap.size() < nLow) {
        storage.push_back({bottoms[j].pos, cd.topPos, w, zOrigin});
        heap.push_back({w, (uint32_t)storage.size() - 1});
        std_push_heap(heap.data(), (int)heap.size(), heap_comp);
        continue;
      }
      const WeightIndex smallest = heap[0];
      if (w <= smallest.weight) continue;
      storage[smallest.index] = {bottoms[j].pos, cd.topPos, w, zOrigin};
      std_pop_heap(heap.data(), (int)heap.size(), heap_comp);
      heap.back() = {w, smallest.index};
      std_push_heap(heap.data(), (int)heap.size(), heap_comp);
    }
  }
  if (c.seedConfirmation) return;
  // toSortedCandidates + filterTripletsMiddleFixed (BroadTripletSeedFilter.cpp:324-393)
  std_sort_heap(heap.data(), (int)heap.size(), heap_comp);
  size_t maxSeeds = heap.size();
  if (maxSeeds > c.maxSeedsPerSpM) maxSeeds = c.maxSeedsPerSpM + 1;
  for (size_t i = 0; i < heap.size() && i < maxSeeds; ++i) {
    const Stored& s = storage[heap[i].index];
    out.push_back({p.copiedFrom[s.b], p.copiedFrom[m], p.copiedFrom[s.t], s.w, s.z});
  }
}

// ---------------------------------------------------------------------------
// seedConfirmation: scalar twin of k_conf_link / k_conf_replay and of the round
// loop in seeding_plugin.cu.  Seeds of a round: per middle (in processing
// order) a short list; Q(sp, w) = best quality over the previous round's seeds
// of middles before w that contain sp.
// ---------------------------------------------------------------------------
struct ConfSeedM {
  uint32_t b, t;
  float weight, zOrigin;
  bool quality;
  bool operator==(const ConfSeedM& o) const {
    return b == o.b && t == o.t && std::memcmp(&weight, &o.weight, 4) == 0;
  }
};
using ConfRound = std::vector<std::vector<ConfSeedM>>;  // [middle in order] -> seeds

struct ConfLists {  // per space point: (middle order, quality)
  std::vector<std::vector<std::pair<uint32_t, float>>> of;
  float best(uint32_t sp, uint32_t w) const {
    float q = std::numeric_limits<float>::lowest();
    for (const auto& e : of[sp]) if (e.first < w && e.second > q) q = e.second;
    return q;
  }
};

void confPush(std::vector<WeightIndex>& heap, size_t nMax, std::vector<ConfSeedM>& storage, const ConfSeedM& sd) {
  if (nMax == 0) return;
  if (heap.size() < nMax) {
    storage.push_back(sd);
    heap.push_back({sd.weight, (uint32_t)storage.size() - 1});
    std_push_heap(heap.data(), (int)heap.size(), heap_comp);
    return;
  }
  const WeightIndex smallest = heap[0];
  if (sd.weight <= smallest.weight) return;
  storage[smallest.index] = sd;
  std_pop_heap(heap.data(), (int)heap.size(), heap_comp);
  heap.back() = {sd.weight, smallest.index};
  std_push_heap(heap.data(), (int)heap.size(), heap_comp);
}

std::vector<ConfSeedM> confReplay(const DeviceConfig& c, const ConfMiddle& cm, uint32_t w, const ConfLists& q) {
  std::vector<WeightIndex> high, low;
  std::vector<ConfSeedM> storage;
  const float bestM = q.best(cm.m, w);
  size_t i = 0;
  const size_t n = cm.rec.size();
  while (i < n) {
    size_t e = i;
    while (e < n && cm.rec[e].group == cm.rec[i].group) ++e;
    const ConfRecord& g = cm.rec[i];
    const uint32_t minTops = (g.needsTwoTops ? 2u : 1u) + (high.empty() ? 0u : 1u);
    if (g.groupSize >= minTops) {
      const float bestB = q.best(g.b, w);
      bool lowHas = false;
      ConfSeedM lowBest{};
      float lowW = std::numeric_limits<float>::lowest();
      for (size_t k = i; k < e; ++k) {
        const ConfRecord& r = cm.rec[k];
        if (r.weight < bestB && r.weight < bestM && r.weight < q.best(r.t, w)) continue;
        if (r.quality) {
          confPush(high, c.maxQualitySeedsPerSpMConf, storage, {r.b, r.t, r.weight, r.zOrigin, true});
        } else if (r.weight > lowW) {  // evaluated for every group; only used while no quality seed exists
          lowW = r.weight;
          lowBest = {r.b, r.t, r.weight, r.zOrigin, false};
          lowHas = true;
        }
      }
      if (lowHas && high.empty()) confPush(low, c.maxSeedsPerSpMConf, storage, lowBest);
    }
    i = e;
  }
  std_sort_heap(high.data(), (int)high.size(), heap_comp);
  std_sort_heap(low.data(), (int)low.size(), heap_comp);
  std::vector<ConfSeedM> sorted, out;
  for (const WeightIndex& h : high) sorted.push_back(storage[h.index]);
  for (const WeightIndex& l : low) sorted.push_back(storage[l.index]);
  size_t maxSeeds = sorted.size();
  if (maxSeeds > c.maxSeedsPerSpM) maxSeeds = c.maxSeedsPerSpM + 1;
  for (const ConfSeedM& sd : sorted) {
    if (out.size() >= maxSeeds) break;
    if (!high.empty() && !sd.quality) continue;
    float qB = q.best(sd.b, w), qM = bestM, qT = q.best(sd.t, w);
    for (const ConfSeedM& e : out) {  // the middle's own earlier seeds are already in the map
      qM = std::max(qM, e.weight);
      if (e.b == sd.b || e.t == sd.b) qB = std::max(qB, e.weight);
      if (e.b == sd.t || e.t == sd.t) qT = std::max(qT, e.weight);
    }
    if (sd.weight < qB && sd.weight < qM && sd.weight < qT) continue;
    out.push_back(sd);
  }
  return out;
}

// returns the number of rounds (>= 2), seeds appended to `out`
int confFixedPoint(const DeviceConfig& c, const Packed& p, size_t nSp, const std::vector<ConfMiddle>& middles,
                   std::vector<SeedOut>& out) {
  ConfRound prev(middles.size()), cur(middles.size());
  int rounds = 0;
  for (;;) {
    ConfLists lists;
    lists.of.resize(nSp);
    if (rounds > 0) {
      for (uint32_t w = 0; w < middles.size(); ++w) {
        for (const ConfSeedM& sd : prev[w]) {
          for (uint32_t sp : {sd.b, middles[w].m, sd.t}) lists.of[sp].push_back({w, sd.weight});
        }
      }
    }
    bool changed = rounds == 0;
    for (uint32_t w = 0; w < middles.size(); ++w) {
      cur[w] = confReplay(c, middles[w], w, lists);
      if (rounds > 0 && !(cur[w] == prev[w])) changed = true;
    }
    ++rounds;
    prev.swap(cur);
    if (!changed) break;
  }
  for (uint32_t w = 0; w < middles.size(); ++w) {
    for (const ConfSeedM& sd : prev[w]) {
      out.push_back({p.copiedFrom[sd.b], p.copiedFrom[middles[w].m], p.copiedFrom[sd.t], sd.weight, sd.zOrigin});
    }
  }
  return rounds;
}

}  // namespace

extern "C" {

// Runs the restructured algorithm on an already built grid (packed arrays +
// per-bin ranges).  Neighbour bins are passed per middle bin as index lists
// into binBegin/binEnd: nbrOffsets has nMiddleBins+1 entries for each side.
// middleBins lists the middle bins in navigation order with their r range.
int64_t model_run(const DeviceConfig* cfg, const uint32_t* copiedFrom, const float* x,
                  const float* y, const float* z, const float* r, const float* varZ,
                  const float* varR, const uint32_t* binBegin, const uint32_t* binEnd,
                  uint32_t nMiddleBins, const uint32_t* middleBins, const float* rangeMin,
                  const float* rangeMax, const uint32_t* botOffsets, const uint32_t* botBins,
                  const uint32_t* topOffsets, const uint32_t* topBins, uint32_t nZWin,
                  const float* zLo, const float* zHi, int tieMode, uint64_t capacity,
                  uint32_t* ob, uint32_t* om, uint32_t* ot, float* oq, float* oz,
                  uint64_t* stats) {
  Packed p{copiedFrom, x, y, z, r, varZ, varR};
  std::vector<SeedOut> out;
  std::vector<ConfMiddle> confMiddles;
  for (int i = 0; i < 8; ++i) stats[i] = 0;
  for (uint32_t g = 0; g < nMiddleBins; ++g) {
    const uint32_t mb = middleBins[g];
    if (binBegin[mb] == binEnd[mb]) continue;
    std::vector<Window> bot, top;
    for (uint32_t k = botOffsets[g]; k < botOffsets[g + 1]; ++k) bot.push_back({binBegin[botBins[k]], binEnd[botBins[k]]});
    for (uint32_t k = topOffsets[g]; k < topOffsets[g + 1]; ++k) top.push_back({binBegin[topBins[k]], binEnd[topBins[k]]});
    // TripletSeeder.cpp:183-195 as a predicate-defined contiguous range
    const uint32_t mlo = firstTrue(binBegin[mb], binEnd[mb], [&](uint32_t i) { return !(r[i] < rangeMin[g]); });
    const uint32_t mhi = firstTrue(mlo, binEnd[mb], [&](uint32_t i) { return r[i] > rangeMax[g]; });
    for (uint32_t m = mlo; m < mhi; ++m) {
      processMiddle(*cfg, p, m, binBegin[mb], bot, top, zLo, zHi, (int)nZWin, tieMode, out, stats, &confMiddles);
    }
  }
  if (cfg->seedConfirmation) {
    size_t nSp = 0;
    for (const ConfMiddle& cm : confMiddles) {
      nSp = std::max<size_t>(nSp, cm.m + 1);
      for (const ConfRecord& rc : cm.rec) nSp = std::max<size_t>(nSp, std::max(rc.b, rc.t) + 1);
    }
    stats[4] = (uint64_t)confFixedPoint(*cfg, p, nSp, confMiddles, out);
  }
  if (out.size() <= capacity) {
    for (size_t i = 0; i < out.size(); ++i) {
      ob[i] = out[i].b; om[i] = out[i].m; ot[i] = out[i].t; oq[i] = out[i].q; oz[i] = out[i].z;
    }
  }
  return (int64_t)out.size();
}

// strip triplet path for the next model_run calls (strip = NULL: back to the pixel path)
void model_set_strips(const float* strip, float cotThetaDiffMax, float toleranceParam) {
  gStrip = strip;
  gCotThetaDiffMax2 = cotThetaDiffMax * cotThetaDiffMax;
  gToleranceParam = toleranceParam;
}

// grid stage of the model: bin index with the replayed atan2f, per-SP
int32_t model_bin_index(const DeviceConfig* cfg, float x, float y, float z, float r) {
  if (cfg->useExtraCuts && !itk_sp_select(r, z)) return -1;
  return grid_bin_index(*cfg, glibc_atan2f(y, x), z, r);
}
float model_atan2f(float y, float x) { return glibc_atan2f(y, x); }

// ---- self checks of the libstdc++ replays against the real libstdc++ -------
// returns the number of trials whose result differs from std::sort
int64_t model_check_std_sort(uint64_t seed, int maxN, int trials, int keyRange) {
  std::mt19937_64 rng(seed);
  int64_t bad = 0;
  for (int t = 0; t < trials; ++t) {
    const int n = (int)(rng() % (uint64_t)(maxN + 1));
    std::vector<SortItem> a(n), b;
    const int mode = (int)(rng() % 5);
    for (int i = 0; i < n; ++i) {
      float key;
      switch (mode) {
        case 0: key = (float)(rng() % (uint64_t)keyRange); break;              // heavy ties
        case 1: key = (float)i; break;                                          // sorted
        case 2: key = (float)(n - i); break;                                    // reversed
        case 3: key = (float)((i * 7919) % std::max(1, n / 3)); break;          // patterned ties
        default: key = (float)(rng() % 1000003) * 1e-3f; break;                 // mostly distinct
      }
      a[i] = {key, (uint32_t)i};
    }
    b = a;
    std_sort(a.data(), n, sortItemLess);
    std::sort(b.begin(), b.end(), sortItemLess);
    for (int i = 0; i < n; ++i) {
      if (a[i].val != b[i].val) { ++bad; break; }
    }
  }
  return bad;
}

// pruned tie replay: canonical order + replay must reproduce std::sort exactly
int64_t model_check_tie_replay(uint64_t seed, int maxN, int trials, int nTieGroups) {
  std::mt19937_64 rng(seed);
  int64_t bad = 0;
  for (int t = 0; t < trials; ++t) {
    const int n = 2 + (int)(rng() % (uint64_t)(maxN - 1));
    std::vector<SortItem> in(n);
    for (int i = 0; i < n; ++i) in[i] = {(float)(rng() % 100000007ull), (uint32_t)i};
    const int groups = (int)(rng() % (uint64_t)(nTieGroups + 1));
    for (int gI = 0; gI < groups; ++gI) {  // plant tie groups of size 2..4
      const int k = 2 + (int)(rng() % 3);
      const float key = in[rng() % n].key;
      for (int j = 0; j < k; ++j) in[rng() % n].key = key;
    }
    std::vector<SortItem> ref = in;
    std::sort(ref.begin(), ref.end(), sortItemLess);
    std::vector<uint32_t> canon(n);
    for (int i = 0; i < n; ++i) canon[i] = i;
    std::sort(canon.begin(), canon.end(), [&](uint32_t a, uint32_t b) {
      if (in[a].key != in[b].key) return in[a].key < in[b].key;
      return a < b;
    });
    std::vector<uint32_t> group(n, 0xFFFFu);
    for (int i = 0; i < n; ++i) {
      const bool tp = i > 0 && in[canon[i]].key == in[canon[i - 1]].key;
      const bool tn = i + 1 < n && in[canon[i]].key == in[canon[i + 1]].key;
      if (tp || tn) {
        int q = i;
        while (q > 0 && in[canon[q - 1]].key == in[canon[i]].key) --q;
        group[canon[i]] = (uint32_t)q;
      }
    }
    if (n > 16) {
      std::vector<SortItem> w(n);
      for (int i = 0; i < n; ++i) w[i] = {in[i].key, (uint32_t)i | (group[i] << 16)};
      std_sort_replay_ties(w.data(), n, sortItemLess, [](const SortItem& e) { return (e.val >> 16) != 0xFFFFu; });
      for (int i = 0; i < n; ++i) if (group[canon[i]] != 0xFFFFu) canon[i] = 0xFFFFFFFFu;
      for (int i = 0; i < n; ++i) {
        const uint32_t g = w[i].val >> 16;
        if (g == 0xFFFFu) continue;
        uint32_t s2 = g;
        while (canon[s2] != 0xFFFFFFFFu) ++s2;
        canon[s2] = w[i].val & 0xFFFFu;
      }
    }
    for (int i = 0; i < n; ++i) {
      if (canon[i] != ref[i].val) { ++bad; break; }
    }
  }
  return bad;
}


}  // extern "C" (templates below need C++ linkage)

// ---------------------------------------------------------------------------
// Scalar emulation of the WARP-parallel pruned replay used on the device
// (seeding_kernels.cuh: warp_sort_replay_ties).  Lanes are emulated by loops;
// the batch partition step must be equivalent to libstdc++'s sequential
// __unguarded_partition.
// ---------------------------------------------------------------------------
namespace {
inline int fnsEmu(uint32_t mask, int k) {  // lane of the k-th (1-based) set bit, -1 if none
  for (int l = 0; l < 32; ++l) {
    if (mask & (1u << l)) { if (--k == 0) return l; }
  }
  return -1;
}
template <typename T, typename Less, typename Flagged>
void warpReplayEmu(T* a, int n, Less less, Flagged flagged) {
  if (n <= 16) return;
  int lg = 0;
  for (unsigned v = (unsigned)n; v > 1; v >>= 1) ++lg;
  int sf[64], sl[64], sd[64], sp = 0;
  sf[0] = 0; sl[0] = n; sd[0] = 2 * lg; sp = 1;
  while (sp > 0) {
    --sp;
    int first = sf[sp], last = sl[sp], depth = sd[sp];
    while (last - first > 16) {
      int c = 0;
      for (int base = first; base < last && c < 2; base += 32) {
        uint32_t m = 0;
        for (int l = 0; l < 32; ++l) { const int i = base + l; if (i < last && flagged(a[i])) m |= 1u << l; }
        c += __builtin_popcount(m);
      }
      if (c < 2) break;
      if (depth == 0) { std_make_heap(a + first, last - first, less); std_sort_heap(a + first, last - first, less); break; }
      --depth;
      const int mid = first + (last - first) / 2;
      {
        const int ia = first + 1, ib = mid, ic = last - 1;
        int pick;
        if (less(a[ia], a[ib])) { if (less(a[ib], a[ic])) pick = ib; else if (less(a[ia], a[ic])) pick = ic; else pick = ia; }
        else if (less(a[ia], a[ic])) pick = ia; else if (less(a[ib], a[ic])) pick = ic; else pick = ib;
        std::swap(a[first], a[pick]);
      }
      const T pivot = a[first];
      int lo = first + 1, hi = last;
      while (hi - lo >= 64) {  // batch step: 32 from the left, 32 from the right
        T el[32], er[32];
        uint32_t mL = 0, mR = 0;
        for (int l = 0; l < 32; ++l) {
          el[l] = a[lo + l]; er[l] = a[hi - 1 - l];
          if (!less(el[l], pivot)) mL |= 1u << l;
          if (!less(pivot, er[l])) mR |= 1u << l;
        }
        const int nL = __builtin_popcount(mL), nR = __builtin_popcount(mR), s2 = std::min(nL, nR);
        for (int l = 0; l < 32; ++l) {
          if (mL & (1u << l)) { const int rk = __builtin_popcount(mL & ((1u << l) - 1u)); if (rk < s2) a[lo + l] = er[fnsEmu(mR, rk + 1)]; }
          if (mR & (1u << l)) { const int rk = __builtin_popcount(mR & ((1u << l) - 1u)); if (rk < s2) a[hi - 1 - l] = el[fnsEmu(mL, rk + 1)]; }
        }
        const int newLo = (nL == s2) ? lo + 32 : lo + fnsEmu(mL, s2 + 1);
        const int newHi = (nR == s2) ? hi - 32 : hi - fnsEmu(mR, s2 + 1);
        lo = newLo; hi = newHi;
      }
      while (true) {  // serial remainder, libstdc++ __unguarded_partition from the same state
        while (less(a[lo], pivot)) ++lo;
        --hi;
        while (less(pivot, a[hi])) --hi;
        if (!(lo < hi)) break;
        std::swap(a[lo], a[hi]);
        ++lo;
      }
      const int cut = lo;
      sf[sp] = cut; sl[sp] = last; sd[sp] = depth; ++sp;
      last = cut;
    }
  }
}
}  // namespace

extern "C" {

int64_t model_check_warp_replay(uint64_t seed, int maxN, int trials, int nTieGroups, int fullSort) {
  std::mt19937_64 rng(seed);
  int64_t bad = 0;
  for (int t = 0; t < trials; ++t) {
    const int n = 2 + (int)(rng() % (uint64_t)(maxN - 1));
    std::vector<SortItem> in(n);
    const int mode = (int)(rng() % 4);
    for (int i = 0; i < n; ++i) {
      float key = (float)(rng() % 100000007ull);
      if (mode == 1) key = (float)i; else if (mode == 2) key = (float)(n - i);
      in[i] = {key, (uint32_t)i};
    }
    const int groups = (int)(rng() % (uint64_t)(nTieGroups + 1));
    for (int gI = 0; gI < groups; ++gI) {
      const int k = 2 + (int)(rng() % 3);
      const float key = in[rng() % n].key;
      for (int j = 0; j < k; ++j) in[rng() % n].key = key;
    }
    std::vector<SortItem> ref = in;
    std::sort(ref.begin(), ref.end(), sortItemLess);
    std::vector<uint32_t> canon(n);
    for (int i = 0; i < n; ++i) canon[i] = i;
    std::sort(canon.begin(), canon.end(), [&](uint32_t x, uint32_t y) { if (in[x].key != in[y].key) return in[x].key < in[y].key; return x < y; });
    std::vector<uint32_t> group(n, 0xFFFFFFFFu);
    for (int i = 0; i < n; ++i) {
      const bool tp = i > 0 && in[canon[i]].key == in[canon[i - 1]].key;
      const bool tn = i + 1 < n && in[canon[i]].key == in[canon[i + 1]].key;
      if (tp || tn || fullSort) {
        int q = i;
        while (q > 0 && in[canon[q - 1]].key == in[canon[i]].key) --q;
        group[canon[i]] = (uint32_t)q;
      }
    }
    if (n > 16) {
      struct W { float key; uint32_t val; uint32_t grp; };
      std::vector<W> w(n);
      for (int i = 0; i < n; ++i) w[i] = {in[i].key, (uint32_t)i, group[i]};
      warpReplayEmu(w.data(), n, [](const W& x, const W& y) { return x.key < y.key; }, [](const W& e) { return e.grp != 0xFFFFFFFFu; });
      for (int i = 0; i < n; ++i) if (group[canon[i]] != 0xFFFFFFFFu) canon[i] = 0xFFFFFFFFu;
      for (int i = 0; i < n; ++i) {
        if (w[i].grp == 0xFFFFFFFFu) continue;
        uint32_t s2 = w[i].grp;
        while (canon[s2] != 0xFFFFFFFFu) ++s2;
        canon[s2] = w[i].val;
      }
    }
    for (int i = 0; i < n; ++i) {
      if (canon[i] != ref[i].val) { ++bad; break; }
    }
  }
  return bad;
}

// adversarial input for median-of-3 quicksort (forces the heap-sort fallback)
int64_t model_check_std_sort_killer(int n) {
  // Musser's median-of-3 killer sequence
  std::vector<SortItem> a(n);
  const int k = n / 2;
  for (int i = 0; i < n; ++i) a[i] = {0.f, (uint32_t)i};
  for (int i = 1; i <= k; ++i) {
    if (i % 2 == 1) { a[i - 1].key = (float)i; a[i].key = (float)(k + i); }
    a[k + i - 1].key = (float)(2 * i);
  }
  std::vector<SortItem> b = a;
  std_sort(a.data(), n, sortItemLess);
  std::sort(b.begin(), b.end(), sortItemLess);
  int64_t bad = 0;
  for (int i = 0; i < n; ++i) bad += a[i].val != b[i].val;
  return bad;
}

int64_t model_check_heap(uint64_t seed, int trials) {
  std::mt19937_64 rng(seed);
  int64_t bad = 0;
  auto comp = [](const std::pair<float, uint32_t>& a, const std::pair<float, uint32_t>& b) { return a.first > b.first; };
  for (int t = 0; t < trials; ++t) {
    const int cap = 1 + (int)(rng() % 8);
    const int nPush = (int)(rng() % 64);
    std::vector<std::pair<float, uint32_t>> ref;
    std::vector<WeightIndex> mine;
    for (int i = 0; i < nPush; ++i) {
      const float w = (float)(rng() % 6);
      if ((int)ref.size() < cap) {
        ref.emplace_back(w, (uint32_t)i); std::push_heap(ref.begin(), ref.end(), comp);
        mine.push_back({w, (uint32_t)i}); std_push_heap(mine.data(), (int)mine.size(), heap_comp);
      } else if (!(w <= ref.front().first)) {
        std::pop_heap(ref.begin(), ref.end(), comp); ref.back() = {w, (uint32_t)i}; std::push_heap(ref.begin(), ref.end(), comp);
        std_pop_heap(mine.data(), (int)mine.size(), heap_comp); mine.back() = {w, (uint32_t)i}; std_push_heap(mine.data(), (int)mine.size(), heap_comp);
      }
    }
    std::sort_heap(ref.begin(), ref.end(), comp);
    std_sort_heap(mine.data(), (int)mine.size(), heap_comp);
    for (size_t i = 0; i < ref.size(); ++i) {
      if (ref[i].second != mine[i].index) { ++bad; break; }
    }
  }
  return bad;
}

int64_t model_check_atan2f(uint64_t seed, int64_t n) {
  uint64_t s = seed | 1;
  int64_t bad = 0;
  auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (int64_t i = 0; i < n; ++i) {
    float x = ((int64_t)(next() & 0xffffff) - 0x800000) * (200.f / 0x800000);
    float y = ((int64_t)(next() & 0xffffff) - 0x800000) * (200.f / 0x800000);
    if (i % 7 == 0) x = u2f((uint32_t)(next() >> 32));
    if (i % 11 == 0) y = u2f((uint32_t)(next() >> 20));
    const float a = std::atan2(y, x), b = glibc_atan2f(y, x);
    if (f2u(a) != f2u(b) && !(a != a && b != b)) ++bad;
  }
  return bad;
}

uint64_t model_sizeof_device_config() { return sizeof(DeviceConfig); }

// classify_pair_flat (branch-free, used by the device scans) against classify_pair (the reference's early exits)
// on random pairs around realistic magnitudes, incl. dU == 0, equal slopes and values at the cut edges.
int64_t model_check_flat_classify(uint64_t seed, int64_t n) {
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  DeviceConfig c{};
  c.minHelixDiameter2 = 2781625.5f;
  c.impactMax = 3.0f;
  int64_t bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    BottomCtx b{};
    const float scale = std::pow(10.f, 4.f * U(rng));  // wide dynamic range of the slope windows
    b.cotThetaB = 3.f * U(rng);
    b.erB = 1e-6f * std::fabs(U(rng)) * scale;
    b.iDeltaRB = 0.02f * std::fabs(U(rng)) + 0.003f;
    b.Ub = 0.03f * U(rng);
    b.Vb = 0.01f * U(rng);
    b.sigmaSquaredPtDependent = (1.f + b.cotThetaB * b.cotThetaB) * 4283.94f;
    b.scatteringInRegion2 = 1.54008687e-3f * (1.f + b.cotThetaB * b.cotThetaB);
    const float rM = 60.f + 60.f * std::fabs(U(rng));
    const float cotT = b.cotThetaB + 0.08f * U(rng) * U(rng);
    const float erT = 1e-6f * std::fabs(U(rng)) * scale;
    const float iDRT = 0.02f * std::fabs(U(rng)) + 0.003f;
    float uT = b.Ub + 0.02f * U(rng) * U(rng) * U(rng);
    float vT = b.Vb + 0.002f * U(rng) * U(rng);
    if (i % 97 == 0) uT = b.Ub;        // dU == 0
    if (i % 101 == 0) vT = b.Vb;       // A == 0
    const int want = classify_pair(c, rM, 2e-4f, 4e-6f, b, cotT, erT, iDRT, uT, vT);
    const int got = classify_pair_flat(c, rM, 2e-4f, 4e-6f, b, cotT, erT, iDRT, uT, vT);
    bad += want != got;
  }
  return bad;
}

}  // extern "C"


// ---------------------------------------------------------------------------
// Orthogonal seeder, k-d tree construction on the device (acts_b200/csrc/orthogonal_kernels.cuh):
//  * k_kd_split replays libstdc++'s bidirectional std::partition by RANKS -- the k-th element of [0, P) that fails the
//    predicate is swapped with the k-th element from the right of [P, n) that satisfies it (P = number of elements
//    that satisfy it) -- instead of walking two cursors towards each other;
//  * the whole construction (boxes, mid-point split above 128 elements, std::sort + median below, the pivot repair)
//    must leave the elements in the order Acts::KDTree leaves them in.
// Both are checked here against the real library calls, without a GPU.
// ---------------------------------------------------------------------------
extern "C" {

// std::partition vs the rank pairing on random keys (duplicates, all-true / all-false included); returns mismatches
int64_t model_check_partition_pairing(uint64_t seed, int maxN, int trials) {
  std::mt19937_64 rng(seed);
  int64_t bad = 0;
  for (int t = 0; t < trials; ++t) {
    const int n = 1 + (int)(rng() % (uint64_t)maxN);
    const int mode = (int)(rng() % 5);
    std::vector<std::pair<float, uint32_t>> a(n);
    for (int i = 0; i < n; ++i) a[i] = {(float)(rng() % (mode == 3 ? 3u : 1000u)), (uint32_t)i};
    const float mid = mode == 0 ? -1.f : (mode == 1 ? 2000.f : (float)(rng() % 1000u));
    auto pred = [mid](const std::pair<float, uint32_t>& v) { return v.first < mid; };
    std::vector<std::pair<float, uint32_t>> ref = a;
    const auto it = std::partition(ref.begin(), ref.end(), pred);
    const int P = (int)(it - ref.begin());
    // the device formulation
    int nTrue = 0;
    for (int i = 0; i < n; ++i) nTrue += pred(a[i]) ? 1 : 0;
    std::vector<int> listL, listR;
    for (int i = 0; i < n; ++i) {
      if (i < nTrue && !pred(a[i])) listL.push_back(i);
      if (i >= nTrue && pred(a[i])) listR.push_back(i);
    }
    if (nTrue != P || listL.size() != listR.size()) { ++bad; continue; }
    const int nMis = (int)listL.size();
    for (int k = 0; k < nMis; ++k) std::swap(a[listL[k]], a[listR[nMis - 1 - k]]);
    for (int i = 0; i < n; ++i) bad += (a[i] != ref[i]) ? 1 : 0;
  }
  return bad;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Strip triplet path (k_seed_middles<..., kStrip>): the device evaluates, for every bottom on its own, the run of
// sorted tops that starts at the first top which is not "outside the cot(theta) window and at or below the bottom"
// (binary search) and ends before the first top that is outside and above.  The reference
// (TripletSeedFinder.cpp:214-238,397-403) walks the tops sequentially, drops the tops below the window for all later
// bottoms (subrange) and stops the loop of a bottom at the first top above it.  Both must enumerate the same
// (bottom, top) pairs in the same order -- checked here on random sorted lists (ties, +-inf windows, empty lists).
// ---------------------------------------------------------------------------
extern "C" {

int64_t model_check_strip_window(uint64_t seed, int maxN, int trials) {
  std::mt19937_64 rng(seed);
  int64_t bad = 0;
  for (int t = 0; t < trials; ++t) {
    const int nB = (int)(rng() % (uint64_t)(maxN + 1)), nT = (int)(rng() % (uint64_t)(maxN + 1));
    const int mode = (int)(rng() % 4);
    auto draw = [&]() { return mode == 2 ? (float)(rng() % 7u) * 0.25f : (float)((double)(rng() % 2000001u) * 1e-5 - 10.0); };
    std::vector<float> cotB(nB), cotT(nT);
    for (float& v : cotB) v = draw();
    for (float& v : cotT) v = draw();
    std::sort(cotB.begin(), cotB.end());
    std::sort(cotT.begin(), cotT.end());
    const float diffMax = mode == 0 ? std::numeric_limits<float>::infinity() : (mode == 1 ? 0.0f : (float)(rng() % 400u) * 0.01f);
    const float diffMax2 = diffMax * diffMax;
    // the reference's sequential form
    std::vector<std::pair<int, int>> ref, dev;
    std::size_t begin = 0;
    for (int j = 0; j < nB; ++j) {
      if (begin == (std::size_t)nT) break;  // TripletSeeder.cpp:29-31
      std::size_t offset = 0;
      for (std::size_t k = begin; k < (std::size_t)nT; ++k) {
        const float d = cotB[j] - cotT[k];
        if (d * d > diffMax2) {
          if (cotB[j] < cotT[k]) break;
          offset = k - begin + 1;
          continue;
        }
        ref.emplace_back(j, (int)k);
      }
      begin += offset;
    }
    // the device form (seeding_kernels.cuh, kStrip scan; strip_outside_window of seed_math.h)
    for (int j = 0; j < nB; ++j) {
      uint32_t lo = 0, hi = (uint32_t)nT;
      while (lo < hi) {
        const uint32_t md = (lo + hi) >> 1;
        if (strip_outside_window(cotB[j], cotT[md], diffMax2) && !(cotB[j] < cotT[md])) lo = md + 1; else hi = md;
      }
      for (uint32_t k = lo; k < (uint32_t)nT; ++k) {
        if (strip_outside_window(cotB[j], cotT[k], diffMax2)) {
          if (cotB[j] < cotT[k]) break;
          continue;
        }
        dev.emplace_back(j, (int)k);
      }
    }
    bad += ref == dev ? 0 : 1;
  }
  return bad;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Phase 3f of k_seed_middles keeps only the nUse + 1 = min(nLow, maxSeedsPerSpM + 1) + 1 largest weights in registers and
// replays the literal bounded heap (capacity nLow, CandidatesForMiddleSp.cpp:44-93) only when two of them are equal.
// Claim checked here: whenever the nUse + 1 largest weights differ, the first nUse entries of the sort_heap'ed
// collector are the nUse largest weights in descending order -- for any nLow >= nUse, any push order, any ties below.
// ---------------------------------------------------------------------------
extern "C" {

int64_t model_check_topk_claim(uint64_t seed, int trials) {
  std::mt19937_64 rng(seed);
  int64_t bad = 0, exercised = 0;
  for (int t = 0; t < trials; ++t) {
    const int n = (int)(rng() % 400u);
    const int nLow = 1 + (int)(rng() % 128u);
    const int maxSeeds = (int)(rng() % 8u);
    const int nUse = std::min(nLow, maxSeeds + 1);
    const uint32_t levels = 1u + (uint32_t)(rng() % (t % 3 == 0 ? 12u : 100000u));  // few levels: many equal weights
    std::vector<float> w(n);
    for (float& v : w) v = (float)(rng() % levels) * 0.5f;
    // literal collector
    std::vector<WeightIndex> heap;
    std::vector<float> storage;
    for (int i = 0; i < n; ++i) {
      if ((int)heap.size() < nLow) {
        storage.push_back(w[i]);
        heap.push_back({w[i], (uint32_t)storage.size() - 1});
        std_push_heap(heap.data(), (int)heap.size(), heap_comp);
        continue;
      }
      const WeightIndex smallest = heap[0];
      if (w[i] <= smallest.weight) continue;
      storage[smallest.index] = w[i];
      std_pop_heap(heap.data(), (int)heap.size(), heap_comp);
      heap.back() = {w[i], smallest.index};
      std_push_heap(heap.data(), (int)heap.size(), heap_comp);
    }
    std_sort_heap(heap.data(), (int)heap.size(), heap_comp);
    // the register selection: nUse + 1 largest, stable in arrival order
    std::vector<float> top(w);
    std::stable_sort(top.begin(), top.end(), [](float a, float b) { return a > b; });
    const int keep = std::min<int>(nUse + 1, n);
    bool distinct = true;
    for (int i = 0; i + 1 < keep; ++i) distinct &= top[i] != top[i + 1];
    if (!distinct) continue;  // the device runs the literal replay here
    ++exercised;
    const int nOut = std::min<int>((int)heap.size(), maxSeeds + 1);
    if (nOut != std::min(n, nUse)) { ++bad; continue; }
    for (int i = 0; i < nOut; ++i) bad += heap[i].weight != top[i] ? 1 : 0;
  }
  return exercised > trials / 10 ? bad : -1;
}

}  // extern "C"
