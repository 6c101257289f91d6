// seeding_oracle.cpp -- CPU ORACLE for the B200 seeding plugin.
//
// TEST INFRASTRUCTURE ONLY.  This file is a dependency-free C++20 restatement
// of the reference algorithm (ACTS GridTripletSeedingAlgorithm and the Core
// seeding classes it drives).  It defines truth for the parity tests and is the
// timed CPU baseline of bench.py; the product (acts_b200/csrc) never links,
// imports or calls it.
//
// PARITY STATUS: *pinned to the reference itself*.  The reference holds no golden
// vector / KAT for this path, but its unmodified sources compile in this container
// against stand-in third-party headers (oracle/Makefile target `ref`, oracle/_ref);
// tests/test_reference_pin.py runs the same inputs through the reference's own
// execute() and through this restatement and demands bit-identical seeds in the same
// order (canonical configurations, the verbatim ITk configuration, fuzzed
// configurations, vertex windows, tie storms, full-size events, the orthogonal seeder).
//
// Arithmetic contract: every cut is IEEE binary32, evaluated in the reference's
// operation order, compiled WITHOUT fma contraction and without -march (the
// reference's default build: cmake/ActsCompilerOptions.cmake:2-14).  Sorting and
// heap operations use the same libstdc++ algorithms the reference calls
// (std::ranges::sort, push_heap, pop_heap, sort_heap), phi uses the host libm
// atan2f (GridTripletSeedingAlgorithm.cpp:219).

#include "../include/acts_b200_seeding.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numbers>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

using Index = std::uint32_t;

// ---------------------------------------------------------------------------
// small helpers restating Acts/Utilities/MathHelpers.hpp:74-80,177-180
// ---------------------------------------------------------------------------
inline float squaref(float v) { return v * v; }
// fastCathetus(h, 1) with an int second argument: fastHypot(int) = std::abs(int)
inline float fastCathetusOne(float h) { return std::sqrt((h - 1) * (h + 1)); }

struct ZWindows {
  const float* lo = nullptr;
  const float* hi = nullptr;
  std::uint32_t n = 0;
};

// Everything the constructor chain derives from the user configuration.
struct Setup {
  b200seed_config cfg{};
  std::vector<std::pair<int, int>> zNeighborsTop, zNeighborsBottom;
  std::vector<float> zBinEdgesF;
  std::vector<std::size_t> zCustomLooping;
  std::vector<std::pair<float, float>> rRangeMiddle;

  // grid (CylindricalSpacePointGrid.cpp:15-99)
  int phiBins = 0;
  double phiMin = 0, phiMax = 0, phiWidth = 0;
  std::vector<double> zEdges, rEdges;
  std::size_t nZ = 0, nR = 0, nGlobal = 0;
  std::vector<std::size_t> navPhi, navZ, navR;

  // doublet finders (GridTripletSeedingAlgorithm.cpp:272-309)
  float dRMinB = 0, dRMaxB = 0, dRMinT = 0, dRMaxT = 0;
  float minHelixDiameter2Doublet = 0;
  // triplet finder (TripletSeedFinder.cpp:456-481)
  float highland = 0, minHelixDiameter2 = 0, sigmapT2perRadius = 0,
        multipleScattering2 = 0;
};

// SpacePointGridPhiBinning.cpp:20-95
int computePhiBins(float minPt, float bFieldInZ, float rMax, float deltaRMax,
                   float impactMax, int phiBinDeflectionCoverage,
                   int maxPhiBins) {
  if (bFieldInZ == 0) {
    return maxPhiBins;
  }
  const float minHelixRadius = minPt / bFieldInZ;
  if (minHelixRadius < rMax * 0.5) {
    throw std::domain_error(
        "phi binning: minimum-pT helix radius below rMax / 2");
  }
  const float outerAngle =
      std::atan(1.f / fastCathetusOne(2 * minHelixRadius / rMax));
  float innerAngle = 0;
  float rMin = rMax;
  if (rMax > deltaRMax) {
    const float innerCircleR = rMax - deltaRMax;
    rMin = innerCircleR;
    innerAngle =
        std::atan(1.f / fastCathetusOne(2 * minHelixRadius / innerCircleR));
  }
  const float sinInner = std::min(1.f, impactMax / rMin);
  const float sinOuter = std::min(1.f, impactMax / rMax);
  const float deltaAngleWithMaxD0 =
      std::abs(std::asin(sinInner) - std::asin(sinOuter));
  const float deltaPhi = (outerAngle - innerAngle + deltaAngleWithMaxD0) /
                         phiBinDeflectionCoverage;
  if (deltaPhi <= 0.f) {
    throw std::domain_error(
        "phi binning: bin width <= 0");
  }
  const int phiBins =
      static_cast<int>(std::ceil(2 * std::numbers::pi / deltaPhi));
  return std::min(phiBins, maxPhiBins);
}

// BinnedGroup.ipp:32-62
void completeNavigation(std::vector<std::size_t>& bins, std::size_t nBins,
                        int axis) {
  if (bins.empty()) {
    bins.resize(nBins);
    std::iota(bins.begin(), bins.end(), std::size_t{1});
    return;
  }
  std::vector<bool> visited(nBins + 1, false);
  for (std::size_t bin : bins) {
    if (bin == 0 || bin > nBins) {
      throw std::invalid_argument("Invalid navigation for axis " +
                                  std::to_string(axis) + ": bin " +
                                  std::to_string(bin) + " is out of range.");
    }
    if (visited[bin]) {
      throw std::invalid_argument("Invalid navigation for axis " +
                                  std::to_string(axis) + ": bin " +
                                  std::to_string(bin) +
                                  " is listed more than once.");
    }
    visited[bin] = true;
  }
}

// derived constants of the doublet / triplet finders
void deriveFinderConstants(Setup& s) {
  const b200seed_config& c = s.cfg;
  // DoubletSeedFinder.cpp:351-357
  {
    const float pTPerHelixRadius = c.bFieldInZ;
    s.minHelixDiameter2Doublet =
        squaref(c.minPt * 2 / pTPerHelixRadius) * c.helixCutTolerance;
  }
  // TripletSeedFinder.cpp:456-481
  {
    const double xOverX0 = c.radLengthPerSeed;
    const double q2OverBeta2 = 1;
    const double t = std::sqrt(xOverX0 * q2OverBeta2);
    // 13.6_MeV: UnitConstants::MeV (double 1e-3) * 13.6L (Units.hpp:149,181-184)
    const double e136 = static_cast<double>(1e-3 * 13.6L);
    s.highland =
        static_cast<float>(e136 * t * (1.0 + 0.038 * 2 * std::log(t)));
    const float maxScatteringAngle = s.highland / c.minPt;
    const float maxScatteringAngle2 = maxScatteringAngle * maxScatteringAngle;
    const float pTPerHelixRadius = c.bFieldInZ;
    s.minHelixDiameter2 =
        squaref(c.minPt * 2 / pTPerHelixRadius) * c.helixCutTolerance;
    const float pT2perRadius = squaref(s.highland / pTPerHelixRadius);
    s.sigmapT2perRadius = pT2perRadius * squaref(2 * c.sigmaScattering);
    s.multipleScattering2 = maxScatteringAngle2 * squaref(c.sigmaScattering);
  }
}

Setup makeSetup(const b200seed_config& c) {
  Setup s;
  s.cfg = c;
  for (std::uint32_t i = 0; i < c.nZBinNeighborsTop; ++i) {
    s.zNeighborsTop.emplace_back(c.zBinNeighborsTop[2 * i],
                                 c.zBinNeighborsTop[2 * i + 1]);
  }
  for (std::uint32_t i = 0; i < c.nZBinNeighborsBottom; ++i) {
    s.zNeighborsBottom.emplace_back(c.zBinNeighborsBottom[2 * i],
                                    c.zBinNeighborsBottom[2 * i + 1]);
  }
  s.zBinEdgesF.assign(c.zBinEdges, c.zBinEdges + c.nZBinEdges);
  s.zCustomLooping.assign(c.zBinsCustomLooping,
                          c.zBinsCustomLooping + c.nZBinsCustomLooping);
  for (std::uint32_t i = 0; i < c.nRRangeMiddleSP; ++i) {
    s.rRangeMiddle.emplace_back(c.rRangeMiddleSP[2 * i],
                                c.rRangeMiddleSP[2 * i + 1]);
  }
  // the vectors now live in the Setup; never read the caller's pointers again
  s.cfg.zBinNeighborsTop = s.cfg.zBinNeighborsBottom = nullptr;
  s.cfg.zBinEdges = nullptr;
  s.cfg.zBinsCustomLooping = nullptr;
  s.cfg.rRangeMiddleSP = nullptr;

  // GridTripletSeedingAlgorithm.cpp:117-126
  for (std::size_t i : s.zCustomLooping) {
    if (i >= s.zBinEdgesF.size()) {
      throw std::invalid_argument(
          "Inconsistent config zBinsCustomLooping does not contain a subset "
          "of bins defined by zBinEdges");
    }
  }

  // CylindricalSpacePointGrid.cpp:18-37 (rMin forced to 0, .cpp:133-134)
  const float gridRMin = 0;
  if (c.phiMin < -std::numbers::pi_v<float> ||
      c.phiMax > std::numbers::pi_v<float>) {
    throw std::runtime_error(
        "CylindricalSpacePointGrid: phiMin and/or phiMax are outside the "
        "allowed phi range");
  }
  if (c.phiMin > c.phiMax) {
    throw std::runtime_error(
        "CylindricalSpacePointGrid: phiMin is bigger then phiMax");
  }
  if (gridRMin > c.rMax) {
    throw std::runtime_error(
        "CylindricalSpacePointGrid: rMin is bigger then rMax");
  }
  if (c.zMin > c.zMax) {
    throw std::runtime_error(
        "CylindricalSpacePointGrid: zMin is bigger than zMax");
  }

  s.phiBins = computePhiBins(c.minPt, c.bFieldInZ, c.rMax, c.deltaRMax,
                             c.impactMax, c.phiBinDeflectionCoverage,
                             c.maxPhiBins);
  // Axis.hpp:40-58 (equidistant, closed)
  s.phiMin = c.phiMin;
  s.phiMax = c.phiMax;
  if (s.phiMin >= s.phiMax) {
    throw std::invalid_argument("Axis: Invalid axis range");
  }
  if (s.phiBins < 1) {
    throw std::invalid_argument(
        "Axis: Invalid binning, at least one bin is needed.");
  }
  s.phiWidth = (s.phiMax - s.phiMin) / static_cast<double>(s.phiBins);

  // CylindricalSpacePointGrid.cpp:49-76
  if (s.zBinEdgesF.empty()) {
    const float zBinSize = c.cotThetaMax * c.deltaRMax;
    const float zBins =
        std::max(1.f, std::floor((c.zMax - c.zMin) / zBinSize));
    for (int bin = 0; bin <= static_cast<int>(zBins); bin++) {
      const double edge = c.zMin + bin * ((c.zMax - c.zMin) / zBins);
      s.zEdges.push_back(edge);
    }
  } else {
    for (float bin : s.zBinEdgesF) {
      s.zEdges.push_back(bin);
    }
  }
  s.rEdges = {gridRMin, c.rMax};  // rBinEdges = {} (.cpp:144, grid :78-85)
  if (s.zEdges.size() < 2 || !std::is_sorted(s.zEdges.begin(), s.zEdges.end())) {
    throw std::invalid_argument("Axis: Invalid z bin edges");
  }
  s.nZ = s.zEdges.size() - 1;
  s.nR = s.rEdges.size() - 1;
  s.nGlobal = (static_cast<std::size_t>(s.phiBins) + 2) * (s.nZ + 2) * (s.nR + 2);

  // navigation (GridTripletSeedingAlgorithm.cpp:152-154, BinnedGroup.ipp:32-62)
  s.navZ = s.zCustomLooping;
  completeNavigation(s.navPhi, static_cast<std::size_t>(s.phiBins), 0);
  completeNavigation(s.navZ, s.nZ, 1);
  completeNavigation(s.navR, s.nR, 2);

  // doublet finder configs (GridTripletSeedingAlgorithm.cpp:272-309)
  s.dRMinB = std::isnan(c.deltaRMinBottom) ? c.deltaRMin : c.deltaRMinBottom;
  s.dRMaxB = std::isnan(c.deltaRMaxBottom) ? c.deltaRMax : c.deltaRMaxBottom;
  s.dRMinT = std::isnan(c.deltaRMinTop) ? c.deltaRMin : c.deltaRMinTop;
  s.dRMaxT = std::isnan(c.deltaRMaxTop) ? c.deltaRMax : c.deltaRMaxTop;
  deriveFinderConstants(s);
  return s;
}

// ---------------------------------------------------------------------------
// grid lookup (SpacePointGridBase.hpp:67-87, Axis.hpp, MultiAxisHelper.hpp:230-245)
// ---------------------------------------------------------------------------
inline std::size_t wrapClosed(int bin, int w) {  // Axis.hpp:216-222
  return static_cast<std::size_t>(1 + (w + ((bin - 1) % w)) % w);
}

// returns nGlobal when the point is outside the grid
std::size_t binIndex(const Setup& s, float phiF, float zF, float rF) {
  const double phi = phiF, z = zF, r = rF;  // Vector3(phi, z, r), CylindricalSpacePointGrid.hpp:117-119
  // isInside on every axis (Axis.hpp:296,598-600)
  if (!((s.phiMin <= phi) && (phi < s.phiMax))) return s.nGlobal;
  if (!((s.zEdges.front() <= z) && (z < s.zEdges.back()))) return s.nGlobal;
  if (!((s.rEdges.front() <= r) && (r < s.rEdges.back()))) return s.nGlobal;
  // Axis.hpp:232-235
  const std::size_t phiBin = wrapClosed(
      static_cast<int>(std::floor((phi - s.phiMin) / s.phiWidth) + 1), s.phiBins);
  // Axis.hpp:535-539 + open wrap :497-503
  auto varBin = [](const std::vector<double>& edges, double v) {
    const auto it = std::ranges::upper_bound(edges, v);
    const int bin = static_cast<int>(std::distance(edges.begin(), it));
    const int nBins = static_cast<int>(edges.size()) - 1;
    return static_cast<std::size_t>(std::max(std::min(bin, nBins + 1), 0));
  };
  const std::size_t zBin = varBin(s.zEdges, z);
  const std::size_t rBin = varBin(s.rEdges, r);
  return (phiBin * (s.nZ + 2) + zBin) * (s.nR + 2) + rBin;
}

// neighbour local-bin lists ------------------------------------------------
// closed axis, Axis.hpp:150-188
std::vector<std::size_t> neighborsClosed(std::size_t idx, std::pair<int, int> sizes,
                                         int nBins) {
  std::vector<std::size_t> out;
  if (idx <= 0 || idx >= static_cast<std::size_t>(nBins + 1)) return out;
  const int max = nBins;
  sizes.first = std::clamp(sizes.first, -max, max);
  sizes.second = std::clamp(sizes.second, -max, max);
  if (std::abs(sizes.first - sizes.second) >= max) {
    sizes.first = 1 - static_cast<int>(idx);
    sizes.second = max - static_cast<int>(idx);
  }
  const int itmin = static_cast<int>(idx) + sizes.first;
  const int itmax = static_cast<int>(idx) + sizes.second;
  const std::size_t itfirst = wrapClosed(itmin, max);
  const std::size_t itlast = wrapClosed(itmax, max);
  if (itfirst <= itlast) {
    for (std::size_t b = itfirst; b < itlast + 1; ++b) out.push_back(b);
  } else {
    for (std::size_t b = itfirst; b < static_cast<std::size_t>(max + 1); ++b) out.push_back(b);
    for (std::size_t b = 1; b < itlast + 1; ++b) out.push_back(b);
  }
  return out;
}
// variable open axis, Axis.hpp:413-423
std::vector<std::size_t> neighborsOpen(std::size_t idx, std::pair<int, int> sizes,
                                       int nBins) {
  std::vector<std::size_t> out;
  const int itmin = std::max(0, static_cast<int>(idx) + sizes.first);
  const int itmax = std::min(nBins + 1, static_cast<int>(idx) + sizes.second);
  for (int b = itmin; b < itmax + 1; ++b) out.push_back(static_cast<std::size_t>(b));
  return out;
}

// GridBinFinder.ipp:42-76 + FlatNeighborHoodIndices (MultiAxisHelper.hpp:30-165):
// lexicographic product, phi outermost, r innermost.
std::vector<std::size_t> findBins(const Setup& s, std::size_t phiLoc, std::size_t zLoc,
                                  std::size_t rLoc, bool top) {
  const auto& zNb = top ? s.zNeighborsTop : s.zNeighborsBottom;
  const std::pair<int, int> phiSize{-s.cfg.numPhiNeighbors, s.cfg.numPhiNeighbors};
  std::pair<int, int> zSize{-1, 1};  // empty vector -> int 1 (GridBinFinder.ipp:30-34)
  if (!zNb.empty()) {
    zSize = zNb.at(zLoc - 1);
  }
  const std::pair<int, int> rSize{0, 0};
  const auto phis = neighborsClosed(phiLoc, phiSize, s.phiBins);
  const auto zs = neighborsOpen(zLoc, zSize, static_cast<int>(s.nZ));
  const auto rs = neighborsOpen(rLoc, rSize, static_cast<int>(s.nR));
  std::vector<std::size_t> out;
  for (auto p : phis)
    for (auto z : zs)
      for (auto r : rs) out.push_back((p * (s.nZ + 2) + z) * (s.nR + 2) + r);
  return out;
}

// ---------------------------------------------------------------------------
// per-event data
// ---------------------------------------------------------------------------
struct Packed {  // reference coreSpacePoints (GridTripletSeedingAlgorithm.cpp:230-253)
  std::vector<Index> copiedFrom;
  std::vector<float> x, y, z, r, varZ, varR;
  std::vector<std::pair<Index, Index>> binRange;  // per global bin
};

struct Doublets {  // DoubletSeedFinder.hpp:26-262
  std::vector<Index> sp;
  std::vector<float> cotTheta, er, iDeltaR, u, v, x, y;
  void clear() {
    sp.clear(); cotTheta.clear(); er.clear(); iDeltaR.clear();
    u.clear(); v.clear(); x.clear(); y.clear();
  }
  std::size_t size() const { return sp.size(); }
  bool empty() const { return sp.empty(); }
};
struct IndexAndCotTheta {
  Index index{};
  float cotTheta{};
};

struct TopCandidates {  // TripletSeedFinder.hpp:24-124
  std::vector<Index> top;
  std::vector<float> curvature, impact;
  void clear() { top.clear(); curvature.clear(); impact.clear(); }
  std::size_t size() const { return top.size(); }
};

struct TripletCandidate {  // detail/CandidatesForMiddleSp.hpp:19-40
  Index bottom{}, middle{}, top{};
  float weight{}, zOrigin{};
  bool isQuality{};
};

// detail/CandidatesForMiddleSp.cpp:31-93
class Collector {
 public:
  using WeightIndex = std::pair<float, std::uint32_t>;
  static constexpr bool comparator(const WeightIndex& a, const WeightIndex& b) {
    return a.first > b.first;
  }
  void configure(std::uint32_t nLow, std::uint32_t nHigh) {
    m_nLow = nLow;
    m_nHigh = nHigh;
  }
  std::uint32_t nHigh() const { return static_cast<std::uint32_t>(m_high.size()); }
  void clear() { m_storage.clear(); m_low.clear(); m_high.clear(); }
  bool push(Index b, Index m, Index t, float w, float z, bool q) {
    return q ? push(m_high, m_nHigh, b, m, t, w, z, q)
             : push(m_low, m_nLow, b, m, t, w, z, q);
  }
  void toSorted(std::vector<TripletCandidate>& out) {
    out.clear();
    std::ranges::sort_heap(m_high, comparator);
    std::ranges::sort_heap(m_low, comparator);
    for (const auto& [w, i] : m_high) out.push_back(m_storage[i]);
    for (const auto& [w, i] : m_low) out.push_back(m_storage[i]);
    clear();
  }

 private:
  bool push(std::vector<WeightIndex>& cont, std::uint32_t nMax, Index b, Index m,
            Index t, float w, float z, bool q) {
    if (nMax == 0) return false;
    if (cont.size() < nMax) {
      m_storage.push_back({b, m, t, w, z, q});
      cont.emplace_back(w, static_cast<std::uint32_t>(m_storage.size() - 1));
      std::ranges::push_heap(cont, comparator);
      return true;
    }
    const auto [smallestWeight, smallestIndex] = cont.front();
    if (w <= smallestWeight) return false;
    m_storage[smallestIndex] = TripletCandidate{b, m, t, w, z, q};
    std::ranges::pop_heap(cont, comparator);
    cont.back() = {w, smallestIndex};
    std::ranges::push_heap(cont, comparator);
    return true;
  }
  std::uint32_t m_nLow = 0, m_nHigh = 0;
  std::vector<TripletCandidate> m_storage;
  std::vector<WeightIndex> m_low, m_high;
};

struct Counters {
  std::uint64_t nInGrid = 0, nMiddles = 0, nPairTests = 0, nBottomDoublets = 0,
                nTopDoublets = 0, nTripletTests = 0, nCandidates = 0, nSeeds = 0;
  std::uint64_t nRTieBins = 0, nCotTieMiddles = 0, nCurvTieGroups = 0,
                nWeightTieMiddles = 0, maxBottoms = 0, maxTops = 0,
                maxCandidatesPerBottom = 0, maxCandidatesPerMiddle = 0,
                maxBinSize = 0;
  // per-middle size histograms (bucket width 128, last bucket = overflow)
  std::uint64_t histBottoms[32] = {0}, histTops[32] = {0}, histCandRound[32] = {0};
};

struct Seed {
  Index b, m, t;
  float quality, vertexZ;
};

struct DoubletDump {  // optional per-middle record for stage-level parity tests
  std::vector<Index> middlePos;
  std::vector<std::uint64_t> first;  // prefix
  std::vector<Index> nBottom;
  std::vector<Index> otherPos;
  std::vector<float> cotTheta, iDeltaR, er, u, v, x, y;
};

enum SortMode : int {
  kFaithful = 0,  // std::ranges::sort like the reference (unstable, libstdc++)
  kStable = 1     // order by (key, position): the canonical parallel order
};

struct Event {
  const Setup* s = nullptr;
  const float *x = nullptr, *y = nullptr, *z = nullptr, *r = nullptr,
              *varZ = nullptr, *varR = nullptr;
  Index n = 0;
  ZWindows zw;
  int sortMode = kFaithful;
  const float* phiOverride = nullptr;  // optional precomputed phi (tests)
  // strip triplet path (TripletSeedFinder::Config::useStripInfo = true): 12 floats per ORIGINAL space point =
  // outerCenter, innerToOuterSeparation, outerHalfVector, innerHalfVector
  // (StripSpacePointCalibrationDetails.hpp:16-29); NULL = pixel path
  const float* strip = nullptr;
  float cotThetaDiffMax = std::numeric_limits<float>::infinity();  // TripletSeedFinder.hpp:175
  // bounded sampling for the timed baseline: only navigation entries g with
  // g % navStride == navPhase are seeded (the grid is always built in full)
  std::uint32_t navStride = 1, navPhase = 0;

  Packed sp;
  Counters cnt;
  std::vector<Seed> seeds;
  DoubletDump* dump = nullptr;

  // per-thread scratch (TripletSeeder::Cache, BroadTripletSeedFilter::Cache)
  Doublets bottomDoublets, topDoublets;
  std::vector<IndexAndCotTheta> sortedBottoms, sortedTops;
  TopCandidates topCandidates;
  std::vector<std::size_t> topSpIndexVec;
  std::vector<float> compatibleSeedR;
  std::vector<TripletCandidate> sortedCandidates;
  Collector collector;
  float rMaxSeedConf = 0;
  std::unordered_map<Index, float> bestSeedQualityMap;
};

// itkFastTrackingSPselect, GridTripletSeedingAlgorithm.cpp:46-62
bool itkFastTrackingSPselect(float rIn, float zIn) {
  float r = rIn;
  float zabs = std::abs(zIn);
  if (zabs > 200. && r < 45.) {
    return false;
  }
  float cotTheta = 27.2899;
  if ((zabs - 150.) > cotTheta * r) {
    return false;
  }
  return true;
}

// experiment cuts on doublets: VertexZCuts (.cpp:69-97) takes the slot when
// windows are given, else itkFastTrackingCuts (.cpp:33-44) with useExtraCuts.
enum class DoubletCuts { None, Itk, VertexZ };

template <typename Key>
void sortIndices(int mode, std::vector<std::size_t>& idx, Key key) {
  if (mode == kFaithful) {
    std::ranges::sort(idx, {}, key);
  } else {
    std::stable_sort(idx.begin(), idx.end(),
                     [&](std::size_t a, std::size_t b) { return key(a) < key(b); });
  }
}

// GridTripletSeedingAlgorithm.cpp:208-253
void buildGrid(Event& ev) {
  const Setup& s = *ev.s;
  std::vector<std::vector<Index>> bins(s.nGlobal);
  for (std::size_t i = 0; i < ev.n; ++i) {
    if (s.cfg.useExtraCuts && !itkFastTrackingSPselect(ev.r[i], ev.z[i])) {
      continue;
    }
    float phi = ev.phiOverride != nullptr ? ev.phiOverride[i]
                                          : std::atan2(ev.y[i], ev.x[i]);
    const std::size_t b = binIndex(s, phi, ev.z[i], ev.r[i]);
    if (b < s.nGlobal) {
      bins[b].push_back(static_cast<Index>(i));
      ++ev.cnt.nInGrid;
    }
  }
  for (auto& bin : bins) {
    if (ev.sortMode == kFaithful) {
      std::ranges::sort(bin, [&](const Index& a, const Index& b) {
        return ev.r[a] < ev.r[b];
      });
    } else {
      std::stable_sort(bin.begin(), bin.end(), [&](const Index& a, const Index& b) {
        return ev.r[a] < ev.r[b];
      });
    }
    bool tie = false;
    for (std::size_t k = 1; k < bin.size(); ++k) {
      tie |= ev.r[bin[k]] == ev.r[bin[k - 1]];
    }
    ev.cnt.nRTieBins += tie ? 1 : 0;
    ev.cnt.maxBinSize = std::max<std::uint64_t>(ev.cnt.maxBinSize, bin.size());
  }
  Packed& p = ev.sp;
  p.binRange.reserve(s.nGlobal);
  for (const auto& bin : bins) {
    const Index begin = static_cast<Index>(p.copiedFrom.size());
    for (Index i : bin) {
      p.copiedFrom.push_back(i);
      p.x.push_back(ev.x[i]);
      p.y.push_back(ev.y[i]);
      p.z.push_back(ev.z[i]);
      p.r.push_back(ev.r[i]);
      p.varZ.push_back(ev.varZ[i]);
      p.varR.push_back(ev.varR[i]);
    }
    p.binRange.emplace_back(begin, static_cast<Index>(p.copiedFrom.size()));
  }
}

struct MiddleInfo {  // DoubletSeedFinder.cpp:359-368
  float uIP, uIP2, cosPhiM, sinPhiM;
};
MiddleInfo computeMiddleInfo(const Packed& p, Index m) {
  const float rM = p.r[m];
  const float uIP = -1 / rM;
  const float cosPhiM = -p.x[m] * uIP;
  const float sinPhiM = -p.y[m] * uIP;
  const float uIP2 = uIP * uIP;
  return {uIP, uIP2, cosPhiM, sinPhiM};
}

inline bool outsideRange(float value, float min, float max) {
  return static_cast<bool>(static_cast<int>(value < min) |
                           static_cast<int>(value > max));
}

bool doubletExperimentCut(const Event& ev, DoubletCuts kind, Index m, Index o,
                          float cotTheta, bool isBottom) {
  const Packed& p = ev.sp;
  if (kind == DoubletCuts::Itk) {  // .cpp:33-44
    const float rMin = 45;
    const float cotThetaMax = 1.5;
    if (isBottom && p.r[o] < rMin &&
        (cotTheta > cotThetaMax || cotTheta < -cotThetaMax)) {
      return false;
    }
    return true;
  }
  // VertexZCuts::operator(), .cpp:78-96
  if (ev.zw.n == 0) return true;
  const float zM = p.z[m];
  const float rM = p.r[m];
  const float zOrigin = zM - rM * cotTheta;
  for (std::uint32_t k = 0; k < ev.zw.n; ++k) {
    if (zOrigin >= ev.zw.lo[k] && zOrigin <= ev.zw.hi[k]) return true;
  }
  return false;
}

// One candidate of DoubletSeedFinder.cpp:41-273 after the deltaR window logic: the cuts of :137-271 and the
// coordinate transform; appends to `out` when the candidate survives.
template <bool isBottom>
inline void doubletCandidate(Event& ev, DoubletCuts cuts, Index m, const MiddleInfo& mi, Index o,
                             float deltaR, Doublets& out) {
  const Setup& s = *ev.s;
  const b200seed_config& c = s.cfg;
  const Packed& p = ev.sp;
  const float impactMax = isBottom ? -c.impactMax : c.impactMax;
  const float xM = p.x[m];
  const float yM = p.y[m];
  const float zM = p.z[m];
  const float rM = p.r[m];
  const float varianceZM = p.varZ[m];
  const float varianceRM = p.varR[m];
  const float vIPAbs = impactMax * mi.uIP2;
  const auto calculateError = [&](float varianceZO, float varianceRO,
                                  float iDeltaR2, float cotTheta) {
    return iDeltaR2 * ((varianceZM + varianceZO) +
                       (cotTheta * cotTheta) * (varianceRM + varianceRO));
  };
  const float xO = p.x[o];
  const float yO = p.y[o];
  const float zO = p.z[o];
  const float varianceZO = p.varZ[o];
  const float varianceRO = p.varR[o];
  ++ev.cnt.nPairTests;

  float deltaZ = 0;
  if constexpr (isBottom) {
    deltaZ = zM - zO;
  } else {
    deltaZ = zO - zM;
  }
  if (outsideRange(deltaZ, c.deltaZMin, c.deltaZMax)) return;

  const float zOriginTimesDeltaR = zM * deltaR - rM * deltaZ;
  if (outsideRange(zOriginTimesDeltaR, c.collisionRegionMin * deltaR,
                   c.collisionRegionMax * deltaR)) {
    return;
  }

  if (!c.interactionPointCut) {
    if (outsideRange(deltaZ, -c.cotThetaMax * deltaR, c.cotThetaMax * deltaR)) {
      return;
    }
    const float deltaX = xO - xM;
    const float deltaY = yO - yM;
    const float xNewFrame = deltaX * mi.cosPhiM + deltaY * mi.sinPhiM;
    const float yNewFrame = deltaY * mi.cosPhiM - deltaX * mi.sinPhiM;
    const float deltaR2 = deltaX * deltaX + deltaY * deltaY;
    const float iDeltaR2 = 1 / deltaR2;
    const float uT = xNewFrame * iDeltaR2;
    const float vT = yNewFrame * iDeltaR2;
    const float iDeltaR = std::sqrt(iDeltaR2);
    const float cotTheta = deltaZ * iDeltaR;
    if (cuts != DoubletCuts::None) {
      if (!doubletExperimentCut(ev, cuts, m, o, cotTheta, isBottom)) return;
    }
    const float er = calculateError(varianceZO, varianceRO, iDeltaR2, cotTheta);
    out.sp.push_back(o); out.cotTheta.push_back(cotTheta);
    out.er.push_back(er); out.iDeltaR.push_back(iDeltaR);
    out.u.push_back(uT); out.v.push_back(vT);
    out.x.push_back(xNewFrame); out.y.push_back(yNewFrame);
    return;
  }

  // interactionPointCut == true, :205-271
  const float deltaX = xO - xM;
  const float deltaY = yO - yM;
  const float xNewFrame = deltaX * mi.cosPhiM + deltaY * mi.sinPhiM;
  const float yNewFrame = deltaY * mi.cosPhiM - deltaX * mi.sinPhiM;
  const float deltaR2 = deltaX * deltaX + deltaY * deltaY;
  const float iDeltaR2 = 1 / deltaR2;
  const float uT = xNewFrame * iDeltaR2;
  const float vT = yNewFrame * iDeltaR2;
  if (std::abs(rM * yNewFrame) > impactMax * xNewFrame) {
    const float vIP = (yNewFrame > 0) ? -vIPAbs : vIPAbs;
    const float aCoef = (vT - vIP) / (uT - mi.uIP);
    const float bCoef = vIP - aCoef * mi.uIP;
    if ((bCoef * bCoef) * s.minHelixDiameter2Doublet > 1 + aCoef * aCoef) {
      return;
    }
  }
  if (outsideRange(deltaZ, -c.cotThetaMax * deltaR, c.cotThetaMax * deltaR)) {
    return;
  }
  const float iDeltaR = std::sqrt(iDeltaR2);
  const float cotTheta = deltaZ * iDeltaR;
  if (cuts != DoubletCuts::None) {
    if (!doubletExperimentCut(ev, cuts, m, o, cotTheta, isBottom)) return;
  }
  const float er = calculateError(varianceZO, varianceRO, iDeltaR2, cotTheta);
  out.sp.push_back(o); out.cotTheta.push_back(cotTheta);
  out.er.push_back(er); out.iDeltaR.push_back(iDeltaR);
  out.u.push_back(uT); out.v.push_back(vT);
  out.x.push_back(xNewFrame); out.y.push_back(yNewFrame);
}

// DoubletSeedFinder.cpp:41-273 for sortedByR = true.  [begin,end) is the
// caller's persistent candidate range; begin is advanced like
// `candidateSps = candidateSps.subrange(offset)`.
template <bool isBottom>
void createDoublets(Event& ev, DoubletCuts cuts, Index m, const MiddleInfo& mi,
                    Index& begin, Index end, Doublets& out) {
  const Setup& s = *ev.s;
  const Packed& p = ev.sp;
  const float cfgDeltaRMin = isBottom ? s.dRMinB : s.dRMinT;
  const float cfgDeltaRMax = isBottom ? s.dRMaxB : s.dRMaxT;
  const float rM = p.r[m];

  // :73-93
  {
    Index offset = 0;
    for (Index o = begin; o < end; ++o) {
      if constexpr (isBottom) {
        if (rM - p.r[o] <= cfgDeltaRMax) break;
      } else {
        if (p.r[o] - rM >= cfgDeltaRMin) break;
      }
      ++offset;
    }
    begin += offset;
  }

  for (Index o = begin; o < end; ++o) {
    const float rO = p.r[o];
    float deltaR = 0;
    if constexpr (isBottom) {
      deltaR = rM - rO;
      if (deltaR < cfgDeltaRMin) break;
    } else {
      deltaR = rO - rM;
      if (deltaR > cfgDeltaRMax) break;
    }
    doubletCandidate<isBottom>(ev, cuts, m, mi, o, deltaR, out);
  }
}

// DoubletSeedFinder.cpp:41-273 for sortedByR = false (spacePointsSortedByRadius = false, the orthogonal
// seeder): every candidate of the subset is tested, the deltaR window is an explicit cut (:126-130).
template <bool isBottom>
void createDoubletsUnsorted(Event& ev, DoubletCuts cuts, Index m, const MiddleInfo& mi,
                            const std::vector<Index>& candidates, Doublets& out) {
  const Setup& s = *ev.s;
  const Packed& p = ev.sp;
  const float cfgDeltaRMin = isBottom ? s.dRMinB : s.dRMinT;
  const float cfgDeltaRMax = isBottom ? s.dRMaxB : s.dRMaxT;
  const float rM = p.r[m];
  for (Index o : candidates) {
    const float rO = p.r[o];
    float deltaR = 0;
    if constexpr (isBottom) {
      deltaR = rM - rO;
    } else {
      deltaR = rO - rM;
    }
    if (outsideRange(deltaR, cfgDeltaRMin, cfgDeltaRMax)) continue;
    doubletCandidate<isBottom>(ev, cuts, m, mi, o, deltaR, out);
  }
}

// DoubletSeedFinder.hpp:94-104
void sortByCotTheta(const Event& ev, const Doublets& d,
                    std::vector<IndexAndCotTheta>& out) {
  out.clear();
  out.reserve(d.size());
  for (Index i = 0; i < d.size(); ++i) out.push_back({i, d.cotTheta[i]});
  if (ev.sortMode == kFaithful) {
    std::ranges::sort(out, {}, [](const IndexAndCotTheta& item) { return item.cotTheta; });
  } else {
    std::stable_sort(out.begin(), out.end(),
                     [](const IndexAndCotTheta& a, const IndexAndCotTheta& b) {
                       return a.cotTheta < b.cotTheta;
                     });
  }
}

// TripletSeedFinder.cpp:34-162 (pixel path, sortedByCotTheta = true).
// [topBegin, topEnd) indexes ev.sortedTops; topBegin is advanced like
// `topDoublets = topDoublets.subrange(topDoubletOffset)`.
void createTripletTopCandidates(Event& ev, Index m, Index bottomDoublet,
                                std::size_t& topBegin, std::size_t topEnd) {
  const Setup& s = *ev.s;
  const Packed& p = ev.sp;
  const Doublets& B = ev.bottomDoublets;
  const Doublets& T = ev.topDoublets;
  TopCandidates& out = ev.topCandidates;

  const float rM = p.r[m];
  const float varianceZM = p.varZ[m];
  const float varianceRM = p.varR[m];

  const float cotThetaB = B.cotTheta[bottomDoublet];
  const float erB = B.er[bottomDoublet];
  const float iDeltaRB = B.iDeltaR[bottomDoublet];
  const float Ub = B.u[bottomDoublet];
  const float Vb = B.v[bottomDoublet];

  const float iSinTheta2 = 1 + cotThetaB * cotThetaB;
  const float sigmaSquaredPtDependent = iSinTheta2 * s.sigmapT2perRadius;
  const float scatteringInRegion2 = s.multipleScattering2 * iSinTheta2;

  std::size_t topDoubletOffset = 0;
  for (std::size_t k = topBegin; k < topEnd; ++k) {
    const std::size_t topDoubletIndex = k - topBegin;
    const Index t = ev.sortedTops[k].index;
    const Index spT = T.sp[t];
    const float cotThetaT = T.cotTheta[t];
    ++ev.cnt.nTripletTests;

    const float cotThetaAvg2 = cotThetaB * cotThetaT;
    const float error2 = T.er[t] + erB +
                         2 * (cotThetaAvg2 * varianceRM + varianceZM) *
                             iDeltaRB * T.iDeltaR[t];
    const float deltaCotTheta = cotThetaB - cotThetaT;
    const float deltaCotTheta2 = deltaCotTheta * deltaCotTheta;

    if (deltaCotTheta2 > error2 + scatteringInRegion2) {
      if (cotThetaB < cotThetaT) break;
      topDoubletOffset = topDoubletIndex + 1;
      continue;
    }
    const float dU = T.u[t] - Ub;
    if (dU == 0) continue;
    const float A = (T.v[t] - Vb) / dU;
    const float S2 = 1 + A * A;
    const float Bc = Vb - A * Ub;
    const float B2 = Bc * Bc;
    if (S2 < B2 * s.minHelixDiameter2) continue;
    const float iHelixDiameter2 = B2 / S2;
    const float p2scatterSigma = iHelixDiameter2 * sigmaSquaredPtDependent;
    if (deltaCotTheta2 > error2 + p2scatterSigma) {
      if (cotThetaB < cotThetaT) break;
      topDoubletOffset = topDoubletIndex;
      continue;
    }
    const float im = std::abs((A - Bc * rM) * rM);
    if (im > s.cfg.impactMax) continue;
    out.top.push_back(spT);
    out.curvature.push_back(Bc / std::sqrt(S2));
    out.impact.push_back(im);
  }
  topBegin += topDoubletOffset;
}

// ---------------------------------------------------------------------------
// Strip triplet path.
// ---------------------------------------------------------------------------
using Vec3 = std::array<float, 3>;
// Utilities/detail/StdArrayLinalg.hpp:57-83
inline float dot3(const Vec3& a, const Vec3& b) {
  float result = 0;
  for (std::size_t i = 0; i < 3; ++i) result += a[i] * b[i];
  return result;
}
inline Vec3 cross3(const Vec3& a, const Vec3& b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
// StripSpacePointCalibrationDetails.hpp:33-49 + detail/StripSpacePointCalibrationImpl.hpp:22-42
struct StripDerived {
  Vec3 iosvCrossIhv, iosvCrossOhv, ihvCrossOhv, oc, ohv;
};
inline StripDerived deriveStrip(const float* d) {
  const Vec3 oc{d[0], d[1], d[2]}, iosv{d[3], d[4], d[5]}, ohv{d[6], d[7], d[8]}, ihv{d[9], d[10], d[11]};
  StripDerived out;
  out.ihvCrossOhv = cross3(ihv, ohv);
  out.iosvCrossOhv = cross3(iosv, ohv);
  out.iosvCrossIhv = cross3(iosv, ihv);
  out.oc = oc;
  out.ohv = ohv;
  return out;
}
// detail/StripSpacePointCalibrationImpl.hpp:44-85
inline bool calibrateStrip(const Vec3& direction, const StripDerived& sp, Vec3& calibrated, float tolerance) {
  const float scale = dot3(direction, sp.ihvCrossOhv);
  const float sInner = dot3(direction, sp.iosvCrossOhv);
  if (std::abs(sInner) > std::abs(scale) * tolerance) return false;
  const float sOuter = dot3(direction, sp.iosvCrossIhv);
  if (std::abs(sOuter) > std::abs(scale) * tolerance) return false;
  const float sOuterNorm = sOuter / scale;
  for (std::size_t i = 0; i < 3; ++i) calibrated[i] = sp.oc[i] + sp.ohv[i] * sOuterNorm;
  return true;
}

// TripletSeedFinder.cpp:164-406 (strip path, sortedByCotTheta = true); window handling as above.
void createStripTripletTopCandidates(Event& ev, Index m, Index bottomDoublet,
                                     std::size_t& topBegin, std::size_t topEnd) {
  const Setup& s = *ev.s;
  const Packed& p = ev.sp;
  const Doublets& B = ev.bottomDoublets;
  const Doublets& T = ev.topDoublets;
  TopCandidates& out = ev.topCandidates;

  const float rM = p.r[m];
  const float cosPhiM = p.x[m] / rM;
  const float sinPhiM = p.y[m] / rM;
  const float varianceZM = p.varZ[m];
  const float varianceRM = p.varR[m];

  const float cotThetaB0 = B.cotTheta[bottomDoublet];
  const float erB = B.er[bottomDoublet];
  const float iDeltaRB = B.iDeltaR[bottomDoublet];
  const float Ub0 = B.u[bottomDoublet];
  const float Vb0 = B.v[bottomDoublet];

  const float iSinTheta2 = 1 + cotThetaB0 * cotThetaB0;
  const float sigmaSquaredPtDependent = iSinTheta2 * s.sigmapT2perRadius;
  const float scatteringInRegion2 = s.multipleScattering2 * iSinTheta2;

  const float sinTheta = 1 / std::sqrt(iSinTheta2);
  const float cosTheta = cotThetaB0 * sinTheta;
  const std::array<float, 2> rot = {cosPhiM * sinTheta, sinPhiM * sinTheta};

  const StripDerived calM = deriveStrip(ev.strip + 12u * p.copiedFrom[m]);
  const StripDerived calB = deriveStrip(ev.strip + 12u * p.copiedFrom[B.sp[bottomDoublet]]);

  std::size_t topDoubletOffset = 0;
  for (std::size_t k = topBegin; k < topEnd; ++k) {
    const std::size_t topDoubletIndex = k - topBegin;
    const Index t = ev.sortedTops[k].index;
    {  // :226-238
      const float cotThetaT = T.cotTheta[t];
      const float deltaCotTheta = cotThetaB0 - cotThetaT;
      const float cotThetaDiffMax2 = ev.cotThetaDiffMax * ev.cotThetaDiffMax;
      if (deltaCotTheta * deltaCotTheta > cotThetaDiffMax2) {
        if (cotThetaB0 < cotThetaT) break;
        topDoubletOffset = topDoubletIndex + 1;
        continue;
      }
    }
    ++ev.cnt.nTripletTests;
    const float dU0 = T.u[t] - Ub0;
    if (dU0 == 0) continue;
    const float A0 = (T.v[t] - Vb0) / dU0;

    const Vec3 directionMiddle = {rot[0] - rot[1] * A0, rot[0] * A0 + rot[1], cosTheta};
    Vec3 rMTransf{};
    if (!calibrateStrip(directionMiddle, calM, rMTransf, s.cfg.toleranceParam)) continue;

    const float zDirectionMiddle = cosTheta * std::sqrt(1 + A0 * A0);

    const float B0 = 2 * (Vb0 - A0 * Ub0);
    const float Cb = 1 - B0 * B.y[bottomDoublet];
    const float Sb = A0 + B0 * B.x[bottomDoublet];
    const Vec3 directionBottom = {rot[0] * Cb - rot[1] * Sb, rot[0] * Sb + rot[1] * Cb, zDirectionMiddle};
    Vec3 rBTransf{};
    if (!calibrateStrip(directionBottom, calB, rBTransf, s.cfg.toleranceParam)) continue;

    const float Ct = 1 - B0 * T.y[t];
    const float St = A0 + B0 * T.x[t];
    const Vec3 directionTop = {rot[0] * Ct - rot[1] * St, rot[0] * St + rot[1] * Ct, zDirectionMiddle};
    const StripDerived calT = deriveStrip(ev.strip + 12u * p.copiedFrom[T.sp[t]]);
    Vec3 rTTransf{};
    if (!calibrateStrip(directionTop, calT, rTTransf, s.cfg.toleranceParam)) continue;

    const float xB = rBTransf[0] - rMTransf[0];
    const float yB = rBTransf[1] - rMTransf[1];
    const float zB = rBTransf[2] - rMTransf[2];
    const float xT = rTTransf[0] - rMTransf[0];
    const float yT = rTTransf[1] - rMTransf[1];
    const float zT = rTTransf[2] - rMTransf[2];

    const float iDeltaRB2 = 1 / (xB * xB + yB * yB);
    const float iDeltaRT2 = 1 / (xT * xT + yT * yT);

    const float cotThetaB = -zB * std::sqrt(iDeltaRB2);
    const float cotThetaT = zT * std::sqrt(iDeltaRT2);

    const float averageCotTheta = 0.5f * (cotThetaB + cotThetaT);
    const float cotThetaAvg2 = averageCotTheta * averageCotTheta;

    const float error2 = T.er[t] + erB +
                         2 * (cotThetaAvg2 * varianceRM + varianceZM) *
                             iDeltaRB * T.iDeltaR[t];

    const float deltaCotTheta = cotThetaB - cotThetaT;
    const float deltaCotTheta2 = deltaCotTheta * deltaCotTheta;
    if (deltaCotTheta2 > error2 + scatteringInRegion2) continue;

    const float rMxy = std::sqrt(rMTransf[0] * rMTransf[0] + rMTransf[1] * rMTransf[1]);
    const float irMxy = 1 / rMxy;
    const float Ax = rMTransf[0] * irMxy;
    const float Ay = rMTransf[1] * irMxy;

    const float Ub = (xB * Ax + yB * Ay) * iDeltaRB2;
    const float Vb = (yB * Ax - xB * Ay) * iDeltaRB2;
    const float Ut = (xT * Ax + yT * Ay) * iDeltaRT2;
    const float Vt = (yT * Ax - xT * Ay) * iDeltaRT2;

    const float dU = Ut - Ub;
    if (dU == 0) continue;
    const float A = (Vt - Vb) / dU;
    const float S2 = 1 + A * A;
    const float Bc = Vb - A * Ub;
    const float B2 = Bc * Bc;
    if (S2 < B2 * s.minHelixDiameter2) continue;

    const float iHelixDiameter2 = B2 / S2;
    const float p2scatterSigma = iHelixDiameter2 * sigmaSquaredPtDependent;
    if (deltaCotTheta2 > error2 + p2scatterSigma) continue;

    const float im = std::abs((A - Bc * rMxy) * rMxy);
    if (im > s.cfg.impactMax) continue;

    out.top.push_back(T.sp[t]);
    out.curvature.push_back(Bc / std::sqrt(S2));
    out.impact.push_back(im);
  }
  topBegin += topDoubletOffset;
}

float getBestSeedQuality(const std::unordered_map<Index, float>& map, Index sp) {
  auto it = map.find(sp);
  if (it != map.end()) return it->second;
  return std::numeric_limits<float>::lowest();
}
void setBestSeedQuality(std::unordered_map<Index, float>& map, Index bottom,
                        Index middle, Index top, float quality) {
  for (Index sp : {top, middle, bottom}) {
    auto it = map.find(sp);
    if (it != map.end()) {
      it->second = std::max(quality, it->second);
    } else {
      map.emplace(sp, quality);
    }
  }
}

// BroadTripletSeedFilter.cpp:63-94
bool sufficientTopDoublets(Event& ev, Index m) {
  const b200seed_config& c = ev.s->cfg;
  if (!c.seedConfirmation) return true;
  const Packed& p = ev.sp;
  const bool isForwardRegion =
      p.z[m] > c.centralSeedConfirmationRange.zMaxSeedConf ||
      p.z[m] < c.centralSeedConfirmationRange.zMinSeedConf;
  const b200seed_seed_confirmation_range& range =
      isForwardRegion ? c.forwardSeedConfirmationRange
                      : c.centralSeedConfirmationRange;
  std::size_t nTopSeedConf =
      p.r[m] > range.rMaxSeedConf ? range.nTopForLargeR : range.nTopForSmallR;
  ev.rMaxSeedConf = range.rMaxSeedConf;
  return !(ev.topDoublets.size() < nTopSeedConf);
}

// BroadTripletSeedFilter.cpp:96-322
void filterTripletTopCandidates(Event& ev, Index m, Index bottomDoublet) {
  const b200seed_config& c = ev.s->cfg;
  const Packed& p = ev.sp;
  const TopCandidates& cand = ev.topCandidates;
  const Index spB = ev.bottomDoublets.sp[bottomDoublet];
  const float cotThetaB = ev.bottomDoublets.cotTheta[bottomDoublet];

  std::size_t minCompatibleTopSPs = 2;
  if (!c.seedConfirmation || p.r[spB] > ev.rMaxSeedConf) {
    minCompatibleTopSPs = 1;
  }
  if (c.seedConfirmation && ev.collector.nHigh() > 0) {
    minCompatibleTopSPs++;
  }
  if (cand.size() < minCompatibleTopSPs) return;
  float zOrigin = p.z[m] - p.r[m] * cotThetaB;

  b200seed_seed_confirmation_range seedConfRange{};
  std::size_t nTopSeedConf = 0;
  if (c.seedConfirmation) {
    const bool isForwardRegion =
        p.z[spB] > c.centralSeedConfirmationRange.zMaxSeedConf ||
        p.z[spB] < c.centralSeedConfirmationRange.zMinSeedConf;
    seedConfRange = isForwardRegion ? c.forwardSeedConfirmationRange
                                    : c.centralSeedConfirmationRange;
    nTopSeedConf = p.r[spB] > seedConfRange.rMaxSeedConf
                       ? seedConfRange.nTopForLargeR
                       : seedConfRange.nTopForSmallR;
  }

  std::size_t maxWeightTopSp = 0;
  bool maxWeightSeed = false;
  float weightMax = std::numeric_limits<float>::lowest();

  auto& order = ev.topSpIndexVec;
  order.resize(cand.size());
  std::iota(order.begin(), order.end(), 0);
  sortIndices(ev.sortMode, order,
              [&cand](const std::size_t t) { return cand.curvature[t]; });
  {
    bool tie = false;
    for (std::size_t k = 1; k < order.size(); ++k) {
      tie |= cand.curvature[order[k]] == cand.curvature[order[k - 1]];
    }
    ev.cnt.nCurvTieGroups += tie ? 1 : 0;
  }

  auto& compatibleSeedR = ev.compatibleSeedR;
  const auto getTopR = [&](Index spT) {
    if (c.useDeltaRinsteadOfTopRadius) {
      // fastHypot(dr, dz) = sqrt(dr*dr + dz*dz), MathHelpers.hpp:86-103
      const float dr = p.r[spT] - p.r[m];
      const float dz = p.z[spT] - p.z[m];
      return std::sqrt(dr * dr + dz * dz);
    }
    return p.r[spT];
  };

  std::size_t beginCompTopIndex = 0;
  for (const std::size_t topSpIndex : order) {
    const Index spT = cand.top[topSpIndex];
    compatibleSeedR.clear();

    float invHelixDiameter = cand.curvature[topSpIndex];
    float lowerLimitCurv = invHelixDiameter - c.deltaInvHelixDiameter;
    float upperLimitCurv = invHelixDiameter + c.deltaInvHelixDiameter;
    float currentTopR = getTopR(spT);
    float impact = cand.impact[topSpIndex];

    float weight = -impact * c.impactWeightFactor;

    for (std::size_t variableCompTopIndex = beginCompTopIndex;
         variableCompTopIndex < order.size(); variableCompTopIndex++) {
      std::size_t compatibleTopSpIndex = order[variableCompTopIndex];
      if (compatibleTopSpIndex == topSpIndex) continue;
      float otherTopR = getTopR(cand.top[compatibleTopSpIndex]);
      if (cand.curvature[compatibleTopSpIndex] < lowerLimitCurv) {
        beginCompTopIndex = variableCompTopIndex + 1;
        continue;
      }
      if (cand.curvature[compatibleTopSpIndex] > upperLimitCurv) break;
      float deltaR = currentTopR - otherTopR;
      if (std::abs(deltaR) < c.deltaRMin) continue;
      bool newCompSeed = true;
      for (const float previousDiameter : compatibleSeedR) {
        if (std::abs(previousDiameter - otherTopR) < c.deltaRMin) {
          newCompSeed = false;
          break;
        }
      }
      if (newCompSeed) {
        compatibleSeedR.push_back(otherTopR);
        weight += c.compatSeedWeight;
      }
      if (compatibleSeedR.size() >= c.compatSeedLimit) break;
    }

    // experimentCuts (ITripletSeedCuts) are never set by this algorithm.

    if (compatibleSeedR.size() > c.numSeedIncrement) {
      weight += c.seedWeightIncrement;
    }
    // absDeltaEtaWeightFactor keeps its default 0 (never set by the algorithm).

    if (c.seedConfirmation) {
      int deltaSeedConf = compatibleSeedR.size() + 1 - nTopSeedConf;
      if (deltaSeedConf < 0 || (ev.collector.nHigh() != 0 && deltaSeedConf == 0)) {
        continue;
      }
      bool seedRangeCuts = p.r[spB] < seedConfRange.seedConfMinBottomRadius ||
                           std::abs(zOrigin) > seedConfRange.seedConfMaxZOrigin;
      if (seedRangeCuts && deltaSeedConf == 0 &&
          impact > seedConfRange.minImpactSeedConf) {
        continue;
      }
      weight += -(std::abs(zOrigin) * c.zOriginWeightFactor) + c.compatSeedWeight;
      // spB.index() etc. are positions in the packed container
      if (weight < getBestSeedQuality(ev.bestSeedQualityMap, spB) &&
          weight < getBestSeedQuality(ev.bestSeedQualityMap, m) &&
          weight < getBestSeedQuality(ev.bestSeedQualityMap, spT)) {
        continue;
      }
      if (deltaSeedConf > 0) {
        ev.collector.push(spB, m, spT, weight, zOrigin, true);
      } else if (weight > weightMax) {
        weightMax = weight;
        maxWeightTopSp = spT;
        maxWeightSeed = true;
      }
    } else {
      ev.collector.push(spB, m, spT, weight, zOrigin, false);
    }
  }

  if (c.seedConfirmation && maxWeightSeed && ev.collector.nHigh() == 0) {
    ev.collector.push(spB, m, static_cast<Index>(maxWeightTopSp), weightMax,
                      zOrigin, false);
  }
}

// BroadTripletSeedFilter.cpp:324-393
void filterTripletsMiddleFixed(Event& ev) {
  const b200seed_config& c = ev.s->cfg;
  const std::size_t numQualitySeeds = ev.collector.nHigh();
  ev.collector.toSorted(ev.sortedCandidates);
  const auto& sorted = ev.sortedCandidates;
  {
    bool tie = false;
    for (std::size_t k = 1; k < sorted.size(); ++k) tie |= sorted[k].weight == sorted[k - 1].weight;
    ev.cnt.nWeightTieMiddles += tie ? 1 : 0;
  }
  std::size_t maxSeeds = sorted.size();
  if (maxSeeds > c.maxSeedsPerSpM) {
    maxSeeds = c.maxSeedsPerSpM + 1;
  }
  std::size_t numTotalSeeds = 0;
  for (const auto& cand : sorted) {
    if (numTotalSeeds >= maxSeeds) break;
    if (c.seedConfirmation) {
      if (numQualitySeeds > 0 && !cand.isQuality) continue;
      if (cand.weight < getBestSeedQuality(ev.bestSeedQualityMap, cand.bottom) &&
          cand.weight < getBestSeedQuality(ev.bestSeedQualityMap, cand.middle) &&
          cand.weight < getBestSeedQuality(ev.bestSeedQualityMap, cand.top)) {
        continue;
      }
    }
    // write-only when seedConfirmation is off, kept so the timed CPU baseline
    // pays what the reference pays (.cpp:374-376)
    setBestSeedQuality(ev.bestSeedQualityMap, cand.bottom, cand.middle,
                       cand.top, cand.weight);
    ev.seeds.push_back({cand.bottom, cand.middle, cand.top, cand.weight, cand.zOrigin});
    ++numTotalSeeds;
  }
}

// TripletSeeder.cpp:44-107 (+ createAndFilterTriplets :21-42).  `makeTops` / `makeBottoms` fill the doublet
// containers from the candidate groups of the caller (grid bins or k-d-tree subsets).
template <typename MakeTops, typename MakeBottoms>
void seedsForMiddleT(Event& ev, Index m, MakeTops&& makeTops, MakeBottoms&& makeBottoms) {
  const MiddleInfo mi = computeMiddleInfo(ev.sp, m);
  ++ev.cnt.nMiddles;

  ev.topDoublets.clear();
  makeTops(mi, ev.topDoublets);
  if (ev.dump != nullptr) {
    // stage-level parity needs both lists even when the reference returns early
    ev.bottomDoublets.clear();
    makeBottoms(mi, ev.bottomDoublets, true);
    DoubletDump& d = *ev.dump;
    d.middlePos.push_back(m);
    d.nBottom.push_back(static_cast<Index>(ev.bottomDoublets.size()));
    for (const Doublets* src : {&ev.bottomDoublets, &ev.topDoublets}) {
      for (std::size_t i = 0; i < src->size(); ++i) {
        d.otherPos.push_back(src->sp[i]);
        d.cotTheta.push_back(src->cotTheta[i]);
        d.iDeltaR.push_back(src->iDeltaR[i]);
        d.er.push_back(src->er[i]);
        d.u.push_back(src->u[i]);
        d.v.push_back(src->v[i]);
        d.x.push_back(src->x[i]);
        d.y.push_back(src->y[i]);
      }
    }
    d.first.push_back(d.otherPos.size());
  }
  if (ev.topDoublets.empty()) return;
  if (!sufficientTopDoublets(ev, m)) return;

  ev.bottomDoublets.clear();
  makeBottoms(mi, ev.bottomDoublets, false);
  if (ev.bottomDoublets.empty()) return;

  ev.cnt.nBottomDoublets += ev.bottomDoublets.size();
  ev.cnt.nTopDoublets += ev.topDoublets.size();
  ev.cnt.maxBottoms = std::max<std::uint64_t>(ev.cnt.maxBottoms, ev.bottomDoublets.size());
  ev.cnt.maxTops = std::max<std::uint64_t>(ev.cnt.maxTops, ev.topDoublets.size());

  sortByCotTheta(ev, ev.bottomDoublets, ev.sortedBottoms);
  sortByCotTheta(ev, ev.topDoublets, ev.sortedTops);
  {
    bool tie = false;
    for (std::size_t k = 1; k < ev.sortedBottoms.size(); ++k)
      tie |= ev.sortedBottoms[k].cotTheta == ev.sortedBottoms[k - 1].cotTheta;
    for (std::size_t k = 1; k < ev.sortedTops.size(); ++k)
      tie |= ev.sortedTops[k].cotTheta == ev.sortedTops[k - 1].cotTheta;
    ev.cnt.nCotTieMiddles += tie ? 1 : 0;
  }

  ev.cnt.histBottoms[std::min<std::size_t>(31, ev.bottomDoublets.size() / 128)]++;
  ev.cnt.histTops[std::min<std::size_t>(31, ev.topDoublets.size() / 128)]++;
  std::size_t topBegin = 0;
  const std::size_t topEnd = ev.sortedTops.size();
  std::uint64_t candThisMiddle = 0;
  std::uint64_t candThisRound = 0, bottomInRound = 0;
  for (const IndexAndCotTheta& b : ev.sortedBottoms) {
    if (bottomInRound == 256) {
      ev.cnt.histCandRound[std::min<std::uint64_t>(31, candThisRound / 32)]++;
      bottomInRound = 0;
      candThisRound = 0;
    }
    ++bottomInRound;
    if (topBegin == topEnd) break;
    ev.topCandidates.clear();
    if (ev.strip != nullptr) {  // TripletSeedFinder.cpp:408-420: useStripInfo picks the strip implementation
      createStripTripletTopCandidates(ev, m, b.index, topBegin, topEnd);
    } else {
      createTripletTopCandidates(ev, m, b.index, topBegin, topEnd);
    }
    ev.cnt.nCandidates += ev.topCandidates.size();
    candThisMiddle += ev.topCandidates.size();
    candThisRound += ev.topCandidates.size();
    ev.cnt.maxCandidatesPerBottom =
        std::max<std::uint64_t>(ev.cnt.maxCandidatesPerBottom, ev.topCandidates.size());
    filterTripletTopCandidates(ev, m, b.index);
  }
  ev.cnt.maxCandidatesPerMiddle = std::max(ev.cnt.maxCandidatesPerMiddle, candThisMiddle);
  ev.cnt.histCandRound[std::min<std::uint64_t>(31, candThisRound / 32)]++;
  filterTripletsMiddleFixed(ev);
}

// the grid algorithm's groups: r-sorted bin ranges, advanced in place (DoubletSeedFinder.cpp:73-93)
void seedsForMiddle(Event& ev, DoubletCuts cuts, Index m,
                    std::vector<std::pair<Index, Index>>& bottomRanges,
                    std::vector<std::pair<Index, Index>>& topRanges) {
  seedsForMiddleT(
      ev, m,
      [&](const MiddleInfo& mi, Doublets& out) {
        for (auto& rng : topRanges) createDoublets<false>(ev, cuts, m, mi, rng.first, rng.second, out);
      },
      [&](const MiddleInfo& mi, Doublets& out, bool forDump) {
        if (forDump) {  // the dump must not advance the persistent ranges
          auto copy = bottomRanges;
          for (auto& rng : copy) createDoublets<true>(ev, cuts, m, mi, rng.first, rng.second, out);
        } else {
          for (auto& rng : bottomRanges) createDoublets<true>(ev, cuts, m, mi, rng.first, rng.second, out);
        }
      });
}

// GridTripletSeedingAlgorithm.cpp:404-421
std::pair<float, float> radiusRangeForMiddle(const Setup& s, float zM,
                                             std::pair<float, float> variable) {
  if (s.cfg.useVariableMiddleSPRange) return variable;
  if (s.rRangeMiddle.empty()) return {s.cfg.rMinMiddle, s.cfg.rMaxMiddle};
  auto pVal = std::ranges::lower_bound(s.zBinEdgesF, zM);
  std::size_t zBin = std::distance(s.zBinEdgesF.begin(), pVal);
  zBin == 0 ? zBin : --zBin;
  return {s.rRangeMiddle.at(zBin).first, s.rRangeMiddle.at(zBin).second};
}

// GridTripletSeedingAlgorithm.cpp:180-402
void runEvent(Event& ev) {
  const Setup& s = *ev.s;
  buildGrid(ev);
  const Packed& p = ev.sp;

  // :257-270
  float minRange = std::numeric_limits<float>::max();
  float maxRange = std::numeric_limits<float>::lowest();
  for (const auto& range : p.binRange) {
    if (range.first == range.second) continue;
    minRange = std::min(p.r[range.first], minRange);
    maxRange = std::max(p.r[range.second - 1], maxRange);
  }
  // :327-330
  const std::pair<float, float> rMiddleSpRange = {
      std::floor(minRange / 2) * 2 + s.cfg.deltaRMiddleMinSPRange,
      std::floor(maxRange / 2) * 2 - s.cfg.deltaRMiddleMaxSPRange};

  DoubletCuts cuts = DoubletCuts::None;
  // .cpp:292-297: a configured `inputVertices` key connects VertexZCuts for every event
  // (an empty window list then accepts every doublet, .cpp:77-79)
  if (s.cfg.useVertexZCuts || ev.zw.n > 0) {
    cuts = DoubletCuts::VertexZ;
  } else if (s.cfg.useExtraCuts) {
    cuts = DoubletCuts::Itk;
  }

  ev.collector.configure(s.cfg.maxSeedsPerSpMConf, s.cfg.maxQualitySeedsPerSpMConf);
  ev.collector.clear();
  ev.bestSeedQualityMap.clear();
  ev.rMaxSeedConf = 0;

  std::vector<std::pair<Index, Index>> bottomRanges, topRanges;
  std::uint32_t navIndex = 0;
  // BinnedGroupIterator (GridIterator.ipp:228-242): phi outermost, then z in
  // navigation order, then r; empty middle bins are skipped.
  for (std::size_t phiLoc : s.navPhi) {
    for (std::size_t zLoc : s.navZ) {
      for (std::size_t rLoc : s.navR) {
        if ((navIndex++ % ev.navStride) != ev.navPhase) continue;
        const std::size_t middleBin = (phiLoc * (s.nZ + 2) + zLoc) * (s.nR + 2) + rLoc;
        const auto middleRange = p.binRange[middleBin];
        if (middleRange.first == middleRange.second) continue;

        bottomRanges.clear();
        for (std::size_t b : findBins(s, phiLoc, zLoc, rLoc, false)) {
          bottomRanges.push_back(p.binRange.at(b));
        }
        topRanges.clear();
        for (std::size_t t : findBins(s, phiLoc, zLoc, rLoc, true)) {
          topRanges.push_back(p.binRange.at(t));
        }

        const std::pair<float, float> radiusRange =
            radiusRangeForMiddle(s, p.z[middleRange.first], rMiddleSpRange);

        // TripletSeeder.cpp:138-201
        const float firstMiddleSpR = p.r[middleRange.first];
        for (auto& rng : bottomRanges) {
          const float value = firstMiddleSpR - s.dRMaxB;
          const auto low = std::lower_bound(p.r.begin() + rng.first,
                                            p.r.begin() + rng.second, value);
          rng.first = static_cast<Index>(low - p.r.begin());
        }
        for (auto& rng : topRanges) {
          const float value = firstMiddleSpR + s.dRMinT;
          const auto low = std::lower_bound(p.r.begin() + rng.first,
                                            p.r.begin() + rng.second, value);
          rng.first = static_cast<Index>(low - p.r.begin());
        }
        for (Index m = middleRange.first; m < middleRange.second; ++m) {
          const float rM = p.r[m];
          if (rM < radiusRange.first) continue;
          if (rM > radiusRange.second) break;
          seedsForMiddle(ev, cuts, m, bottomRanges, topRanges);
        }
      }
    }
  }
  ev.cnt.nSeeds = ev.seeds.size();
  // :394-398 remap to the caller's indices
  for (Seed& sd : ev.seeds) {
    sd.b = p.copiedFrom[sd.b];
    sd.m = p.copiedFrom[sd.m];
    sd.t = p.copiedFrom[sd.t];
  }
}

// ===========================================================================
// OrthogonalTripletSeedingAlgorithm (Examples/Algorithms/TrackFinding/src/
// OrthogonalTripletSeedingAlgorithm.cpp:62-317): the same doublet / triplet / filter
// stages behind a k-d-tree candidate provider.
// ===========================================================================
namespace orth {

constexpr std::size_t kDims = 3;      // CylindricalSpacePointKDTree.hpp:33-36 (phi, r, z)
constexpr std::size_t kLeafSize = 4;  // CylindricalSpacePointKDTree.hpp:42
enum Dim { DimPhi = 0, DimR = 1, DimZ = 2 };

using Coord = std::array<float, kDims>;
using Pair = std::pair<Coord, Index>;  // KDTree::pair_t

// RangeXD<3, float> (Core/include/Acts/Utilities/RangeXD.hpp): semi-open [min, max) per dimension
struct Range {
  float mn[kDims], mx[kDims];
  Range() {  // RangeXD():  lowest .. max
    for (std::size_t i = 0; i < kDims; ++i) {
      mn[i] = std::numeric_limits<float>::lowest();
      mx[i] = std::numeric_limits<float>::max();
    }
  }
  void shrinkMin(std::size_t i, float v) { mn[i] = std::max(mn[i], v); }
  void shrinkMax(std::size_t i, float v) { mx[i] = std::min(mx[i], v); }
  void shrink(std::size_t i, float lo, float hi) { shrinkMin(i, lo); shrinkMax(i, hi); }
  bool degenerate() const {
    for (std::size_t i = 0; i < kDims; ++i) if (mn[i] >= mx[i]) return true;
    return false;
  }
  bool contains(const Coord& v) const {
    for (std::size_t i = 0; i < kDims; ++i) if (!(mn[i] <= v[i] && v[i] < mx[i])) return false;
    return true;
  }
  bool covers(const Range& o) const {  // operator>=
    for (std::size_t i = 0; i < kDims; ++i) if (!(mn[i] <= o.mn[i] && mx[i] >= o.mx[i])) return false;
    return true;
  }
  bool intersects(const Range& r) const {  // operator&&
    for (std::size_t i = 0; i < kDims; ++i) if (!(mn[i] < r.mx[i] && r.mn[i] < mx[i])) return false;
    return true;
  }
};

// KDTree<3, SpacePointIndex, float, std::array, 4> (Core/include/Acts/Utilities/KDTree.hpp:40-503): nodes hold
// [begin, end) of the one element vector, which the constructors permute in place.
struct Node {
  std::size_t begin = 0, end = 0;
  bool internal = false;
  Range range;
  int lhs = -1, rhs = -1;
};
struct Tree {
  std::vector<Pair> elems;
  std::vector<Node> nodes;

  static Range boundingBox(const std::vector<Pair>& e, std::size_t b, std::size_t en) {  // :236-264
    Coord mn{}, mx{};
    for (std::size_t i = 0; i < kDims; ++i) {
      mn[i] = std::numeric_limits<float>::max();
      mx[i] = std::numeric_limits<float>::lowest();
    }
    for (std::size_t i = b; i != en; ++i) {
      for (std::size_t j = 0; j < kDims; ++j) {
        mn[j] = std::min(mn[j], e[i].first[j]);
        mx[j] = std::max(mx[j], e[i].first[j]);
      }
    }
    Range r;
    for (std::size_t j = 0; j < kDims; ++j) {
      r.mn[j] = mn[j];
      r.mx[j] = std::nextafter(mx[j], std::numeric_limits<float>::max());  // nextRepresentable, :222-234
    }
    return r;
  }

  // KDTreeNode constructor, :271-350
  int build(std::size_t b, std::size_t e, bool internal, std::size_t d) {
    const int id = static_cast<int>(nodes.size());
    nodes.emplace_back();
    nodes[id].begin = b;
    nodes[id].end = e;
    nodes[id].internal = internal;
    nodes[id].range = boundingBox(elems, b, e);
    if (!internal) return id;
    constexpr std::size_t maxExactMedian = 128;
    const auto first = elems.begin() + static_cast<std::ptrdiff_t>(b);
    const auto last = elems.begin() + static_cast<std::ptrdiff_t>(e);
    auto pivot = first;
    if (e - b > maxExactMedian) {
      const float mid = static_cast<float>(0.5) * (nodes[id].range.mx[d] + nodes[id].range.mn[d]);
      pivot = std::partition(first, last, [=](const Pair& i) { return i.first[d] < mid; });
    } else {
      std::sort(first, last, [d](const Pair& a, const Pair& bb) { return a.first[d] < bb.first[d]; });
      pivot = first + (std::distance(first, last) / 2);
    }
    if (pivot == first || pivot == std::prev(last)) {
      pivot = std::next(first, kLeafSize);
    }
    const std::size_t p = static_cast<std::size_t>(pivot - elems.begin());
    const std::size_t lhsSize = p - b, rhsSize = e - p;
    const int l = build(b, p, lhsSize > kLeafSize, (d + 1) % kDims);
    nodes[id].lhs = l;
    const int r = build(p, e, rhsSize > kLeafSize, (d + 1) % kDims);
    nodes[id].rhs = r;
    return id;
  }
  void construct() {  // KDTree(vector_t&&), :66-80
    nodes.clear();
    nodes.reserve(elems.size());
    build(0, elems.size(), elems.size() > kLeafSize, 0);
  }

  // KDTreeNode::rangeSearchMapDiscard, :352-396
  void search(int id, const Range& r, std::vector<Index>& out) const {
    const Node& n = nodes[static_cast<std::size_t>(id)];
    const bool contained = r.covers(n.range);
    if (n.internal) {
      if (contained) {
        for (std::size_t i = n.begin; i != n.end; ++i) out.push_back(elems[i].second);
        return;
      }
      if (nodes[static_cast<std::size_t>(n.lhs)].range.intersects(r)) search(n.lhs, r, out);
      if (nodes[static_cast<std::size_t>(n.rhs)].range.intersects(r)) search(n.rhs, r, out);
    } else {
      for (std::size_t i = n.begin; i != n.end; ++i) {
        if (contained || r.contains(elems[i].first)) out.push_back(elems[i].second);
      }
    }
  }
};

// CylindricalSpacePointKDTree::Options (hpp:45-70)
struct Options {
  float rMax, zMin, zMax, phiMin, phiMax, deltaRMin, deltaRMax, collisionRegionMin, collisionRegionMax, cotThetaMax,
      deltaPhiMax;
  float deltaZMax = std::numeric_limits<float>::infinity();  // hpp default; the algorithm never sets it
};

// CylindricalSpacePointKDTree.cpp:17-96
Range validTupleOrthoRangeLH(const Options& o, float pL, float rL, float zL) {
  const float colMin = o.collisionRegionMin;
  const float colMax = o.collisionRegionMax;
  Range res;
  res.shrinkMin(DimPhi, o.phiMin);
  res.shrinkMax(DimPhi, o.phiMax);
  res.shrinkMax(DimR, o.rMax);
  res.shrinkMin(DimZ, o.zMin);
  res.shrinkMax(DimZ, o.zMax);
  res.shrinkMin(DimR, rL + o.deltaRMin);
  res.shrinkMax(DimR, rL + o.deltaRMax);
  const float zMax = (res.mx[DimR] / rL) * (zL - colMin) + colMin;
  const float zMin = colMax - (res.mx[DimR] / rL) * (colMax - zL);
  if (zL > colMin) {
    res.shrinkMax(DimZ, zMax);
  } else if (zL < colMax) {
    res.shrinkMin(DimZ, zMin);
  }
  res.shrinkMin(DimZ, zL - o.cotThetaMax * (res.mx[DimR] - rL));
  res.shrinkMax(DimZ, zL + o.cotThetaMax * (res.mx[DimR] - rL));
  res.shrinkMin(DimPhi, pL - o.deltaPhiMax);
  res.shrinkMax(DimPhi, pL + o.deltaPhiMax);
  res.shrinkMin(DimZ, zL - o.deltaZMax);
  res.shrinkMax(DimZ, zL + o.deltaZMax);
  return res;
}

// CylindricalSpacePointKDTree.cpp:98-157
Range validTupleOrthoRangeHL(const Options& o, float pM, float rM, float zM) {
  Range res;
  res.shrinkMin(DimPhi, o.phiMin);
  res.shrinkMax(DimPhi, o.phiMax);
  res.shrinkMax(DimR, o.rMax);
  res.shrinkMin(DimZ, o.zMin);
  res.shrinkMax(DimZ, o.zMax);
  res.shrinkMin(DimR, rM - o.deltaRMax);
  res.shrinkMax(DimR, rM - o.deltaRMin);
  const float fracR = res.mn[DimR] / rM;
  const float zMin = (zM - o.collisionRegionMin) * fracR + o.collisionRegionMin;
  const float zMax = (zM - o.collisionRegionMax) * fracR + o.collisionRegionMax;
  res.shrinkMin(DimZ, std::min(zMin, zM));
  res.shrinkMax(DimZ, std::max(zMax, zM));
  res.shrinkMin(DimPhi, pM - o.deltaPhiMax);
  res.shrinkMax(DimPhi, pM + o.deltaPhiMax);
  res.shrinkMin(DimZ, zM - o.deltaZMax);
  res.shrinkMax(DimZ, zM + o.deltaZMax);
  return res;
}

struct Candidates {  // hpp:73-93
  std::vector<Index> bottom_lh_v, bottom_hl_v, top_lh_v, top_hl_v;
  void clear() { bottom_lh_v.clear(); bottom_hl_v.clear(); top_lh_v.clear(); top_hl_v.clear(); }
};

// the four search boxes of one middle (CylindricalSpacePointKDTree.cpp:159-217), shared with the tests
struct Boxes {
  Range bottom_lh, top_lh, bottom_hl, top_hl;
};
Boxes searchBoxes(const Options& lh, const Options& hl, float pM, float rM, float zM) {
  const Range bottom_r = validTupleOrthoRangeHL(hl, pM, rM, zM);
  const Range top_r = validTupleOrthoRangeLH(lh, pM, rM, zM);
  const float cotTheta = std::max(std::abs(zM / rM), lh.cotThetaMax);
  const float deltaRMaxTop = top_r.mx[DimR] - rM;
  const float deltaRMaxBottom = rM - bottom_r.mn[DimR];
  Boxes b;
  b.bottom_lh = bottom_r;
  b.bottom_lh.shrink(DimZ, zM - cotTheta * deltaRMaxBottom, zM);
  b.top_lh = top_r;
  b.top_lh.shrink(DimZ, zM, zM + cotTheta * deltaRMaxTop);
  b.bottom_hl = bottom_r;
  b.bottom_hl.shrink(DimZ, zM, zM + cotTheta * deltaRMaxBottom);
  b.top_hl = top_r;
  b.top_hl.shrink(DimZ, zM - cotTheta * deltaRMaxTop, zM);
  return b;
}

// CylindricalSpacePointKDTree::validTuples, .cpp:159-263
void validTuples(const Tree& tree, const Options& lh, const Options& hl, float pM, float rM, float zM,
                 std::size_t nTopSeedConf, Candidates& c) {
  const Boxes b = searchBoxes(lh, hl, pM, rM, zM);
  if (!b.bottom_lh.degenerate() && !b.top_lh.degenerate()) tree.search(0, b.top_lh, c.top_lh_v);
  if (!b.bottom_hl.degenerate() && !b.top_hl.degenerate()) tree.search(0, b.top_hl, c.top_hl_v);
  const bool searchBotLh = c.top_lh_v.size() >= nTopSeedConf;
  const bool searchBotHl = c.top_hl_v.size() >= nTopSeedConf;
  if (!c.top_lh_v.empty() && searchBotLh) tree.search(0, b.bottom_lh, c.bottom_lh_v);
  if (!c.top_hl_v.empty() && searchBotHl) tree.search(0, b.bottom_hl, c.bottom_hl_v);
}

struct OrthOptions {
  float zOutermostLayersMin, zOutermostLayersMax, deltaPhiMax;
};

// what the tests read back about the tree
struct TreeDump {
  std::vector<Index> order;  // element order after construction (= middle iteration order), core indices
};

// OrthogonalTripletSeedingAlgorithm::execute, .cpp:101-317
void runEvent(Event& ev, const OrthOptions& oo, TreeDump* dump) {
  const Setup& s = *ev.s;
  const b200seed_config& c = s.cfg;
  Packed& p = ev.sp;
  Tree tree;
  std::vector<float> phi;
  double extentRMin = std::numeric_limits<double>::max(), extentRMax = std::numeric_limits<double>::lowest();
  bool extentSet = false;
  for (Index i = 0; i < ev.n; ++i) {
    if (c.useExtraCuts && !itkFastTrackingSPselect(ev.r[i], ev.z[i])) continue;  // :123-127
    const Index newIndex = static_cast<Index>(p.copiedFrom.size());
    p.copiedFrom.push_back(i);
    p.x.push_back(ev.x[i]); p.y.push_back(ev.y[i]); p.z.push_back(ev.z[i]); p.r.push_back(ev.r[i]);
    p.varZ.push_back(ev.varZ[i]); p.varR.push_back(ev.varR[i]);
    const float ph = ev.phiOverride != nullptr ? ev.phiOverride[i] : std::atan2(ev.y[i], ev.x[i]);  // :136
    phi.push_back(ph);
    tree.elems.push_back({Coord{ph, ev.r[i], ev.z[i]}, newIndex});  // :140, insert(index, phi, r, z)
    // Extent::extend (Core/src/Geometry/Extent.cpp:33-56): AxisR = VectorHelpers::perp of the double vector
    const double xd = ev.x[i], yd = ev.y[i];
    const double perp = std::sqrt(xd * xd + yd * yd);
    if (!extentSet) { extentRMin = perp; extentRMax = perp; extentSet = true; }
    extentRMin = std::min(extentRMin, perp);
    extentRMax = std::max(extentRMax, perp);
  }
  ev.cnt.nInGrid = p.copiedFrom.size();
  if (tree.elems.empty()) return;  // (KDTree of an empty vector: a single empty leaf, no middle)
  tree.construct();
  if (dump != nullptr) for (const Pair& e : tree.elems) dump->order.push_back(e.second);

  // :152-171
  Options lh{};
  lh.rMax = c.rMax; lh.zMin = c.zMin; lh.zMax = c.zMax; lh.phiMin = c.phiMin; lh.phiMax = c.phiMax;
  lh.deltaRMin = std::isnan(c.deltaRMinBottom) ? c.deltaRMin : c.deltaRMinBottom;
  lh.deltaRMax = std::isnan(c.deltaRMaxBottom) ? c.deltaRMax : c.deltaRMaxBottom;
  lh.collisionRegionMin = c.collisionRegionMin; lh.collisionRegionMax = c.collisionRegionMax;
  lh.cotThetaMax = c.cotThetaMax; lh.deltaPhiMax = oo.deltaPhiMax;
  Options hl = lh;
  hl.deltaRMin = std::isnan(c.deltaRMinTop) ? c.deltaRMin : c.deltaRMinTop;
  hl.deltaRMax = std::isnan(c.deltaRMaxTop) ? c.deltaRMax : c.deltaRMaxTop;

  // :227-232 (double arithmetic, stored in a Range1D<float>)
  const float rMiddleMin = static_cast<float>(std::floor(extentRMin / 2) * 2 + c.deltaRMiddleMinSPRange);
  const float rMiddleMax = static_cast<float>(std::floor(extentRMax / 2) * 2 - c.deltaRMiddleMaxSPRange);

  const DoubletCuts cuts = c.useExtraCuts ? DoubletCuts::Itk : DoubletCuts::None;  // :191-193
  ev.collector.configure(c.maxSeedsPerSpMConf, c.maxQualitySeedsPerSpMConf);
  ev.collector.clear();
  ev.bestSeedQualityMap.clear();
  ev.rMaxSeedConf = 0;

  Candidates cand;
  for (const Pair& middle : tree.elems) {  // :247
    const Index m = middle.second;
    const float rM = p.r[m];
    if (c.useVariableMiddleSPRange) {
      if (rM < rMiddleMin || rM > rMiddleMax) continue;
    } else {
      if (rM > c.rMaxMiddle || rM < c.rMinMiddle) continue;
    }
    const float zM = p.z[m];
    if (zM < oo.zOutermostLayersMin || zM > oo.zOutermostLayersMax) continue;
    const float phiM = phi[m];
    if (phiM > c.phiMax || phiM < c.phiMin) continue;

    std::size_t nTopSeedConf = 0;
    if (c.seedConfirmation) {  // :276-288
      const b200seed_seed_confirmation_range& rng =
          (zM > c.centralSeedConfirmationRange.zMaxSeedConf || zM < c.centralSeedConfirmationRange.zMinSeedConf)
              ? c.forwardSeedConfirmationRange
              : c.centralSeedConfirmationRange;
      nTopSeedConf = rM > rng.rMaxSeedConf ? rng.nTopForLargeR : rng.nTopForSmallR;
    }
    cand.clear();
    validTuples(tree, lh, hl, phiM, rM, zM, nTopSeedConf, cand);
    for (int group = 0; group < 2; ++group) {  // :294-306: increasing-z candidates, then decreasing-z
      const std::vector<Index>& bottoms = group == 0 ? cand.bottom_lh_v : cand.bottom_hl_v;
      const std::vector<Index>& tops = group == 0 ? cand.top_lh_v : cand.top_hl_v;
      seedsForMiddleT(
          ev, m,
          [&](const MiddleInfo& mi, Doublets& out) { createDoubletsUnsorted<false>(ev, cuts, m, mi, tops, out); },
          [&](const MiddleInfo& mi, Doublets& out, bool) { createDoubletsUnsorted<true>(ev, cuts, m, mi, bottoms, out); });
    }
  }
  ev.cnt.nSeeds = ev.seeds.size();
  for (Seed& sd : ev.seeds) {  // :310-314
    sd.b = p.copiedFrom[sd.b];
    sd.m = p.copiedFrom[sd.m];
    sd.t = p.copiedFrom[sd.t];
  }
}

// the derived constants of the three finders with the orthogonal algorithm's own deltaR resolution
// (.cpp:173-205: note the isnan tests on deltaRMax{Bottom,Top} for BOTH bounds)
Setup makeSetup(const b200seed_config& c) {
  Setup s;
  s.cfg = c;
  s.cfg.zBinNeighborsTop = s.cfg.zBinNeighborsBottom = nullptr;
  s.cfg.zBinEdges = nullptr;
  s.cfg.zBinsCustomLooping = nullptr;
  s.cfg.rRangeMiddleSP = nullptr;
  s.dRMinB = std::isnan(c.deltaRMaxBottom) ? c.deltaRMin : c.deltaRMinBottom;
  s.dRMaxB = std::isnan(c.deltaRMaxBottom) ? c.deltaRMax : c.deltaRMaxBottom;
  s.dRMinT = std::isnan(c.deltaRMaxTop) ? c.deltaRMin : c.deltaRMinTop;
  s.dRMaxT = std::isnan(c.deltaRMaxTop) ? c.deltaRMax : c.deltaRMaxTop;
  deriveFinderConstants(s);
  return s;
}

}  // namespace orth

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

}  // namespace

// ---------------------------------------------------------------------------
// C interface (ctypes / bench.py)
// ---------------------------------------------------------------------------
extern "C" {

struct oracle_counters {
  std::uint64_t nSpacePoints, nInGrid, nMiddles, nPairTests, nBottomDoublets,
      nTopDoublets, nTripletTests, nCandidates, nSeeds, nRTieBins,
      nCotTieMiddles, nCurvTieGroups, nWeightTieMiddles, maxBottoms, maxTops,
      maxCandidatesPerBottom, maxCandidatesPerMiddle, maxBinSize;
};

struct oracle_handle {
  Setup setup;
  bool orthogonal = false;
  orth::OrthOptions orthOptions{};
};

const char* oracle_last_error() { return g_error.c_str(); }

// Reference defaults, GridTripletSeedingAlgorithm.hpp:34-244 (restated
// independently of the product's b200seed_config_init; a test compares them).
void oracle_config_init(b200seed_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->abi_version = B200SEED_ABI_VERSION;
  c->struct_size = sizeof(b200seed_config);
  const float nan = std::numeric_limits<float>::quiet_NaN();
  const float inf = std::numeric_limits<float>::infinity();
  c->bFieldInZ = 2 * 0.000299792458;
  c->minPt = 0.4;
  c->cotThetaMax = 10.01788;
  c->impactMax = 20;
  c->deltaRMin = 5;
  c->deltaRMax = 270;
  c->deltaRMinTop = c->deltaRMaxTop = c->deltaRMinBottom = c->deltaRMaxBottom = nan;
  c->rMin = 0;
  c->rMax = 600;
  c->zMin = -2800;
  c->zMax = 2800;
  c->phiMin = -std::numbers::pi_v<float>;
  c->phiMax = std::numbers::pi_v<float>;
  c->phiBinDeflectionCoverage = 1;
  c->maxPhiBins = 10000;
  c->numPhiNeighbors = 1;
  c->rMinMiddle = 60;
  c->rMaxMiddle = 120;
  c->useVariableMiddleSPRange = 0;
  c->deltaRMiddleMinSPRange = 10;
  c->deltaRMiddleMaxSPRange = 10;
  c->deltaZMin = -inf;
  c->deltaZMax = inf;
  c->interactionPointCut = 0;
  c->collisionRegionMin = -150;
  c->collisionRegionMax = +150;
  c->helixCutTolerance = 1;
  c->sigmaScattering = 5;
  c->radLengthPerSeed = 0.05;
  c->toleranceParam = 1.1;
  c->deltaInvHelixDiameter = 0.00003;
  c->compatSeedWeight = 200;
  c->impactWeightFactor = 1;
  c->zOriginWeightFactor = 1;
  c->maxSeedsPerSpM = 5;
  c->compatSeedLimit = 2;
  c->seedWeightIncrement = 0;
  c->numSeedIncrement = inf;
  c->seedConfirmation = 0;
  for (b200seed_seed_confirmation_range* r :
       {&c->centralSeedConfirmationRange, &c->forwardSeedConfirmationRange}) {
    r->zMinSeedConf = std::numeric_limits<float>::lowest();
    r->zMaxSeedConf = std::numeric_limits<float>::max();
    r->rMaxSeedConf = std::numeric_limits<float>::max();
    r->nTopForLargeR = 0;
    r->nTopForSmallR = 0;
    r->seedConfMinBottomRadius = 60.;
    r->seedConfMaxZOrigin = 150.;
    r->minImpactSeedConf = 1.;
  }
  c->maxSeedsPerSpMConf = 5;
  c->maxQualitySeedsPerSpMConf = 5;
  c->useDeltaRinsteadOfTopRadius = 0;
  c->useExtraCuts = 0;
  c->useVertexZCuts = 0;
  c->vertexZNSigma = 3.0;
  c->vertexZMargin = 0.0;
  c->relaxedFloat = 0;
}

int oracle_create(const b200seed_config* cfg, oracle_handle** out) {
  try {
    auto* h = new oracle_handle{makeSetup(*cfg)};
    *out = h;
    return B200SEED_OK;
  } catch (const std::invalid_argument& e) {
    return fail(B200SEED_ERR_INVALID_ARGUMENT, e.what());
  } catch (const std::domain_error& e) {
    return fail(B200SEED_ERR_DOMAIN, e.what());
  } catch (const std::runtime_error& e) {
    return fail(B200SEED_ERR_RUNTIME, e.what());
  } catch (const std::exception& e) {
    return fail(B200SEED_ERR_RUNTIME, e.what());
  }
}
void oracle_destroy(oracle_handle* h) { delete h; }

// OrthogonalTripletSeedingAlgorithm::Config defaults (hpp:38-186): the shared fields have the grid
// algorithm's defaults, plus zOutermostLayers and deltaPhiMax.
void oracle_orthogonal_config_init(b200seed_config* c, b200seed_orthogonal_options* o) {
  oracle_config_init(c);
  o->zOutermostLayersMin = -2700;
  o->zOutermostLayersMax = 2700;
  o->deltaPhiMax = 0.085;
}
int oracle_create_orthogonal(const b200seed_config* cfg, const b200seed_orthogonal_options* opt, oracle_handle** out) {
  try {
    auto* h = new oracle_handle{orth::makeSetup(*cfg)};
    h->orthogonal = true;
    h->orthOptions = {opt->zOutermostLayersMin, opt->zOutermostLayersMax, opt->deltaPhiMax};
    *out = h;
    return B200SEED_OK;
  } catch (const std::exception& e) {
    return fail(B200SEED_ERR_RUNTIME, e.what());
  }
}

int oracle_get_info(const oracle_handle* h, b200seed_info* info) {
  const Setup& s = h->setup;
  std::memset(info, 0, sizeof(*info));
  info->phiBins = s.phiBins;
  info->zBins = static_cast<int>(s.nZ);
  info->rBins = static_cast<int>(s.nR);
  info->nGlobalBins = static_cast<int>(s.nGlobal);
  info->minHelixDiameter2 = s.minHelixDiameter2;
  info->highland = s.highland;
  info->sigmapT2perRadius = s.sigmapT2perRadius;
  info->multipleScattering2 = s.multipleScattering2;
  info->deltaRMinBottom = s.dRMinB;
  info->deltaRMaxBottom = s.dRMaxB;
  info->deltaRMinTop = s.dRMinT;
  info->deltaRMaxTop = s.dRMaxT;
  return B200SEED_OK;
}

// z edges / neighbour lists for the tests of the utility layer
int oracle_z_edges(const oracle_handle* h, double* edges, std::uint32_t cap) {
  const auto& e = h->setup.zEdges;
  for (std::uint32_t i = 0; i < e.size() && i < cap; ++i) edges[i] = e[i];
  return static_cast<int>(e.size());
}
int oracle_find_bins(const oracle_handle* h, std::uint32_t phiLoc, std::uint32_t zLoc,
                     std::uint32_t rLoc, int top, std::uint64_t* out, std::uint32_t cap) {
  const auto v = findBins(h->setup, phiLoc, zLoc, rLoc, top != 0);
  for (std::uint32_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
  return static_cast<int>(v.size());
}
int oracle_neighbors_closed(std::uint32_t idx, int first, int second, int nBins,
                            std::uint64_t* out, std::uint32_t cap) {
  const auto v = neighborsClosed(idx, {first, second}, nBins);
  for (std::uint32_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
  return static_cast<int>(v.size());
}
int oracle_neighbors_open(std::uint32_t idx, int first, int second, int nBins,
                          std::uint64_t* out, std::uint32_t cap) {
  const auto v = neighborsOpen(idx, {first, second}, nBins);
  for (std::uint32_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
  return static_cast<int>(v.size());
}
// global bin of a point, or -1 when outside (phi given explicitly)
std::int64_t oracle_bin_index(const oracle_handle* h, float phi, float z, float r) {
  const std::size_t b = binIndex(h->setup, phi, z, r);
  return b < h->setup.nGlobal ? static_cast<std::int64_t>(b) : -1;
}
float oracle_atan2f(float y, float x) { return std::atan2(y, x); }

struct oracle_event_result {
  Event ev;
  DoubletDump dump;
  orth::TreeDump treeDump;
};

// GridTripletSeedingAlgorithm.cpp:187-206: one z window per vertex,
// [z - half, z + half] with half = vertexZNSigma * sqrt(cov(2, 2)) + vertexZMargin in double,
// narrowed to float at the end.
void oracle_vertex_windows(const b200seed_config* cfg, std::uint32_t nVertices, const double* vertexZ,
                           const double* vertexVarZ, float* lo, float* hi) {
  for (std::uint32_t i = 0; i < nVertices; ++i) {
    const double z = vertexZ[i];
    const double sigmaZ = std::sqrt(vertexVarZ[i]);
    const double half = cfg->vertexZNSigma * sigmaZ + cfg->vertexZMargin;
    lo[i] = static_cast<float>(z - half);
    hi[i] = static_cast<float>(z + half);
  }
}

// Run one event.  Returns an opaque result the caller reads with the
// accessors below and frees with oracle_result_free.
int oracle_run(const oracle_handle* h, std::uint32_t n, const float* x, const float* y,
               const float* z, const float* r, const float* varZ, const float* varR,
               std::uint32_t nZWindows, const float* zLo, const float* zHi,
               int sortMode, int dumpDoublets, const float* phiOverride,
               oracle_event_result** out) {
  try {
    auto* res = new oracle_event_result;
    Event& ev = res->ev;
    ev.s = &h->setup;
    ev.x = x; ev.y = y; ev.z = z; ev.r = r; ev.varZ = varZ; ev.varR = varR;
    ev.n = n;
    ev.zw = {zLo, zHi, nZWindows};
    ev.sortMode = sortMode;
    ev.phiOverride = phiOverride;
    if (dumpDoublets != 0) {
      res->dump.first.push_back(0);
      ev.dump = &res->dump;
    }
    if (h->orthogonal) {
      orth::runEvent(ev, h->orthOptions, &res->treeDump);
    } else {
      runEvent(ev);
    }
    *out = res;
    return B200SEED_OK;
  } catch (const std::exception& e) {
    return fail(B200SEED_ERR_RUNTIME, e.what());
  }
}
// One event through the strip triplet path (TripletSeedFinder::Config::useStripInfo = true, cotThetaDiffMax):
// strip = 12 floats per space point (outerCenter, innerToOuterSeparation, outerHalfVector, innerHalfVector).
int oracle_run_strips(const oracle_handle* h, std::uint32_t n, const float* x, const float* y,
                      const float* z, const float* r, const float* varZ, const float* varR,
                      const float* strip, float cotThetaDiffMax, int sortMode, oracle_event_result** out) {
  try {
    if (h->orthogonal) return fail(B200SEED_ERR_UNSUPPORTED, "strip triplet path: grid handles only");
    if (strip == nullptr && n != 0) return fail(B200SEED_ERR_INVALID_ARGUMENT, "strip details missing");
    auto* res = new oracle_event_result;
    Event& ev = res->ev;
    ev.s = &h->setup;
    ev.x = x; ev.y = y; ev.z = z; ev.r = r; ev.varZ = varZ; ev.varR = varR;
    ev.n = n;
    ev.zw = {nullptr, nullptr, 0};
    ev.sortMode = sortMode;
    static const float kNoStrip[12] = {};
    ev.strip = strip != nullptr ? strip : kNoStrip;
    ev.cotThetaDiffMax = cotThetaDiffMax;
    runEvent(ev);
    *out = res;
    return B200SEED_OK;
  } catch (const std::exception& e) {
    return fail(B200SEED_ERR_RUNTIME, e.what());
  }
}
void oracle_result_free(oracle_event_result* r) { delete r; }
// element order of the k-d tree after construction (core space point indices), orthogonal handles only
std::uint64_t oracle_result_tree_order(const oracle_event_result* r, std::uint32_t* order) {
  if (order != nullptr) std::copy(r->treeDump.order.begin(), r->treeDump.order.end(), order);
  return r->treeDump.order.size();
}

std::uint64_t oracle_result_num_seeds(const oracle_event_result* r) { return r->ev.seeds.size(); }
void oracle_result_seeds(const oracle_event_result* r, std::uint32_t* b, std::uint32_t* m,
                         std::uint32_t* t, float* quality, float* vertexZ) {
  const auto& s = r->ev.seeds;
  for (std::size_t i = 0; i < s.size(); ++i) {
    b[i] = s[i].b; m[i] = s[i].m; t[i] = s[i].t;
    quality[i] = s[i].quality; vertexZ[i] = s[i].vertexZ;
  }
}
void oracle_result_counters(const oracle_event_result* r, oracle_counters* c) {
  const Counters& k = r->ev.cnt;
  *c = {r->ev.n, k.nInGrid, k.nMiddles, k.nPairTests, k.nBottomDoublets,
        k.nTopDoublets, k.nTripletTests, k.nCandidates, k.nSeeds, k.nRTieBins,
        k.nCotTieMiddles, k.nCurvTieGroups, k.nWeightTieMiddles, k.maxBottoms,
        k.maxTops, k.maxCandidatesPerBottom, k.maxCandidatesPerMiddle, k.maxBinSize};
}
// histograms: bottoms / tops per middle (bucket 128), candidates per round of
// 256 cotTheta-sorted bottoms (bucket 32); 32 buckets each
void oracle_result_histograms(const oracle_event_result* r, std::uint64_t* out) {
  for (int i = 0; i < 32; ++i) {
    out[i] = r->ev.cnt.histBottoms[i];
    out[32 + i] = r->ev.cnt.histTops[i];
    out[64 + i] = r->ev.cnt.histCandRound[i];
  }
}
std::uint64_t oracle_result_grid_size(const oracle_event_result* r) { return r->ev.sp.copiedFrom.size(); }
void oracle_result_grid(const oracle_event_result* r, std::uint32_t* copiedFrom, float* x,
                        float* y, float* z, float* rr, float* varZ, float* varR,
                        std::uint32_t* binBegin, std::uint32_t* binEnd) {
  const Packed& p = r->ev.sp;
  const std::size_t n = p.copiedFrom.size();
  std::memcpy(copiedFrom, p.copiedFrom.data(), n * 4);
  std::memcpy(x, p.x.data(), n * 4);
  std::memcpy(y, p.y.data(), n * 4);
  std::memcpy(z, p.z.data(), n * 4);
  std::memcpy(rr, p.r.data(), n * 4);
  std::memcpy(varZ, p.varZ.data(), n * 4);
  std::memcpy(varR, p.varR.data(), n * 4);
  for (std::size_t b = 0; b < p.binRange.size(); ++b) {
    binBegin[b] = p.binRange[b].first;
    binEnd[b] = p.binRange[b].second;
  }
}
std::uint64_t oracle_result_dump_middles(const oracle_event_result* r) { return r->dump.middlePos.size(); }
std::uint64_t oracle_result_dump_doublets(const oracle_event_result* r) { return r->dump.otherPos.size(); }
void oracle_result_dump(const oracle_event_result* r, std::uint32_t* middlePos,
                        std::uint64_t* first, std::uint32_t* nBottom, std::uint32_t* otherPos,
                        float* cotTheta, float* iDeltaR, float* er, float* u, float* v,
                        float* xNew, float* yNew) {
  const DoubletDump& d = r->dump;
  std::memcpy(middlePos, d.middlePos.data(), d.middlePos.size() * 4);
  std::memcpy(first, d.first.data(), d.first.size() * 8);
  std::memcpy(nBottom, d.nBottom.data(), d.nBottom.size() * 4);
  const std::size_t n = d.otherPos.size();
  std::memcpy(otherPos, d.otherPos.data(), n * 4);
  std::memcpy(cotTheta, d.cotTheta.data(), n * 4);
  std::memcpy(iDeltaR, d.iDeltaR.data(), n * 4);
  std::memcpy(er, d.er.data(), n * 4);
  std::memcpy(u, d.u.data(), n * 4);
  std::memcpy(v, d.v.data(), n * 4);
  std::memcpy(xNew, d.x.data(), n * 4);
  std::memcpy(yNew, d.y.data(), n * 4);
}

// ---------------------------------------------------------------------------
// estimateTrackParamsFromSeed, Core/src/Seeding/EstimateTrackParamsFromSeed.cpp:20-160,
// restated with explicit 3-vectors.  Like Eigen's Transform<double,3,Affine>::inverse()
// the inverse of the frame is computed as a general 3x3 inverse (cofactors /
// determinant), not as a transpose.
// ---------------------------------------------------------------------------
void oracle_estimate_params(std::uint64_t nSeeds, const std::uint32_t* bottom, const std::uint32_t* middle,
                            const std::uint32_t* top, const float* x, const float* y, const float* z,
                            const double* bField, double* out) {
  using V3 = std::array<double, 3>;
  auto sub = [](const V3& a, const V3& b) { return V3{a[0] - b[0], a[1] - b[1], a[2] - b[2]}; };
  auto cross = [](const V3& a, const V3& b) {
    return V3{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  };
  auto norm = [](const V3& a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
  auto normalized = [&](const V3& a) { const double n = norm(a); return V3{a[0] / n, a[1] / n, a[2] / n}; };
  for (std::uint64_t i = 0; i < nSeeds; ++i) {
    const V3 sp0{x[bottom[i]], y[bottom[i]], z[bottom[i]]};
    const V3 sp1{x[middle[i]], y[middle[i]], z[middle[i]]};
    const V3 sp2{x[top[i]], y[top[i]], z[top[i]]};
    const V3 b{bField[0], bField[1], bField[2]};
    // estimationFrameLocalToGlobal :20-41
    const V3 relVec = sub(sp1, sp0);
    const V3 newZ = normalized(b);
    const V3 newY = normalized(cross(newZ, relVec));
    const V3 newX = cross(newY, newZ);
    // rotation with columns newX, newY, newZ: R[r][c]
    const double R[3][3] = {{newX[0], newY[0], newZ[0]}, {newX[1], newY[1], newZ[1]}, {newX[2], newY[2], newZ[2]}};
    // general inverse
    const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                       R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    double I[3][3];
    I[0][0] = (R[1][1] * R[2][2] - R[1][2] * R[2][1]) / det;
    I[0][1] = (R[0][2] * R[2][1] - R[0][1] * R[2][2]) / det;
    I[0][2] = (R[0][1] * R[1][2] - R[0][2] * R[1][1]) / det;
    I[1][0] = (R[1][2] * R[2][0] - R[1][0] * R[2][2]) / det;
    I[1][1] = (R[0][0] * R[2][2] - R[0][2] * R[2][0]) / det;
    I[1][2] = (R[0][2] * R[1][0] - R[0][0] * R[1][2]) / det;
    I[2][0] = (R[1][0] * R[2][1] - R[1][1] * R[2][0]) / det;
    I[2][1] = (R[0][1] * R[2][0] - R[0][0] * R[2][1]) / det;
    I[2][2] = (R[0][0] * R[1][1] - R[0][1] * R[1][0]) / det;
    auto toLocal = [&](const V3& p) {
      const V3 d = sub(p, sp0);
      return V3{I[0][0] * d[0] + I[0][1] * d[1] + I[0][2] * d[2], I[1][0] * d[0] + I[1][1] * d[1] + I[1][2] * d[2],
                I[2][0] * d[0] + I[2][1] * d[1] + I[2][2] * d[2]};
    };
    const V3 local1 = toLocal(sp1), local2 = toLocal(sp2);
    // performConformalMapping :75-86
    const double n1 = local1[0] * local1[0] + local1[1] * local1[1];
    const double n2 = local2[0] * local2[0] + local2[1] * local2[1];
    const double uv1[2] = {local1[0] / n1, local1[1] / n1}, uv2[2] = {local2[0] / n2, local2[1] / n2};
    const double duv[2] = {uv2[0] - uv1[0], uv2[1] - uv1[1]};
    const double A = duv[1] / duv[0];
    const double B = uv1[1] - A * uv1[0];
    const double bOverS = (uv1[1] * uv2[0] - uv2[1] * uv1[0]) / std::sqrt(duv[0] * duv[0] + duv[1] * duv[1]);
    // computeDzDs :43-65
    auto localPhi = [&](const V3& l) {
      const double rx = 2 * B * l[0] - (-A), ry = 2 * B * l[1] - 1;
      return std::atan2(ry, rx);
    };
    const double dPhi = localPhi(local2) - localPhi(local1);
    const double dZ = local2[2] - local1[2];
    static const double eps = std::sqrt(std::numeric_limits<double>::epsilon()) * 6;  // MathHelpers.hpp:256-266
    const double hx = dPhi / 2;
    const double sincCorrection = std::abs(hx) < eps ? 1.0 : std::sin(hx) / hx;
    const double dd[2] = {local2[0] - local1[0], local2[1] - local1[1]};
    const double dzds = sincCorrection * dZ / std::sqrt(dd[0] * dd[0] + dd[1] * dd[1]);
    // computeLocalTangent at local0 = 0 :88-97
    const double r[2] = {2 * B * 0.0 - (-A), 2 * B * 0.0 - 1};
    const V3 tl = normalized(V3{-r[1], r[0], std::sqrt(r[0] * r[0] + r[1] * r[1]) * dzds});
    const V3 direction{R[0][0] * tl[0] + R[0][1] * tl[1] + R[0][2] * tl[2], R[1][0] * tl[0] + R[1][1] * tl[1] + R[1][2] * tl[2],
                       R[2][0] * tl[0] + R[2][1] * tl[1] + R[2][2] * tl[2]};
    const double qOverPt = 2 * bOverS / norm(b);
    double* o = out + 8 * i;
    o[0] = sp0[0]; o[1] = sp0[1]; o[2] = sp0[2]; o[3] = 0.0;
    o[4] = direction[0]; o[5] = direction[1]; o[6] = direction[2];
    o[7] = qOverPt / std::sqrt(1 * 1 + dzds * dzds);  // fastHypot(1, dzds)
  }
}

// Timed multi-event run for the CPU baseline: events are handed out
// dynamically to nThreads workers like the Sequencer's tbb::parallel_for over
// events (Examples/Framework/src/Framework/Sequencer.cpp:472-475), each with
// its own scratch.  Returns the total number of seeds (so the work cannot be
// optimised away); per-event seed counts go to seedCounts when not NULL.
// navStride > 1 seeds only every navStride-th middle bin of an event (a bounded
// sample: 1/navStride of the event's seeding work, grid build in full).
std::int64_t oracle_run_many(const oracle_handle* h, std::uint32_t nEvents,
                             const std::uint32_t* spOffsets, const float* x,
                             const float* y, const float* z, const float* r,
                             const float* varZ, const float* varR, int nThreads,
                             std::uint32_t navStride, std::uint64_t* seedCounts) {
  std::atomic<std::uint32_t> next{0};
  std::atomic<std::int64_t> total{0};
  std::atomic<bool> failed{false};
  auto worker = [&]() {
    for (;;) {
      const std::uint32_t e = next.fetch_add(1);
      if (e >= nEvents) break;
      try {
        Event ev;
        ev.s = &h->setup;
        const std::uint32_t o = spOffsets[e];
        ev.x = x + o; ev.y = y + o; ev.z = z + o; ev.r = r + o;
        ev.varZ = varZ + o; ev.varR = varR + o;
        ev.n = spOffsets[e + 1] - o;
        ev.navStride = navStride == 0 ? 1 : navStride;
        ev.navPhase = e % ev.navStride;
        if (h->orthogonal) {
          orth::runEvent(ev, h->orthOptions, nullptr);
        } else {
          runEvent(ev);
        }
        total += static_cast<std::int64_t>(ev.seeds.size());
        if (seedCounts != nullptr) seedCounts[e] = ev.seeds.size();
      } catch (...) {
        failed = true;
      }
    }
  };
  if (nThreads <= 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nThreads; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
  return failed ? -1 : total.load();
}

// ---------------------------------------------------------------------------
// Pixel space points from measurements ("next" row f4), restating
// createPixelSpacePoint (Examples/Algorithms/TrackFinding/src/SpacePointMaker.cpp:44-76)
// with the matrices written out the way the reference forms them:
//   global   = transform * Vector3(loc0, loc1, 0)            PlaneSurface.cpp:72-75
//   rot      = transform.matrix().block<3,3>(0,0)            Surface.cpp:243-247
//   jacXyzToZr (2x3), jac = jacXyzToZr * rot.topLeftCorner<3,2>(), jac * localCov * jac^T
//                                                            PixelSpacePointBuilder.cpp:17-42
// transforms: row-major 3x4 per surface.  Returns -1 on a bad surface index.
// ---------------------------------------------------------------------------
int oracle_make_pixel_spacepoints(std::uint32_t n, const std::uint32_t* surface, const double* loc0,
                                  const double* loc1, const double* cov00, const double* cov01,
                                  const double* cov11, std::uint32_t nSurfaces, const double* transforms,
                                  float* x, float* y, float* z, float* r, float* varZ, float* varR) {
  for (std::uint32_t i = 0; i < n; ++i) {
    if (surface[i] >= nSurfaces) return -1;
    const double* T = transforms + static_cast<std::size_t>(surface[i]) * 12;
    const double local3[3] = {loc0[i], loc1[i], 0.};
    double global[3];
    for (int a = 0; a < 3; ++a) {
      double acc = 0;
      for (int c = 0; c < 3; ++c) acc += T[a * 4 + c] * local3[c];
      global[a] = acc + T[a * 4 + 3];
    }
    const double scale = 1 / std::sqrt(global[0] * global[0] + global[1] * global[1]);
    double jacXyzToZr[2][3] = {{0, 0, 0}, {0, 0, 0}};
    jacXyzToZr[0][2] = 1;
    jacXyzToZr[1][0] = scale * global[0];
    jacXyzToZr[1][1] = scale * global[1];
    double jac[2][2];
    for (int a = 0; a < 2; ++a) {
      for (int b = 0; b < 2; ++b) {
        double acc = 0;
        for (int k = 0; k < 3; ++k) acc += jacXyzToZr[a][k] * T[k * 4 + b];
        jac[a][b] = acc;
      }
    }
    const double C2[2][2] = {{cov00[i], cov01[i]}, {cov01[i], cov11[i]}};
    double jc[2][2];
    for (int a = 0; a < 2; ++a) {
      for (int b = 0; b < 2; ++b) jc[a][b] = jac[a][0] * C2[0][b] + jac[a][1] * C2[1][b];
    }
    const double vz = jc[0][0] * jac[0][0] + jc[0][1] * jac[0][1];
    const double vr = jc[1][0] * jac[1][0] + jc[1][1] * jac[1][1];
    x[i] = static_cast<float>(global[0]);
    y[i] = static_cast<float>(global[1]);
    z[i] = static_cast<float>(global[2]);
    r[i] = static_cast<float>(std::sqrt(global[0] * global[0] + global[1] * global[1]));
    varZ[i] = static_cast<float>(vz);
    varR[i] = static_cast<float>(vr);
  }
  return 0;
}

}  // extern "C"
