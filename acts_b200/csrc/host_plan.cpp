// host_plan.cpp -- see host_plan.hpp.  Compiled for the host only, without
// fast-math and without FMA contraction, so every float below rounds exactly
// like the reference's constructor code does on x86-64.
#include "host_plan.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numbers>

namespace B200SEED_NS {

namespace {

constexpr float kPiF = std::numbers::pi_v<float>;

struct Fail {
  int code;
  std::string msg;
};

float sq(float v) { return v * v; }

// Acts::fastCathetus(h, 1) with an int argument (MathHelpers.hpp:86-103,177-180)
float cathetusOne(float h) { return std::sqrt((h - 1) * (h + 1)); }

// Number of phi bins of the grid: the arithmetic of SpacePointGridPhiBinning.cpp:20-95 (same operations, same
// float / double types), condensed -- the azimuth a minimum-pT helix sweeps between (rMax - deltaRMax) and rMax plus
// the spread an impact parameter of impactMax adds, divided by phiBinDeflectionCoverage, is one bin width.
int phiBinCount(const b200seed_config& c) {
  if (c.bFieldInZ == 0) return c.maxPhiBins;
  const float helixR = c.minPt / c.bFieldInZ;
  if (helixR < c.rMax * 0.5) {
    throw Fail{B200SEED_ERR_DOMAIN, "phi binning: the minimum-pT helix (minPt / bFieldInZ) does not reach rMax"};
  }
  auto sweep = [&](float radius) { return std::atan(1.f / cathetusOne(2 * helixR / radius)); };
  const float sweepOuter = sweep(c.rMax);
  float sweepInner = 0;
  float innerR = c.rMax;
  if (c.rMax > c.deltaRMax) {
    const float r0 = c.rMax - c.deltaRMax;
    innerR = r0;
    sweepInner = sweep(r0);
  }
  const float impactSpread = std::abs(std::asin(std::min(1.f, c.impactMax / innerR)) - std::asin(std::min(1.f, c.impactMax / c.rMax)));
  const float binWidth = (sweepOuter - sweepInner + impactSpread) / c.phiBinDeflectionCoverage;
  if (binWidth <= 0.f) {
    throw Fail{B200SEED_ERR_DOMAIN, "phi binning: non-positive bin width (check rMax, deltaRMax, impactMax, phiBinDeflectionCoverage)"};
  }
  return std::min(static_cast<int>(std::ceil(2 * std::numbers::pi / binWidth)), c.maxPhiBins);
}

int wrapClosed(int bin, int w) { return 1 + (w + ((bin - 1) % w)) % w; }

// Axis.hpp:150-188 (closed axis neighbourhood, wrap order preserved)
std::vector<int> closedNeighbors(int idx, int first, int second, int nBins) {
  std::vector<int> out;
  if (idx <= 0 || idx >= nBins + 1) return out;
  const int max = nBins;
  first = std::clamp(first, -max, max);
  second = std::clamp(second, -max, max);
  if (std::abs(first - second) >= max) {
    first = 1 - idx;
    second = max - idx;
  }
  const int itfirst = wrapClosed(idx + first, max);
  const int itlast = wrapClosed(idx + second, max);
  if (itfirst <= itlast) {
    for (int b = itfirst; b <= itlast; ++b) out.push_back(b);
  } else {
    for (int b = itfirst; b <= max; ++b) out.push_back(b);
    for (int b = 1; b <= itlast; ++b) out.push_back(b);
  }
  return out;
}
// Axis.hpp:413-423 (variable open axis: under/overflow bins are legal neighbours)
std::vector<int> openNeighbors(int idx, int first, int second, int nBins) {
  std::vector<int> out;
  const int itmin = std::max(0, idx + first);
  const int itmax = std::min(nBins + 1, idx + second);
  for (int b = itmin; b <= itmax; ++b) out.push_back(b);
  return out;
}

void completeNavigation(std::vector<std::size_t>& bins, std::size_t nBins, int axis) {
  if (bins.empty()) {
    for (std::size_t b = 1; b <= nBins; ++b) bins.push_back(b);
    return;
  }
  std::vector<bool> visited(nBins + 1, false);
  for (std::size_t bin : bins) {
    if (bin == 0 || bin > nBins) {
      throw Fail{B200SEED_ERR_INVALID_ARGUMENT,
                 "Invalid navigation for axis " + std::to_string(axis) + ": bin " +
                     std::to_string(bin) +
                     " is out of range. Local bin indices are 1-based and must lie "
                     "within [1, " + std::to_string(nBins) + "]."};
    }
    if (visited[bin]) {
      throw Fail{B200SEED_ERR_INVALID_ARGUMENT,
                 "Invalid navigation for axis " + std::to_string(axis) + ": bin " +
                     std::to_string(bin) + " is listed more than once."};
    }
    visited[bin] = true;
  }
}

void build(const b200seed_config& c, HostPlan& plan, const b200seed_orthogonal_options* orthOpt = nullptr) {
  const bool orthogonal = orthOpt != nullptr;
  plan.orthogonal = orthogonal;
  if (c.abi_version != B200SEED_ABI_VERSION || c.struct_size != sizeof(b200seed_config)) {
    throw Fail{B200SEED_ERR_INVALID_ARGUMENT,
               "b200seed_config: abi_version/struct_size mismatch (use b200seed_config_init)"};
  }
  if ((c.nZBinNeighborsTop && !c.zBinNeighborsTop) ||
      (c.nZBinNeighborsBottom && !c.zBinNeighborsBottom) || (c.nZBinEdges && !c.zBinEdges) ||
      (c.nZBinsCustomLooping && !c.zBinsCustomLooping) || (c.nRRangeMiddleSP && !c.rRangeMiddleSP)) {
    throw Fail{B200SEED_ERR_INVALID_ARGUMENT, "b200seed_config: null array with non-zero count"};
  }
  const float gridRMin = 0;
  if (!orthogonal) {
  // GridTripletSeedingAlgorithm.cpp:117-126
  for (uint32_t i = 0; i < c.nZBinsCustomLooping; ++i) {
    if (c.zBinsCustomLooping[i] >= c.nZBinEdges) {
      throw Fail{B200SEED_ERR_INVALID_ARGUMENT,
                 "Inconsistent config zBinsCustomLooping does not contain a subset "
                 "of bins defined by zBinEdges"};
    }
  }
  // CylindricalSpacePointGrid.cpp:18-37; the algorithm forces the grid's rMin to 0
  if (c.phiMin < -kPiF || c.phiMax > kPiF) {
    throw Fail{B200SEED_ERR_RUNTIME,
               "CylindricalSpacePointGrid: phiMin (" + std::to_string(c.phiMin) +
                   ") and/or phiMax (" + std::to_string(c.phiMax) +
                   ") are outside the allowed phi range, defined as "
                   "[-std::numbers::pi_v<float>, std::numbers::pi_v<float>]"};
  }
  if (c.phiMin > c.phiMax) {
    throw Fail{B200SEED_ERR_RUNTIME, "CylindricalSpacePointGrid: phiMin is bigger then phiMax"};
  }
  if (gridRMin > c.rMax) {
    throw Fail{B200SEED_ERR_RUNTIME, "CylindricalSpacePointGrid: rMin is bigger then rMax"};
  }
  if (c.zMin > c.zMax) {
    throw Fail{B200SEED_ERR_RUNTIME, "CylindricalSpacePointGrid: zMin is bigger than zMax"};
  }

  }
  DeviceConfig& d = plan.dev;
  std::memset(&d, 0, sizeof(d));
  plan.navBins.clear();
  plan.botOffsets = {0};
  plan.topOffsets = {0};
  plan.botBins.clear();
  plan.topBins.clear();
  plan.maxNeighborBins = 0;
  if (!orthogonal) {
  const int phiBins = phiBinCount(c);

  // phi axis, Axis.hpp:40-58
  d.phiMin = c.phiMin;
  d.phiMax = c.phiMax;
  if (d.phiMin >= d.phiMax) {
    throw Fail{B200SEED_ERR_INVALID_ARGUMENT,
               "Axis: Invalid axis range, min edge needs to be smaller than max edge."};
  }
  if (phiBins < 1) {
    throw Fail{B200SEED_ERR_INVALID_ARGUMENT, "Axis: Invalid binning, at least one bin is needed."};
  }
  d.phiBins = phiBins;
  d.phiWidth = (d.phiMax - d.phiMin) / static_cast<double>(phiBins);

  // z axis, CylindricalSpacePointGrid.cpp:49-76
  std::vector<double> zValues;
  if (c.nZBinEdges == 0) {
    const float zBinSize = c.cotThetaMax * c.deltaRMax;
    const float zBins = std::max(1.f, std::floor((c.zMax - c.zMin) / zBinSize));
    for (int bin = 0; bin <= static_cast<int>(zBins); bin++) {
      const double edge = c.zMin + bin * ((c.zMax - c.zMin) / zBins);
      zValues.push_back(edge);
    }
  } else {
    for (uint32_t i = 0; i < c.nZBinEdges; ++i) zValues.push_back(c.zBinEdges[i]);
  }
  if (zValues.size() < 2 || !std::is_sorted(zValues.begin(), zValues.end())) {
    throw Fail{B200SEED_ERR_INVALID_ARGUMENT, "Axis: z bin edges must be sorted and at least two"};
  }
  if (zValues.size() > static_cast<std::size_t>(kMaxZEdges)) {
    throw Fail{B200SEED_ERR_UNSUPPORTED, "more than " + std::to_string(kMaxZEdges - 1) + " z bins"};
  }
  d.nZ = static_cast<int>(zValues.size()) - 1;
  for (std::size_t i = 0; i < zValues.size(); ++i) d.zEdges[i] = zValues[i];
  // r axis: {rMin = 0, rMax} (rBinEdges = {}, .cpp:144)
  d.rAxisMin = gridRMin;
  d.rAxisMax = c.rMax;
  if (!(d.rAxisMin < d.rAxisMax)) {
    throw Fail{B200SEED_ERR_INVALID_ARGUMENT, "Axis: invalid r axis range"};
  }
  d.nR = 1;
  d.nGlobalBins = (d.phiBins + 2) * (d.nZ + 2) * (d.nR + 2);

  // navigation
  std::vector<std::size_t> navPhi, navZ(c.zBinsCustomLooping, c.zBinsCustomLooping + c.nZBinsCustomLooping), navR;
  completeNavigation(navPhi, static_cast<std::size_t>(d.phiBins), 0);
  completeNavigation(navZ, static_cast<std::size_t>(d.nZ), 1);
  completeNavigation(navR, static_cast<std::size_t>(d.nR), 2);

  // GridBinFinder values: a non-empty per-bin vector must cover every z bin
  // (the reference only asserts this, GridBinFinder.ipp:56-58)
  for (uint32_t n : {c.nZBinNeighborsTop, c.nZBinNeighborsBottom}) {
    if (n != 0 && n < static_cast<uint32_t>(d.nZ)) {
      throw Fail{B200SEED_ERR_INVALID_ARGUMENT,
                 "zBinNeighbors{Top,Bottom} must be empty or hold one pair per z bin"};
    }
  }
  if (c.numPhiNeighbors < 0) {
    throw Fail{B200SEED_ERR_INVALID_ARGUMENT, "numPhiNeighbors must be >= 0"};
  }

  auto findBins = [&](int phiLoc, int zLoc, int rLoc, bool top) {
    const int32_t* zNb = top ? c.zBinNeighborsTop : c.zBinNeighborsBottom;
    const uint32_t nZNb = top ? c.nZBinNeighborsTop : c.nZBinNeighborsBottom;
    int zFirst = -1, zSecond = 1;  // empty vector -> int 1 (GridBinFinder.ipp:30-34)
    if (nZNb != 0) {
      zFirst = zNb[2 * (zLoc - 1)];
      zSecond = zNb[2 * (zLoc - 1) + 1];
    }
    // Under-/overflow bins of the open z and r axes are legal neighbours in the
    // reference (Axis.hpp:413-423) but can never hold a space point because
    // insert() requires isInside on every axis (SpacePointGridBase.hpp:67-87):
    // dropping them keeps the emission order of everything that can exist.
    std::vector<uint32_t> out;
    for (int p : closedNeighbors(phiLoc, -c.numPhiNeighbors, c.numPhiNeighbors, d.phiBins))
      for (int z : openNeighbors(zLoc, zFirst, zSecond, d.nZ))
        for (int r : openNeighbors(rLoc, 0, 0, d.nR)) {
          if (z < 1 || z > d.nZ || r < 1 || r > d.nR) continue;
          out.push_back(static_cast<uint32_t>((p * (d.nZ + 2) + z) * (d.nR + 2) + r));
        }
    return out;
  };

  plan.navBins.clear();
  plan.botOffsets = {0};
  plan.topOffsets = {0};
  plan.botBins.clear();
  plan.topBins.clear();
  plan.maxNeighborBins = 0;
  for (std::size_t p : navPhi)
    for (std::size_t z : navZ)
      for (std::size_t r : navR) {
        plan.navBins.push_back(static_cast<uint32_t>((p * (d.nZ + 2) + z) * (d.nR + 2) + r));
        const auto bot = findBins(static_cast<int>(p), static_cast<int>(z), static_cast<int>(r), false);
        const auto top = findBins(static_cast<int>(p), static_cast<int>(z), static_cast<int>(r), true);
        plan.botBins.insert(plan.botBins.end(), bot.begin(), bot.end());
        plan.topBins.insert(plan.topBins.end(), top.begin(), top.end());
        plan.botOffsets.push_back(static_cast<uint32_t>(plan.botBins.size()));
        plan.topOffsets.push_back(static_cast<uint32_t>(plan.topBins.size()));
        plan.maxNeighborBins = std::max<uint32_t>(plan.maxNeighborBins, static_cast<uint32_t>(std::max(bot.size(), top.size())));
      }
  if (plan.maxNeighborBins > static_cast<uint32_t>(kMaxNeighborBins)) {
    throw Fail{B200SEED_ERR_UNSUPPORTED,
               "more than " + std::to_string(kMaxNeighborBins) + " neighbour bins per middle bin"};
  }

  }
  // doublet finders, GridTripletSeedingAlgorithm.cpp:272-309
  d.dRMinB = std::isnan(c.deltaRMinBottom) ? c.deltaRMin : c.deltaRMinBottom;
  d.dRMaxB = std::isnan(c.deltaRMaxBottom) ? c.deltaRMax : c.deltaRMaxBottom;
  d.dRMinT = std::isnan(c.deltaRMinTop) ? c.deltaRMin : c.deltaRMinTop;
  d.dRMaxT = std::isnan(c.deltaRMaxTop) ? c.deltaRMax : c.deltaRMaxTop;
  if (orthogonal) {
    // OrthogonalTripletSeedingAlgorithm.cpp:176-199: the doublet finders test deltaRMax{Bottom,Top} for NaN to
    // pick BOTH bounds; the tree options (.cpp:158-171) test each bound on its own, the "low-high" set with the
    // Bottom values and the "high-low" set with the Top values
    d.dRMinB = std::isnan(c.deltaRMaxBottom) ? c.deltaRMin : c.deltaRMinBottom;
    d.dRMinT = std::isnan(c.deltaRMaxTop) ? c.deltaRMin : c.deltaRMinTop;
    OrthDeviceConfig& o = plan.orth;
    o.rMax = c.rMax; o.zMin = c.zMin; o.zMax = c.zMax; o.phiMin = c.phiMin; o.phiMax = c.phiMax;
    o.lhDeltaRMin = std::isnan(c.deltaRMinBottom) ? c.deltaRMin : c.deltaRMinBottom;
    o.lhDeltaRMax = std::isnan(c.deltaRMaxBottom) ? c.deltaRMax : c.deltaRMaxBottom;
    o.hlDeltaRMin = std::isnan(c.deltaRMinTop) ? c.deltaRMin : c.deltaRMinTop;
    o.hlDeltaRMax = std::isnan(c.deltaRMaxTop) ? c.deltaRMax : c.deltaRMaxTop;
    o.collisionRegionMin = c.collisionRegionMin; o.collisionRegionMax = c.collisionRegionMax;
    o.cotThetaMax = c.cotThetaMax;
    o.deltaPhiMax = orthOpt->deltaPhiMax;
    o.deltaZMax = std::numeric_limits<float>::infinity();  // Options::deltaZMax default, never set by the algorithm
    o.zOutermostLayersMin = orthOpt->zOutermostLayersMin;
    o.zOutermostLayersMax = orthOpt->zOutermostLayersMax;
  }
  d.deltaZMin = c.deltaZMin;
  d.deltaZMax = c.deltaZMax;
  d.collisionRegionMin = c.collisionRegionMin;
  d.collisionRegionMax = c.collisionRegionMax;
  d.cotThetaMax = c.cotThetaMax;
  d.impactMax = c.impactMax;
  d.interactionPointCut = c.interactionPointCut ? 1 : 0;
  d.useExtraCuts = c.useExtraCuts ? 1 : 0;
  d.doubletCuts = c.useExtraCuts ? kCutsItk : kCutsNone;  // VertexZ is chosen per event
  {
    // DoubletSeedFinder.cpp:351-357
    const float pTPerHelixRadius = c.bFieldInZ;
    d.minHelixDiameter2Doublet = sq(c.minPt * 2 / pTPerHelixRadius) * c.helixCutTolerance;
  }
  float highland = 0;
  {
    // TripletSeedFinder.cpp:456-481
    const double xOverX0 = c.radLengthPerSeed;
    const double q2OverBeta2 = 1;
    const double t = std::sqrt(xOverX0 * q2OverBeta2);
    const double e136 = static_cast<double>(1e-3 * 13.6L);  // 13.6_MeV, Units.hpp:149,181-184
    highland = static_cast<float>(e136 * t * (1.0 + 0.038 * 2 * std::log(t)));
    const float maxScatteringAngle = highland / c.minPt;
    const float maxScatteringAngle2 = maxScatteringAngle * maxScatteringAngle;
    const float pTPerHelixRadius = c.bFieldInZ;
    d.minHelixDiameter2 = sq(c.minPt * 2 / pTPerHelixRadius) * c.helixCutTolerance;
    const float pT2perRadius = sq(highland / pTPerHelixRadius);
    d.sigmapT2perRadius = pT2perRadius * sq(2 * c.sigmaScattering);
    d.multipleScattering2 = maxScatteringAngle2 * sq(c.sigmaScattering);
  }

  // filter, GridTripletSeedingAlgorithm.cpp:156-173
  d.deltaInvHelixDiameter = c.deltaInvHelixDiameter;
  d.filterDeltaRMin = c.deltaRMin;
  d.compatSeedWeight = c.compatSeedWeight;
  d.impactWeightFactor = c.impactWeightFactor;
  d.seedWeightIncrement = c.seedWeightIncrement;
  d.numSeedIncrement = c.numSeedIncrement;
  if (c.compatSeedLimit > static_cast<uint64_t>(kMaxCompatSeedLimit)) {
    throw Fail{B200SEED_ERR_UNSUPPORTED,
               "compatSeedLimit > " + std::to_string(kMaxCompatSeedLimit)};
  }
  d.compatSeedLimit = static_cast<uint32_t>(c.compatSeedLimit);
  d.maxSeedsPerSpM = c.maxSeedsPerSpM;
  if (c.maxSeedsPerSpMConf > static_cast<uint32_t>(kMaxHeapBig)) {
    throw Fail{B200SEED_ERR_UNSUPPORTED, "maxSeedsPerSpMConf > " + std::to_string(kMaxHeapBig)};
  }
  // a middle returns at most min(collector capacity, maxSeedsPerSpM + 1) seeds (BroadTripletSeedFilter.cpp:336-348)
  if (std::min<uint64_t>(c.maxSeedsPerSpMConf, static_cast<uint64_t>(c.maxSeedsPerSpM) + 1) > static_cast<uint64_t>(kMaxHeap)) {
    throw Fail{B200SEED_ERR_UNSUPPORTED, "min(maxSeedsPerSpMConf, maxSeedsPerSpM + 1) > " + std::to_string(kMaxHeap)};
  }
  d.maxSeedsPerSpMConf = c.maxSeedsPerSpMConf;
  d.useDeltaRinsteadOfTopRadius = c.useDeltaRinsteadOfTopRadius ? 1 : 0;
  // seed confirmation, GridTripletSeedingAlgorithm.cpp:163-171
  d.seedConfirmation = c.seedConfirmation ? 1 : 0;
  d.zOriginWeightFactor = c.zOriginWeightFactor;
  if (c.seedConfirmation && c.maxQualitySeedsPerSpMConf > static_cast<uint32_t>(kMaxHeapBig)) {
    throw Fail{B200SEED_ERR_UNSUPPORTED, "maxQualitySeedsPerSpMConf > " + std::to_string(kMaxHeapBig)};
  }
  if (c.seedConfirmation && std::min<uint64_t>(static_cast<uint64_t>(c.maxSeedsPerSpMConf) + c.maxQualitySeedsPerSpMConf,
                                               static_cast<uint64_t>(c.maxSeedsPerSpM) + 1) > 2u * static_cast<uint64_t>(kMaxHeap)) {
    throw Fail{B200SEED_ERR_UNSUPPORTED, "seedConfirmation: more than " + std::to_string(2 * kMaxHeap) + " seeds per middle"};
  }
  d.maxQualitySeedsPerSpMConf = c.maxQualitySeedsPerSpMConf;
  {
    const b200seed_seed_confirmation_range* src[2] = {&c.centralSeedConfirmationRange, &c.forwardSeedConfirmationRange};
    ConfRange* dst[2] = {&d.confCentral, &d.confForward};
    for (int i = 0; i < 2; ++i) {
      if (c.seedConfirmation && (src[i]->nTopForLargeR > 0x7fffffffull || src[i]->nTopForSmallR > 0x7fffffffull)) {
        throw Fail{B200SEED_ERR_UNSUPPORTED, "nTopForLargeR / nTopForSmallR beyond 2^31 - 1"};
      }
      dst[i]->zMinSeedConf = src[i]->zMinSeedConf;
      dst[i]->zMaxSeedConf = src[i]->zMaxSeedConf;
      dst[i]->rMaxSeedConf = src[i]->rMaxSeedConf;
      dst[i]->nTopForLargeR = static_cast<uint32_t>(std::min<uint64_t>(src[i]->nTopForLargeR, 0x7fffffffull));
      dst[i]->nTopForSmallR = static_cast<uint32_t>(std::min<uint64_t>(src[i]->nTopForSmallR, 0x7fffffffull));
      dst[i]->seedConfMinBottomRadius = src[i]->seedConfMinBottomRadius;
      dst[i]->seedConfMaxZOrigin = src[i]->seedConfMaxZOrigin;
      dst[i]->minImpactSeedConf = src[i]->minImpactSeedConf;
    }
  }

  // middle r range, GridTripletSeedingAlgorithm.cpp:404-421
  d.useVariableMiddleSPRange = c.useVariableMiddleSPRange ? 1 : 0;
  d.rMinMiddle = c.rMinMiddle;
  d.rMaxMiddle = c.rMaxMiddle;
  d.deltaRMiddleMinSPRange = c.deltaRMiddleMinSPRange;
  d.deltaRMiddleMaxSPRange = c.deltaRMiddleMaxSPRange;
  d.nRRangeMiddleSP = orthogonal ? 0 : static_cast<int32_t>(c.nRRangeMiddleSP);
  if (!orthogonal && c.nRRangeMiddleSP > static_cast<uint32_t>(kMaxZEdges)) {
    throw Fail{B200SEED_ERR_UNSUPPORTED, "rRangeMiddleSP too long"};
  }
  for (uint32_t i = 0; !orthogonal && i < 2 * c.nRRangeMiddleSP; ++i) d.rRangeMiddleSP[i] = c.rRangeMiddleSP[i];
  d.nZBinEdgesF = orthogonal ? 0 : static_cast<int32_t>(c.nZBinEdges);
  for (uint32_t i = 0; !orthogonal && i < c.nZBinEdges; ++i) d.zBinEdgesF[i] = c.zBinEdges[i];
  if (!orthogonal && c.nRRangeMiddleSP != 0 && !c.useVariableMiddleSPRange) {
    // the reference indexes rRangeMiddleSP[zBin] unchecked (.cpp:415-420) with
    // zBin = max(lower_bound(zBinEdges, zM) - 1, 0) <= max(nZBinEdges, 1) - 1
    const uint32_t needed = c.nZBinEdges > 1 ? c.nZBinEdges - 1 : 1;
    if (c.nRRangeMiddleSP < needed) {
      throw Fail{B200SEED_ERR_INVALID_ARGUMENT, "rRangeMiddleSP shorter than the number of z bins"};
    }
  }

  // BroadTripletSeedFilter.cpp:336-348: at most maxSeedsPerSpM + 1 of the collector's candidates
  const uint32_t collectorMax = c.maxSeedsPerSpMConf + (c.seedConfirmation ? c.maxQualitySeedsPerSpMConf : 0u);
  plan.seedsPerMiddle = std::min<uint32_t>(
      collectorMax, c.maxSeedsPerSpM == std::numeric_limits<uint32_t>::max() ? collectorMax : c.maxSeedsPerSpM + 1);
  plan.relaxedFloat = c.relaxedFloat != 0;
  plan.useVertexZCuts = c.useVertexZCuts != 0;
  plan.vertexZNSigma = c.vertexZNSigma;
  plan.vertexZMargin = c.vertexZMargin;
  plan.toleranceParam = c.toleranceParam;

  b200seed_info& info = plan.info;
  std::memset(&info, 0, sizeof(info));
  info.phiBins = d.phiBins;
  info.zBins = d.nZ;
  info.rBins = d.nR;
  info.nGlobalBins = d.nGlobalBins;
  info.minHelixDiameter2 = d.minHelixDiameter2;
  info.highland = highland;
  info.sigmapT2perRadius = d.sigmapT2perRadius;
  info.multipleScattering2 = d.multipleScattering2;
  info.deltaRMinBottom = d.dRMinB;
  info.deltaRMaxBottom = d.dRMaxB;
  info.deltaRMinTop = d.dRMinT;
  info.deltaRMaxTop = d.dRMaxT;
}

}  // namespace

bool make_host_plan(const b200seed_config& cfg, HostPlan& plan, PlanError& err) {
  try {
    build(cfg, plan);
    return true;
  } catch (const Fail& f) {
    err.code = f.code;
    err.message = f.msg;
    return false;
  }
}

bool make_orthogonal_plan(const b200seed_config& cfg, const b200seed_orthogonal_options& opt, HostPlan& plan, PlanError& err) {
  try {
    build(cfg, plan, &opt);
    return true;
  } catch (const Fail& f) {
    err.code = f.code;
    err.message = f.msg;
    return false;
  }
}

void config_defaults(b200seed_config& c) {
  std::memset(&c, 0, sizeof(c));
  c.abi_version = B200SEED_ABI_VERSION;
  c.struct_size = sizeof(b200seed_config);
  constexpr double T = 0.000299792458;  // Units.hpp:168
  c.bFieldInZ = 2 * T;
  c.minPt = 0.4;
  c.cotThetaMax = 10.01788;
  c.impactMax = 20;
  c.deltaRMin = 5;
  c.deltaRMax = 270;
  c.deltaRMinTop = std::numeric_limits<float>::quiet_NaN();
  c.deltaRMaxTop = std::numeric_limits<float>::quiet_NaN();
  c.deltaRMinBottom = std::numeric_limits<float>::quiet_NaN();
  c.deltaRMaxBottom = std::numeric_limits<float>::quiet_NaN();
  c.rMin = 0;
  c.rMax = 600;
  c.zMin = -2800;
  c.zMax = 2800;
  c.phiMin = -kPiF;
  c.phiMax = kPiF;
  c.phiBinDeflectionCoverage = 1;
  c.maxPhiBins = 10000;
  c.numPhiNeighbors = 1;
  c.rMinMiddle = 60;
  c.rMaxMiddle = 120;
  c.useVariableMiddleSPRange = 0;
  c.deltaRMiddleMinSPRange = 10;
  c.deltaRMiddleMaxSPRange = 10;
  c.deltaZMin = -std::numeric_limits<float>::infinity();
  c.deltaZMax = std::numeric_limits<float>::infinity();
  c.interactionPointCut = 0;
  c.collisionRegionMin = -150;
  c.collisionRegionMax = +150;
  c.helixCutTolerance = 1;
  c.sigmaScattering = 5;
  c.radLengthPerSeed = 0.05;
  c.toleranceParam = 1.1;
  c.deltaInvHelixDiameter = 0.00003;
  c.compatSeedWeight = 200;
  c.impactWeightFactor = 1;
  c.zOriginWeightFactor = 1;
  c.maxSeedsPerSpM = 5;
  c.compatSeedLimit = 2;
  c.seedWeightIncrement = 0;
  c.numSeedIncrement = std::numeric_limits<float>::infinity();
  c.seedConfirmation = 0;
  for (b200seed_seed_confirmation_range* r :
       {&c.centralSeedConfirmationRange, &c.forwardSeedConfirmationRange}) {
    r->zMinSeedConf = std::numeric_limits<float>::lowest();
    r->zMaxSeedConf = std::numeric_limits<float>::max();
    r->rMaxSeedConf = std::numeric_limits<float>::max();
    r->nTopForLargeR = 0;
    r->nTopForSmallR = 0;
    r->seedConfMinBottomRadius = 60.;
    r->seedConfMaxZOrigin = 150.;
    r->minImpactSeedConf = 1.;
  }
  c.maxSeedsPerSpMConf = 5;
  c.maxQualitySeedsPerSpMConf = 5;
  c.useDeltaRinsteadOfTopRadius = 0;
  c.useExtraCuts = 0;
  c.useVertexZCuts = 0;
  c.vertexZNSigma = 3.0;
  c.vertexZMargin = 0.0;
  c.relaxedFloat = 0;
}

}  // namespace B200SEED_NS
