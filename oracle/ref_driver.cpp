// ref_driver.cpp -- C entry points around the UNMODIFIED reference seeding sources.
//
// TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's
// cpu_baseline / --impl reference legs).  The product package never loads this.
//
// oracle/_ref/libseeding_ref.so (recipe: oracle/Makefile, target `ref`) is built from
// the reference's own translation units, compiled where they lie under
// /root/reference, nothing copied:
//   Examples/Algorithms/TrackFinding/src/GridTripletSeedingAlgorithm.cpp   (the boundary, execute())
//   Examples/Algorithms/TrackFinding/src/OrthogonalTripletSeedingAlgorithm.cpp + Core/src/Seeding/
//                     CylindricalSpacePointKDTree.cpp + Core/src/Geometry/Extent.cpp (the k-d-tree seeder)
//   Core/src/Seeding/{DoubletSeedFinder,TripletSeedFinder,BroadTripletSeedFilter,TripletSeeder,
//                     CylindricalSpacePointGrid}.cpp, Core/src/Seeding/detail/{CandidatesForMiddleSp,
//                     SpacePointGridPhiBinning}.cpp
//   Core/src/Utilities/Logger.cpp, Examples/Framework/src/Framework/{IAlgorithm,SequenceElement,
//                     DataHandle,WhiteBoard}.cpp
// against stand-in headers for the third-party pieces that are absent from this
// image (oracle/ref_shim: the corner of Eigen / boost::container / boost::mp11 the
// seeding path instantiates, and ActsExamples/EventData/Vertex.hpp, whose real
// version drags in the track-parameter headers).  This file only does what the
// Examples Sequencer does around the algorithm: put the input on a WhiteBoard,
// call execute(), read the SeedContainer back.
//
// Flags: the reference's default build (RelWithDebInfo: -O2 -g, C++20, no -march,
// no fast-math; cmake/ActsCompilerOptions.cmake:2-14).
#include "../include/acts_b200_seeding.h"

#include "Acts/EventData/SeedContainer.hpp"
#include "Acts/EventData/SpacePointContainer.hpp"
#include "Acts/Seeding/BroadTripletSeedFilter.hpp"
#include "Acts/Seeding/CylindricalSpacePointGrid.hpp"
#include "Acts/Seeding/DoubletSeedFinder.hpp"
#include "Acts/Seeding/TripletSeedFinder.hpp"
#include "Acts/Seeding/TripletSeeder.hpp"
#include "Acts/Utilities/Logger.hpp"
#include "ActsExamples/EventData/Seed.hpp"
#include "ActsExamples/EventData/SpacePoint.hpp"
#include "ActsExamples/EventData/Vertex.hpp"
#include "ActsExamples/Framework/AlgorithmContext.hpp"
#include "ActsExamples/Framework/DataHandle.hpp"
#include "ActsExamples/Framework/IAlgorithm.hpp"
#include "ActsExamples/Framework/WhiteBoard.hpp"
// ref_run_strips (bottom of this file) reads what the reference's constructor derived (m_gridConfig, m_filterConfig,
// m_seedFinder, ...): the class is included with its private section opened.  Nothing else depends on it.
#define private public
#include "ActsExamples/TrackFinding/GridTripletSeedingAlgorithm.hpp"
#undef private
#include "ActsExamples/TrackFinding/OrthogonalTripletSeedingAlgorithm.hpp"

#include <atomic>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

using Algorithm = ActsExamples::GridTripletSeedingAlgorithm;

thread_local std::string g_error;

// Writes the event input to the WhiteBoard and reads the seeds back, through the
// framework's own data handles (what a reader / writer algorithm would do).
class Harness final : public ActsExamples::IAlgorithm {
 public:
  Harness(const std::string& spKey, const std::string& vtxKey, const std::string& seedKey)
      : ActsExamples::IAlgorithm("RefDriverHarness", Acts::getDefaultLogger("RefDriverHarness", Acts::Logging::WARNING)) {
    m_spacePoints.initialize(spKey);
    if (!vtxKey.empty()) m_vertices.initialize(vtxKey);
    m_seeds.initialize(seedKey);
  }
  ActsExamples::ProcessCode execute(const ActsExamples::AlgorithmContext&) const override {
    return ActsExamples::ProcessCode::SUCCESS;
  }
  void put(ActsExamples::WhiteBoard& wb, ActsExamples::SpacePointContainer&& sps) const { m_spacePoints(wb, std::move(sps)); }
  void put(ActsExamples::WhiteBoard& wb, ActsExamples::VertexContainer&& v) const { m_vertices(wb, std::move(v)); }
  const ActsExamples::SeedContainer& seeds(const ActsExamples::WhiteBoard& wb) const { return m_seeds(wb); }

 private:
  ActsExamples::WriteDataHandle<ActsExamples::SpacePointContainer> m_spacePoints{this, "OutputSpacePoints"};
  ActsExamples::WriteDataHandle<ActsExamples::VertexContainer> m_vertices{this, "OutputVertices"};
  ActsExamples::ReadDataHandle<ActsExamples::SeedContainer> m_seeds{this, "InputSeeds"};
};

Acts::SeedConfirmationRangeConfig toRange(const b200seed_seed_confirmation_range& r) {
  Acts::SeedConfirmationRangeConfig o;
  o.zMinSeedConf = r.zMinSeedConf;
  o.zMaxSeedConf = r.zMaxSeedConf;
  o.rMaxSeedConf = r.rMaxSeedConf;
  o.nTopForLargeR = r.nTopForLargeR;
  o.nTopForSmallR = r.nTopForSmallR;
  o.seedConfMinBottomRadius = r.seedConfMinBottomRadius;
  o.seedConfMaxZOrigin = r.seedConfMaxZOrigin;
  o.minImpactSeedConf = r.minImpactSeedConf;
  return o;
}

Algorithm::Config toConfig(const b200seed_config& c) {
  Algorithm::Config o;
  o.inputSpacePoints = "spacepoints";
  o.outputSeeds = "seeds";
  o.bFieldInZ = c.bFieldInZ;
  o.minPt = c.minPt;
  o.cotThetaMax = c.cotThetaMax;
  o.impactMax = c.impactMax;
  o.deltaRMin = c.deltaRMin;
  o.deltaRMax = c.deltaRMax;
  o.deltaRMinTop = c.deltaRMinTop;
  o.deltaRMaxTop = c.deltaRMaxTop;
  o.deltaRMinBottom = c.deltaRMinBottom;
  o.deltaRMaxBottom = c.deltaRMaxBottom;
  o.rMin = c.rMin;
  o.rMax = c.rMax;
  o.zMin = c.zMin;
  o.zMax = c.zMax;
  o.phiMin = c.phiMin;
  o.phiMax = c.phiMax;
  o.phiBinDeflectionCoverage = c.phiBinDeflectionCoverage;
  o.maxPhiBins = c.maxPhiBins;
  for (uint32_t i = 0; i < c.nZBinNeighborsTop; ++i) o.zBinNeighborsTop.emplace_back(c.zBinNeighborsTop[2 * i], c.zBinNeighborsTop[2 * i + 1]);
  for (uint32_t i = 0; i < c.nZBinNeighborsBottom; ++i) o.zBinNeighborsBottom.emplace_back(c.zBinNeighborsBottom[2 * i], c.zBinNeighborsBottom[2 * i + 1]);
  o.numPhiNeighbors = c.numPhiNeighbors;
  o.zBinEdges.assign(c.zBinEdges, c.zBinEdges + c.nZBinEdges);
  for (uint32_t i = 0; i < c.nZBinsCustomLooping; ++i) o.zBinsCustomLooping.push_back(static_cast<std::size_t>(c.zBinsCustomLooping[i]));
  o.rMinMiddle = c.rMinMiddle;
  o.rMaxMiddle = c.rMaxMiddle;
  o.useVariableMiddleSPRange = c.useVariableMiddleSPRange != 0;
  for (uint32_t i = 0; i < c.nRRangeMiddleSP; ++i) o.rRangeMiddleSP.push_back({c.rRangeMiddleSP[2 * i], c.rRangeMiddleSP[2 * i + 1]});
  o.deltaRMiddleMinSPRange = c.deltaRMiddleMinSPRange;
  o.deltaRMiddleMaxSPRange = c.deltaRMiddleMaxSPRange;
  o.deltaZMin = c.deltaZMin;
  o.deltaZMax = c.deltaZMax;
  o.interactionPointCut = c.interactionPointCut != 0;
  o.collisionRegionMin = c.collisionRegionMin;
  o.collisionRegionMax = c.collisionRegionMax;
  o.helixCutTolerance = c.helixCutTolerance;
  o.sigmaScattering = c.sigmaScattering;
  o.radLengthPerSeed = c.radLengthPerSeed;
  o.toleranceParam = c.toleranceParam;
  o.deltaInvHelixDiameter = c.deltaInvHelixDiameter;
  o.compatSeedWeight = c.compatSeedWeight;
  o.impactWeightFactor = c.impactWeightFactor;
  o.zOriginWeightFactor = c.zOriginWeightFactor;
  o.maxSeedsPerSpM = c.maxSeedsPerSpM;
  o.compatSeedLimit = static_cast<std::size_t>(c.compatSeedLimit);
  o.seedWeightIncrement = c.seedWeightIncrement;
  o.numSeedIncrement = c.numSeedIncrement;
  o.seedConfirmation = c.seedConfirmation != 0;
  o.centralSeedConfirmationRange = toRange(c.centralSeedConfirmationRange);
  o.forwardSeedConfirmationRange = toRange(c.forwardSeedConfirmationRange);
  o.maxSeedsPerSpMConf = c.maxSeedsPerSpMConf;
  o.maxQualitySeedsPerSpMConf = c.maxQualitySeedsPerSpMConf;
  o.useDeltaRinsteadOfTopRadius = c.useDeltaRinsteadOfTopRadius != 0;
  o.useExtraCuts = c.useExtraCuts != 0;
  if (c.useVertexZCuts != 0) o.inputVertices = "vertices";
  o.vertexZNSigma = c.vertexZNSigma;
  o.vertexZMargin = c.vertexZMargin;
  return o;
}

// OrthogonalTripletSeedingAlgorithm::Config from the shared b200seed_config fields + the three extra members
ActsExamples::OrthogonalTripletSeedingAlgorithm::Config toOrthogonalConfig(const b200seed_config& c,
                                                                           const b200seed_orthogonal_options& x) {
  ActsExamples::OrthogonalTripletSeedingAlgorithm::Config o;
  o.inputSpacePoints = "spacepoints";
  o.outputSeeds = "seeds";
  o.bFieldInZ = c.bFieldInZ;
  o.minPt = c.minPt;
  o.cotThetaMax = c.cotThetaMax;
  o.impactMax = c.impactMax;
  o.deltaRMin = c.deltaRMin;
  o.deltaRMax = c.deltaRMax;
  o.deltaRMinTop = c.deltaRMinTop;
  o.deltaRMaxTop = c.deltaRMaxTop;
  o.deltaRMinBottom = c.deltaRMinBottom;
  o.deltaRMaxBottom = c.deltaRMaxBottom;
  o.rMin = c.rMin;
  o.rMax = c.rMax;
  o.zMin = c.zMin;
  o.zMax = c.zMax;
  o.phiMin = c.phiMin;
  o.phiMax = c.phiMax;
  o.rMinMiddle = c.rMinMiddle;
  o.rMaxMiddle = c.rMaxMiddle;
  o.useVariableMiddleSPRange = c.useVariableMiddleSPRange != 0;
  o.deltaRMiddleMinSPRange = c.deltaRMiddleMinSPRange;
  o.deltaRMiddleMaxSPRange = c.deltaRMiddleMaxSPRange;
  o.zOutermostLayers = {x.zOutermostLayersMin, x.zOutermostLayersMax};
  o.deltaZMin = c.deltaZMin;
  o.deltaZMax = c.deltaZMax;
  o.deltaPhiMax = x.deltaPhiMax;
  o.interactionPointCut = c.interactionPointCut != 0;
  o.collisionRegionMin = c.collisionRegionMin;
  o.collisionRegionMax = c.collisionRegionMax;
  o.helixCutTolerance = c.helixCutTolerance;
  o.sigmaScattering = c.sigmaScattering;
  o.radLengthPerSeed = c.radLengthPerSeed;
  o.toleranceParam = c.toleranceParam;
  o.deltaInvHelixDiameter = c.deltaInvHelixDiameter;
  o.compatSeedWeight = c.compatSeedWeight;
  o.impactWeightFactor = c.impactWeightFactor;
  o.zOriginWeightFactor = c.zOriginWeightFactor;
  o.maxSeedsPerSpM = c.maxSeedsPerSpM;
  o.compatSeedLimit = static_cast<std::size_t>(c.compatSeedLimit);
  o.seedWeightIncrement = c.seedWeightIncrement;
  o.numSeedIncrement = c.numSeedIncrement;
  o.seedConfirmation = c.seedConfirmation != 0;
  o.centralSeedConfirmationRange = toRange(c.centralSeedConfirmationRange);
  o.forwardSeedConfirmationRange = toRange(c.forwardSeedConfirmationRange);
  o.maxSeedsPerSpMConf = c.maxSeedsPerSpMConf;
  o.maxQualitySeedsPerSpMConf = c.maxQualitySeedsPerSpMConf;
  o.useDeltaRinsteadOfTopRadius = c.useDeltaRinsteadOfTopRadius != 0;
  o.useExtraCuts = c.useExtraCuts != 0;
  return o;
}

struct RefHandle {
  std::unique_ptr<ActsExamples::IAlgorithm> algorithm;
  std::unique_ptr<Harness> harness;
  bool withVertices = false;
};

struct RefResult {
  std::vector<uint32_t> bottom, middle, top;
  std::vector<float> quality, vertexZ;
};

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return B200SEED_OK;
  } catch (const std::invalid_argument& e) {
    g_error = e.what();
    return B200SEED_ERR_INVALID_ARGUMENT;
  } catch (const std::domain_error& e) {
    g_error = e.what();
    return B200SEED_ERR_DOMAIN;
  } catch (const std::exception& e) {
    g_error = e.what();
    return B200SEED_ERR_RUNTIME;
  }
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }

// Constructs the reference algorithm.  cfg->useVertexZCuts != 0 configures `inputVertices`
// (GridTripletSeedingAlgorithm.hpp:239-243): every ref_run then passes (z, sigma_z^2)
// per vertex and the reference builds the windows itself (.cpp:187-206).
int ref_create(const b200seed_config* cfg, void** out) {
  *out = nullptr;
  auto h = std::make_unique<RefHandle>();
  const int rc = guarded([&] {
    h->withVertices = cfg->useVertexZCuts != 0;
    h->algorithm = std::make_unique<Algorithm>(toConfig(*cfg),
                                               Acts::getDefaultLogger("GridTripletSeeding", Acts::Logging::WARNING));
    h->harness = std::make_unique<Harness>("spacepoints", h->withVertices ? "vertices" : "", "seeds");
  });
  if (rc == B200SEED_OK) *out = h.release();
  return rc;
}

// The reference's OrthogonalTripletSeedingAlgorithm (k-d-tree candidate provider); used with ref_run like a grid handle.
int ref_create_orthogonal(const b200seed_config* cfg, const b200seed_orthogonal_options* opt, void** out) {
  *out = nullptr;
  auto h = std::make_unique<RefHandle>();
  const int rc = guarded([&] {
    h->algorithm = std::make_unique<ActsExamples::OrthogonalTripletSeedingAlgorithm>(
        toOrthogonalConfig(*cfg, *opt), Acts::getDefaultLogger("OrthogonalTripletSeeding", Acts::Logging::WARNING));
    h->harness = std::make_unique<Harness>("spacepoints", "", "seeds");
  });
  if (rc == B200SEED_OK) *out = h.release();
  return rc;
}

void ref_destroy(void* h) { delete static_cast<RefHandle*>(h); }

// One event through GridTripletSeedingAlgorithm::execute.  *result owns the seeds.
int ref_run(void* handle, uint32_t n, const float* x, const float* y, const float* z, const float* r, const float* varZ,
            const float* varR, uint32_t nVertices, const double* vertexZ, const double* vertexVarZ, void** result) {
  auto* h = static_cast<RefHandle*>(handle);
  *result = nullptr;
  auto res = std::make_unique<RefResult>();
  const int rc = guarded([&] {
    // the columns SpacePointMaker creates and the seeding reads (SpacePointMaker.cpp:260-264)
    ActsExamples::SpacePointContainer sps(Acts::SpacePointColumns::X | Acts::SpacePointColumns::Y | Acts::SpacePointColumns::Z |
                                          Acts::SpacePointColumns::R | Acts::SpacePointColumns::VarianceZ |
                                          Acts::SpacePointColumns::VarianceR);
    sps.reserve(n);
    for (uint32_t i = 0; i < n; ++i) {
      auto sp = sps.createSpacePoint();
      sp.x() = x[i];
      sp.y() = y[i];
      sp.z() = z[i];
      sp.r() = r[i];
      sp.varianceZ() = varZ[i];
      sp.varianceR() = varR[i];
    }
    ActsExamples::WhiteBoard wb(Acts::getDefaultLogger("WhiteBoard", Acts::Logging::WARNING));
    h->harness->put(wb, std::move(sps));
    if (h->withVertices) {
      ActsExamples::VertexContainer vertices;
      for (uint32_t i = 0; i < nVertices; ++i) vertices.emplace_back(vertexZ[i], vertexVarZ[i]);
      h->harness->put(wb, std::move(vertices));
    }
    ActsExamples::AlgorithmContext ctx(0, 0, wb, 0);
    if (h->algorithm->execute(ctx) != ActsExamples::ProcessCode::SUCCESS) throw std::runtime_error("execute did not return SUCCESS");
    const ActsExamples::SeedContainer& seeds = h->harness->seeds(wb);
    res->bottom.reserve(seeds.size());
    for (const auto& seed : seeds) {
      const auto idx = seed.spacePointIndices();
      if (idx.size() != 3) throw std::runtime_error("seed without three space points");
      res->bottom.push_back(idx[0]);
      res->middle.push_back(idx[1]);
      res->top.push_back(idx[2]);
      res->quality.push_back(seed.quality());
      res->vertexZ.push_back(seed.vertexZ());
    }
  });
  if (rc == B200SEED_OK) *result = res.release();
  return rc;
}

// Timed-baseline entry: nEvents events (concatenated columns, spOffsets[nEvents + 1]) through ONE
// algorithm object from nThreads worker threads, one event per execute() call, every call with its own
// WhiteBoard -- what Sequencer::run does with tbb::parallel_for (Sequencer.cpp:472-525).  seedCounts[e] =
// number of seeds of event e.  Returns the total number of seeds, or -1 on error.
int64_t ref_run_many(void* handle, uint32_t nEvents, const uint32_t* spOffsets, const float* x, const float* y,
                     const float* z, const float* r, const float* varZ, const float* varR, int nThreads,
                     uint64_t* seedCounts) {
  auto* h = static_cast<RefHandle*>(handle);
  if (h->withVertices) {
    g_error = "ref_run_many: handle configured with vertices";
    return -1;
  }
  if (nThreads < 1) nThreads = 1;
  std::atomic<uint32_t> next{0};
  std::atomic<int> failed{0};
  auto worker = [&] {
    for (;;) {
      const uint32_t e = next.fetch_add(1);
      if (e >= nEvents) return;
      try {
        const uint32_t b = spOffsets[e], n = spOffsets[e + 1] - b;
        ActsExamples::SpacePointContainer sps(Acts::SpacePointColumns::X | Acts::SpacePointColumns::Y |
                                              Acts::SpacePointColumns::Z | Acts::SpacePointColumns::R |
                                              Acts::SpacePointColumns::VarianceZ | Acts::SpacePointColumns::VarianceR);
        sps.reserve(n);
        for (uint32_t i = 0; i < n; ++i) {
          auto sp = sps.createSpacePoint();
          sp.x() = x[b + i];
          sp.y() = y[b + i];
          sp.z() = z[b + i];
          sp.r() = r[b + i];
          sp.varianceZ() = varZ[b + i];
          sp.varianceR() = varR[b + i];
        }
        ActsExamples::WhiteBoard wb(Acts::getDefaultLogger("WhiteBoard", Acts::Logging::WARNING));
        h->harness->put(wb, std::move(sps));
        ActsExamples::AlgorithmContext ctx(0, e, wb, 0);
        if (h->algorithm->execute(ctx) != ActsExamples::ProcessCode::SUCCESS) throw std::runtime_error("execute failed");
        seedCounts[e] = h->harness->seeds(wb).size();
      } catch (...) {
        failed.store(1);
        return;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nThreads; ++t) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  if (failed.load() != 0) {
    g_error = "ref_run_many: an event failed";
    return -1;
  }
  int64_t total = 0;
  for (uint32_t e = 0; e < nEvents; ++e) total += static_cast<int64_t>(seedCounts[e]);
  return total;
}

// ---------------------------------------------------------------------------
// Strip triplet path (Core/src/Seeding/TripletSeedFinder.cpp:164-406).
// GridTripletSeedingAlgorithm::execute hard-wires TripletSeedFinder::Config::useStripInfo = false (.cpp:315) and
// copies no strip column into its core container, so the unmodified execute() cannot reach that path.
// ref_run_strips drives the same Core objects in the sequence execute() does -- the algorithm object's own grid /
// filter configuration and TripletSeeder (built by the reference's constructor), CylindricalSpacePointGrid,
// DoubletSeedFinder, BroadTripletSeedFilter, TripletSeeder::createSeedsFromGroups -- but with a TripletSeedFinder
// created with useStripInfo = true (+ cotThetaDiffMax) and the StripCalibrationDetails column filled.
// Everything that computes is the reference's; this function is the glue between its public Core interfaces.
// strip: 12 floats per space point = outerCenter, innerToOuterSeparation, outerHalfVector, innerHalfVector
// (StripSpacePointCalibrationDetails.hpp:16-29).  useStripInfo = 0 runs the pixel path through the same glue: the
// result must then equal ref_run's (tests/test_reference_pin.py checks it).  Plain doublet cuts only (the ITk doublet cut and the vertex
// windows are file-local to the algorithm's source).
int ref_run_strips(void* handle, uint32_t n, const float* x, const float* y, const float* z, const float* r,
                   const float* varZ, const float* varR, const float* strip, float cotThetaDiffMax, int useStripInfo,
                   void** result) {
  auto* h = static_cast<RefHandle*>(handle);
  *result = nullptr;
  auto res = std::make_unique<RefResult>();
  const int rc = guarded([&] {
    const auto* alg = dynamic_cast<const Algorithm*>(h->algorithm.get());
    if (alg == nullptr) throw std::invalid_argument("ref_run_strips: not a grid handle");
    const Algorithm::Config& c = alg->m_cfg;
    if (c.useExtraCuts || !c.inputVertices.empty()) throw std::invalid_argument("ref_run_strips: plain doublet cuts only");

    Acts::CylindricalSpacePointGrid grid(alg->m_gridConfig, Acts::getDefaultLogger("Grid", Acts::Logging::WARNING));
    for (uint32_t i = 0; i < n; ++i) grid.insert(i, std::atan2(y[i], x[i]), z[i], r[i]);
    for (std::size_t b = 0; b < grid.numberOfBins(); ++b) {
      std::ranges::sort(grid.at(b), [&](const Acts::SpacePointIndex& p, const Acts::SpacePointIndex& q) { return r[p] < r[q]; });
    }

    Acts::SpacePointContainer core(Acts::SpacePointColumns::CopiedFromIndex | Acts::SpacePointColumns::PackedXY |
                                   Acts::SpacePointColumns::PackedZR | Acts::SpacePointColumns::VarianceZ |
                                   Acts::SpacePointColumns::VarianceR | Acts::SpacePointColumns::StripCalibrationDetails);
    core.reserve(grid.numberOfSpacePoints());
    std::vector<Acts::SpacePointIndexRange> ranges;
    float rLow = std::numeric_limits<float>::max(), rHigh = std::numeric_limits<float>::lowest();
    for (std::size_t b = 0; b < grid.numberOfBins(); ++b) {
      const std::uint32_t first = core.size();
      for (Acts::SpacePointIndex i : grid.at(b)) {
        auto sp = core.createSpacePoint();
        sp.copiedFromIndex() = i;
        sp.xy() = std::array<float, 2>{x[i], y[i]};
        sp.zr() = std::array<float, 2>{z[i], r[i]};
        sp.varianceZ() = varZ[i];
        sp.varianceR() = varR[i];
        const float* d = strip + 12u * i;
        Acts::OuterStripSpacePointCalibrationDetails det;
        det.outerCenter = {d[0], d[1], d[2]};
        det.innerToOuterSeparation = {d[3], d[4], d[5]};
        det.outerHalfVector = {d[6], d[7], d[8]};
        det.innerHalfVector = {d[9], d[10], d[11]};
        sp.outerStripCalibrationDetails() = det;
      }
      const std::uint32_t last = core.size();
      ranges.emplace_back(first, last);
      if (first != last) {
        rLow = std::min(rLow, core[first].zr()[1]);
        rHigh = std::max(rHigh, core[last - 1].zr()[1]);
      }
    }

    Acts::DoubletSeedFinder::Config dc;
    dc.spacePointsSortedByRadius = true;
    dc.candidateDirection = Acts::Direction::Backward();
    dc.deltaRMin = std::isnan(c.deltaRMinBottom) ? c.deltaRMin : c.deltaRMinBottom;
    dc.deltaRMax = std::isnan(c.deltaRMaxBottom) ? c.deltaRMax : c.deltaRMaxBottom;
    dc.deltaZMin = c.deltaZMin;
    dc.deltaZMax = c.deltaZMax;
    dc.impactMax = c.impactMax;
    dc.interactionPointCut = c.interactionPointCut;
    dc.collisionRegionMin = c.collisionRegionMin;
    dc.collisionRegionMax = c.collisionRegionMax;
    dc.cotThetaMax = c.cotThetaMax;
    dc.minPt = c.minPt;
    dc.helixCutTolerance = c.helixCutTolerance;
    auto bottomFinder = Acts::DoubletSeedFinder::create(Acts::DoubletSeedFinder::DerivedConfig(dc, c.bFieldInZ));
    dc.candidateDirection = Acts::Direction::Forward();
    dc.deltaRMin = std::isnan(c.deltaRMinTop) ? c.deltaRMin : c.deltaRMinTop;
    dc.deltaRMax = std::isnan(c.deltaRMaxTop) ? c.deltaRMax : c.deltaRMaxTop;
    auto topFinder = Acts::DoubletSeedFinder::create(Acts::DoubletSeedFinder::DerivedConfig(dc, c.bFieldInZ));

    Acts::TripletSeedFinder::Config tc;
    // the one switch execute() does not offer (0: the glue itself can be checked -- same seeds as ref_run)
    tc.useStripInfo = useStripInfo != 0;
    tc.sortedByCotTheta = true;
    tc.minPt = c.minPt;
    tc.sigmaScattering = c.sigmaScattering;
    tc.radLengthPerSeed = c.radLengthPerSeed;
    tc.impactMax = c.impactMax;
    tc.helixCutTolerance = c.helixCutTolerance;
    tc.toleranceParam = c.toleranceParam;
    tc.cotThetaDiffMax = cotThetaDiffMax;
    auto tripletFinder = Acts::TripletSeedFinder::create(Acts::TripletSeedFinder::DerivedConfig(tc, c.bFieldInZ));

    const Acts::Range1D<float> variableRange = {std::floor(rLow / 2) * 2 + c.deltaRMiddleMinSPRange,
                                                std::floor(rHigh / 2) * 2 - c.deltaRMiddleMaxSPRange};
    Acts::BroadTripletSeedFilter::State filterState;
    Acts::BroadTripletSeedFilter::Cache filterCache;
    Acts::BroadTripletSeedFilter filter(alg->m_filterConfig, filterState, filterCache, *alg->m_filterLogger);
    Acts::TripletSeeder::Cache cache;
    Acts::SeedContainer seeds;
    std::vector<Acts::SpacePointContainer::ConstRange> below, above;
    for (const auto [bottomBins, middleBin, topBins] : grid.binnedGroup()) {
      const auto middles = core.range(ranges.at(middleBin)).asConst();
      if (middles.empty()) continue;
      below.clear();
      above.clear();
      for (const auto b : bottomBins) below.push_back(core.range(ranges.at(b)).asConst());
      for (const auto t : topBins) above.push_back(core.range(ranges.at(t)).asConst());
      const auto rRangeMiddle = alg->retrieveRadiusRangeForMiddle(middles.front(), variableRange);
      alg->m_seedFinder->createSeedsFromGroups(cache, *bottomFinder, *topFinder, *tripletFinder, filter, core, below, middles,
                                               above, rRangeMiddle, seeds);
    }
    for (const auto& seed : seeds) {
      const auto idx = seed.spacePointIndices();
      if (idx.size() != 3) throw std::runtime_error("seed without three space points");
      res->bottom.push_back(core.at(idx[0]).copiedFromIndex());
      res->middle.push_back(core.at(idx[1]).copiedFromIndex());
      res->top.push_back(core.at(idx[2]).copiedFromIndex());
      res->quality.push_back(seed.quality());
      res->vertexZ.push_back(seed.vertexZ());
    }
  });
  if (rc == B200SEED_OK) *result = res.release();
  return rc;
}

uint64_t ref_result_num_seeds(const void* r) { return static_cast<const RefResult*>(r)->bottom.size(); }

void ref_result_seeds(const void* rv, uint32_t* b, uint32_t* m, uint32_t* t, float* q, float* vz) {
  const auto* r = static_cast<const RefResult*>(rv);
  const std::size_t n = r->bottom.size();
  if (n == 0) return;
  std::memcpy(b, r->bottom.data(), n * sizeof(uint32_t));
  std::memcpy(m, r->middle.data(), n * sizeof(uint32_t));
  std::memcpy(t, r->top.data(), n * sizeof(uint32_t));
  std::memcpy(q, r->quality.data(), n * sizeof(float));
  std::memcpy(vz, r->vertexZ.data(), n * sizeof(float));
}

void ref_result_free(void* r) { delete static_cast<RefResult*>(r); }

}  // extern "C"
