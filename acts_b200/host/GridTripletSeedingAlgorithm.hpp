// GridTripletSeedingAlgorithm.hpp -- C++20 host-side mirror of the reference
// interface over the C ABI of the B200 plugin.
//
// Mirrors ActsExamples::GridTripletSeedingAlgorithm
// (Examples/Algorithms/TrackFinding/include/ActsExamples/TrackFinding/
// GridTripletSeedingAlgorithm.hpp:32-288): the nested Config has the same field
// names, types, defaults and units; the constructor validates like the reference
// constructor chain and throws the same exception classes
// (std::invalid_argument / std::runtime_error / std::domain_error); execute()
// takes the six float columns of the input SpacePointContainer and returns the
// columns of the output SeedContainer (Core/include/Acts/EventData/
// SeedContainer.hpp:208-213) in the reference's order.  Inside ACTS this class
// body is what a Plugins/B200Seeding algorithm would contain, with the column
// spans taken from SpacePointContainer::xColumn() etc..
//
// execute() is const and RE-ENTRANT like the reference's (the Sequencer enters it from several TBB worker
// threads at once, one event each: Examples/Framework/src/Framework/Sequencer.cpp:472-525; the reference keeps
// its per-thread scratch in a thread_local cache, GridTripletSeedingAlgorithm.cpp:337).  Here every call borrows
// an engine slot -- a C-ABI handle (own CUDA stream and device workspaces) plus page-locked seed buffers that
// are reused from call to call -- from a pool owned by the algorithm object; slots are created on demand up to
// Config::maxConcurrentEvents, further callers wait for a free one.
//
// Header only; link against libacts_b200_seeding.so.  No CPU fallback.
#pragma once

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <numbers>
#include <span>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/acts_b200_seeding.h"

namespace ActsB200 {

/// Acts::SeedConfirmationRangeConfig
struct SeedConfirmationRangeConfig {
  float zMinSeedConf = std::numeric_limits<float>::lowest();
  float zMaxSeedConf = std::numeric_limits<float>::max();
  float rMaxSeedConf = std::numeric_limits<float>::max();
  std::size_t nTopForLargeR = 0;
  std::size_t nTopForSmallR = 0;
  float seedConfMinBottomRadius = 60.f;
  float seedConfMaxZOrigin = 150.f;
  float minImpactSeedConf = 1.f;
};

/// Output seed columns (Acts::SeedContainer): indices refer to the caller's
/// space point columns.
struct SeedColumns {
  std::vector<std::uint32_t> bottom, middle, top;
  std::vector<float> quality, vertexZ;
  std::size_t size() const { return quality.size(); }
};

/// Input space point columns (Acts::SpacePointContainer X, Y, Z, R, VarianceZ, VarianceR).
struct SpacePointColumns {
  std::span<const float> x, y, z, r, varianceZ, varianceR;
};

class GridTripletSeedingAlgorithm final {
 public:
  struct Config {
    // identical to the reference Config, GridTripletSeedingAlgorithm.hpp:34-244
    float bFieldInZ = static_cast<float>(2 * 0.000299792458);
    float minPt = 0.4f;
    float cotThetaMax = 10.01788f;
    float impactMax = 20.f;
    float deltaRMin = 5.f;
    float deltaRMax = 270.f;
    float deltaRMinTop = std::numeric_limits<float>::quiet_NaN();
    float deltaRMaxTop = std::numeric_limits<float>::quiet_NaN();
    float deltaRMinBottom = std::numeric_limits<float>::quiet_NaN();
    float deltaRMaxBottom = std::numeric_limits<float>::quiet_NaN();
    float rMin = 0.f;
    float rMax = 600.f;
    float zMin = -2800.f;
    float zMax = 2800.f;
    float phiMin = -std::numbers::pi_v<float>;
    float phiMax = std::numbers::pi_v<float>;
    int phiBinDeflectionCoverage = 1;
    int maxPhiBins = 10000;
    std::vector<std::pair<int, int>> zBinNeighborsTop;
    std::vector<std::pair<int, int>> zBinNeighborsBottom;
    int numPhiNeighbors = 1;
    std::vector<float> zBinEdges;
    std::vector<std::size_t> zBinsCustomLooping;
    float rMinMiddle = 60.f;
    float rMaxMiddle = 120.f;
    bool useVariableMiddleSPRange = false;
    std::vector<std::vector<float>> rRangeMiddleSP;
    float deltaRMiddleMinSPRange = 10.f;
    float deltaRMiddleMaxSPRange = 10.f;
    float deltaZMin = -std::numeric_limits<float>::infinity();
    float deltaZMax = std::numeric_limits<float>::infinity();
    bool interactionPointCut = false;
    float collisionRegionMin = -150.f;
    float collisionRegionMax = +150.f;
    float helixCutTolerance = 1.f;
    float sigmaScattering = 5.f;
    float radLengthPerSeed = 0.05f;
    float toleranceParam = 1.1f;
    float deltaInvHelixDiameter = 0.00003f;
    float compatSeedWeight = 200.f;
    float impactWeightFactor = 1.f;
    float zOriginWeightFactor = 1.f;
    unsigned int maxSeedsPerSpM = 5;
    std::size_t compatSeedLimit = 2;
    float seedWeightIncrement = 0.f;
    float numSeedIncrement = std::numeric_limits<float>::infinity();
    bool seedConfirmation = false;
    SeedConfirmationRangeConfig centralSeedConfirmationRange;
    SeedConfirmationRangeConfig forwardSeedConfirmationRange;
    std::uint32_t maxSeedsPerSpMConf = 5;
    std::uint32_t maxQualitySeedsPerSpMConf = 5;
    bool useDeltaRinsteadOfTopRadius = false;
    bool useExtraCuts = false;
    /// Vertex-z constraint (reference: `inputVertices` names a whiteboard collection; a non-empty name enables the
    /// cut for every event, GridTripletSeedingAlgorithm.hpp:239-243, .cpp:292-297).  Here the vertices are passed
    /// to execute() directly; the name only switches the cut on.
    std::string inputVertices;
    double vertexZNSigma = 3.0;
    double vertexZMargin = 0.0;
    /// engine option: events seeded at the same time by one algorithm object (engine slots in the pool)
    unsigned int maxConcurrentEvents = 4;
    /// engine option: CUDA device ordinal
    int device = 0;
    /// engine option: float fast path (FMA contraction, approximate division); seeds not guaranteed identical
    bool relaxedFloat = false;
  };

  explicit GridTripletSeedingAlgorithm(const Config& cfg) : m_cfg(cfg) {
    b200seed_config& c = m_abi;
    b200seed_config_init(&c);
    c.bFieldInZ = cfg.bFieldInZ; c.minPt = cfg.minPt; c.cotThetaMax = cfg.cotThetaMax; c.impactMax = cfg.impactMax;
    c.deltaRMin = cfg.deltaRMin; c.deltaRMax = cfg.deltaRMax;
    c.deltaRMinTop = cfg.deltaRMinTop; c.deltaRMaxTop = cfg.deltaRMaxTop;
    c.deltaRMinBottom = cfg.deltaRMinBottom; c.deltaRMaxBottom = cfg.deltaRMaxBottom;
    c.rMin = cfg.rMin; c.rMax = cfg.rMax; c.zMin = cfg.zMin; c.zMax = cfg.zMax;
    c.phiMin = cfg.phiMin; c.phiMax = cfg.phiMax;
    c.phiBinDeflectionCoverage = cfg.phiBinDeflectionCoverage; c.maxPhiBins = cfg.maxPhiBins;
    for (const auto& [a, b] : cfg.zBinNeighborsTop) { m_zTop.push_back(a); m_zTop.push_back(b); }
    for (const auto& [a, b] : cfg.zBinNeighborsBottom) { m_zBottom.push_back(a); m_zBottom.push_back(b); }
    c.zBinNeighborsTop = m_zTop.data(); c.nZBinNeighborsTop = static_cast<std::uint32_t>(cfg.zBinNeighborsTop.size());
    c.zBinNeighborsBottom = m_zBottom.data(); c.nZBinNeighborsBottom = static_cast<std::uint32_t>(cfg.zBinNeighborsBottom.size());
    c.numPhiNeighbors = cfg.numPhiNeighbors;
    c.zBinEdges = m_cfg.zBinEdges.data(); c.nZBinEdges = static_cast<std::uint32_t>(m_cfg.zBinEdges.size());
    for (std::size_t v : cfg.zBinsCustomLooping) m_looping.push_back(v);
    c.zBinsCustomLooping = m_looping.data(); c.nZBinsCustomLooping = static_cast<std::uint32_t>(m_looping.size());
    c.rMinMiddle = cfg.rMinMiddle; c.rMaxMiddle = cfg.rMaxMiddle;
    c.useVariableMiddleSPRange = cfg.useVariableMiddleSPRange;
    for (const auto& v : cfg.rRangeMiddleSP) {
      if (v.size() < 2) throw std::invalid_argument("rRangeMiddleSP entries need {rMin, rMax}");
      m_rRange.push_back(v[0]); m_rRange.push_back(v[1]);
    }
    c.rRangeMiddleSP = m_rRange.data(); c.nRRangeMiddleSP = static_cast<std::uint32_t>(cfg.rRangeMiddleSP.size());
    c.deltaRMiddleMinSPRange = cfg.deltaRMiddleMinSPRange; c.deltaRMiddleMaxSPRange = cfg.deltaRMiddleMaxSPRange;
    c.deltaZMin = cfg.deltaZMin; c.deltaZMax = cfg.deltaZMax; c.interactionPointCut = cfg.interactionPointCut;
    c.collisionRegionMin = cfg.collisionRegionMin; c.collisionRegionMax = cfg.collisionRegionMax;
    c.helixCutTolerance = cfg.helixCutTolerance; c.sigmaScattering = cfg.sigmaScattering;
    c.radLengthPerSeed = cfg.radLengthPerSeed; c.toleranceParam = cfg.toleranceParam;
    c.deltaInvHelixDiameter = cfg.deltaInvHelixDiameter; c.compatSeedWeight = cfg.compatSeedWeight;
    c.impactWeightFactor = cfg.impactWeightFactor; c.zOriginWeightFactor = cfg.zOriginWeightFactor;
    c.maxSeedsPerSpM = cfg.maxSeedsPerSpM; c.compatSeedLimit = cfg.compatSeedLimit;
    c.seedWeightIncrement = cfg.seedWeightIncrement; c.numSeedIncrement = cfg.numSeedIncrement;
    c.seedConfirmation = cfg.seedConfirmation;
    copyRange(cfg.centralSeedConfirmationRange, c.centralSeedConfirmationRange);
    copyRange(cfg.forwardSeedConfirmationRange, c.forwardSeedConfirmationRange);
    c.maxSeedsPerSpMConf = cfg.maxSeedsPerSpMConf; c.maxQualitySeedsPerSpMConf = cfg.maxQualitySeedsPerSpMConf;
    c.useDeltaRinsteadOfTopRadius = cfg.useDeltaRinsteadOfTopRadius; c.useExtraCuts = cfg.useExtraCuts;
    c.relaxedFloat = cfg.relaxedFloat;
    c.useVertexZCuts = !cfg.inputVertices.empty();
    c.vertexZNSigma = cfg.vertexZNSigma;
    c.vertexZMargin = cfg.vertexZMargin;
    if (cfg.maxConcurrentEvents == 0) throw std::invalid_argument("maxConcurrentEvents must be at least 1");
    // the reference constructor chain throws here; so does this one (host-side validation, then the first slot)
    b200seed_info info{};
    check(b200seed_plan_info(&c, &info));
    m_pool = std::make_unique<Pool>();
    m_pool->slots.push_back(newSlot());
    m_pool->free.push_back(m_pool->slots.back().get());
  }
  ~GridTripletSeedingAlgorithm() = default;
  GridTripletSeedingAlgorithm(const GridTripletSeedingAlgorithm&) = delete;
  GridTripletSeedingAlgorithm& operator=(const GridTripletSeedingAlgorithm&) = delete;

  /// Run the seeding algorithm on one event (reference: execute(ctx), .cpp:180-402).  Thread-safe and const.
  /// `zWindows` are optional per-event vertex z-windows given directly (VertexZCuts, .cpp:69-97).
  SeedColumns execute(const SpacePointColumns& sp, std::span<const std::pair<float, float>> zWindows = {}) const {
    checkColumns(sp);
    std::vector<float> lo, hi;
    for (const auto& [a, b] : zWindows) { lo.push_back(a); hi.push_back(b); }
    Lease lease(*m_pool, *this);
    return run(*lease.slot, sp, [&](Slot& slot, b200seed_seeds* s) {
      return b200seed_run(slot.handle, static_cast<std::uint32_t>(sp.x.size()), sp.x.data(), sp.y.data(), sp.z.data(), sp.r.data(),
                          sp.varianceZ.data(), sp.varianceR.data(), static_cast<std::uint32_t>(lo.size()), lo.data(), hi.data(), s);
    });
  }

  /// One event with its reconstructed vertices (reference: m_inputVertices(ctx), .cpp:187-206): z position and the
  /// (2, 2) element of the covariance of every vertex.  Needs Config::inputVertices to be set, like the reference.
  SeedColumns execute(const SpacePointColumns& sp, std::span<const double> vertexZ, std::span<const double> vertexVarZ) const {
    checkColumns(sp);
    if (m_cfg.inputVertices.empty()) throw std::invalid_argument("vertices given but Config::inputVertices is empty");
    if (vertexZ.size() != vertexVarZ.size()) throw std::invalid_argument("vertex columns differ in length");
    Lease lease(*m_pool, *this);
    return run(*lease.slot, sp, [&](Slot& slot, b200seed_seeds* s) {
      return b200seed_run_vertices(slot.handle, static_cast<std::uint32_t>(sp.x.size()), sp.x.data(), sp.y.data(), sp.z.data(),
                                   sp.r.data(), sp.varianceZ.data(), sp.varianceR.data(), static_cast<std::uint32_t>(vertexZ.size()),
                                   vertexZ.data(), vertexVarZ.data(), s);
    });
  }

  /// The strip triplet path (Acts::TripletSeedFinder::Config::useStripInfo = true, TripletSeedFinder.cpp:164-406; the
  /// reference algorithm hard-wires false, .cpp:315 -- this is the entry for callers that build their own finder).
  /// `stripDetails`: 12 floats per space point = outerCenter, innerToOuterSeparation, outerHalfVector, innerHalfVector
  /// (Acts::OuterStripSpacePointCalibrationDetails); `cotThetaDiffMax`: TripletSeedFinder::Config::cotThetaDiffMax.
  SeedColumns executeStrips(const SpacePointColumns& sp, std::span<const float> stripDetails,
                            float cotThetaDiffMax = std::numeric_limits<float>::infinity()) const {
    checkColumns(sp);
    if (stripDetails.size() != 12 * sp.x.size()) throw std::invalid_argument("stripDetails: 12 floats per space point");
    Lease lease(*m_pool, *this);
    return run(*lease.slot, sp, [&](Slot& slot, b200seed_seeds* s) {
      return b200seed_run_strips(slot.handle, static_cast<std::uint32_t>(sp.x.size()), sp.x.data(), sp.y.data(), sp.z.data(),
                                 sp.r.data(), sp.varianceZ.data(), sp.varianceR.data(), stripDetails.data(), cotThetaDiffMax, s);
    });
  }

  const Config& config() const { return m_cfg; }
  /// engine slots created so far (<= Config::maxConcurrentEvents)
  std::size_t slotsInUse() const { std::lock_guard<std::mutex> g(m_pool->mutex); return m_pool->slots.size(); }

 private:
  static void copyRange(const SeedConfirmationRangeConfig& a, b200seed_seed_confirmation_range& b) {
    b.zMinSeedConf = a.zMinSeedConf; b.zMaxSeedConf = a.zMaxSeedConf; b.rMaxSeedConf = a.rMaxSeedConf;
    b.nTopForLargeR = a.nTopForLargeR; b.nTopForSmallR = a.nTopForSmallR;
    b.seedConfMinBottomRadius = a.seedConfMinBottomRadius; b.seedConfMaxZOrigin = a.seedConfMaxZOrigin;
    b.minImpactSeedConf = a.minImpactSeedConf;
  }
  /// status code -> the exception class the reference throws
  static void check(int rc) {
    if (rc == B200SEED_OK) return;
    const std::string msg = b200seed_last_error();
    switch (rc) {
      case B200SEED_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
      case B200SEED_ERR_DOMAIN: throw std::domain_error(msg);
      default: throw std::runtime_error(msg);
    }
  }

  /// One engine slot: a handle (stream + device workspaces) and page-locked seed columns, reused from call to call.
  struct Slot {
    b200seed_handle* handle = nullptr;
    void* pinned = nullptr;
    std::size_t capacity = 0;  // seeds
    ~Slot() {
      if (pinned != nullptr) b200seed_free_pinned(pinned);
      if (handle != nullptr) b200seed_destroy(handle);
    }
    void reserve(std::size_t seeds) {
      if (seeds <= capacity) return;
      if (pinned != nullptr) b200seed_free_pinned(pinned);
      capacity = seeds + seeds / 4;
      pinned = b200seed_alloc_pinned(capacity * 20);
      if (pinned == nullptr) { capacity = 0; throw std::runtime_error("page-locked allocation failed"); }
    }
    b200seed_seeds columns() {
      auto* base = static_cast<std::uint32_t*>(pinned);
      return {base, base + capacity, base + 2 * capacity, reinterpret_cast<float*>(base + 3 * capacity),
              reinterpret_cast<float*>(base + 4 * capacity), capacity, 0};
    }
  };
  struct Pool {
    std::mutex mutex;
    std::condition_variable cv;
    std::vector<std::unique_ptr<Slot>> slots;
    std::vector<Slot*> free;
  };
  struct Lease {  // borrows a slot for the duration of one execute()
    Pool& pool;
    Slot* slot = nullptr;
    Lease(Pool& p, const GridTripletSeedingAlgorithm& alg) : pool(p) {
      std::unique_lock<std::mutex> lock(pool.mutex);
      for (;;) {
        if (!pool.free.empty()) { slot = pool.free.back(); pool.free.pop_back(); return; }
        if (pool.slots.size() < alg.m_cfg.maxConcurrentEvents) {
          pool.slots.push_back(nullptr);  // reserve the place, create outside the lock
          const std::size_t at = pool.slots.size() - 1;
          lock.unlock();
          std::unique_ptr<Slot> fresh;
          try { fresh = alg.newSlot(); } catch (...) { lock.lock(); pool.slots.erase(pool.slots.begin() + at); pool.cv.notify_one(); throw; }
          lock.lock();
          slot = fresh.get();
          for (auto& sl : pool.slots) { if (sl == nullptr) { sl = std::move(fresh); break; } }
          return;
        }
        pool.cv.wait(lock);
      }
    }
    ~Lease() {
      { std::lock_guard<std::mutex> g(pool.mutex); pool.free.push_back(slot); }
      pool.cv.notify_one();
    }
  };

  std::unique_ptr<Slot> newSlot() const {
    auto slot = std::make_unique<Slot>();
    check(b200seed_create(&m_abi, m_cfg.device, &slot->handle));
    return slot;
  }

  static void checkColumns(const SpacePointColumns& sp) {
    const std::size_t n = sp.x.size();
    if (sp.y.size() != n || sp.z.size() != n || sp.r.size() != n || sp.varianceZ.size() != n || sp.varianceR.size() != n) {
      throw std::invalid_argument("space point columns differ in length");
    }
  }

  template <typename Call>
  static SeedColumns run(Slot& slot, const SpacePointColumns& sp, Call&& call) {
    slot.reserve(std::max<std::size_t>(16, 2 * sp.x.size()));
    for (;;) {
      b200seed_seeds s = slot.columns();
      const int rc = call(slot, &s);
      if (rc == B200SEED_ERR_CAPACITY) { slot.reserve(static_cast<std::size_t>(s.size)); continue; }
      check(rc);
      SeedColumns out;
      const std::size_t n = static_cast<std::size_t>(s.size);
      out.bottom.assign(s.bottom, s.bottom + n);
      out.middle.assign(s.middle, s.middle + n);
      out.top.assign(s.top, s.top + n);
      out.quality.assign(s.quality, s.quality + n);
      out.vertexZ.assign(s.vertexZ, s.vertexZ + n);
      return out;
    }
  }

  Config m_cfg;
  b200seed_config m_abi{};  // points into the vectors below
  std::vector<std::int32_t> m_zTop, m_zBottom;
  std::vector<std::uint64_t> m_looping;
  std::vector<float> m_rRange;
  std::unique_ptr<Pool> m_pool;  // mutable state behind the const interface, guarded by its mutex
};

}  // namespace ActsB200
