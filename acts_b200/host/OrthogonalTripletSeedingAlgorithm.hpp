// OrthogonalTripletSeedingAlgorithm.hpp -- C++20 host-side mirror of
// ActsExamples::OrthogonalTripletSeedingAlgorithm over the C ABI of the B200 plugin.
//
// Reference: Examples/Algorithms/TrackFinding/include/ActsExamples/TrackFinding/
// OrthogonalTripletSeedingAlgorithm.hpp:33-186 (Config: same field names, types, defaults, units) and
// src/OrthogonalTripletSeedingAlgorithm.cpp:62-317 (ctor, execute).  execute() takes the six float columns of the
// input SpacePointContainer and returns the columns of the output SeedContainer in the reference's order (middles in
// k-d-tree element order, the increasing-z group of a middle before its decreasing-z group).
//
// execute() is const and re-entrant like the reference's (thread_local caches, .cpp:240-242): every call borrows an
// engine slot (C-ABI handle + page-locked seed buffers) from a pool owned by the object, like
// ActsB200::GridTripletSeedingAlgorithm.  Header only; link against libacts_b200_seeding.so.  No CPU fallback.
#pragma once

#include <condition_variable>
#include <mutex>

#include "GridTripletSeedingAlgorithm.hpp"

namespace ActsB200 {

class OrthogonalTripletSeedingAlgorithm final {
 public:
  struct Config {
    // identical to the reference Config, OrthogonalTripletSeedingAlgorithm.hpp:38-186
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
    float rMinMiddle = 60.f;
    float rMaxMiddle = 120.f;
    bool useVariableMiddleSPRange = false;
    std::vector<std::vector<float>> rRangeMiddleSP;  // (declared by the reference, never read by its execute())
    float deltaRMiddleMinSPRange = 10.f;
    float deltaRMiddleMaxSPRange = 10.f;
    std::pair<float, float> zOutermostLayers{-2700.f, 2700.f};
    float deltaZMin = -std::numeric_limits<float>::infinity();
    float deltaZMax = std::numeric_limits<float>::infinity();
    float deltaPhiMax = 0.085f;
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
    /// engine options (not in the reference)
    unsigned int maxConcurrentEvents = 4;
    int device = 0;
  };

  explicit OrthogonalTripletSeedingAlgorithm(const Config& cfg) : m_cfg(cfg) {
    b200seed_config& c = m_abi;
    b200seed_orthogonal_config_init(&c, &m_opt);
    c.bFieldInZ = cfg.bFieldInZ; c.minPt = cfg.minPt; c.cotThetaMax = cfg.cotThetaMax; c.impactMax = cfg.impactMax;
    c.deltaRMin = cfg.deltaRMin; c.deltaRMax = cfg.deltaRMax;
    c.deltaRMinTop = cfg.deltaRMinTop; c.deltaRMaxTop = cfg.deltaRMaxTop;
    c.deltaRMinBottom = cfg.deltaRMinBottom; c.deltaRMaxBottom = cfg.deltaRMaxBottom;
    c.rMin = cfg.rMin; c.rMax = cfg.rMax; c.zMin = cfg.zMin; c.zMax = cfg.zMax;
    c.phiMin = cfg.phiMin; c.phiMax = cfg.phiMax;
    c.rMinMiddle = cfg.rMinMiddle; c.rMaxMiddle = cfg.rMaxMiddle;
    c.useVariableMiddleSPRange = cfg.useVariableMiddleSPRange;
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
    m_opt.zOutermostLayersMin = cfg.zOutermostLayers.first;
    m_opt.zOutermostLayersMax = cfg.zOutermostLayers.second;
    m_opt.deltaPhiMax = cfg.deltaPhiMax;
    if (cfg.maxConcurrentEvents == 0) throw std::invalid_argument("maxConcurrentEvents must be at least 1");
    m_free.push_back(newSlot());  // validates the configuration (throws like the reference's finder constructors)
    m_created = 1;
  }
  ~OrthogonalTripletSeedingAlgorithm() {
    for (Slot* s : m_free) delete s;
  }
  OrthogonalTripletSeedingAlgorithm(const OrthogonalTripletSeedingAlgorithm&) = delete;
  OrthogonalTripletSeedingAlgorithm& operator=(const OrthogonalTripletSeedingAlgorithm&) = delete;

  /// Run the seeding algorithm on one event (reference: execute(ctx), .cpp:101-317).  Thread-safe and const.
  SeedColumns execute(const SpacePointColumns& sp) const {
    const std::size_t n = sp.x.size();
    if (sp.y.size() != n || sp.z.size() != n || sp.r.size() != n || sp.varianceZ.size() != n || sp.varianceR.size() != n) {
      throw std::invalid_argument("space point columns differ in length");
    }
    Slot* slot = acquire();
    struct Release {
      const OrthogonalTripletSeedingAlgorithm& a;
      Slot* s;
      ~Release() { a.release(s); }
    } guard{*this, slot};
    slot->reserve(std::max<std::size_t>(16, 4 * n));
    for (;;) {
      b200seed_seeds s = slot->columns();
      const int rc = b200seed_run(slot->handle, static_cast<std::uint32_t>(n), sp.x.data(), sp.y.data(), sp.z.data(), sp.r.data(),
                                  sp.varianceZ.data(), sp.varianceR.data(), 0, nullptr, nullptr, &s);
      if (rc == B200SEED_ERR_CAPACITY) { slot->reserve(static_cast<std::size_t>(s.size)); continue; }
      check(rc);
      SeedColumns out;
      const std::size_t k = static_cast<std::size_t>(s.size);
      out.bottom.assign(s.bottom, s.bottom + k);
      out.middle.assign(s.middle, s.middle + k);
      out.top.assign(s.top, s.top + k);
      out.quality.assign(s.quality, s.quality + k);
      out.vertexZ.assign(s.vertexZ, s.vertexZ + k);
      return out;
    }
  }

  const Config& config() const { return m_cfg; }

 private:
  struct Slot {
    b200seed_handle* handle = nullptr;
    void* pinned = nullptr;
    std::size_t capacity = 0;
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
  static void copyRange(const SeedConfirmationRangeConfig& a, b200seed_seed_confirmation_range& b) {
    b.zMinSeedConf = a.zMinSeedConf; b.zMaxSeedConf = a.zMaxSeedConf; b.rMaxSeedConf = a.rMaxSeedConf;
    b.nTopForLargeR = a.nTopForLargeR; b.nTopForSmallR = a.nTopForSmallR;
    b.seedConfMinBottomRadius = a.seedConfMinBottomRadius; b.seedConfMaxZOrigin = a.seedConfMaxZOrigin;
    b.minImpactSeedConf = a.minImpactSeedConf;
  }
  static void check(int rc) {
    if (rc == B200SEED_OK) return;
    const std::string msg = b200seed_last_error();
    switch (rc) {
      case B200SEED_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
      case B200SEED_ERR_DOMAIN: throw std::domain_error(msg);
      default: throw std::runtime_error(msg);
    }
  }
  Slot* newSlot() const {
    auto slot = std::make_unique<Slot>();
    check(b200seed_create_orthogonal(&m_abi, &m_opt, m_cfg.device, &slot->handle));
    return slot.release();
  }
  Slot* acquire() const {
    std::unique_lock<std::mutex> lock(m_mutex);
    for (;;) {
      if (!m_free.empty()) { Slot* s = m_free.back(); m_free.pop_back(); return s; }
      if (m_created < m_cfg.maxConcurrentEvents) {
        ++m_created;
        lock.unlock();
        try { return newSlot(); } catch (...) { lock.lock(); --m_created; m_cv.notify_one(); throw; }
      }
      m_cv.wait(lock);
    }
  }
  void release(Slot* s) const {
    { std::lock_guard<std::mutex> g(m_mutex); m_free.push_back(s); }
    m_cv.notify_one();
  }

  Config m_cfg;
  b200seed_config m_abi{};
  b200seed_orthogonal_options m_opt{};
  // the slot pool: mutable state behind the const interface, guarded by its mutex
  mutable std::mutex m_mutex;
  mutable std::condition_variable m_cv;
  mutable std::vector<Slot*> m_free;
  mutable unsigned int m_created = 0;
};

}  // namespace ActsB200
