// host_plan.hpp -- host-side planning of the B200 seeding engine.
//
// Replaces the constructor chain of the reference algorithm:
//   GridTripletSeedingAlgorithm ctor      GridTripletSeedingAlgorithm.cpp:101-178
//   CylindricalSpacePointGrid ctor        Core/src/Seeding/CylindricalSpacePointGrid.cpp:15-99
//   computeSpacePointGridPhiBins          Core/src/Seeding/detail/SpacePointGridPhiBinning.cpp:20-95
//   DoubletSeedFinder::DerivedConfig      Core/src/Seeding/DoubletSeedFinder.cpp:351-357
//   TripletSeedFinder::DerivedConfig      Core/src/Seeding/TripletSeedFinder.cpp:456-481
//   BinnedGroup / GridBinFinder tables    BinnedGroup.ipp:32-62, GridBinFinder.ipp:42-76
// All derived constants are computed here, on the host, with the reference's
// exact expressions and types, and handed to the device as plain numbers.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/acts_b200_seeding.h"
#include "seed_math.h"

namespace B200SEED_NS {

struct PlanError {
  int code = B200SEED_OK;
  std::string message;
};

struct HostPlan {
  DeviceConfig dev{};
  b200seed_info info{};
  // middle bins in the reference's visiting order (phi outermost, z in
  // zBinsCustomLooping order, r) as global bin indices
  std::vector<uint32_t> navBins;
  // neighbour bins of every navigation entry, flattened, in the reference's
  // order (never-fillable under/overflow bins removed)
  std::vector<uint32_t> botOffsets, botBins, topOffsets, topBins;
  uint32_t maxNeighborBins = 0;
  // per-middle output slots: min(maxSeedsPerSpM + 1, maxSeedsPerSpMConf)
  uint32_t seedsPerMiddle = 0;
  bool relaxedFloat = false;
  // OrthogonalTripletSeedingAlgorithm: no grid, the k-d-tree provider's search options instead
  bool orthogonal = false;
  OrthDeviceConfig orth{};
  // Config::inputVertices / vertexZNSigma / vertexZMargin (GridTripletSeedingAlgorithm.hpp:239-243)
  bool useVertexZCuts = false;
  double vertexZNSigma = 3.0, vertexZMargin = 0.0;
  float toleranceParam = 1.1f;  // TripletSeedFinder::Config::toleranceParam (strip triplet path only)
};

// Validates like the reference (same exception classes mapped to status codes)
// and fills the plan.  Returns false and sets err on failure.
bool make_host_plan(const b200seed_config& cfg, HostPlan& plan, PlanError& err);

// The same for OrthogonalTripletSeedingAlgorithm (ctor .cpp:62-99, the option / finder set-up of execute()
// .cpp:152-225): derived constants of the three finders and the two CylindricalSpacePointKDTree::Options.
bool make_orthogonal_plan(const b200seed_config& cfg, const b200seed_orthogonal_options& opt, HostPlan& plan, PlanError& err);

// Reference defaults (GridTripletSeedingAlgorithm.hpp:34-244).
void config_defaults(b200seed_config& cfg);

}  // namespace B200SEED_NS
