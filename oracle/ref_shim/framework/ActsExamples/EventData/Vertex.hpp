// Stand-in for ActsExamples/EventData/Vertex.hpp (TEST INFRASTRUCTURE, see
// ../../../Eigen/Core).  The reference header pulls Acts/Vertexing/Vertex.hpp and
// with it the track-parameter / surface headers, which need far more of Eigen
// than the seeding path does.  GridTripletSeedingAlgorithm.cpp:187-206 only reads
// `position().z()` and `covariance()(2, 2)` of every vertex, so the container
// element here carries exactly a 3-position and a 3x3 covariance.
#pragma once

#include "Acts/Definitions/Algebra.hpp"

#include <vector>

namespace ActsExamples {

class SeedingVertex {
 public:
  SeedingVertex() = default;
  SeedingVertex(double z, double varZ) {
    m_position[2] = z;
    m_covariance(2, 2) = varZ;
  }
  const Acts::Vector3& position() const { return m_position; }
  const Acts::SquareMatrix3& covariance() const { return m_covariance; }

 private:
  Acts::Vector3 m_position;
  Acts::SquareMatrix3 m_covariance;
};

using VertexContainer = std::vector<SeedingVertex>;

}  // namespace ActsExamples
