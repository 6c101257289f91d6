// Exercises the C++ host mirror (acts_b200/host/GridTripletSeedingAlgorithm.hpp).
//   host_mirror_main errors              -> exception mapping (no GPU needed beyond plan validation)
//   host_mirror_main run <in.bin> <out.bin>  -> seeds one event read from a raw float file
// File format in: uint32 n, then x[n] y[n] z[n] r[n] varZ[n] varR[n] (float32).
// File format out: uint64 nSeeds, then bottom, middle, top (uint32) and quality, vertexZ (float32).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "../../acts_b200/host/GridTripletSeedingAlgorithm.hpp"

using Alg = ActsB200::GridTripletSeedingAlgorithm;

static Alg::Config pu200() {
  // CI/physmon/workflows/physmon_trackfinding_ttbar_pu200.py:104-114
  Alg::Config c;
  c.rMax = 200; c.deltaRMin = 1; c.deltaRMax = 300; c.deltaRMinTop = 1; c.deltaRMaxTop = 300;
  c.deltaRMinBottom = 1; c.deltaRMaxBottom = 300; c.collisionRegionMin = -250; c.collisionRegionMax = 250;
  c.zMin = -2000; c.zMax = 2000; c.maxSeedsPerSpM = 1; c.sigmaScattering = 5; c.radLengthPerSeed = 0.1f;
  c.minPt = 0.5f; c.impactMax = 3; c.rMin = 33;
  return c;
}

template <typename E>
static bool throwsAs(Alg::Config c) {
  try {
    Alg a(c);
  } catch (const E&) {
    return true;
  } catch (const std::exception& e) {
    std::cerr << "unexpected exception: " << e.what() << "\n";
    return false;
  }
  return false;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::strcmp(argv[1], "errors") == 0) {
    auto a = pu200(); a.minPt = 0.010f;                       // std::domain_error (phi binning)
    auto b = pu200(); b.phiMin = -4.f;                        // std::runtime_error (grid range)
    auto c = pu200(); c.zBinEdges = {-1, 0, 1}; c.zBinsCustomLooping = {1, 3};  // std::invalid_argument
    const bool ok = throwsAs<std::domain_error>(a) && throwsAs<std::runtime_error>(b) && throwsAs<std::invalid_argument>(c);
    std::puts(ok ? "errors ok" : "errors FAILED");
    return ok ? 0 : 1;
  }
  if (argc >= 4 && std::strcmp(argv[1], "run") == 0) {
    std::ifstream in(argv[2], std::ios::binary);
    std::uint32_t n = 0;
    in.read(reinterpret_cast<char*>(&n), 4);
    std::vector<std::vector<float>> col(6, std::vector<float>(n));
    for (auto& v : col) in.read(reinterpret_cast<char*>(v.data()), 4ull * n);
    Alg alg(pu200());
    const auto seeds = alg.execute({col[0], col[1], col[2], col[3], col[4], col[5]});
    std::ofstream out(argv[3], std::ios::binary);
    const std::uint64_t ns = seeds.size();
    out.write(reinterpret_cast<const char*>(&ns), 8);
    out.write(reinterpret_cast<const char*>(seeds.bottom.data()), 4 * ns);
    out.write(reinterpret_cast<const char*>(seeds.middle.data()), 4 * ns);
    out.write(reinterpret_cast<const char*>(seeds.top.data()), 4 * ns);
    out.write(reinterpret_cast<const char*>(seeds.quality.data()), 4 * ns);
    out.write(reinterpret_cast<const char*>(seeds.vertexZ.data()), 4 * ns);
    std::printf("seeds %llu\n", static_cast<unsigned long long>(ns));
    return 0;
  }
  std::puts("usage: host_mirror_main errors | run <in.bin> <out.bin>");
  return 2;
}
