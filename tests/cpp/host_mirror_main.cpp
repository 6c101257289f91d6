// Exercises the C++ host mirror (acts_b200/host/GridTripletSeedingAlgorithm.hpp).
//   host_mirror_main errors              -> exception mapping (no GPU needed beyond plan validation)
//   host_mirror_main run <in.bin> <out.bin>  -> seeds one event read from a raw float file
//   host_mirror_main mt <out.bin> <threads> <in0.bin> <in1.bin> ...
//        -> ONE algorithm object, <threads> worker threads calling execute() concurrently (every thread seeds every
//           event twice, in a different order, the way the Sequencer's workers enter execute()); all results of an
//           event must be identical; the seeds of every event are written one after the other
//   host_mirror_main runv <in.bin> <out.bin> <nSigma> <margin> z0 var0 z1 var1 ...  -> Config::inputVertices path
//   host_mirror_main orth <out.bin> <threads> <in0.bin> ...  -> ActsB200::OrthogonalTripletSeedingAlgorithm, one object,
//        <threads> worker threads, every event seeded by every thread; seeds of every event written one after the other
// File format in: uint32 n, then x[n] y[n] z[n] r[n] varZ[n] varR[n] (float32).
// File format out: uint64 nSeeds, then bottom, middle, top (uint32) and quality, vertexZ (float32).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

#include "../../acts_b200/host/GridTripletSeedingAlgorithm.hpp"
#include "../../acts_b200/host/OrthogonalTripletSeedingAlgorithm.hpp"

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

using Columns = std::vector<std::vector<float>>;

static Columns readEvent(const char* path) {
  std::ifstream in(path, std::ios::binary);
  std::uint32_t n = 0;
  in.read(reinterpret_cast<char*>(&n), 4);
  Columns col(6, std::vector<float>(n));
  for (auto& v : col) in.read(reinterpret_cast<char*>(v.data()), 4ull * n);
  return col;
}

static void writeSeeds(std::ofstream& out, const ActsB200::SeedColumns& seeds) {
  const std::uint64_t ns = seeds.size();
  out.write(reinterpret_cast<const char*>(&ns), 8);
  out.write(reinterpret_cast<const char*>(seeds.bottom.data()), 4 * ns);
  out.write(reinterpret_cast<const char*>(seeds.middle.data()), 4 * ns);
  out.write(reinterpret_cast<const char*>(seeds.top.data()), 4 * ns);
  out.write(reinterpret_cast<const char*>(seeds.quality.data()), 4 * ns);
  out.write(reinterpret_cast<const char*>(seeds.vertexZ.data()), 4 * ns);
}

static bool sameSeeds(const ActsB200::SeedColumns& a, const ActsB200::SeedColumns& b) {
  return a.bottom == b.bottom && a.middle == b.middle && a.top == b.top &&
         a.size() == b.size() && std::memcmp(a.quality.data(), b.quality.data(), 4 * a.size()) == 0 &&
         std::memcmp(a.vertexZ.data(), b.vertexZ.data(), 4 * a.size()) == 0;
}

int main(int argc, char** argv) {
  if (argc >= 5 && std::strcmp(argv[1], "mt") == 0) {
    const int nThreads = std::atoi(argv[3]);
    std::vector<Columns> evs;
    for (int i = 4; i < argc; ++i) evs.push_back(readEvent(argv[i]));
    const Alg alg(pu200());  // const: execute() is a const member, entered concurrently
    std::vector<ActsB200::SeedColumns> first(evs.size());
    for (std::size_t e = 0; e < evs.size(); ++e) {
      first[e] = alg.execute({evs[e][0], evs[e][1], evs[e][2], evs[e][3], evs[e][4], evs[e][5]});
    }
    std::atomic<int> bad{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < nThreads; ++t) {
      pool.emplace_back([&, t] {
        try {
          for (int rep = 0; rep < 2; ++rep) {
            for (std::size_t k = 0; k < evs.size(); ++k) {
              const std::size_t e = (k * 7 + static_cast<std::size_t>(t) * 3 + static_cast<std::size_t>(rep)) % evs.size();
              const auto seeds = alg.execute({evs[e][0], evs[e][1], evs[e][2], evs[e][3], evs[e][4], evs[e][5]});
              if (!sameSeeds(seeds, first[e])) ++bad;
            }
          }
        } catch (const std::exception& ex) {
          std::cerr << "thread " << t << ": " << ex.what() << "\n";
          ++bad;
        }
      });
    }
    for (auto& th : pool) th.join();
    std::ofstream out(argv[2], std::ios::binary);
    for (const auto& s : first) writeSeeds(out, s);
    std::printf("threads %d events %zu slots %zu mismatches %d\n", nThreads, evs.size(), alg.slotsInUse(), bad.load());
    return bad.load() == 0 ? 0 : 1;
  }
  if (argc >= 6 && std::strcmp(argv[1], "runv") == 0) {
    const Columns col = readEvent(argv[2]);
    auto cfg = pu200();
    cfg.inputVertices = "vertices";
    cfg.vertexZNSigma = std::atof(argv[4]);
    cfg.vertexZMargin = std::atof(argv[5]);
    std::vector<double> vz, vv;
    for (int i = 6; i + 1 < argc; i += 2) { vz.push_back(std::atof(argv[i])); vv.push_back(std::atof(argv[i + 1])); }
    const Alg alg(cfg);
    const auto seeds = alg.execute({col[0], col[1], col[2], col[3], col[4], col[5]}, vz, vv);
    std::ofstream out(argv[3], std::ios::binary);
    writeSeeds(out, seeds);
    std::printf("seeds %zu\n", seeds.size());
    return 0;
  }
  if (argc >= 5 && std::strcmp(argv[1], "strips") == 0) {  // strips <event> <out> <cotThetaDiffMax>: the event file carries 12 more floats per point
    std::ifstream in(argv[2], std::ios::binary);
    std::uint32_t n = 0;
    in.read(reinterpret_cast<char*>(&n), 4);
    Columns col(6, std::vector<float>(n));
    for (auto& v : col) in.read(reinterpret_cast<char*>(v.data()), 4ull * n);
    std::vector<float> details(12ull * n);
    in.read(reinterpret_cast<char*>(details.data()), 48ull * n);
    const Alg alg(pu200());
    const auto seeds = alg.executeStrips({col[0], col[1], col[2], col[3], col[4], col[5]}, details, static_cast<float>(std::atof(argv[4])));
    std::ofstream out(argv[3], std::ios::binary);
    writeSeeds(out, seeds);
    std::printf("seeds %zu\n", seeds.size());
    return 0;
  }
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
  if (argc >= 5 && std::strcmp(argv[1], "orth") == 0) {
    using Orth = ActsB200::OrthogonalTripletSeedingAlgorithm;
    Orth::Config c;  // the <mu>=200 cut set on the orthogonal defaults (acts_b200/config.py: orthogonal_config)
    c.rMax = 200; c.deltaRMin = 1; c.deltaRMax = 300; c.deltaRMinTop = 1; c.deltaRMaxTop = 300;
    c.deltaRMinBottom = 1; c.deltaRMaxBottom = 300; c.collisionRegionMin = -250; c.collisionRegionMax = 250;
    c.zMin = -2000; c.zMax = 2000; c.maxSeedsPerSpM = 1; c.sigmaScattering = 5; c.radLengthPerSeed = 0.1f;
    c.minPt = 0.5f; c.impactMax = 3;
    c.maxConcurrentEvents = 2;
    bool confRejected = false;
    try {
      Orth::Config bad = c;
      bad.seedConfirmation = true;
      Orth rejected(bad);
    } catch (const std::runtime_error&) {
      confRejected = true;
    }
    const int nThreads = std::atoi(argv[3]);
    std::vector<Columns> evs;
    for (int i = 4; i < argc; ++i) evs.push_back(readEvent(argv[i]));
    const Orth alg(c);
    std::vector<ActsB200::SeedColumns> first(evs.size());
    for (std::size_t e = 0; e < evs.size(); ++e) {
      const auto& col = evs[e];
      first[e] = alg.execute({col[0], col[1], col[2], col[3], col[4], col[5]});
    }
    std::atomic<int> mismatches{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < nThreads; ++t) {
      pool.emplace_back([&, t] {
        for (std::size_t k = 0; k < evs.size(); ++k) {
          const std::size_t e = (k + static_cast<std::size_t>(t)) % evs.size();
          const auto& col = evs[e];
          if (!sameSeeds(alg.execute({col[0], col[1], col[2], col[3], col[4], col[5]}), first[e])) ++mismatches;
        }
      });
    }
    for (auto& th : pool) th.join();
    std::ofstream out(argv[2], std::ios::binary);
    for (const auto& s : first) writeSeeds(out, s);
    std::printf("threads %d events %zu mismatches %d\n", nThreads, evs.size(), mismatches.load());
    return mismatches.load() == 0 ? 0 : 1;
  }
  std::puts("usage: host_mirror_main errors | run <in.bin> <out.bin>");
  return 2;
}
