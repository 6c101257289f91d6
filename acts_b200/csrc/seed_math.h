// seed_math.h -- exact-arithmetic building blocks of the B200 seeding engine.
//
// Everything in this header is `host + device`: the CUDA kernels
// (seeding_kernels.cu) use it on the GPU and tests/model/ compiles the very same
// functions with g++ to check them against the CPU oracle without a GPU.
//
// Exactness contract: the reference evaluates every cut
// in IEEE-754 binary32, round-to-nearest, one rounding per operation, no FMA.
// All float arithmetic below therefore goes through fmul/fadd/fsub/fdiv/fsqrt,
// which map to __fmul_rn & co. on the device (never contracted into FMA by
// ptxas, IEEE-correct division and square root) and to plain operators on the
// host (compiled with -ffp-contract=off, no -march).
//
// Reference lines are cited per function (paths relative to the ACTS tree).
#pragma once

#include <stdint.h>

// Two engines are compiled from the same sources into one library (build.py):
// the exact one (namespace b200seed) and the relaxedFloat fast path
// (-DB200SEED_RELAXED -DB200SEED_NS=b200seed_rx, FMA contraction and
// approximate division / square root allowed).
#ifndef B200SEED_NS
#define B200SEED_NS b200seed
#endif

#if defined(__CUDACC__)
#define B2S_HD __host__ __device__ __forceinline__
#else
#define B2S_HD inline
#endif

#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cstring>
#endif

namespace B200SEED_NS {

// ---------------------------------------------------------------------------
// IEEE binary32 primitives
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__) && defined(B200SEED_RELAXED)
// relaxedFloat: plain operators, the compiler may contract them into FMAs and
// use approximate reciprocals (-fmad=true -prec-div=false -prec-sqrt=false -ftz=true)
B2S_HD float fmul(float a, float b) { return a * b; }
B2S_HD float fadd(float a, float b) { return a + b; }
B2S_HD float fsub(float a, float b) { return a - b; }
B2S_HD float fdiv(float a, float b) { return a / b; }
B2S_HD float fsqrt(float a) { return sqrtf(a); }
B2S_HD double dadd(double a, double b) { return a + b; }
B2S_HD double dsub(double a, double b) { return a - b; }
B2S_HD double dmul(double a, double b) { return a * b; }
B2S_HD double ddiv(double a, double b) { return a / b; }
B2S_HD uint32_t f2u(float f) { return __float_as_uint(f); }
B2S_HD float u2f(uint32_t u) { return __uint_as_float(u); }
B2S_HD float fabs_(float a) { return fabsf(a); }
B2S_HD double dfloor(double a) { return floor(a); }
#elif defined(__CUDA_ARCH__)
B2S_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
B2S_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
B2S_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
B2S_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
B2S_HD float fsqrt(float a) { return __fsqrt_rn(a); }
B2S_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
B2S_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
B2S_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
B2S_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
B2S_HD uint32_t f2u(float f) { return __float_as_uint(f); }
B2S_HD float u2f(uint32_t u) { return __uint_as_float(u); }
B2S_HD float fabs_(float a) { return fabsf(a); }
B2S_HD double dfloor(double a) { return floor(a); }
#else
B2S_HD float fmul(float a, float b) { return a * b; }
B2S_HD float fadd(float a, float b) { return a + b; }
B2S_HD float fsub(float a, float b) { return a - b; }
B2S_HD float fdiv(float a, float b) { return a / b; }
B2S_HD float fsqrt(float a) { return std::sqrt(a); }
B2S_HD double dadd(double a, double b) { return a + b; }
B2S_HD double dsub(double a, double b) { return a - b; }
B2S_HD double dmul(double a, double b) { return a * b; }
B2S_HD double ddiv(double a, double b) { return a / b; }
B2S_HD uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
B2S_HD float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
B2S_HD float fabs_(float a) { return std::fabs(a); }
B2S_HD double dfloor(double a) { return std::floor(a); }
#endif

// ---------------------------------------------------------------------------
// atan2f exactly as glibc 2.39 (x86-64, generic fdlibm-derived
// sysdeps/ieee754/flt-32/e_atan2f.c + s_atanf.c) computes it.  The reference
// bins space points with std::atan2(float, float)
// (GridTripletSeedingAlgorithm.cpp:219) and glibc's result is NOT correctly
// rounded, so the device must replay the same operation sequence.  Validated
// bit-for-bit against the host libm in tests/ (2e8 samples, zero mismatches).
// ---------------------------------------------------------------------------
B2S_HD float glibc_atanf(float x) {
  // atanhi[] / atanlo[] of s_atanf.c, selected by `id` below
  const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f,
              hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
  const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f,
              lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f,
              aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
              aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f,
              aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
              aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f,
              aT10 = 1.6285819933e-02f;
  const int32_t hx = (int32_t)f2u(x);
  const int32_t ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return fadd(x, x);
    if (hx > 0) return fadd(hi3, lo3);
    return fsub(-hi3, lo3);
  }
  if (ix < 0x3ee00000) {  // |x| < 0.4375
    if (ix < 0x31000000) return x;  // |x| < 2^-29
    id = -1;
  } else {
    x = fabs_(x);
    if (ix < 0x3f980000) {    // |x| < 1.1875
      if (ix < 0x3f300000) {  // 7/16 <= |x| < 11/16
        id = 0;
        x = fdiv(fsub(fmul(2.0f, x), 1.0f), fadd(2.0f, x));
      } else {  // 11/16 <= |x| < 19/16
        id = 1;
        x = fdiv(fsub(x, 1.0f), fadd(x, 1.0f));
      }
    } else {
      if (ix < 0x401c0000) {  // |x| < 2.4375
        id = 2;
        x = fdiv(fsub(x, 1.5f), fadd(1.0f, fmul(1.5f, x)));
      } else {  // 2.4375 <= |x| < 2^25
        id = 3;
        x = fdiv(-1.0f, x);
      }
    }
  }
  const float z = fmul(x, x);
  const float w = fmul(z, z);
  float s1 = fadd(aT8, fmul(w, aT10));
  s1 = fadd(aT6, fmul(w, s1));
  s1 = fadd(aT4, fmul(w, s1));
  s1 = fadd(aT2, fmul(w, s1));
  s1 = fadd(aT0, fmul(w, s1));
  s1 = fmul(z, s1);
  float s2 = fadd(aT7, fmul(w, aT9));
  s2 = fadd(aT5, fmul(w, s2));
  s2 = fadd(aT3, fmul(w, s2));
  s2 = fadd(aT1, fmul(w, s2));
  s2 = fmul(w, s2);
  if (id < 0) return fsub(x, fmul(x, fadd(s1, s2)));
  const float ahi = id == 0 ? hi0 : (id == 1 ? hi1 : (id == 2 ? hi2 : hi3));
  const float alo = id == 0 ? lo0 : (id == 1 ? lo1 : (id == 2 ? lo2 : lo3));
  const float r = fsub(ahi, fsub(fsub(fmul(x, fadd(s1, s2)), alo), x));
  return (hx < 0) ? -r : r;
}

B2S_HD float glibc_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f,
              pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  const int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return fadd(x, y);
  if (hx == 0x3f800000) return glibc_atanf(y);
  const int32_t m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0:
      case 1: return y;
      case 2: return fadd(pi, tiny);
      default: return fsub(-pi, tiny);
    }
  }
  if (ix == 0) return (hy < 0) ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return fadd(pi_o_4, tiny);
        case 1: return fsub(-pi_o_4, tiny);
        case 2: return fadd(fmul(3.0f, pi_o_4), tiny);
        default: return fsub(fmul(-3.0f, pi_o_4), tiny);
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return fadd(pi, tiny);
        default: return fsub(-pi, tiny);
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
  const int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60) {
    z = fadd(pi_o_2, fmul(0.5f, pi_lo));
  } else if (hx < 0 && k < -60) {
    z = 0.0f;
  } else {
    z = glibc_atanf(fabs_(fdiv(y, x)));
  }
  switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return fsub(pi, fsub(z, pi_lo));
    default: return fsub(fsub(z, pi_lo), pi);
  }
}

// ---------------------------------------------------------------------------
// Device-side constants (derived on the host with the reference's expressions)
// ---------------------------------------------------------------------------
constexpr int kMaxZEdges = 64;        // z bins + 1 supported on the device
constexpr int kMaxNeighborBins = 64;  // non-empty-able neighbour bins per side per middle bin
constexpr int kMaxCompatSeedLimit = 8;
constexpr int kMaxHeap = 16;          // seeds a middle can return: min(maxSeedsPerSpMConf, maxSeedsPerSpM + 1) (one lane each)
constexpr int kMaxHeapBig = 128;      // collector capacities supported (maxSeedsPerSpMConf, maxQualitySeedsPerSpMConf; itk.py:504-505 uses 100)

enum DoubletCutKind : int { kCutsNone = 0, kCutsItk = 1, kCutsVertexZ = 2 };

// SeedConfirmationRangeConfig (SeedConfirmationRangeConfig.hpp) with the counts narrowed to 32 bits
struct ConfRange {
  float zMinSeedConf, zMaxSeedConf, rMaxSeedConf;
  uint32_t nTopForLargeR, nTopForSmallR;
  float seedConfMinBottomRadius, seedConfMaxZOrigin, minImpactSeedConf;
};

struct DeviceConfig {
  // grid axes (Axis.hpp, doubles like the reference)
  double phiMin, phiMax, phiWidth;
  double zEdges[kMaxZEdges];
  double rAxisMin, rAxisMax;
  int32_t phiBins, nZ, nR, nGlobalBins;
  // space point selector + doublet experiment cuts
  int32_t useExtraCuts;
  int32_t doubletCuts;  // DoubletCutKind
  // doublet finder (DoubletSeedFinder::DerivedConfig)
  float dRMinB, dRMaxB, dRMinT, dRMaxT;
  float deltaZMin, deltaZMax;
  float collisionRegionMin, collisionRegionMax;
  float cotThetaMax, impactMax;
  float minHelixDiameter2Doublet;
  int32_t interactionPointCut;
  // triplet finder (TripletSeedFinder::DerivedConfig)
  float minHelixDiameter2, sigmapT2perRadius, multipleScattering2;
  // filter (BroadTripletSeedFilter::Config)
  float deltaInvHelixDiameter, filterDeltaRMin, compatSeedWeight,
      impactWeightFactor, seedWeightIncrement, numSeedIncrement;
  uint32_t compatSeedLimit;  // clamped to kMaxCompatSeedLimit (checked on host)
  uint32_t maxSeedsPerSpM;
  uint32_t maxSeedsPerSpMConf;  // heap capacity nLow
  int32_t useDeltaRinsteadOfTopRadius;
  // seed confirmation (BroadTripletSeedFilter.cpp:63-94,107-136,253-301)
  int32_t seedConfirmation;
  uint32_t maxQualitySeedsPerSpMConf;  // heap capacity nHigh
  float zOriginWeightFactor;
  ConfRange confCentral, confForward;
  // middle r range
  int32_t useVariableMiddleSPRange;
  float rMinMiddle, rMaxMiddle;
  float deltaRMiddleMinSPRange, deltaRMiddleMaxSPRange;
  int32_t nRRangeMiddleSP;  // 0 = use (rMinMiddle, rMaxMiddle)
  float rRangeMiddleSP[2 * kMaxZEdges];
  int32_t nZBinEdgesF;  // user zBinEdges as float (lower_bound in .cpp:415)
  float zBinEdgesF[kMaxZEdges];
};

// CylindricalSpacePointKDTree::Options of the orthogonal seeder (CylindricalSpacePointKDTree.hpp:45-70,
// OrthogonalTripletSeedingAlgorithm.cpp:152-171): the "low-high" set (top search) and the "high-low" set (bottom
// search) differ only in the deltaR window; + the middle selection of .cpp:268-275.
struct OrthDeviceConfig {
  float rMax, zMin, zMax, phiMin, phiMax;
  float lhDeltaRMin, lhDeltaRMax, hlDeltaRMin, hlDeltaRMax;
  float collisionRegionMin, collisionRegionMax, cotThetaMax, deltaPhiMax, deltaZMax;
  float zOutermostLayersMin, zOutermostLayersMax;
};

// ---------------------------------------------------------------------------
// Grid lookup: SpacePointGridBase.hpp:67-87, Axis.hpp:216-235,296,497-539,598-600,
// MultiAxisHelper.hpp:230-245.  Returns -1 when the point is outside the grid.
// ---------------------------------------------------------------------------
B2S_HD int32_t grid_bin_index(const DeviceConfig& c, float phiF, float zF, float rF) {
  const double phi = (double)phiF, z = (double)zF, r = (double)rF;
  if (!((c.phiMin <= phi) && (phi < c.phiMax))) return -1;
  if (!((c.zEdges[0] <= z) && (z < c.zEdges[c.nZ]))) return -1;
  if (!((c.rAxisMin <= r) && (r < c.rAxisMax))) return -1;
  const int raw = (int)(dadd(dfloor(ddiv(dsub(phi, c.phiMin), c.phiWidth)), 1.0));
  const int w = c.phiBins;
  const int phiBin = 1 + (w + ((raw - 1) % w)) % w;
  // std::upper_bound over the edges: first edge > z
  int lo = 0, hi = c.nZ + 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (c.zEdges[mid] > z) hi = mid; else lo = mid + 1;
  }
  int zBin = lo;
  if (zBin > c.nZ + 1) zBin = c.nZ + 1;
  // r axis has the two edges {rAxisMin, rAxisMax}: inside => bin 1
  const int rBin = 1;
  return (phiBin * (c.nZ + 2) + zBin) * (c.nR + 2) + rBin;
}

// itkFastTrackingSPselect, GridTripletSeedingAlgorithm.cpp:46-62
B2S_HD bool itk_sp_select(float r, float z) {
  const float zabs = fabs_(z);
  if ((double)zabs > 200. && (double)r < 45.) return false;
  const float cotTheta = 27.2899f;
  if (dsub((double)zabs, 150.) > (double)fmul(cotTheta, r)) return false;
  return true;
}

// ---------------------------------------------------------------------------
// Doublets: DoubletSeedFinder.cpp:41-273,359-368
// ---------------------------------------------------------------------------
struct MiddleSp {
  float x, y, z, r, varZ, varR;
  float uIP, uIP2, cosPhiM, sinPhiM;
};

B2S_HD void middle_info(MiddleSp& m) {  // :359-368
  m.uIP = fdiv(-1.0f, m.r);
  m.cosPhiM = fmul(-m.x, m.uIP);
  m.sinPhiM = fmul(-m.y, m.uIP);
  m.uIP2 = fmul(m.uIP, m.uIP);
}

B2S_HD bool outside_range(float v, float lo, float hi) { return (v < lo) | (v > hi); }

// Cuts that only need (z, r) of the other space point (:104-165 for
// interactionPointCut == false; the cotTheta cut moves after the IP cut when
// it is true).  deltaR is assumed to be inside [deltaRMin, deltaRMax] already
// (the r window is found by binary search).
// (`kBottom`: the side as a value -- a compile-time constant in the grid kernels, a per-lane bool in the
// orthogonal seeder's walkers, where the lanes of a warp work on different sides: operands are selected, the
// operations are the same)
B2S_HD bool doublet_zr_cuts_side(const bool kBottom, const DeviceConfig& c, const MiddleSp& m, float zO, float rO,
                                 float& deltaR, float& deltaZ) {
  deltaR = fsub(kBottom ? m.r : rO, kBottom ? rO : m.r);
  deltaZ = fsub(kBottom ? m.z : zO, kBottom ? zO : m.z);
  if (outside_range(deltaZ, c.deltaZMin, c.deltaZMax)) return false;
  const float zOriginTimesDeltaR = fsub(fmul(m.z, deltaR), fmul(m.r, deltaZ));
  if (outside_range(zOriginTimesDeltaR, fmul(c.collisionRegionMin, deltaR),
                    fmul(c.collisionRegionMax, deltaR))) {
    return false;
  }
  if (!c.interactionPointCut) {
    if (outside_range(deltaZ, fmul(-c.cotThetaMax, deltaR), fmul(c.cotThetaMax, deltaR))) {
      return false;
    }
  }
  return true;
}
template <bool kBottom>
B2S_HD bool doublet_zr_cuts(const DeviceConfig& c, const MiddleSp& m, float zO, float rO,
                            float& deltaR, float& deltaZ) {
  return doublet_zr_cuts_side(kBottom, c, m, zO, rO, deltaR, deltaZ);
}

struct DoubletRec {
  float cotTheta, iDeltaR, er, u, v, xNew, yNew;
};

// Second half (:168-201 resp. :205-271): coordinate transform, optional IP /
// curvature cut, experiment cuts, error term.
B2S_HD bool doublet_finish_side(const bool kBottom, const DeviceConfig& c, const MiddleSp& m, float deltaR, float deltaZ,
                                float xO, float yO, float rO, float varZO, float varRO,
                                const float* zWinLo, const float* zWinHi, int nZWin,
                                DoubletRec& out, bool sortedWindows = false) {
  const float deltaX = fsub(xO, m.x);
  const float deltaY = fsub(yO, m.y);
  const float xNewFrame = fadd(fmul(deltaX, m.cosPhiM), fmul(deltaY, m.sinPhiM));
  const float yNewFrame = fsub(fmul(deltaY, m.cosPhiM), fmul(deltaX, m.sinPhiM));
  const float deltaR2 = fadd(fmul(deltaX, deltaX), fmul(deltaY, deltaY));
  const float iDeltaR2 = fdiv(1.0f, deltaR2);
  const float uT = fmul(xNewFrame, iDeltaR2);
  const float vT = fmul(yNewFrame, iDeltaR2);
  if (c.interactionPointCut) {
    const float impactMax = kBottom ? -c.impactMax : c.impactMax;
    const float vIPAbs = fmul(impactMax, m.uIP2);
    if (fabs_(fmul(m.r, yNewFrame)) > fmul(impactMax, xNewFrame)) {
      const float vIP = (yNewFrame > 0) ? -vIPAbs : vIPAbs;
      const float aCoef = fdiv(fsub(vT, vIP), fsub(uT, m.uIP));
      const float bCoef = fsub(vIP, fmul(aCoef, m.uIP));
      if (fmul(fmul(bCoef, bCoef), c.minHelixDiameter2Doublet) >
          fadd(1.0f, fmul(aCoef, aCoef))) {
        return false;
      }
    }
    if (outside_range(deltaZ, fmul(-c.cotThetaMax, deltaR), fmul(c.cotThetaMax, deltaR))) {
      return false;
    }
  }
  const float iDeltaR = fsqrt(iDeltaR2);
  const float cotTheta = fmul(deltaZ, iDeltaR);
  if (c.doubletCuts == kCutsItk) {  // itkFastTrackingCuts, .cpp:33-44
    if (kBottom && rO < 45.0f && (cotTheta > 1.5f || cotTheta < -1.5f)) return false;
  } else if (c.doubletCuts == kCutsVertexZ && nZWin > 0) {  // VertexZCuts, .cpp:78-96
    const float zOrigin = fsub(m.z, fmul(m.r, cotTheta));
    bool inside = false;
    if (sortedWindows) {
      // the device gets the windows merged into disjoint intervals in ascending order (same union, so the same
      // "inside any window" answer): the first interval that ends at or after zOrigin decides
      int lo = 0, hi = nZWin;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (zWinHi[mid] < zOrigin) lo = mid + 1; else hi = mid;
      }
      inside = lo < nZWin && zOrigin >= zWinLo[lo];
    } else {
      for (int k = 0; k < nZWin; ++k) {
        if (zOrigin >= zWinLo[k] && zOrigin <= zWinHi[k]) { inside = true; break; }
      }
    }
    if (!inside) return false;
  }
  // calculateError :67-71
  out.er = fmul(iDeltaR2, fadd(fadd(m.varZ, varZO),
                               fmul(fmul(cotTheta, cotTheta), fadd(m.varR, varRO))));
  out.cotTheta = cotTheta;
  out.iDeltaR = iDeltaR;
  out.u = uT;
  out.v = vT;
  out.xNew = xNewFrame;
  out.yNew = yNewFrame;
  return true;
}
template <bool kBottom>
B2S_HD bool doublet_finish(const DeviceConfig& c, const MiddleSp& m, float deltaR, float deltaZ,
                           float xO, float yO, float rO, float varZO, float varRO,
                           const float* zWinLo, const float* zWinHi, int nZWin,
                           DoubletRec& out, bool sortedWindows = false) {
  return doublet_finish_side(kBottom, c, m, deltaR, deltaZ, xO, yO, rO, varZO, varRO, zWinLo, zWinHi, nZWin, out, sortedWindows);
}

// ---------------------------------------------------------------------------
// Orthogonal seeder: the search boxes of one middle space point
// (CylindricalSpacePointKDTree.cpp:17-217).  RangeXD<3, float> is semi-open per dimension
// (phi, r, z); shrinkMin / shrinkMax are std::max / std::min with the running bound first.
// ---------------------------------------------------------------------------
struct KdBox {
  float mn[3], mx[3];  // phi, r, z
};
B2S_HD float std_max(float a, float b) { return (a < b) ? b : a; }  // std::max(a, b)
B2S_HD float std_min(float a, float b) { return (b < a) ? b : a; }  // std::min(a, b)
B2S_HD void kd_shrink_min(KdBox& b, int d, float v) { b.mn[d] = std_max(b.mn[d], v); }
B2S_HD void kd_shrink_max(KdBox& b, int d, float v) { b.mx[d] = std_min(b.mx[d], v); }
B2S_HD bool kd_degenerate(const KdBox& b) { return (b.mn[0] >= b.mx[0]) | (b.mn[1] >= b.mx[1]) | (b.mn[2] >= b.mx[2]); }
B2S_HD void kd_box_init(KdBox& b) {
  for (int d = 0; d < 3; ++d) { b.mn[d] = -3.402823466e+38f; b.mx[d] = 3.402823466e+38f; }
}

// validTupleOrthoRangeLH, .cpp:17-96 (the box of the TOP candidates, "low-high" options)
B2S_HD void kd_range_lh(const OrthDeviceConfig& o, float pL, float rL, float zL, KdBox& res) {
  const float colMin = o.collisionRegionMin, colMax = o.collisionRegionMax;
  kd_box_init(res);
  kd_shrink_min(res, 0, o.phiMin);
  kd_shrink_max(res, 0, o.phiMax);
  kd_shrink_max(res, 1, o.rMax);
  kd_shrink_min(res, 2, o.zMin);
  kd_shrink_max(res, 2, o.zMax);
  kd_shrink_min(res, 1, fadd(rL, o.lhDeltaRMin));
  kd_shrink_max(res, 1, fadd(rL, o.lhDeltaRMax));
  const float zMax = fadd(fmul(fdiv(res.mx[1], rL), fsub(zL, colMin)), colMin);
  const float zMin = fsub(colMax, fmul(fdiv(res.mx[1], rL), fsub(colMax, zL)));
  if (zL > colMin) {
    kd_shrink_max(res, 2, zMax);
  } else if (zL < colMax) {
    kd_shrink_min(res, 2, zMin);
  }
  kd_shrink_min(res, 2, fsub(zL, fmul(o.cotThetaMax, fsub(res.mx[1], rL))));
  kd_shrink_max(res, 2, fadd(zL, fmul(o.cotThetaMax, fsub(res.mx[1], rL))));
  kd_shrink_min(res, 0, fsub(pL, o.deltaPhiMax));
  kd_shrink_max(res, 0, fadd(pL, o.deltaPhiMax));
  kd_shrink_min(res, 2, fsub(zL, o.deltaZMax));
  kd_shrink_max(res, 2, fadd(zL, o.deltaZMax));
}

// validTupleOrthoRangeHL, .cpp:98-157 (the box of the BOTTOM candidates, "high-low" options)
B2S_HD void kd_range_hl(const OrthDeviceConfig& o, float pM, float rM, float zM, KdBox& res) {
  kd_box_init(res);
  kd_shrink_min(res, 0, o.phiMin);
  kd_shrink_max(res, 0, o.phiMax);
  kd_shrink_max(res, 1, o.rMax);
  kd_shrink_min(res, 2, o.zMin);
  kd_shrink_max(res, 2, o.zMax);
  kd_shrink_min(res, 1, fsub(rM, o.hlDeltaRMax));
  kd_shrink_max(res, 1, fsub(rM, o.hlDeltaRMin));
  const float fracR = fdiv(res.mn[1], rM);
  const float zMin = fadd(fmul(fsub(zM, o.collisionRegionMin), fracR), o.collisionRegionMin);
  const float zMax = fadd(fmul(fsub(zM, o.collisionRegionMax), fracR), o.collisionRegionMax);
  kd_shrink_min(res, 2, std_min(zMin, zM));
  kd_shrink_max(res, 2, std_max(zMax, zM));
  kd_shrink_min(res, 0, fsub(pM, o.deltaPhiMax));
  kd_shrink_max(res, 0, fadd(pM, o.deltaPhiMax));
  kd_shrink_min(res, 2, fsub(zM, o.deltaZMax));
  kd_shrink_max(res, 2, fadd(zM, o.deltaZMax));
}

// validTuples, .cpp:159-217: dir 0 = monotonically increasing z ("lh" lists), dir 1 = decreasing z ("hl" lists)
B2S_HD void kd_search_boxes(const OrthDeviceConfig& o, float pM, float rM, float zM, int dir, KdBox& bottom, KdBox& top) {
  kd_range_hl(o, pM, rM, zM, bottom);
  kd_range_lh(o, pM, rM, zM, top);
  const float cotTheta = std_max(fabs_(fdiv(zM, rM)), o.cotThetaMax);
  const float deltaRMaxTop = fsub(top.mx[1], rM);
  const float deltaRMaxBottom = fsub(rM, bottom.mn[1]);
  if (dir == 0) {
    kd_shrink_min(bottom, 2, fsub(zM, fmul(cotTheta, deltaRMaxBottom)));
    kd_shrink_max(bottom, 2, zM);
    kd_shrink_min(top, 2, zM);
    kd_shrink_max(top, 2, fadd(zM, fmul(cotTheta, deltaRMaxTop)));
  } else {
    kd_shrink_min(bottom, 2, zM);
    kd_shrink_max(bottom, 2, fadd(zM, fmul(cotTheta, deltaRMaxBottom)));
    kd_shrink_min(top, 2, fsub(zM, fmul(cotTheta, deltaRMaxTop)));
    kd_shrink_max(top, 2, zM);
  }
}

// ---------------------------------------------------------------------------
// Triplets: TripletSeedFinder.cpp:34-162 (pixel path)
// ---------------------------------------------------------------------------
struct BottomCtx {
  float cotThetaB, erB, iDeltaRB, Ub, Vb;
  float sigmaSquaredPtDependent, scatteringInRegion2;
};
B2S_HD void bottom_ctx(const DeviceConfig& c, BottomCtx& b) {  // :54-65
  const float iSinTheta2 = fadd(1.0f, fmul(b.cotThetaB, b.cotThetaB));
  b.sigmaSquaredPtDependent = fmul(iSinTheta2, c.sigmapT2perRadius);
  b.scatteringInRegion2 = fmul(c.multipleScattering2, iSinTheta2);
}

enum PairClass : int {
  kPairFailA = 0,  // slope cut with the min-pT scattering term  (:96-107)
  kPairFailB = 1,  // slope cut with the measured-pT term        (:135-143)
  kPairSkip = 2,   // dU == 0, helix diameter or impact cut      (:109-125,148-151)
  kPairEmit = 3    // candidate {top, curvature, impact}         (:155)
};

template <bool kWithCurvature>
B2S_HD int eval_pair_t(const DeviceConfig& c, float rM, float varZM, float varRM,
                       const BottomCtx& b, float cotThetaT, float erT, float iDeltaRT,
                       float uT, const float* vTPtr, float& curvature, float& impact) {
  const float cotThetaAvg2 = fmul(b.cotThetaB, cotThetaT);
  // erT + erB + ((2 * (cotAvg2*varRM + varZM)) * iDeltaRB) * iDeltaRT, left to right
  const float corr = fmul(fmul(fmul(2.0f, fadd(fmul(cotThetaAvg2, varRM), varZM)), b.iDeltaRB), iDeltaRT);
  const float error2 = fadd(fadd(erT, b.erB), corr);
  const float deltaCotTheta = fsub(b.cotThetaB, cotThetaT);
  const float deltaCotTheta2 = fmul(deltaCotTheta, deltaCotTheta);
  if (deltaCotTheta2 > fadd(error2, b.scatteringInRegion2)) return kPairFailA;
  const float dU = fsub(uT, b.Ub);
  if (dU == 0) return kPairSkip;
  const float vT = *vTPtr;  // only pairs that pass the first slope cut read it
  const float A = fdiv(fsub(vT, b.Vb), dU);
  const float S2 = fadd(1.0f, fmul(A, A));
  const float B = fsub(b.Vb, fmul(A, b.Ub));
  const float B2 = fmul(B, B);
  if (S2 < fmul(B2, c.minHelixDiameter2)) return kPairSkip;
  const float iHelixDiameter2 = fdiv(B2, S2);
  const float p2scatterSigma = fmul(iHelixDiameter2, b.sigmaSquaredPtDependent);
  if (deltaCotTheta2 > fadd(error2, p2scatterSigma)) return kPairFailB;
  const float im = fabs_(fmul(fsub(A, fmul(B, rM)), rM));
  if (im > c.impactMax) return kPairSkip;
  if (kWithCurvature) {
    curvature = fdiv(B, fsqrt(S2));
    impact = im;
  }
  return kPairEmit;
}
B2S_HD int eval_pair(const DeviceConfig& c, float rM, float varZM, float varRM,
                     const BottomCtx& b, float cotThetaT, float erT, float iDeltaRT,
                     float uT, float vT, float& curvature, float& impact) {
  return eval_pair_t<true>(c, rM, varZM, varRM, b, cotThetaT, erT, iDeltaRT, uT, &vT, curvature, impact);
}
// classification only (the scans): no curvature division / square root
B2S_HD int classify_pair(const DeviceConfig& c, float rM, float varZM, float varRM,
                         const BottomCtx& b, float cotThetaT, float erT, float iDeltaRT,
                         float uT, float vT) {
  float cu, im;
  return eval_pair_t<false>(c, rM, varZM, varRM, b, cotThetaT, erT, iDeltaRT, uT, &vT, cu, im);
}
// Branch-free classification for the scans: every term of eval_pair_t is evaluated unconditionally (the same
// operations on the same inputs, so the same bits) and the class is selected at the end.  A warp of walkers takes
// every path of the branchy version in almost every step anyway; straight-line code lets the independent chains
// (first slope cut | conformal line A, S2, B | measured-pT slope cut | impact) overlap instead of waiting for one
// another.  dU == 0 divides by zero (inf / NaN, discarded by the selection; IEEE division does not trap).
B2S_HD int classify_pair_flat(const DeviceConfig& c, float rM, float varZM, float varRM,
                              const BottomCtx& b, float cotThetaT, float erT, float iDeltaRT,
                              float uT, float vT) {
  const float cotThetaAvg2 = fmul(b.cotThetaB, cotThetaT);
  const float corr = fmul(fmul(fmul(2.0f, fadd(fmul(cotThetaAvg2, varRM), varZM)), b.iDeltaRB), iDeltaRT);
  const float error2 = fadd(fadd(erT, b.erB), corr);
  const float deltaCotTheta = fsub(b.cotThetaB, cotThetaT);
  const float deltaCotTheta2 = fmul(deltaCotTheta, deltaCotTheta);
  const bool failA = deltaCotTheta2 > fadd(error2, b.scatteringInRegion2);
  const float dU = fsub(uT, b.Ub);
  const float A = fdiv(fsub(vT, b.Vb), dU);
  const float S2 = fadd(1.0f, fmul(A, A));
  const float B = fsub(b.Vb, fmul(A, b.Ub));
  const float B2 = fmul(B, B);
  const bool skipHelix = (dU == 0) | (S2 < fmul(B2, c.minHelixDiameter2));
  const float iHelixDiameter2 = fdiv(B2, S2);
  const float p2scatterSigma = fmul(iHelixDiameter2, b.sigmaSquaredPtDependent);
  const bool failB = deltaCotTheta2 > fadd(error2, p2scatterSigma);
  const float im = fabs_(fmul(fsub(A, fmul(B, rM)), rM));
  const bool skipImpact = im > c.impactMax;
  return failA ? kPairFailA : (skipHelix ? kPairSkip : (failB ? kPairFailB : (skipImpact ? kPairSkip : kPairEmit)));
}

// the scans: v of the top is read only when the pair gets past the first slope cut
B2S_HD int classify_pair_lazy(const DeviceConfig& c, float rM, float varZM, float varRM,
                              const BottomCtx& b, float cotThetaT, float erT, float iDeltaRT,
                              float uT, const float* vTPtr) {
  float cu, im;
  return eval_pair_t<false>(c, rM, varZM, varRM, b, cotThetaT, erT, iDeltaRT, uT, vTPtr, cu, im);
}

// ---------------------------------------------------------------------------
// Triplets, strip path: TripletSeedFinder.cpp:164-406 (useStripInfo = true) with
// detail/StripSpacePointCalibrationImpl.hpp:22-85.  A space point carries the
// derived calibration details (three cross products, outer strip centre and half
// vector) as one 64-byte record, computed once per event by k_gather_strips.
// ---------------------------------------------------------------------------
struct StripDerived {  // StripSpacePointCalibrationDetails.hpp:33-49
  float ihvXohv[3];   // innerCrossOuterHalfVector
  float iosvXohv[3];  // innerToOuterSeparationCrossOuterHalfVector
  float iosvXihv[3];  // innerToOuterSeparationCrossInnerHalfVector
  float oc[3], ohv[3];
  float pad;
};
static_assert(sizeof(StripDerived) == 64, "four 16-byte words per space point");
// Utilities/detail/StdArrayLinalg.hpp:70-83
B2S_HD void strip_cross(const float* a, const float* b, float* out) {
  out[0] = fsub(fmul(a[1], b[2]), fmul(a[2], b[1]));
  out[1] = fsub(fmul(a[2], b[0]), fmul(a[0], b[2]));
  out[2] = fsub(fmul(a[0], b[1]), fmul(a[1], b[0]));
}
// raw: outerCenter, innerToOuterSeparation, outerHalfVector, innerHalfVector (12 floats)
B2S_HD void strip_derive(const float* raw, StripDerived& o) {  // StripSpacePointCalibrationImpl.hpp:22-42
  const float* oc = raw;
  const float* iosv = raw + 3;
  const float* ohv = raw + 6;
  const float* ihv = raw + 9;
  strip_cross(ihv, ohv, o.ihvXohv);
  strip_cross(iosv, ohv, o.iosvXohv);
  strip_cross(iosv, ihv, o.iosvXihv);
  for (int i = 0; i < 3; ++i) { o.oc[i] = oc[i]; o.ohv[i] = ohv[i]; }
  o.pad = 0.0f;
}
// stdArrayDot: result = 0; result += a[i] * b[i] (StdArrayLinalg.hpp:60-67)
B2S_HD float strip_dot(float dx, float dy, float dz, const float* b) {
  return fadd(fadd(fadd(0.0f, fmul(dx, b[0])), fmul(dy, b[1])), fmul(dz, b[2]));
}
// calibrateOuterStripSpacePoint, StripSpacePointCalibrationImpl.hpp:44-85
B2S_HD bool strip_calibrate(float dx, float dy, float dz, const StripDerived& sp, float tolerance, float* out) {
  const float scale = strip_dot(dx, dy, dz, sp.ihvXohv);
  const float limit = fmul(fabs_(scale), tolerance);
  const float sInner = strip_dot(dx, dy, dz, sp.iosvXohv);
  if (fabs_(sInner) > limit) return false;
  const float sOuter = strip_dot(dx, dy, dz, sp.iosvXihv);
  if (fabs_(sOuter) > limit) return false;
  const float sOuterNorm = fdiv(sOuter, scale);
  for (int i = 0; i < 3; ++i) out[i] = fadd(sp.oc[i], fmul(sp.ohv[i], sOuterNorm));
  return true;
}

struct StripBottomCtx {  // loop invariants of one (middle, bottom doublet), :172-212
  float cotThetaB0, erB, iDeltaRB, Ub0, Vb0, xB, yB;
  float sigmaSquaredPtDependent, scatteringInRegion2;
  float rot0, rot1, cosTheta;
};
B2S_HD void strip_bottom_ctx(const DeviceConfig& c, float cosPhiM, float sinPhiM, StripBottomCtx& b) {
  const float iSinTheta2 = fadd(1.0f, fmul(b.cotThetaB0, b.cotThetaB0));
  b.sigmaSquaredPtDependent = fmul(iSinTheta2, c.sigmapT2perRadius);
  b.scatteringInRegion2 = fmul(c.multipleScattering2, iSinTheta2);
  const float sinTheta = fdiv(1.0f, fsqrt(iSinTheta2));
  b.cosTheta = fmul(b.cotThetaB0, sinTheta);
  b.rot0 = fmul(cosPhiM, sinTheta);
  b.rot1 = fmul(sinPhiM, sinTheta);
}
// the cot(theta) pre-filter of :226-238: true = the top is outside the window of this bottom
B2S_HD bool strip_outside_window(float cotThetaB0, float cotThetaT, float cotThetaDiffMax2) {
  const float deltaCotTheta = fsub(cotThetaB0, cotThetaT);
  return fmul(deltaCotTheta, deltaCotTheta) > cotThetaDiffMax2;
}
// One (bottom, top) pair inside the window, :240-395.  true = candidate {top, curvature, impact}.
B2S_HD bool eval_strip_pair(const DeviceConfig& c, float toleranceParam, float varZM, float varRM, const StripBottomCtx& b,
                            const StripDerived& calM, const StripDerived& calB, const StripDerived& calT, float erT,
                            float iDeltaRT, float uT, float vT, float xT, float yT, float& curvature, float& impact) {
  const float dU0 = fsub(uT, b.Ub0);
  if (dU0 == 0) return false;
  const float A0 = fdiv(fsub(vT, b.Vb0), dU0);
  float rM[3], rB[3], rT[3];
  if (!strip_calibrate(fsub(b.rot0, fmul(b.rot1, A0)), fadd(fmul(b.rot0, A0), b.rot1), b.cosTheta, calM, toleranceParam, rM)) return false;
  const float zDirectionMiddle = fmul(b.cosTheta, fsqrt(fadd(1.0f, fmul(A0, A0))));
  const float B0 = fmul(2.0f, fsub(b.Vb0, fmul(A0, b.Ub0)));
  const float Cb = fsub(1.0f, fmul(B0, b.yB));
  const float Sb = fadd(A0, fmul(B0, b.xB));
  if (!strip_calibrate(fsub(fmul(b.rot0, Cb), fmul(b.rot1, Sb)), fadd(fmul(b.rot0, Sb), fmul(b.rot1, Cb)), zDirectionMiddle, calB,
                       toleranceParam, rB)) return false;
  const float Ct = fsub(1.0f, fmul(B0, yT));
  const float St = fadd(A0, fmul(B0, xT));
  if (!strip_calibrate(fsub(fmul(b.rot0, Ct), fmul(b.rot1, St)), fadd(fmul(b.rot0, St), fmul(b.rot1, Ct)), zDirectionMiddle, calT,
                       toleranceParam, rT)) return false;
  const float xB = fsub(rB[0], rM[0]), yB = fsub(rB[1], rM[1]), zB = fsub(rB[2], rM[2]);
  const float xTn = fsub(rT[0], rM[0]), yTn = fsub(rT[1], rM[1]), zT = fsub(rT[2], rM[2]);
  const float iDeltaRB2 = fdiv(1.0f, fadd(fmul(xB, xB), fmul(yB, yB)));
  const float iDeltaRT2 = fdiv(1.0f, fadd(fmul(xTn, xTn), fmul(yTn, yTn)));
  const float cotThetaB = fmul(-zB, fsqrt(iDeltaRB2));
  const float cotThetaT = fmul(zT, fsqrt(iDeltaRT2));
  const float averageCotTheta = fmul(0.5f, fadd(cotThetaB, cotThetaT));
  const float cotThetaAvg2 = fmul(averageCotTheta, averageCotTheta);
  const float corr = fmul(fmul(fmul(2.0f, fadd(fmul(cotThetaAvg2, varRM), varZM)), b.iDeltaRB), iDeltaRT);
  const float error2 = fadd(fadd(erT, b.erB), corr);
  const float deltaCotTheta = fsub(cotThetaB, cotThetaT);
  const float deltaCotTheta2 = fmul(deltaCotTheta, deltaCotTheta);
  if (deltaCotTheta2 > fadd(error2, b.scatteringInRegion2)) return false;
  const float rMxy = fsqrt(fadd(fmul(rM[0], rM[0]), fmul(rM[1], rM[1])));
  const float irMxy = fdiv(1.0f, rMxy);
  const float Ax = fmul(rM[0], irMxy), Ay = fmul(rM[1], irMxy);
  const float Ub = fmul(fadd(fmul(xB, Ax), fmul(yB, Ay)), iDeltaRB2);
  const float Vb = fmul(fsub(fmul(yB, Ax), fmul(xB, Ay)), iDeltaRB2);
  const float Ut = fmul(fadd(fmul(xTn, Ax), fmul(yTn, Ay)), iDeltaRT2);
  const float Vt = fmul(fsub(fmul(yTn, Ax), fmul(xTn, Ay)), iDeltaRT2);
  const float dU = fsub(Ut, Ub);
  if (dU == 0) return false;
  const float A = fdiv(fsub(Vt, Vb), dU);
  const float S2 = fadd(1.0f, fmul(A, A));
  const float B = fsub(Vb, fmul(A, Ub));
  const float B2 = fmul(B, B);
  if (S2 < fmul(B2, c.minHelixDiameter2)) return false;
  const float iHelixDiameter2 = fdiv(B2, S2);
  const float p2scatterSigma = fmul(iHelixDiameter2, b.sigmaSquaredPtDependent);
  if (deltaCotTheta2 > fadd(error2, p2scatterSigma)) return false;
  const float im = fabs_(fmul(fsub(A, fmul(B, rMxy)), rMxy));
  if (im > c.impactMax) return false;
  curvature = fdiv(B, fsqrt(S2));
  impact = im;
  return true;
}

// ---------------------------------------------------------------------------
// Seed filter: BroadTripletSeedFilter.cpp:96-322 (seedConfirmation == false).
// The candidates of one (middle, bottom) pair are given in curvature-sorted
// order (curv[], topR[], impact[] indexed by sorted rank).  The weight of
// candidate `k` only depends on the other candidates (the reference's
// beginCompTopIndex is an exact monotone skip), so every candidate can be
// weighted independently.
// ---------------------------------------------------------------------------
template <typename CurvAt, typename TopRAt>
B2S_HD float filter_weight(const DeviceConfig& c, int n, int k, float impact, CurvAt curvAt,
                           TopRAt topRAt, uint32_t& nCompatOut) {
  const float invHelixDiameter = curvAt(k);
  const float lowerLimitCurv = fsub(invHelixDiameter, c.deltaInvHelixDiameter);
  const float upperLimitCurv = fadd(invHelixDiameter, c.deltaInvHelixDiameter);
  const float currentTopR = topRAt(k);
  float weight = fmul(-impact, c.impactWeightFactor);
  float compatR[kMaxCompatSeedLimit];
  uint32_t nCompat = 0;
  for (int o = 0; o < n; ++o) {
    if (o == k) continue;
    const float curvO = curvAt(o);
    if (curvO < lowerLimitCurv) continue;
    if (curvO > upperLimitCurv) break;
    const float otherTopR = topRAt(o);
    const float deltaR = fsub(currentTopR, otherTopR);
    if (fabs_(deltaR) < c.filterDeltaRMin) continue;
    bool newCompSeed = true;
    for (uint32_t q = 0; q < nCompat; ++q) {
      if (fabs_(fsub(compatR[q], otherTopR)) < c.filterDeltaRMin) {
        newCompSeed = false;
        break;
      }
    }
    if (newCompSeed) {
      if (nCompat < (uint32_t)kMaxCompatSeedLimit) compatR[nCompat] = otherTopR;
      ++nCompat;
      weight = fadd(weight, c.compatSeedWeight);
    }
    if (nCompat >= c.compatSeedLimit) break;
  }
  if ((float)nCompat > c.numSeedIncrement) {
    weight = fadd(weight, c.seedWeightIncrement);
  }
  nCompatOut = nCompat;
  return weight;
}
template <typename CurvAt, typename TopRAt>
B2S_HD float filter_weight(const DeviceConfig& c, int n, int k, float impact, CurvAt curvAt,
                           TopRAt topRAt) {
  uint32_t nCompat;
  return filter_weight(c, n, k, impact, curvAt, topRAt, nCompat);
}

// Seed confirmation: the region of a space point and its minimum number of tops
// (BroadTripletSeedFilter.cpp:72-82 for the middle, :120-134 for the bottom)
B2S_HD const ConfRange& conf_range(const DeviceConfig& c, float z) {
  const bool isForwardRegion = z > c.confCentral.zMaxSeedConf || z < c.confCentral.zMinSeedConf;
  return isForwardRegion ? c.confForward : c.confCentral;
}
B2S_HD uint32_t conf_n_top(const ConfRange& range, float r) {
  return r > range.rMaxSeedConf ? range.nTopForLargeR : range.nTopForSmallR;
}
// The confirmation part of the per-candidate filter (:253-276) that does not
// depend on the collector or on bestSeedQualityMap.  Returns false when the
// candidate is dropped; otherwise `weight` gets its z-origin term and
// `deltaSeedConf` (>= 0) is set.
B2S_HD bool conf_candidate(const DeviceConfig& c, const ConfRange& rangeB, float rB, float zOrigin,
                           float impact, uint32_t nCompat, float& weight, int& deltaSeedConf) {
  deltaSeedConf = (int)(nCompat + 1u - conf_n_top(rangeB, rB));
  if (deltaSeedConf < 0) return false;
  const bool seedRangeCuts = rB < rangeB.seedConfMinBottomRadius || fabs_(zOrigin) > rangeB.seedConfMaxZOrigin;
  if (seedRangeCuts && deltaSeedConf == 0 && impact > rangeB.minImpactSeedConf) return false;
  weight = fadd(weight, fadd(-fmul(fabs_(zOrigin), c.zOriginWeightFactor), c.compatSeedWeight));
  return true;
}

// ---------------------------------------------------------------------------
// libstdc++ (GCC 13, bits/stl_heap.h) binary-heap primitives, restated so that
// the bounded candidate heap of CandidatesForMiddleSp.cpp:44-93 behaves
// identically on ties.  `comp(a, b)` is the std comparator (a.first > b.first
// for the reference's min-heap on weight).
// ---------------------------------------------------------------------------
struct WeightIndex {
  float weight;
  uint32_t index;
};
B2S_HD bool heap_comp(const WeightIndex& a, const WeightIndex& b) { return a.weight > b.weight; }

template <typename T, typename Comp>
B2S_HD void std_push_heap_impl(T* first, int holeIndex, int topIndex, T value, Comp comp) {
  int parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && comp(first[parent], value)) {
    first[holeIndex] = first[parent];
    holeIndex = parent;
    parent = (holeIndex - 1) / 2;
  }
  first[holeIndex] = value;
}
template <typename T, typename Comp>
B2S_HD void std_adjust_heap(T* first, int holeIndex, int len, T value, Comp comp) {
  const int topIndex = holeIndex;
  int secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (comp(first[secondChild], first[secondChild - 1])) secondChild--;
    first[holeIndex] = first[secondChild];
    holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
    secondChild = 2 * (secondChild + 1);
    first[holeIndex] = first[secondChild - 1];
    holeIndex = secondChild - 1;
  }
  std_push_heap_impl(first, holeIndex, topIndex, value, comp);
}
// std::push_heap(first, first + n): the new element is first[n-1]
template <typename T, typename Comp>
B2S_HD void std_push_heap(T* first, int n, Comp comp) {
  std_push_heap_impl(first, n - 1, 0, first[n - 1], comp);
}
// std::pop_heap(first, first + n): moves the top to first[n-1]
template <typename T, typename Comp>
B2S_HD void std_pop_heap(T* first, int n, Comp comp) {
  if (n > 1) {
    const T value = first[n - 1];
    first[n - 1] = first[0];
    std_adjust_heap(first, 0, n - 1, value, comp);
  }
}
template <typename T, typename Comp>
B2S_HD void std_sort_heap(T* first, int n, Comp comp) {
  while (n > 1) {
    std_pop_heap(first, n, comp);
    --n;
  }
}
template <typename T, typename Comp>
B2S_HD void std_make_heap(T* first, int len, Comp comp) {
  if (len < 2) return;
  int parent = (len - 2) / 2;
  while (true) {
    const T value = first[parent];
    std_adjust_heap(first, parent, len, value, comp);
    if (parent == 0) return;
    parent--;
  }
}

// ---------------------------------------------------------------------------
// libstdc++ std::sort (GCC 13 bits/stl_algo.h: introsort with median-of-3,
// depth limit 2*floor(log2 n), heap-sort fallback, final insertion sort with
// threshold 16) restated with an explicit stack.  Used wherever the reference
// calls the UNSTABLE std::ranges::sort and the keys contain ties, so that tie
// order -- and with it window / heap decisions -- match the reference
// (GridTripletSeedingAlgorithm.cpp:224, DoubletSeedFinder.hpp:101,
// BroadTripletSeedFilter.cpp:145).  `less(a, b)` compares two elements.
// ---------------------------------------------------------------------------
template <typename T, typename Less>
B2S_HD void std_unguarded_linear_insert(T* a, int last, Less less) {
  const T val = a[last];
  int next = last - 1;
  while (less(val, a[next])) {
    a[last] = a[next];
    last = next;
    --next;
  }
  a[last] = val;
}
template <typename T, typename Less>
B2S_HD void std_insertion_sort(T* a, int first, int last, Less less) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (less(a[i], a[first])) {
      const T val = a[i];
      for (int k = i; k > first; --k) a[k] = a[k - 1];
      a[first] = val;
    } else {
      std_unguarded_linear_insert(a, i, less);
    }
  }
}
template <typename T, typename Less>
B2S_HD void std_sort(T* a, int n, Less less) {
  if (n <= 0) return;
  constexpr int kThreshold = 16;
  int lg = 0;
  for (unsigned v = (unsigned)n; v > 1; v >>= 1) ++lg;
  // explicit stack of (first, last, depth); sub-ranges are disjoint so the
  // processing order does not change the result
  int stackFirst[64], stackLast[64], stackDepth[64];
  int sp = 0;
  stackFirst[0] = 0; stackLast[0] = n; stackDepth[0] = 2 * lg; sp = 1;
  while (sp > 0) {
    --sp;
    int first = stackFirst[sp], last = stackLast[sp], depth = stackDepth[sp];
    while (last - first > kThreshold) {
      if (depth == 0) {
        // std::__partial_sort(first, last, last): make_heap + sort_heap
        std_make_heap(a + first, last - first, less);
        std_sort_heap(a + first, last - first, less);
        break;
      }
      --depth;
      // __unguarded_partition_pivot
      const int mid = first + (last - first) / 2;
      {  // __move_median_to_first(first, first + 1, mid, last - 1)
        const int ia = first + 1, ib = mid, ic = last - 1;
        int pick;
        if (less(a[ia], a[ib])) {
          if (less(a[ib], a[ic])) pick = ib;
          else if (less(a[ia], a[ic])) pick = ic;
          else pick = ia;
        } else if (less(a[ia], a[ic])) pick = ia;
        else if (less(a[ib], a[ic])) pick = ic;
        else pick = ib;
        const T t = a[first]; a[first] = a[pick]; a[pick] = t;
      }
      int lo = first + 1, hi = last;
      while (true) {  // __unguarded_partition(first + 1, last, pivot = first)
        while (less(a[lo], a[first])) ++lo;
        --hi;
        while (less(a[first], a[hi])) --hi;
        if (!(lo < hi)) break;
        const T t = a[lo]; a[lo] = a[hi]; a[hi] = t;
        ++lo;
      }
      const int cut = lo;
      // recurse on [cut, last), loop on [first, cut)
      stackFirst[sp] = cut; stackLast[sp] = last; stackDepth[sp] = depth; ++sp;
      last = cut;
    }
  }
  // __final_insertion_sort
  if (n > kThreshold) {
    std_insertion_sort(a, 0, kThreshold, less);
    for (int i = kThreshold; i != n; ++i) std_unguarded_linear_insert(a, i, less);
  } else {
    std_insertion_sort(a, 0, n, less);
  }
}

// Pruned replay of libstdc++ std::sort that only resolves TIE ORDER.
//
// For an element whose key is unique the final position is forced (number of
// smaller keys), whatever the algorithm does.  What the unstable introsort
// decides is the internal order of every group of equal keys.  That order can
// be obtained by replaying the introsort partitions only on ranges that still
// contain at least two flagged (tied) elements: a range without two tied
// elements cannot change any tie order, elements never leave their range, and
// the final insertion sort is stable (it never swaps equal keys).  After the
// pruned replay the left-to-right order of the flagged elements in `a` IS their
// order in the reference's std::sort output.
//
// `a` must hold the n elements in the reference's INPUT order; `flagged(e)`
// tells whether an element belongs to a tie group.  Cost ~2-3 n element visits
// for a few tie groups instead of n log n.
template <typename T, typename Less, typename Flagged>
B2S_HD void std_sort_replay_ties(T* a, int n, Less less, Flagged flagged) {
  if (n <= 16) return;  // plain insertion sort: stable, input order decides
  constexpr int kThreshold = 16;
  constexpr int kMaxTracked = 24;
  // positions of the flagged elements, kept up to date across swaps so that
  // "does this range still hold two tied elements" costs O(#tied) instead of a scan
  int fpos[kMaxTracked];
  int nf = 0;
  bool tracked = true;
  for (int i = 0; i < n; ++i) {
    if (flagged(a[i])) {
      if (nf < kMaxTracked) fpos[nf++] = i; else { tracked = false; }
    }
  }
  auto countIn = [&](int first, int last) {
    int c = 0;
    if (tracked) {
      for (int k = 0; k < nf; ++k) c += (fpos[k] >= first && fpos[k] < last) ? 1 : 0;
    } else {
      for (int i = first; i < last && c < 2; ++i) c += flagged(a[i]) ? 1 : 0;
    }
    return c;
  };
  auto swapAt = [&](int i, int j) {
    const T t = a[i]; a[i] = a[j]; a[j] = t;
    if (tracked && (flagged(t) || flagged(a[i]))) {
      for (int k = 0; k < nf; ++k) {
        if (fpos[k] == i) fpos[k] = j; else if (fpos[k] == j) fpos[k] = i;
      }
    }
  };
  int lg = 0;
  for (unsigned v = (unsigned)n; v > 1; v >>= 1) ++lg;
  int stackFirst[64], stackLast[64], stackDepth[64];
  int sp = 0;
  stackFirst[0] = 0; stackLast[0] = n; stackDepth[0] = 2 * lg; sp = 1;
  while (sp > 0) {
    --sp;
    int first = stackFirst[sp], last = stackLast[sp], depth = stackDepth[sp];
    while (last - first > kThreshold) {
      if (countIn(first, last) < 2) break;  // nothing left to decide in this range
      if (depth == 0) {
        std_make_heap(a + first, last - first, less);
        std_sort_heap(a + first, last - first, less);
        break;  // (positions inside a heap-sorted range no longer matter: it is final)
      }
      --depth;
      const int mid = first + (last - first) / 2;
      {
        const int ia = first + 1, ib = mid, ic = last - 1;
        int pick;
        if (less(a[ia], a[ib])) {
          if (less(a[ib], a[ic])) pick = ib;
          else if (less(a[ia], a[ic])) pick = ic;
          else pick = ia;
        } else if (less(a[ia], a[ic])) pick = ia;
        else if (less(a[ib], a[ic])) pick = ic;
        else pick = ib;
        swapAt(first, pick);
      }
      const T pivot = a[first];
      int lo = first + 1, hi = last;
      while (true) {
        while (less(a[lo], pivot)) ++lo;
        --hi;
        while (less(pivot, a[hi])) --hi;
        if (!(lo < hi)) break;
        swapAt(lo, hi);
        ++lo;
      }
      const int cut = lo;
      stackFirst[sp] = cut; stackLast[sp] = last; stackDepth[sp] = depth; ++sp;
      last = cut;
    }
  }
}

// ---------------------------------------------------------------------------
// Acts::estimateTrackParamsFromSeed (free parameters, FP64):
// Core/src/Seeding/EstimateTrackParamsFromSeed.cpp:20-160.  out = {pos0 (3),
// time, direction (3), q/p}.  Not bit-exact by contract: the reference goes
// through Eigen (general affine inverse, expression templates) and the host
// libm; the north star's tolerance for estimated parameters is 1e-9 relative.
// ---------------------------------------------------------------------------
B2S_HD void estimate_free_params(const double sp0[3], double t0, const double sp1[3], const double sp2[3],
                                 const double bField[3], double out[8]) {
  // estimationFrameLocalToGlobal, :20-41
  const double rel[3] = {sp1[0] - sp0[0], sp1[1] - sp0[1], sp1[2] - sp0[2]};
  const double bNorm = sqrt(bField[0] * bField[0] + bField[1] * bField[1] + bField[2] * bField[2]);
  const double zA[3] = {bField[0] / bNorm, bField[1] / bNorm, bField[2] / bNorm};
  double yA[3] = {zA[1] * rel[2] - zA[2] * rel[1], zA[2] * rel[0] - zA[0] * rel[2], zA[0] * rel[1] - zA[1] * rel[0]};
  const double yN = sqrt(yA[0] * yA[0] + yA[1] * yA[1] + yA[2] * yA[2]);
  yA[0] /= yN; yA[1] /= yN; yA[2] /= yN;
  const double xA[3] = {yA[1] * zA[2] - yA[2] * zA[1], yA[2] * zA[0] - yA[0] * zA[2], yA[0] * zA[1] - yA[1] * zA[0]};
  // local = R^-1 (p - sp0); R = [xA yA zA] is a rotation, its inverse is the transpose
  const double d2[3] = {sp2[0] - sp0[0], sp2[1] - sp0[1], sp2[2] - sp0[2]};
  const double l1[3] = {xA[0] * rel[0] + xA[1] * rel[1] + xA[2] * rel[2], yA[0] * rel[0] + yA[1] * rel[1] + yA[2] * rel[2],
                        zA[0] * rel[0] + zA[1] * rel[1] + zA[2] * rel[2]};
  const double l2[3] = {xA[0] * d2[0] + xA[1] * d2[1] + xA[2] * d2[2], yA[0] * d2[0] + yA[1] * d2[1] + yA[2] * d2[2],
                        zA[0] * d2[0] + zA[1] * d2[1] + zA[2] * d2[2]};
  // performConformalMapping, :75-86
  const double n1 = l1[0] * l1[0] + l1[1] * l1[1], n2 = l2[0] * l2[0] + l2[1] * l2[1];
  const double u1 = l1[0] / n1, v1 = l1[1] / n1, u2 = l2[0] / n2, v2 = l2[1] / n2;
  const double du = u2 - u1, dv = v2 - v1;
  const double A = dv / du;
  const double B = v1 - A * u1;
  const double bOverS = (v1 * u2 - v2 * u1) / sqrt(du * du + dv * dv);
  // computeDzDs, :43-65
  const double r1x = 2 * B * l1[0] + A, r1y = 2 * B * l1[1] - 1;
  const double r2x = 2 * B * l2[0] + A, r2y = 2 * B * l2[1] - 1;
  const double dPhi = atan2(r2y, r2x) - atan2(r1y, r1x);
  const double dZ = l2[2] - l1[2];
  const double hx = dPhi / 2;
  // Acts::sinc, MathHelpers.hpp:256-266
  const double eps = 1.4901161193847656e-08 * 6;  // sqrt(DBL_EPSILON) * 6
  const double sincC = fabs(hx) < eps ? 1.0 : sin(hx) / hx;
  const double ddx = l2[0] - l1[0], ddy = l2[1] - l1[1];
  const double dzds = sincC * dZ / sqrt(ddx * ddx + ddy * ddy);
  // computeLocalTangent at local0 = (0, 0): r = (A, -1), t = (-r.y, r.x, |r| dzds), :88-97
  double t[3] = {1.0, A, sqrt(A * A + 1.0) * dzds};
  const double tn = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  t[0] /= tn; t[1] /= tn; t[2] /= tn;
  out[0] = sp0[0]; out[1] = sp0[1]; out[2] = sp0[2];
  out[3] = t0;
  out[4] = xA[0] * t[0] + yA[0] * t[1] + zA[0] * t[2];
  out[5] = xA[1] * t[0] + yA[1] * t[1] + zA[1] * t[2];
  out[6] = xA[2] * t[0] + yA[2] * t[1] + zA[2] * t[2];
  const double qOverPt = 2 * bOverS / bNorm;  // :138-141
  out[7] = qOverPt / sqrt(1.0 + dzds * dzds);
}

// ---------------------------------------------------------------------------
// Pixel space point from one measurement on a planar surface ("next" row f4):
// createPixelSpacePoint (Examples/Algorithms/TrackFinding/src/SpacePointMaker.cpp:44-76),
// PlaneSurface::localToGlobal (Core/src/Surfaces/PlaneSurface.cpp:72-75),
// Surface::referenceFrame (Core/src/Surfaces/Surface.cpp:243-247) and
// PixelSpacePointBuilder::computeCovarianceZR (Core/src/SpacePointFormation/PixelSpacePointBuilder.cpp:17-42).
// T is the row-major 3x4 affine local->global transform of the surface; the six outputs are the
// float columns of the SpacePointContainer.  FP64 like the reference (Eigen), rounded to float at the end.
// ---------------------------------------------------------------------------
B2S_HD void pixel_space_point(const double T[12], double loc0, double loc1, double c00, double c01, double c11,
                              float out[6]) {
  // global = linear * (loc0, loc1, 0) + translation
  const double gx = T[0] * loc0 + T[1] * loc1 + T[3];
  const double gy = T[4] * loc0 + T[5] * loc1 + T[7];
  const double gz = T[8] * loc0 + T[9] * loc1 + T[11];
  const double rr = sqrt(gx * gx + gy * gy);  // fastHypot, MathHelpers.hpp:93-96
  const double scale = 1 / rr;
  // jacXyzToZr = [[0, 0, 1], [scale x, scale y, 0]];  jac = jacXyzToZr * rot.topLeftCorner<3, 2>()
  const double jx = scale * gx, jy = scale * gy;
  const double j00 = T[8], j01 = T[9];
  const double j10 = jx * T[0] + jy * T[4], j11 = jx * T[1] + jy * T[5];
  // diag(jac * cov * jac^T)
  const double varZ = j00 * (c00 * j00 + c01 * j01) + j01 * (c01 * j00 + c11 * j01);
  const double varR = j10 * (c00 * j10 + c01 * j11) + j11 * (c01 * j10 + c11 * j11);
  out[0] = (float)gx; out[1] = (float)gy; out[2] = (float)gz; out[3] = (float)rr;
  out[4] = (float)varZ; out[5] = (float)varR;
}

// Monotone (non-decreasing in the key) bucket of a cot(theta) key for the
// in-block bucket sort.  Keys are expected inside [-cotThetaMax, cotThetaMax]
// but any value is clamped.
B2S_HD int cot_bucket(float cot, float cotMax, float scale, int nBuckets) {
  float t = fmul(fadd(cot, cotMax), scale);
  if (!(t > 0.0f)) t = 0.0f;
  int b = (int)t;
  if (b >= nBuckets) b = nBuckets - 1;
  return b;
}

}  // namespace B200SEED_NS
