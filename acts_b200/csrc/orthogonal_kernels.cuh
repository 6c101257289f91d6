// orthogonal_kernels.cuh -- the k-d-tree candidate provider of OrthogonalTripletSeedingAlgorithm
// (Examples/Algorithms/TrackFinding/src/OrthogonalTripletSeedingAlgorithm.cpp:101-317,
// Core/src/Seeding/CylindricalSpacePointKDTree.cpp:159-263, Core/include/Acts/Utilities/KDTree.hpp:352-396)
// in front of the same doublet arena / per-middle seeding kernel as the grid algorithm.
//
//   k_orth_gather       packed copy of the selected space points in the tree's ELEMENT ORDER (= the order the
//                       reference visits the middles in and receives range-search results in) + the middle
//                       selection of .cpp:253-275: every accepted middle yields TWO work items, the
//                       increasing-z group ("lh") and the decreasing-z group ("hl"), in that order (.cpp:294-306)
//   k_orth_fill_work    work list from the exclusive scan of the per-position item counts
//   k_doublets_kd<fill> the two-pass doublet stage of k_doublets with the r-sorted bin windows replaced by a
//                       range search: one THREAD per work item walks the event's tree (nodes in pre-order with skip
//                       indices, no stack) and tests the elements of every reported node in element order; deltaR
//                       is an explicit cut (DoubletSeedFinder.cpp:126-130, spacePointsSortedByRadius = false)
#pragma once

#include "kd_tree_host.hpp"
#include "seeding_kernels.cuh"

namespace B200SEED_NS {

struct OrthParams {
  DeviceConfig cfg;
  OrthDeviceConfig orth;
  uint32_t nEvents, nCoreTotal;
  const uint32_t* spOffsets;    // [nEvents + 1] the caller's events
  const uint32_t* coreOffsets;  // [nEvents + 1] selected space points (= tree elements) per event
  const float* rMiddleRange;    // [2 * nEvents] variable middle range per event
  const uint32_t* posOrig;      // [nCoreTotal] element position -> index inside the caller's event
  const float* posPhi;          // [nCoreTotal]
  const KdNodeDev* nodes;
  const float *x, *y, *z, *r, *varZ, *varR;
  uint32_t* pIdx;
  float2 *pXY, *pZR, *pVar;
  uint32_t* itemCount;  // [nCoreTotal] 0 or 2
  uint32_t* workStart;  // [nCoreTotal + 1] exclusive scan of itemCount
  uint32_t* workPos;    // [nWork] packed position of the middle
  uint32_t* workEG;     // [nWork] event * 2 + direction
};

__device__ __forceinline__ uint32_t orth_event_of(const uint32_t* offsets, uint32_t nEvents, uint32_t p) {
  uint32_t lo = 0, hi = nEvents;  // last offset <= p
  while (lo < hi) {
    const uint32_t mid = (lo + hi + 1) >> 1;
    if (__ldg(offsets + mid) <= p) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_orth_gather(const __grid_constant__ OrthParams p) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < p.nCoreTotal; q += stride) {
    const uint32_t e = orth_event_of(p.coreOffsets, p.nEvents, q);
    const uint32_t orig = __ldg(p.posOrig + q);
    const uint32_t src = __ldg(p.spOffsets + e) + orig;
    const float z = __ldg(p.z + src), r = __ldg(p.r + src);
    p.pIdx[q] = orig;
    p.pXY[q] = make_float2(__ldg(p.x + src), __ldg(p.y + src));
    p.pZR[q] = make_float2(z, r);
    p.pVar[q] = make_float2(__ldg(p.varZ + src), __ldg(p.varR + src));
    // middle selection, .cpp:253-275
    bool ok;
    if (p.cfg.useVariableMiddleSPRange) {
      ok = !(r < __ldg(p.rMiddleRange + 2 * e) || r > __ldg(p.rMiddleRange + 2 * e + 1));
    } else {
      ok = !(r > p.cfg.rMaxMiddle || r < p.cfg.rMinMiddle);
    }
    ok = ok && !(z < p.orth.zOutermostLayersMin || z > p.orth.zOutermostLayersMax);
    const float phi = __ldg(p.posPhi + q);
    ok = ok && !(phi > p.orth.phiMax || phi < p.orth.phiMin);
    p.itemCount[q] = ok ? 2u : 0u;
  }
}

__global__ void __launch_bounds__(256) k_orth_fill_work(const __grid_constant__ OrthParams p) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < p.nCoreTotal; q += stride) {
    if (p.itemCount[q] == 0u) continue;
    const uint32_t e = orth_event_of(p.coreOffsets, p.nEvents, q);
    const uint32_t w = p.workStart[q];
    p.workPos[w] = q;
    p.workPos[w + 1] = q;
    p.workEG[w] = 2u * e;
    p.workEG[w + 1] = 2u * e + 1u;
  }
}

struct KdDoubletParams {
  DoubletParams d;  // work list, slot sizes, arena, headers, class lists (binStart / nav tables unused)
  OrthDeviceConfig orth;
  const KdNodeDev* nodes;  // node e = root of event e
  const float* posPhi;
};

// One side (bottom or top candidates) of one work item, walked by ONE THREAD: the range search + the doublet cuts.
// Count pass: returns the number of survivors of the deltaR window and the (z, r) cuts (the slot size).  Fill pass:
// every survivor is finished on the spot and written to the slot in emission order; returns the number of doublets.
// The 32 work items of a warp are consecutive in the work list = neighbours in the tree's element order (both z
// directions of a middle, then the next middle), so their walks visit nearly the same nodes (L1 hits) and have
// similar lengths.
template <bool kBottom, bool kFill>
__device__ __forceinline__ uint32_t kd_side(const KdDoubletParams& p, const MiddleSp& mid, const KdBox& box,
                                            const KdNodeDev* nodes, uint32_t rootNode,
                                            DoubletRecord* recOut, float* keyOut, float& cotMin, float& cotMax) {
  const DeviceConfig& cfg = p.d.cfg;
  const float dRMin = kBottom ? cfg.dRMinB : cfg.dRMinT, dRMax = kBottom ? cfg.dRMaxB : cfg.dRMaxT;
  uint32_t n = 0;
  float mn = 3.0e38f, mx = -3.0e38f;
  // KDTreeNode::rangeSearchMapDiscard as a pre-order scan: a node that does not overlap the box is skipped with
  // its subtree (the reference tests the overlap before it descends into a child, KDTree.hpp:377-385); a leaf,
  // or an internal node the box covers completely, reports its elements in element order (:363-374,386-395).
  uint32_t id = rootNode;
  while (id != kKdEnd) {
    const float4* nd = reinterpret_cast<const float4*>(nodes + id);
    const float4 a = __ldg(nd), b = __ldg(nd + 1);
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(nd + 2));
    // a = {mnPhi, mnR, mnZ, mxPhi}, b = {mxR, mxZ, begin, end}, c = {skip, internal, lhs, -}
    const bool overlaps = (a.x < box.mx[0]) & (box.mn[0] < a.w) & (a.y < box.mx[1]) & (box.mn[1] < b.x) &
                          (a.z < box.mx[2]) & (box.mn[2] < b.y);
    if (!overlaps) { id = c.x; continue; }  // rope: the next node outside this subtree (kKdEnd after the last)
    const bool contained = (box.mn[0] <= a.x) & (box.mx[0] >= a.w) & (box.mn[1] <= a.y) & (box.mx[1] >= b.x) &
                           (box.mn[2] <= a.z) & (box.mx[2] >= b.y);
    if (c.y != 0u && !contained) { id = c.z; continue; }  // descend: left child (the right one is the left one's rope)
    const uint32_t e0 = __float_as_uint(b.z), e1 = __float_as_uint(b.w);  // packed positions (batch-wide)
    for (uint32_t o = e0; o < e1; ++o) {
      const float2 zr = ldg2(p.d.pZR + o);
      if (!contained) {
        const float phi = __ldg(p.posPhi + o);
        const bool inside = (box.mn[0] <= phi) & (phi < box.mx[0]) & (box.mn[1] <= zr.y) & (zr.y < box.mx[1]) &
                            (box.mn[2] <= zr.x) & (zr.x < box.mx[2]);
        if (!inside) continue;
      }
      float dR, dZ;
      if (!doublet_zr_cuts<kBottom>(cfg, mid, zr.x, zr.y, dR, dZ)) continue;
      if (outside_range(dR, dRMin, dRMax)) continue;  // (first in the reference: both are pure rejections)
      if (!kFill) {
        ++n;
      } else {
        const float2 xy = ldg2(p.d.pXY + o), var = ldg2(p.d.pVar + o);
        DoubletRec rec;
        if (!doublet_finish<kBottom>(cfg, mid, dR, dZ, xy.x, xy.y, zr.y, var.x, var.y, nullptr, nullptr, 0, rec, true)) continue;
        float4* dst = reinterpret_cast<float4*>(recOut + n);
        dst[0] = make_float4(__uint_as_float(o), rec.cotTheta, rec.iDeltaR, rec.er);
        dst[1] = make_float4(rec.u, rec.v, rec.xNew, rec.yNew);
        keyOut[n] = rec.cotTheta;
        mn = fminf(mn, rec.cotTheta);
        mx = fmaxf(mx, rec.cotTheta);
        ++n;
      }
    }
    id = c.x;
  }
  cotMin = mn;
  cotMax = mx;
  return n;
}

constexpr int kKdThreads = 128;

template <bool kFill>
__global__ void __launch_bounds__(kKdThreads) k_doublets_kd(const __grid_constant__ KdDoubletParams kp) {
  const DoubletParams& p = kp.d;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t itemFirst = kFill ? p.itemFirst : 0u;
  const uint32_t itemEnd = kFill ? p.itemEnd : *p.nWorkPtr;
  const uint32_t nItems = itemEnd - itemFirst;
  unsigned long long cntMiddles = 0, cntB = 0, cntT = 0;
  uint32_t maxFoot = 0, maxB = 0, maxT = 0;
  for (;;) {
    uint32_t it = 0;
    if (lane == 0) it = atomicAdd(p.workCounter, 32u);  // a warp takes 32 consecutive items
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= nItems) break;
    it += lane;
    if (it < nItems) {
      const uint32_t w = itemFirst + it;
      uint32_t capT = 0, capB = 0;
      bool live = true;
      if (kFill) {
        capT = __ldg(p.capT + w);
        capB = __ldg(p.capB + w);
        if (capT == 0u || capB == 0u) {
          MiddleHeader h{};
          h.capB = capB;
          p.hdr[w] = h;
          p.slotCount[w] = 0;
          live = false;
        }
      }
      if (live) {
        const uint32_t m = __ldg(p.workPos + w);
        const uint32_t eg = __ldg(p.workEG + w);
        const uint32_t ev = eg >> 1;
        const int dir = (int)(eg & 1u);
        const KdNodeDev* nodes = kp.nodes;
        const uint32_t rootNode = ev;  // node e is the root of event e
        MiddleSp mid;
        {
          const float2 mxy = ldg2(p.pXY + m), mzr = ldg2(p.pZR + m), mvar = ldg2(p.pVar + m);
          mid.x = mxy.x; mid.y = mxy.y; mid.z = mzr.x; mid.r = mzr.y; mid.varZ = mvar.x; mid.varR = mvar.y;
          middle_info(mid);
        }
        KdBox boxB, boxT;
        kd_search_boxes(kp.orth, __ldg(kp.posPhi + m), mid.r, mid.z, dir, boxB, boxT);
        const bool searchable = !kd_degenerate(boxB) && !kd_degenerate(boxT);  // CylindricalSpacePointKDTree.cpp:226,241
        if (!kFill) {
          ++cntMiddles;
          float a, b;
          if (searchable) {
            capT = kd_side<false, false>(kp, mid, boxT, nodes, rootNode, nullptr, nullptr, a, b);
            if (capT != 0u) capB = kd_side<true, false>(kp, mid, boxB, nodes, rootNode, nullptr, nullptr, a, b);
          }
          if (capB == 0u) capT = 0u;
          if (capB > kMaxListLength || capT > kMaxListLength) {
            atomicOr(p.status, kStatusOverflowDoublets);
            capB = 0; capT = 0;
          }
          capB = (capB + 3u) & ~3u;  // slots and their two halves start on 16-byte boundaries of the key array (TMA)
          capT = (capT + 3u) & ~3u;
          p.capB[w] = capB;
          p.capT[w] = capT;
          if (capB != 0u) {
            const uint32_t foot = seed_carve(capB, capT).minBytes;
            maxFoot = foot > maxFoot ? foot : maxFoot;
            maxB = capB > maxB ? capB : maxB;
            maxT = capT > maxT ? capT : maxT;
          }
        } else {
          const unsigned long long slot = p.slotPrefix[w] - p.slotPrefix[p.itemFirst];
          DoubletRecord* recSlot = p.rec + slot;
          float* keySlot = p.key + slot;
          float mnT = 0.f, mxT = 0.f, mnB = 0.f, mxB = 0.f;
          const uint32_t nT = kd_side<false, true>(kp, mid, boxT, nodes, rootNode, recSlot + capB, keySlot + capB, mnT, mxT);
          uint32_t nB = 0;
          if (nT != 0u) nB = kd_side<true, true>(kp, mid, boxB, nodes, rootNode, recSlot, keySlot, mnB, mxB);
          const bool go = nT != 0u && nB != 0u;
          MiddleHeader h{};
          h.capB = capB;
          h.offset = (uint32_t)slot;
          if (go) {
            h.nB = nB; h.nT = nT;
            h.cotMinB = float_to_ordered(mnB); h.cotMaxB = float_to_ordered(mxB);
            h.cotMinT = float_to_ordered(mnT); h.cotMaxT = float_to_ordered(mxT);
            const SeedCarve cv = seed_carve(nB, nT);
            p.carve[w] = cv;
            const uint32_t foot = cv.minBytes;
            int c = 0;
            while (c < kSpillClass && foot > p.classBytes[c]) ++c;
            p.classList[(size_t)c * p.classStride + atomicAdd(p.classCount + c, 1u)] = w;
            cntB += nB;
            cntT += nT;
          } else {
            p.slotCount[w] = 0;
          }
          p.hdr[w] = h;
        }
      }
    }
  }
  // per-warp totals -> global counters
  for (int d = 16; d > 0; d >>= 1) {
    cntMiddles += __shfl_xor_sync(0xffffffffu, cntMiddles, d);
    cntB += __shfl_xor_sync(0xffffffffu, cntB, d);
    cntT += __shfl_xor_sync(0xffffffffu, cntT, d);
    maxFoot = max(maxFoot, __shfl_xor_sync(0xffffffffu, maxFoot, d));
    maxB = max(maxB, __shfl_xor_sync(0xffffffffu, maxB, d));
    maxT = max(maxT, __shfl_xor_sync(0xffffffffu, maxT, d));
  }
  if (lane == 0) {
    if (!kFill) {
      if (cntMiddles != 0ull) atomicAdd(p.counters + kCntMiddles, cntMiddles);
      if (maxFoot != 0u) {
        atomicMax(p.planWords + 0, maxFoot);
        atomicMax(p.planWords + 1, maxB);
        atomicMax(p.planWords + 2, maxT);
      }
    } else {
      if (cntB != 0ull) atomicAdd(p.counters + kCntBottomDoublets, cntB);
      if (cntT != 0ull) atomicAdd(p.counters + kCntTopDoublets, cntT);
    }
  }
}

}  // namespace B200SEED_NS
