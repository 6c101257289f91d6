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
//                       range search: every LANE is a walker with its own work item (rope walk, no stack; lane
//                       refill) and tests the elements of every reported node in element order; deltaR is an
//                       explicit cut (DoubletSeedFinder.cpp:126-130, spacePointsSortedByRadius = false)
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

// ---------------------------------------------------------------------------
// k-d tree construction on the device (Core/include/Acts/Utilities/KDTree.hpp:66-80,236-350): an in-place replay
// of the reference's node-by-node std::partition / std::sort over the element arrays of the whole batch.
//   k_kd_select   selection (.cpp:123-127), phi = replayed atan2f, extent of r (Extent::extend), element arrays in
//                 insertion order (ordered compaction through the scan of the selection flags when a selector is on)
//   k_kd_roots    one task per event; radius range of the middles (.cpp:227-232)
//   k_kd_split    one block per node with more than 128 elements: bounding box, split value = middle of the box in
//                 the node's dimension, std::partition replayed exactly -- libstdc++'s bidirectional partition swaps
//                 the k-th misplaced element from the left with the k-th misplaced element from the right, so the
//                 block ranks both kinds with a scan and performs the swaps in parallel --, child tasks.  One launch
//                 per tree level (the host reads the number of tasks of the next level).
//   k_kd_small    one warp per subtree of at most 128 elements, in shared memory: per sub-level every lane owns a node,
//                 sorts it by the node's dimension with the libstdc++ introsort replay (std_sort, seed_math.h), splits
//                 at the median and emits its children; leaves (at most 4 elements) are finished on the spot.
// Nodes carry ropes (`skip`) and their left child, so the searches need neither a stack nor any node order.
// ---------------------------------------------------------------------------
struct KdTask {
  uint32_t begin, end;  // packed element positions
  uint32_t node;        // the node record to fill
  uint32_t skip;        // its rope
  uint32_t dim;         // split dimension (phi, r, z cycle)
};
constexpr uint32_t kKdExactMedian = 128, kKdLeaf = 4;
constexpr int kKdSplitThreads = 1024;
constexpr int kKdSmallWarps = 8;

struct KdBuildParams {
  DeviceConfig cfg;
  uint32_t nEvents, nTotal;
  const uint32_t* spOffsets;
  const float *x, *y, *z, *r;
  uint32_t* selFlag;      // [nTotal] (selector on)
  uint32_t* selScan;      // [nTotal + 1] exclusive scan of selFlag
  float* phiTmp;          // [nTotal] phi in the caller's order (selector on)
  uint32_t* coreOffsets;  // [nEvents + 1]
  float *ePhi, *eR, *eZ;  // element arrays
  uint32_t* eIdx;         // element -> index inside the caller's event
  unsigned long long* extent;  // [nEvents] min, then [nEvents] max of perp, as bits (non-negative doubles order like integers)
  float* rMiddleRange;         // [2 * nEvents]
  KdNodeDev* nodes;
  uint32_t* counters;  // [0] nodes allocated, [1] small tasks, [2 + (level & 1)] tasks of the next level
  KdTask *tasksIn, *tasksOut, *tasksSmall;
  uint32_t nTasksIn, outSlot;
  uint32_t *listL, *listR;  // [nTotal] positions of the misplaced elements of every node, by rank
};

__global__ void __launch_bounds__(256) k_kd_select(const __grid_constant__ KdBuildParams p) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const bool selector = p.cfg.useExtraCuts != 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.nTotal; i += stride) {
    const uint32_t e = orth_event_of(p.spOffsets, p.nEvents, i);
    const float x = __ldg(p.x + i), y = __ldg(p.y + i), z = __ldg(p.z + i), r = __ldg(p.r + i);
    const bool sel = !(selector && !itk_sp_select(r, z));
#ifdef B200SEED_RELAXED
    const float phi = atan2f(y, x);
#else
    const float phi = glibc_atan2f(y, x);
#endif
    if (selector) {
      p.selFlag[i] = sel ? 1u : 0u;
      p.phiTmp[i] = phi;
    } else {  // every space point is an element: insertion order = the caller's order
      p.ePhi[i] = phi; p.eR[i] = r; p.eZ[i] = z;
      p.eIdx[i] = i - __ldg(p.spOffsets + e);
    }
    if (sel) {
      const double xd = (double)x, yd = (double)y;
      const double perp = sqrt(dadd(dmul(xd, xd), dmul(yd, yd)));  // VectorHelpers::perp of the double vector
      const unsigned long long bits = (unsigned long long)__double_as_longlong(perp);
      atomicMin(p.extent + e, bits);
      atomicMax(p.extent + p.nEvents + e, bits);
    }
  }
}

__global__ void __launch_bounds__(256) k_kd_compact(const __grid_constant__ KdBuildParams p) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.nTotal; i += stride) {
    if (p.selFlag[i] == 0u) continue;
    const uint32_t e = orth_event_of(p.spOffsets, p.nEvents, i);
    const uint32_t q = p.selScan[i];
    p.ePhi[q] = p.phiTmp[i]; p.eR[q] = __ldg(p.r + i); p.eZ[q] = __ldg(p.z + i);
    p.eIdx[q] = i - __ldg(p.spOffsets + e);
  }
}

__global__ void __launch_bounds__(256) k_kd_roots(const __grid_constant__ KdBuildParams p) {
  const bool selector = p.cfg.useExtraCuts != 0;
  for (uint32_t e = threadIdx.x; e <= p.nEvents; e += blockDim.x) {
    const uint32_t o = __ldg(p.spOffsets + e);
    p.coreOffsets[e] = selector ? p.selScan[o] : o;
  }
  if (threadIdx.x == 0) p.counters[0] = p.nEvents;  // node e = root of event e
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < p.nEvents; e += blockDim.x) {
    KdTask t;
    t.begin = p.coreOffsets[e];
    t.end = p.coreOffsets[e + 1];
    t.node = e;
    t.skip = kKdEnd;
    t.dim = 0;
    if (t.end - t.begin > kKdExactMedian) {
      p.tasksOut[atomicAdd(p.counters + 2 + p.outSlot, 1u)] = t;
    } else {
      p.tasksSmall[atomicAdd(p.counters + 1, 1u)] = t;
    }
    // .cpp:227-232 in double, stored as float (Range1D<float>)
    const bool any = t.end != t.begin;
    const double rLo = any ? __longlong_as_double((long long)p.extent[e]) : 0.0;
    const double rHi = any ? __longlong_as_double((long long)p.extent[p.nEvents + e]) : 0.0;
    p.rMiddleRange[2 * e] = (float)dadd(dmul(dfloor(ddiv(rLo, 2.0)), 2.0), (double)p.cfg.deltaRMiddleMinSPRange);
    p.rMiddleRange[2 * e + 1] = (float)dsub(dmul(dfloor(ddiv(rHi, 2.0)), 2.0), (double)p.cfg.deltaRMiddleMaxSPRange);
  }
}

__device__ __forceinline__ float kd_next_up(float v) { return nextafterf(v, 3.402823466e+38f); }  // KDTree::nextRepresentable

__global__ void __launch_bounds__(kKdSplitThreads) k_kd_split(const __grid_constant__ KdBuildParams p) {
  __shared__ float sMin[3][32], sMax[3][32];
  __shared__ uint32_t scratch[34];
  __shared__ uint32_t sLhs;
  const KdTask t = p.tasksIn[blockIdx.x];
  const uint32_t b = t.begin, e = t.end, d = t.dim;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* const key[3] = {p.ePhi, p.eR, p.eZ};
  // bounding box (KDTree::boundingBox)
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
  float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  for (uint32_t i = b + tid; i < e; i += blockDim.x) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float v = key[j][i];
      mn[j] = std_min(mn[j], v);
      mx[j] = std_max(mx[j], v);
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    for (int s = 16; s > 0; s >>= 1) {
      mn[j] = fminf(mn[j], __shfl_xor_sync(0xffffffffu, mn[j], s));
      mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], s));
    }
    if (lane == 0) { sMin[j][warp] = mn[j]; sMax[j][warp] = mx[j]; }
  }
  __syncthreads();
  const uint32_t nWarps = blockDim.x >> 5;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float a = lane < nWarps ? sMin[j][lane] : 3.402823466e+38f, c = lane < nWarps ? sMax[j][lane] : -3.402823466e+38f;
    for (int s = 16; s > 0; s >>= 1) {
      a = fminf(a, __shfl_xor_sync(0xffffffffu, a, s));
      c = fmaxf(c, __shfl_xor_sync(0xffffffffu, c, s));
    }
    mn[j] = a;
    mx[j] = kd_next_up(c);
  }
  // split value: the middle of the box in this dimension (KDTree.hpp:299-305)
  const float mid = fmul(0.5f, fadd(mx[d], mn[d]));
  const float* kd = key[d];
  uint32_t cnt = 0;
  for (uint32_t i = b + tid; i < e; i += blockDim.x) cnt += kd[i] < mid ? 1u : 0u;
  uint32_t nTrue;
  block_scan_exclusive(cnt, scratch, nTrue, OpSum());
  const uint32_t P = b + nTrue;  // std::partition's return value
  // std::partition (bidirectional): the k-th element of [b, P) that fails the predicate is swapped with the k-th
  // element from the right of [P, e) that satisfies it
  uint32_t carryL = 0, carryR = 0;
  for (uint32_t base = b; base < e; base += blockDim.x) {
    const uint32_t i = base + tid;
    const bool in = i < e;
    const bool pred = in && kd[i] < mid;
    const bool isL = in && i < P && !pred, isR = in && i >= P && pred;
    uint32_t total;
    const uint32_t excl = block_scan_exclusive((isL ? 1u : 0u) | (isR ? 0x10000u : 0u), scratch, total, OpSum());
    if (isL) p.listL[b + carryL + (excl & 0xFFFFu)] = i;
    if (isR) p.listR[b + carryR + (excl >> 16)] = i;
    carryL += total & 0xFFFFu;
    carryR += total >> 16;
  }
  __syncthreads();
  const uint32_t nMis = carryL;  // == carryR
  for (uint32_t k = tid; k < nMis; k += blockDim.x) {
    const uint32_t i = p.listL[b + k], j = p.listR[b + nMis - 1u - k];
    const float a0 = p.ePhi[i], a1 = p.eR[i], a2 = p.eZ[i];
    const uint32_t a3 = p.eIdx[i];
    p.ePhi[i] = p.ePhi[j]; p.eR[i] = p.eR[j]; p.eZ[i] = p.eZ[j]; p.eIdx[i] = p.eIdx[j];
    p.ePhi[j] = a0; p.eR[j] = a1; p.eZ[j] = a2; p.eIdx[j] = a3;
  }
  if (tid == 0) {
    uint32_t pivot = P;
    if (pivot == b || pivot == e - 1u) pivot = b + kKdLeaf;  // KDTree.hpp:322-324
    const uint32_t lhs = atomicAdd(p.counters + 0, 2u);
    KdNodeDev nd{};
    for (int j = 0; j < 3; ++j) { nd.mn[j] = mn[j]; nd.mx[j] = mx[j]; }
    nd.begin = b; nd.end = e; nd.skip = t.skip; nd.internal = 1u; nd.lhs = lhs;
    p.nodes[t.node] = nd;
    KdTask c[2];
    c[0].begin = b; c[0].end = pivot; c[0].node = lhs; c[0].skip = lhs + 1u; c[0].dim = (d + 1u) % 3u;
    c[1].begin = pivot; c[1].end = e; c[1].node = lhs + 1u; c[1].skip = t.skip; c[1].dim = (d + 1u) % 3u;
    for (int k = 0; k < 2; ++k) {
      if (c[k].end - c[k].begin > kKdExactMedian) {
        p.tasksOut[atomicAdd(p.counters + 2 + p.outSlot, 1u)] = c[k];
      } else {
        p.tasksSmall[atomicAdd(p.counters + 1, 1u)] = c[k];
      }
    }
    sLhs = lhs;
  }
  (void)sLhs;
}

struct KdElem {
  float c[3];
  uint32_t idx;
};

__device__ __forceinline__ void kd_bbox_smem(const KdElem* el, uint32_t n, float* mn, float* mx) {
  for (int j = 0; j < 3; ++j) { mn[j] = 3.402823466e+38f; mx[j] = -3.402823466e+38f; }
  for (uint32_t i = 0; i < n; ++i) {
    for (int j = 0; j < 3; ++j) {
      mn[j] = std_min(mn[j], el[i].c[j]);
      mx[j] = std_max(mx[j], el[i].c[j]);
    }
  }
  for (int j = 0; j < 3; ++j) mx[j] = kd_next_up(mx[j]);
}

__global__ void __launch_bounds__(kKdSmallWarps * 32) k_kd_small(const __grid_constant__ KdBuildParams p) {
  __shared__ KdElem sElem[kKdSmallWarps][kKdExactMedian];
  __shared__ KdTask sTask[kKdSmallWarps][2][32];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t nSmall = p.counters[1];
  const uint32_t w = blockIdx.x * kKdSmallWarps + wib;
  if (w >= nSmall) return;
  const KdTask root = p.tasksSmall[w];
  const uint32_t b0 = root.begin, n0 = root.end - root.begin;
  KdElem* el = sElem[wib];
  for (uint32_t i = lane; i < n0; i += 32) {
    KdElem v;
    v.c[0] = p.ePhi[b0 + i]; v.c[1] = p.eR[b0 + i]; v.c[2] = p.eZ[b0 + i]; v.idx = p.eIdx[b0 + i];
    el[i] = v;
  }
  if (lane == 0) sTask[wib][0][0] = root;
  __syncwarp();
  uint32_t nCur = 1;
  int cur = 0;
  while (nCur > 0) {
    const bool have = lane < nCur;
    KdTask t{};
    if (have) t = sTask[wib][cur][lane];
    const uint32_t n = t.end - t.begin;
    const bool internal = have && n > kKdLeaf;
    float mn[3], mx[3];
    uint32_t pivot = t.begin;
    if (have) {
      KdElem* mine = el + (t.begin - b0);
      kd_bbox_smem(mine, n, mn, mx);
      if (internal) {
        const uint32_t d = t.dim;
        std_sort(mine, (int)n, [d](const KdElem& a, const KdElem& c) { return a.c[d] < c.c[d]; });  // KDTree.hpp:310-316
        pivot = t.begin + n / 2u;
        if (pivot == t.begin || pivot == t.end - 1u) pivot = t.begin + kKdLeaf;
      }
    }
    // node ids of the children: two per internal node
    const uint32_t intMask = __ballot_sync(0xffffffffu, internal);
    uint32_t base = 0;
    if (lane == 0 && intMask != 0u) base = atomicAdd(p.counters + 0, 2u * (uint32_t)__popc(intMask));
    base = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t lhs = base + 2u * (uint32_t)__popc(intMask & ((1u << lane) - 1u));
    if (have) {
      KdNodeDev nd{};
      for (int j = 0; j < 3; ++j) { nd.mn[j] = mn[j]; nd.mx[j] = mx[j]; }
      nd.begin = t.begin; nd.end = t.end; nd.skip = t.skip; nd.internal = internal ? 1u : 0u; nd.lhs = internal ? lhs : 0u;
      p.nodes[t.node] = nd;
    }
    // children: leaves are finished here, internal children go to the next sub-level
    KdTask c[2];
    bool next[2] = {false, false};
    if (internal) {
      c[0].begin = t.begin; c[0].end = pivot; c[0].node = lhs; c[0].skip = lhs + 1u; c[0].dim = (t.dim + 1u) % 3u;
      c[1].begin = pivot; c[1].end = t.end; c[1].node = lhs + 1u; c[1].skip = t.skip; c[1].dim = (t.dim + 1u) % 3u;
      for (int k = 0; k < 2; ++k) {
        const uint32_t cn = c[k].end - c[k].begin;
        if (cn > kKdLeaf) {
          next[k] = true;
        } else {
          KdNodeDev leaf{};
          float lmn[3], lmx[3];
          kd_bbox_smem(el + (c[k].begin - b0), cn, lmn, lmx);
          for (int j = 0; j < 3; ++j) { leaf.mn[j] = lmn[j]; leaf.mx[j] = lmx[j]; }
          leaf.begin = c[k].begin; leaf.end = c[k].end; leaf.skip = c[k].skip; leaf.internal = 0u; leaf.lhs = 0u;
          p.nodes[c[k].node] = leaf;
        }
      }
    }
    const uint32_t m0 = __ballot_sync(0xffffffffu, next[0]), m1 = __ballot_sync(0xffffffffu, next[1]);
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t before = (uint32_t)__popc(m0 & lt) + (uint32_t)__popc(m1 & lt);
    if (next[0]) sTask[wib][cur ^ 1][before] = c[0];
    if (next[1]) sTask[wib][cur ^ 1][before + (next[0] ? 1u : 0u)] = c[1];
    nCur = (uint32_t)__popc(m0) + (uint32_t)__popc(m1);  // <= 32: a sub-level of a <= 128-element subtree with nodes of > 4 elements
    cur ^= 1;
    __syncwarp();
  }
  for (uint32_t i = lane; i < n0; i += 32) {
    const KdElem v = el[i];
    p.ePhi[b0 + i] = v.c[0]; p.eR[b0 + i] = v.c[1]; p.eZ[b0 + i] = v.c[2]; p.eIdx[b0 + i] = v.idx;
  }
}

struct KdDoubletParams {
  DoubletParams d;  // work list, slot sizes, arena, headers, class lists (binStart / nav tables unused)
  OrthDeviceConfig orth;
  const KdNodeDev* nodes;  // node e = root of event e
  const float* posPhi;
  // Hand-off from the count pass to the fill pass: the positions of the survivors of every walk, in emission order,
  // as a chain of 32-word pages (31 hits + the index of the next page; the list ends with kKdHitEnd).  hitHead[2 w]
  // / [2 w + 1] = first page of the top / bottom walk of item w (kKdHitNone: arena full -- that walk is repeated by
  // the fill pass).  The fill pass then reads positions instead of searching the tree a second time.
  uint32_t* hitArena;
  uint32_t hitPages;    // capacity in pages (0: no hand-off)
  uint32_t* hitCursor;  // next free page
  uint32_t* hitHead;    // [2 * items]
};
constexpr uint32_t kKdHitEnd = 0xFFFFFFFFu, kKdHitNone = 0xFFFFFFFFu;

// The two-pass doublet stage of the orthogonal seeder.  Every LANE is a walker with its own work item: it runs the
// range search of the item's top box, then (if tops were found) of its bottom box, one micro-step per loop iteration
// -- either one node visit or one element test -- so that the 32 walkers of a warp reconverge after every step; a lane
// whose item is finished takes the next unassigned item at once (lane refill), so warps stay full although the items
// differ widely (the decreasing-z group of a forward middle is nearly empty, the increasing-z group is not).
//   KDTreeNode::rangeSearchMapDiscard as a rope walk: a node that does not overlap the box is skipped with its
//   subtree (the reference tests the overlap before it descends into a child, KDTree.hpp:377-385); a leaf, or an
//   internal node the box covers completely, reports its elements in element order (:363-374,386-395).
// Count pass: survivors of the deltaR window and the (z, r) cuts per side (the slot size).  Fill pass: every survivor
// is finished on the spot and written to the item's slot in emission order.
constexpr int kKdThreads = 128;
#ifndef B200SEED_KD_SCAN_BELOW
#define B200SEED_KD_SCAN_BELOW 16  // measured against 4 (the reference leaf size: nothing scanned), 32, 64: tools/ab_kd.sh
#endif
constexpr int kKdScanBelow = B200SEED_KD_SCAN_BELOW;

template <bool kFill>
__global__ void __launch_bounds__(kKdThreads) k_doublets_kd(const __grid_constant__ KdDoubletParams kp) {
  const DoubletParams& p = kp.d;
  const DeviceConfig& cfg = p.cfg;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t ltMask = (1u << lane) - 1u;
  const uint32_t itemFirst = kFill ? p.itemFirst : 0u;
  const uint32_t itemEnd = kFill ? p.itemEnd : *p.nWorkPtr;
  const uint32_t nItems = itemEnd - itemFirst;
  unsigned long long cntMiddles = 0, cntB = 0, cntT = 0;
  uint32_t maxFoot = 0, maxB = 0, maxT = 0;
  // walker state
  int phase = 0;  // 0 idle, 1 top walk, 2 bottom walk, 3 retired
  uint32_t w = 0, m = 0, rootNode = 0;
  int dir = 0;
  MiddleSp mid{};
  KdBox box{};
  float dRMin = 0.f, dRMax = 0.f;
  uint32_t id = kKdEnd, o = 0, oEnd = 0;
  bool contained = false;
  uint32_t n = 0, nTop = 0, capT = 0, capB = 0;
  float mn = 0.f, mx = 0.f, mnT = 0.f, mxT = 0.f;
  DoubletRecord* recSlot = nullptr;
  float* keySlot = nullptr;
  DoubletRecord* recOut = nullptr;
  float* keyOut = nullptr;
  uint32_t slotLo = 0;
  bool exhausted = false;  // warp-uniform: the work list has been handed out
  const KdNodeDev* nodes = kp.nodes;
  // hit list of the current walk: count pass -- page being written / words used / overflow; fill pass -- page being
  // read / next word, listMode = this walk replays a list
  uint32_t hitPage = kKdHitNone, hitFill = 0;
  bool listMode = false;
  auto hitAlloc = [&]() -> uint32_t {
    const uint32_t pg = atomicAdd(kp.hitCursor, 1u);
    return pg < kp.hitPages ? pg : kKdHitNone;
  };
  auto hitPut = [&](uint32_t v) {  // count pass: append one word to the current walk's list
    if (hitPage == kKdHitNone) return;
    if (hitFill == 31u) {
      const uint32_t np = hitAlloc();
      kp.hitArena[32ull * hitPage + 31u] = np;  // (kKdHitNone: the reader never gets here, the head is withdrawn below)
      hitPage = np;
      hitFill = 0;
      if (np == kKdHitNone) return;
    }
    kp.hitArena[32ull * hitPage + hitFill++] = v;
  };

  auto startWalk = [&](bool bottom) {
    KdBox boxB, boxT;
    kd_search_boxes(kp.orth, __ldg(kp.posPhi + m), mid.r, mid.z, dir, boxB, boxT);
    box = bottom ? boxB : boxT;
    dRMin = bottom ? cfg.dRMinB : cfg.dRMinT;
    dRMax = bottom ? cfg.dRMaxB : cfg.dRMaxT;
    id = rootNode; o = 0; oEnd = 0; contained = false;
    n = 0; mn = 3.0e38f; mx = -3.0e38f;
    phase = bottom ? 2 : 1;
    listMode = false;
    if (kp.hitPages != 0u) {
      uint32_t* head = kp.hitHead + 2ull * w + (bottom ? 1u : 0u);
      if (!kFill) {
        hitPage = hitAlloc();
        hitFill = 0;
        *head = hitPage;
      } else {
        hitPage = *head;
        hitFill = 0;
        listMode = hitPage != kKdHitNone;
      }
    }
  };
  auto endList = [&](bool bottom) {  // count pass: close the list of the walk that just ended (or withdraw it)
    if (kp.hitPages == 0u) return;
    hitPut(kKdHitEnd);
    if (hitPage == kKdHitNone) kp.hitHead[2ull * w + (bottom ? 1u : 0u)] = kKdHitNone;
  };
  auto finishCount = [&]() {  // count pass: slot sizes of the item
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
    phase = 0;
  };
  auto finishFill = [&](uint32_t nB, float mnB, float mxB) {  // fill pass: header, carve-up, class list
    const bool go = nTop != 0u && nB != 0u;
    MiddleHeader h{};
    h.capB = capB;
    h.offset = slotLo;
    if (go) {
      h.nB = nB; h.nT = nTop;
      h.cotMinB = float_to_ordered(mnB); h.cotMaxB = float_to_ordered(mxB);
      h.cotMinT = float_to_ordered(mnT); h.cotMaxT = float_to_ordered(mxT);
      const SeedCarve cv = seed_carve(nB, nTop);
      p.carve[w] = cv;
      const uint32_t foot = cv.minBytes;
      int c = 0;
      while (c < kSpillClass && foot > p.classBytes[c]) ++c;
      p.classList[(size_t)c * p.classStride + atomicAdd(p.classCount + c, 1u)] = w;
      cntB += nB;
      cntT += nTop;
    } else {
      p.slotCount[w] = 0;
    }
    p.hdr[w] = h;
    phase = 0;
  };

  for (;;) {
    const uint32_t idle = __ballot_sync(0xffffffffu, phase == 0);
    if (idle != 0u) {
      if (!exhausted) {  // hand the next items to the idle lanes
        const uint32_t nIdle = (uint32_t)__popc(idle);
        const int leader = __ffs(idle) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(p.workCounter, nIdle);
        base = __shfl_sync(0xffffffffu, base, leader);
        exhausted = base + nIdle >= nItems;
        if (phase == 0) {
          const uint32_t it = base + (uint32_t)__popc(idle & ltMask);
          if (it >= nItems) {
            phase = 3;
          } else {
            w = itemFirst + it;
            bool live = true;
            capT = 0; capB = 0;
            if (kFill) {
              capT = __ldg(p.capT + w);
              capB = __ldg(p.capB + w);
              if (capT == 0u || capB == 0u) {
                MiddleHeader h{};
                h.capB = capB;
                p.hdr[w] = h;
                p.slotCount[w] = 0;
                live = false;  // stays idle: takes another item in the next round
              } else if (kp.hitPages != 0u && kp.hitHead[2ull * w] != kKdHitNone && kp.hitHead[2ull * w + 1u] != kKdHitNone) {
                live = false;  // both walks left their hit lists: k_doublets_kd_lists fills this item
              }
            }
            if (live) {
              m = __ldg(p.workPos + w);
              const uint32_t eg = __ldg(p.workEG + w);
              rootNode = eg >> 1;  // node e is the root of event e
              dir = (int)(eg & 1u);
              const float2 mxy = ldg2(p.pXY + m), mzr = ldg2(p.pZR + m), mvar = ldg2(p.pVar + m);
              mid.x = mxy.x; mid.y = mxy.y; mid.z = mzr.x; mid.r = mzr.y; mid.varZ = mvar.x; mid.varR = mvar.y;
              middle_info(mid);
              KdBox boxB, boxT;
              kd_search_boxes(kp.orth, __ldg(kp.posPhi + m), mid.r, mid.z, dir, boxB, boxT);
              const bool searchable = !kd_degenerate(boxB) && !kd_degenerate(boxT);  // CylindricalSpacePointKDTree.cpp:226,241
              if (!kFill) {
                ++cntMiddles;
                if (searchable) {
                  startWalk(false);
                } else {
                  finishCount();
                }
              } else {
                const unsigned long long slot = p.slotPrefix[w] - p.slotPrefix[p.itemFirst];
                slotLo = (uint32_t)slot;
                recSlot = p.rec + slot;
                keySlot = p.key + slot;
                nTop = 0;
                if (searchable) {
                  startWalk(false);
                  recOut = recSlot + capB;
                  keyOut = keySlot + capB;
                } else {
                  finishFill(0u, 0.f, 0.f);
                }
              }
            }
          }
        }
        continue;
      }
      if (phase == 0) phase = 3;  // nothing left to hand out
    }
    const bool walking = phase == 1 || phase == 2;
    if (__ballot_sync(0xffffffffu, walking) == 0u) break;
    if (!walking) continue;
    const bool bottom = phase == 2;
    // fill pass, list mode: the next word of the count pass's hit list is either a survivor, a page link or the end
    bool listEnd = false;
    uint32_t eo = o;
    bool elemStep = o < oEnd;
    if (kFill && listMode) {
      const uint32_t v = __ldg(kp.hitArena + 32ull * hitPage + hitFill);
      elemStep = false;
      if (hitFill == 31u) {
        hitPage = v;
        hitFill = 0;
        continue;
      }
      ++hitFill;
      if (v == kKdHitEnd) { listEnd = true; } else { eo = v; elemStep = true; }
    }
    if (elemStep) {
      const float2 zr = ldg2(p.pZR + eo);
      bool inside = contained || (kFill && listMode);
      if (!inside) {
        const float phi = __ldg(kp.posPhi + eo);
        inside = (box.mn[0] <= phi) & (phi < box.mx[0]) & (box.mn[1] <= zr.y) & (zr.y < box.mx[1]) &
                 (box.mn[2] <= zr.x) & (zr.x < box.mx[2]);
      }
      float dR, dZ;
      // (deltaR is tested first in the reference, DoubletSeedFinder.cpp:126-130: both are pure rejections)
      const bool pass = inside && doublet_zr_cuts_side(bottom, cfg, mid, zr.x, zr.y, dR, dZ) && !outside_range(dR, dRMin, dRMax);
      if (pass) {
        if (!kFill) {
          ++n;
          hitPut(eo);
        } else {
          const uint32_t o = eo;  // (shadows the walker's position for the stores below)
          const float2 xy = ldg2(p.pXY + o), var = ldg2(p.pVar + o);
          DoubletRec rec;
          if (doublet_finish_side(bottom, cfg, mid, dR, dZ, xy.x, xy.y, zr.y, var.x, var.y, nullptr, nullptr, 0, rec, true)) {
            float4* dst = reinterpret_cast<float4*>(recOut + n);
            dst[0] = make_float4(__uint_as_float(o), rec.cotTheta, rec.iDeltaR, rec.er);
            dst[1] = make_float4(rec.u, rec.v, rec.xNew, rec.yNew);
            keyOut[n] = rec.cotTheta;
            mn = fminf(mn, rec.cotTheta);
            mx = fmaxf(mx, rec.cotTheta);
            ++n;
          }
        }
      }
      if (!(kFill && listMode)) ++o;
    } else if (!listEnd && id != kKdEnd && !(kFill && listMode)) {
      const float4* nd = reinterpret_cast<const float4*>(nodes + id);
      const float4 a = __ldg(nd), b = __ldg(nd + 1);
      const uint4 c = __ldg(reinterpret_cast<const uint4*>(nd + 2));
      // a = {mnPhi, mnR, mnZ, mxPhi}, b = {mxR, mxZ, begin, end}, c = {rope, internal, lhs, -}
      const bool overlaps = (a.x < box.mx[0]) & (box.mn[0] < a.w) & (a.y < box.mx[1]) & (box.mn[1] < b.x) &
                            (a.z < box.mx[2]) & (box.mn[2] < b.y);
      const bool cont = (box.mn[0] <= a.x) & (box.mx[0] >= a.w) & (box.mn[1] <= a.y) & (box.mx[1] >= b.x) &
                        (box.mn[2] <= a.z) & (box.mx[2] >= b.y);
      // Descend only into nodes of more than kKdScanBelow elements: a smaller subtree is scanned element by
      // element instead (every reported element passes the exact `contains` test of the reference's leaves, so
      // the reported set and its order are the same; the walk trades node visits -- divergent, dependent loads --
      // for element tests at consecutive addresses).
      const uint32_t e0 = __float_as_uint(b.z), e1 = __float_as_uint(b.w);
      const bool descend = overlaps & (c.y != 0u) & !cont & (e1 - e0 > (uint32_t)kKdScanBelow);  // left child; the right one is the left one's rope
      id = descend ? c.z : c.x;
      if (overlaps & !descend) {
        o = e0;  // packed positions (batch-wide)
        oEnd = e1;
        contained = cont;
      }
    } else {  // the walk is over
      if (!kFill) endList(bottom);
      if (!bottom) {
        if (!kFill) {
          capT = n;
          if (capT != 0u) startWalk(true); else finishCount();
        } else {
          nTop = n; mnT = mn; mxT = mx;
          bool go = nTop != 0u;
          // BroadTripletSeedFilter.cpp:63-94 (sufficientTopDoublets; it subsumes the candidate-count test of
          // CylindricalSpacePointKDTree.cpp:245-246: doublets <= candidates)
          if (go && p.conf) go = !(nTop < conf_n_top(conf_range(cfg, mid.z), mid.r));
          if (go) {
            startWalk(true);
            recOut = recSlot;
            keyOut = keySlot;
          } else {
            nTop = 0;
            finishFill(0u, 0.f, 0.f);
          }
        }
      } else {
        if (!kFill) {
          capB = n;
          finishCount();
        } else {
          finishFill(n, mn, mx);
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

// Fill pass of the items whose two tree walks left hit lists (KdDoubletParams::hitHead): one WARP per item reads a
// page of the list with one coalesced request (31 positions + the link), gathers the space points, finishes the
// doublets and writes the survivors compacted in list order -- the arena slot, header, carve-up and class list are
// those of k_doublets_kd<true>, which keeps the items without lists.  ticket: a work counter of its own.
constexpr int kKdListWarps = 8;
__global__ void __launch_bounds__(kKdListWarps * 32) k_doublets_kd_lists(const __grid_constant__ KdDoubletParams kp, uint32_t* ticket) {
  const DoubletParams& p = kp.d;
  const DeviceConfig& cfg = p.cfg;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t ltMask = (1u << lane) - 1u;
  const uint32_t nItems = p.itemEnd - p.itemFirst;
  unsigned long long cntB = 0, cntT = 0;  // lane 0
  auto side = [&](bool bottom, const MiddleSp& mid, uint32_t page, DoubletRecord* recOut, float* keyOut, float& cotMin, float& cotMax) -> uint32_t {
    uint32_t n = 0;
    float mn = 3.0e38f, mx = -3.0e38f;
    for (;;) {
      const uint32_t v = __ldg(kp.hitArena + 32ull * page + lane);
      const uint32_t link = __shfl_sync(0xffffffffu, v, 31);
      const uint32_t endMask = __ballot_sync(0xffffffffu, lane < 31u && v == kKdHitEnd);
      const uint32_t nValid = endMask != 0u ? (uint32_t)(__ffs(endMask) - 1) : 31u;
      bool ok = false;
      DoubletRec rec;
      if (lane < nValid) {
        const float2 zr = ldg2(p.pZR + v), xy = ldg2(p.pXY + v), var = ldg2(p.pVar + v);
        float dR, dZ;
        doublet_zr_cuts_side(bottom, cfg, mid, zr.x, zr.y, dR, dZ);  // (a hit passed them in the count pass: dR, dZ)
        ok = doublet_finish_side(bottom, cfg, mid, dR, dZ, xy.x, xy.y, zr.y, var.x, var.y, nullptr, nullptr, 0, rec, true);
      }
      const uint32_t mask = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const uint32_t d = n + (uint32_t)__popc(mask & ltMask);
        float4* dst = reinterpret_cast<float4*>(recOut + d);
        dst[0] = make_float4(__uint_as_float(v), rec.cotTheta, rec.iDeltaR, rec.er);
        dst[1] = make_float4(rec.u, rec.v, rec.xNew, rec.yNew);
        keyOut[d] = rec.cotTheta;
        mn = fminf(mn, rec.cotTheta);
        mx = fmaxf(mx, rec.cotTheta);
      }
      n += (uint32_t)__popc(mask);
      if (endMask != 0u) break;
      page = link;
    }
    for (int d = 16; d > 0; d >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    cotMin = mn;
    cotMax = mx;
    return n;
  };
  for (;;) {
    uint32_t it = 0;
    if (lane == 0) it = atomicAdd(ticket, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= nItems) break;
    const uint32_t w = p.itemFirst + it;
    const uint32_t capT = __ldg(p.capT + w), capB = __ldg(p.capB + w);
    if (capT == 0u || capB == 0u) continue;  // (k_doublets_kd<true> writes the empty header)
    const uint32_t headT = kp.hitHead[2ull * w], headB = kp.hitHead[2ull * w + 1u];
    if (headT == kKdHitNone || headB == kKdHitNone) continue;  // (k_doublets_kd<true> walks the tree again)
    const uint32_t m = __ldg(p.workPos + w);
    MiddleSp mid;
    {
      const float2 mxy = ldg2(p.pXY + m), mzr = ldg2(p.pZR + m), mvar = ldg2(p.pVar + m);
      mid.x = mxy.x; mid.y = mxy.y; mid.z = mzr.x; mid.r = mzr.y; mid.varZ = mvar.x; mid.varR = mvar.y;
      middle_info(mid);
    }
    const unsigned long long slot = p.slotPrefix[w] - p.slotPrefix[p.itemFirst];
    DoubletRecord* recSlot = p.rec + slot;
    float* keySlot = p.key + slot;
    float mnT, mxT, mnB = 0.f, mxB = 0.f;
    uint32_t nTop = side(false, mid, headT, recSlot + capB, keySlot + capB, mnT, mxT);
    bool go = nTop != 0u;
    if (go && p.conf) go = !(nTop < conf_n_top(conf_range(cfg, mid.z), mid.r));  // BroadTripletSeedFilter.cpp:63-94
    uint32_t nB = 0;
    if (go) nB = side(true, mid, headB, recSlot, keySlot, mnB, mxB);
    go = go && nB != 0u;
    if (lane == 0) {  // header, carve-up, class list: as k_doublets_kd<true>::finishFill
      MiddleHeader h{};
      h.capB = capB;
      h.offset = (uint32_t)slot;
      if (go) {
        h.nB = nB; h.nT = nTop;
        h.cotMinB = float_to_ordered(mnB); h.cotMaxB = float_to_ordered(mxB);
        h.cotMinT = float_to_ordered(mnT); h.cotMaxT = float_to_ordered(mxT);
        const SeedCarve cv = seed_carve(nB, nTop);
        p.carve[w] = cv;
        int c = 0;
        while (c < kSpillClass && cv.minBytes > p.classBytes[c]) ++c;
        p.classList[(size_t)c * p.classStride + atomicAdd(p.classCount + c, 1u)] = w;
        cntB += nB;
        cntT += nTop;
      } else {
        p.slotCount[w] = 0;
      }
      p.hdr[w] = h;
    }
  }
  if (lane == 0) {
    if (cntB != 0ull) atomicAdd(p.counters + kCntBottomDoublets, cntB);
    if (cntT != 0ull) atomicAdd(p.counters + kCntTopDoublets, cntT);
  }
}

}  // namespace B200SEED_NS
