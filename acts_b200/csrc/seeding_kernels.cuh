// seeding_kernels.cuh -- CUDA kernels of the B200 seeding engine (sm_100a).
//
// Stage map (reference lines they replace; paths relative to the ACTS tree):
//   k_bin_count / k_scan / k_scatter / k_sort_bins
//        grid fill, per-bin r sort, packed SoA copy
//        GridTripletSeedingAlgorithm.cpp:208-253, SpacePointGridBase.hpp:67-87
//   k_middle_ranges / k_fill_work
//        BinnedGroup iteration + middle r range, TripletSeeder.cpp:183-195,
//        GridTripletSeedingAlgorithm.cpp:346-371,404-421
//   k_seed_middles
//        doublets, cotTheta sort, triplets, filter, per-middle selection
//        TripletSeeder.cpp:21-107,138-181, DoubletSeedFinder.cpp:41-273,
//        TripletSeedFinder.cpp:34-162, BroadTripletSeedFilter.cpp:96-393,
//        CandidatesForMiddleSp.cpp:44-93
//   k_tile_sums / k_compact_seeds / k_event_offsets
//        SeedContainer fill + index remap, GridTripletSeedingAlgorithm.cpp:394-398
//
// All events of a batch are processed by the same launches: space points are
// concatenated, bins are (event, bin) pairs and the middle work list spans the
// whole batch, so the grid is always sized for the machine, not for one event.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "seed_math.h"

namespace b200seed {

constexpr uint32_t kInvalidBin = 0xFFFFFFFFu;
constexpr int kSeedThreads = 512;
constexpr int kSortThreads = 256;
constexpr int kScanThreads = 1024;
constexpr int kTile = 2048;  // elements per block in the tiled scan

enum CounterSlot : int {
  kCntInGrid = 0,
  kCntMiddles,
  kCntBottomDoublets,
  kCntTopDoublets,
  kCntTripletTests,
  kCntCandidates,
  kCntSeeds,
  kCntTieMiddles,
  kCntSlots
};

enum StatusBits : int {
  kStatusOverflowDoublets = 1,
  kStatusOverflowPool = 2,
  kStatusBinTooLarge = 4
};

struct GridParams {
  DeviceConfig cfg;
  uint32_t nEvents, nTotal, nBins;
  const uint32_t* spOffsets;  // [nEvents + 1]
  const float *x, *y, *z, *r, *varZ, *varR;
  const float* phi;  // optional precomputed phi (NULL: replay atan2f on device)
  uint32_t* binOf;     // [nTotal]
  uint32_t* binCount;  // [nEvents * nBins]
  uint32_t* binStart;  // [nEvents * nBins + 1]
  uint32_t* binCursor; // [nEvents * nBins]
  uint32_t* tmpIdx;    // [nTotal] event-local index, unsorted inside a bin
  // packed copy (reference coreSpacePoints)
  uint32_t* pIdx;
  float2 *pXY, *pZR, *pVar;
  unsigned long long* sortScratch;  // [2 * nTotal] for bins larger than smem
  uint32_t sortSmemCap;             // elements that fit the sort kernel's smem
  int exactTies;                    // replay libstdc++ std::sort tie order
  int* status;
  unsigned long long* counters;
};

struct WorkParams {
  DeviceConfig cfg;
  uint32_t nEvents, nBins, nNav;
  const uint32_t* binStart;
  const float2* pZR;
  const uint32_t* navBins;
  uint32_t* midLo;      // [nEvents * nNav]
  uint32_t* midCount;   // [nEvents * nNav]
  uint32_t* workStart;  // [nEvents * nNav + 1]
  uint32_t* workPos;    // [nTotal]
  uint32_t* workEG;     // [nTotal]
};

struct SeedParams {
  DeviceConfig cfg;
  const float2 *pXY, *pZR, *pVar;
  const uint32_t* binStart;
  const uint32_t *navBins, *botOffsets, *botBins, *topOffsets, *topBins;
  const uint32_t *workPos, *workEG;
  const uint32_t* nWorkPtr;   // number of items of THIS launch
  const uint32_t* workList;   // NULL: items are 0..n-1; else indices into workPos/workEG
  uint32_t* overflowList;     // middles that did not fit this launch's scratch (NULL: error)
  uint32_t* overflowCount;
  uint32_t nNav, nBins;
  const float *zWinLo, *zWinHi;
  int nZWin;
  uint32_t* workCounter;
  uint32_t *slotB, *slotM, *slotT;
  float *slotQ, *slotZ;
  uint32_t* slotCount;
  uint32_t seedsPerMiddle;
  uint32_t capB, capT, capPool, nBuckets;
  int exactTies;  // replay libstdc++ std::sort inside groups of equal cotTheta
  unsigned long long* counters;
  int* status;
};

struct CompactParams {
  const uint32_t* nWorkPtr;
  const uint32_t* slotCount;
  uint32_t* tileSums;    // [nTiles + 1]
  uint32_t* tilePrefix;  // [nTiles + 1]
  const uint32_t *slotB, *slotM, *slotT;
  const float *slotQ, *slotZ;
  uint32_t seedsPerMiddle;
  const uint32_t* pIdx;
  uint32_t *outB, *outM, *outT;
  float *outQ, *outZ;
  unsigned long long outCapacity;
  unsigned long long* seedOffsets;  // [nEvents + 1]
  const uint32_t* workStart;
  uint32_t* seedStart;  // [nWork + 1] exclusive scan of slotCount
  uint32_t nEvents, nNav;
  unsigned long long* counters;
};

// ---------------------------------------------------------------------------
// block-level primitives
// ---------------------------------------------------------------------------
struct OpSum {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a + b; }
};
struct OpMax {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// Exclusive scan over the threads of a block (identity 0).  `scratch` holds
// blockDim.x / 32 + 1 words of shared memory.  Contains two __syncthreads().
template <typename Op>
__device__ __forceinline__ uint32_t block_scan_exclusive(uint32_t v, uint32_t* scratch, uint32_t& total, Op op) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl = op(incl, o);
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t t = lane < nWarps ? scratch[lane] : 0u;
    uint32_t ti = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, ti, d);
      if (lane >= (uint32_t)d) ti = op(ti, o);
    }
    const uint32_t excl = __shfl_up_sync(0xffffffffu, ti, 1);
    if (lane < nWarps) scratch[lane] = lane == 0 ? 0u : excl;
    if (lane == 31) scratch[32] = ti;  // nWarps <= 32: lane 31 holds the block total
  }
  __syncthreads();
  uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0u;
  total = scratch[32];
  const uint32_t out = op(scratch[warp], excl);
  __syncthreads();  // scratch may be reused right after the call
  return out;
}

// Ordered warp-aggregated append: lanes with `pred` get consecutive slots.
// Must be called by all 32 lanes of the warp.
__device__ __forceinline__ uint32_t warp_append(uint32_t* counter, bool pred) {
  const uint32_t mask = __ballot_sync(0xffffffffu, pred);
  if (mask == 0u) return 0u;
  const uint32_t lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

template <typename Pred>
__device__ __forceinline__ uint32_t first_true(uint32_t lo, uint32_t hi, Pred pred) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (pred(mid)) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ float2 ldg2(const float2* p) { return __ldg(p); }

// ---------------------------------------------------------------------------
// Grid stage
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bin_count(const __grid_constant__ GridParams p) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.nTotal; i += stride) {
    // event of this space point: last offset <= i
    uint32_t lo = 0, hi = p.nEvents;
    while (lo < hi) {
      const uint32_t mid = (lo + hi + 1) >> 1;
      if (__ldg(p.spOffsets + mid) <= i) lo = mid; else hi = mid - 1;
    }
    const float x = __ldg(p.x + i), y = __ldg(p.y + i), z = __ldg(p.z + i), r = __ldg(p.r + i);
    int32_t bin = -1;
    if (!(p.cfg.useExtraCuts && !itk_sp_select(r, z))) {
      const float phi = p.phi != nullptr ? __ldg(p.phi + i) : glibc_atan2f(y, x);
      bin = grid_bin_index(p.cfg, phi, z, r);
    }
    uint32_t gb = kInvalidBin;
    if (bin >= 0) {
      gb = lo * p.nBins + (uint32_t)bin;
      atomicAdd(p.binCount + gb, 1u);
    }
    p.binOf[i] = gb;
  }
}

// out[0..n] = exclusive scan of in[0..n) (out[n] = total), single block.
__global__ void __launch_bounds__(kScanThreads) k_scan(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n) {
  __shared__ uint32_t scratch[34];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0u;
    uint32_t total;
    const uint32_t excl = block_scan_exclusive(v, scratch, total, OpSum());
    const uint32_t c = carry;
    if (i < n) out[i] = c + excl;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

__global__ void __launch_bounds__(256) k_scatter(const __grid_constant__ GridParams p) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.nTotal; i += stride) {
    const uint32_t gb = p.binOf[i];
    if (gb == kInvalidBin) continue;
    const uint32_t e = gb / p.nBins;
    const uint32_t pos = p.binStart[gb] + atomicAdd(p.binCursor + gb, 1u);
    p.tmpIdx[pos] = i - __ldg(p.spOffsets + e);
  }
}

// In-place bitonic sort of n 64-bit keys stored in a buffer of `padded` (power
// of two, >= n) entries; entries [n, padded) must hold ~0ull.
__device__ __forceinline__ void block_bitonic_sort(unsigned long long* keys, uint32_t padded) {
  for (uint32_t k = 2; k <= padded; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < padded; i += blockDim.x) {
        const uint32_t ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

struct RItem {
  float key;
  uint32_t val;
};
__device__ __forceinline__ bool ritem_less(const RItem& a, const RItem& b) { return a.key < b.key; }

// One block per (event, bin): order the bin by r and write the packed copy.
//   exactTies = 1: entries are first put in insertion order (ascending original
//     index, the order grid.insert produced, .cpp:211-221) and then sorted with
//     the replayed libstdc++ std::sort (.cpp:223-228) so that equal radii end up
//     in the reference's order.
//   exactTies = 0: canonical (r, original index) order.
__global__ void __launch_bounds__(kSortThreads) k_sort_bins(const __grid_constant__ GridParams p) {
  extern __shared__ unsigned long long smemKeys[];
  const uint32_t gb = blockIdx.x;
  const uint32_t b0 = p.binStart[gb], b1 = p.binStart[gb + 1];
  const uint32_t n = b1 - b0;
  if (n == 0) return;
  const uint32_t e = gb / p.nBins;
  const uint32_t evBase = __ldg(p.spOffsets + e);
  uint32_t padded = 1;
  while (padded < n) padded <<= 1;
  unsigned long long* keys = padded <= p.sortSmemCap ? smemKeys : p.sortScratch + 2ull * b0;
  for (uint32_t i = threadIdx.x; i < padded; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < n) {
      const uint32_t idx = p.tmpIdx[b0 + i];
      const float r = __ldg(p.r + evBase + idx);
      const uint32_t rb = (r == 0.0f) ? 0u : __float_as_uint(r);  // r >= 0 inside the grid
      k = p.exactTies ? (((unsigned long long)idx << 32) | rb) : (((unsigned long long)rb << 32) | idx);
    }
    keys[i] = k;
  }
  __syncthreads();
  block_bitonic_sort(keys, padded);
  if (p.exactTies) {
    RItem* items = reinterpret_cast<RItem*>(keys);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long k = keys[i];
      RItem it;
      it.key = __uint_as_float((uint32_t)(k & 0xffffffffu));
      it.val = (uint32_t)(k >> 32);
      items[i] = it;  // same 8 bytes the key came from
    }
    __syncthreads();
    if (threadIdx.x == 0) std_sort(items, (int)n, ritem_less);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t idx = items[i].val;
      const uint32_t src = evBase + idx;
      p.pIdx[b0 + i] = idx;
      p.pXY[b0 + i] = make_float2(__ldg(p.x + src), __ldg(p.y + src));
      p.pZR[b0 + i] = make_float2(__ldg(p.z + src), __ldg(p.r + src));
      p.pVar[b0 + i] = make_float2(__ldg(p.varZ + src), __ldg(p.varR + src));
    }
  } else {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t idx = (uint32_t)(keys[i] & 0xffffffffu);
      const uint32_t src = evBase + idx;
      p.pIdx[b0 + i] = idx;
      p.pXY[b0 + i] = make_float2(__ldg(p.x + src), __ldg(p.y + src));
      p.pZR[b0 + i] = make_float2(__ldg(p.z + src), __ldg(p.r + src));
      p.pVar[b0 + i] = make_float2(__ldg(p.varZ + src), __ldg(p.varR + src));
    }
  }
}

// ---------------------------------------------------------------------------
// Middle work list
// ---------------------------------------------------------------------------
// GridTripletSeedingAlgorithm.cpp:404-421
__device__ __forceinline__ float2 radius_range_for_middle(const DeviceConfig& c, float zFirst, float2 variableRange) {
  if (c.useVariableMiddleSPRange) return variableRange;
  if (c.nRRangeMiddleSP == 0) return make_float2(c.rMinMiddle, c.rMaxMiddle);
  int lo = 0, hi = c.nZBinEdgesF;  // std::lower_bound: first edge >= zFirst
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (c.zBinEdgesF[mid] < zFirst) lo = mid + 1; else hi = mid;
  }
  int zBin = lo;
  if (zBin != 0) --zBin;
  return make_float2(c.rRangeMiddleSP[2 * zBin], c.rRangeMiddleSP[2 * zBin + 1]);
}

// One block per event.
__global__ void __launch_bounds__(256) k_middle_ranges(const __grid_constant__ WorkParams p) {
  __shared__ float sMin[256], sMax[256];
  const uint32_t e = blockIdx.x;
  const uint32_t* bs = p.binStart + (size_t)e * p.nBins;
  float2 variable = make_float2(0.f, 0.f);
  if (p.cfg.useVariableMiddleSPRange) {
    // .cpp:257-270,327-330
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    for (uint32_t b = threadIdx.x; b < p.nBins; b += blockDim.x) {
      const uint32_t b0 = bs[b], b1 = bs[b + 1];
      if (b0 == b1) continue;
      mn = fminf(mn, p.pZR[b0].y);
      mx = fmaxf(mx, p.pZR[b1 - 1].y);
    }
    sMin[threadIdx.x] = mn;
    sMax[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) {
        sMin[threadIdx.x] = fminf(sMin[threadIdx.x], sMin[threadIdx.x + s]);
        sMax[threadIdx.x] = fmaxf(sMax[threadIdx.x], sMax[threadIdx.x + s]);
      }
      __syncthreads();
    }
    variable.x = fadd(fmul(floorf(fdiv(sMin[0], 2.0f)), 2.0f), p.cfg.deltaRMiddleMinSPRange);
    variable.y = fsub(fmul(floorf(fdiv(sMax[0], 2.0f)), 2.0f), p.cfg.deltaRMiddleMaxSPRange);
  }
  for (uint32_t g = threadIdx.x; g < p.nNav; g += blockDim.x) {
    const uint32_t bin = p.navBins[g];
    const uint32_t b0 = bs[bin], b1 = bs[bin + 1];
    uint32_t lo = b0, hi = b0;
    if (b0 != b1) {
      const float2 range = radius_range_for_middle(p.cfg, p.pZR[b0].x, variable);
      // TripletSeeder.cpp:183-195: skip r < min, stop at r > max (bin is r-sorted)
      lo = first_true(b0, b1, [&](uint32_t i) { return !(p.pZR[i].y < range.x); });
      hi = first_true(lo, b1, [&](uint32_t i) { return p.pZR[i].y > range.y; });
    }
    p.midLo[(size_t)e * p.nNav + g] = lo;
    p.midCount[(size_t)e * p.nNav + g] = hi - lo;
  }
}

// One warp per (event, navigation entry): materialise the middle work list in
// the reference's processing order.
__global__ void __launch_bounds__(256) k_fill_work(const __grid_constant__ WorkParams p) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= p.nEvents * p.nNav) return;
  const uint32_t start = p.workStart[warp], count = p.midCount[warp], lo = p.midLo[warp];
  for (uint32_t i = lane; i < count; i += 32) {
    p.workPos[start + i] = lo + i;
    p.workEG[start + i] = warp;
  }
}

// ---------------------------------------------------------------------------
// Seeding kernel: one block per middle space point (persistent blocks pulling
// work items from an atomic counter).
// ---------------------------------------------------------------------------
struct Cand {
  float curv;
  float impactOrWeight;
  float topR;
  uint32_t tOwner;  // sorted top rank | owner thread << 16
};
__device__ __forceinline__ bool cand_less(const Cand& a, const Cand& b) { return a.curv < b.curv; }

struct StoredSeed {
  uint32_t bottomPos, topPos;
  float weight, zOrigin;
};

struct SeedShared {
  MiddleSp mid;
  uint32_t w, m, eg;
  uint32_t nB, nT, poolCount;
  uint32_t tie, tieTmp, bad;
  uint32_t carry;
  uint32_t nBotWin, nTopWin;
  uint32_t winBs[kMaxNeighborBins], winBe[kMaxNeighborBins], winBp[kMaxNeighborBins + 1];
  uint32_t winTs[kMaxNeighborBins], winTe[kMaxNeighborBins], winTp[kMaxNeighborBins + 1];
  uint32_t scratch[34];
  WeightIndex heap[kMaxHeap];
  StoredSeed storage[kMaxHeap];
  int heapSize;
  float heapMin;
  unsigned long long cnt[kCntSlots];
};

__device__ __forceinline__ uint32_t seq_to_pos(uint32_t seq, const uint32_t* prefix, const uint32_t* start, uint32_t nWin) {
  uint32_t k = 0;
  while (k + 1 < nWin && prefix[k + 1] <= seq) ++k;
  return start[k] + (seq - prefix[k]);
}

// Shared-memory carve-up.  Persistent arrays first, then one arena whose
// content changes with the phase of a middle:
//   phases 1-2 : tCot, tSeq, tSorted (tops before their records exist)
//   tie replay : W + two u16 arrays -- for the tops inside the (not yet written)
//                sorted-top arrays, for the bottoms inside the arena
//   phase 3    : pool (+ links) and pool2
struct SeedSmem {
  SeedShared* sh;
  float* bCot; uint32_t* bSeq; uint16_t* bSorted;       // bottoms, unsorted + rank -> index
  float *sCot, *sIDR, *sEr, *sU, *sV; uint32_t* sPos;   // tops in sorted order
  uint32_t* buckets;                                    // [nBuckets + 1]
  unsigned char* arena;
  float* tCot; uint32_t* tSeq; uint16_t* tSorted;       // arena, phases 1-2
  Cand* pool; uint16_t* poolNext; Cand* pool2;          // arena, phase 3
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

__host__ __device__ inline size_t seed_arena_bytes(uint32_t capB, uint32_t capT, uint32_t capPool) {
  const size_t tops = align16(4ull * capT) * 2 + align16(2ull * capT);
  const size_t pools = align16(sizeof(Cand) * (size_t)capPool) * 2 + align16(2ull * capPool);
  const size_t tieB = align16(8ull * capB) + align16(2ull * capB) * 2;  // W + seqSorted + grpOf (bottoms)
  size_t a = tops > pools ? tops : pools;
  return a > tieB ? a : tieB;
}

__host__ __device__ inline size_t seed_smem_bytes(uint32_t capB, uint32_t capT, uint32_t capPool, uint32_t nBuckets) {
  size_t s = align16(sizeof(SeedShared));
  s += align16(4ull * capB) * 2 + align16(2ull * capB);
  s += align16(4ull * capT) * 6;
  s += align16(4ull * (nBuckets + 1));
  s += seed_arena_bytes(capB, capT, capPool);
  return s;
}

__device__ __forceinline__ SeedSmem carve_seed_smem(unsigned char* base, uint32_t capB, uint32_t capT, uint32_t capPool, uint32_t nBuckets) {
  SeedSmem s;
  unsigned char* q = base;
  s.sh = reinterpret_cast<SeedShared*>(q); q += align16(sizeof(SeedShared));
  s.bCot = reinterpret_cast<float*>(q); q += align16(4ull * capB);
  s.bSeq = reinterpret_cast<uint32_t*>(q); q += align16(4ull * capB);
  s.bSorted = reinterpret_cast<uint16_t*>(q); q += align16(2ull * capB);
  s.sCot = reinterpret_cast<float*>(q); q += align16(4ull * capT);
  s.sIDR = reinterpret_cast<float*>(q); q += align16(4ull * capT);
  s.sEr = reinterpret_cast<float*>(q); q += align16(4ull * capT);
  s.sU = reinterpret_cast<float*>(q); q += align16(4ull * capT);
  s.sV = reinterpret_cast<float*>(q); q += align16(4ull * capT);
  s.sPos = reinterpret_cast<uint32_t*>(q); q += align16(4ull * capT);
  s.buckets = reinterpret_cast<uint32_t*>(q); q += align16(4ull * (nBuckets + 1));
  s.arena = q;
  s.tCot = reinterpret_cast<float*>(q); q += align16(4ull * capT);
  s.tSeq = reinterpret_cast<uint32_t*>(q); q += align16(4ull * capT);
  s.tSorted = reinterpret_cast<uint16_t*>(q);
  q = s.arena;
  s.pool = reinterpret_cast<Cand*>(q); q += align16(sizeof(Cand) * (size_t)capPool);
  s.pool2 = reinterpret_cast<Cand*>(q); q += align16(sizeof(Cand) * (size_t)capPool);
  s.poolNext = reinterpret_cast<uint16_t*>(q);
  return s;
}

// Doublet search for one side (DoubletSeedFinder.cpp:41-273): every thread
// strides over the r windows, survivors are appended as (cotTheta, seq).
template <bool kBottom>
__device__ __forceinline__ void find_doublets(const SeedParams& p, SeedShared& sh, uint32_t nWin, const uint32_t* winS,
                                              const uint32_t* winE, const uint32_t* winP, float* cotOut, uint32_t* seqOut,
                                              uint32_t* counter, uint32_t cap) {
  const MiddleSp mid = sh.mid;
  for (uint32_t k = 0; k < nWin; ++k) {
    const uint32_t s = winS[k], e = winE[k], pre = winP[k];
    for (uint32_t base = s; base < e; base += blockDim.x) {
      const uint32_t o = base + threadIdx.x;
      bool pass = false;
      DoubletRec rec;
      if (o < e) {
        const float2 zr = ldg2(p.pZR + o);
        float dR, dZ;
        if (doublet_zr_cuts<kBottom>(p.cfg, mid, zr.x, zr.y, dR, dZ)) {
          const float2 xy = ldg2(p.pXY + o);
          const float2 var = ldg2(p.pVar + o);
          pass = doublet_finish<kBottom>(p.cfg, mid, dR, dZ, xy.x, xy.y, zr.y, var.x, var.y, p.zWinLo, p.zWinHi, p.nZWin, rec);
        }
      }
      const uint32_t slot = warp_append(counter, pass);
      if (pass && slot < cap) {
        cotOut[slot] = rec.cotTheta;
        seqOut[slot] = pre + (o - s);
      }
    }
  }
}

// Keys of the in-block bucket sort.  bucket() must be monotone in the order
// and after(u, v) tells whether element u sorts strictly after element v.
struct CotKey {  // order by (cotTheta, emission index)
  const float* cot;
  const uint32_t* seq;
  float cotMax, scale;
  int nBk;
  __device__ __forceinline__ int bucket(uint32_t i) const { return cot_bucket(cot[i], cotMax, scale, nBk); }
  __device__ __forceinline__ bool after(uint32_t u, uint32_t v) const {
    const float cu = cot[u], cv = cot[v];
    return cu > cv || (cu == cv && seq[u] > seq[v]);
  }
};
struct SeqKey {  // order by emission index (unique)
  const uint32_t* seq;
  uint32_t total;
  int nBk;
  __device__ __forceinline__ int bucket(uint32_t i) const {
    return (int)(((unsigned long long)seq[i] * (unsigned long long)nBk) / (unsigned long long)total);
  }
  __device__ __forceinline__ bool after(uint32_t u, uint32_t v) const { return seq[u] > seq[v]; }
};

// Bucket sort of n elements: sorted[rank] = element index.
template <typename Key>
__device__ __forceinline__ void block_bucket_sort(uint32_t n, const Key key, uint16_t* sorted, uint32_t* buckets,
                                                  uint32_t nBk, uint32_t* scratch) {
  for (uint32_t i = threadIdx.x; i <= nBk; i += blockDim.x) buckets[i] = 0;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(buckets + key.bucket(i), 1u);
  __syncthreads();
  // exclusive scan of the bucket counts (in place), blockDim.x buckets per pass
  __shared__ uint32_t runCarry;
  if (threadIdx.x == 0) runCarry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nBk; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nBk ? buckets[i] : 0u;
    uint32_t total;
    const uint32_t excl = block_scan_exclusive(v, scratch, total, OpSum());
    const uint32_t c = runCarry;
    if (i < nBk) buckets[i] = c + excl;
    __syncthreads();
    if (threadIdx.x == 0) runCarry = c + total;
    __syncthreads();
  }
  // scatter: after this loop buckets[b] is the END of bucket b
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    sorted[atomicAdd(buckets + key.bucket(i), 1u)] = (uint16_t)i;
  }
  __syncthreads();
  // order every bucket; buckets hold O(1) entries
  for (uint32_t b = threadIdx.x; b < nBk; b += blockDim.x) {
    const uint32_t s = b == 0 ? 0u : buckets[b - 1], e = buckets[b];
    for (uint32_t i = s + 1; i < e; ++i) {
      const uint16_t v = sorted[i];
      uint32_t j = i;
      while (j > s && key.after(sorted[j - 1], v)) {
        sorted[j] = sorted[j - 1];
        --j;
      }
      sorted[j] = v;
    }
  }
  __syncthreads();
}

// true (block-uniform) when two neighbours of the sorted list have equal keys
__device__ __forceinline__ bool block_has_ties(uint32_t n, const float* cot, const uint16_t* sorted, uint32_t* flag) {
  if (threadIdx.x == 0) *flag = 0;
  __syncthreads();
  bool t = false;
  for (uint32_t i = threadIdx.x + 1; i < n; i += blockDim.x) t |= cot[sorted[i]] == cot[sorted[i - 1]];
  if (t) *flag = 1;
  __syncthreads();
  const bool r = *flag != 0;
  __syncthreads();
  return r;
}

struct TieItem {
  float key;
  uint32_t val;  // element index | tie group (first canonical rank) << 16; group 0xFFFF = unique key
};
__device__ __forceinline__ bool tie_less(const TieItem& a, const TieItem& b) { return a.key < b.key; }
__device__ __forceinline__ bool tie_flagged(const TieItem& a) { return (a.val >> 16) != 0xFFFFu; }

// Exact order inside groups of equal cotTheta: the reference sorts the doublets
// with the unstable std::ranges::sort (DoubletSeedFinder.hpp:94-104) starting
// from the emission order.  `sorted` holds the canonical (cot, seq) order; the
// pruned replay of libstdc++'s introsort (seed_math.h) run by one thread on the
// emission-ordered copy W decides which member of a tie group takes which of the
// group's slots.  seqSorted / grpOf are n-entry u16 scratch arrays.
__device__ __forceinline__ void block_fix_ties(uint32_t n, const float* cot, const uint32_t* seq, uint32_t totalCand,
                                               uint16_t* sorted, TieItem* W, uint16_t* seqSorted, uint16_t* grpOf,
                                               uint32_t* buckets, uint32_t nBk, uint32_t* scratch) {
  SeqKey sk{seq, totalCand > 0 ? totalCand : 1u, (int)nBk};
  block_bucket_sort(n, sk, seqSorted, buckets, nBk, scratch);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint16_t e = sorted[i];
    const float c = cot[e];
    const bool tp = i > 0 && cot[sorted[i - 1]] == c;
    const bool tn = i + 1 < n && cot[sorted[i + 1]] == c;
    uint16_t g = 0xFFFFu;
    if (tp || tn) {
      uint32_t q = i;
      while (q > 0 && cot[sorted[q - 1]] == c) --q;
      g = (uint16_t)q;
    }
    grpOf[e] = g;
  }
  __syncthreads();
  for (uint32_t r = threadIdx.x; r < n; r += blockDim.x) {
    const uint16_t e = seqSorted[r];
    TieItem it;
    it.key = cot[e];
    it.val = (uint32_t)e | ((uint32_t)grpOf[e] << 16);
    W[r] = it;
  }
  __syncthreads();
  if (threadIdx.x == 0) std_sort_replay_ties(W, (int)n, tie_less, tie_flagged);
  __syncthreads();
  // member k (in W order) of group q goes to canonical slot q + k
  for (uint32_t r = threadIdx.x; r < n; r += blockDim.x) {
    const TieItem it = W[r];
    const uint32_t g = it.val >> 16;
    if (g == 0xFFFFu) continue;
    uint32_t before = 0;
    for (uint32_t q = 0; q < r; ++q) before += (W[q].val >> 16) == g ? 1u : 0u;
    sorted[g + before] = (uint16_t)(it.val & 0xFFFFu);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSeedThreads, 2) k_seed_middles(const __grid_constant__ SeedParams p) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const SeedSmem S = carve_seed_smem(smemRaw, p.capB, p.capT, p.capPool, p.nBuckets);
  SeedShared& sh = *S.sh;
  const uint32_t tid = threadIdx.x;
  const DeviceConfig& cfg = p.cfg;
  const uint32_t nWork = *p.nWorkPtr;

  if (tid < (uint32_t)kCntSlots) sh.cnt[tid] = 0ull;

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      const uint32_t item = atomicAdd(p.workCounter, 1u);
      sh.w = item < nWork ? (p.workList != nullptr ? p.workList[item] : item) : 0xFFFFFFFFu;
    }
    __syncthreads();
    const uint32_t w = sh.w;
    if (w == 0xFFFFFFFFu) break;

    // ---- phase 0: middle, r windows ------------------------------------
    const uint32_t m = __ldg(p.workPos + w);
    const uint32_t eg = __ldg(p.workEG + w);
    const uint32_t ev = eg / p.nNav, g = eg - ev * p.nNav;
    const uint32_t* bs = p.binStart + (size_t)ev * p.nBins;
    const uint32_t botBeg = __ldg(p.botOffsets + g), nBot = __ldg(p.botOffsets + g + 1) - botBeg;
    const uint32_t topBeg = __ldg(p.topOffsets + g), nTop = __ldg(p.topOffsets + g + 1) - topBeg;
    const float2 mzr = ldg2(p.pZR + m);
    const float rM = mzr.y;
    if (tid == 0) {
      const float2 mxy = ldg2(p.pXY + m), mvar = ldg2(p.pVar + m);
      MiddleSp mid;
      mid.x = mxy.x; mid.y = mxy.y; mid.z = mzr.x; mid.r = mzr.y; mid.varZ = mvar.x; mid.varR = mvar.y;
      middle_info(mid);
      sh.mid = mid;
      sh.m = m;
      sh.nB = 0; sh.nT = 0; sh.tie = 0; sh.bad = 0; sh.carry = 0; sh.heapSize = 0; sh.poolCount = 0;
      sh.nBotWin = nBot; sh.nTopWin = nTop;
    }
    {
      // first middle space point of the bin (TripletSeeder.cpp:157-181 pre-trim)
      const uint32_t mb0 = bs[__ldg(p.navBins + g)];
      const float firstMiddleR = ldg2(p.pZR + mb0).y;
      if (tid < nBot) {
        const uint32_t bin = __ldg(p.botBins + botBeg + tid);
        const uint32_t b0 = bs[bin], b1 = bs[bin + 1];
        const float trimValue = fsub(firstMiddleR, cfg.dRMaxB);
        const uint32_t trim = first_true(b0, b1, [&](uint32_t i) { return !(ldg2(p.pZR + i).y < trimValue); });
        const uint32_t s = first_true(trim, b1, [&](uint32_t i) { return fsub(rM, ldg2(p.pZR + i).y) <= cfg.dRMaxB; });
        const uint32_t e = first_true(s, b1, [&](uint32_t i) { return fsub(rM, ldg2(p.pZR + i).y) < cfg.dRMinB; });
        sh.winBs[tid] = s; sh.winBe[tid] = e;
      } else if (tid >= (uint32_t)kMaxNeighborBins && tid < (uint32_t)kMaxNeighborBins + nTop) {
        const uint32_t k = tid - (uint32_t)kMaxNeighborBins;
        const uint32_t bin = __ldg(p.topBins + topBeg + k);
        const uint32_t b0 = bs[bin], b1 = bs[bin + 1];
        const float trimValue = fadd(firstMiddleR, cfg.dRMinT);
        const uint32_t trim = first_true(b0, b1, [&](uint32_t i) { return !(ldg2(p.pZR + i).y < trimValue); });
        const uint32_t s = first_true(trim, b1, [&](uint32_t i) { return fsub(ldg2(p.pZR + i).y, rM) >= cfg.dRMinT; });
        const uint32_t e = first_true(s, b1, [&](uint32_t i) { return fsub(ldg2(p.pZR + i).y, rM) > cfg.dRMaxT; });
        sh.winTs[k] = s; sh.winTe[k] = e;
      }
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      for (uint32_t k = 0; k < nBot; ++k) { sh.winBp[k] = acc; acc += sh.winBe[k] - sh.winBs[k]; }
      sh.winBp[nBot] = acc;
      acc = 0;
      for (uint32_t k = 0; k < nTop; ++k) { sh.winTp[k] = acc; acc += sh.winTe[k] - sh.winTs[k]; }
      sh.winTp[nTop] = acc;
      sh.cnt[kCntMiddles] += 1;
    }
    __syncthreads();

    // ---- phase 1: doublets (tops first, TripletSeeder.cpp:52-82) ---------
    find_doublets<false>(p, sh, nTop, sh.winTs, sh.winTe, sh.winTp, S.tCot, S.tSeq, &sh.nT, p.capT);
    __syncthreads();
    const uint32_t nT = sh.nT;
    if (nT == 0) {
      if (tid == 0) p.slotCount[w] = 0;
      continue;
    }
    find_doublets<true>(p, sh, nBot, sh.winBs, sh.winBe, sh.winBp, S.bCot, S.bSeq, &sh.nB, p.capB);
    __syncthreads();
    const uint32_t nB = sh.nB;
    if (nB == 0 || nT > p.capT || nB > p.capB) {
      if (tid == 0) {
        p.slotCount[w] = 0;
        if (nB != 0) {  // does not fit this launch's scratch: hand over to the large-capacity launch
          if (p.overflowList != nullptr) {
            p.overflowList[atomicAdd(p.overflowCount, 1u)] = w;
            sh.cnt[kCntMiddles] -= 1;  // counted again by the launch that completes it
          } else {
            atomicOr(p.status, kStatusOverflowDoublets);
          }
        }
      }
      continue;
    }

    // ---- phase 2: order both lists like DoubletSeedFinder.hpp:94-104 -------
    const float bkScale = (float)p.nBuckets / (2.0f * cfg.cotThetaMax);
    {
      CotKey tk{S.tCot, S.tSeq, cfg.cotThetaMax, bkScale, (int)p.nBuckets};
      block_bucket_sort(nT, tk, S.tSorted, S.buckets, p.nBuckets, sh.scratch);
      const bool tieT = block_has_ties(nT, S.tCot, S.tSorted, &sh.tieTmp);
      if (tieT && tid == 0) sh.tie = 1;
      if (tieT && p.exactTies && nT > 16) {
        // scratch inside the sorted-top arrays, which are written only after this
        block_fix_ties(nT, S.tCot, S.tSeq, sh.winTp[nTop], S.tSorted, reinterpret_cast<TieItem*>(S.sCot),
                       reinterpret_cast<uint16_t*>(S.sEr), reinterpret_cast<uint16_t*>(S.sU), S.buckets,
                       p.nBuckets, sh.scratch);
      }
    }
    // tops: full records in sorted order (recomputed from the space points)
    for (uint32_t t = tid; t < nT; t += blockDim.x) {
      const uint32_t idx = S.tSorted[t];
      const uint32_t pos = seq_to_pos(S.tSeq[idx], sh.winTp, sh.winTs, nTop);
      const float2 zr = ldg2(p.pZR + pos), xy = ldg2(p.pXY + pos), var = ldg2(p.pVar + pos);
      float dR, dZ;
      DoubletRec rec;
      doublet_zr_cuts<false>(cfg, sh.mid, zr.x, zr.y, dR, dZ);
      doublet_finish<false>(cfg, sh.mid, dR, dZ, xy.x, xy.y, zr.y, var.x, var.y, p.zWinLo, p.zWinHi, p.nZWin, rec);
      S.sCot[t] = rec.cotTheta; S.sIDR[t] = rec.iDeltaR; S.sEr[t] = rec.er; S.sU[t] = rec.u; S.sV[t] = rec.v;
      S.sPos[t] = pos;
    }
    __syncthreads();  // tCot / tSeq are dead from here on
    {
      CotKey bk{S.bCot, S.bSeq, cfg.cotThetaMax, bkScale, (int)p.nBuckets};
      block_bucket_sort(nB, bk, S.bSorted, S.buckets, p.nBuckets, sh.scratch);
      const bool tieB = block_has_ties(nB, S.bCot, S.bSorted, &sh.tieTmp);
      if (tieB && tid == 0) sh.tie = 1;
      if (tieB && p.exactTies && nB > 16) {
        // scratch in the arena (tCot / tSeq / tSorted are dead, the pools not yet alive)
        unsigned char* a = S.arena;
        TieItem* W = reinterpret_cast<TieItem*>(a);
        uint16_t* seqSorted = reinterpret_cast<uint16_t*>(a + align16(8ull * p.capB));
        uint16_t* grpOf = reinterpret_cast<uint16_t*>(a + align16(8ull * p.capB) + align16(2ull * p.capB));
        block_fix_ties(nB, S.bCot, S.bSeq, sh.winBp[nBot], S.bSorted, W, seqSorted, grpOf, S.buckets, p.nBuckets,
                       sh.scratch);
      }
    }
    __syncthreads();  // the arena now belongs to the candidate pools

    // ---- phase 3: triplets + filter, one thread per bottom ---------------
    const MiddleSp mid = sh.mid;
    unsigned long long myTests = 0, myCands = 0;
    for (uint32_t base = 0; base < nB; base += blockDim.x) {
      const uint32_t j = base + tid;
      const bool valid = j < nB;
      BottomCtx bc;
      uint32_t H = 0, brk = 0;
      if (valid) {
        const uint32_t idx = S.bSorted[j];
        const uint32_t pos = seq_to_pos(S.bSeq[idx], sh.winBp, sh.winBs, nBot);
        const float2 zr = ldg2(p.pZR + pos), xy = ldg2(p.pXY + pos), var = ldg2(p.pVar + pos);
        float dR, dZ;
        DoubletRec rec;
        doublet_zr_cuts<true>(cfg, mid, zr.x, zr.y, dR, dZ);
        doublet_finish<true>(cfg, mid, dR, dZ, xy.x, xy.y, zr.y, var.x, var.y, p.zWinLo, p.zWinHi, p.nZWin, rec);
        bc.cotThetaB = rec.cotTheta; bc.erB = rec.er; bc.iDeltaRB = rec.iDeltaR; bc.Ub = rec.u; bc.Vb = rec.v;
        bottom_ctx(cfg, bc);
        // |P_j|: tops with cotT <= cotB
        uint32_t lo = 0, hi = nT;
        while (lo < hi) {
          const uint32_t md = (lo + hi) >> 1;
          if (bc.cotThetaB < S.sCot[md]) hi = md; else lo = md + 1;
        }
        float cu, im;
        for (int t = (int)lo - 1; t >= 0; --t) {  // last failing top of the prefix -> H_j
          const int cls = eval_pair(cfg, mid.r, mid.varZ, mid.varR, bc, S.sCot[t], S.sEr[t], S.sIDR[t], S.sU[t], S.sV[t], cu, im);
          if (cls == kPairFailA) { H = (uint32_t)t + 1; break; }
          if (cls == kPairFailB) { H = (uint32_t)t; break; }
        }
        uint32_t k = lo;
        for (; k < nT; ++k) {  // first failing top beyond the prefix -> brk_j
          const int cls = eval_pair(cfg, mid.r, mid.varZ, mid.varR, bc, S.sCot[k], S.sEr[k], S.sIDR[k], S.sU[k], S.sV[k], cu, im);
          if (cls == kPairFailA || cls == kPairFailB) break;
        }
        brk = k;
      }
      // window start = running max of H over the preceding bottoms
      uint32_t blockMax;
      const uint32_t exclMax = block_scan_exclusive(H, sh.scratch, blockMax, OpMax());
      const uint32_t carry = sh.carry;
      const uint32_t start = exclMax > carry ? exclMax : carry;
      __syncthreads();
      if (tid == 0) {
        sh.carry = blockMax > carry ? blockMax : carry;
        sh.poolCount = 0;
      }
      __syncthreads();

      // emission into the linked pool
      uint32_t nMine = 0, head = 0xFFFFu, tail = 0xFFFFu;
      if (valid) {
        for (uint32_t t = start; t < brk; ++t) {
          float cu, im;
          ++myTests;
          const int cls = eval_pair(cfg, mid.r, mid.varZ, mid.varR, bc, S.sCot[t], S.sEr[t], S.sIDR[t], S.sU[t], S.sV[t], cu, im);
          if (cls != kPairEmit) continue;
          const uint32_t slot = atomicAdd(&sh.poolCount, 1u);
          ++nMine;
          if (slot < p.capPool) {
            const uint32_t tp = S.sPos[t];
            const float2 tzr = ldg2(p.pZR + tp);
            float topR = tzr.y;
            if (cfg.useDeltaRinsteadOfTopRadius) {
              const float dr = fsub(tzr.y, mid.r), dz = fsub(tzr.x, mid.z);
              topR = fsqrt(fadd(fmul(dr, dr), fmul(dz, dz)));
            }
            Cand c;
            c.curv = cu; c.impactOrWeight = im; c.topR = topR; c.tOwner = t | (tid << 16);
            S.pool[slot] = c;
            S.poolNext[slot] = 0xFFFFu;
            if (tail != 0xFFFFu) S.poolNext[tail] = (uint16_t)slot; else head = slot;
            tail = slot;
          }
        }
      }
      __syncthreads();
      const uint32_t poolCount = sh.poolCount;
      if (poolCount > p.capPool) {
        if (tid == 0) {
          sh.bad = 1;
          if (p.overflowList != nullptr) {
            p.overflowList[atomicAdd(p.overflowCount, 1u)] = w;
            sh.cnt[kCntMiddles] -= 1;
          } else {
            atomicOr(p.status, kStatusOverflowPool);
          }
        }
        __syncthreads();
        break;
      }
      uint32_t totalCands;
      const uint32_t seg = block_scan_exclusive(nMine, sh.scratch, totalCands, OpSum());
      if (poolCount == 0) continue;  // uniform
      myCands += nMine;
      if (nMine > 0) {
        uint32_t cur = head;
        for (uint32_t i = 0; i < nMine; ++i) {
          S.pool2[seg + i] = S.pool[cur];
          cur = S.poolNext[cur];
        }
        // BroadTripletSeedFilter.cpp:143-148: sort by curvature (libstdc++ replay)
        Cand* mine = S.pool2 + seg;
        if (nMine > 1) std_sort(mine, (int)nMine, cand_less);
        for (uint32_t k = 0; k < nMine; ++k) {
          const float wgt = filter_weight(
              cfg, (int)nMine, (int)k, mine[k].impactOrWeight, [&](int i) { return mine[i].curv; },
              [&](int i) { return mine[i].topR; });
          mine[k].impactOrWeight = wgt;
        }
      }
      __syncthreads();
      // bounded heap pushes in the reference's order (bottom-major, curvature
      // order): CandidatesForMiddleSp.cpp:44-75
      if (tid < 32) {
        const int nLow = (int)cfg.maxSeedsPerSpMConf;
        for (uint32_t c0 = 0; c0 < poolCount && nLow > 0; c0 += 32) {
          const uint32_t i = c0 + tid;
          const int hs = sh.heapSize;
          const float hmin = sh.heapMin;
          float wgt = 0.f;
          bool want = false;
          if (i < poolCount) {
            wgt = S.pool2[i].impactOrWeight;
            want = (hs < nLow) || (wgt > hmin);
          }
          uint32_t mask = __ballot_sync(0xffffffffu, want);
          if (tid == 0) {
            while (mask != 0u) {
              const uint32_t q = c0 + (uint32_t)(__ffs(mask) - 1);
              mask &= mask - 1u;
              const Cand c = S.pool2[q];
              const float wq = c.impactOrWeight;
              const uint32_t owner = c.tOwner >> 16, tRank = c.tOwner & 0xFFFFu;
              const uint32_t bIdx = S.bSorted[base + owner];
              if (sh.heapSize < nLow) {
                StoredSeed sd;
                sd.bottomPos = seq_to_pos(S.bSeq[bIdx], sh.winBp, sh.winBs, nBot);
                sd.topPos = S.sPos[tRank];
                sd.weight = wq;
                sd.zOrigin = fsub(mid.z, fmul(mid.r, S.bCot[bIdx]));
                const int slotI = sh.heapSize;
                sh.storage[slotI] = sd;
                sh.heap[slotI].weight = wq;
                sh.heap[slotI].index = (uint32_t)slotI;
                sh.heapSize = slotI + 1;
                std_push_heap(sh.heap, sh.heapSize, heap_comp);
              } else {
                const WeightIndex smallest = sh.heap[0];
                if (wq <= smallest.weight) continue;
                StoredSeed sd;
                sd.bottomPos = seq_to_pos(S.bSeq[bIdx], sh.winBp, sh.winBs, nBot);
                sd.topPos = S.sPos[tRank];
                sd.weight = wq;
                sd.zOrigin = fsub(mid.z, fmul(mid.r, S.bCot[bIdx]));
                sh.storage[smallest.index] = sd;
                std_pop_heap(sh.heap, sh.heapSize, heap_comp);
                sh.heap[sh.heapSize - 1].weight = wq;
                sh.heap[sh.heapSize - 1].index = smallest.index;
                std_push_heap(sh.heap, sh.heapSize, heap_comp);
              }
              sh.heapMin = sh.heap[0].weight;
            }
          }
          __syncwarp();
        }
      }
      __syncthreads();
    }  // rounds over bottoms

    // per-thread counters -> block counters
    {
      unsigned long long t = myTests, c = myCands;
      for (int d = 16; d > 0; d >>= 1) {
        t += __shfl_down_sync(0xffffffffu, t, d);
        c += __shfl_down_sync(0xffffffffu, c, d);
      }
      if ((tid & 31) == 0 && !sh.bad) {
        atomicAdd(&sh.cnt[kCntTripletTests], t);
        atomicAdd(&sh.cnt[kCntCandidates], c);
      }
    }
    __syncthreads();

    // ---- phase 4: per-middle selection (BroadTripletSeedFilter.cpp:324-393)
    if (tid == 0) {
      uint32_t nOut = 0;
      if (!sh.bad) {
        std_sort_heap(sh.heap, sh.heapSize, heap_comp);
        uint32_t maxSeeds = (uint32_t)sh.heapSize;
        if (maxSeeds > cfg.maxSeedsPerSpM) maxSeeds = cfg.maxSeedsPerSpM + 1;
        for (uint32_t i = 0; i < (uint32_t)sh.heapSize && i < maxSeeds; ++i) {
          const StoredSeed sd = sh.storage[sh.heap[i].index];
          const size_t o = (size_t)w * p.seedsPerMiddle + i;
          p.slotB[o] = sd.bottomPos;
          p.slotM[o] = m;
          p.slotT[o] = sd.topPos;
          p.slotQ[o] = sd.weight;
          p.slotZ[o] = sd.zOrigin;
          ++nOut;
        }
      }
      p.slotCount[w] = nOut;
      if (!sh.bad) {
        sh.cnt[kCntBottomDoublets] += nB;
        sh.cnt[kCntTopDoublets] += nT;
        sh.cnt[kCntSeeds] += nOut;
        sh.cnt[kCntTieMiddles] += sh.tie;
      }
    }
  }
  __syncthreads();
  if (tid < (uint32_t)kCntSlots && tid != (uint32_t)kCntInGrid && sh.cnt[tid] != 0ull) {
    atomicAdd(p.counters + tid, sh.cnt[tid]);
  }
}

// ---------------------------------------------------------------------------
// Seed compaction (ordered): tiled exclusive scan of the per-middle counts
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tile_sums(const __grid_constant__ CompactParams p) {
  __shared__ uint32_t scratch[34];
  const uint32_t nWork = *p.nWorkPtr;
  const uint32_t tile = blockIdx.x;
  uint32_t s = 0;
  for (uint32_t i = threadIdx.x; i < (uint32_t)kTile; i += blockDim.x) {
    const uint32_t w = tile * kTile + i;
    if (w < nWork) s += p.slotCount[w];
  }
  uint32_t total;
  block_scan_exclusive(s, scratch, total, OpSum());
  if (threadIdx.x == 0) p.tileSums[tile] = total;
}

__global__ void __launch_bounds__(256) k_compact_seeds(const __grid_constant__ CompactParams p) {
  __shared__ uint32_t scratch[34];
  __shared__ uint32_t carry;
  const uint32_t nWork = *p.nWorkPtr;
  const uint32_t tile = blockIdx.x;
  if (threadIdx.x == 0) carry = p.tilePrefix[tile];
  __syncthreads();
  for (uint32_t base = 0; base < (uint32_t)kTile; base += blockDim.x) {
    const uint32_t w = tile * kTile + base + threadIdx.x;
    const uint32_t n = w < nWork ? p.slotCount[w] : 0u;
    uint32_t total;
    const uint32_t excl = block_scan_exclusive(n, scratch, total, OpSum());
    const uint32_t c = carry;
    if (w < nWork) {
      const uint32_t o = c + excl;
      p.seedStart[w] = o;
      for (uint32_t i = 0; i < n; ++i) {
        if ((unsigned long long)(o + i) < p.outCapacity) {
          const size_t s = (size_t)w * p.seedsPerMiddle + i;
          p.outB[o + i] = p.pIdx[p.slotB[s]];
          p.outM[o + i] = p.pIdx[p.slotM[s]];
          p.outT[o + i] = p.pIdx[p.slotT[s]];
          p.outQ[o + i] = p.slotQ[s];
          p.outZ[o + i] = p.slotZ[s];
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (tile * (uint32_t)kTile < nWork && (tile + 1) * (uint32_t)kTile >= nWork && threadIdx.x == 0) {
    p.seedStart[nWork] = carry;
  }
}

__global__ void k_event_offsets(const __grid_constant__ CompactParams p) {
  const uint32_t nWork = *p.nWorkPtr;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e <= p.nEvents; e += gridDim.x * blockDim.x) {
    const uint32_t w = e < p.nEvents ? p.workStart[(size_t)e * p.nNav] : nWork;
    // seedStart[nWork] is only written when nWork > 0
    p.seedOffsets[e] = nWork == 0 ? 0ull : (unsigned long long)p.seedStart[w];
  }
}

// phi = atan2f(y, x) replay, for validation against the host libm
__global__ void k_atan2f(const float* __restrict__ y, const float* __restrict__ x, float* __restrict__ out, unsigned long long n) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    out[i] = glibc_atan2f(y[i], x[i]);
  }
}

}  // namespace b200seed
