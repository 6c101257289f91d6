// seeding_kernels.cuh -- CUDA kernels of the B200 seeding engine (sm_100a).
//
// Stage map (reference lines they replace; paths relative to the ACTS tree):
//   k_bin_count / k_scan / k_scatter / k_sort_bins
//        grid fill, per-bin r sort, packed SoA copy
//        GridTripletSeedingAlgorithm.cpp:208-253, SpacePointGridBase.hpp:67-87
//   k_middle_ranges / k_fill_work
//        BinnedGroup iteration + middle r range, TripletSeeder.cpp:183-195,
//        GridTripletSeedingAlgorithm.cpp:346-371,404-421
//   k_doublets<count> / k_cap_* / k_plan_chunks / k_doublets<fill>
//        doublet search, two passes (count, then fill into an HBM arena)
//        TripletSeeder.cpp:52-82,138-181, DoubletSeedFinder.cpp:41-273
//   k_seed_middles
//        cotTheta sort, triplets, filter, per-middle selection
//        TripletSeeder.cpp:21-42,89-107, DoubletSeedFinder.hpp:94-104,
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

namespace B200SEED_NS {

#ifndef B200SEED_SPLIT_WALKERS
#define B200SEED_SPLIT_WALKERS 0  // scans: 1 = separate backward / forward walkers per bottom, 0 = one merged walk
#endif
#ifndef B200SEED_PREFETCH_SLOT
#define B200SEED_PREFETCH_SLOT 1  // L2 prefetch of the middle's arena slot at the top of the middle
#endif
#ifndef B200SEED_CLASSIFY
#define B200SEED_CLASSIFY classify_pair_flat  // or classify_pair (early exits)
#endif
#ifndef B200SEED_REFILL
#define B200SEED_REFILL 12  // idle lanes of a warp that trigger a refill of the scan walkers
#endif
#ifndef B200SEED_FLAT_WALK
#define B200SEED_FLAT_WALK 1  // scans: straight-line walker bookkeeping (selects) instead of the backward / forward branches
#endif
#ifndef B200SEED_MERGE_3B3C
#define B200SEED_MERGE_3B3C 1  // window starts and the gap list in one pass (two block barriers and one sweep less)
#endif
#ifndef B200SEED_OPAQUE_EMIT
#define B200SEED_OPAQUE_EMIT 1  // candidate emission: one shared-memory atomic per lane (no vote / leader / shuffle aggregation)
#endif
#ifndef B200SEED_WORK_CHUNK
#define B200SEED_WORK_CHUNK 8  // consecutive work items a block takes at a time (1 = none)
#endif
constexpr uint32_t kInvalidBin = 0xFFFFFFFFu;
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
  kStatusBinTooLarge = 4,
  kStatusOverflowRecords = 8,
  kStatusArenaTooSmall = 16
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
  unsigned long long* sortScratch;  // [4 * nTotal] for bins larger than smem
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
  // phi-sector split of a single event over several GPUs: only middle bins
  // whose phi bin lies in [phiFirst, phiFirst + phiCount) are seeded
  uint32_t phiFirst, phiCount;
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
  const uint32_t* eventFirstItem;  // NULL: event e starts at workStart[e * nNav]; else at workStart[eventFirstItem[e]]
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

// Exclusive scan (sum) of an array of n words in place; returns the total
// (to every thread).  Every thread scans kItems consecutive words serially, one
// block scan combines the per-thread sums: a single pass for n <= 8 * blockDim.x.
__device__ __forceinline__ uint32_t block_scan_array(uint32_t* a, uint32_t n, uint32_t* scratch) {
  constexpr uint32_t kItems = 8;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n; base += blockDim.x * kItems) {
    const uint32_t i0 = base + threadIdx.x * kItems;
    uint32_t v[kItems];
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < kItems; ++k) {
      v[k] = (i0 + k) < n ? a[i0 + k] : 0u;
      sum += v[k];
    }
    uint32_t total;
    uint32_t run = block_scan_exclusive(sum, scratch, total, OpSum()) + carry;
#pragma unroll
    for (uint32_t k = 0; k < kItems; ++k) {
      if ((i0 + k) < n) a[i0 + k] = run;
      run += v[k];
    }
    carry += total;
  }
  __syncthreads();
  return carry;
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

// Item of the tie replays: sort key + payload (top bit of val may carry a flag).
struct TieItem {
  float key;
  uint32_t val;
};
__device__ __forceinline__ bool tie_less(const TieItem& a, const TieItem& b) { return a.key < b.key; }

// Pruned replay of libstdc++'s introsort (see std_sort_replay_ties in
// seed_math.h) executed by ONE WARP: all 32 lanes call it with the same
// arguments on an array in shared or global memory.  The Hoare partition of
// libstdc++ pairs the k-th misplaced element from the left with the k-th
// misplaced element from the right; here 32 + 32 elements are classified per
// step with ballots and swapped by rank, the last < 64 elements of a partition
// run through the literal sequential loop (tests/model: model_check_warp_replay
// checks this formulation against std::sort).
template <typename Flagged>
__device__ __forceinline__ void warp_sort_replay_ties(TieItem* a, int n, Flagged flagged) {
  if (n <= 16) return;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t ltMask = (1u << lane) - 1u;
  int lg = 0;
  for (unsigned v = (unsigned)n; v > 1; v >>= 1) ++lg;
  int sf[64], sl[64], sd[64];
  int sp = 0;
  sf[0] = 0; sl[0] = n; sd[0] = 2 * lg; sp = 1;
  while (sp > 0) {
    --sp;
    int first = sf[sp], last = sl[sp], depth = sd[sp];
    while (last - first > 16) {
      int c = 0;
      for (int base = first; base < last && c < 2; base += 32) {
        const int i = base + (int)lane;
        const bool f = i < last && flagged(a[i]);
        c += __popc(__ballot_sync(0xffffffffu, f));
      }
      if (c < 2) break;  // nothing left to decide in this range
      if (depth == 0) {
        if (lane == 0) {
          std_make_heap(a + first, last - first, tie_less);
          std_sort_heap(a + first, last - first, tie_less);
        }
        __syncwarp();
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      {
        const int ia = first + 1, ib = mid, ic = last - 1;
        const TieItem va = a[ia], vb = a[ib], vc = a[ic];
        int pick;
        if (tie_less(va, vb)) {
          if (tie_less(vb, vc)) pick = ib;
          else if (tie_less(va, vc)) pick = ic;
          else pick = ia;
        } else if (tie_less(va, vc)) pick = ia;
        else if (tie_less(vb, vc)) pick = ic;
        else pick = ib;
        __syncwarp();
        if (lane == 0) {
          const TieItem t = a[first]; a[first] = a[pick]; a[pick] = t;
        }
        __syncwarp();
      }
      const TieItem pivot = a[first];
      int lo = first + 1, hi = last;
      while (hi - lo >= 64) {
        const TieItem el = a[lo + (int)lane], er = a[hi - 1 - (int)lane];
        const bool ml = !tie_less(el, pivot), mr = !tie_less(pivot, er);
        const uint32_t mL = __ballot_sync(0xffffffffu, ml), mR = __ballot_sync(0xffffffffu, mr);
        const int nL = __popc(mL), nR = __popc(mR), s = nL < nR ? nL : nR;
        const int rkL = __popc(mL & ltMask), rkR = __popc(mR & ltMask);
        const bool swapL = ml && rkL < s, swapR = mr && rkR < s;
        const int srcR = swapL ? (int)__fns(mR, 0, rkL + 1) : (int)lane;
        const int srcL = swapR ? (int)__fns(mL, 0, rkR + 1) : (int)lane;
        TieItem fromRight, fromLeft;
        fromRight.key = __shfl_sync(0xffffffffu, er.key, srcR);
        fromRight.val = __shfl_sync(0xffffffffu, er.val, srcR);
        fromLeft.key = __shfl_sync(0xffffffffu, el.key, srcL);
        fromLeft.val = __shfl_sync(0xffffffffu, el.val, srcL);
        if (swapL) a[lo + (int)lane] = fromRight;
        if (swapR) a[hi - 1 - (int)lane] = fromLeft;
        const int newLo = nL == s ? lo + 32 : lo + (int)__fns(mL, 0, s + 1);
        const int newHi = nR == s ? hi - 32 : hi - (int)__fns(mR, 0, s + 1);
        lo = newLo;
        hi = newHi;
        __syncwarp();
      }
      int cut = 0;
      if (lane == 0) {  // libstdc++ __unguarded_partition resumed from the same (lo, hi) state
        while (true) {
          while (tie_less(a[lo], pivot)) ++lo;
          --hi;
          while (tie_less(pivot, a[hi])) --hi;
          if (!(lo < hi)) break;
          const TieItem t = a[lo]; a[lo] = a[hi]; a[hi] = t;
          ++lo;
        }
        cut = lo;
      }
      cut = __shfl_sync(0xffffffffu, cut, 0);
      __syncwarp();
      sf[sp] = cut; sl[sp] = last; sd[sp] = depth; ++sp;
      last = cut;
    }
  }
  __syncwarp();
}

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
#ifdef B200SEED_RELAXED
      const float phi = p.phi != nullptr ? __ldg(p.phi + i) : atan2f(y, x);  // CUDA libm, not glibc's rounding
#else
      const float phi = p.phi != nullptr ? __ldg(p.phi + i) : glibc_atan2f(y, x);
#endif
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

// Bucket sort of the n entries of a bin (shared or global memory): `put(i, q)` stores entry i at position q,
// `bucketOf(i)` is monotone non-decreasing in the sort key and < nBk, `sortBucket(s, e)` orders positions [s, e)
// (entries of one bucket; one thread).  Replaces a 66-stage bitonic sort of 2048 padded keys by two atomic
// passes, one scan and an insertion sort inside the buckets.
constexpr uint32_t kSortBuckets = 2048;
template <typename BucketOf, typename Put, typename SortBucket>
__device__ __forceinline__ void block_bucket_sort(uint32_t n, uint32_t* buckets, uint32_t* scratch, BucketOf bucketOf, Put put,
                                                  SortBucket sortBucket) {
  for (uint32_t i = threadIdx.x; i <= kSortBuckets; i += blockDim.x) buckets[i] = 0;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(buckets + bucketOf(i), 1u);
  __syncthreads();
  block_scan_array(buckets, kSortBuckets, scratch);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) put(i, atomicAdd(buckets + bucketOf(i), 1u));
  __syncthreads();  // buckets[b] is now the END of bucket b
  for (uint32_t b = threadIdx.x; b < kSortBuckets; b += blockDim.x) {
    const uint32_t s0 = b == 0 ? 0u : buckets[b - 1], e0 = buckets[b];
    if (e0 - s0 >= 2u) sortBucket(s0, e0);
  }
  __syncthreads();
}

__device__ __forceinline__ bool bin_flagged(const TieItem& a) { return (a.val >> 31) != 0u; }

// One block per (event, bin): order the bin by r and write the packed copy.
// The reference sorts the bin's indices (inserted in ascending original index,
// .cpp:211-221) with the unstable std::ranges::sort (.cpp:223-228).  Here:
//   1. bitonic sort by (r, original index)              -> canonical order
//   2. exactTies only, bins with equal radii only: the elements are put in
//      insertion order (second bitonic sort, by original index) and one warp
//      replays libstdc++'s introsort on them (pruned to the ranges that still
//      hold tied elements); tied elements take the slots of their group in the
//      order the replay leaves them in
//   3. gather the six columns into the packed copy
// `keys` and `work` hold `padded` (power of two >= n) 8-byte entries each, in
// shared memory when they fit, else in the global scratch.
__global__ void __launch_bounds__(kSortThreads) k_sort_bins(const __grid_constant__ GridParams p) {
  extern __shared__ unsigned long long smemKeys[];
  __shared__ uint32_t sFlag, sMin, sMax;
  __shared__ uint32_t sScratch[34];
  const uint32_t gb = blockIdx.x;
  const uint32_t b0 = p.binStart[gb], b1 = p.binStart[gb + 1];
  const uint32_t n = b1 - b0;
  if (n == 0) return;
  const uint32_t e = gb / p.nBins;
  const uint32_t evBase = __ldg(p.spOffsets + e);
  const uint32_t nEvent = __ldg(p.spOffsets + e + 1) - evBase;
  // shared memory: keys[cap], work[cap], buckets[kSortBuckets + 1]; a bin beyond the capacity uses its 32 n bytes of
  // the global scratch the same way
  const bool inSmem = n <= p.sortSmemCap;
  unsigned long long* keys = inSmem ? smemKeys : p.sortScratch + 4ull * b0;
  unsigned long long* work = inSmem ? smemKeys + p.sortSmemCap : p.sortScratch + 4ull * b0 + n;
  uint32_t* buckets = inSmem ? reinterpret_cast<uint32_t*>(smemKeys + 2ull * p.sortSmemCap)
                             : reinterpret_cast<uint32_t*>(p.sortScratch + 4ull * b0 + 2ull * n);
  if (threadIdx.x == 0) { sFlag = 0; sMin = 0xFFFFFFFFu; sMax = 0u; }
  __syncthreads();
  {
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t idx = p.tmpIdx[b0 + i];
      const float r = __ldg(p.r + evBase + idx);
      const uint32_t rb = (r == 0.0f) ? 0u : __float_as_uint(r);  // r >= 0 inside the grid: the bit pattern orders like the value
      work[i] = ((unsigned long long)rb << 32) | idx;
      mn = mn < rb ? mn : rb;
      mx = mx > rb ? mx : rb;
    }
    for (int d = 16; d > 0; d >>= 1) {
      const uint32_t om = __shfl_xor_sync(0xffffffffu, mn, d), ox = __shfl_xor_sync(0xffffffffu, mx, d);
      mn = mn < om ? mn : om;
      mx = mx > ox ? mx : ox;
    }
    if ((threadIdx.x & 31u) == 0u) { atomicMin(&sMin, mn); atomicMax(&sMax, mx); }
  }
  __syncthreads();
  // 1. canonical order (r, original index): work -> keys
  {
    const float rMin = __uint_as_float(sMin), rMax = __uint_as_float(sMax);
    const float scale = rMax > rMin ? (float)kSortBuckets / (rMax - rMin) : 0.0f;
    block_bucket_sort(
        n, buckets, sScratch,
        [&](uint32_t i) {
          float t = (__uint_as_float((uint32_t)(work[i] >> 32)) - rMin) * scale;  // monotone in r
          if (!(t > 0.0f)) t = 0.0f;
          const uint32_t b = (uint32_t)t;
          return b < kSortBuckets ? b : kSortBuckets - 1u;
        },
        [&](uint32_t i, uint32_t q) { keys[q] = work[i]; },
        [&](uint32_t s0, uint32_t e0) {
          unsigned long long* a = keys + s0;
          const uint32_t m = e0 - s0;
          if (m <= 48u) {
            for (uint32_t i = 1; i < m; ++i) {
              const unsigned long long v = a[i];
              uint32_t j = i;
              while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; --j; }
              a[j] = v;
            }
          } else {  // many equal or nearly equal radii (quantised coordinates): heapsort, O(m log m) for any input
            auto sift = [&](uint32_t root, uint32_t end) {
              const unsigned long long v = a[root];
              for (;;) {
                uint32_t c = 2u * root + 1u;
                if (c >= end) break;
                if (c + 1u < end && a[c + 1u] > a[c]) ++c;
                if (!(a[c] > v)) break;
                a[root] = a[c];
                root = c;
              }
              a[root] = v;
            };
            for (uint32_t i = m / 2u; i-- > 0u;) sift(i, m);
            for (uint32_t end = m - 1u; end > 0u; --end) {
              const unsigned long long t = a[0]; a[0] = a[end]; a[end] = t;
              sift(0u, end);
            }
          }
        });
  }
  if (p.exactTies && n > 16) {
    bool t = false;
    for (uint32_t i = threadIdx.x + 1; i < n; i += blockDim.x) t |= (keys[i] >> 32) == (keys[i - 1] >> 32);
    if (t) sFlag = 1;
    __syncthreads();
    if (sFlag != 0) {  // block-uniform
      // 2. insertion order (ascending original index): the replay items {r, canonical rank | tied flag} are scattered
      // straight into that order (original indices are spread evenly over [0, nEvent))
      TieItem* W = reinterpret_cast<TieItem*>(work);
      auto idxOfItem = [&](const TieItem& it) { return (uint32_t)(keys[it.val & 0x7fffffffu] & 0xffffffffull); };
      block_bucket_sort(
          n, buckets, sScratch,
          [&](uint32_t i) {
            const uint32_t b = (uint32_t)(((keys[i] & 0xffffffffull) * kSortBuckets) / nEvent);
            return b < kSortBuckets ? b : kSortBuckets - 1u;
          },
          [&](uint32_t i, uint32_t q) {
            const uint32_t rb = (uint32_t)(keys[i] >> 32);
            const bool tied = (i > 0 && (uint32_t)(keys[i - 1] >> 32) == rb) || (i + 1 < n && (uint32_t)(keys[i + 1] >> 32) == rb);
            TieItem it;
            it.key = __uint_as_float(rb);
            it.val = i | (tied ? 0x80000000u : 0u);
            W[q] = it;
          },
          [&](uint32_t s0, uint32_t e0) {
            for (uint32_t i = s0 + 1; i < e0; ++i) {
              const TieItem v = W[i];
              const uint32_t vi = idxOfItem(v);
              uint32_t j = i;
              while (j > s0 && idxOfItem(W[j - 1]) > vi) { W[j] = W[j - 1]; --j; }
              W[j] = v;
            }
          });
      __syncthreads();
      if (threadIdx.x < 32) warp_sort_replay_ties(W, (int)n, bin_flagged);
      __syncthreads();
      // Tied element k (in replay order) of a group takes the group's k-th canonical slot.  One warp walks W in
      // order: the group of an element is the first canonical rank with its radius (lower bound over the sorted
      // keys), its place inside the group comes from a ballot scan + a running member count per group (kept in
      // binOf[b0 + g], free since k_scatter) -- linear in n however many elements tie.  The reassignments are
      // staged in tmpIdx (free from here on), then patched into the keys.
      uint32_t* grpCount = p.binOf + b0;
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) grpCount[i] = 0;
      __syncthreads();
      if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        for (uint32_t base = 0; base < n; base += 32) {
          const uint32_t i = base + lane;
          TieItem it;
          it.key = 0.f;
          it.val = 0u;
          if (i < n) it = W[i];
          const bool flagged = i < n && bin_flagged(it);
          const uint32_t rank = it.val & 0x7fffffffu;
          const uint32_t rb = __float_as_uint(it.key);
          uint32_t g = 0;
          if (flagged) {
            uint32_t lo = 0, hi = rank;
            while (lo < hi) {
              const uint32_t mid = (lo + hi) >> 1;
              if ((uint32_t)(keys[mid] >> 32) < rb) lo = mid + 1; else hi = mid;
            }
            g = lo;
          }
          const uint32_t same = __match_any_sync(0xffffffffu, flagged ? g : 0x80000000u + lane);  // unflagged lanes match themselves only
          const uint32_t before = (uint32_t)__popc(same & ((1u << lane) - 1u));
          uint32_t carried = 0;
          if (flagged && before == 0u) {  // first member of its group in this step: one lane per group
            carried = grpCount[g];
            grpCount[g] = carried + (uint32_t)__popc(same);
          }
          carried = __shfl_sync(0xffffffffu, carried, __ffs(same) - 1);
          if (flagged) p.tmpIdx[b0 + g + carried + before] = (uint32_t)(keys[rank] & 0xffffffffull);
          __syncwarp();
        }
      }
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t rb = (uint32_t)(keys[i] >> 32);
        const bool tied = (i > 0 && (uint32_t)(keys[i - 1] >> 32) == rb) || (i + 1 < n && (uint32_t)(keys[i + 1] >> 32) == rb);
        if (tied) {
          const uint32_t idx = p.tmpIdx[b0 + i];
          // all members of a group share rb, so only the index half changes
          work[i] = ((unsigned long long)rb << 32) | idx;
        } else {
          work[i] = keys[i];
        }
      }
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) keys[i] = work[i];
      __syncthreads();
    }
  }
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t idx = (uint32_t)(keys[i] & 0xffffffffull);
    const uint32_t src = evBase + idx;
    p.pIdx[b0 + i] = idx;
    p.pXY[b0 + i] = make_float2(__ldg(p.x + src), __ldg(p.y + src));
    p.pZR[b0 + i] = make_float2(__ldg(p.z + src), __ldg(p.r + src));
    p.pVar[b0 + i] = make_float2(__ldg(p.varZ + src), __ldg(p.varR + src));
  }
}

// Strip triplet path: the derived calibration details of every space point of the packed copy
// (deriveOuterStripSpacePointCalibrationDetails, StripSpacePointCalibrationImpl.hpp:22-42; the reference derives them
// per (bottom, top) pair, TripletSeedFinder.cpp:205-213,282-286 -- same inputs, same operations, same bits).
// raw: 12 floats per ORIGINAL space point of the (single) event.
__global__ void __launch_bounds__(256) k_gather_strips(const uint32_t* __restrict__ pIdx, const uint32_t* __restrict__ nPackedPtr,
                                                       const float* __restrict__ raw, StripDerived* __restrict__ out) {
  const uint32_t nPacked = *nPackedPtr;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < nPacked; pos += stride) {
    float d[12];
    const float4* src = reinterpret_cast<const float4*>(raw + 12ull * pIdx[pos]);  // 48 bytes per point: 16-byte aligned
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float4 v = __ldg(src + q);
      d[4 * q] = v.x; d[4 * q + 1] = v.y; d[4 * q + 2] = v.z; d[4 * q + 3] = v.w;
    }
    StripDerived o;
    strip_derive(d, o);
    const uint4* w = reinterpret_cast<const uint4*>(&o);
    uint4* dst = reinterpret_cast<uint4*>(out + pos);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = w[q];
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
    // local phi bin (1-based) of this middle bin: global = (phi * (nZ + 2) + z) * (nR + 2) + r
    const uint32_t phiBin = bin / (uint32_t)((p.cfg.nZ + 2) * (p.cfg.nR + 2));
    const bool inSector = phiBin >= p.phiFirst && phiBin - p.phiFirst < p.phiCount;
    if (b0 != b1 && inSector) {
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

// First index in [lo, hi) where the monotone predicate holds (hi if none),
// searched by a whole warp: 32 probes per step instead of one.
template <typename Pred>
__device__ __forceinline__ uint32_t warp_first_true(uint32_t lo, uint32_t hi, Pred pred) {
  const uint32_t lane = threadIdx.x & 31;
  while (hi - lo > 32) {
    const uint32_t n = hi - lo;
    const uint32_t step = (n + 31) / 32;
    uint32_t q = lo + (lane + 1) * step - 1;
    if (q > hi - 1) q = hi - 1;
    const uint32_t mask = __ballot_sync(0xffffffffu, pred(q));
    if (mask == 0u) return hi;  // the last probe is hi - 1
    const int f = __ffs(mask) - 1;
    const uint32_t qf = __shfl_sync(0xffffffffu, q, f);
    const uint32_t qp = __shfl_sync(0xffffffffu, q, f > 0 ? f - 1 : 0);
    lo = f > 0 ? qp + 1 : lo;
    hi = qf;  // pred(qf) holds: the answer is qf unless an earlier index in [lo, qf) holds
    if (hi == lo) return qf;
    // search [lo, qf); if nothing is found there the answer is qf
    uint32_t inner = hi;
    {
      uint32_t l2 = lo, h2 = hi;
      while (h2 - l2 > 32) {
        const uint32_t n2 = h2 - l2;
        const uint32_t st2 = (n2 + 31) / 32;
        uint32_t q2 = l2 + (lane + 1) * st2 - 1;
        if (q2 > h2 - 1) q2 = h2 - 1;
        const uint32_t m2 = __ballot_sync(0xffffffffu, pred(q2));
        if (m2 == 0u) { l2 = h2; break; }
        const int f2 = __ffs(m2) - 1;
        const uint32_t qf2 = __shfl_sync(0xffffffffu, q2, f2);
        const uint32_t qp2 = __shfl_sync(0xffffffffu, q2, f2 > 0 ? f2 - 1 : 0);
        l2 = f2 > 0 ? qp2 + 1 : l2;
        h2 = qf2;
        inner = qf2;
      }
      if (l2 < h2) {
        const uint32_t i = l2 + lane;
        const uint32_t m3 = __ballot_sync(0xffffffffu, i < h2 && pred(i));
        if (m3 != 0u) inner = l2 + (uint32_t)(__ffs(m3) - 1);
      }
    }
    return inner;
  }
  const uint32_t i = lo + lane;
  const uint32_t mask = __ballot_sync(0xffffffffu, i < hi && pred(i));
  return mask != 0u ? lo + (uint32_t)(__ffs(mask) - 1) : hi;
}

// float <-> int image that preserves the order (for shared-memory atomicMin/Max)
__device__ __forceinline__ int float_to_ordered(float f) {
  const int k = __float_as_int(f);
  return k >= 0 ? k : k ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// first_true with a hint: the answer is expected within 32 elements of `lo` (windows of consecutive middles of a bin)
template <typename Pred>
__device__ __forceinline__ uint32_t warp_first_true_near(uint32_t lo, uint32_t hi, Pred pred) {
  if (lo >= hi) return hi;
  const uint32_t q = lo + (threadIdx.x & 31);
  const uint32_t mask = __ballot_sync(0xffffffffu, q < hi && pred(q));
  if (mask != 0u) return lo + (uint32_t)(__ffs(mask) - 1);
  if (hi - lo <= 32u) return hi;
  return warp_first_true(lo + 32u, hi, pred);
}


// ---------------------------------------------------------------------------
// Doublet stage (DoubletSeedFinder.cpp:41-273, TripletSeeder.cpp:52-82,138-181):
// two passes, count then fill, so that the output is allocation-free.
//   k_doublets<false>  one warp per middle: r windows of every neighbour bin, the
//                      (z, r) cuts of every candidate -> capT / capB = survivors
//                      per side (= the doublet counts unless a cut of the second
//                      half -- interaction-point cut, experiment cuts -- rejects
//                      more: an upper bound that sizes the middle's arena slot)
//   k_cap_*            tiled exclusive scan of the slot sizes (64-bit prefix)
//   k_plan_chunks      chunks of consecutive work items that fit the arena
//   k_doublets<true>   the same sweep again; survivors are queued per warp so that
//                      the coordinate transform runs on dense warps; every doublet
//                      is written as one 32-byte record {sp, cotTheta, iDeltaR, er,
//                      u, v, x', y'} (the reference's DoubletsForMiddleSp columns,
//                      DoubletSeedFinder.hpp:26-262) + its cotTheta key, in the
//                      reference's emission order (neighbour bins in order, ascending
//                      position inside a bin); the middle gets a header and is
//                      appended to the work list of the shared-memory class its
//                      exact list sizes fit (k_seed_middles)
// ---------------------------------------------------------------------------
struct __align__(16) DoubletRecord {
  uint32_t pos;  // packed position of the other space point
  float cotTheta, iDeltaR, er, u, v, xNew, yNew;
};
static_assert(sizeof(DoubletRecord) == 32, "one 32-byte sector per doublet");

struct __align__(16) MiddleHeader {
  uint32_t nB, nT;     // doublets of the two sides (0 / 0: the middle does not reach the triplet stage)
  uint32_t capB;       // slot layout: bottoms at [offset, offset + nB), tops at [offset + capB, offset + capB + nT)
  uint32_t offset;     // first record of the middle's slot, relative to the chunk's arena
  int cotMinB, cotMaxB, cotMinT, cotMaxT;  // ordered-int images of the cotTheta ranges (bucket sort scaling)
};
static_assert(sizeof(MiddleHeader) == 32, "header is loaded as two 16-byte words");

// Shared-memory classes of k_seed_middles: a middle goes to the smallest class its lists fit.
constexpr int kNumSeedClasses = 6;  // 5 shared-memory classes (6, 4, 3, 2, 1 resident blocks per SM) + the spill class (lists in global memory)
constexpr int kChunkCounterWords = 32;  // per arena chunk: [0] fill ticket, [1 + k] ticket of class k, [16 + k] list length of class k
constexpr int kSpillClass = kNumSeedClasses - 1;
constexpr uint32_t kMaxListLength = 65534;  // 16-bit ranks
constexpr uint32_t kMaxChunks = 4096;

// Byte offsets of the per-middle arrays of k_seed_middles inside the block's dynamic shared memory (or, for
// the spill class, the block's global scratch).  Everything is sized by the EXACT list lengths of the middle.
struct SeedCarve {
  uint32_t oRankB;    // u16[nB]   sorted rank -> index of the bottom in the arena slot
  uint32_t oTstar;    // u16[nB]   |P_j|, later the last failing top of the prefix
  uint32_t oTops;     // sorted tops: float4[nT] {cotTheta, er, iDeltaR, u}, float[nT] v, u32[nT] pos
  uint32_t oBuckets;  // u32[nBk + 1]
  uint32_t nBk;
  // region a (until the tops are gathered)
  uint32_t oKeyB, oKeyT;  // float[nB], float[nT]
  uint32_t oRankT;        // u16[nT]
  uint32_t oTie, oTieGrp; // TieItem[max(nB, nT)], u16[max(nB, nT)]
  // region b (after the tops are gathered), overlays region a
  uint32_t oHval;     // u16[nB]
  uint32_t oCnt;      // u32[nB + 1]
  uint32_t oPool;     // u32[P] emission records, then Cand[P]
  uint32_t endA;
  uint32_t minBytes;  // with the smallest pool the class assignment guarantees
  uint32_t pad;       // 16 words: stored per middle by the fill pass, loaded as four 16-byte words by the seeding kernel
};
static_assert(sizeof(SeedCarve) == 64, "four 16-byte words");

B2S_HD uint32_t carve_align(uint32_t v) { return (v + 15u) & ~15u; }
B2S_HD uint32_t seed_pool_min(uint32_t nB) { return nB / 2u + 32u; }  // candidates per middle: 0.29 nB on average at <mu>=200 (a middle whose pool overflows moves up one class)
constexpr uint32_t kPoolEntryBytes = 20;  // u32 emission record + 16-byte Cand

B2S_HD SeedCarve seed_carve(uint32_t nB, uint32_t nT) {
  SeedCarve c;
  uint32_t o = 0;
  c.oRankB = o; o = carve_align(o + 2u * nB);
  c.oTstar = o; o = carve_align(o + 2u * nB);
  c.oTops = o; o = carve_align(o + 24u * nT);
  uint32_t nBk = 64;
  while (nBk < 4096u && nBk * 2u < nB + nT) nBk <<= 1;
  c.nBk = nBk;
  c.oBuckets = o; o = carve_align(o + 4u * (nBk + 1u));
  uint32_t a = o;
  c.oKeyB = a; a = carve_align(a + 4u * nB);
  c.oKeyT = a; a = carve_align(a + 4u * nT);
  c.oRankT = a; a = carve_align(a + 2u * nT);
  const uint32_t n = nB > nT ? nB : nT;
  c.oTie = a; a = carve_align(a + 8u * n);
  c.oTieGrp = a; a = carve_align(a + 2u * n);
  c.endA = a;
  uint32_t b = o;
  c.oHval = b; b = carve_align(b + 2u * nB);
  c.oCnt = b; b = carve_align(b + 4u * (nB + 1u));
  c.oPool = b;
  const uint32_t withPool = b + kPoolEntryBytes * seed_pool_min(nB);
  c.minBytes = a > withPool ? a : withPool;
  c.pad = 0;
  return c;
}

struct DoubletParams {
  DeviceConfig cfg;
  const float2 *pXY, *pZR, *pVar;
  const uint32_t* binStart;
  const uint32_t *navBins, *botOffsets, *botBins, *topOffsets, *topBins;
  const uint32_t *workPos, *workEG;
  const uint32_t* nWorkPtr;
  uint32_t nNav, nBins;
  // VertexZCuts: per event disjoint windows in ascending order; event e owns [zWinOffsets[e], zWinOffsets[e + 1])
  // (zWinOffsets == NULL: the nZWin windows apply to every event)
  const float *zWinLo, *zWinHi;
  const uint32_t* zWinOffsets;
  int nZWin;
  uint32_t* workCounter;  // ticket counter of this launch
  uint32_t *capB, *capT;  // [nWork] (z, r) survivors per side
  uint32_t* planWords;    // [0] largest footprint, [1] largest capB, [2] largest capT (atomicMax)
  // fill pass
  uint32_t itemFirst, itemEnd;           // the chunk: work items [itemFirst, itemEnd)
  const unsigned long long* slotPrefix;  // [nWork + 1] exclusive prefix of capB + capT
  DoubletRecord* rec;                    // arena of the chunk
  float* key;
  MiddleHeader* hdr;                     // [nWork]
  SeedCarve* carve;                      // [nWork] shared-memory layout of the middle's arrays in k_seed_middles
  uint32_t* slotCount;                   // [nWork] seeds per middle (zeroed here for middles without triplet stage)
  uint32_t* classList;                   // [kNumSeedClasses][classStride]
  uint32_t* classCount;                  // [kNumSeedClasses]
  uint32_t classStride;
  uint32_t classBytes[kNumSeedClasses];  // capacity of every shared-memory class (spill: unused)
  int conf;                              // seedConfirmation: sufficientTopDoublets after the top search
  unsigned long long* counters;
  int* status;
  // Hand-off from the count pass to the fill pass: the r windows of every neighbour bin of a middle and the survivor
  // bit masks of the (z, r) cuts (one word per 32 candidates), so that the fill pass neither searches the windows
  // nor loads and tests the candidates again.  Per middle: [2 * nWin window words][top mask words][bottom mask words]
  // at maskOff[w] (0xFFFFFFFF: the arena was full -- that middle is recomputed by the fill pass).
  uint32_t* maskArena;
  uint32_t maskCapacity;  // words
  uint32_t* maskCursor;
  uint32_t* maskOff;      // [nWork]
};

// r windows of the neighbour bins of one middle, searched by one warp (TripletSeeder.cpp:157-195,
// DoubletSeedFinder.cpp:73-93,105-122: monotone predicates over an r-sorted bin, evaluated with the
// reference's float expressions).  `near` = the arrays hold the windows of the preceding middle of the
// same bin (ascending r): they are lower bounds of the new ones.
struct WarpWindows {
  uint32_t s[2 * kMaxNeighborBins], e[2 * kMaxNeighborBins], hi[2 * kMaxNeighborBins];  // bottoms first, then tops
};

__device__ __forceinline__ void warp_windows(const DoubletParams& p, const uint32_t* bs, uint32_t botBeg, uint32_t nBot,
                                             uint32_t topBeg, uint32_t nTop, float rM, float firstMiddleR, bool near,
                                             WarpWindows& W) {
  const DeviceConfig& cfg = p.cfg;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t k = 0; k < nBot + nTop; ++k) {
    uint32_t s, e, b1;
    if (k < nBot) {
      auto inMax = [&](uint32_t i) { return fsub(rM, ldg2(p.pZR + i).y) <= cfg.dRMaxB; };
      auto belowMin = [&](uint32_t i) { return fsub(rM, ldg2(p.pZR + i).y) < cfg.dRMinB; };
      if (near) {
        b1 = W.hi[k];
        s = warp_first_true_near(W.s[k], b1, inMax);
        const uint32_t e0 = W.e[k] > s ? W.e[k] : s;
        e = warp_first_true_near(e0, b1, belowMin);
      } else {
        const uint32_t bin = __ldg(p.botBins + botBeg + k);
        const uint32_t b0 = bs[bin];
        b1 = bs[bin + 1];
        const float trimValue = fsub(firstMiddleR, cfg.dRMaxB);
        const uint32_t trim = warp_first_true(b0, b1, [&](uint32_t i) { return !(ldg2(p.pZR + i).y < trimValue); });
        s = warp_first_true(trim, b1, inMax);
        e = warp_first_true(s, b1, belowMin);
      }
    } else {
      auto inMin = [&](uint32_t i) { return fsub(ldg2(p.pZR + i).y, rM) >= cfg.dRMinT; };
      auto aboveMax = [&](uint32_t i) { return fsub(ldg2(p.pZR + i).y, rM) > cfg.dRMaxT; };
      if (near) {
        b1 = W.hi[k];
        s = warp_first_true_near(W.s[k], b1, inMin);
        const uint32_t e0 = W.e[k] > s ? W.e[k] : s;
        e = warp_first_true_near(e0, b1, aboveMax);
      } else {
        const uint32_t bin = __ldg(p.topBins + topBeg + (k - nBot));
        const uint32_t b0 = bs[bin];
        b1 = bs[bin + 1];
        const float trimValue = fadd(firstMiddleR, cfg.dRMinT);
        const uint32_t trim = warp_first_true(b0, b1, [&](uint32_t i) { return !(ldg2(p.pZR + i).y < trimValue); });
        s = warp_first_true(trim, b1, inMin);
        e = warp_first_true(s, b1, aboveMax);
      }
    }
    __syncwarp();
    if (lane == 0) { W.s[k] = s; W.e[k] = e; W.hi[k] = b1; }
  }
  __syncwarp();
}

constexpr int kDoubletWarps = 8;
#ifndef B200SEED_MASK_PREFETCH
#define B200SEED_MASK_PREFETCH 0
#endif
constexpr int kDoubletQueue = 160;  // < 32 left over + up to 128 new survivors per step

// One side of one middle.  Count pass: returns the number of (z, r) survivors.  Fill pass: survivors are
// queued (warp-private queue of kDoubletQueue positions) and finished 32 at a time; returns the number of doublets written.
template <bool kBottom, bool kFill>
__device__ __forceinline__ uint32_t doublet_side(const DoubletParams& p, const MiddleSp& mid, const uint32_t* winS,
                                                 const uint32_t* winE, uint32_t nWin, uint32_t* queue,
                                                 DoubletRecord* recOut, float* keyOut, float& cotMin, float& cotMax,
                                                 const float* zLo, const float* zHi, int nZ, uint32_t* masks = nullptr,
                                                 uint32_t nMaskWords = 0) {
  // `masks`: count pass -- where to write the survivor words (NULL: nowhere); fill pass -- where to read them from
  // (NULL: load and test the candidates like the count pass did)
  const DeviceConfig& cfg = p.cfg;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t ltMask = (1u << lane) - 1u;
  uint32_t myCount = 0;  // count pass: per lane
  uint32_t wi = 0;       // mask word index
  uint32_t qn = 0, nOut = 0;
  float mn = 3.0e38f, mx = -3.0e38f;
  auto drain = [&](uint32_t n) {  // finish the first n (<= 32) queued candidates
    bool ok = false;
    DoubletRecord out;
    if (lane < n) {
      const uint32_t o = queue[lane];
      const float2 zr = ldg2(p.pZR + o), xy = ldg2(p.pXY + o), var = ldg2(p.pVar + o);
      float dR, dZ;
      doublet_zr_cuts<kBottom>(cfg, mid, zr.x, zr.y, dR, dZ);
      DoubletRec rec;
      ok = doublet_finish<kBottom>(cfg, mid, dR, dZ, xy.x, xy.y, zr.y, var.x, var.y, zLo, zHi, nZ, rec, true);
      out.pos = o; out.cotTheta = rec.cotTheta; out.iDeltaR = rec.iDeltaR; out.er = rec.er;
      out.u = rec.u; out.v = rec.v; out.xNew = rec.xNew; out.yNew = rec.yNew;
    }
    __syncwarp();  // the queue entries of this step are read: later steps (or the other side of the middle) may overwrite them
    const uint32_t mask = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const uint32_t d = nOut + (uint32_t)__popc(mask & ltMask);
      float4* dst = reinterpret_cast<float4*>(recOut + d);
      dst[0] = make_float4(__uint_as_float(out.pos), out.cotTheta, out.iDeltaR, out.er);
      dst[1] = make_float4(out.u, out.v, out.xNew, out.yNew);
      keyOut[d] = out.cotTheta;
      mn = fminf(mn, out.cotTheta);
      mx = fmaxf(mx, out.cotTheta);
    }
    nOut += (uint32_t)__popc(mask);
    const uint32_t rest = qn - n;
    for (uint32_t base = 0; base < rest; base += 32u) {  // move the rest to the front (reads of a step precede its writes)
      const uint32_t i = base + lane;
      const uint32_t carry = i < rest ? queue[i + n] : 0u;
      __syncwarp();
      if (i < rest) queue[i] = carry;
      __syncwarp();
    }
    qn = rest;
  };
#if B200SEED_MASK_PREFETCH
  uint4 mNext = make_uint4(0u, 0u, 0u, 0u);  // the survivor words of the next step, requested one step ahead
  if (kFill && masks != nullptr && nMaskWords != 0u) mNext = __ldg(reinterpret_cast<const uint4*>(masks));
#endif
  for (uint32_t k = 0; k < nWin; ++k) {
    const uint32_t s = winS[k], e = winE[k];
    for (uint32_t base = s; base < e; base += 128u) {
      if (kFill && masks != nullptr) {  // the count pass left the survivor words: no candidate is loaded or tested again
#if B200SEED_MASK_PREFETCH
        const uint4 m4 = mNext;
        wi += 4;
        if (wi < nMaskWords) mNext = __ldg(reinterpret_cast<const uint4*>(masks + wi));
#else
        const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(masks + wi));  // (4 words per step, 16-byte aligned)
        wi += 4;
#endif
        const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t mask = mw[u];
          if ((mask >> lane) & 1u) queue[qn + (uint32_t)__popc(mask & ltMask)] = base + 32u * (uint32_t)u + lane;
          qn += (uint32_t)__popc(mask);
        }
      } else {
      float2 zr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t o = base + 32u * (uint32_t)u + lane;
        zr[u] = o < e ? ldg2(p.pZR + o) : make_float2(0.f, 0.f);
      }
      uint32_t mw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t o = base + 32u * (uint32_t)u + lane;
        float dR, dZ;
        const bool pass = o < e && doublet_zr_cuts<kBottom>(cfg, mid, zr[u].x, zr[u].y, dR, dZ);
        if (!kFill) {
          if (masks != nullptr) {
            mw[u] = __ballot_sync(0xffffffffu, pass);
            myCount += lane == 0 ? (uint32_t)__popc(mw[u]) : 0u;
          } else {
            myCount += pass ? 1u : 0u;
          }
        } else {
          const uint32_t mask = __ballot_sync(0xffffffffu, pass);
          if (pass) queue[qn + (uint32_t)__popc(mask & ltMask)] = o;
          qn += (uint32_t)__popc(mask);
        }
      }
      if (!kFill && masks != nullptr) {
        if (lane == 0) *reinterpret_cast<uint4*>(masks + wi) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
        wi += 4;
      }
      }
      if (kFill) {
        __syncwarp();
        while (qn >= 32u) drain(32u);  // qn < 32 + 128 before: kDoubletQueue entries suffice
      }
    }
  }
  if (!kFill) {
    for (int d = 16; d > 0; d >>= 1) myCount += __shfl_xor_sync(0xffffffffu, myCount, d);
    return myCount;
  }
  if (qn > 0u) drain(qn);
  for (int d = 16; d > 0; d >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  }
  cotMin = mn;
  cotMax = mx;
  return nOut;
}

// Measured and not kept for the fill pass (16-event batch, 9.6 ms as it stands; profiles/README.md, session 4):
//   register caps -- __launch_bounds__(256, 3 / 4 / 5): 13.3 / 11.9 / 13.1 ms; an explicit (256, 1) lets ptxas take 173 registers
//   a ring queue instead of moving the left-over entries to the front + the mask words of 32 steps loaded by one
//   coalesced request and broadcast by shuffles: 10.1 ms (108 registers), 10.7 / 11.4 ms capped to 80 / 96
//   two queued candidates per lane and drain step (both gathers in flight before the first use): 15.1 ms (173 registers)
template <bool kFill>
__global__ void __launch_bounds__(kDoubletWarps * 32) k_doublets(const __grid_constant__ DoubletParams p) {
  __shared__ WarpWindows sWin[kDoubletWarps];
  __shared__ uint32_t sQueue[kDoubletWarps][kDoubletQueue];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const DeviceConfig& cfg = p.cfg;
  WarpWindows& W = sWin[wib];
  uint32_t* queue = sQueue[wib];
  const uint32_t itemFirst = kFill ? p.itemFirst : 0u;
  const uint32_t itemEnd = kFill ? p.itemEnd : *p.nWorkPtr;
  const uint32_t nItems = itemEnd - itemFirst;
  const uint32_t nWarpsGrid = gridDim.x * (uint32_t)kDoubletWarps;
  unsigned long long cntMiddles = 0, cntB = 0, cntT = 0;
  uint32_t maxFoot = 0, maxB = 0, maxT = 0;
  uint32_t prevEG = 0xFFFFFFFFu, prevW = 0xFFFFFFF0u;
  for (;;) {
    // a warp takes B200SEED_WORK_CHUNK consecutive items (middles of one bin in ascending r: window hints), single
    // items near the end of the list keep the tail balanced
    uint32_t it0 = 0, it1 = 0;
    if (lane == 0) {
      const uint32_t seen = *reinterpret_cast<volatile uint32_t*>(p.workCounter);
      const uint32_t chunk = (seen < nItems && nItems - seen > nWarpsGrid * (uint32_t)(2 * B200SEED_WORK_CHUNK))
                                 ? (uint32_t)B200SEED_WORK_CHUNK : 1u;
      it0 = atomicAdd(p.workCounter, chunk);
      it1 = it0 + chunk < nItems ? it0 + chunk : nItems;
    }
    it0 = __shfl_sync(0xffffffffu, it0, 0);
    it1 = __shfl_sync(0xffffffffu, it1, 0);
    if (it0 >= nItems) break;
    for (uint32_t it = it0; it < it1; ++it) {
      const uint32_t w = itemFirst + it;
      uint32_t capT = 0, capB = 0;
      if (kFill) {
        capT = __ldg(p.capT + w);
        capB = __ldg(p.capB + w);
        if (capT == 0u || capB == 0u) {  // no triplet stage (TripletSeeder.cpp:62,79)
          if (lane == 0) {
            MiddleHeader h{};
            h.capB = capB;
            p.hdr[w] = h;
            p.slotCount[w] = 0;
          }
          continue;
        }
      }
      const uint32_t m = __ldg(p.workPos + w);
      const uint32_t eg = __ldg(p.workEG + w);
      const uint32_t ev = eg / p.nNav, g = eg - ev * p.nNav;
      const uint32_t* bs = p.binStart + (size_t)ev * p.nBins;
      const uint32_t botBeg = __ldg(p.botOffsets + g), nBot = __ldg(p.botOffsets + g + 1) - botBeg;
      const uint32_t topBeg = __ldg(p.topOffsets + g), nTop = __ldg(p.topOffsets + g + 1) - topBeg;
      MiddleSp mid;
      {
        const float2 mxy = ldg2(p.pXY + m), mzr = ldg2(p.pZR + m), mvar = ldg2(p.pVar + m);
        mid.x = mxy.x; mid.y = mxy.y; mid.z = mzr.x; mid.r = mzr.y; mid.varZ = mvar.x; mid.varR = mvar.y;
        middle_info(mid);
      }
      const uint32_t nWin = nBot + nTop;
      const uint32_t nWinWords = (2u * nWin + 3u) & ~3u;  // the mask words that follow stay 16-byte aligned
      uint32_t moff = 0xFFFFFFFFu;
      if (kFill && p.maskArena != nullptr) moff = __ldg(p.maskOff + w);
      if (moff != 0xFFFFFFFFu) {
        // the count pass left this middle's windows (and survivor masks) in the arena
        for (uint32_t k = lane; k < nWin; k += 32u) {
          const uint2 se = __ldg(reinterpret_cast<const uint2*>(p.maskArena + moff) + k);
          W.s[k] = se.x;
          W.e[k] = se.y;
        }
        __syncwarp();
        prevEG = 0xFFFFFFFFu;  // (W.hi is not restored: the next searched middle starts from scratch)
      } else {
        const float firstMiddleR = ldg2(p.pZR + bs[__ldg(p.navBins + g)]).y;
        const bool near = eg == prevEG && w == prevW + 1u;
        prevEG = eg;
        prevW = w;
        warp_windows(p, bs, botBeg, nBot, topBeg, nTop, mid.r, firstMiddleR, near, W);
      }
      // 128-candidate steps of the top windows (= 4 mask words each): where the bottom masks start
      uint32_t stepsT = 0, stepsB = 0;
      if (p.maskArena != nullptr) {
        for (uint32_t k = lane; k < nWin; k += 32u) {
          const uint32_t st = (W.e[k] - W.s[k] + 127u) >> 7;
          if (k < nBot) stepsB += st; else stepsT += st;
        }
        for (int d = 16; d > 0; d >>= 1) {
          stepsT += __shfl_xor_sync(0xffffffffu, stepsT, d);
          stepsB += __shfl_xor_sync(0xffffffffu, stepsB, d);
        }
      }
      if (!kFill) {
        ++cntMiddles;
        float a, b;
        uint32_t *maskT = nullptr, *maskB = nullptr;
        if (p.maskArena != nullptr) {
          const uint32_t words = nWinWords + 4u * (stepsT + stepsB);
          uint32_t off = 0;
          if (lane == 0) off = atomicAdd(p.maskCursor, words);
          off = __shfl_sync(0xffffffffu, off, 0);
          const bool fits = off <= p.maskCapacity && words <= p.maskCapacity - off;
          if (fits) {
            for (uint32_t k = lane; k < nWin; k += 32u) {
              reinterpret_cast<uint2*>(p.maskArena + off)[k] = make_uint2(W.s[k], W.e[k]);
            }
            maskT = p.maskArena + off + nWinWords;
            maskB = maskT + 4u * stepsT;
          }
          if (lane == 0) p.maskOff[w] = fits ? off : 0xFFFFFFFFu;
        }
        capT = doublet_side<false, false>(p, mid, W.s + nBot, W.e + nBot, nTop, queue, nullptr, nullptr, a, b, nullptr, nullptr, 0, maskT);
        if (capT != 0u) capB = doublet_side<true, false>(p, mid, W.s, W.e, nBot, queue, nullptr, nullptr, a, b, nullptr, nullptr, 0, maskB);
        if (capB == 0u) capT = 0u;  // a middle without bottoms or without tops costs nothing later
        if (capB > kMaxListLength || capT > kMaxListLength) {
          if (lane == 0) atomicOr(p.status, kStatusOverflowDoublets);
          capB = 0; capT = 0;
        }
        capB = (capB + 3u) & ~3u;  // slots and their two halves start on 16-byte boundaries of the key array (TMA)
        capT = (capT + 3u) & ~3u;
        if (lane == 0) { p.capB[w] = capB; p.capT[w] = capT; }
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
        const uint32_t zw0 = p.zWinOffsets != nullptr ? __ldg(p.zWinOffsets + ev) : 0u;
        const int nZ = p.zWinOffsets != nullptr ? (int)(__ldg(p.zWinOffsets + ev + 1) - zw0) : p.nZWin;
        const float *zLo = p.zWinLo + zw0, *zHi = p.zWinHi + zw0;
        uint32_t *maskT = nullptr, *maskB = nullptr;
        if (moff != 0xFFFFFFFFu) {
          maskT = p.maskArena + moff + nWinWords;
          maskB = maskT + 4u * stepsT;
        }
        const uint32_t nT = doublet_side<false, true>(p, mid, W.s + nBot, W.e + nBot, nTop, queue, recSlot + capB, keySlot + capB, mnT, mxT, zLo, zHi, nZ, maskT, 4u * stepsT);
        bool go = nT != 0u;
        if (go && p.conf) go = !(nT < conf_n_top(conf_range(cfg, mid.z), mid.r));  // BroadTripletSeedFilter.cpp:63-94
        uint32_t nB = 0;
        if (go) nB = doublet_side<true, true>(p, mid, W.s, W.e, nBot, queue, recSlot, keySlot, mnB, mxB, zLo, zHi, nZ, maskB, 4u * stepsB);
        go = go && nB != 0u;
        if (lane == 0) {
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
          } else {
            p.slotCount[w] = 0;
          }
          p.hdr[w] = h;
        }
        if (go) { cntB += nB; cntT += nT; }
      }
    }
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

// slot sizes -> 64-bit exclusive prefix (tiled: per-tile sums, scan of the tile sums, per-tile scan)
struct SlotScanParams {
  const uint32_t* nWorkPtr;
  const uint32_t *capB, *capT;
  unsigned long long* tileSums;    // [nTiles + 1]
  unsigned long long* tilePrefix;  // [nTiles + 1]
  unsigned long long* slotPrefix;  // [nWork + 1]
  // chunk plan
  unsigned long long arenaRecords;  // capacity of the arena in records
  uint32_t* chunkBounds;            // [kMaxChunks + 1]
  uint32_t* planWords;              // [3] number of chunks, [4] nWork
  int* status;
};

__global__ void __launch_bounds__(256) k_cap_tile_sums(const __grid_constant__ SlotScanParams p) {
  __shared__ uint32_t scratch[34];
  const uint32_t nWork = *p.nWorkPtr;
  uint32_t s = 0;
  for (uint32_t i = threadIdx.x; i < (uint32_t)kTile; i += blockDim.x) {
    const uint32_t w = blockIdx.x * kTile + i;
    if (w < nWork) s += p.capB[w] + p.capT[w];  // <= 2048 * 2 * 65534: fits 32 bits
  }
  uint32_t total;
  block_scan_exclusive(s, scratch, total, OpSum());
  if (threadIdx.x == 0) p.tileSums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_u64(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, uint32_t n) {
  // n tile sums (a few hundred): one warp, serial over chunks of 32
  if (threadIdx.x >= 32) return;
  const uint32_t lane = threadIdx.x;
  unsigned long long carry = 0;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t i = base + lane;
    unsigned long long v = i < n ? in[i] : 0ull, incl = v;
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (uint32_t)d) incl += o;
    }
    if (i < n) out[i] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) out[n] = carry;
}

__global__ void __launch_bounds__(256) k_cap_prefix(const __grid_constant__ SlotScanParams p) {
  __shared__ uint32_t scratch[34];
  __shared__ uint32_t carry;
  const uint32_t nWork = *p.nWorkPtr;
  const uint32_t tile = blockIdx.x;
  if (tile * (uint32_t)kTile >= nWork && !(tile == 0 && nWork == 0)) return;
  const unsigned long long tileBase = p.tilePrefix[tile];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < (uint32_t)kTile; base += blockDim.x) {
    const uint32_t w = tile * kTile + base + threadIdx.x;
    const uint32_t n = w < nWork ? p.capB[w] + p.capT[w] : 0u;
    uint32_t total;
    const uint32_t excl = block_scan_exclusive(n, scratch, total, OpSum());
    const uint32_t c = carry;
    if (w < nWork) p.slotPrefix[w] = tileBase + c + excl;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && (tile + 1) * (uint32_t)kTile >= nWork) p.slotPrefix[nWork] = tileBase + carry;
}

// Greedy chunks of consecutive work items whose slots fit the arena (one thread: a binary search per chunk).
__global__ void k_plan_chunks(const __grid_constant__ SlotScanParams p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint32_t nWork = *p.nWorkPtr;
  uint32_t first = 0, n = 0;
  p.chunkBounds[0] = 0;
  while (first < nWork) {
    const unsigned long long limit = p.slotPrefix[first] + p.arenaRecords;
    uint32_t lo = first + 1, hi = nWork;  // the last end in [first + 1, nWork] with slotPrefix[end] <= limit
    if (p.slotPrefix[lo] > limit) {
      atomicOr(p.status, kStatusArenaTooSmall);  // a single middle larger than the arena (host sizes it for 2 * kMaxListLength)
    } else {
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo + 1) >> 1);
        if (p.slotPrefix[mid] <= limit) lo = mid; else hi = mid - 1;
      }
    }
    if (n == kMaxChunks) {
      atomicOr(p.status, kStatusArenaTooSmall);
      break;
    }
    p.chunkBounds[++n] = lo;
    first = lo;
  }
  p.planWords[3] = n;
  p.planWords[4] = nWork;
  // records of the largest chunk (the arena the host reserves): read back with the plan, no second round trip
  unsigned long long largest = 0;
  for (uint32_t c = 0; c < n; ++c) {
    const unsigned long long r = p.slotPrefix[p.chunkBounds[c + 1]] - p.slotPrefix[p.chunkBounds[c]];
    largest = r > largest ? r : largest;
  }
  p.planWords[5] = (uint32_t)(largest & 0xFFFFFFFFull);
  p.planWords[6] = (uint32_t)(largest >> 32);
}

// ---------------------------------------------------------------------------
// Seeding kernel: one block per middle space point (persistent blocks pulling
// the work list of one shared-memory class).  The doublets of the middle come
// from the arena written by k_doublets<true>; every array below is sized by the
// middle's exact list lengths (SeedCarve).  Per middle:
//   phase 1  cotTheta keys of both lists -> shared memory
//   phase 2  order both lists like the reference's sortByCotTheta (bucket sort
//            by (cotTheta, emission index) + pruned libstdc++ introsort replay for
//            ties); sorted tops are gathered into shared memory, |P_j| (tops with
//            cotT <= cotB_j) is found for every bottom
//   phase 3  a) H_j / brk_j scans, one lane per bottom with lane refill: a lane
//               that finishes its bottom takes the next one, every pair is
//               evaluated once and candidates are emitted on the spot
//                                                  (TripletSeedFinder.cpp:34-162)
//            b) window starts = exclusive prefix max of H
//            c) the few pairs in [start_j, t*_j) the scans did not touch
//            d) candidates grouped by bottom (counting sort), each group in
//               curvature order                  (BroadTripletSeedFilter.cpp:143-148)
//            e) one thread per candidate: weight  (BroadTripletSeedFilter.cpp:162-251)
//            f) bounded heap replay in the reference's push order
//               (CandidatesForMiddleSp.cpp:44-75)
//   phase 4  sort_heap, keep min(n, maxSeedsPerSpM + 1), write the seed slots
// ---------------------------------------------------------------------------
struct SeedParams {
  DeviceConfig cfg;
  const float2 *pXY, *pZR, *pVar;
  const uint32_t* workPos;
  // the chunk's doublets
  const MiddleHeader* hdr;
  const SeedCarve* carve;
  const DoubletRecord* rec;
  const float* key;
  // work list of this launch (one shared-memory class of one chunk)
  const uint32_t* workList;
  const uint32_t* nWorkPtr;
  uint32_t* workCounter;
  uint32_t* overflowList;   // middles whose candidate pool did not fit: next class (NULL: error)
  uint32_t* overflowCount;
  uint32_t arrayBytes;      // bytes of dynamic shared memory (spill class: of global scratch) per block
  unsigned char* spillScratch;
  uint32_t *slotB, *slotM, *slotT;
  float *slotQ, *slotZ;
  uint32_t* slotCount;
  uint32_t seedsPerMiddle;
  int exactTies;  // replay libstdc++ std::sort inside groups of equal cotTheta
  unsigned long long* counters;
  int* status;
  // seedConfirmation only: the weighted candidates of every middle, in the reference's push order
  uint4* rec4;           // {bottom pos, top pos, weight bits, meta}
  float* recZ;           // zOrigin
  uint32_t *recBegin, *recCount;  // per work item
  uint32_t* recCounter;  // bump allocator (keeps counting past the capacity: tells the host what to reserve)
  uint32_t recCapacity;  // strip triplet path only (TripletSeedFinder::Config::useStripInfo): derived calibration details per packed
  // position, cotThetaDiffMax^2, toleranceParam
  const StripDerived* pStrip;
  float cotThetaDiffMax2;
  float toleranceParam;
  // maxSeedsPerSpMConf > kMaxHeap: the literal heap replay keeps its arrays (16 bytes per entry) in the dynamic shared
  // memory behind the per-middle arrays (at arrayBytes; spill class: at 0) instead of SeedShared
  uint32_t bigHeap;
};

// meta word of a candidate record
constexpr uint32_t kRecGroupMask = 0xFFFFu;      // sorted rank of the bottom = group id
constexpr int kRecGroupSizeShift = 16;           // min(#candidates of the group, 3)
constexpr uint32_t kRecNeedsTwoTops = 1u << 18;  // bottom radius <= rMaxSeedConf of the middle's region
constexpr uint32_t kRecQuality = 1u << 19;       // deltaSeedConf > 0
constexpr uint32_t kRecKeep = 1u << 31;          // in-kernel only

struct Cand {
  float curv;
  float impactOrWeight;
  float topR;
  uint32_t tOwner;  // sorted top rank | sorted bottom rank << 16
};
__device__ __forceinline__ bool cand_less(const Cand& a, const Cand& b) { return a.curv < b.curv; }

struct StoredSeed {
  uint32_t tOwner;  // candidate identity (sorted top rank | sorted bottom rank << 16)
  float weight;
};

struct SeedShared {
  uint32_t w;
  uint32_t poolCount, nSurv;
  uint32_t tie, tieB, tieT;
  uint32_t nextBottom;
  uint32_t runCarry;
  uint32_t scratch[34];
  WeightIndex heap[kMaxHeap];
  StoredSeed storage[kMaxHeap];
  int heapSize;
  int heapSorted;
  float heapMin;
  unsigned long long cnt[kCntSlots];
};

// Combined bucket sort of the bottom and top doublet lists of one middle by
// (cotTheta, emission index): the bottoms use buckets [0, nBk / 2), the tops
// [nBk / 2, nBk), each list spread over its own cotTheta range.  rankB[0, nB) /
// rankT[0, nT) are the lists in sorted order (values index the per-side key
// arrays = the arena slot).  *tieB / *tieT are set when a list holds equal
// neighbours (equal keys always share a bucket).
__device__ __forceinline__ void block_sort_both(uint32_t nB, const float* cotB, float minB, float maxB, uint32_t nT,
                                                const float* cotT, float minT, float maxT, uint16_t* rankB,
                                                uint16_t* rankT, uint32_t* buckets, uint32_t nBk, uint32_t* scratch,
                                                uint32_t* tieB, uint32_t* tieT) {
  const uint32_t half = nBk >> 1;
  const float scaleB = maxB > minB ? (float)half / (maxB - minB) : 0.0f;
  const float scaleT = maxT > minT ? (float)half / (maxT - minT) : 0.0f;
  auto bucketOf = [&](uint32_t e) {
    const bool bottom = e < nB;
    const float c = bottom ? cotB[e] : cotT[e - nB];
    float t = bottom ? fmul(fsub(c, minB), scaleB) : fmul(fsub(c, minT), scaleT);  // monotone in c
    if (!(t > 0.0f)) t = 0.0f;
    uint32_t b = (uint32_t)t;
    if (b >= half) b = half - 1;
    return bottom ? b : half + b;
  };
  const uint32_t n = nB + nT;
  for (uint32_t i = threadIdx.x; i <= nBk; i += blockDim.x) buckets[i] = 0;
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) atomicAdd(buckets + bucketOf(e), 1u);
  __syncthreads();
  block_scan_array(buckets, nBk, scratch);
  for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
    const uint32_t q = atomicAdd(buckets + bucketOf(e), 1u);  // bottoms fill [0, nB), tops [nB, n)
    if (e < nB) rankB[q] = (uint16_t)e; else rankT[q - nB] = (uint16_t)(e - nB);
  }
  __syncthreads();  // buckets[b] is now the END of bucket b
  for (uint32_t b = threadIdx.x; b < nBk; b += blockDim.x) {
    const uint32_t s = b == 0 ? 0u : buckets[b - 1], e = buckets[b];
    if (e - s < 2) continue;
    const bool bottom = b < half;
    const float* cot = bottom ? cotB : cotT;
    uint16_t* rank = bottom ? rankB : rankT - nB;
    bool tie = false;
    for (uint32_t i = s + 1; i < e; ++i) {
      const uint16_t v = rank[i];
      const float cv = cot[v];
      uint32_t j = i;
      while (j > s) {
        const uint16_t u = rank[j - 1];
        const float cu = cot[u];
        tie |= cu == cv;
        if (cu > cv || (cu == cv && u > v)) {
          rank[j] = u;
          --j;
        } else {
          break;
        }
      }
      rank[j] = v;
    }
    if (tie) *(bottom ? tieB : tieT) = 1u;
  }
  __syncthreads();
}

// cotTheta tie items: val = element index | tie group (first canonical rank) << 16; group 0xFFFF = unique key
__device__ __forceinline__ bool tie_flagged(const TieItem& a) { return (a.val >> 16) != 0xFFFFu; }

// Exact order inside groups of equal cotTheta: the reference sorts the doublets
// with the unstable std::ranges::sort (DoubletSeedFinder.hpp:94-104) starting
// from the emission order = the order of the arena slot.  `sorted` holds the
// canonical (cot, index) order; the pruned replay of libstdc++'s introsort
// (warp_sort_replay_ties) run by one warp on the emission-ordered copy W decides
// which member of a tie group takes which of the group's slots.
__device__ __forceinline__ void block_fix_ties(uint32_t n, const float* cot, uint16_t* sorted, TieItem* W, uint16_t* grpOf) {
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
    TieItem it;
    it.key = cot[r];
    it.val = r | ((uint32_t)grpOf[r] << 16);
    W[r] = it;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    warp_sort_replay_ties(W, (int)n, tie_flagged);
    // member k (in W order) of group g goes to canonical slot g + k.  The ranks come from a ballot scan over
    // W; grpOf (dead once W is built) is reused as the running member count of every group, indexed by g.
    const uint32_t lane = threadIdx.x;
    for (uint32_t i = lane; i < n; i += 32) grpOf[i] = 0;
    __syncwarp();
    for (uint32_t base = 0; base < n; base += 32) {
      const uint32_t r = base + lane;
      TieItem it;
      it.key = 0.f;
      it.val = 0xFFFF0000u;
      if (r < n) it = W[r];
      const uint32_t g = it.val >> 16;
      const bool flagged = g != 0xFFFFu;
      const uint32_t same = __match_any_sync(0xffffffffu, flagged ? g : 0x80000000u + lane);  // unflagged lanes match themselves only
      const uint32_t before = (uint32_t)__popc(same & ((1u << lane) - 1u));
      uint32_t carried = 0;
      if (flagged && before == 0u) {  // first member of its group in this step: one lane per group, no race
        carried = grpOf[g];
        grpOf[g] = (uint16_t)(carried + (uint32_t)__popc(same));
      }
      carried = __shfl_sync(0xffffffffu, carried, __ffs(same) - 1);
      if (flagged) sorted[g + carried + before] = (uint16_t)(it.val & 0xFFFFu);
      __syncwarp();
    }
  }
  __syncthreads();
}

// The block size is a launch parameter (any multiple of 32 up to 1024): every class runs the same code with the
// thread count that fills the SM next to its shared-memory footprint (seeding_plugin.cu: kSeedClassShape).
#ifndef B200SEED_SEED_REGS
#define B200SEED_SEED_REGS 56  // 36 warps per SM; measured against 64 (32 warps) and 48 (42 warps, spills)
#endif
// Per-middle array inside the block's dynamic shared memory, addressed by a 32-bit byte offset from the shared
// window (one register per array; the accesses are LDS / STS with the offset folded into the address), or, for the
// spill class, a plain pointer into the block's global scratch.
template <typename T, bool kSpill>
struct Arr {
  uint32_t off;
  T* glob;
  __device__ __forceinline__ Arr(unsigned char* base, uint32_t byteOffset) : off(byteOffset), glob(reinterpret_cast<T*>(base + byteOffset)) {}
  __device__ __forceinline__ T* ptr() const {
    if constexpr (kSpill) {
      return glob;
    } else {
      extern __shared__ __align__(16) unsigned char smemRaw[];
      return reinterpret_cast<T*>(smemRaw + off);
    }
  }
  __device__ __forceinline__ T& operator[](uint32_t i) const { return ptr()[i]; }
};

// ---- TMA bulk copies (cp.async.bulk, 1-D, global -> shared) completing on an mbarrier ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bytes: multiple of 16; dst and src 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra MBAR_WAIT_%=;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// generic-proxy accesses to shared memory before, async-proxy (TMA) writes to the same bytes after
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// kStrip: the strip triplet path (TripletSeedFinder.cpp:164-406) in place of phases 3a-3c; everything else is shared.
template <bool kConf, bool kSpill, bool kStrip = false>
__global__ void __maxnreg__(B200SEED_SEED_REGS) k_seed_middles(const __grid_constant__ SeedParams p) {
  const uint32_t THREADS = blockDim.x;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ SeedShared sh;
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31;
  const uint32_t ltMask = (1u << lane) - 1u;
  const DeviceConfig& cfg = p.cfg;
  const uint32_t nWork = *p.nWorkPtr;
  unsigned char* const base = kSpill ? p.spillScratch + (size_t)blockIdx.x * p.arrayBytes : smemRaw;

  __shared__ __align__(8) uint64_t keyBar;  // completion of the TMA copies of the cotTheta keys
  uint32_t keyParity = 0;
  if (tid < (uint32_t)kCntSlots) sh.cnt[tid] = 0ull;
  if (tid == 0) {
    const uint32_t item = atomicAdd(p.workCounter, 1u);
    sh.w = item < nWork ? p.workList[item] : 0xFFFFFFFFu;
    if (!kSpill) mbar_init(&keyBar, 1);
  }

  for (;;) {
    __syncthreads();
    const uint32_t w = sh.w;
    if (w == 0xFFFFFFFFu) break;
    if (!kSpill) fence_proxy_async();  // this thread's accesses to the previous middle's arrays, before the TMA writes below
    __syncthreads();  // everybody has read sh.w before thread 0 fetches the next item

    // ---- phase 0: header, middle, carve-up ---------------------------------
    const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(p.hdr + w));
    const int4 h1 = __ldg(reinterpret_cast<const int4*>(p.hdr + w) + 1);
    const uint32_t nB = h0.x, nT = h0.y;
    const DoubletRecord* recB = p.rec + h0.w;
    const DoubletRecord* recT = recB + h0.z;
    const float* gKeyB = p.key + h0.w;
    const float* gKeyT = gKeyB + h0.z;
    const uint32_t m = __ldg(p.workPos + w);
    MiddleSp mid;
    {
      const float2 mxy = ldg2(p.pXY + m), mzr = ldg2(p.pZR + m), mvar = ldg2(p.pVar + m);
      mid.x = mxy.x; mid.y = mxy.y; mid.z = mzr.x; mid.r = mzr.y; mid.varZ = mvar.x; mid.varR = mvar.y;
      middle_info(mid);
    }
    SeedCarve cv;  // computed once per middle by the fill pass
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.carve + w);
      uint4* dst = reinterpret_cast<uint4*>(&cv);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = __ldg(src + q);
    }
    const Arr<uint16_t, kSpill> rankB(base, cv.oRankB), tstar(base, cv.oTstar);
    const Arr<float4, kSpill> sA(base, cv.oTops);  // sorted tops: {cotTheta, er, iDeltaR, u}, then v, then pos
    const Arr<float, kSpill> sV(base, cv.oTops + 16u * nT);
    const Arr<uint32_t, kSpill> sPos(base, cv.oTops + 20u * nT);
    const Arr<uint32_t, kSpill> buckets(base, cv.oBuckets);
    const Arr<float, kSpill> keyB(base, cv.oKeyB), keyT(base, cv.oKeyT);
    const Arr<uint16_t, kSpill> rankT(base, cv.oRankT);
    const Arr<TieItem, kSpill> tieW(base, cv.oTie);
    const Arr<uint16_t, kSpill> tieGrp(base, cv.oTieGrp);
    const Arr<uint16_t, kSpill> hval(base, cv.oHval);
    const Arr<uint32_t, kSpill> cnt(base, cv.oCnt);
    const uint32_t poolCap = (p.arrayBytes - cv.oPool) / kPoolEntryBytes;
    const Arr<uint32_t, kSpill> pool(base, cv.oPool);
    const Arr<Cand, kSpill> pool2(base, carve_align(cv.oPool + 4u * poolCap));

    if (tid == 0) {
      sh.tie = 0; sh.tieB = 0; sh.tieT = 0; sh.heapSize = 0; sh.poolCount = 0; sh.nextBottom = 0; sh.heapSorted = 0;
      const uint32_t item = atomicAdd(p.workCounter, 1u);  // next item: the latency hides behind this middle
      sh.w = item < nWork ? p.workList[item] : 0xFFFFFFFFu;
    }

#if B200SEED_PREFETCH_SLOT
    // The records of the slot are gathered later (tops after the sort, bottoms one by one in the scans): ask the
    // L2 for the whole slot now, one 128-byte line (4 records) per request.
    {
      const uint32_t nLines = (h0.z + nT + 3u) >> 2;  // bottoms [0, capB) and tops [capB, capB + nT)
      for (uint32_t i = tid; i < nLines; i += THREADS) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(recB + 4u * i));
      }
    }
#endif
    // ---- phase 1: keys ---------------------------------------------------------
    // Two TMA bulk copies stage the contiguous cotTheta key spans of the middle's arena slot in shared memory
    // (slots and list capacities are multiples of four records, so both spans are 16-byte aligned and padded).
    if constexpr (!kSpill) {
      if (tid == 0) {
        const uint32_t bytesB = ((nB + 3u) & ~3u) * 4u, bytesT = ((nT + 3u) & ~3u) * 4u;
        mbar_expect_tx(&keyBar, bytesB + bytesT);
        tma_bulk_g2s(keyB.ptr(), gKeyB, bytesB, &keyBar);
        tma_bulk_g2s(keyT.ptr(), gKeyT, bytesT, &keyBar);
      }
      mbar_wait(&keyBar, keyParity);
      keyParity ^= 1u;
    } else {
      for (uint32_t i = tid; i < nB; i += THREADS) keyB[i] = __ldg(gKeyB + i);
      for (uint32_t i = tid; i < nT; i += THREADS) keyT[i] = __ldg(gKeyT + i);
      __syncthreads();
    }

    // ---- phase 2: order both lists like DoubletSeedFinder.hpp:94-104 -------
    block_sort_both(nB, keyB.ptr(), ordered_to_float(h1.x), ordered_to_float(h1.y), nT, keyT.ptr(), ordered_to_float(h1.z),
                    ordered_to_float(h1.w), rankB.ptr(), rankT.ptr(), buckets.ptr(), cv.nBk, sh.scratch, &sh.tieB, &sh.tieT);
    {
      const bool tieB = sh.tieB != 0, tieT = sh.tieT != 0;  // block-uniform (read after the sort's last barrier)
      if ((tieB || tieT) && tid == 0) sh.tie = 1;
      if (tieB && p.exactTies && nB > 16) block_fix_ties(nB, keyB.ptr(), rankB.ptr(), tieW.ptr(), tieGrp.ptr());
      if (tieT && p.exactTies && nT > 16) block_fix_ties(nT, keyT.ptr(), rankT.ptr(), tieW.ptr(), tieGrp.ptr());
    }
    // tops: full records in sorted order
    for (uint32_t t = tid; t < nT; t += THREADS) {
      const float4* src = reinterpret_cast<const float4*>(recT + rankT[t]);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      // (strip path: the index of the top in the arena slot instead -- its record holds the position AND x', y')
      sPos[t] = kStrip ? (uint32_t)rankT[t] : __float_as_uint(a.x);
      sA[t] = make_float4(a.y, a.w, a.z, b.x);
      sV[t] = b.y;
    }
    __syncthreads();
    // |P_j|: tops with cotT <= cotB_j
    if constexpr (!kStrip)
    for (uint32_t j = tid; j < nB; j += THREADS) {
      const float c = keyB[rankB[j]];
      uint32_t lo = 0, hi = nT;
      while (lo < hi) {
        const uint32_t md = (lo + hi) >> 1;
        if (c < sA[md].x) hi = md; else lo = md + 1;
      }
      tstar[j] = (uint16_t)lo;
    }
    __syncthreads();  // region a (keys, rankT, tie scratch) is dead, region b starts

    // ---- phase 3a: H_j / brk_j scans, candidates emitted on the spot ------
    uint32_t myTests = 0;
    auto bottomCtx = [&](uint32_t j, BottomCtx& bc) {
      const float4* src = reinterpret_cast<const float4*>(recB + rankB[j]);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      bc.cotThetaB = a.y; bc.iDeltaRB = a.z; bc.erB = a.w; bc.Ub = b.x; bc.Vb = b.y;
      bottom_ctx(cfg, bc);
    };
    auto emit = [&](uint32_t j, uint32_t t) {
      // a plain shared-memory atomic per candidate: they are rare (0.03 per pair test), the compiler's warp
      // aggregation of atomicAdd costs more than it saves here
      uint32_t slot;
#if B200SEED_OPAQUE_EMIT
      // the address carries a per-lane zero: ptxas cannot prove it warp-uniform and emits one ATOMS per lane
      // instead of its vote / leader / popc / shuffle aggregation sequence
      const uint32_t addr = smem_u32(&sh.poolCount) + (j >> 16);  // j < 65536: + 0, but per-lane in the compiler's eyes
      asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(addr) : "memory");
#else
      asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(smem_u32(&sh.poolCount)) : "memory");
#endif
      if (slot < poolCap) pool[slot] = t | (j << 16);
    };
    // packed position of sorted top t
    auto topPos = [&](uint32_t t) -> uint32_t {
      if constexpr (kStrip) return __ldg(&recT[sPos[t]].pos); else return sPos[t];
    };
    // ---- strip path: phases 3a-3c are one plain scan ---------------------------------------------
    // The window of a bottom is the run of sorted tops with (cotB - cotT)^2 <= cotThetaDiffMax^2 (:226-238: tops
    // below the window are dropped for good, the loop ends at the first top above it; both tests are monotone in
    // the sorted order, so the window does not depend on the bottoms before).  One thread per bottom.
    float cosPhiM = 0.f, sinPhiM = 0.f;
    auto stripBottom = [&](uint32_t j, StripBottomCtx& sb) -> uint32_t {
      const float4* src = reinterpret_cast<const float4*>(recB + rankB[j]);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      sb.cotThetaB0 = a.y; sb.iDeltaRB = a.z; sb.erB = a.w; sb.Ub0 = b.x; sb.Vb0 = b.y; sb.xB = b.z; sb.yB = b.w;
      strip_bottom_ctx(cfg, cosPhiM, sinPhiM, sb);
      return __float_as_uint(a.x);
    };
    auto stripEval = [&](const StripBottomCtx& sb, const StripDerived& calM, const StripDerived& calB, uint32_t t, float& curv,
                         float& im) -> bool {
      const DoubletRecord* rt = recT + sPos[t];
      const float4 a = __ldg(reinterpret_cast<const float4*>(rt)), b = __ldg(reinterpret_cast<const float4*>(rt) + 1);
      StripDerived calT;
      {
        const uint4* src = reinterpret_cast<const uint4*>(p.pStrip + __float_as_uint(a.x));
        uint4* dst = reinterpret_cast<uint4*>(&calT);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = __ldg(src + q);
      }
      return eval_strip_pair(cfg, p.toleranceParam, mid.varZ, mid.varR, sb, calM, calB, calT, a.w, a.z, b.x, b.y, b.z, b.w, curv, im);
    };
    auto loadStrip = [&](uint32_t pos, StripDerived& out) {
      const uint4* src = reinterpret_cast<const uint4*>(p.pStrip + pos);
      uint4* dst = reinterpret_cast<uint4*>(&out);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = __ldg(src + q);
    };
    if constexpr (kStrip) {
      cosPhiM = fdiv(mid.x, mid.r);  // :171-172
      sinPhiM = fdiv(mid.y, mid.r);
      StripDerived calM;
      loadStrip(m, calM);
      for (uint32_t j = tid; j < nB; j += THREADS) {
        StripBottomCtx sb;
        const uint32_t posB = stripBottom(j, sb);
        StripDerived calB;
        loadStrip(posB, calB);
        uint32_t lo = 0, hi = nT;  // first top that is not below the window
        while (lo < hi) {
          const uint32_t md = (lo + hi) >> 1;
          const float cT = sA[md].x;
          if (strip_outside_window(sb.cotThetaB0, cT, p.cotThetaDiffMax2) && !(sb.cotThetaB0 < cT)) lo = md + 1; else hi = md;
        }
        for (uint32_t t = lo; t < nT; ++t) {
          const float cT = sA[t].x;
          if (strip_outside_window(sb.cotThetaB0, cT, p.cotThetaDiffMax2)) {
            if (sb.cotThetaB0 < cT) break;
            continue;
          }
          ++myTests;
          float cu, im;
          if (stripEval(sb, calM, calB, t, cu, im)) emit(j, t);
        }
        hval[j] = 0;
      }
      __syncthreads();
    } else {
#if B200SEED_SPLIT_WALKERS
    {
      // Two walkers per bottom, one lane each: the backward one goes down from |P_j| - 1 to the last failing top
      // of the prefix (-> H_j, t*_j), the forward one up from |P_j| to the first failing top beyond it (brk_j).
      // A lane that finishes takes the next unassigned walker as soon as kRefill lanes of its warp are idle, so the
      // warp stays full although the walks differ in length.  Both walkers of a bottom start in the same step (they
      // read |P_j| from tstar[j] before the backward one can overwrite it with t*_j).
      constexpr uint32_t kRefill = B200SEED_REFILL;
      bool active = false, exhausted = false;  // exhausted is warp-uniform
      uint32_t j = 0;
      int t = 0, tEnd = 0, step = 0;
      BottomCtx bc{};
      for (;;) {
        const uint32_t idleMask = __ballot_sync(0xffffffffu, !active);
        const uint32_t nIdle = (uint32_t)__popc(idleMask);
        if (!exhausted && nIdle >= kRefill) {
          const uint32_t nPairs = nIdle >> 1;
          uint32_t first = 0;
          if (lane == 0) first = atomicAdd(&sh.nextBottom, nPairs);
          first = __shfl_sync(0xffffffffu, first, 0);
          exhausted = first + nPairs >= nB;
          if (!active) {
            const uint32_t rank = (uint32_t)__popc(idleMask & ltMask);
            const uint32_t jn = first + (rank >> 1);
            if (rank < 2u * nPairs && jn < nB) {
              j = jn;
              bottomCtx(j, bc);
              const int P = (int)tstar[j];
              if (rank & 1u) { t = P; step = 1; tEnd = (int)nT; } else { t = P - 1; step = -1; tEnd = -1; }
              active = t != tEnd;
              if (!active && step < 0) { hval[j] = 0; tstar[j] = 0; }  // empty prefix
            }
          }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) {
          if (exhausted) break;
          continue;
        }
        if (active) {
          ++myTests;
          const float4 a = sA[t];
          const int cls = B200SEED_CLASSIFY(cfg, mid.r, mid.varZ, mid.varR, bc, a.x, a.y, a.z, a.w, sV[t]);
          if (cls == kPairEmit) emit(j, (uint32_t)t);
          if (cls <= kPairFailB) {  // the walk ends at the first failing top
            if (step < 0) {
              hval[j] = (uint16_t)(cls == kPairFailA ? t + 1 : t);
              tstar[j] = (uint16_t)t;
            }
            active = false;
          } else {
            t += step;
            if (t == tEnd) {
              if (step < 0) { hval[j] = 0; tstar[j] = 0; }  // no failing top in the prefix
              active = false;
            }
          }
        }
      }
    }
#else
    {
      // One lane per bottom; a lane that finishes takes the next unassigned bottom as soon as kRefill lanes
      // of its warp are idle, so the warp stays full although the windows differ in length.  Per bottom one
      // merged walk: down from |P_j| - 1 to the last failing top of the prefix (-> H_j, t*_j), then up from
      // |P_j| to the first failing top beyond it (brk_j).
      constexpr uint32_t kRefill = B200SEED_REFILL;
      // the sorted tops are read through two pinned 32-bit shared-memory addresses (one LEA per load)
      uint32_t aBase = 0, vBase = 0;
      if constexpr (!kSpill) {
        aBase = smem_u32(sA.ptr());
        vBase = smem_u32(sV.ptr());
        asm volatile("" : "+r"(aBase), "+r"(vBase));  // keep them in registers: no re-derivation from the carve-up in the loop
      }
      bool active = false, exhausted = false;  // exhausted is warp-uniform
      uint32_t j = 0, H = 0, ts = 0;
      int t = 0, tUp = 0, step = 0;
      BottomCtx bc{};
      for (;;) {
        const uint32_t actMask = __ballot_sync(0xffffffffu, active);
        if (!exhausted && (uint32_t)__popc(actMask) <= 32u - kRefill) {
          const uint32_t idleMask = ~actMask;
          const uint32_t nIdle = (uint32_t)__popc(idleMask);
          uint32_t first = 0;
          if (lane == 0) first = atomicAdd(&sh.nextBottom, nIdle);
          first = __shfl_sync(0xffffffffu, first, 0);
          exhausted = first + nIdle >= nB;
          if (!active) {
            const uint32_t jn = first + (uint32_t)__popc(idleMask & ltMask);
            if (jn < nB) {
              j = jn;
              bottomCtx(j, bc);
              const int P = (int)tstar[j];
              H = 0; ts = 0;
              tUp = P;
              if (P > 0) { t = P - 1; step = -1; } else { t = 0; step = 1; }
              active = true;  // nT > 0
#ifdef B200SEED_ABLATE_SCAN  // timing experiment only: no pair is tested
              active = false; hval[j] = 0; tstar[j] = 0;
#endif
            }
          }
          continue;  // re-vote with the new walkers
        }
        if (actMask == 0u) break;  // nothing walks and nothing is left to hand out
        if (active) {
          ++myTests;
          float4 a;
          float vT;
          if constexpr (!kSpill) {
            asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(aBase + 16u * (uint32_t)t));
            asm("ld.shared.f32 %0, [%1];" : "=f"(vT) : "r"(vBase + 4u * (uint32_t)t));
          } else {
            a = sA[t];
            vT = sV[t];
          }
          const int cls = B200SEED_CLASSIFY(cfg, mid.r, mid.varZ, mid.varR, bc, a.x, a.y, a.z, a.w, vT);
          if (cls == kPairEmit) emit(j, (uint32_t)t);
          const bool fail = cls <= kPairFailB;
          bool done = false;
#if B200SEED_FLAT_WALK
          {
            // straight-line bookkeeping: backward and forward walkers of a warp take the same instructions
            const bool bwd = step < 0;
            const bool bf = bwd & fail;
            H = bf ? (uint32_t)t + (cls == kPairFailA ? 1u : 0u) : H;
            ts = bf ? (uint32_t)t : ts;
            const bool turn = bwd & (fail | (t == 0));  // the prefix is done: turn around
            const int tn = turn ? tUp : t + step;
            step = turn ? 1 : step;
            done = ((!bwd) & fail) | (tn >= (int)nT);
            t = tn;
          }
#else
          if (step < 0) {
            if (fail) { H = (uint32_t)(cls == kPairFailA ? t + 1 : t); ts = (uint32_t)t; }
            if (fail || t == 0) {  // the prefix is done: turn around
              t = tUp;
              step = 1;
              done = t >= (int)nT;
            } else {
              --t;
            }
          } else {
            ++t;
            done = fail || t >= (int)nT;
          }
#endif
          if (done) {
            hval[j] = (uint16_t)H;
            tstar[j] = (uint16_t)ts;
            active = false;
          }
        }
      }
    }
#endif
    __syncthreads();

#if B200SEED_MERGE_3B3C
    // ---- phase 3b + 3c: window start = exclusive running max of H; the few bottoms with pairs in
    // [start_j, t*_j) that the scans did not touch are listed while the starts are written back
    {
      uint16_t* gapList = reinterpret_cast<uint16_t*>(pool2.ptr());  // pool2 is free until phase 3d
      if (tid == 0) sh.nSurv = 0;  // ordered before the appends below by the barriers of the block scan
      const uint32_t chunk = (nB + THREADS - 1) / THREADS;
      const uint32_t c0 = tid * chunk, c1 = (c0 + chunk < nB) ? c0 + chunk : nB;
      uint32_t localMax = 0;
      for (uint32_t j = c0; j < c1; ++j) localMax = localMax > hval[j] ? localMax : (uint32_t)hval[j];
      uint32_t blockMax;
      uint32_t run = block_scan_exclusive(localMax, sh.scratch, blockMax, OpMax());
      for (uint32_t j = c0; j < c1; ++j) {
        const uint32_t h = hval[j];
        hval[j] = (uint16_t)run;
        if (run < tstar[j]) gapList[atomicAdd(&sh.nSurv, 1u)] = (uint16_t)j;  // rare
        run = run > h ? run : h;
      }
      __syncthreads();
      const uint32_t nGap = sh.nSurv;
      for (uint32_t q = tid; q < nGap; q += THREADS) {
        const uint32_t j = gapList[q];
        const uint32_t s = hval[j], te = tstar[j];
        BottomCtx bc;
        bottomCtx(j, bc);
        for (uint32_t t = s; t < te; ++t) {
          ++myTests;
          const float4 a = sA[t];
          const int cls = B200SEED_CLASSIFY(cfg, mid.r, mid.varZ, mid.varR, bc, a.x, a.y, a.z, a.w, sV[t]);
          if (cls == kPairEmit) emit(j, t);
        }
      }
    }
#else
    // ---- phase 3b: window start = exclusive running max of H --------------
    {
      const uint32_t chunk = (nB + THREADS - 1) / THREADS;
      const uint32_t c0 = tid * chunk, c1 = (c0 + chunk < nB) ? c0 + chunk : nB;
      uint32_t localMax = 0;
      for (uint32_t j = c0; j < c1; ++j) localMax = localMax > hval[j] ? localMax : (uint32_t)hval[j];
      uint32_t blockMax;
      uint32_t run = block_scan_exclusive(localMax, sh.scratch, blockMax, OpMax());
      for (uint32_t j = c0; j < c1; ++j) {
        const uint32_t h = hval[j];
        hval[j] = (uint16_t)run;
        run = run > h ? run : h;
      }
    }
    __syncthreads();

    // ---- phase 3c: pairs in [start_j, t*_j) that the scans did not touch --
    // few bottoms have such a gap: list them first, then one thread per listed bottom
    {
      uint16_t* gapList = reinterpret_cast<uint16_t*>(pool2.ptr());  // pool2 is free until phase 3d
      if (tid == 0) sh.nSurv = 0;
      __syncthreads();
      for (uint32_t b0 = 0; b0 < nB; b0 += THREADS) {
        const uint32_t j = b0 + tid;
        const bool has = j < nB && hval[j] < tstar[j];
        const uint32_t slot = warp_append(&sh.nSurv, has);
        if (has) gapList[slot] = (uint16_t)j;
      }
      __syncthreads();
      const uint32_t nGap = sh.nSurv;
      for (uint32_t q = tid; q < nGap; q += THREADS) {
        const uint32_t j = gapList[q];
        const uint32_t s = hval[j], te = tstar[j];
        BottomCtx bc;
        bottomCtx(j, bc);
        for (uint32_t t = s; t < te; ++t) {
          ++myTests;
          const float4 a = sA[t];
          const int cls = B200SEED_CLASSIFY(cfg, mid.r, mid.varZ, mid.varR, bc, a.x, a.y, a.z, a.w, sV[t]);
          if (cls == kPairEmit) emit(j, t);
        }
      }
    }
#endif
    }  // !kStrip
    for (uint32_t j = tid; j <= nB; j += THREADS) cnt[j] = 0;
    __syncthreads();
    const uint32_t poolCount = sh.poolCount;
    if (poolCount > poolCap) {
      if (tid == 0) {
        p.slotCount[w] = 0;
        if (p.overflowList != nullptr) {
          p.overflowList[atomicAdd(p.overflowCount, 1u)] = w;
        } else {
          atomicOr(p.status, kStatusOverflowPool);
        }
      }
      continue;
    }

    // ---- phase 3d: group the candidates by bottom, curvature order inside --
    // a candidate from the scans is only real when its top is inside the window
    for (uint32_t e = tid; e < poolCount; e += THREADS) {
      const uint32_t to = pool[e];
      const uint32_t j = to >> 16, t = to & 0xFFFFu;
      if (t >= hval[j]) atomicAdd(&cnt[j], 1u);
    }
    __syncthreads();
    const uint32_t nValid = block_scan_array(cnt.ptr(), nB, sh.scratch);
    for (uint32_t e = tid; e < poolCount; e += THREADS) {
      const uint32_t to = pool[e];
      const uint32_t j = to >> 16, t = to & 0xFFFFu;
      if (t < hval[j]) continue;
      // the candidate's curvature and impact (TripletSeedFinder.cpp:148-155), one thread per candidate
      Cand c;
      if constexpr (kStrip) {
        StripBottomCtx sb;
        const uint32_t posB = stripBottom(j, sb);
        StripDerived calM, calB;
        loadStrip(m, calM);
        loadStrip(posB, calB);
        stripEval(sb, calM, calB, t, c.curv, c.impactOrWeight);
      } else {
        BottomCtx bc;
        bottomCtx(j, bc);
        const float4 a = sA[t];
        eval_pair(cfg, mid.r, mid.varZ, mid.varR, bc, a.x, a.y, a.z, a.w, sV[t], c.curv, c.impactOrWeight);
      }
      const float2 tzr = ldg2(p.pZR + topPos(t));
      c.topR = tzr.y;
      if (cfg.useDeltaRinsteadOfTopRadius) {
        const float dr = fsub(tzr.y, mid.r), dz = fsub(tzr.x, mid.z);
        c.topR = fsqrt(fadd(fmul(dr, dr), fmul(dz, dz)));
      }
      c.tOwner = to;
      pool2[atomicAdd(&cnt[j], 1u)] = c;
    }
    __syncthreads();  // cnt[j] is now the END of bottom j's group
    for (uint32_t j = tid; j < nB; j += THREADS) {
      const uint32_t s = j == 0 ? 0u : cnt[j - 1], e = cnt[j];
      if (e - s < 2) continue;
      Cand* grp = pool2.ptr() + s;
      const int n = (int)(e - s);
      // the reference's input order is ascending top rank (emission order) ...
      for (int i = 1; i < n; ++i) {
        const Cand v = grp[i];
        int q = i;
        while (q > 0 && (grp[q - 1].tOwner & 0xFFFFu) > (v.tOwner & 0xFFFFu)) { grp[q] = grp[q - 1]; --q; }
        grp[q] = v;
      }
      // ... sorted by curvature with std::ranges::sort (BroadTripletSeedFilter.cpp:143-148)
      std_sort(grp, n, cand_less);
    }
    __syncthreads();

    if constexpr (kConf) {
      // ---- seedConfirmation: weights and the map-independent cuts of every candidate
      // (BroadTripletSeedFilter.cpp:253-276), then the survivors go to global memory in
      // the reference's push order; k_conf_replay applies the collector and
      // bestSeedQualityMap logic on them.
      const float rMaxSeedConfMid = conf_range(cfg, mid.z).rMaxSeedConf;  // state().rMaxSeedConf, :84
      uint32_t myKept = 0;
      for (uint32_t i = tid; i < nValid; i += THREADS) {
        const Cand c = pool2[i];
        const uint32_t j = c.tOwner >> 16;
        const uint32_t s = j == 0 ? 0u : cnt[j - 1], e = cnt[j];
        const Cand* grp = pool2.ptr() + s;
        uint32_t nCompat;
        float wgt = filter_weight(
            cfg, (int)(e - s), (int)(i - s), c.impactOrWeight, [&](int q) { return grp[q].curv; },
            [&](int q) { return grp[q].topR; }, nCompat);
        const float4 br = __ldg(reinterpret_cast<const float4*>(recB + rankB[j]));  // {pos, cotTheta, ...}
        const float2 bzr = ldg2(p.pZR + __float_as_uint(br.x));
        const float zOrigin = fsub(mid.z, fmul(mid.r, br.y));  // :119
        int deltaSeedConf;
        const bool keepIt = conf_candidate(cfg, conf_range(cfg, bzr.x), bzr.y, zOrigin, c.impactOrWeight, nCompat, wgt, deltaSeedConf);
        uint32_t meta = j | ((e - s < 3u ? e - s : 3u) << kRecGroupSizeShift);
        if (!(bzr.y > rMaxSeedConfMid)) meta |= kRecNeedsTwoTops;  // :107-110
        if (deltaSeedConf > 0) meta |= kRecQuality;
        if (keepIt) { meta |= kRecKeep; ++myKept; }
        pool2[i].impactOrWeight = wgt;  // nobody reads another candidate's impact
        pool[i] = meta;                 // the emission records are dead after phase 3d
      }
      uint32_t nKept;
      block_scan_exclusive(myKept, sh.scratch, nKept, OpSum());
      if (tid == 0) {
        const uint32_t first = nKept != 0 ? atomicAdd(p.recCounter, nKept) : 0u;
        sh.runCarry = first;
        const bool fits = (unsigned long long)first + nKept <= (unsigned long long)p.recCapacity;
        if (!fits) atomicOr(p.status, kStatusOverflowRecords);
        p.recBegin[w] = first;
        p.recCount[w] = fits ? nKept : 0u;
        p.slotCount[w] = 0;
        sh.cnt[kCntCandidates] += nValid;
        sh.cnt[kCntTieMiddles] += sh.tie;
      }
      __syncthreads();
      const uint32_t recBase = sh.runCarry;
      if ((unsigned long long)recBase + nKept <= (unsigned long long)p.recCapacity) {
        uint32_t carry = 0;
        for (uint32_t b0 = 0; b0 < nValid; b0 += THREADS) {
          const uint32_t i = b0 + tid;
          const uint32_t meta = i < nValid ? pool[i] : 0u;
          const bool keepIt = (meta & kRecKeep) != 0u;
          uint32_t total;
          const uint32_t rank = carry + block_scan_exclusive(keepIt ? 1u : 0u, sh.scratch, total, OpSum());
          carry += total;
          if (keepIt) {
            const Cand c = pool2[i];
            const uint32_t j = meta & kRecGroupMask;
            const size_t o = (size_t)recBase + rank;
            const float4 br = __ldg(reinterpret_cast<const float4*>(recB + rankB[j]));
            p.rec4[o] = make_uint4(__float_as_uint(br.x), topPos(c.tOwner & 0xFFFFu), __float_as_uint(c.impactOrWeight), meta & ~kRecKeep);
            p.recZ[o] = fsub(mid.z, fmul(mid.r, br.y));
          }
        }
      }
      uint32_t t = myTests;
      for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
      if (lane == 0) atomicAdd(&sh.cnt[kCntTripletTests], (unsigned long long)t);
    } else {
    // ---- phase 3e: one thread per candidate: weight -----------------------
    for (uint32_t i = tid; i < nValid; i += THREADS) {
      const Cand c = pool2[i];
      const uint32_t j = c.tOwner >> 16;
      const uint32_t s = j == 0 ? 0u : cnt[j - 1], e = cnt[j];
      const Cand* grp = pool2.ptr() + s;
      const float wgt = filter_weight(
          cfg, (int)(e - s), (int)(i - s), c.impactOrWeight, [&](int q) { return grp[q].curv; },
          [&](int q) { return grp[q].topR; });
      pool2[i].impactOrWeight = wgt;  // nobody reads another candidate's impact
    }
    __syncthreads();

    // ---- phase 3f + 4: bounded heap (CandidatesForMiddleSp.cpp:44-93) and the per-middle
    // selection (BroadTripletSeedFilter.cpp:324-393), all by warp 0 -----------------------
    // The heap keeps the nLow largest weights; its final sort_heap order is
    // unique unless weights tie.  Warp 0 therefore selects the nLow + 1 largest
    // weights in registers (lane r holds the r-th largest, stable in arrival
    // order).  Only if two of them are equal -- the cases where the reference's
    // result depends on the heap's history -- the literal heap replay runs.
    if (tid < 32) {
      const int nLow = (int)cfg.maxSeedsPerSpMConf;
      // Only the first min(nLow, maxSeedsPerSpM + 1) = K entries of the sorted collector are used (:336-348).  They
      // are the K largest weights in descending order whenever the K + 1 largest differ: an entry with fewer than
      // K <= nLow others at or above its weight is never refused and never the evicted minimum.
      const int nUse = cfg.maxSeedsPerSpM < (uint32_t)nLow ? (int)cfg.maxSeedsPerSpM + 1 : nLow;
      const int keep = nUse + 1;  // <= kMaxHeap + 1 <= 32 lanes (host_plan.cpp)
      WeightIndex* heapArr = sh.heap;
      StoredSeed* storArr = sh.storage;
      if (p.bigHeap != 0u) {
        unsigned char* tail = smemRaw + (kSpill ? 0u : p.arrayBytes);
        heapArr = reinterpret_cast<WeightIndex*>(tail);
        storArr = reinterpret_cast<StoredSeed*>(tail + 8u * (uint32_t)nLow);
      }
      float myW = -3.402823466e+38f;
      uint32_t myId = 0xFFFFFFFFu;
      int filled = 0;
      for (uint32_t c0 = 0; c0 < nValid && nLow > 0; c0 += 32) {
        const uint32_t i = c0 + lane;
        float wgt = 0.f;
        uint32_t id = 0;
        if (i < nValid) { wgt = pool2[i].impactOrWeight; id = pool2[i].tOwner; }
        const float kth = __shfl_sync(0xffffffffu, myW, keep - 1);
        uint32_t mask = __ballot_sync(0xffffffffu, i < nValid && (filled < keep || wgt > kth));
        while (mask != 0u) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1u;
          const float v = __shfl_sync(0xffffffffu, wgt, src);
          const uint32_t vid = __shfl_sync(0xffffffffu, id, src);
          // stable insertion: after every entry with weight >= v
          const uint32_t ge = __ballot_sync(0xffffffffu, (int)lane < filled && myW >= v);
          const int pos = __popc(ge);
          if (pos >= keep) continue;
          const float upW = __shfl_up_sync(0xffffffffu, myW, 1);
          const uint32_t upId = __shfl_up_sync(0xffffffffu, myId, 1);
          if ((int)lane > pos && (int)lane < keep) { myW = upW; myId = upId; }
          if ((int)lane == pos) { myW = v; myId = vid; }
          if (filled < keep) ++filled;
        }
      }
      // ties among the K + 1 largest -> the order / membership depends on the heap history
      const float nextW = __shfl_down_sync(0xffffffffu, myW, 1);
      const bool tieHere = (int)lane + 1 < filled && myW == nextW;
      const bool anyTie = __any_sync(0xffffffffu, tieHere);
      int inHeap = filled < nLow ? filled : nLow;
      if (anyTie) {
        // literal replay in the reference's push order (bottom-major, curvature order)
        if (lane == 0) { sh.heapSize = 0; }
        __syncwarp();
        for (uint32_t c0 = 0; c0 < nValid && nLow > 0; c0 += 32) {
          const uint32_t i = c0 + lane;
          const int hs = sh.heapSize;
          const float hmin = sh.heapMin;
          bool want = false;
          if (i < nValid) want = (hs < nLow) || (pool2[i].impactOrWeight > hmin);
          uint32_t mask = __ballot_sync(0xffffffffu, want);
          if (lane == 0) {
            while (mask != 0u) {
              const uint32_t q = c0 + (uint32_t)(__ffs(mask) - 1);
              mask &= mask - 1u;
              const Cand c = pool2[q];
              const float wq = c.impactOrWeight;
              StoredSeed sd;
              sd.tOwner = c.tOwner;
              sd.weight = wq;
              if (sh.heapSize < nLow) {
                const int slotI = sh.heapSize;
                storArr[slotI] = sd;
                heapArr[slotI].weight = wq;
                heapArr[slotI].index = (uint32_t)slotI;
                sh.heapSize = slotI + 1;
                std_push_heap(heapArr, sh.heapSize, heap_comp);
              } else {
                const WeightIndex smallest = heapArr[0];
                if (wq <= smallest.weight) continue;
                storArr[smallest.index] = sd;
                std_pop_heap(heapArr, sh.heapSize, heap_comp);
                heapArr[sh.heapSize - 1].weight = wq;
                heapArr[sh.heapSize - 1].index = smallest.index;
                std_push_heap(heapArr, sh.heapSize, heap_comp);
              }
              sh.heapMin = heapArr[0].weight;
            }
          }
          __syncwarp();
        }
        if (lane == 0) std_sort_heap(heapArr, sh.heapSize, heap_comp);
        __syncwarp();
        inHeap = sh.heapSize;
        if ((int)lane < inHeap) {
          const StoredSeed sd = storArr[heapArr[lane].index];
          myW = sd.weight;
          myId = sd.tOwner;
        }
      }
      // phase 4: lane i writes seed i
      uint32_t maxSeeds = (uint32_t)inHeap;
      if (maxSeeds > cfg.maxSeedsPerSpM) maxSeeds = cfg.maxSeedsPerSpM + 1;
      if (lane < maxSeeds) {
        const uint32_t j = myId >> 16, tRank = myId & 0xFFFFu;
        const float4 br = __ldg(reinterpret_cast<const float4*>(recB + rankB[j]));
        const size_t o = (size_t)w * p.seedsPerMiddle + lane;
        p.slotB[o] = __float_as_uint(br.x);
        p.slotM[o] = m;
        p.slotT[o] = topPos(tRank);
        p.slotQ[o] = myW;
        p.slotZ[o] = fsub(mid.z, fmul(mid.r, br.y));  // zOrigin, BroadTripletSeedFilter.cpp:119
      }
      uint32_t t = myTests;
      for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
      if (lane == 0) {
        p.slotCount[w] = maxSeeds;
        atomicAdd(&sh.cnt[kCntTripletTests], (unsigned long long)t);
        sh.cnt[kCntCandidates] += nValid;
        sh.cnt[kCntSeeds] += maxSeeds;
        sh.cnt[kCntTieMiddles] += sh.tie;
      }
    } else {
      uint32_t t = myTests;
      for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
      if (lane == 0) atomicAdd(&sh.cnt[kCntTripletTests], (unsigned long long)t);
    }
    }  // !kConf
  }
  __syncthreads();
  if (tid < (uint32_t)kCntSlots && sh.cnt[tid] != 0ull) atomicAdd(p.counters + tid, sh.cnt[tid]);
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
    const uint32_t w = e < p.nEvents ? p.workStart[p.eventFirstItem != nullptr ? (size_t)p.eventFirstItem[e] : (size_t)e * p.nNav] : nWork;
    // seedStart[nWork] is only written when nWork > 0
    p.seedOffsets[e] = nWork == 0 ? 0ull : (unsigned long long)p.seedStart[w];
  }
}

// ---------------------------------------------------------------------------
// seedConfirmation = true.  The reference threads one mutable map through the
// whole event: bestSeedQualityMap[sp] = best quality of the seeds emitted so far
// that contain sp (BroadTripletSeedFilter.cpp:33-50,278-285,364-376), read by
// every later middle.  The device resolves that order dependence by fixed-point
// iteration over rounds: round r replays the collector logic of every middle
// (one warp each) against Q_r(sp, w) = max quality over the seeds that round
// r - 1 emitted for middles BEFORE w (round 0: empty map).  The seeds of the
// first middle in order are final after round 0, and by induction everything is
// final once a round reproduces its predecessor; the fixed point is exactly the
// sequential result.  Seeds live in two slot sets (previous / current); the
// per-space-point lists of the previous round's seeds are linked lists built by
// k_conf_link (node = 3 * slot + role).
// ---------------------------------------------------------------------------
struct ConfParams {
  DeviceConfig cfg;
  const uint32_t* nWorkPtr;
  const uint32_t* workPos;
  const uint4* rec;
  const float* recZ;
  const uint32_t *recBegin, *recCount;
  const uint32_t *prevB, *prevT;  // the previous round's seeds
  const float* prevQ;
  const uint32_t* prevCount;
  uint32_t *curB, *curM, *curT;   // this round's seeds
  float *curQ, *curZ;
  uint32_t* curCount;
  int* head;        // [nTotal] first node of a space point, -1: none
  int* next;        // [3 * nWork * seedsPerMiddle]
  uint32_t seedsPerMiddle;
  int round;
  uint32_t* changed;  // [rounds] number of middles whose seeds differ from the previous round
  // Space points contained in a seed (old or new) of a middle whose output changed in the previous round: a middle
  // that reads none of them sees the same map as one round ago and keeps its seeds without a replay.
  const uint8_t* dirty;  // [nTotal] written by the previous round
  uint8_t* dirtyNext;    // [nTotal] for the next round (cleared before this round)
  const float* prevZ;
};

__device__ __forceinline__ bool conf_converged(const ConfParams& p) {
  return p.round >= 2 && p.changed[p.round - 1] == 0u;
}

__global__ void __launch_bounds__(256) k_conf_link(const __grid_constant__ ConfParams p) {
  if (conf_converged(p)) return;
  const uint32_t nSlots = *p.nWorkPtr * p.seedsPerMiddle;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < nSlots; slot += gridDim.x * blockDim.x) {
    const uint32_t w = slot / p.seedsPerMiddle, i = slot - w * p.seedsPerMiddle;
    if (i >= p.prevCount[w]) continue;
    const uint32_t sp[3] = {p.prevB[slot], p.workPos[w], p.prevT[slot]};
#pragma unroll
    for (int role = 0; role < 3; ++role) {
      const int node = (int)(slot * 3u + (uint32_t)role);
      p.next[node] = atomicExch(p.head + sp[role], node);
    }
  }
}

// getBestSeedQuality (.cpp:24-31) as of just before work item w
__device__ __forceinline__ float conf_best(const ConfParams& p, uint32_t sp, uint32_t w) {
  float best = -3.402823466e+38f;  // numeric_limits<float>::lowest()
  for (int node = p.head[sp]; node >= 0; node = p.next[node]) {
    const uint32_t slot = (uint32_t)node / 3u;
    if (slot / p.seedsPerMiddle < w) {
      const float q = p.prevQ[slot];
      best = q > best ? q : best;  // std::max(quality, it->second), .cpp:41
    }
  }
  return best;
}

struct ConfSeed {
  uint32_t b, t;
  float weight, zOrigin;
  uint32_t isQuality;
};

// CandidatesForMiddleSp::push (detail/CandidatesForMiddleSp.cpp:44-78), one of the two heaps
__device__ __forceinline__ void conf_push(WeightIndex* heap, int& size, int nMax, ConfSeed* storage, int& nStored,
                                          const ConfSeed& sd) {
  if (nMax == 0) return;
  if (size < nMax) {
    storage[nStored] = sd;
    heap[size].weight = sd.weight;
    heap[size].index = (uint32_t)nStored;
    ++nStored;
    ++size;
    std_push_heap(heap, size, heap_comp);
    return;
  }
  const WeightIndex smallest = heap[0];
  if (sd.weight <= smallest.weight) return;
  storage[smallest.index] = sd;
  std_pop_heap(heap, size, heap_comp);
  heap[size - 1].weight = sd.weight;
  heap[size - 1].index = smallest.index;
  std_push_heap(heap, size, heap_comp);
}

constexpr int kConfWarps = 4;

// kCap: capacity of either collector heap (kMaxHeap for the usual 5 / 5, kMaxHeapBig for itk.py's strip block 100 / 100)
template <int kCap>
__global__ void __launch_bounds__(kConfWarps * 32) k_conf_replay(const __grid_constant__ ConfParams p) {
  if (conf_converged(p)) return;
  __shared__ WeightIndex heapHigh[kConfWarps][kCap], heapLow[kConfWarps][kCap];
  __shared__ ConfSeed storageAll[kConfWarps][2 * kCap];
  const DeviceConfig& cfg = p.cfg;
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WeightIndex* hHigh = heapHigh[wib];
  WeightIndex* hLow = heapLow[wib];
  ConfSeed* storage = storageAll[wib];
  const int maxHigh = (int)cfg.maxQualitySeedsPerSpMConf, maxLow = (int)cfg.maxSeedsPerSpMConf;
  const uint32_t nWork = *p.nWorkPtr;
  const uint32_t warpsPerGrid = gridDim.x * kConfWarps;
  const float kLowest = -3.402823466e+38f;
  uint32_t myChanged = 0;
  for (uint32_t w = blockIdx.x * kConfWarps + wib; w < nWork; w += warpsPerGrid) {
    const uint32_t n = p.recCount[w];
    const uint32_t K = p.seedsPerMiddle;
    const size_t slot0 = (size_t)w * K;
    uint32_t nOut = 0;
    if (n != 0 && p.round >= 2) {
      const uint32_t recBase = p.recBegin[w];
      bool touched = p.dirty[p.workPos[w]] != 0;
      for (uint32_t i = lane; i < n; i += 32) {
        const uint4 r = p.rec[recBase + i];
        touched |= (p.dirty[r.x] | p.dirty[r.y]) != 0;
      }
      if (!__any_sync(0xffffffffu, touched)) {  // same inputs as in the previous round: same seeds
        const uint32_t c = p.prevCount[w];
        if (lane < c) {
          p.curB[slot0 + lane] = p.prevB[slot0 + lane];
          p.curM[slot0 + lane] = p.workPos[w];
          p.curT[slot0 + lane] = p.prevT[slot0 + lane];
          p.curQ[slot0 + lane] = p.prevQ[slot0 + lane];
          p.curZ[slot0 + lane] = p.prevZ[slot0 + lane];
        }
        if (lane == 0) p.curCount[w] = c;
        continue;
      }
    }
    if (n != 0) {
      const uint32_t recBase = p.recBegin[w];
      const uint32_t m = p.workPos[w];
      const float bestM = conf_best(p, m, w);
      int nHigh = 0, nLow = 0, nStored = 0;  // maintained by lane 0, broadcast where the warp needs them
      // state of the (middle, bottom) group being processed
      uint32_t curGroup = 0xFFFFFFFFu;
      bool groupSkip = false;
      float bestB = kLowest;
      bool lowHas = false;  // maxWeightSeed / weightMax / maxWeightTopSp, .cpp:138-140
      float lowW = kLowest, lowZ = 0.f;
      uint32_t lowT = 0, lowB = 0;
      auto closeGroup = [&]() {  // .cpp:305-321: the best lower-quality seed, only without quality seeds
        if (lane == 0 && lowHas && nHigh == 0) {
          ConfSeed sd{lowB, lowT, lowW, lowZ, 0u};
          conf_push(hLow, nLow, maxLow, storage, nStored, sd);
        }
        lowHas = false;
        lowW = kLowest;
      };
      for (uint32_t c0 = 0; c0 < n; c0 += 32) {
        const uint32_t i = c0 + lane;
        const bool valid = i < n;
        uint4 r = make_uint4(0, 0, 0, 0);
        float z = 0.f;
        if (valid) { r = p.rec[recBase + i]; z = p.recZ[recBase + i]; }
        const uint32_t grp = valid ? (r.w & kRecGroupMask) : 0xFFFFFFFFu;
        const float wgt = __uint_as_float(r.z);
        const float bestT = valid ? conf_best(p, r.y, w) : kLowest;
        uint32_t todo = __ballot_sync(0xffffffffu, valid);
        while (todo != 0u) {
          const int first = __ffs(todo) - 1;
          const uint32_t g = __shfl_sync(0xffffffffu, grp, first);
          const uint32_t piece = __ballot_sync(0xffffffffu, valid && grp == g) & todo;  // groups are contiguous
          todo &= ~piece;
          if (g != curGroup) {
            closeGroup();
            curGroup = g;
            const uint32_t meta = __shfl_sync(0xffffffffu, r.w, first);
            const uint32_t bPos = __shfl_sync(0xffffffffu, r.x, first);
            const int nHighNow = __shfl_sync(0xffffffffu, nHigh, 0);
            // minCompatibleTopSPs, .cpp:106-118
            const uint32_t minTops = ((meta & kRecNeedsTwoTops) ? 2u : 1u) + (nHighNow > 0 ? 1u : 0u);
            groupSkip = ((meta >> kRecGroupSizeShift) & 3u) < minTops;
            bestB = groupSkip ? kLowest : conf_best(p, bPos, w);
          }
          if (groupSkip) continue;
          const bool mine = ((piece >> lane) & 1u) != 0u;
          // .cpp:278-285
          const bool pass = mine && !(wgt < bestB && wgt < bestM && wgt < bestT);
          const bool quality = (r.w & kRecQuality) != 0u;
          uint32_t hi = __ballot_sync(0xffffffffu, pass && quality);
          const uint32_t lo = __ballot_sync(0xffffffffu, pass && !quality);
          if (lo != 0u) {
            // first maximum of the piece; an earlier piece of the group wins ties (weight > weightMax, .cpp:296)
            float best = (pass && !quality) ? wgt : kLowest;
            uint32_t bestLane = (pass && !quality) ? lane : 32u;
            for (int d = 16; d > 0; d >>= 1) {
              const float ow = __shfl_xor_sync(0xffffffffu, best, d);
              const uint32_t ol = __shfl_xor_sync(0xffffffffu, bestLane, d);
              if (ol != 32u && (bestLane == 32u || ow > best || (ow == best && ol < bestLane))) { best = ow; bestLane = ol; }
            }
            const uint32_t bt = __shfl_sync(0xffffffffu, r.y, bestLane & 31u);
            const uint32_t bb = __shfl_sync(0xffffffffu, r.x, bestLane & 31u);
            const float bz = __shfl_sync(0xffffffffu, z, bestLane & 31u);
            if (!lowHas || best > lowW) {
              if (best > lowW) { lowHas = true; lowW = best; lowT = bt; lowB = bb; lowZ = bz; }
            }
          }
          while (hi != 0u) {  // quality seeds go to the collector in order, .cpp:287-295
            const int src = __ffs(hi) - 1;
            hi &= hi - 1u;
            ConfSeed sd;
            sd.b = __shfl_sync(0xffffffffu, r.x, src);
            sd.t = __shfl_sync(0xffffffffu, r.y, src);
            sd.weight = __shfl_sync(0xffffffffu, wgt, src);
            sd.zOrigin = __shfl_sync(0xffffffffu, z, src);
            sd.isQuality = 1u;
            if (lane == 0) conf_push(hHigh, nHigh, maxHigh, storage, nStored, sd);
          }
        }
      }
      closeGroup();
      __syncwarp();
      // ---- filterTripletsMiddleFixed, .cpp:324-393 (lane 0)
      if (lane == 0) {
        std_sort_heap(hHigh, nHigh, heap_comp);
        std_sort_heap(hLow, nLow, heap_comp);
        const uint32_t total = (uint32_t)(nHigh + nLow);
        uint32_t maxSeeds = total;
        if (maxSeeds > cfg.maxSeedsPerSpM) maxSeeds = cfg.maxSeedsPerSpM + 1;
        for (uint32_t k = 0; k < total && nOut < maxSeeds; ++k) {
          const ConfSeed sd = storage[k < (uint32_t)nHigh ? hHigh[k].index : hLow[k - (uint32_t)nHigh].index];
          if (nHigh > 0 && sd.isQuality == 0u) continue;
          // the map as of now: the previous middles' seeds plus this middle's own earlier ones (.cpp:375)
          float qB = conf_best(p, sd.b, w), qM = bestM, qT = conf_best(p, sd.t, w);
          for (uint32_t e = 0; e < nOut; ++e) {
            const float qe = p.curQ[slot0 + e];
            const uint32_t eb = p.curB[slot0 + e], et = p.curT[slot0 + e];
            qM = qe > qM ? qe : qM;
            if (eb == sd.b || et == sd.b) qB = qe > qB ? qe : qB;
            if (eb == sd.t || et == sd.t) qT = qe > qT ? qe : qT;
          }
          if (sd.weight < qB && sd.weight < qM && sd.weight < qT) continue;
          p.curB[slot0 + nOut] = sd.b;
          p.curM[slot0 + nOut] = m;
          p.curT[slot0 + nOut] = sd.t;
          p.curQ[slot0 + nOut] = sd.weight;
          p.curZ[slot0 + nOut] = sd.zOrigin;
          ++nOut;
        }
      }
    }
    if (lane == 0) {
      p.curCount[w] = nOut;
      bool same = p.round > 0 && p.prevCount[w] == nOut;
      for (uint32_t e = 0; same && e < nOut; ++e) {
        same = p.prevB[slot0 + e] == p.curB[slot0 + e] && p.prevT[slot0 + e] == p.curT[slot0 + e] &&
               __float_as_uint(p.prevQ[slot0 + e]) == __float_as_uint(p.curQ[slot0 + e]);
      }
      if (!same) {
        ++myChanged;
        p.dirtyNext[p.workPos[w]] = 1;
        if (p.round > 0) {
          for (uint32_t e = 0; e < p.prevCount[w]; ++e) { p.dirtyNext[p.prevB[slot0 + e]] = 1; p.dirtyNext[p.prevT[slot0 + e]] = 1; }
        }
        for (uint32_t e = 0; e < nOut; ++e) { p.dirtyNext[p.curB[slot0 + e]] = 1; p.dirtyNext[p.curT[slot0 + e]] = 1; }
      }
    }
  }
  if (lane == 0 && myChanged != 0u) atomicAdd(p.changed + p.round, myChanged);
}

// Free track parameters of every seed (Acts::estimateTrackParamsFromSeed,
// Core/src/Seeding/EstimateTrackParamsFromSeed.cpp:114-160), FP64, one thread per seed.
__global__ void __launch_bounds__(256) k_estimate_params(const uint32_t* __restrict__ b, const uint32_t* __restrict__ m,
                                                         const uint32_t* __restrict__ t, const float* __restrict__ x,
                                                         const float* __restrict__ y, const float* __restrict__ z,
                                                         double bx, double by, double bz, double* __restrict__ out,
                                                         unsigned long long n, uint32_t nSpacePoints, int* status) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t ib = b[i], im = m[i], it = t[i];
    if (ib >= nSpacePoints || im >= nSpacePoints || it >= nSpacePoints) {  // caller-supplied indices: never read outside the columns
      atomicOr(status, 1);
#pragma unroll
      for (int k = 0; k < 8; ++k) out[i * 8 + k] = 0.0;
      continue;
    }
    const double s0[3] = {(double)x[ib], (double)y[ib], (double)z[ib]};
    const double s1[3] = {(double)x[im], (double)y[im], (double)z[im]};
    const double s2[3] = {(double)x[it], (double)y[it], (double)z[it]};
    const double bf[3] = {bx, by, bz};
    double o[8];
    estimate_free_params(s0, 0.0, s1, s2, bf, o);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[i * 8 + k] = o[k];
  }
}

// Pixel space points from measurements (SpacePointMaker.cpp:44-76), FP64, one thread per measurement
struct SpacePointMakerParams {
  const uint32_t* surface;  // [n] index into transforms
  const double *loc0, *loc1, *cov00, *cov01, *cov11;
  const double* transforms;  // [nSurfaces][12] row-major 3x4
  float *x, *y, *z, *r, *varZ, *varR;
  uint32_t n, nSurfaces;
  int* status;
};
__global__ void __launch_bounds__(256) k_pixel_spacepoints(const __grid_constant__ SpacePointMakerParams p) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    const uint32_t sf = p.surface[i];
    if (sf >= p.nSurfaces) {
      atomicOr(p.status, 1);
      continue;
    }
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = __ldg(p.transforms + (size_t)sf * 12 + k);
    float o[6];
    pixel_space_point(T, p.loc0[i], p.loc1[i], p.cov00[i], p.cov01[i], p.cov11[i], o);
    p.x[i] = o[0]; p.y[i] = o[1]; p.z[i] = o[2]; p.r[i] = o[3]; p.varZ[i] = o[4]; p.varR[i] = o[5];
  }
}

// phi = atan2f(y, x) replay, for validation against the host libm
__global__ void k_atan2f(const float* __restrict__ y, const float* __restrict__ x, float* __restrict__ out, unsigned long long n) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    out[i] = glibc_atan2f(y[i], x[i]);
  }
}

}  // namespace B200SEED_NS
