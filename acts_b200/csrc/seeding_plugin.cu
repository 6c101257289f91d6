// seeding_plugin.cu -- C ABI of the B200 seeding plugin (include/acts_b200_seeding.h).
//
// Host side: validates the configuration (host_plan.cpp), owns the device
// workspaces and enqueues the kernel sequence of seeding_kernels.cuh on one
// stream.  There is no CPU implementation of any stage in this library: without
// a CUDA device every computing entry point fails with B200SEED_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/acts_b200_seeding.h"
#include "host_plan.hpp"
#include <atomic>
#include <thread>

#include "seeding_kernels.cuh"
#include "orthogonal_kernels.cuh"

using namespace B200SEED_NS;

namespace {

thread_local std::string g_lastError;

int fail(int code, const std::string& msg) {
  g_lastError = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t err__ = (expr);                                                          \
    if (err__ != cudaSuccess) {                                                          \
      return fail(B200SEED_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
    }                                                                                    \
  } while (0)

struct DevBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (ptr != nullptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    const size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() {
    if (ptr != nullptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(ptr); }
};

// confState words: [0] record counter, [kConfChangedBase + r] middles changed in round r
constexpr int kConfChangedBase = 8;
constexpr int kConfMaxRounds = 1016;
constexpr int kConfStateWords = kConfChangedBase + kConfMaxRounds;
constexpr int kConfRoundsPerBatch = 12;  // default number of rounds enqueued before the host looks at the convergence flags

uint32_t env_u32(const char* name, uint32_t def) {
  const char* v = std::getenv(name);
  if (v == nullptr || *v == 0) return def;
  return static_cast<uint32_t>(std::strtoul(v, nullptr, 10));
}

}  // namespace

struct b200seed_handle {
  HostPlan plan;
  int device = 0;
  cudaStream_t stream = nullptr;
  int smCount = 0, ccMajor = 0, ccMinor = 0;
  // engine tunables
  // per-middle shared-memory capacities are compile-time (seeding_kernels.cuh):
  // tier 0 takes every middle at 2 blocks per SM, middles that do not fit are
  // re-queued to tier 1 (1 block per SM)
  uint32_t sortSmemCap = 2560;  // elements of a bin sorted in shared memory (16 bytes each + the bucket counters: 4 blocks per SM); larger bins use the global scratch
  int exactTies = 1;
  uint32_t phiFirst = 1, phiCount = 0xFFFFFFFFu;  // middle phi-bin sector (default: all)
  // shared-memory classes of the seeding kernel (seeding_kernels.cuh): blocks per SM, dynamic bytes per block
  int classBlocksPerSM[kNumSeedClasses] = {};
  int classThreads[kNumSeedClasses] = {};
  uint32_t classBytes[kNumSeedClasses] = {};
  int doubletBlocksPerSM[2] = {1, 1};  // count / fill
  int kdBlocksPerSM[2] = {1, 1};       // orthogonal seeder: k_doublets_kd count / fill
  size_t arenaMaxBytes = (size_t)8192 << 20;  // B200SEED_ARENA_MB: doublet arena per chunk of middles
  // constant tables
  DevBuf navBins, botOffsets, botBins, topOffsets, topBins;
  // per-batch workspaces
  DevBuf inOffsets, inX, inY, inZ, inR, inVarZ, inVarR;  // staging of host inputs
  DevBuf binOf, binCount, binStart, binCursor, tmpIdx, pIdx, pXY, pZR, pVar, sortScratch;
  DevBuf midLo, midCount, workStart, workPos, workEG, workCounter;
  // doublet stage: slot sizes and prefix, chunk plan, arena, per-middle headers, per-class work lists
  DevBuf capB, capT, slotPrefix, capTileSums, capTilePrefix, planDev, hdr, carve, classList, arenaRec[2], arenaKey[2], spillScratch;
  uint32_t heapBytes = 0;  // maxSeedsPerSpMConf > kMaxHeap: dynamic shared memory behind the per-middle arrays for the literal heap replay
  DevBuf inStrip, pStrip;  // strip triplet path: raw details of the event (12 floats per point), derived details per packed position
  const float* pendingStrip = nullptr;  // set by b200seed_run_strips around its run_host_batch call
  float pendingCotThetaDiffMax = 0.f;
  DevBuf maskArena, maskOff;  // count pass -> fill pass: windows + survivor masks per middle (B200SEED_MASK_WORDS_PER_SP, 0 = off)
  uint32_t maskWordsPerSp = 192;
  // Consecutive chunks alternate between two internal streams (and two arena halves): the tail of one chunk's
  // seeding kernels overlaps the fill pass of the next.  B200SEED_CHUNK_STREAMS=1 serialises them (stage timing).
  int chunkStreams = 2;
  cudaStream_t chunkStream[2] = {nullptr, nullptr};
  cudaEvent_t evPlan = nullptr, evChunkEnd[2] = {nullptr, nullptr};
  // The seeding kernels of the shared-memory classes of one chunk are independent: each runs on its own stream, so
  // the tail of one class overlaps the start of the next (B200SEED_CLASS_STREAMS=0: one after the other).
  int classStreams = 1;
  cudaStream_t classStream[2][kNumSeedClasses] = {};
  cudaEvent_t evFill[2] = {nullptr, nullptr}, evClass[2][kNumSeedClasses] = {};
  uint32_t* hPlan = nullptr;  // pinned: planWords[8] + chunkBounds[kMaxChunks + 1]
  std::vector<uint32_t> lastChunkBounds;  // of the last call (debug_doublets)
  DevBuf slotB, slotM, slotT, slotQ, slotZ, slotCount, seedStart, tileSums, tilePrefix;
  DevBuf outB, outM, outT, outQ, outZ, seedOffsets;  // device outputs of the host API
  DevBuf counters, status, zWin, zWinOffsets;
  // orthogonal seeder: the event trees built by the host layer (kd_tree_host.hpp)
  DevBuf orthPosOrig, orthPosPhi, orthNodes, orthCoreOffsets, orthNodeOffsets, orthRRange;
  DevBuf orthElemR, orthElemZ, orthScratch, orthTasks, orthExtent;  // device construction of the trees
  DevBuf kdHitArena, kdHitHead;  // count pass -> fill pass: survivor positions of every tree walk (B200SEED_KD_HIT_MB, 0 = off)
  uint32_t kdHitMB = 4096;
  bool kdHostBuild = false;  // B200SEED_KD_HOST=1: build the trees in the host layer (kd_tree_host.hpp) instead
  uint32_t itemsMax = 0;  // upper bound of the work items of the last call (grid: space points, orthogonal: 2 x)
  uint32_t zWinCapacity = 1;  // windows per column of zWin (lo column, then hi column)
  DoubletParams lastDoublets{};  // of the last call (debug_doublets re-runs the fill pass chunk by chunk)
  // seedConfirmation: candidate records, second slot set, per-space-point seed lists, {record counter, changed[round]}
  DevBuf rec, recZ, recBegin, recCount, slot2B, slot2M, slot2T, slot2Q, slot2Z, slot2Count, confHead, confNext, confState, confDirty;
  uint32_t* hConfState = nullptr;  // pinned mirror of confState
  int confRoundsPerBatch = kConfRoundsPerBatch;  // B200SEED_CONF_ROUNDS (tests exercise the continuation path with 2)
  uint32_t recPerSpacePoint = 32;                // B200SEED_REC_PER_SP: first guess of the record pool
  size_t recCapacity = 0;
  int confRoundsLaunched = 0;
  ConfParams confParams{};
  CompactParams compactParams{};
  struct EnqueueArgs {
    uint32_t nEvents = 0, nTotal = 0;
    const uint32_t* dOffsets = nullptr;
    const float *x = nullptr, *y = nullptr, *z = nullptr, *r = nullptr, *varZ = nullptr, *varR = nullptr, *dPhi = nullptr;
    const float *hx = nullptr, *hy = nullptr, *hz = nullptr, *hr = nullptr;  // host copies of the columns, when the caller gave them
    const uint32_t* hOffsets = nullptr;
    int nZWin = 0;
    bool vertexCuts = false;            // VertexZCuts connected (cfg.useVertexZCuts, or windows given)
    const float* dStrip = nullptr;      // strip triplet path: raw calibration details on the device (NULL: pixel path)
    float cotThetaDiffMax = 0.f;
    const uint32_t* dZWinOffsets = nullptr;  // per-event window ranges (NULL: all events share [0, nZWin))
    uint32_t *outB = nullptr, *outM = nullptr, *outT = nullptr;
    float *outQ = nullptr, *outZ = nullptr;
    unsigned long long outCapacity = 0;
    unsigned long long* dSeedOffsets = nullptr;
    cudaStream_t stream = nullptr;
  } last;
  // pinned host mirrors
  unsigned long long* hCounters = nullptr;  // [kCntSlots]
  int* hStatus = nullptr;
  unsigned long long* hSeedTotal = nullptr;
  // state of the last call
  uint32_t lastEvents = 0, lastTotal = 0;
  int lastZWin = 0;
  unsigned long long lastCapacity = 0;
  const unsigned long long* lastSeedOffsets = nullptr;
  b200seed_counters lastCounters{};
  uint32_t lastConfRounds = 0;
  bool pending = false;
  uint64_t launches = 0;
  // stage boundaries of the last call: start | grid | work list | seeding | compaction
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // inside the seeding stage: after the count pass + chunk plan, then (after fill, after seeding) per chunk
  cudaEvent_t evCount = nullptr;
  std::vector<cudaEvent_t> evChunk;
  uint32_t chunksTimed = 0;
  float stageMs[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // grid, work list, seeding, compaction, doublet count, doublet fill, seed middles
};

namespace {

int upload(DevBuf& buf, const std::vector<uint32_t>& v, cudaStream_t s) {
  CUDA_TRY(buf.reserve(std::max<size_t>(4, v.size() * 4)));
  if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(buf.ptr, v.data(), v.size() * 4, cudaMemcpyHostToDevice, s));
  return B200SEED_OK;
}

int ensure_workspace(b200seed_handle* h, uint32_t nEvents, uint32_t nTotal) {
  const size_t nBinsAll = (size_t)nEvents * (size_t)h->plan.dev.nGlobalBins;
  const size_t nNavAll = (size_t)nEvents * h->plan.navBins.size();
  // per work item arrays: one item per space point at most (grid), two per space point (orthogonal: both z directions)
  const size_t nT = std::max<size_t>(h->plan.orthogonal ? 2 * (size_t)nTotal : (size_t)nTotal, 1);
  const size_t K = std::max<uint32_t>(h->plan.seedsPerMiddle, 1);
  CUDA_TRY(h->binOf.reserve(nT * 4));
  CUDA_TRY(h->binCount.reserve((nBinsAll + 1) * 4));
  CUDA_TRY(h->binStart.reserve((nBinsAll + 1) * 4));
  CUDA_TRY(h->binCursor.reserve((nBinsAll + 1) * 4));
  CUDA_TRY(h->tmpIdx.reserve(nT * 4));
  CUDA_TRY(h->pIdx.reserve(nT * 4));
  CUDA_TRY(h->pXY.reserve(nT * 8));
  CUDA_TRY(h->pZR.reserve(nT * 8));
  CUDA_TRY(h->pVar.reserve(nT * 8));
  CUDA_TRY(h->sortScratch.reserve(nT * 32));
  CUDA_TRY(h->midLo.reserve((nNavAll + 1) * 4));
  CUDA_TRY(h->midCount.reserve((nNavAll + 1) * 4));
  CUDA_TRY(h->workStart.reserve((nNavAll + 1) * 4));
  CUDA_TRY(h->workPos.reserve(nT * 4));
  CUDA_TRY(h->workEG.reserve(nT * 4));
  CUDA_TRY(h->workCounter.reserve(kChunkCounterWords * 4 * ((size_t)kMaxChunks + 2)));
  CUDA_TRY(h->capB.reserve(nT * 4));
  CUDA_TRY(h->capT.reserve(nT * 4));
  CUDA_TRY(h->slotPrefix.reserve((nT + 1) * 8));
  CUDA_TRY(h->hdr.reserve(nT * sizeof(MiddleHeader)));
  CUDA_TRY(h->carve.reserve(nT * sizeof(SeedCarve)));
  CUDA_TRY(h->classList.reserve(nT * 4 * (kNumSeedClasses + 1) * 2));
  if (h->planDev.ptr == nullptr) {  // the whole plan buffer is copied to the host each call, the bounds beyond the last chunk unread: define them once
    CUDA_TRY(h->planDev.reserve(((size_t)kMaxChunks + 1 + 8) * 4));
    CUDA_TRY(cudaMemset(h->planDev.ptr, 0, h->planDev.bytes));
  }
  CUDA_TRY(h->slotB.reserve(nT * K * 4));
  CUDA_TRY(h->slotM.reserve(nT * K * 4));
  CUDA_TRY(h->slotT.reserve(nT * K * 4));
  CUDA_TRY(h->slotQ.reserve(nT * K * 4));
  CUDA_TRY(h->slotZ.reserve(nT * K * 4));
  CUDA_TRY(h->slotCount.reserve(nT * 4));
  CUDA_TRY(h->seedStart.reserve((nT + 1) * 4));
  const size_t nTiles = (nT + kTile - 1) / kTile;
  CUDA_TRY(h->tileSums.reserve((nTiles + 1) * 4));
  CUDA_TRY(h->tilePrefix.reserve((nTiles + 1) * 4));
  CUDA_TRY(h->capTileSums.reserve((nTiles + 1) * 8));
  CUDA_TRY(h->capTilePrefix.reserve((nTiles + 1) * 8));
  CUDA_TRY(h->counters.reserve(kCntSlots * 8));
  CUDA_TRY(h->status.reserve(16));
  if (h->plan.dev.seedConfirmation) {
    if (h->recCapacity < (size_t)h->recPerSpacePoint * nT) h->recCapacity = (size_t)h->recPerSpacePoint * nT;  // grown on demand by finish()
    h->recCapacity = std::min<size_t>(h->recCapacity, 0xFFFFFFF0u);
    CUDA_TRY(h->rec.reserve(h->recCapacity * 16));
    CUDA_TRY(h->recZ.reserve(h->recCapacity * 4));
    CUDA_TRY(h->recBegin.reserve(nT * 4));
    CUDA_TRY(h->recCount.reserve(nT * 4));
    CUDA_TRY(h->slot2B.reserve(nT * K * 4));
    CUDA_TRY(h->slot2M.reserve(nT * K * 4));
    CUDA_TRY(h->slot2T.reserve(nT * K * 4));
    CUDA_TRY(h->slot2Q.reserve(nT * K * 4));
    CUDA_TRY(h->slot2Z.reserve(nT * K * 4));
    CUDA_TRY(h->slot2Count.reserve(nT * 4));
    CUDA_TRY(h->confHead.reserve(nT * 4));
    CUDA_TRY(h->confNext.reserve(nT * K * 3 * 4));
    CUDA_TRY(h->confState.reserve(kConfStateWords * 4));
    CUDA_TRY(h->confDirty.reserve(2 * nT));
  }
  return B200SEED_OK;
}

// Shared-memory classes of k_seed_middles: {threads, blocks per SM aimed at}.  A middle goes to the smallest class
// whose dynamic shared memory holds its lists (exact sizes, SeedCarve); the last class keeps them in global memory.
struct SeedClassShape { int threads, blocksPerSM; };
// (56 registers per thread: 36 warps fit an SM; thread counts measured, tools/sweep_regs.sh)
// Finer classes (9 ... 1 blocks per SM) were measured and lost: every class kernel has its own tail.
constexpr SeedClassShape kSeedClassShape[kNumSeedClasses] = {{192, 6}, {288, 4}, {384, 3}, {576, 2}, {1024, 1}, {1024, 1}};

using SeedKernel = void (*)(const SeedParams);
template <bool kConf, bool kStrip>
SeedKernel seed_kernel(int c) {
  return c == kSpillClass ? k_seed_middles<kConf, true, kStrip> : k_seed_middles<kConf, false, kStrip>;
}
SeedKernel seed_kernel(bool conf, int c, bool strip = false) {
  if (strip) return conf ? seed_kernel<true, true>(c) : seed_kernel<false, true>(c);
  return conf ? seed_kernel<true, false>(c) : seed_kernel<false, false>(c);
}

// Rounds [first, first + count) of the seedConfirmation fixed point (seeding_kernels.cuh, k_conf_replay).
// Round r reads the seeds of slot set (r + 1) & 1 and writes set r & 1.
int enqueue_conf_rounds(b200seed_handle* h, int first, int count, cudaStream_t s) {
  const uint32_t nTotal = std::max<uint32_t>(h->last.nTotal, 1);
  const uint32_t K = std::max<uint32_t>(h->plan.seedsPerMiddle, 1);
  uint32_t* setB[2] = {h->slotB.as<uint32_t>(), h->slot2B.as<uint32_t>()};
  uint32_t* setM[2] = {h->slotM.as<uint32_t>(), h->slot2M.as<uint32_t>()};
  uint32_t* setT[2] = {h->slotT.as<uint32_t>(), h->slot2T.as<uint32_t>()};
  float* setQ[2] = {h->slotQ.as<float>(), h->slot2Q.as<float>()};
  float* setZ[2] = {h->slotZ.as<float>(), h->slot2Z.as<float>()};
  uint32_t* setCount[2] = {h->slotCount.as<uint32_t>(), h->slot2Count.as<uint32_t>()};
  ConfParams cp = h->confParams;
  const int linkBlocks = std::max(1, std::min<int>((int)(((size_t)nTotal * K + 255) / 256), h->smCount * 8));
  const int replayBlocks = h->smCount * 16;
  for (int r = first; r < first + count; ++r) {
    const int cur = r & 1, prev = cur ^ 1;
    cp.round = r;
    cp.prevB = setB[prev]; cp.prevT = setT[prev]; cp.prevQ = setQ[prev]; cp.prevCount = setCount[prev];
    cp.curB = setB[cur]; cp.curM = setM[cur]; cp.curT = setT[cur]; cp.curQ = setQ[cur]; cp.curZ = setZ[cur];
    cp.curCount = setCount[cur];
    cp.prevZ = setZ[prev];
    cp.dirty = h->confDirty.as<uint8_t>() + (size_t)(r & 1) * nTotal;
    cp.dirtyNext = h->confDirty.as<uint8_t>() + (size_t)((r + 1) & 1) * nTotal;
    CUDA_TRY(cudaMemsetAsync(cp.dirtyNext, 0, nTotal, s));
    CUDA_TRY(cudaMemsetAsync(h->confHead.ptr, 0xFF, (size_t)nTotal * 4, s));
    if (r > 0) {
      k_conf_link<<<linkBlocks, 256, 0, s>>>(cp);
      ++h->launches;
    }
    if (std::max(h->plan.dev.maxSeedsPerSpMConf, h->plan.dev.maxQualitySeedsPerSpMConf) > (uint32_t)kMaxHeap) {
      k_conf_replay<kMaxHeapBig><<<replayBlocks, kConfWarps * 32, 0, s>>>(cp);
    } else {
      k_conf_replay<kMaxHeap><<<replayBlocks, kConfWarps * 32, 0, s>>>(cp);
    }
    ++h->launches;
  }
  h->confRoundsLaunched = first + count;
  return B200SEED_OK;
}

// Ordered seed compaction + the copies of everything the host reads after the sync.
int enqueue_tail(b200seed_handle* h, cudaStream_t s) {
  const CompactParams& cp = h->compactParams;
  const uint32_t nEvents = h->last.nEvents;
  const uint32_t nBinsAll = nEvents * (uint32_t)h->plan.dev.nGlobalBins;
  const uint32_t nTiles = std::max<uint32_t>(1, (std::max(h->itemsMax, h->last.nTotal) + kTile - 1) / kTile);
  k_tile_sums<<<nTiles, 256, 0, s>>>(cp);
  k_scan<<<1, kScanThreads, 0, s>>>(cp.tileSums, cp.tilePrefix, nTiles);
  k_compact_seeds<<<nTiles, 256, 0, s>>>(cp);
  k_event_offsets<<<1, 256, 0, s>>>(cp);
  h->launches += 4;
  CUDA_TRY(cudaEventRecord(h->ev[4], s));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(h->hCounters, h->counters.ptr, kCntSlots * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->hStatus, h->status.ptr, 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->hSeedTotal, h->last.dSeedOffsets + nEvents, 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->hCounters + kCntSlots, h->binStart.as<uint32_t>() + nBinsAll, 4, cudaMemcpyDeviceToHost, s));
  if (h->plan.dev.seedConfirmation) {
    CUDA_TRY(cudaMemcpyAsync(h->hConfState, h->confState.ptr, kConfStateWords * 4, cudaMemcpyDeviceToHost, s));
  }
  return B200SEED_OK;
}

// Orthogonal seeder, front of the pipeline (replaces the grid stage and the bin-wise work list): the host layer
// builds every event's k-d tree (kd_tree_host.hpp, one thread per event), the device gathers the packed copy in
// element order and lists two work items per accepted middle.
// Tree construction on the device (k_kd_select / k_kd_roots / k_kd_split per level / k_kd_small).  Leaves the element
// arrays (orthPosPhi = phi, orthPosOrig = index in the caller's event), the node array, coreOffsets and the middle
// radius ranges on the device; returns the number of elements of the batch.
int orth_build_device(b200seed_handle* h, uint64_t& launches, uint32_t& nCoreOut) {
  const b200seed_handle::EnqueueArgs& a = h->last;
  const uint32_t nEvents = a.nEvents, nTotal = a.nTotal;
  cudaStream_t s = a.stream;
  const size_t nT = std::max<size_t>(nTotal, 1);
  const bool selector = h->plan.dev.useExtraCuts != 0;
  const size_t maxActive = nT / 129 + nEvents + 2, maxSmall = nT / 2 + nEvents + 8;
  CUDA_TRY(h->orthPosOrig.reserve(nT * 4));
  CUDA_TRY(h->orthPosPhi.reserve(nT * 4));
  CUDA_TRY(h->orthElemR.reserve(nT * 4));
  CUDA_TRY(h->orthElemZ.reserve(nT * 4));
  CUDA_TRY(h->orthScratch.reserve((nT + 1) * 4 * 5));
  CUDA_TRY(h->orthNodes.reserve((2 * nT + nEvents + 2) * sizeof(KdNodeDev)));
  CUDA_TRY(h->orthCoreOffsets.reserve(((size_t)nEvents + 1) * 4));
  CUDA_TRY(h->orthRRange.reserve(std::max<size_t>(2 * (size_t)nEvents, 1) * 4));
  CUDA_TRY(h->orthTasks.reserve((2 * maxActive + maxSmall) * sizeof(KdTask)));
  CUDA_TRY(h->orthExtent.reserve(std::max<size_t>(2 * (size_t)nEvents, 1) * 8 + 64));
  KdBuildParams bp{};
  bp.cfg = h->plan.dev;
  bp.nEvents = nEvents; bp.nTotal = nTotal;
  bp.spOffsets = a.dOffsets;
  bp.x = a.x; bp.y = a.y; bp.z = a.z; bp.r = a.r;
  uint32_t* scratch = h->orthScratch.as<uint32_t>();
  bp.selFlag = scratch;
  bp.selScan = scratch + (nT + 1);
  bp.phiTmp = reinterpret_cast<float*>(scratch + 2 * (nT + 1));
  bp.listL = scratch + 3 * (nT + 1);
  bp.listR = scratch + 4 * (nT + 1);
  bp.coreOffsets = h->orthCoreOffsets.as<uint32_t>();
  bp.ePhi = h->orthPosPhi.as<float>(); bp.eR = h->orthElemR.as<float>(); bp.eZ = h->orthElemZ.as<float>();
  bp.eIdx = h->orthPosOrig.as<uint32_t>();
  bp.extent = h->orthExtent.as<unsigned long long>();
  bp.counters = reinterpret_cast<uint32_t*>(h->orthExtent.as<unsigned char>() + std::max<size_t>(2 * (size_t)nEvents, 1) * 8);
  bp.rMiddleRange = h->orthRRange.as<float>();
  bp.nodes = h->orthNodes.as<KdNodeDev>();
  KdTask* lists[2] = {h->orthTasks.as<KdTask>(), h->orthTasks.as<KdTask>() + maxActive};
  bp.tasksSmall = h->orthTasks.as<KdTask>() + 2 * maxActive;
  CUDA_TRY(cudaMemsetAsync(bp.extent, 0xFF, (size_t)nEvents * 8, s));
  CUDA_TRY(cudaMemsetAsync(bp.extent + nEvents, 0, (size_t)nEvents * 8, s));
  CUDA_TRY(cudaMemsetAsync(bp.counters, 0, 16, s));
  const int blocks = std::max(1, std::min<int>((int)((nTotal + 255) / 256), h->smCount * 8));
  if (nTotal > 0) {
    k_kd_select<<<blocks, 256, 0, s>>>(bp);
    ++launches;
    if (selector) {
      k_scan<<<1, kScanThreads, 0, s>>>(bp.selFlag, bp.selScan, nTotal);
      k_kd_compact<<<blocks, 256, 0, s>>>(bp);
      launches += 2;
    }
  } else if (selector) {
    CUDA_TRY(cudaMemsetAsync(bp.selScan, 0, 4, s));
  }
  int slot = 0;
  bp.tasksOut = lists[slot];
  bp.outSlot = (uint32_t)slot;
  k_kd_roots<<<1, 256, 0, s>>>(bp);
  ++launches;
  uint32_t hc[4] = {0, 0, 0, 0};
  std::vector<uint32_t> coreOffsets((size_t)nEvents + 1);
  CUDA_TRY(cudaMemcpyAsync(hc, bp.counters, 16, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(coreOffsets.data(), bp.coreOffsets, coreOffsets.size() * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  uint32_t nActive = hc[2 + slot];
  uint32_t level = 0;
  while (nActive > 0) {  // one launch per tree level that still has nodes of more than 128 elements
    if (nActive > maxActive) return fail(B200SEED_ERR_RUNTIME, "k-d tree construction: task list overflow");
    // A node of more than 128 points with identical (phi, r, z) never splits: the reference's constructor recurses
    // without end on it (KDTree.hpp:299-324: the whole range becomes the left child again).  Refuse instead of looping.
    if (++level > nTotal / 4 + 64) {
      return fail(B200SEED_ERR_INVALID_ARGUMENT, "k-d tree construction does not terminate: more than 128 space points with identical (phi, r, z)");
    }
    bp.tasksIn = lists[slot];
    bp.nTasksIn = nActive;
    bp.tasksOut = lists[slot ^ 1];
    bp.outSlot = (uint32_t)(slot ^ 1);
    CUDA_TRY(cudaMemsetAsync(bp.counters + 2 + (slot ^ 1), 0, 4, s));
    k_kd_split<<<nActive, kKdSplitThreads, 0, s>>>(bp);
    ++launches;
    CUDA_TRY(cudaMemcpyAsync(hc, bp.counters, 16, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    slot ^= 1;
    nActive = hc[2 + slot];
  }
  const uint32_t nSmall = hc[1];
  if (nSmall > maxSmall) return fail(B200SEED_ERR_RUNTIME, "k-d tree construction: small-task list overflow");
  if (nSmall > 0) {
    k_kd_small<<<(nSmall + kKdSmallWarps - 1) / kKdSmallWarps, kKdSmallWarps * 32, 0, s>>>(bp);
    ++launches;
  }
  CUDA_TRY(cudaGetLastError());
  nCoreOut = coreOffsets[nEvents];
  CUDA_TRY(cudaMemcpyAsync(h->binStart.ptr, bp.coreOffsets + nEvents, 4, cudaMemcpyDeviceToDevice, s));  // reported as nInGrid
  return B200SEED_OK;
}

int orth_front_tail(b200seed_handle* h, GridParams& gp, WorkParams& wp, KdDoubletParams& kdp, uint64_t& launches, uint32_t nCore);

int orth_front(b200seed_handle* h, GridParams& gp, WorkParams& wp, KdDoubletParams& kdp, uint64_t& launches) {
  if (!h->kdHostBuild) {
    uint32_t nCore = 0;
    const int rc = orth_build_device(h, launches, nCore);
    if (rc != B200SEED_OK) return rc;
    CUDA_TRY(h->midCount.reserve(((size_t)nCore + 1) * 4));
    CUDA_TRY(h->workStart.reserve(((size_t)nCore + 2) * 4));
    return orth_front_tail(h, gp, wp, kdp, launches, nCore);
  }
  const b200seed_handle::EnqueueArgs& a = h->last;
  const uint32_t nEvents = a.nEvents, nTotal = a.nTotal;
  cudaStream_t s = a.stream;
  const HostPlan& plan = h->plan;
  // host copies of x, y, z, r (the caller's, or read back when the inputs are device resident)
  std::vector<float> back[4];
  std::vector<uint32_t> backOffsets;
  const float* hc[4] = {a.hx, a.hy, a.hz, a.hr};
  const uint32_t* hOff = a.hOffsets;
  if (hOff == nullptr || (nTotal > 0 && (hc[0] == nullptr || hc[1] == nullptr || hc[2] == nullptr || hc[3] == nullptr))) {
    const float* dc[4] = {a.x, a.y, a.z, a.r};
    backOffsets.resize((size_t)nEvents + 1);
    CUDA_TRY(cudaMemcpyAsync(backOffsets.data(), a.dOffsets, ((size_t)nEvents + 1) * 4, cudaMemcpyDeviceToHost, s));
    for (int k = 0; k < 4; ++k) {
      back[k].resize(std::max<size_t>(nTotal, 1));
      if (nTotal > 0) CUDA_TRY(cudaMemcpyAsync(back[k].data(), dc[k], (size_t)nTotal * 4, cudaMemcpyDeviceToHost, s));
      hc[k] = back[k].data();
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    hOff = backOffsets.data();
  }
  std::vector<KdEventTree> trees(nEvents);
  {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const uint32_t nThreads = std::min<uint32_t>(nEvents, std::min<unsigned>(hw, 32u));
    std::atomic<uint32_t> next{0};
    auto worker = [&] {
      for (;;) {
        const uint32_t e = next.fetch_add(1);
        if (e >= nEvents) return;
        const uint32_t b = hOff[e], n = hOff[e + 1] - b;
        build_kd_event(plan.dev, n, hc[0] + b, hc[1] + b, hc[2] + b, hc[3] + b, trees[e]);
      }
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < nThreads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
  }
  for (const KdEventTree& t : trees) {
    if (t.runaway) {
      return fail(B200SEED_ERR_INVALID_ARGUMENT, "k-d tree construction does not terminate: more than 128 space points with identical (phi, r, z)");
    }
  }
  // batch-wide node array: node e = root of event e, the other nodes of event e follow from base[e]
  std::vector<uint32_t> coreOffsets(nEvents + 1, 0), nodeBase(nEvents + 1, nEvents);
  for (uint32_t e = 0; e < nEvents; ++e) {
    coreOffsets[e + 1] = coreOffsets[e] + (uint32_t)trees[e].posOrig.size();
    nodeBase[e + 1] = nodeBase[e] + (uint32_t)std::max<size_t>(trees[e].nodes.size(), 1) - 1;
  }
  const uint32_t nCore = coreOffsets[nEvents], nNodes = nodeBase[nEvents];
  std::vector<uint32_t> posOrig(std::max<uint32_t>(nCore, 1));
  std::vector<float> posPhi(std::max<uint32_t>(nCore, 1)), rRange(2 * (size_t)nEvents);
  std::vector<KdNodeDev> nodes(std::max<uint32_t>(nNodes, 1));
  for (uint32_t e = 0; e < nEvents; ++e) {
    std::copy(trees[e].posOrig.begin(), trees[e].posOrig.end(), posOrig.begin() + coreOffsets[e]);
    std::copy(trees[e].posPhi.begin(), trees[e].posPhi.end(), posPhi.begin() + coreOffsets[e]);
    const uint32_t nLocal = (uint32_t)trees[e].nodes.size();
    auto global = [&](uint32_t local) { return local == 0 ? e : (local >= nLocal ? kKdEnd : nodeBase[e] + local - 1); };
    if (nLocal == 0) {  // an event without selected space points: an empty leaf that overlaps no box
      KdNodeDev leaf{};
      for (int j = 0; j < 3; ++j) { leaf.mn[j] = std::numeric_limits<float>::max(); leaf.mx[j] = std::numeric_limits<float>::lowest(); }
      leaf.begin = leaf.end = coreOffsets[e];
      leaf.skip = kKdEnd;
      nodes[e] = leaf;
    }
    for (uint32_t k = 0; k < nLocal; ++k) {
      KdNodeDev nd = trees[e].nodes[k];
      nd.begin += coreOffsets[e];
      nd.end += coreOffsets[e];
      nd.skip = global(nd.skip);
      nd.lhs = nd.internal != 0u ? global(nd.lhs) : 0u;
      nodes[global(k)] = nd;
    }
    rRange[2 * e] = trees[e].rMiddleMin;
    rRange[2 * e + 1] = trees[e].rMiddleMax;
  }
  CUDA_TRY(h->orthPosOrig.reserve(posOrig.size() * 4));
  CUDA_TRY(h->orthPosPhi.reserve(posPhi.size() * 4));
  CUDA_TRY(h->orthNodes.reserve(nodes.size() * sizeof(KdNodeDev)));
  CUDA_TRY(h->orthCoreOffsets.reserve(coreOffsets.size() * 4));
  CUDA_TRY(h->orthRRange.reserve(std::max<size_t>(rRange.size(), 1) * 4));
  CUDA_TRY(h->midCount.reserve(((size_t)nCore + 1) * 4));
  CUDA_TRY(h->workStart.reserve(((size_t)nCore + 2) * 4));
  CUDA_TRY(cudaMemcpyAsync(h->orthPosOrig.ptr, posOrig.data(), posOrig.size() * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->orthPosPhi.ptr, posPhi.data(), posPhi.size() * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->orthNodes.ptr, nodes.data(), nodes.size() * sizeof(KdNodeDev), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->orthCoreOffsets.ptr, coreOffsets.data(), coreOffsets.size() * 4, cudaMemcpyHostToDevice, s));
  if (!rRange.empty()) CUDA_TRY(cudaMemcpyAsync(h->orthRRange.ptr, rRange.data(), rRange.size() * 4, cudaMemcpyHostToDevice, s));
  // the number of selected space points is what the grid path reports as nInGrid (read back from binStart[0])
  CUDA_TRY(cudaMemcpyAsync(h->binStart.ptr, &coreOffsets[nEvents], 4, cudaMemcpyHostToDevice, s));

  CUDA_TRY(cudaStreamSynchronize(s));  // the staging vectors go out of scope
  return orth_front_tail(h, gp, wp, kdp, launches, nCore);
}

// gather in element order, middle selection, work list (two items per accepted middle)
int orth_front_tail(b200seed_handle* h, GridParams& gp, WorkParams& wp, KdDoubletParams& kdp, uint64_t& launches, uint32_t nCore) {
  const b200seed_handle::EnqueueArgs& a = h->last;
  const uint32_t nEvents = a.nEvents;
  cudaStream_t s = a.stream;
  const HostPlan& plan = h->plan;
  CUDA_TRY(cudaEventRecord(h->ev[1], s));
  OrthParams op{};
  op.cfg = plan.dev;
  op.orth = plan.orth;
  op.nEvents = nEvents; op.nCoreTotal = nCore;
  op.spOffsets = a.dOffsets;
  op.coreOffsets = h->orthCoreOffsets.as<uint32_t>();
  op.rMiddleRange = h->orthRRange.as<float>();
  op.posOrig = h->orthPosOrig.as<uint32_t>();
  op.posPhi = h->orthPosPhi.as<float>();
  op.nodes = h->orthNodes.as<KdNodeDev>();
  op.x = a.x; op.y = a.y; op.z = a.z; op.r = a.r; op.varZ = a.varZ; op.varR = a.varR;
  op.pIdx = gp.pIdx; op.pXY = gp.pXY; op.pZR = gp.pZR; op.pVar = gp.pVar;
  op.itemCount = h->midCount.as<uint32_t>();
  op.workStart = h->workStart.as<uint32_t>();
  op.workPos = h->workPos.as<uint32_t>();
  op.workEG = h->workEG.as<uint32_t>();
  wp.workStart = op.workStart;
  wp.workPos = op.workPos;
  wp.workEG = op.workEG;
  const int blocks = std::max(1, std::min<int>((int)((nCore + 255) / 256), h->smCount * 8));
  if (nCore > 0) {
    k_orth_gather<<<blocks, 256, 0, s>>>(op);
    ++launches;
  }
  k_scan<<<1, kScanThreads, 0, s>>>(op.itemCount, op.workStart, nCore);
  ++launches;
  if (nCore > 0) {
    k_orth_fill_work<<<blocks, 256, 0, s>>>(op);
    ++launches;
  }
  kdp.orth = plan.orth;
  kdp.nodes = op.nodes;
  kdp.posPhi = op.posPhi;
  kdp.d.nNav = nCore;  // read by enqueue(): the work-item total sits at workStart[nCore]
  return B200SEED_OK;
}

// Enqueue the whole pipeline for a batch whose inputs already live on the device (h->last holds the arguments).
int enqueue(b200seed_handle* h) {
  const b200seed_handle::EnqueueArgs& a = h->last;
  const uint32_t nEvents = a.nEvents, nTotal = a.nTotal;
  const uint32_t* dOffsets = a.dOffsets;
  const float *x = a.x, *y = a.y, *z = a.z, *r = a.r, *varZ = a.varZ, *varR = a.varR, *dPhi = a.dPhi;
  const int nZWin = a.nZWin;
  cudaStream_t s = a.stream;
  int rc = ensure_workspace(h, nEvents, nTotal);
  if (rc != B200SEED_OK) return rc;
  const HostPlan& plan = h->plan;
  const uint32_t nBins = (uint32_t)plan.dev.nGlobalBins;
  const uint32_t nNav = (uint32_t)plan.navBins.size();
  const uint32_t nBinsAll = nEvents * nBins, nNavAll = nEvents * nNav;
  uint64_t launches = 0;

  CUDA_TRY(cudaMemsetAsync(h->binCount.ptr, 0, ((size_t)nBinsAll + 1) * 4, s));
  CUDA_TRY(cudaMemsetAsync(h->binCursor.ptr, 0, ((size_t)nBinsAll + 1) * 4, s));
  CUDA_TRY(cudaMemsetAsync(h->workCounter.ptr, 0, kChunkCounterWords * 4 * ((size_t)kMaxChunks + 2), s));
  CUDA_TRY(cudaMemsetAsync(h->counters.ptr, 0, kCntSlots * 8, s));
  CUDA_TRY(cudaMemsetAsync(h->status.ptr, 0, 16, s));

  CUDA_TRY(cudaEventRecord(h->ev[0], s));
  GridParams gp{};
  gp.cfg = plan.dev;
  gp.nEvents = nEvents; gp.nTotal = nTotal; gp.nBins = nBins;
  gp.spOffsets = dOffsets;
  gp.x = x; gp.y = y; gp.z = z; gp.r = r; gp.varZ = varZ; gp.varR = varR;
  gp.phi = dPhi;
  gp.binOf = h->binOf.as<uint32_t>();
  gp.binCount = h->binCount.as<uint32_t>();
  gp.binStart = h->binStart.as<uint32_t>();
  gp.binCursor = h->binCursor.as<uint32_t>();
  gp.tmpIdx = h->tmpIdx.as<uint32_t>();
  gp.pIdx = h->pIdx.as<uint32_t>();
  gp.pXY = h->pXY.as<float2>();
  gp.pZR = h->pZR.as<float2>();
  gp.pVar = h->pVar.as<float2>();
  gp.sortScratch = h->sortScratch.as<unsigned long long>();
  gp.sortSmemCap = h->sortSmemCap;
  gp.exactTies = h->exactTies;
  gp.status = h->status.as<int>();
  gp.counters = h->counters.as<unsigned long long>();

  const bool orthogonal = plan.orthogonal;
  const uint32_t itemsMax = orthogonal ? 2u * nTotal : nTotal;
  h->itemsMax = itemsMax;
  const uint32_t* nWorkPtr = nullptr;       // device: number of work items of this call
  const uint32_t* eventFirstItem = nullptr;  // orthogonal: index into workStart of every event's first item
  KdDoubletParams kdp{};
  WorkParams wp{};
  wp.cfg = plan.dev;
  wp.nEvents = nEvents; wp.nBins = nBins; wp.nNav = nNav;
  wp.binStart = gp.binStart;
  wp.pZR = gp.pZR;
  wp.navBins = h->navBins.as<uint32_t>();
  wp.midLo = h->midLo.as<uint32_t>();
  wp.midCount = h->midCount.as<uint32_t>();
  wp.workStart = h->workStart.as<uint32_t>();
  wp.workPos = h->workPos.as<uint32_t>();
  wp.workEG = h->workEG.as<uint32_t>();
  wp.phiFirst = h->phiFirst;
  wp.phiCount = h->phiCount;
  if (!orthogonal) {
  const int elemBlocks = std::max(1, std::min<int>((int)((nTotal + 255) / 256), h->smCount * 8));
  if (nTotal > 0) {
    k_bin_count<<<elemBlocks, 256, 0, s>>>(gp);
    ++launches;
  }
  k_scan<<<1, kScanThreads, 0, s>>>(gp.binCount, gp.binStart, nBinsAll);
  ++launches;
  if (nTotal > 0) {
    k_scatter<<<elemBlocks, 256, 0, s>>>(gp);
    k_sort_bins<<<nBinsAll, kSortThreads, (size_t)h->sortSmemCap * 16 + ((size_t)kSortBuckets + 1) * 4, s>>>(gp);
    launches += 2;
  }
  if (a.dStrip != nullptr && nTotal > 0) {  // strip triplet path: derived calibration details in packed order
    CUDA_TRY(h->pStrip.reserve((size_t)nTotal * sizeof(StripDerived)));
    k_gather_strips<<<elemBlocks, 256, 0, s>>>(gp.pIdx, gp.binStart + nBinsAll, a.dStrip, h->pStrip.as<StripDerived>());
    ++launches;
  }

  CUDA_TRY(cudaEventRecord(h->ev[1], s));
  k_middle_ranges<<<nEvents, 256, 0, s>>>(wp);
  k_scan<<<1, kScanThreads, 0, s>>>(wp.midCount, wp.workStart, nNavAll);
  k_fill_work<<<(nNavAll * 32 + 255) / 256, 256, 0, s>>>(wp);
  launches += 3;

  nWorkPtr = wp.workStart + nNavAll;
  } else {
    rc = orth_front(h, gp, wp, kdp, launches);
    if (rc != B200SEED_OK) return rc;
    nWorkPtr = wp.workStart + kdp.d.nNav;  // (orth_front parks the element count there)
    eventFirstItem = h->orthCoreOffsets.as<uint32_t>();
  }
  CUDA_TRY(cudaEventRecord(h->ev[2], s));
  const bool conf = plan.dev.seedConfirmation != 0;
  const uint32_t nWorkMax = std::max<uint32_t>(itemsMax, 1);
  uint32_t* wc = h->workCounter.as<uint32_t>();  // 16 words per launch group: [0] count pass, [1 + c] chunk c
  uint32_t* planWords = h->planDev.as<uint32_t>();
  CUDA_TRY(cudaMemsetAsync(planWords, 0, 8 * 4, s));

  // ---- doublet stage, pass 1: slot sizes ------------------------------------
  DoubletParams dp{};
  dp.cfg = plan.dev;
  if (h->last.vertexCuts) dp.cfg.doubletCuts = kCutsVertexZ;  // takes the experimentCuts slot, .cpp:291-296
  dp.pXY = gp.pXY; dp.pZR = gp.pZR; dp.pVar = gp.pVar;
  dp.binStart = gp.binStart;
  dp.navBins = h->navBins.as<uint32_t>();
  dp.botOffsets = h->botOffsets.as<uint32_t>();
  dp.botBins = h->botBins.as<uint32_t>();
  dp.topOffsets = h->topOffsets.as<uint32_t>();
  dp.topBins = h->topBins.as<uint32_t>();
  dp.workPos = wp.workPos; dp.workEG = wp.workEG;
  dp.nWorkPtr = nWorkPtr;
  dp.nNav = nNav; dp.nBins = nBins;
  dp.zWinLo = h->zWin.as<float>();
  dp.zWinHi = h->zWin.as<float>() + h->zWinCapacity;
  dp.zWinOffsets = a.dZWinOffsets;
  dp.nZWin = nZWin;
  dp.workCounter = wc;
  dp.capB = h->capB.as<uint32_t>(); dp.capT = h->capT.as<uint32_t>();
  dp.planWords = planWords;
  dp.slotPrefix = h->slotPrefix.as<unsigned long long>();
  dp.hdr = h->hdr.as<MiddleHeader>();
  dp.carve = h->carve.as<SeedCarve>();
  dp.slotCount = h->slotCount.as<uint32_t>();
  dp.classList = h->classList.as<uint32_t>();
  dp.classStride = nWorkMax;
  for (int c = 0; c < kNumSeedClasses; ++c) dp.classBytes[c] = h->classBytes[c];
  dp.conf = conf ? 1 : 0;
  dp.counters = gp.counters;
  dp.status = gp.status;
  if (!orthogonal && h->maskWordsPerSp != 0u) {
    const size_t words = std::min<size_t>(std::max<size_t>((size_t)nTotal * h->maskWordsPerSp, 1u << 16), 0xFFFFFFF0u);
    CUDA_TRY(h->maskArena.reserve(words * 4));
    CUDA_TRY(h->maskOff.reserve((size_t)nWorkMax * 4));
    dp.maskArena = h->maskArena.as<uint32_t>();
    dp.maskCapacity = (uint32_t)words;
    dp.maskCursor = planWords + 7;  // zeroed with the plan words
    dp.maskOff = h->maskOff.as<uint32_t>();
  }
  if (orthogonal) {
    kdp.d = dp;
    if (h->kdHitMB != 0u) {
      const size_t bytes = (size_t)h->kdHitMB << 20;
      if (h->kdHitArena.bytes < bytes) {
        // (zeroed once: the list kernel reads whole pages, i.e. also the words behind the end mark of a walk's last page)
        CUDA_TRY(h->kdHitArena.reserve(bytes));
        CUDA_TRY(cudaMemsetAsync(h->kdHitArena.ptr, 0, h->kdHitArena.bytes, s));
      }
      CUDA_TRY(h->kdHitHead.reserve((size_t)nWorkMax * 8));
      kdp.hitArena = h->kdHitArena.as<uint32_t>();
      kdp.hitPages = (uint32_t)std::min<size_t>(bytes / 128, 0xFFFFFFF0u);
      kdp.hitCursor = planWords + 7;  // zeroed with the plan words
      kdp.hitHead = h->kdHitHead.as<uint32_t>();
    }
    k_doublets_kd<false><<<h->smCount * h->kdBlocksPerSM[0], kKdThreads, 0, s>>>(kdp);
  } else {
    k_doublets<false><<<h->smCount * h->doubletBlocksPerSM[0], kDoubletWarps * 32, 0, s>>>(dp);
  }
  ++launches;

  // ---- slot prefix, chunk plan; the host reads the plan (the one synchronisation inside a call) ---
  const uint32_t nTiles = std::max<uint32_t>(1, (itemsMax + kTile - 1) / kTile);
  const int nStreams = h->chunkStreams;
  const unsigned long long arenaRecordsMax =
      std::max<unsigned long long>(h->arenaMaxBytes / 36ull / (unsigned long long)nStreams, 4ull * kMaxListLength);
  SlotScanParams ssp{};
  ssp.nWorkPtr = dp.nWorkPtr;
  ssp.capB = dp.capB; ssp.capT = dp.capT;
  ssp.tileSums = h->capTileSums.as<unsigned long long>();
  ssp.tilePrefix = h->capTilePrefix.as<unsigned long long>();
  ssp.slotPrefix = h->slotPrefix.as<unsigned long long>();
  ssp.arenaRecords = arenaRecordsMax;
  ssp.chunkBounds = planWords + 8;
  ssp.planWords = planWords;
  ssp.status = gp.status;
  k_cap_tile_sums<<<nTiles, 256, 0, s>>>(ssp);
  k_scan_u64<<<1, 32, 0, s>>>(ssp.tileSums, ssp.tilePrefix, nTiles);
  k_cap_prefix<<<nTiles, 256, 0, s>>>(ssp);
  k_plan_chunks<<<1, 1, 0, s>>>(ssp);
  launches += 4;
  CUDA_TRY(cudaEventRecord(h->evCount, s));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(h->hPlan, planWords, (8 + (size_t)kMaxChunks + 1) * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h->hStatus, h->status.ptr, 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (*h->hStatus & kStatusOverflowDoublets) {
    return fail(B200SEED_ERR_OVERFLOW, "a middle space point with more than " + std::to_string(kMaxListLength) +
                                           " doublet candidates on one side (16-bit ranks)");
  }
  if (*h->hStatus & kStatusArenaTooSmall) {
    return fail(B200SEED_ERR_OVERFLOW, "doublet arena too small for this batch: raise B200SEED_ARENA_MB");
  }
  const uint32_t nChunks = h->hPlan[3];
  const uint32_t* bounds = h->hPlan + 8;
  h->lastChunkBounds.assign(bounds, bounds + nChunks + 1);
  // arena: the largest chunk (its slots are the prefix difference, read back below with the plan)
  const unsigned long long chunkRecordsMax = (unsigned long long)h->hPlan[5] | ((unsigned long long)h->hPlan[6] << 32);
  for (int a = 0; a < (nChunks > 1 ? nStreams : 1); ++a) {
    CUDA_TRY(h->arenaRec[a].reserve(std::max<size_t>(64, (size_t)chunkRecordsMax * sizeof(DoubletRecord))));
    CUDA_TRY(h->arenaKey[a].reserve(std::max<size_t>(64, (size_t)chunkRecordsMax * 4)));
  }
  // spill class: per-block scratch for the largest lists of the batch + a candidate pool
  const uint32_t maxCapB = h->hPlan[1], maxCapT = h->hPlan[2];
  const SeedCarve spillCarve = seed_carve(std::max<uint32_t>(maxCapB, 1), std::max<uint32_t>(maxCapT, 1));
  const uint32_t spillPool = std::max<uint32_t>(4u * seed_pool_min(maxCapB), 1u << 18);
  const uint32_t spillBytes = carve_align(spillCarve.oPool + kPoolEntryBytes * spillPool + 64);
  const bool spillLikely = h->hPlan[0] > h->classBytes[kSpillClass - 1];
  const int spillBlocks = spillLikely ? h->smCount : 16;
  CUDA_TRY(h->spillScratch.reserve((size_t)spillBytes * (size_t)spillBlocks * (size_t)nStreams));

  SeedParams sp{};
  sp.cfg = dp.cfg;
  sp.pXY = gp.pXY; sp.pZR = gp.pZR; sp.pVar = gp.pVar;
  sp.workPos = wp.workPos;
  sp.hdr = dp.hdr;
  sp.carve = dp.carve;
  sp.spillScratch = h->spillScratch.as<unsigned char>();
  sp.slotB = h->slotB.as<uint32_t>(); sp.slotM = h->slotM.as<uint32_t>(); sp.slotT = h->slotT.as<uint32_t>();
  sp.slotQ = h->slotQ.as<float>(); sp.slotZ = h->slotZ.as<float>();
  sp.slotCount = h->slotCount.as<uint32_t>();
  sp.seedsPerMiddle = std::max<uint32_t>(plan.seedsPerMiddle, 1);
  sp.exactTies = h->exactTies;
  const bool strip = a.dStrip != nullptr;
  sp.pStrip = h->pStrip.as<StripDerived>();
  sp.cotThetaDiffMax2 = a.cotThetaDiffMax * a.cotThetaDiffMax;  // :232-233 (binary32 product, also in the relaxed engine)
  sp.toleranceParam = plan.toleranceParam;
  sp.bigHeap = h->heapBytes != 0u ? 1u : 0u;
  sp.counters = gp.counters;
  sp.status = gp.status;
  if (conf) {
    sp.rec4 = h->rec.as<uint4>();
    sp.recZ = h->recZ.as<float>();
    sp.recBegin = h->recBegin.as<uint32_t>();
    sp.recCount = h->recCount.as<uint32_t>();
    sp.recCounter = h->confState.as<uint32_t>();
    sp.recCapacity = (uint32_t)h->recCapacity;
    CUDA_TRY(cudaMemsetAsync(h->recCount.ptr, 0, (size_t)nWorkMax * 4, s));
    CUDA_TRY(cudaMemsetAsync(h->confState.ptr, 0, kConfStateWords * 4, s));
  }
  // ---- per chunk: doublet fill, then the seeding kernel of every shared-memory class --------
  while (h->evChunk.size() < 2 * (size_t)nChunks) {
    cudaEvent_t e = nullptr;
    CUDA_TRY(cudaEventCreate(&e));
    h->evChunk.push_back(e);
  }
  h->chunksTimed = nChunks;
  const bool overlap = nStreams > 1 && nChunks > 1;
  if (overlap) {
    CUDA_TRY(cudaEventRecord(h->evPlan, s));
    for (int a = 0; a < nStreams; ++a) CUDA_TRY(cudaStreamWaitEvent(h->chunkStream[a], h->evPlan, 0));
  }
  for (uint32_t c = 0; c < nChunks; ++c) {
    const int a = overlap ? (int)(c % (uint32_t)nStreams) : 0;
    cudaStream_t cs = overlap ? h->chunkStream[a] : s;
    uint32_t* cw = wc + kChunkCounterWords * (c + 1);
    // lists of the chunk: [k] class k (filled by the doublet pass), [kNumSeedClasses] middles whose candidate pool
    // overflowed in a shared-memory class (re-run with the largest shared memory, then, if need be, by the spill class)
    uint32_t* lists = dp.classList + (size_t)a * (kNumSeedClasses + 1) * dp.classStride;
    DoubletParams dpc = dp;
    dpc.itemFirst = bounds[c];
    dpc.itemEnd = bounds[c + 1];
    dpc.workCounter = cw;
    dpc.classCount = cw + 16;
    dpc.classList = lists;
    dpc.rec = h->arenaRec[a].as<DoubletRecord>();
    dpc.key = h->arenaKey[a].as<float>();
    if (orthogonal) {
      KdDoubletParams kdc = kdp;
      kdc.d = dpc;
      if (kdc.hitPages != 0u) {  // items whose walks left hit lists: one warp per item, no tree walk (cw[8]: its ticket)
        k_doublets_kd_lists<<<h->smCount * 4, kKdListWarps * 32, 0, cs>>>(kdc, cw + 8);
        ++launches;
      }
      k_doublets_kd<true><<<h->smCount * h->kdBlocksPerSM[1], kKdThreads, 0, cs>>>(kdc);
    } else {
      k_doublets<true><<<h->smCount * h->doubletBlocksPerSM[1], kDoubletWarps * 32, 0, cs>>>(dpc);
    }
    CUDA_TRY(cudaEventRecord(h->evChunk[2 * c], cs));
    sp.rec = dpc.rec;
    sp.key = dpc.key;
    sp.spillScratch = h->spillScratch.as<unsigned char>() + (size_t)a * spillBytes * (size_t)spillBlocks;
    const int kLargest = kSpillClass - 1;
    auto launchClass = [&](int k, int list, int ticket, int overflowTo, cudaStream_t st) {
      const bool spill = k == kSpillClass;
      sp.workList = lists + (size_t)list * dp.classStride;
      sp.nWorkPtr = cw + 16 + list;
      sp.workCounter = cw + 1 + ticket;
      sp.overflowList = overflowTo < 0 ? nullptr : lists + (size_t)overflowTo * dp.classStride;
      sp.overflowCount = overflowTo < 0 ? nullptr : cw + 16 + overflowTo;
      sp.arrayBytes = spill ? spillBytes : h->classBytes[k];
      const int blocks = spill ? spillBlocks : h->smCount * h->classBlocksPerSM[k];
      seed_kernel(conf, k, strip)<<<blocks, h->classThreads[k], (spill ? 0 : h->classBytes[k]) + h->heapBytes, st>>>(sp);
    };
    const bool fan = h->classStreams != 0;
    if (fan) CUDA_TRY(cudaEventRecord(h->evFill[a], cs));
    for (int k = 0; k <= kLargest; ++k) {  // the shared-memory classes, concurrently
      cudaStream_t st = fan ? h->classStream[a][k] : cs;
      if (fan) CUDA_TRY(cudaStreamWaitEvent(st, h->evFill[a], 0));
      launchClass(k, k, k, k == kLargest ? kSpillClass : kNumSeedClasses, st);
      if (fan) {
        CUDA_TRY(cudaEventRecord(h->evClass[a][k], st));
        CUDA_TRY(cudaStreamWaitEvent(cs, h->evClass[a][k], 0));
      }
    }
    launchClass(kLargest, kNumSeedClasses, kNumSeedClasses, kSpillClass, cs);  // the rare pool overflows
    launchClass(kSpillClass, kSpillClass, kSpillClass, -1, cs);
    launches += 1;  // (the re-run of the overflow list)
    CUDA_TRY(cudaEventRecord(h->evChunk[2 * c + 1], cs));
    launches += 1 + kNumSeedClasses;
    if (c + 1 == nChunks) h->lastDoublets = dpc;
  }
  if (overlap) {  // the caller's stream continues when both chunk streams are done
    for (int a = 0; a < nStreams; ++a) {
      CUDA_TRY(cudaEventRecord(h->evChunkEnd[a], h->chunkStream[a]));
      CUDA_TRY(cudaStreamWaitEvent(s, h->evChunkEnd[a], 0));
    }
  }
  if (nChunks == 0) h->lastDoublets = dp;
  CUDA_TRY(cudaGetLastError());
  sp.nWorkPtr = nWorkPtr;
  h->launches = launches;
  if (conf) {
    ConfParams& cf = h->confParams;
    cf = ConfParams{};
    cf.cfg = plan.dev;
    cf.nWorkPtr = sp.nWorkPtr;
    cf.workPos = sp.workPos;
    cf.rec = sp.rec4; cf.recZ = sp.recZ; cf.recBegin = sp.recBegin; cf.recCount = sp.recCount;
    cf.head = h->confHead.as<int>();
    cf.next = h->confNext.as<int>();
    cf.seedsPerMiddle = sp.seedsPerMiddle;
    cf.changed = h->confState.as<uint32_t>() + kConfChangedBase;
    rc = enqueue_conf_rounds(h, 0, h->confRoundsPerBatch, s);
    if (rc != B200SEED_OK) return rc;
  }

  CUDA_TRY(cudaEventRecord(h->ev[3], s));
  CompactParams& cp = h->compactParams;
  cp = CompactParams{};
  cp.nWorkPtr = sp.nWorkPtr;
  cp.slotCount = sp.slotCount;
  cp.tileSums = h->tileSums.as<uint32_t>();
  cp.tilePrefix = h->tilePrefix.as<uint32_t>();
  cp.slotB = sp.slotB; cp.slotM = sp.slotM; cp.slotT = sp.slotT; cp.slotQ = sp.slotQ; cp.slotZ = sp.slotZ;
  cp.seedsPerMiddle = sp.seedsPerMiddle;
  cp.pIdx = gp.pIdx;
  cp.outB = a.outB; cp.outM = a.outM; cp.outT = a.outT; cp.outQ = a.outQ; cp.outZ = a.outZ;
  cp.outCapacity = a.outCapacity;
  cp.seedOffsets = a.dSeedOffsets;
  cp.workStart = wp.workStart;
  cp.eventFirstItem = eventFirstItem;
  cp.seedStart = h->seedStart.as<uint32_t>();
  cp.nEvents = nEvents; cp.nNav = nNav;
  cp.counters = gp.counters;
  rc = enqueue_tail(h, s);
  if (rc != B200SEED_OK) return rc;
  h->lastEvents = nEvents;
  h->lastTotal = nTotal;
  h->lastZWin = nZWin;
  h->lastCapacity = a.outCapacity;
  h->pending = true;
  return B200SEED_OK;
}

int finish(b200seed_handle* h, cudaStream_t s, b200seed_seeds* out) {
  CUDA_TRY(cudaStreamSynchronize(s));
  if (h->plan.dev.seedConfirmation) {
    // the two host decisions of the seedConfirmation path: grow the record pool, run more rounds
    for (;;) {
      if (*h->hStatus & kStatusOverflowRecords) {
        const size_t need = h->hConfState[0];
        if (need > 0xFFFFFFF0u) return fail(B200SEED_ERR_OVERFLOW, "more than 2^32 candidate records in one batch: split it");
        h->recCapacity = need + need / 8 + 1024;
        int rc = enqueue(h);
        if (rc != B200SEED_OK) return rc;
        CUDA_TRY(cudaStreamSynchronize(s));
        continue;
      }
      const int rounds = h->confRoundsLaunched;
      int used = rounds;
      for (int r = 1; r < rounds; ++r) {
        if (h->hConfState[kConfChangedBase + r] == 0u) { used = r + 1; break; }
      }
      h->lastConfRounds = (uint32_t)used;
      if (h->hConfState[kConfChangedBase + rounds - 1] == 0u) break;  // a round reproduced its predecessor
      if (rounds + h->confRoundsPerBatch > kConfMaxRounds) {
        return fail(B200SEED_ERR_RUNTIME, "seedConfirmation fixed point not reached after " + std::to_string(rounds) + " rounds");
      }
      int rc = enqueue_conf_rounds(h, rounds, h->confRoundsPerBatch, s);
      if (rc == B200SEED_OK) rc = enqueue_tail(h, s);
      if (rc != B200SEED_OK) return rc;
      CUDA_TRY(cudaStreamSynchronize(s));
    }
  }
  h->pending = false;
  for (int i = 0; i < 4; ++i) {
    if (cudaEventElapsedTime(&h->stageMs[i], h->ev[i], h->ev[i + 1]) != cudaSuccess) h->stageMs[i] = 0.f;
  }
  {
    float t = 0.f;
    h->stageMs[4] = cudaEventElapsedTime(&t, h->ev[2], h->evCount) == cudaSuccess ? t : 0.f;
    h->stageMs[5] = 0.f;
    h->stageMs[6] = 0.f;
    const uint32_t lag = (h->chunkStreams > 1 && h->chunksTimed > 1) ? (uint32_t)h->chunkStreams : 1u;
    for (uint32_t c = 0; c < h->chunksTimed; ++c) {  // with two chunk streams the sums overlap in time
      cudaEvent_t before = c < lag ? h->evCount : h->evChunk[2 * (c - lag) + 1];
      if (cudaEventElapsedTime(&t, before, h->evChunk[2 * c]) == cudaSuccess) h->stageMs[5] += t;
      if (cudaEventElapsedTime(&t, h->evChunk[2 * c], h->evChunk[2 * c + 1]) == cudaSuccess) h->stageMs[6] += t;
    }
  }
  (void)cudaGetLastError();
  b200seed_counters& c = h->lastCounters;
  c.nSpacePoints = h->lastTotal;
  c.nInGrid = *reinterpret_cast<uint32_t*>(h->hCounters + kCntSlots);
  c.nMiddles = h->hCounters[kCntMiddles];
  c.nBottomDoublets = h->hCounters[kCntBottomDoublets];
  c.nTopDoublets = h->hCounters[kCntTopDoublets];
  c.nTripletTests = h->hCounters[kCntTripletTests];
  c.nCandidates = h->hCounters[kCntCandidates];
  c.nSeeds = *h->hSeedTotal;
  c.nTieMiddles = h->hCounters[kCntTieMiddles];
  c.nKernelLaunches = h->launches;
  c.nConfirmationRounds = h->plan.dev.seedConfirmation ? h->lastConfRounds : 0;
  if (out != nullptr) out->size = *h->hSeedTotal;
  const int st = *h->hStatus;
  if (st & (kStatusOverflowDoublets | kStatusOverflowPool)) {
    return fail(B200SEED_ERR_OVERFLOW,
                "a middle space point exceeds the engine's limits (more than " + std::to_string(kMaxListLength) +
                    " doublets on one side, or more triplet candidates than the spill class's pool holds)");
  }
  if (*h->hSeedTotal > h->lastCapacity) {
    return fail(B200SEED_ERR_CAPACITY, "seed buffers too small: need " + std::to_string(*h->hSeedTotal));
  }
  return B200SEED_OK;
}

}  // namespace

extern "C" {

const char* b200seed_last_error(void) { return g_lastError.c_str(); }

void* b200seed_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes > 0 ? bytes : 1) != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  return p;
}

void b200seed_free_pinned(void* p) {
  if (p != nullptr) cudaFreeHost(p);
}

int b200seed_config_init(b200seed_config* cfg) {
  if (cfg == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "cfg is NULL");
  config_defaults(*cfg);
  return B200SEED_OK;
}

// Host-only planning, usable without a GPU (validation + derived constants).
int b200seed_plan_info(const b200seed_config* cfg, b200seed_info* info) {
  if (cfg == nullptr || info == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  HostPlan plan;
  PlanError err;
  if (!make_host_plan(*cfg, plan, err)) return fail(err.code, err.message);
  *info = plan.info;
  return B200SEED_OK;
}

// Host-only: device constants + navigation / neighbour tables of a config
// (used by the GPU-less model tests).  Arrays may be NULL to query the sizes:
// sizes[0] = nNav, sizes[1] = #bottom bins, sizes[2] = #top bins,
// sizes[3] = sizeof(DeviceConfig), sizes[4] = seeds per middle.
int b200seed_plan_tables(const b200seed_config* cfg, void* deviceConfig, uint64_t deviceConfigBytes,
                         uint32_t* navBins, uint32_t* botOffsets, uint32_t* botBins,
                         uint32_t* topOffsets, uint32_t* topBins, uint64_t* sizes) {
  if (cfg == nullptr || sizes == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  HostPlan plan;
  PlanError err;
  if (!make_host_plan(*cfg, plan, err)) return fail(err.code, err.message);
  sizes[0] = plan.navBins.size();
  sizes[1] = plan.botBins.size();
  sizes[2] = plan.topBins.size();
  sizes[3] = sizeof(DeviceConfig);
  sizes[4] = plan.seedsPerMiddle;
  if (deviceConfig != nullptr) {
    if (deviceConfigBytes != sizeof(DeviceConfig)) return fail(B200SEED_ERR_INVALID_ARGUMENT, "DeviceConfig size mismatch");
    std::memcpy(deviceConfig, &plan.dev, sizeof(DeviceConfig));
  }
  auto copy = [](uint32_t* dst, const std::vector<uint32_t>& v) {
    if (dst != nullptr && !v.empty()) std::memcpy(dst, v.data(), v.size() * 4);
  };
  copy(navBins, plan.navBins);
  copy(botOffsets, plan.botOffsets);
  copy(botBins, plan.botBins);
  copy(topOffsets, plan.topOffsets);
  copy(topBins, plan.topBins);
  return B200SEED_OK;
}

static int create_impl(const b200seed_config* cfg, const b200seed_orthogonal_options* orthOpt, int device, b200seed_handle** out);

int b200seed_create(const b200seed_config* cfg, int device, b200seed_handle** out) {
  return create_impl(cfg, nullptr, device, out);
}

int b200seed_orthogonal_config_init(b200seed_config* cfg, b200seed_orthogonal_options* opt) {
  if (cfg == nullptr || opt == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  config_defaults(*cfg);  // the shared members have the grid algorithm's defaults (OrthogonalTripletSeedingAlgorithm.hpp:38-186)
  opt->zOutermostLayersMin = -2700.f;
  opt->zOutermostLayersMax = 2700.f;
  opt->deltaPhiMax = 0.085f;
  return B200SEED_OK;
}

int b200seed_create_orthogonal(const b200seed_config* cfg, const b200seed_orthogonal_options* opt, int device,
                               b200seed_handle** out) {
  if (opt == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  return create_impl(cfg, opt, device, out);
}

static int create_impl(const b200seed_config* cfg, const b200seed_orthogonal_options* orthOpt, int device, b200seed_handle** out) {
  if (cfg == nullptr || out == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = nullptr;
  auto* h = new b200seed_handle;
  PlanError err;
  if (!(orthOpt != nullptr ? make_orthogonal_plan(*cfg, *orthOpt, h->plan, err) : make_host_plan(*cfg, h->plan, err))) {
    delete h;
    return fail(err.code, err.message);
  }
#ifdef B200SEED_RELAXED
  const bool engineRelaxed = true;
#else
  const bool engineRelaxed = false;
#endif
  if (h->plan.relaxedFloat != engineRelaxed) {  // seeding_abi.cpp picks the engine by cfg->relaxedFloat
    delete h;
    return fail(B200SEED_ERR_INVALID_ARGUMENT, "relaxedFloat does not match the engine this handle was routed to");
  }
  int nDev = 0;
  cudaError_t ce = cudaGetDeviceCount(&nDev);
  if (ce != cudaSuccess || nDev <= 0) {
    delete h;
    return fail(B200SEED_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(ce));
  }
  if (device < 0 || device >= nDev) {
    delete h;
    return fail(B200SEED_ERR_INVALID_ARGUMENT, "device index out of range");
  }
  h->device = device;
  auto cleanup = [&](int code) {
    b200seed_destroy(h);
    return code;
  };
#define CREATE_TRY(expr)                                                                       \
  do {                                                                                         \
    cudaError_t err__ = (expr);                                                                \
    if (err__ != cudaSuccess) {                                                                \
      return cleanup(fail(B200SEED_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__))); \
    }                                                                                          \
  } while (0)
  CREATE_TRY(cudaSetDevice(device));
  cudaDeviceProp prop{};
  CREATE_TRY(cudaGetDeviceProperties(&prop, device));
  h->smCount = prop.multiProcessorCount;
  h->ccMajor = prop.major;
  h->ccMinor = prop.minor;
  h->plan.info.smCount = h->smCount;
  h->plan.info.ccMajor = h->ccMajor;
  h->plan.info.ccMinor = h->ccMinor;
  CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 5; ++i) CREATE_TRY(cudaEventCreate(&h->ev[i]));
  CREATE_TRY(cudaMallocHost(&h->hCounters, (kCntSlots + 2) * 8));
  CREATE_TRY(cudaMallocHost(&h->hStatus, 16));
  CREATE_TRY(cudaMallocHost(&h->hSeedTotal, 16));

  h->confRoundsPerBatch = (int)std::min<uint32_t>(std::max<uint32_t>(env_u32("B200SEED_CONF_ROUNDS", kConfRoundsPerBatch), 2u), 64u);
  h->recPerSpacePoint = std::max<uint32_t>(env_u32("B200SEED_REC_PER_SP", 32), 1u);
  // the tie-order replay of the unstable sorts only makes sense when the keys are the reference's bit for bit
  h->exactTies = engineRelaxed ? 0 : (int)env_u32("B200SEED_EXACT_TIES", 1);
  h->arenaMaxBytes = (size_t)std::max<uint32_t>(env_u32("B200SEED_ARENA_MB", 8192), 64u) << 20;
  CREATE_TRY(cudaEventCreate(&h->evCount));
  CREATE_TRY(cudaEventCreateWithFlags(&h->evPlan, cudaEventDisableTiming));
  h->chunkStreams = env_u32("B200SEED_CHUNK_STREAMS", 2) >= 2 ? 2 : 1;
  h->kdHostBuild = env_u32("B200SEED_KD_HOST", 0) != 0;
  h->kdHitMB = env_u32("B200SEED_KD_HIT_MB", 4096);
  h->maskWordsPerSp = env_u32("B200SEED_MASK_WORDS_PER_SP", 192);
  h->classStreams = env_u32("B200SEED_CLASS_STREAMS", 1) != 0 ? 1 : 0;
  for (int a = 0; a < 2; ++a) {
    CREATE_TRY(cudaEventCreateWithFlags(&h->evFill[a], cudaEventDisableTiming));
    for (int k = 0; k < kNumSeedClasses; ++k) {
      CREATE_TRY(cudaStreamCreateWithFlags(&h->classStream[a][k], cudaStreamNonBlocking));
      CREATE_TRY(cudaEventCreateWithFlags(&h->evClass[a][k], cudaEventDisableTiming));
    }
  }
  for (int a = 0; a < 2; ++a) {
    CREATE_TRY(cudaStreamCreateWithFlags(&h->chunkStream[a], cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreateWithFlags(&h->evChunkEnd[a], cudaEventDisableTiming));
  }
  CREATE_TRY(cudaMallocHost(&h->hPlan, (8 + (size_t)kMaxChunks + 1 + 8) * 4));
  {
    // dynamic shared memory of every class: what is left of the SM for N resident blocks (1 KB per block is
    // reserved by the system, the kernel's static shared memory comes on top)
    const bool conf = h->plan.dev.seedConfirmation != 0;
    if (conf) CREATE_TRY(cudaMallocHost(&h->hConfState, kConfStateWords * 4));
    // (with seedConfirmation the collector lives in k_conf_replay, not in the seeding kernel)
    h->heapBytes = (!conf && h->plan.dev.maxSeedsPerSpMConf > (uint32_t)kMaxHeap) ? ((16u * h->plan.dev.maxSeedsPerSpMConf + 127u) & ~127u) : 0u;
    for (int c = 0; c < kNumSeedClasses; ++c) {
      SeedKernel k = seed_kernel(conf, c);
      cudaFuncAttributes fa{};
      CREATE_TRY(cudaFuncGetAttributes(&fa, k));
      uint32_t bytes = 0;
      if (c != kSpillClass) {
        const size_t perBlock = (size_t)prop.sharedMemPerMultiprocessor / (size_t)kSeedClassShape[c].blocksPerSM;
        size_t dyn = perBlock - 1024 - fa.sharedSizeBytes;
        dyn = std::min<size_t>(dyn, (size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes);
        bytes = (uint32_t)(dyn & ~(size_t)127) - h->heapBytes;
        // The classes share one kernel function and the attribute is per function and process-wide: every
        // handle sets the SAME value (the opt-in maximum), so a handle created while another thread launches
        // the largest class can never lower the limit under that launch.
        CREATE_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(((size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes) & ~(size_t)127)));
        if (!h->plan.orthogonal) {  // the strip variant of the class (same static shared memory, same limit)
          CREATE_TRY(cudaFuncSetAttribute(seed_kernel(conf, c, true), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(((size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes) & ~(size_t)127)));
        }
      }
      h->classBytes[c] = bytes;
      h->classThreads[c] = kSeedClassShape[c].threads;
      if (const char* v = std::getenv("B200SEED_CLASS_THREADS")) {  // kernel experiments: "160,256,320,512,1024,1024"
        std::vector<int> t;
        for (const char* q = v; *q != 0;) {
          t.push_back(std::atoi(q));
          const char* comma = std::strchr(q, ',');
          if (comma == nullptr) break;
          q = comma + 1;
        }
        if ((int)t.size() == kNumSeedClasses && t[c] >= 32 && t[c] <= 1024 && t[c] % 32 == 0) h->classThreads[c] = t[c];
      }
      int b = 0;
      CREATE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k, h->classThreads[c], bytes + h->heapBytes));
      h->classBlocksPerSM[c] = std::max(1, b);
    }
    int b = 0;
    CREATE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_doublets<false>, kDoubletWarps * 32, 0));
    h->doubletBlocksPerSM[0] = std::max(1, b);
    CREATE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_doublets<true>, kDoubletWarps * 32, 0));
    h->doubletBlocksPerSM[1] = std::max(1, b);
    CREATE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_doublets_kd<false>, kKdThreads, 0));
    h->kdBlocksPerSM[0] = std::max(1, b);
    CREATE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_doublets_kd<true>, kKdThreads, 0));
    h->kdBlocksPerSM[1] = std::max(1, b);
  }
  h->sortSmemCap = std::min<uint32_t>(std::max<uint32_t>(env_u32("B200SEED_SORT_CAP", 2560), 520), 8192);  // (>= 513: a larger bin's 32 n bytes of scratch also hold the bucket counters)
  CREATE_TRY(cudaFuncSetAttribute(k_sort_bins, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8192 * 16 + (kSortBuckets + 1) * 4)));
  if (std::getenv("B200SEED_VERBOSE") != nullptr) {
    for (int c = 0; c < kNumSeedClasses; ++c) {
      std::fprintf(stderr, "b200seed: seed class %d: %d threads, %u bytes of shared memory, %d blocks per SM\n", c,
                   h->classThreads[c], h->classBytes[c], h->classBlocksPerSM[c]);
    }
    std::fprintf(stderr, "b200seed: doublet kernels: %d / %d blocks per SM (count / fill)\n", h->doubletBlocksPerSM[0],
                 h->doubletBlocksPerSM[1]);
  }

  int rc = upload(h->navBins, h->plan.navBins, h->stream);
  if (rc == B200SEED_OK) rc = upload(h->botOffsets, h->plan.botOffsets, h->stream);
  if (rc == B200SEED_OK) rc = upload(h->botBins, h->plan.botBins, h->stream);
  if (rc == B200SEED_OK) rc = upload(h->topOffsets, h->plan.topOffsets, h->stream);
  if (rc == B200SEED_OK) rc = upload(h->topBins, h->plan.topBins, h->stream);
  if (rc != B200SEED_OK) return cleanup(rc);
  CREATE_TRY(cudaStreamSynchronize(h->stream));
#undef CREATE_TRY
  *out = h;
  return B200SEED_OK;
}

void b200seed_destroy(b200seed_handle* h) {
  if (h == nullptr) return;
  cudaSetDevice(h->device);
  if (h->stream != nullptr) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  for (DevBuf* b : {&h->navBins, &h->botOffsets, &h->botBins, &h->topOffsets, &h->topBins, &h->inOffsets,
                    &h->inX, &h->inY, &h->inZ, &h->inR, &h->inVarZ, &h->inVarR, &h->binOf, &h->binCount,
                    &h->binStart, &h->binCursor, &h->tmpIdx, &h->pIdx, &h->pXY, &h->pZR, &h->pVar,
                    &h->sortScratch, &h->midLo, &h->midCount, &h->workStart, &h->workPos, &h->workEG,
                    &h->workCounter, &h->capB, &h->capT, &h->slotPrefix, &h->capTileSums, &h->capTilePrefix, &h->planDev,
                    &h->hdr, &h->carve, &h->classList, &h->maskArena, &h->maskOff, &h->inStrip, &h->pStrip, &h->kdHitArena, &h->kdHitHead, &h->arenaRec[0], &h->arenaRec[1], &h->arenaKey[0], &h->arenaKey[1], &h->spillScratch,
                    &h->zWinOffsets, &h->slotB, &h->slotM, &h->slotT, &h->slotQ, &h->slotZ, &h->slotCount,
                    &h->seedStart, &h->tileSums, &h->tilePrefix, &h->outB, &h->outM, &h->outT, &h->outQ,
                    &h->outZ, &h->seedOffsets, &h->counters, &h->status, &h->zWin, &h->rec, &h->recZ, &h->recBegin,
                    &h->recCount, &h->slot2B, &h->slot2M, &h->slot2T, &h->slot2Q, &h->slot2Z, &h->slot2Count,
                    &h->confHead, &h->confNext, &h->confState, &h->confDirty, &h->orthPosOrig, &h->orthPosPhi,
                    &h->orthNodes, &h->orthCoreOffsets, &h->orthNodeOffsets, &h->orthRRange, &h->orthElemR, &h->orthElemZ,
                    &h->orthScratch, &h->orthTasks, &h->orthExtent}) {
    b->release();
  }
  if (h->hConfState != nullptr) cudaFreeHost(h->hConfState);
  if (h->hPlan != nullptr) cudaFreeHost(h->hPlan);
  if (h->evCount != nullptr) cudaEventDestroy(h->evCount);
  if (h->evPlan != nullptr) cudaEventDestroy(h->evPlan);
  for (int a = 0; a < 2; ++a) {
    if (h->chunkStream[a] != nullptr) { cudaStreamSynchronize(h->chunkStream[a]); cudaStreamDestroy(h->chunkStream[a]); }
    if (h->evChunkEnd[a] != nullptr) cudaEventDestroy(h->evChunkEnd[a]);
    if (h->evFill[a] != nullptr) cudaEventDestroy(h->evFill[a]);
    for (int k = 0; k < kNumSeedClasses; ++k) {
      if (h->classStream[a][k] != nullptr) cudaStreamDestroy(h->classStream[a][k]);
      if (h->evClass[a][k] != nullptr) cudaEventDestroy(h->evClass[a][k]);
    }
  }
  for (cudaEvent_t e : h->evChunk) cudaEventDestroy(e);
  for (int i = 0; i < 5; ++i) {
    if (h->ev[i] != nullptr) cudaEventDestroy(h->ev[i]);
  }
  if (h->hCounters != nullptr) cudaFreeHost(h->hCounters);
  if (h->hStatus != nullptr) cudaFreeHost(h->hStatus);
  if (h->hSeedTotal != nullptr) cudaFreeHost(h->hSeedTotal);
  delete h;
}

int b200seed_get_info(const b200seed_handle* h, b200seed_info* info) {
  if (h == nullptr || info == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  *info = h->plan.info;
  return B200SEED_OK;
}

int b200seed_get_counters(const b200seed_handle* h, b200seed_counters* c) {
  if (h == nullptr || c == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  *c = h->lastCounters;
  return B200SEED_OK;
}

// Restrict the following calls on this handle to the middle space points whose
// phi bin lies in [firstPhiBin, firstPhiBin + nPhiBins) (1-based local bins;
// nPhiBins = 0 restores "all").  Used to split ONE event over several GPUs:
// every GPU builds the full grid, seeds its own sector (neighbour bins are read
// across the sector border like everywhere else) and the per-sector seed lists,
// concatenated in sector order, are the reference's output (phi is the
// outermost navigation axis, GridIterator.ipp:228-242).
int b200seed_set_phi_sector(b200seed_handle* h, uint32_t firstPhiBin, uint32_t nPhiBins) {
  if (h == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL handle");
  if (h->plan.orthogonal) return fail(B200SEED_ERR_UNSUPPORTED, "not available on an orthogonal-seeder handle (no grid, no vertex windows)");
  if (nPhiBins != 0 && h->plan.dev.seedConfirmation) {
    return fail(B200SEED_ERR_UNSUPPORTED,
                "seedConfirmation couples the middles of an event through bestSeedQualityMap: no phi-sector split");
  }
  if (nPhiBins == 0 && firstPhiBin == B200SEED_PHI_SECTOR_EMPTY) {  // a rank without bins (more ranks than phi bins): no middles at all
    h->phiFirst = 0xFFFFFFFFu;
    h->phiCount = 0;
    return B200SEED_OK;
  }
  if (nPhiBins == 0) {
    h->phiFirst = 1;
    h->phiCount = 0xFFFFFFFFu;
    return B200SEED_OK;
  }
  if (firstPhiBin < 1 || firstPhiBin > (uint32_t)h->plan.dev.phiBins) {
    return fail(B200SEED_ERR_INVALID_ARGUMENT, "phi sector outside [1, phiBins]");
  }
  h->phiFirst = firstPhiBin;
  h->phiCount = nPhiBins;
  return B200SEED_OK;
}

// GPU time of the four stages of the last completed call, in milliseconds:
// [0] grid (bin, scan, scatter, sort + packed copy), [1] middle work list,
// [2] seeding kernel (both capacity tiers), [3] ordered seed compaction.
int b200seed_get_stage_times(const b200seed_handle* h, float* ms) {
  if (h == nullptr || ms == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  for (int i = 0; i < 4; ++i) ms[i] = h->stageMs[i];
  return B200SEED_OK;
}

int b200seed_get_stage_times_ex(const b200seed_handle* h, float* ms, uint32_t n) {
  if (h == nullptr || ms == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  for (uint32_t i = 0; i < n && i < 7; ++i) ms[i] = h->stageMs[i];
  return B200SEED_OK;
}

int b200seed_run_batch_device(b200seed_handle* h, uint32_t nEvents, uint32_t nSpacePointsTotal,
                              const uint32_t* spOffsets, const float* x, const float* y, const float* z,
                              const float* r, const float* varZ, const float* varR, uint64_t* seedOffsets,
                              b200seed_seeds* out, void* cudaStream) {
  if (h == nullptr || out == nullptr || spOffsets == nullptr || seedOffsets == nullptr || nEvents == 0) {
    return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument or empty batch");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = cudaStream != nullptr ? static_cast<cudaStream_t>(cudaStream) : h->stream;
  if (h->pending && h->plan.dev.seedConfirmation) {
    // the convergence of the seedConfirmation rounds (and the size of the record pool) is a host decision taken at
    // the sync: a previous asynchronous call is completed first, so that it can never pass unchecked
    int rc = finish(h, h->last.stream != nullptr ? h->last.stream : h->stream, nullptr);
    if (rc != B200SEED_OK) return rc;
  }
  h->lastSeedOffsets = reinterpret_cast<const unsigned long long*>(seedOffsets);
  b200seed_handle::EnqueueArgs& a = h->last;
  a = b200seed_handle::EnqueueArgs{};
  a.nEvents = nEvents; a.nTotal = nSpacePointsTotal; a.dOffsets = spOffsets;
  a.x = x; a.y = y; a.z = z; a.r = r; a.varZ = varZ; a.varR = varR;
  a.outB = static_cast<uint32_t*>(out->bottom); a.outM = static_cast<uint32_t*>(out->middle);
  a.outT = static_cast<uint32_t*>(out->top);
  a.outQ = static_cast<float*>(out->quality); a.outZ = static_cast<float*>(out->vertexZ);
  a.outCapacity = out->capacity;
  a.dSeedOffsets = reinterpret_cast<unsigned long long*>(seedOffsets);
  a.stream = s;
  a.vertexCuts = h->plan.useVertexZCuts;  // no windows on this entry: every doublet passes the vertex cut
  return enqueue(h);
}

int b200seed_sync(b200seed_handle* h, b200seed_seeds* out) {
  if (h == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL handle");
  CUDA_TRY(cudaSetDevice(h->device));
  // the stream of the last call is ordered with the handle's stream through
  // the device-wide synchronisation below (the caller may have used its own)
  CUDA_TRY(cudaDeviceSynchronize());
  return finish(h, h->last.stream != nullptr ? h->last.stream : h->stream, out);
}

// measurements of one event as the source of the space points (b200seed_run_measurements)
struct MeasurementSource {
  const uint32_t* surface;
  const double* col[5];  // loc0, loc1, cov00, cov01, cov11
  uint32_t nSurfaces;
  const double* transforms;
  float* spOut[6];  // optional host copies of the space point columns
};

// VertexZCuts windows of a call: `offsets` == NULL: the nZWin windows apply to every event of the batch,
// else event e owns windows [offsets[e], offsets[e + 1]).
struct WindowSource {
  uint32_t nZWin = 0;
  const float* lo = nullptr;
  const float* hi = nullptr;
  const uint32_t* offsets = nullptr;
};

// The windows of one event merged into disjoint intervals in ascending order: the union -- and with it the
// answer of VertexZCuts::operator() (.cpp:78-96), "inside any window" -- is unchanged, and the device can
// find the deciding interval by binary search however many vertices there are.
static void merge_windows(const float* lo, const float* hi, uint32_t n, std::vector<float>& outLo, std::vector<float>& outHi,
                          size_t first) {
  std::vector<std::pair<float, float>> w;
  w.reserve(n);
  for (uint32_t i = 0; i < n; ++i) {
    if (lo[i] <= hi[i]) w.emplace_back(lo[i], hi[i]);  // an empty (or NaN) window contains nothing
  }
  std::sort(w.begin(), w.end());
  for (const auto& [a, b] : w) {
    if (outLo.size() > first && a <= outHi.back()) {  // overlaps (or touches) the current interval
      if (b > outHi.back()) outHi.back() = b;
    } else {
      outLo.push_back(a);
      outHi.push_back(b);
    }
  }
}

static int run_host_batch(b200seed_handle* h, uint32_t nEvents, const uint32_t* spOffsets, const float* x,
                          const float* y, const float* z, const float* r, const float* varZ,
                          const float* varR, const float* phi, const WindowSource& win, uint64_t* seedOffsets,
                          b200seed_seeds* out, const MeasurementSource* meas = nullptr) {
  if (h == nullptr || out == nullptr || spOffsets == nullptr || nEvents == 0) {
    return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument or empty batch");
  }
  const uint32_t nZWin = win.nZWin;
  if (nZWin > 0 && (win.lo == nullptr || win.hi == nullptr)) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL z window column");
  if (h->plan.orthogonal && (nZWin > 0 || win.offsets != nullptr || phi != nullptr)) {
    return fail(B200SEED_ERR_UNSUPPORTED, "vertex z windows / a caller-supplied phi column are not part of the orthogonal seeder");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const uint32_t nTotal = spOffsets[nEvents] - spOffsets[0];
  if (spOffsets[0] != 0) return fail(B200SEED_ERR_INVALID_ARGUMENT, "spOffsets[0] must be 0");
  const size_t colBytes = std::max<size_t>(4, (size_t)nTotal * 4);
  CUDA_TRY(h->inOffsets.reserve(((size_t)nEvents + 1) * 4));
  DevBuf* cols[6] = {&h->inX, &h->inY, &h->inZ, &h->inR, &h->inVarZ, &h->inVarR};
  const float* src[6] = {x, y, z, r, varZ, varR};
  DevBuf measBuf[7], measStatus;
  struct Guard {
    std::vector<DevBuf*> bufs;
    ~Guard() { for (DevBuf* b : bufs) b->release(); }
  } measGuard;
  for (int i = 0; i < 6; ++i) {
    CUDA_TRY(cols[i]->reserve(colBytes));
    if (nTotal > 0 && meas == nullptr) {
      if (src[i] == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL space point column");
      CUDA_TRY(cudaMemcpyAsync(cols[i]->ptr, src[i], (size_t)nTotal * 4, cudaMemcpyHostToDevice, s));
    }
  }
  if (meas != nullptr && nTotal > 0) {
    // the space points are made on the device, straight into the input columns of the pipeline
    for (DevBuf& b : measBuf) measGuard.bufs.push_back(&b);
    measGuard.bufs.push_back(&measStatus);
    CUDA_TRY(measBuf[5].reserve((size_t)nTotal * 4));
    CUDA_TRY(cudaMemcpyAsync(measBuf[5].ptr, meas->surface, (size_t)nTotal * 4, cudaMemcpyHostToDevice, s));
    for (int k = 0; k < 5; ++k) {
      CUDA_TRY(measBuf[k].reserve((size_t)nTotal * 8));
      CUDA_TRY(cudaMemcpyAsync(measBuf[k].ptr, meas->col[k], (size_t)nTotal * 8, cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(measBuf[6].reserve((size_t)meas->nSurfaces * 96));
    CUDA_TRY(cudaMemcpyAsync(measBuf[6].ptr, meas->transforms, (size_t)meas->nSurfaces * 96, cudaMemcpyHostToDevice, s));
    CUDA_TRY(measStatus.reserve(16));
    CUDA_TRY(cudaMemsetAsync(measStatus.ptr, 0, 16, s));
    SpacePointMakerParams mp{};
    mp.surface = measBuf[5].as<uint32_t>();
    mp.loc0 = measBuf[0].as<double>(); mp.loc1 = measBuf[1].as<double>();
    mp.cov00 = measBuf[2].as<double>(); mp.cov01 = measBuf[3].as<double>(); mp.cov11 = measBuf[4].as<double>();
    mp.transforms = measBuf[6].as<double>();
    mp.x = h->inX.as<float>(); mp.y = h->inY.as<float>(); mp.z = h->inZ.as<float>(); mp.r = h->inR.as<float>();
    mp.varZ = h->inVarZ.as<float>(); mp.varR = h->inVarR.as<float>();
    mp.n = nTotal; mp.nSurfaces = meas->nSurfaces;
    mp.status = measStatus.as<int>();
    const int blocks = (int)std::min<uint64_t>(((uint64_t)nTotal + 255) / 256, (uint64_t)h->smCount * 8);
    k_pixel_spacepoints<<<blocks, 256, 0, s>>>(mp);
    CUDA_TRY(cudaGetLastError());
    int st = 0;
    CUDA_TRY(cudaMemcpyAsync(&st, measStatus.ptr, 4, cudaMemcpyDeviceToHost, s));
    for (int i = 0; i < 6; ++i) {
      if (meas->spOut[i] != nullptr) {
        CUDA_TRY(cudaMemcpyAsync(meas->spOut[i], cols[i]->ptr, (size_t)nTotal * 4, cudaMemcpyDeviceToHost, s));
      }
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    if (st != 0) return fail(B200SEED_ERR_INVALID_ARGUMENT, "surface index out of range");
  }
  CUDA_TRY(cudaMemcpyAsync(h->inOffsets.ptr, spOffsets, ((size_t)nEvents + 1) * 4, cudaMemcpyHostToDevice, s));
  float* dPhi = nullptr;
  if (phi != nullptr && nTotal > 0) {
    CUDA_TRY(h->binOf.reserve(colBytes));  // make sure workspaces exist before aliasing
    CUDA_TRY(h->tmpIdx.reserve(colBytes));
    // phi staging shares no workspace with the pipeline: use sortScratch's tail
    CUDA_TRY(h->sortScratch.reserve((size_t)nTotal * 32 + colBytes));
    dPhi = reinterpret_cast<float*>(h->sortScratch.as<unsigned char>() + (size_t)nTotal * 32);
    CUDA_TRY(cudaMemcpyAsync(dPhi, phi, (size_t)nTotal * 4, cudaMemcpyHostToDevice, s));
  }
  int rc = ensure_workspace(h, nEvents, nTotal);
  if (rc != B200SEED_OK) return rc;
  // vertex z windows: merged per event, uploaded as {lo column, hi column} (+ per-event offsets)
  uint32_t nMerged = 0;
  const uint32_t* dWinOffsets = nullptr;
  {
    std::vector<float> mLo, mHi;
    std::vector<uint32_t> mOff;
    if (win.offsets != nullptr) {
      mOff.push_back(0);
      for (uint32_t e = 0; e < nEvents; ++e) {
        const uint32_t a = win.offsets[e], b = win.offsets[e + 1];
        if (b < a || b > nZWin) return fail(B200SEED_ERR_INVALID_ARGUMENT, "z window offsets are not ascending / exceed the window count");
        merge_windows(win.lo + a, win.hi + a, b - a, mLo, mHi, mLo.size());
        mOff.push_back((uint32_t)mLo.size());
      }
    } else {
      merge_windows(win.lo, win.hi, nZWin, mLo, mHi, 0);
    }
    nMerged = (uint32_t)mLo.size();
    h->zWinCapacity = std::max<uint32_t>(nMerged, 1);
    CUDA_TRY(h->zWin.reserve((size_t)h->zWinCapacity * 8));
    if (nMerged > 0) {
      CUDA_TRY(cudaMemcpyAsync(h->zWin.ptr, mLo.data(), (size_t)nMerged * 4, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemcpyAsync(h->zWin.as<float>() + h->zWinCapacity, mHi.data(), (size_t)nMerged * 4, cudaMemcpyHostToDevice, s));
    }
    if (!mOff.empty()) {
      CUDA_TRY(h->zWinOffsets.reserve(mOff.size() * 4));
      CUDA_TRY(cudaMemcpyAsync(h->zWinOffsets.ptr, mOff.data(), mOff.size() * 4, cudaMemcpyHostToDevice, s));
      dWinOffsets = h->zWinOffsets.as<uint32_t>();
    }
    CUDA_TRY(cudaStreamSynchronize(s));  // the staging vectors go out of scope
  }
  const size_t K = std::max<uint32_t>(h->plan.seedsPerMiddle, 1);
  const size_t maxSeeds = std::max<size_t>(1, (size_t)nTotal * K * (h->plan.orthogonal ? 2 : 1));
  CUDA_TRY(h->outB.reserve(maxSeeds * 4));
  CUDA_TRY(h->outM.reserve(maxSeeds * 4));
  CUDA_TRY(h->outT.reserve(maxSeeds * 4));
  CUDA_TRY(h->outQ.reserve(maxSeeds * 4));
  CUDA_TRY(h->outZ.reserve(maxSeeds * 4));
  CUDA_TRY(h->seedOffsets.reserve(((size_t)nEvents + 1) * 8));
  {
    b200seed_handle::EnqueueArgs& a = h->last;
    a = b200seed_handle::EnqueueArgs{};
    a.nEvents = nEvents; a.nTotal = nTotal; a.dOffsets = h->inOffsets.as<uint32_t>();
    a.x = h->inX.as<float>(); a.y = h->inY.as<float>(); a.z = h->inZ.as<float>(); a.r = h->inR.as<float>();
    a.varZ = h->inVarZ.as<float>(); a.varR = h->inVarR.as<float>(); a.dPhi = dPhi;
    if (meas == nullptr) { a.hx = x; a.hy = y; a.hz = z; a.hr = r; a.hOffsets = spOffsets; }
    a.nZWin = (int)nMerged;
    a.dZWinOffsets = dWinOffsets;
    a.vertexCuts = h->plan.useVertexZCuts || nZWin > 0 || win.offsets != nullptr;
    a.dStrip = h->pendingStrip;
    a.cotThetaDiffMax = h->pendingCotThetaDiffMax;
    a.outB = h->outB.as<uint32_t>(); a.outM = h->outM.as<uint32_t>(); a.outT = h->outT.as<uint32_t>();
    a.outQ = h->outQ.as<float>(); a.outZ = h->outZ.as<float>();
    a.outCapacity = maxSeeds;
    a.dSeedOffsets = h->seedOffsets.as<unsigned long long>();
    a.stream = s;
  }
  rc = enqueue(h);
  if (rc != B200SEED_OK) return rc;
  b200seed_seeds tmp{};
  rc = finish(h, s, &tmp);
  out->size = tmp.size;
  if (rc != B200SEED_OK) return rc;
  if (tmp.size > out->capacity) {
    return fail(B200SEED_ERR_CAPACITY, "seed buffers too small: need " + std::to_string(tmp.size));
  }
  if (tmp.size > 0) {
    CUDA_TRY(cudaMemcpyAsync(out->bottom, h->outB.ptr, tmp.size * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out->middle, h->outM.ptr, tmp.size * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out->top, h->outT.ptr, tmp.size * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out->quality, h->outQ.ptr, tmp.size * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out->vertexZ, h->outZ.ptr, tmp.size * 4, cudaMemcpyDeviceToHost, s));
  }
  if (seedOffsets != nullptr) {
    CUDA_TRY(cudaMemcpyAsync(seedOffsets, h->seedOffsets.ptr, ((size_t)nEvents + 1) * 8, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return B200SEED_OK;
}

int b200seed_run_measurements(b200seed_handle* h, uint32_t n, const uint32_t* surface, const double* loc0,
                              const double* loc1, const double* cov00, const double* cov01, const double* cov11,
                              uint32_t nSurfaces, const double* transforms, uint32_t nZWindows, const float* zWindowLo,
                              const float* zWindowHi, float* x, float* y, float* z, float* r, float* varZ, float* varR,
                              b200seed_seeds* out) {
  if (n > 0 && (surface == nullptr || loc0 == nullptr || loc1 == nullptr || cov00 == nullptr || cov01 == nullptr ||
                cov11 == nullptr || transforms == nullptr || nSurfaces == 0)) {
    return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL measurement column or empty surface table");
  }
  MeasurementSource ms{surface, {loc0, loc1, cov00, cov01, cov11}, nSurfaces, transforms, {x, y, z, r, varZ, varR}};
  const uint32_t offsets[2] = {0, n};
  WindowSource win;
  win.nZWin = nZWindows; win.lo = zWindowLo; win.hi = zWindowHi;
  return run_host_batch(h, 1, offsets, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, win, nullptr, out, &ms);
}

int b200seed_run(b200seed_handle* h, uint32_t nSpacePoints, const float* x, const float* y, const float* z,
                 const float* r, const float* varZ, const float* varR, uint32_t nZWindows,
                 const float* zWindowLo, const float* zWindowHi, b200seed_seeds* out) {
  const uint32_t offsets[2] = {0, nSpacePoints};
  WindowSource win;
  win.nZWin = nZWindows; win.lo = zWindowLo; win.hi = zWindowHi;
  return run_host_batch(h, 1, offsets, x, y, z, r, varZ, varR, nullptr, win, nullptr, out);
}

// One event through the strip triplet path (TripletSeedFinder::Config::useStripInfo = true,
// TripletSeedFinder.cpp:164-406): the raw calibration details are staged on the device, the grid stage gathers the
// derived details into packed order (k_gather_strips) and the seeding kernel runs its strip variant.
int b200seed_run_strips(b200seed_handle* h, uint32_t nSpacePoints, const float* x, const float* y, const float* z,
                        const float* r, const float* varZ, const float* varR, const float* stripDetails,
                        float cotThetaDiffMax, b200seed_seeds* out) {
  if (h == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL handle");
  if (h->plan.orthogonal) return fail(B200SEED_ERR_UNSUPPORTED, "strip triplet path: grid handles only");
  if (nSpacePoints > 0 && stripDetails == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL strip details");
  if (cotThetaDiffMax != cotThetaDiffMax) return fail(B200SEED_ERR_INVALID_ARGUMENT, "cotThetaDiffMax is NaN");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(h->inStrip.reserve(std::max<size_t>((size_t)nSpacePoints * 48, 16)));
  if (nSpacePoints > 0) {
    CUDA_TRY(cudaMemcpyAsync(h->inStrip.ptr, stripDetails, (size_t)nSpacePoints * 48, cudaMemcpyHostToDevice, h->stream));
  }
  const uint32_t offsets[2] = {0, nSpacePoints};
  WindowSource win;
  h->pendingStrip = h->inStrip.as<float>();
  h->pendingCotThetaDiffMax = cotThetaDiffMax;
  const int rc = run_host_batch(h, 1, offsets, x, y, z, r, varZ, varR, nullptr, win, nullptr, out);
  h->pendingStrip = nullptr;
  return rc;
}

// GridTripletSeedingAlgorithm.cpp:187-206: one z window per vertex, [z - half, z + half] with
// half = vertexZNSigma * sqrt(cov(2, 2)) + vertexZMargin, evaluated in double and narrowed to float.
int b200seed_vertex_windows(const b200seed_handle* h, uint32_t nVertices, const double* vertexZ, const double* vertexVarZ,
                            float* windowLo, float* windowHi) {
  if (h == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL handle");
  if (nVertices > 0 && (vertexZ == nullptr || vertexVarZ == nullptr || windowLo == nullptr || windowHi == nullptr)) {
    return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL vertex column");
  }
  for (uint32_t i = 0; i < nVertices; ++i) {
    const double zv = vertexZ[i];
    const double sigmaZ = std::sqrt(vertexVarZ[i]);
    const double half = h->plan.vertexZNSigma * sigmaZ + h->plan.vertexZMargin;
    windowLo[i] = static_cast<float>(zv - half);
    windowHi[i] = static_cast<float>(zv + half);
  }
  return B200SEED_OK;
}

// One event with its reconstructed vertices (Config::inputVertices): the windows are built like the reference does.
int b200seed_run_vertices(b200seed_handle* h, uint32_t nSpacePoints, const float* x, const float* y, const float* z,
                          const float* r, const float* varZ, const float* varR, uint32_t nVertices,
                          const double* vertexZ, const double* vertexVarZ, b200seed_seeds* out) {
  if (h == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL handle");
  if (!h->plan.useVertexZCuts) {
    return fail(B200SEED_ERR_INVALID_ARGUMENT, "b200seed_run_vertices needs a handle created with useVertexZCuts = 1 (inputVertices configured)");
  }
  std::vector<float> lo(nVertices), hi(nVertices);
  int rc = b200seed_vertex_windows(h, nVertices, vertexZ, vertexVarZ, lo.data(), hi.data());
  if (rc != B200SEED_OK) return rc;
  const uint32_t offsets[2] = {0, nSpacePoints};
  const uint32_t winOffsets[2] = {0, nVertices};
  WindowSource win;
  win.nZWin = nVertices; win.lo = lo.data(); win.hi = hi.data(); win.offsets = winOffsets;
  return run_host_batch(h, 1, offsets, x, y, z, r, varZ, varR, nullptr, win, nullptr, out);
}

// A batch of events, each with its own z windows: event e owns windows [windowOffsets[e], windowOffsets[e + 1]).
int b200seed_run_batch_windows(b200seed_handle* h, uint32_t nEvents, const uint32_t* spOffsets, const float* x,
                               const float* y, const float* z, const float* r, const float* varZ, const float* varR,
                               const uint32_t* windowOffsets, const float* zWindowLo, const float* zWindowHi,
                               uint64_t* seedOffsets, b200seed_seeds* out) {
  if (windowOffsets == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL window offsets");
  WindowSource win;
  win.nZWin = windowOffsets[nEvents]; win.lo = zWindowLo; win.hi = zWindowHi; win.offsets = windowOffsets;
  return run_host_batch(h, nEvents, spOffsets, x, y, z, r, varZ, varR, nullptr, win, seedOffsets, out);
}

// b200seed_run with a caller-provided phi column (optional precomputed
// azimuth, SURVEY.md section 7 "atan2f parity"; used by tests to separate the
// atan2f replay from the rest of the path).
int b200seed_run_with_phi(b200seed_handle* h, uint32_t nSpacePoints, const float* x, const float* y,
                          const float* z, const float* r, const float* varZ, const float* varR,
                          const float* phi, b200seed_seeds* out) {
  const uint32_t offsets[2] = {0, nSpacePoints};
  return run_host_batch(h, 1, offsets, x, y, z, r, varZ, varR, phi, WindowSource{}, nullptr, out);
}

int b200seed_run_batch(b200seed_handle* h, uint32_t nEvents, const uint32_t* spOffsets, const float* x,
                       const float* y, const float* z, const float* r, const float* varZ, const float* varR,
                       uint64_t* seedOffsets, b200seed_seeds* out) {
  return run_host_batch(h, nEvents, spOffsets, x, y, z, r, varZ, varR, nullptr, WindowSource{}, seedOffsets, out);
}

int b200seed_debug_grid(b200seed_handle* h, uint64_t capacity, uint32_t* copiedFromIndex, float* x, float* y,
                        float* z, float* r, float* varZ, float* varR, uint64_t binCapacity,
                        uint32_t* binBegin, uint32_t* binEnd) {
  if (h == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL handle");
  if (h->plan.orthogonal) return fail(B200SEED_ERR_UNSUPPORTED, "not available on an orthogonal-seeder handle (no grid, no vertex windows)");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const size_t n = h->lastCounters.nInGrid;
  const size_t nBinsAll = (size_t)h->lastEvents * (size_t)h->plan.dev.nGlobalBins;
  if (capacity < n || binCapacity < nBinsAll) return fail(B200SEED_ERR_CAPACITY, "debug_grid buffers too small");
  std::vector<float2> xy(n), zr(n), var(n);
  std::vector<uint32_t> starts(nBinsAll + 1);
  if (n > 0) {
    CUDA_TRY(cudaMemcpy(copiedFromIndex, h->pIdx.ptr, n * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(xy.data(), h->pXY.ptr, n * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(zr.data(), h->pZR.ptr, n * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(var.data(), h->pVar.ptr, n * 8, cudaMemcpyDeviceToHost));
  }
  CUDA_TRY(cudaMemcpy(starts.data(), h->binStart.ptr, (nBinsAll + 1) * 4, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) {
    x[i] = xy[i].x; y[i] = xy[i].y; z[i] = zr[i].x; r[i] = zr[i].y; varZ[i] = var[i].x; varR[i] = var[i].y;
  }
  for (size_t b = 0; b < nBinsAll; ++b) {
    binBegin[b] = starts[b];
    binEnd[b] = starts[b + 1];
  }
  return B200SEED_OK;
}

// Stage-level view of the PRODUCTION doublet stage: the fill pass of the last call is run again chunk by chunk and
// the arena is copied out (bottoms first, then tops, per middle, in the reference's emission order).  Middles that
// do not reach the triplet stage (no tops, no bottoms, or -- seedConfirmation -- too few tops) have no bottoms here:
// like the reference, the engine does not search them (TripletSeeder.cpp:62-69).
int b200seed_debug_doublets(b200seed_handle* h, b200seed_doublets* out) {
  if (h == nullptr || out == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  if (h->plan.orthogonal) return fail(B200SEED_ERR_UNSUPPORTED, "not available on an orthogonal-seeder handle (no grid, no vertex windows)");
  if (h->lastEvents == 0) return fail(B200SEED_ERR_INVALID_ARGUMENT, "no previous run on this handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaDeviceSynchronize());
  cudaStream_t s = h->stream;
  const uint32_t nWork = (uint32_t)h->lastCounters.nMiddles;
  const uint32_t nChunks = h->lastChunkBounds.empty() ? 0u : (uint32_t)h->lastChunkBounds.size() - 1;
  DoubletParams dp = h->lastDoublets;
  std::vector<MiddleHeader> hdr(nWork);
  std::vector<uint32_t> capT(nWork);
  uint64_t total = 0;
  out->nMiddles = nWork;
  // the headers of every chunk are still in place after the run
  if (nWork > 0) {
    CUDA_TRY(cudaMemcpy(hdr.data(), h->hdr.ptr, (size_t)nWork * sizeof(MiddleHeader), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(capT.data(), h->capT.ptr, (size_t)nWork * 4, cudaMemcpyDeviceToHost));
  }
  for (uint32_t w = 0; w < nWork; ++w) total += (uint64_t)hdr[w].nB + hdr[w].nT;
  out->nDoublets = total;
  out->gpuMilliseconds = h->stageMs[4] + h->stageMs[5];
  if (out->middlePos == nullptr) return B200SEED_OK;  // size query
  if (out->middleCapacity < nWork || out->doubletCapacity < total) return fail(B200SEED_ERR_CAPACITY, "debug_doublets buffers too small");
  if (nWork > 0) CUDA_TRY(cudaMemcpy(out->middlePos, h->workPos.ptr, (size_t)nWork * 4, cudaMemcpyDeviceToHost));
  uint64_t o = 0;
  uint32_t* scratchCounters = h->workCounter.as<uint32_t>() + kChunkCounterWords * ((size_t)kMaxChunks + 1);
  std::vector<DoubletRecord> rec;
  for (uint32_t c = 0; c < nChunks; ++c) {
    const uint32_t w0 = h->lastChunkBounds[c], w1 = h->lastChunkBounds[c + 1];
    CUDA_TRY(cudaMemsetAsync(scratchCounters, 0, kChunkCounterWords * 4, s));
    dp.itemFirst = w0;
    dp.itemEnd = w1;
    dp.workCounter = scratchCounters;
    dp.classCount = scratchCounters + 16;
    dp.counters = h->counters.as<unsigned long long>();  // scribbled on: the host copy of the last run is what counts
    dp.rec = h->arenaRec[0].as<DoubletRecord>();
    dp.key = h->arenaKey[0].as<float>();
    dp.classList = h->classList.as<uint32_t>();
    k_doublets<true><<<h->smCount * h->doubletBlocksPerSM[1], kDoubletWarps * 32, 0, s>>>(dp);
    CUDA_TRY(cudaStreamSynchronize(s));
    CUDA_TRY(cudaGetLastError());
    uint64_t slots = 0;
    for (uint32_t w = w0; w < w1; ++w) slots = std::max<uint64_t>(slots, (uint64_t)hdr[w].offset + hdr[w].capB + capT[w]);
    rec.resize(std::max<uint64_t>(slots, 1));
    if (slots > 0) CUDA_TRY(cudaMemcpy(rec.data(), h->arenaRec[0].ptr, slots * sizeof(DoubletRecord), cudaMemcpyDeviceToHost));
    for (uint32_t w = w0; w < w1; ++w) {
      out->firstDoublet[w] = o;
      out->nBottom[w] = hdr[w].nB;
      for (int side = 0; side < 2; ++side) {
        const DoubletRecord* src = rec.data() + hdr[w].offset + (side == 0 ? 0u : hdr[w].capB);
        const uint32_t n = side == 0 ? hdr[w].nB : hdr[w].nT;
        for (uint32_t i = 0; i < n; ++i, ++o) {
          out->otherPos[o] = src[i].pos;
          out->cotTheta[o] = src[i].cotTheta; out->iDeltaR[o] = src[i].iDeltaR; out->er[o] = src[i].er;
          out->u[o] = src[i].u; out->v[o] = src[i].v; out->xNew[o] = src[i].xNew; out->yNew[o] = src[i].yNew;
        }
      }
    }
  }
  out->firstDoublet[nWork] = o;
  return B200SEED_OK;
}

int b200seed_estimate_params(b200seed_handle* h, uint64_t nSeeds, const uint32_t* bottom, const uint32_t* middle,
                             const uint32_t* top, uint32_t nSpacePoints, const float* x, const float* y, const float* z,
                             const double* bField, double* freeParams) {
  if (h == nullptr || bField == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  if (nSeeds == 0) return B200SEED_OK;
  if (bottom == nullptr || middle == nullptr || top == nullptr || x == nullptr || y == nullptr || z == nullptr ||
      freeParams == nullptr) {
    return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const double bNorm2 = bField[0] * bField[0] + bField[1] * bField[1] + bField[2] * bField[2];
  if (!(bNorm2 > 0.0)) return fail(B200SEED_ERR_INVALID_ARGUMENT, "estimate_params: magnetic field of zero (or NaN) norm");
  DevBuf idx[3], col[3], out, st;
  struct Guard {
    std::vector<DevBuf*> bufs;
    ~Guard() { for (DevBuf* b : bufs) b->release(); }
  } guard;
  guard.bufs = {&idx[0], &idx[1], &idx[2], &col[0], &col[1], &col[2], &out, &st};
  CUDA_TRY(st.reserve(16));
  CUDA_TRY(cudaMemsetAsync(st.ptr, 0, 16, s));
  const uint32_t* hi[3] = {bottom, middle, top};
  const float* hc[3] = {x, y, z};
  for (int k = 0; k < 3; ++k) {
    CUDA_TRY(idx[k].reserve(nSeeds * 4));
    CUDA_TRY(col[k].reserve(std::max<size_t>(4, (size_t)nSpacePoints * 4)));
    CUDA_TRY(cudaMemcpyAsync(idx[k].ptr, hi[k], nSeeds * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(col[k].ptr, hc[k], (size_t)nSpacePoints * 4, cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(out.reserve(nSeeds * 64));
  const int blocks = (int)std::min<uint64_t>((nSeeds + 255) / 256, (uint64_t)h->smCount * 8);
  k_estimate_params<<<blocks, 256, 0, s>>>(idx[0].as<uint32_t>(), idx[1].as<uint32_t>(), idx[2].as<uint32_t>(),
                                           col[0].as<float>(), col[1].as<float>(), col[2].as<float>(), bField[0], bField[1],
                                           bField[2], out.as<double>(), nSeeds, nSpacePoints, st.as<int>());
  CUDA_TRY(cudaGetLastError());
  int bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&bad, st.ptr, 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(freeParams, out.ptr, nSeeds * 64, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (bad != 0) return fail(B200SEED_ERR_INVALID_ARGUMENT, "estimate_params: a seed refers to a space point index >= nSpacePoints");
  return B200SEED_OK;
}

int b200seed_make_pixel_spacepoints(b200seed_handle* h, uint32_t n, const uint32_t* surface, const double* loc0,
                                    const double* loc1, const double* cov00, const double* cov01, const double* cov11,
                                    uint32_t nSurfaces, const double* transforms, float* x, float* y, float* z,
                                    float* r, float* varZ, float* varR) {
  if (h == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL handle");
  if (n == 0) return B200SEED_OK;
  const double* dcol[5] = {loc0, loc1, cov00, cov01, cov11};
  float* ocol[6] = {x, y, z, r, varZ, varR};
  for (const double* c : dcol) if (c == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL measurement column");
  for (float* c : ocol) if (c == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL space point column");
  if (surface == nullptr || transforms == nullptr || nSurfaces == 0) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL / empty surface table");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  DevBuf dsurf, din[5], dtr, dout[6], dstat;
  struct Guard {
    std::vector<DevBuf*> bufs;
    ~Guard() { for (DevBuf* b : bufs) b->release(); }
  } guard;
  guard.bufs = {&dsurf, &dtr, &dstat};
  for (DevBuf& b : din) guard.bufs.push_back(&b);
  for (DevBuf& b : dout) guard.bufs.push_back(&b);
  CUDA_TRY(dsurf.reserve((size_t)n * 4));
  CUDA_TRY(cudaMemcpyAsync(dsurf.ptr, surface, (size_t)n * 4, cudaMemcpyHostToDevice, s));
  for (int k = 0; k < 5; ++k) {
    CUDA_TRY(din[k].reserve((size_t)n * 8));
    CUDA_TRY(cudaMemcpyAsync(din[k].ptr, dcol[k], (size_t)n * 8, cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(dtr.reserve((size_t)nSurfaces * 96));
  CUDA_TRY(cudaMemcpyAsync(dtr.ptr, transforms, (size_t)nSurfaces * 96, cudaMemcpyHostToDevice, s));
  for (DevBuf& b : dout) CUDA_TRY(b.reserve((size_t)n * 4));
  CUDA_TRY(dstat.reserve(16));
  CUDA_TRY(cudaMemsetAsync(dstat.ptr, 0, 16, s));
  SpacePointMakerParams mp{};
  mp.surface = dsurf.as<uint32_t>();
  mp.loc0 = din[0].as<double>(); mp.loc1 = din[1].as<double>();
  mp.cov00 = din[2].as<double>(); mp.cov01 = din[3].as<double>(); mp.cov11 = din[4].as<double>();
  mp.transforms = dtr.as<double>();
  mp.x = dout[0].as<float>(); mp.y = dout[1].as<float>(); mp.z = dout[2].as<float>(); mp.r = dout[3].as<float>();
  mp.varZ = dout[4].as<float>(); mp.varR = dout[5].as<float>();
  mp.n = n; mp.nSurfaces = nSurfaces;
  mp.status = dstat.as<int>();
  const int blocks = (int)std::min<uint64_t>(((uint64_t)n + 255) / 256, (uint64_t)h->smCount * 8);
  k_pixel_spacepoints<<<blocks, 256, 0, s>>>(mp);
  CUDA_TRY(cudaGetLastError());
  for (int k = 0; k < 6; ++k) CUDA_TRY(cudaMemcpyAsync(ocol[k], dout[k].ptr, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  int st = 0;
  CUDA_TRY(cudaMemcpyAsync(&st, dstat.ptr, 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (st != 0) return fail(B200SEED_ERR_INVALID_ARGUMENT, "surface index out of range");
  return B200SEED_OK;
}

int b200seed_debug_atan2f(b200seed_handle* h, uint64_t n, const float* y, const float* x, float* phi) {
  if (h == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL handle");
  if (n == 0) return B200SEED_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  if (y == nullptr || x == nullptr || phi == nullptr) return fail(B200SEED_ERR_INVALID_ARGUMENT, "NULL argument");
  DevBuf dy, dx, dp;
  struct Guard {
    std::vector<DevBuf*> bufs;
    ~Guard() { for (DevBuf* b : bufs) b->release(); }
  } guard;
  guard.bufs = {&dy, &dx, &dp};
  CUDA_TRY(dy.reserve(n * 4));
  CUDA_TRY(dx.reserve(n * 4));
  CUDA_TRY(dp.reserve(n * 4));
  CUDA_TRY(cudaMemcpyAsync(dy.ptr, y, n * 4, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(dx.ptr, x, n * 4, cudaMemcpyHostToDevice, h->stream));
  k_atan2f<<<h->smCount * 8, 256, 0, h->stream>>>(dy.as<float>(), dx.as<float>(), dp.as<float>(), n);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(phi, dp.ptr, n * 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return B200SEED_OK;
}

}  // extern "C"
