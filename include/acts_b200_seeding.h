/*
 * acts_b200_seeding.h -- C ABI of the B200 triplet-seeding plugin.
 *
 * This is the drop-in boundary for ACTS's grid triplet seeding hot path.
 * Everything here is plain C (pointers + sizes, int status codes, no C++ or
 * torch types) so that a `Plugins/B200Seeding` library inside ACTS, a ctypes
 * harness, or any other FFI can bind it.  Reference interfaces replaced
 * (paths relative to the ACTS source tree):
 *
 *   b200seed_config      <- ActsExamples::GridTripletSeedingAlgorithm::Config
 *                           Examples/Algorithms/TrackFinding/include/ActsExamples/
 *                           TrackFinding/GridTripletSeedingAlgorithm.hpp:34-244
 *                           (same field names, same defaults, same units: mm, GeV,
 *                           bFieldInZ in GeV/mm i.e. 1 T = 0.000299792458)
 *   b200seed_create      <- GridTripletSeedingAlgorithm ctor
 *                           GridTripletSeedingAlgorithm.cpp:101-178 (config fan-out,
 *                           validation) + CylindricalSpacePointGrid ctor
 *                           Core/src/Seeding/CylindricalSpacePointGrid.cpp:15-99
 *                           + DerivedConfig ctors DoubletSeedFinder.cpp:351-357,
 *                           TripletSeedFinder.cpp:456-481
 *   b200seed_run         <- GridTripletSeedingAlgorithm::execute
 *                           GridTripletSeedingAlgorithm.cpp:180-402
 *   b200seed_run_batch   <- Sequencer event loop calling execute once per event
 *                           Examples/Framework/src/Framework/Sequencer.cpp:472-525
 *   b200seed_seeds       <- Acts::SeedContainer columns (spacePoints, quality,
 *                           vertexZ) Core/include/Acts/EventData/SeedContainer.hpp:208-213
 *
 * Threading contract: one handle owns one CUDA device, one stream and its
 * device workspaces.  A handle must not be used from two threads at once;
 * create one handle per worker thread (the reference uses a thread_local
 * cache the same way, GridTripletSeedingAlgorithm.cpp:337).
 *
 * There is no CPU fallback: every entry point that computes returns
 * B200SEED_ERR_CUDA when no usable device is present.
 */
#ifndef ACTS_B200_SEEDING_H
#define ACTS_B200_SEEDING_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SEED_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------- */
enum {
  B200SEED_OK = 0,
  B200SEED_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument in the reference  */
  B200SEED_ERR_RUNTIME = 2,          /* std::runtime_error (grid range checks)   */
  B200SEED_ERR_DOMAIN = 3,           /* std::domain_error (phi binning)          */
  B200SEED_ERR_UNSUPPORTED = 4,      /* valid reference config not yet on device */
  B200SEED_ERR_CUDA = 5,             /* CUDA runtime / no device                 */
  B200SEED_ERR_CAPACITY = 6,         /* caller-provided seed buffers too small   */
  B200SEED_ERR_OVERFLOW = 7          /* internal per-middle scratch exhausted    */
};

/* Acts::SeedConfirmationRangeConfig
 * (Core/include/Acts/Seeding/SeedConfirmationRangeConfig.hpp) */
typedef struct b200seed_seed_confirmation_range {
  float zMinSeedConf;
  float zMaxSeedConf;
  float rMaxSeedConf;
  uint64_t nTopForLargeR;
  uint64_t nTopForSmallR;
  float seedConfMinBottomRadius;
  float seedConfMaxZOrigin;
  float minImpactSeedConf;
} b200seed_seed_confirmation_range;

/* Flattened GridTripletSeedingAlgorithm::Config.  Field names, meaning and
 * defaults are the reference's (GridTripletSeedingAlgorithm.hpp:34-244).
 * Vectors become (pointer, count) pairs; the library copies them in
 * b200seed_create, the caller keeps ownership.  NaN in deltaR{Min,Max}{Top,Bottom}
 * means "use deltaRMin/deltaRMax" exactly like the reference. */
typedef struct b200seed_config {
  uint32_t abi_version; /* must be B200SEED_ABI_VERSION */
  uint32_t struct_size; /* sizeof(b200seed_config) of the caller */

  /* general seeding parameters */
  float bFieldInZ;
  float minPt;
  float cotThetaMax;
  float impactMax;
  float deltaRMin;
  float deltaRMax;
  float deltaRMinTop;
  float deltaRMaxTop;
  float deltaRMinBottom;
  float deltaRMaxBottom;

  /* grid */
  float rMin; /* ignored by the grid like in the reference (.cpp:133-134) */
  float rMax;
  float zMin;
  float zMax;
  float phiMin;
  float phiMax;
  int32_t phiBinDeflectionCoverage;
  int32_t maxPhiBins;
  const int32_t* zBinNeighborsTop; /* nZBinNeighborsTop pairs (first, second) */
  uint32_t nZBinNeighborsTop;
  const int32_t* zBinNeighborsBottom;
  uint32_t nZBinNeighborsBottom;
  int32_t numPhiNeighbors;
  const float* zBinEdges;
  uint32_t nZBinEdges;
  const uint64_t* zBinsCustomLooping;
  uint32_t nZBinsCustomLooping;

  /* middle space point region of interest */
  float rMinMiddle;
  float rMaxMiddle;
  uint8_t useVariableMiddleSPRange;
  const float* rRangeMiddleSP; /* nRRangeMiddleSP pairs (rMin, rMax) per z bin */
  uint32_t nRRangeMiddleSP;
  float deltaRMiddleMinSPRange;
  float deltaRMiddleMaxSPRange;

  /* doublet cuts */
  float deltaZMin;
  float deltaZMax;
  uint8_t interactionPointCut;
  float collisionRegionMin;
  float collisionRegionMax;
  float helixCutTolerance;

  /* triplet cuts */
  float sigmaScattering;
  float radLengthPerSeed;
  float toleranceParam;

  /* seed filter */
  float deltaInvHelixDiameter;
  float compatSeedWeight;
  float impactWeightFactor;
  float zOriginWeightFactor;
  uint32_t maxSeedsPerSpM;
  uint64_t compatSeedLimit;
  float seedWeightIncrement;
  float numSeedIncrement;
  uint8_t seedConfirmation;
  b200seed_seed_confirmation_range centralSeedConfirmationRange;
  b200seed_seed_confirmation_range forwardSeedConfirmationRange;
  uint32_t maxSeedsPerSpMConf;
  uint32_t maxQualitySeedsPerSpMConf;
  uint8_t useDeltaRinsteadOfTopRadius;

  /* other */
  uint8_t useExtraCuts; /* ITk fast-tracking SP selector + doublet cut (.cpp:33-62) */
  /* Vertex-z constraint (Config::inputVertices / vertexZNSigma / vertexZMargin,
   * GridTripletSeedingAlgorithm.hpp:239-243).  useVertexZCuts = 1 stands for a
   * non-empty `inputVertices` key: the VertexZCuts functor then takes the doublet
   * experiment-cut slot for EVERY event (.cpp:292-297) -- the ITk doublet cut of
   * useExtraCuts is not applied, and an event without vertices accepts every
   * doublet (.cpp:77-79).  b200seed_run_vertices builds the windows from
   * (z, var z) like .cpp:187-206 with the two parameters below. */
  uint8_t useVertexZCuts;
  double vertexZNSigma;
  double vertexZMargin;

  /* Engine options (no reference counterpart).
   * relaxedFloat = 0: bit-exact binary32 replay of the reference cuts (default).
   * relaxedFloat = 1: FMA contraction + approximate reciprocals ("float fast
   * path" of the north star); results are NOT guaranteed identical. */
  uint8_t relaxedFloat;
} b200seed_config;

/* Caller-owned seed output columns (host memory for b200seed_run*, device
 * memory for the *_device entry points).  Indices refer to the caller's
 * ORIGINAL space point arrays (reference .cpp:394-398).  Seeds of one event
 * are written in the reference's order: bin groups in navigation order, middle
 * space points by ascending radius, per-middle candidates by descending weight. */
typedef struct b200seed_seeds {
  uint32_t* bottom;
  uint32_t* middle;
  uint32_t* top;
  float* quality;
  float* vertexZ;
  uint64_t capacity; /* elements available in each column */
  uint64_t size;     /* out: seeds written (or required, on ERR_CAPACITY) */
} b200seed_seeds;

/* Derived quantities, for known-answer tests against the reference's own
 * expressions (SpacePointGridPhiBinning.cpp:20-95, DoubletSeedFinder.cpp:351-357,
 * TripletSeedFinder.cpp:456-481). */
typedef struct b200seed_info {
  int32_t phiBins;
  int32_t zBins;
  int32_t rBins;
  int32_t nGlobalBins; /* (phiBins+2)*(zBins+2)*(rBins+2) */
  float minHelixDiameter2;
  float highland;
  float sigmapT2perRadius;
  float multipleScattering2;
  float deltaRMinBottom, deltaRMaxBottom, deltaRMinTop, deltaRMaxTop;
  int32_t smCount;
  int32_t ccMajor, ccMinor;
} b200seed_info;

/* Per-call counters of the last b200seed_run* on this handle (whole batch). */
typedef struct b200seed_counters {
  uint64_t nSpacePoints;   /* N  : input space points                       */
  uint64_t nInGrid;        /* N' : space points inside the grid             */
  uint64_t nMiddles;       /* M  : middle space points in the r window      */
  uint64_t nBottomDoublets;/* D_b                                           */
  uint64_t nTopDoublets;   /* D_t                                           */
  uint64_t nTripletTests;  /* (bottom, top) pairs evaluated                 */
  uint64_t nCandidates;    /* T  : triplet candidates passed to the filter  */
  uint64_t nSeeds;         /* S                                             */
  uint64_t nTieMiddles;    /* middles that needed the exact-tie slow path   */
  uint64_t nKernelLaunches;/* kernels launched by the call                  */
  uint64_t nConfirmationRounds; /* seedConfirmation: rounds until the bestSeedQualityMap fixed point (0 otherwise) */
} b200seed_counters;

typedef struct b200seed_handle b200seed_handle;

/* Fill *cfg with the reference defaults (GridTripletSeedingAlgorithm.hpp:34-244). */
int b200seed_config_init(b200seed_config* cfg);

/* Host-only planning (no GPU needed): validation exactly like the reference
 * constructor chain + the derived constants.  Same status codes as
 * b200seed_create. */
int b200seed_plan_info(const b200seed_config* cfg, b200seed_info* info);

/* Host-only: device constant block and navigation / neighbour-bin tables of a
 * config.  Arrays may be NULL to query the sizes: sizes[0] = #navigation
 * entries (middle bins in visiting order), sizes[1] / sizes[2] = flattened
 * bottom / top neighbour bins, sizes[3] = bytes of the device constant block,
 * sizes[4] = seed slots per middle. */
int b200seed_plan_tables(const b200seed_config* cfg, void* deviceConfig, uint64_t deviceConfigBytes,
                         uint32_t* navBins, uint32_t* botOffsets, uint32_t* botBins,
                         uint32_t* topOffsets, uint32_t* topBins, uint64_t* sizes);

/* Validate the config exactly like the reference constructor chain, compute
 * the derived constants on the host with the reference's expressions, build
 * the bin-neighbour tables and allocate device workspaces on `device`. */
int b200seed_create(const b200seed_config* cfg, int device, b200seed_handle** out);
void b200seed_destroy(b200seed_handle* h);

/* ---- OrthogonalTripletSeedingAlgorithm (the k-d-tree candidate provider feeding the same
 * doublet / triplet / filter stages; Examples/Algorithms/TrackFinding/include/ActsExamples/
 * TrackFinding/OrthogonalTripletSeedingAlgorithm.hpp:33-186, src/...cpp:62-317) -----------------
 * Its Config is a subset of the grid algorithm's fields (same names, other defaults) plus the
 * three below.  Grid-only members of the config struct -- phiBinDeflectionCoverage, maxPhiBins,
 * zBinNeighbors*, numPhiNeighbors, zBinEdges, zBinsCustomLooping, rRangeMiddleSP, vertex cuts --
 * are ignored.  seedConfirmation = true is supported (one filter state across both groups of a middle and
 * across middles, OrthogonalTripletSeedingAlgorithm.cpp:234-238).
 * The handle is used with b200seed_run / b200seed_run_batch / b200seed_sync like a grid handle;
 * seeds come in the reference's order (middles in k-d-tree order, the increasing-z group of a
 * middle before its decreasing-z group). */
typedef struct b200seed_orthogonal_options {
  float zOutermostLayersMin; /* Config::zOutermostLayers.first  (hpp:102-103) */
  float zOutermostLayersMax; /* Config::zOutermostLayers.second */
  float deltaPhiMax;         /* Config::deltaPhiMax (hpp:110) */
} b200seed_orthogonal_options;
/* Reference defaults of OrthogonalTripletSeedingAlgorithm::Config (hpp:38-186). */
int b200seed_orthogonal_config_init(b200seed_config* cfg, b200seed_orthogonal_options* opt);
int b200seed_create_orthogonal(const b200seed_config* cfg, const b200seed_orthogonal_options* opt, int device,
                               b200seed_handle** out);

/* Page-locked host memory for the caller's column / seed buffers (cudaMallocHost / cudaFreeHost): copies from and
 * to such buffers run at the full link rate and asynchronously.  NULL on failure. */
void* b200seed_alloc_pinned(size_t bytes);
void b200seed_free_pinned(void* p);

/* Error text of the last failing call on this thread (never NULL). */
const char* b200seed_last_error(void);

int b200seed_get_info(const b200seed_handle* h, b200seed_info* info);
int b200seed_get_counters(const b200seed_handle* h, b200seed_counters* c);

/* One event, host buffers (pageable or pinned).  x,y,z,r,varZ,varR are the six
 * float columns of the input SpacePointContainer (SpacePointMaker.cpp:260-264).
 * zWindowLo/Hi (nZWindows entries, may be NULL/0) are the per-event vertex
 * z-windows of the reference's VertexZCuts (.cpp:69-97,187-206).
 * Synchronous: returns after the seeds are in `out`. */
int b200seed_run(b200seed_handle* h, uint32_t nSpacePoints, const float* x,
                 const float* y, const float* z, const float* r,
                 const float* varZ, const float* varR, uint32_t nZWindows,
                 const float* zWindowLo, const float* zWindowHi,
                 b200seed_seeds* out);

/* Config::inputVertices (GridTripletSeedingAlgorithm.hpp:239-243, .cpp:187-206): one event together with its
 * reconstructed vertices (z position and its variance, the (2, 2) element of the vertex covariance).  The z
 * windows are built exactly like the reference does (double arithmetic, narrowed to float); any number of
 * vertices, including none (every doublet then passes).  The handle must have useVertexZCuts = 1. */
int b200seed_run_vertices(b200seed_handle* h, uint32_t nSpacePoints, const float* x,
                          const float* y, const float* z, const float* r,
                          const float* varZ, const float* varR, uint32_t nVertices,
                          const double* vertexZ, const double* vertexVarZ,
                          b200seed_seeds* out);

/* The strip triplet path: TripletSeedFinder::Config::useStripInfo = true
 * (Core/src/Seeding/TripletSeedFinder.cpp:164-406, createStripTripletTopCandidates; calibration
 * Core/include/Acts/SpacePointFormation/detail/StripSpacePointCalibrationImpl.hpp:22-85).  One event whose space
 * points carry the StripCalibrationDetails column (SpacePointContainer.hpp:249-256): stripDetails holds 12 floats per
 * space point = outerCenter, innerToOuterSeparation, outerHalfVector, innerHalfVector
 * (StripSpacePointCalibrationDetails.hpp:16-29).  cotThetaDiffMax is TripletSeedFinder::Config::cotThetaDiffMax
 * (TripletSeedFinder.hpp:171-175; INFINITY = the reference's default: no pre-filter); the tolerance of the module
 * check is the config's toleranceParam.  Grid handles only (B200SEED_ERR_UNSUPPORTED on an orthogonal handle); the
 * doublet stage, the seed filter and the output order are those of b200seed_run. */
int b200seed_run_strips(b200seed_handle* h, uint32_t nSpacePoints, const float* x,
                        const float* y, const float* z, const float* r,
                        const float* varZ, const float* varR, const float* stripDetails,
                        float cotThetaDiffMax, b200seed_seeds* out);

/* The window construction of b200seed_run_vertices on its own (for the batch entry below). */
int b200seed_vertex_windows(const b200seed_handle* h, uint32_t nVertices,
                            const double* vertexZ, const double* vertexVarZ,
                            float* windowLo, float* windowHi);

/* b200seed_run_batch with per-event z windows: event e owns windows
 * [windowOffsets[e], windowOffsets[e + 1]) of (zWindowLo, zWindowHi). */
int b200seed_run_batch_windows(b200seed_handle* h, uint32_t nEvents,
                               const uint32_t* spOffsets, const float* x,
                               const float* y, const float* z, const float* r,
                               const float* varZ, const float* varR,
                               const uint32_t* windowOffsets, const float* zWindowLo,
                               const float* zWindowHi, uint64_t* seedOffsets,
                               b200seed_seeds* out);

/* b200seed_run with a caller-provided azimuth column (phi[i] replaces the
 * on-device atan2f(y, x) replay; everything else is identical). */
int b200seed_run_with_phi(b200seed_handle* h, uint32_t nSpacePoints, const float* x,
                          const float* y, const float* z, const float* r,
                          const float* varZ, const float* varR, const float* phi,
                          b200seed_seeds* out);

/* A batch of independent events, host buffers.  Event e owns space points
 * [spOffsets[e], spOffsets[e+1]) of the concatenated columns; its seeds are
 * written to out columns [seedOffsets[e], seedOffsets[e+1]) and their indices
 * are relative to the event's own first space point (so each event looks
 * exactly like a b200seed_run call).  seedOffsets has nEvents+1 entries. */
int b200seed_run_batch(b200seed_handle* h, uint32_t nEvents,
                       const uint32_t* spOffsets, const float* x,
                       const float* y, const float* z, const float* r,
                       const float* varZ, const float* varR,
                       uint64_t* seedOffsets, b200seed_seeds* out);

/* Same as b200seed_run_batch but every pointer (spOffsets, columns,
 * seedOffsets, out->columns) is DEVICE memory on the handle's device and the
 * work is enqueued on `cudaStream` (a cudaStream_t passed as void*; NULL = the
 * handle's own stream).  Asynchronous: out->size is only valid after
 * b200seed_sync.  Used to time the path with inputs resident in HBM.  Several
 * calls may be enqueued before one b200seed_sync (counters / errors then refer
 * to the last one); with seedConfirmation = true a call first completes the
 * previous one, because the fixed-point rounds are checked on the host. */
int b200seed_run_batch_device(b200seed_handle* h, uint32_t nEvents,
                              uint32_t nSpacePointsTotal,
                              const uint32_t* spOffsets, const float* x,
                              const float* y, const float* z, const float* r,
                              const float* varZ, const float* varR,
                              uint64_t* seedOffsets, b200seed_seeds* out,
                              void* cudaStream);
/* Wait for the last *_device call; fills out->size, returns deferred errors. */
int b200seed_sync(b200seed_handle* h, b200seed_seeds* out);

/* Single-event split over several GPUs: restrict
 * the following calls on this handle to the middle space points whose phi bin
 * (1-based) lies in [firstPhiBin, firstPhiBin + nPhiBins); nPhiBins = 0 restores
 * "all".  Sector results concatenated in sector order equal the unsplit result.
 * Valid because seedConfirmation = false has no cross-middle state; refused
 * with B200SEED_ERR_UNSUPPORTED on a seedConfirmation handle. */
int b200seed_set_phi_sector(b200seed_handle* h, uint32_t firstPhiBin, uint32_t nPhiBins);
/* firstPhiBin = B200SEED_PHI_SECTOR_EMPTY with nPhiBins = 0: the EMPTY sector (a rank that got no bins because
 * there are more ranks than phi bins seeds nothing, instead of falling back to "all"). */
#define B200SEED_PHI_SECTOR_EMPTY 0xFFFFFFFFu

/* GPU time of the stages of the last completed call, milliseconds (CUDA events
 * on the launching stream): ms[0] grid build, ms[1] middle work list,
 * ms[2] seeding kernel, ms[3] ordered seed compaction. */
int b200seed_get_stage_times(const b200seed_handle* h, float* ms);
/* The first n (<= 7) of: ms[0..3] as above, ms[4] doublet count pass + slot scan + chunk plan,
 * ms[5] doublet fill pass (all chunks), ms[6] per-middle seeding kernels (all chunks);
 * ms[2] = ms[4] + ms[5] + ms[6]. */
int b200seed_get_stage_times_ex(const b200seed_handle* h, float* ms, uint32_t n);

/* "Next" row f1: free track parameters of every seed, FP64 on the
 * device.  Replaces Acts::estimateTrackParamsFromSeed(sp0, sp1, sp2, bField)
 * (Core/src/Seeding/EstimateTrackParamsFromSeed.cpp:106-160): freeParams holds 8 doubles
 * per seed {x, y, z of the bottom space point, time = 0, direction (3), q/p}.  bottom /
 * middle / top index the float columns x, y, z (host memory); bField is {Bx, By, Bz} in
 * the reference's native units (1 T = 0.000299792458).  Agreement with the reference
 * arithmetic: 1e-9 relative (not bit-exact: the reference goes through Eigen). */
int b200seed_estimate_params(b200seed_handle* h, uint64_t nSeeds, const uint32_t* bottom,
                             const uint32_t* middle, const uint32_t* top, uint32_t nSpacePoints,
                             const float* x, const float* y, const float* z, const double* bField,
                             double* freeParams);

/* "Next" row f4: pixel space points from measurements, FP64 on the
 * device.  Replaces createPixelSpacePoint of the upstream SpacePointMaker
 * (Examples/Algorithms/TrackFinding/src/SpacePointMaker.cpp:44-76): global position
 * PlaneSurface::localToGlobal (Core/src/Surfaces/PlaneSurface.cpp:72-75), r = fastHypot(x, y),
 * (varZ, varR) = diag of PixelSpacePointBuilder::computeCovarianceZR
 * (Core/src/SpacePointFormation/PixelSpacePointBuilder.cpp:17-42).  Measurement i lies on
 * surface[i]; transforms holds nSurfaces row-major 3x4 affine local->global matrices
 * (rotation | translation) of planar surfaces; (loc0, loc1) and (cov00, cov01, cov11) are the
 * local position and covariance of the measurement.  The six outputs are the float columns
 * b200seed_run takes.  All pointers are host memory. */
int b200seed_make_pixel_spacepoints(b200seed_handle* h, uint32_t n, const uint32_t* surface,
                                    const double* loc0, const double* loc1, const double* cov00,
                                    const double* cov01, const double* cov11, uint32_t nSurfaces,
                                    const double* transforms, float* x, float* y, float* z,
                                    float* r, float* varZ, float* varR);

/* b200seed_make_pixel_spacepoints + b200seed_run without the host round trip: the space
 * points of ONE event are made on the device straight into the input columns of the
 * seeding pipeline (SpacePointMaker -> GridTripletSeedingAlgorithm of the reference
 * chain, reconstruction.py:1001-1077).  x .. varR are optional (NULL) host outputs of the
 * space point columns; seed indices refer to the measurement order. */
int b200seed_run_measurements(b200seed_handle* h, uint32_t n, const uint32_t* surface,
                              const double* loc0, const double* loc1, const double* cov00,
                              const double* cov01, const double* cov11, uint32_t nSurfaces,
                              const double* transforms, uint32_t nZWindows,
                              const float* zWindowLo, const float* zWindowHi, float* x, float* y,
                              float* z, float* r, float* varZ, float* varR, b200seed_seeds* out);

/* ---- stage-level introspection (parity tests of the grid / doublet stages) */

/* After a run: the packed, bin-ordered, r-sorted space point copy of the LAST
 * event batch (reference coreSpacePoints, .cpp:230-253).  All arrays are host
 * memory with nInGrid elements except binBegin/binEnd (nEvents*nGlobalBins). */
int b200seed_debug_grid(b200seed_handle* h, uint64_t capacity,
                        uint32_t* copiedFromIndex, float* x, float* y, float* z,
                        float* r, float* varZ, float* varR,
                        uint64_t binCapacity, uint32_t* binBegin,
                        uint32_t* binEnd);

/* Two-pass (count, scan, fill) materialised doublet search for the LAST event
 * batch: for every processed middle (in work order) the compatible bottom and
 * top doublets in the reference's emission order
 * (DoubletSeedFinder.cpp:41-273).  Call with NULL columns to obtain sizes.
 * middlePos / otherPos are positions in the packed copy of b200seed_debug_grid. */
typedef struct b200seed_doublets {
  uint64_t nMiddles;
  uint64_t nDoublets;     /* total entries in the doublet columns            */
  uint32_t* middlePos;    /* [nMiddles]                                      */
  uint64_t* firstDoublet; /* [nMiddles+1] prefix offsets into the columns    */
  uint32_t* nBottom;      /* [nMiddles] bottoms come first, then tops        */
  uint32_t* otherPos;     /* [nDoublets]                                     */
  float* cotTheta;        /* [nDoublets]                                     */
  float* iDeltaR;
  float* er;
  float* u;
  float* v;
  float* xNew;
  float* yNew;
  uint64_t middleCapacity;
  uint64_t doubletCapacity;
  float gpuMilliseconds;  /* out: GPU time of count + scan + fill (CUDA events) */
} b200seed_doublets;
int b200seed_debug_doublets(b200seed_handle* h, b200seed_doublets* out);

/* phi = atan2f(y, x) exactly as glibc 2.39 computes it, evaluated ON THE
 * DEVICE for n host values (used to validate the device replay against the
 * host libm the reference calls at .cpp:219). */
int b200seed_debug_atan2f(b200seed_handle* h, uint64_t n, const float* y,
                          const float* x, float* phi);

#ifdef __cplusplus
}
#endif
#endif /* ACTS_B200_SEEDING_H */
