// seeding_abi.cpp -- the exported C ABI (include/acts_b200_seeding.h).
//
// The library contains two engines built from the same sources (engine_symbols.h):
//   b200ex_*  exact binary32 replay of the reference cuts (bit-identical seeds)
//   b200rx_*  relaxedFloat fast path (FMA contraction, approximate division / sqrt,
//             CUDA atan2f, no tie-order replay) -- results NOT guaranteed identical
// b200seed_create routes a handle to one of them by cfg->relaxedFloat; every other
// call is forwarded to the engine the handle belongs to.  Host-only planning
// (validation, derived constants, tables) is the same code in both; the exact one serves it.
#include <new>

#include "../../include/acts_b200_seeding.h"

extern "C" {

#define B200SEED_DECLARE_ENGINE(P)                                                                          \
  struct P##handle;                                                                                         \
  int P##config_init(b200seed_config*);                                                                     \
  int P##plan_info(const b200seed_config*, b200seed_info*);                                                 \
  int P##plan_tables(const b200seed_config*, void*, uint64_t, uint32_t*, uint32_t*, uint32_t*, uint32_t*,   \
                     uint32_t*, uint64_t*);                                                                 \
  int P##create(const b200seed_config*, int, P##handle**);                                                  \
  int P##create_orthogonal(const b200seed_config*, const b200seed_orthogonal_options*, int, P##handle**);   \
  int P##orthogonal_config_init(b200seed_config*, b200seed_orthogonal_options*);                            \
  void P##destroy(P##handle*);                                                                              \
  const char* P##last_error(void);                                                                          \
  void* P##alloc_pinned(size_t);                                                                            \
  void P##free_pinned(void*);                                                                               \
  int P##get_info(const P##handle*, b200seed_info*);                                                        \
  int P##get_counters(const P##handle*, b200seed_counters*);                                                \
  int P##run(P##handle*, uint32_t, const float*, const float*, const float*, const float*, const float*,    \
             const float*, uint32_t, const float*, const float*, b200seed_seeds*);                          \
  int P##run_with_phi(P##handle*, uint32_t, const float*, const float*, const float*, const float*,         \
                      const float*, const float*, const float*, b200seed_seeds*);                           \
  int P##run_batch(P##handle*, uint32_t, const uint32_t*, const float*, const float*, const float*,         \
                   const float*, const float*, const float*, uint64_t*, b200seed_seeds*);                   \
  int P##run_batch_device(P##handle*, uint32_t, uint32_t, const uint32_t*, const float*, const float*,      \
                          const float*, const float*, const float*, const float*, uint64_t*,                \
                          b200seed_seeds*, void*);                                                          \
  int P##sync(P##handle*, b200seed_seeds*);                                                                 \
  int P##set_phi_sector(P##handle*, uint32_t, uint32_t);                                                    \
  int P##get_stage_times(const P##handle*, float*);                                                         \
  int P##get_stage_times_ex(const P##handle*, float*, uint32_t);                                            \
  int P##run_vertices(P##handle*, uint32_t, const float*, const float*, const float*, const float*,         \
                      const float*, const float*, uint32_t, const double*, const double*, b200seed_seeds*); \
  int P##vertex_windows(const P##handle*, uint32_t, const double*, const double*, float*, float*);          \
  int P##run_strips(P##handle*, uint32_t, const float*, const float*, const float*, const float*,           \
                    const float*, const float*, const float*, float, b200seed_seeds*);                      \
  int P##run_batch_windows(P##handle*, uint32_t, const uint32_t*, const float*, const float*, const float*, \
                           const float*, const float*, const float*, const uint32_t*, const float*,         \
                           const float*, uint64_t*, b200seed_seeds*);                                       \
  int P##estimate_params(P##handle*, uint64_t, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t, \
                         const float*, const float*, const float*, const double*, double*);                 \
  int P##make_pixel_spacepoints(P##handle*, uint32_t, const uint32_t*, const double*, const double*,        \
                                const double*, const double*, const double*, uint32_t, const double*,       \
                                float*, float*, float*, float*, float*, float*);                            \
  int P##run_measurements(P##handle*, uint32_t, const uint32_t*, const double*, const double*, const double*, \
                          const double*, const double*, uint32_t, const double*, uint32_t, const float*,     \
                          const float*, float*, float*, float*, float*, float*, float*, b200seed_seeds*);    \
  int P##debug_grid(P##handle*, uint64_t, uint32_t*, float*, float*, float*, float*, float*, float*,        \
                    uint64_t, uint32_t*, uint32_t*);                                                        \
  int P##debug_doublets(P##handle*, b200seed_doublets*);                                                    \
  int P##debug_atan2f(P##handle*, uint64_t, const float*, const float*, float*);

B200SEED_DECLARE_ENGINE(b200ex_)
B200SEED_DECLARE_ENGINE(b200rx_)

}  // extern "C"

struct b200seed_handle {
  b200ex_handle* exact = nullptr;
  b200rx_handle* relaxed = nullptr;
};

namespace {

thread_local int g_lastEngine = 0;  // whose last_error() the caller should see: 0 exact, 1 relaxed, 2 this file
thread_local const char* g_ownError = "";

int own_error(const char* msg) {
  g_lastEngine = 2;
  g_ownError = msg;
  return B200SEED_ERR_INVALID_ARGUMENT;
}

}  // namespace

// forward a call to the engine of the handle
#define B200SEED_FORWARD(h, call_exact, call_relaxed)                \
  do {                                                               \
    if ((h) == nullptr) return own_error("NULL handle");             \
    if ((h)->relaxed != nullptr) {                                   \
      g_lastEngine = 1;                                              \
      return call_relaxed;                                           \
    }                                                                \
    g_lastEngine = 0;                                                \
    return call_exact;                                               \
  } while (0)

extern "C" {

const char* b200seed_last_error(void) {
  if (g_lastEngine == 2) return g_ownError;
  return g_lastEngine == 1 ? b200rx_last_error() : b200ex_last_error();
}

void* b200seed_alloc_pinned(size_t bytes) { return b200ex_alloc_pinned(bytes); }
void b200seed_free_pinned(void* p) { b200ex_free_pinned(p); }

int b200seed_config_init(b200seed_config* cfg) {
  g_lastEngine = 0;
  return b200ex_config_init(cfg);
}

int b200seed_plan_info(const b200seed_config* cfg, b200seed_info* info) {
  g_lastEngine = 0;
  return b200ex_plan_info(cfg, info);
}

int b200seed_plan_tables(const b200seed_config* cfg, void* deviceConfig, uint64_t deviceConfigBytes,
                         uint32_t* navBins, uint32_t* botOffsets, uint32_t* botBins, uint32_t* topOffsets,
                         uint32_t* topBins, uint64_t* sizes) {
  g_lastEngine = 0;
  return b200ex_plan_tables(cfg, deviceConfig, deviceConfigBytes, navBins, botOffsets, botBins, topOffsets, topBins, sizes);
}

int b200seed_create(const b200seed_config* cfg, int device, b200seed_handle** out) {
  if (cfg == nullptr || out == nullptr) return own_error("NULL argument");
  *out = nullptr;
  b200seed_handle* h = new (std::nothrow) b200seed_handle;
  if (h == nullptr) return own_error("out of memory");
  int rc;
  // struct_size / abi_version are validated by the engine; relaxedFloat is the last member of the struct
  const bool relaxed = cfg->struct_size == sizeof(b200seed_config) && cfg->relaxedFloat != 0;
  if (relaxed) {
    g_lastEngine = 1;
    rc = b200rx_create(cfg, device, &h->relaxed);
  } else {
    g_lastEngine = 0;
    rc = b200ex_create(cfg, device, &h->exact);
  }
  if (rc != B200SEED_OK) {
    delete h;
    return rc;
  }
  *out = h;
  return B200SEED_OK;
}

int b200seed_orthogonal_config_init(b200seed_config* cfg, b200seed_orthogonal_options* opt) {
  g_lastEngine = 0;
  return b200ex_orthogonal_config_init(cfg, opt);
}

int b200seed_create_orthogonal(const b200seed_config* cfg, const b200seed_orthogonal_options* opt, int device,
                               b200seed_handle** out) {
  if (cfg == nullptr || opt == nullptr || out == nullptr) return own_error("NULL argument");
  *out = nullptr;
  b200seed_handle* h = new (std::nothrow) b200seed_handle;
  if (h == nullptr) return own_error("out of memory");
  int rc;
  const bool relaxed = cfg->struct_size == sizeof(b200seed_config) && cfg->relaxedFloat != 0;
  if (relaxed) {
    g_lastEngine = 1;
    rc = b200rx_create_orthogonal(cfg, opt, device, &h->relaxed);
  } else {
    g_lastEngine = 0;
    rc = b200ex_create_orthogonal(cfg, opt, device, &h->exact);
  }
  if (rc != B200SEED_OK) {
    delete h;
    return rc;
  }
  *out = h;
  return B200SEED_OK;
}

void b200seed_destroy(b200seed_handle* h) {
  if (h == nullptr) return;
  if (h->exact != nullptr) b200ex_destroy(h->exact);
  if (h->relaxed != nullptr) b200rx_destroy(h->relaxed);
  delete h;
}

int b200seed_get_info(const b200seed_handle* h, b200seed_info* info) {
  B200SEED_FORWARD(h, b200ex_get_info(h->exact, info), b200rx_get_info(h->relaxed, info));
}

int b200seed_get_counters(const b200seed_handle* h, b200seed_counters* c) {
  B200SEED_FORWARD(h, b200ex_get_counters(h->exact, c), b200rx_get_counters(h->relaxed, c));
}

int b200seed_run(b200seed_handle* h, uint32_t n, const float* x, const float* y, const float* z, const float* r,
                 const float* varZ, const float* varR, uint32_t nZWindows, const float* zWindowLo,
                 const float* zWindowHi, b200seed_seeds* out) {
  B200SEED_FORWARD(h, b200ex_run(h->exact, n, x, y, z, r, varZ, varR, nZWindows, zWindowLo, zWindowHi, out),
                   b200rx_run(h->relaxed, n, x, y, z, r, varZ, varR, nZWindows, zWindowLo, zWindowHi, out));
}

int b200seed_run_with_phi(b200seed_handle* h, uint32_t n, const float* x, const float* y, const float* z,
                          const float* r, const float* varZ, const float* varR, const float* phi,
                          b200seed_seeds* out) {
  B200SEED_FORWARD(h, b200ex_run_with_phi(h->exact, n, x, y, z, r, varZ, varR, phi, out),
                   b200rx_run_with_phi(h->relaxed, n, x, y, z, r, varZ, varR, phi, out));
}

int b200seed_run_batch(b200seed_handle* h, uint32_t nEvents, const uint32_t* spOffsets, const float* x,
                       const float* y, const float* z, const float* r, const float* varZ, const float* varR,
                       uint64_t* seedOffsets, b200seed_seeds* out) {
  B200SEED_FORWARD(h, b200ex_run_batch(h->exact, nEvents, spOffsets, x, y, z, r, varZ, varR, seedOffsets, out),
                   b200rx_run_batch(h->relaxed, nEvents, spOffsets, x, y, z, r, varZ, varR, seedOffsets, out));
}

int b200seed_run_batch_device(b200seed_handle* h, uint32_t nEvents, uint32_t nTotal, const uint32_t* spOffsets,
                              const float* x, const float* y, const float* z, const float* r, const float* varZ,
                              const float* varR, uint64_t* seedOffsets, b200seed_seeds* out, void* cudaStream) {
  B200SEED_FORWARD(
      h, b200ex_run_batch_device(h->exact, nEvents, nTotal, spOffsets, x, y, z, r, varZ, varR, seedOffsets, out, cudaStream),
      b200rx_run_batch_device(h->relaxed, nEvents, nTotal, spOffsets, x, y, z, r, varZ, varR, seedOffsets, out, cudaStream));
}

int b200seed_sync(b200seed_handle* h, b200seed_seeds* out) {
  B200SEED_FORWARD(h, b200ex_sync(h->exact, out), b200rx_sync(h->relaxed, out));
}

int b200seed_set_phi_sector(b200seed_handle* h, uint32_t firstPhiBin, uint32_t nPhiBins) {
  B200SEED_FORWARD(h, b200ex_set_phi_sector(h->exact, firstPhiBin, nPhiBins),
                   b200rx_set_phi_sector(h->relaxed, firstPhiBin, nPhiBins));
}

int b200seed_get_stage_times(const b200seed_handle* h, float* ms) {
  B200SEED_FORWARD(h, b200ex_get_stage_times(h->exact, ms), b200rx_get_stage_times(h->relaxed, ms));
}

int b200seed_get_stage_times_ex(const b200seed_handle* h, float* ms, uint32_t n) {
  B200SEED_FORWARD(h, b200ex_get_stage_times_ex(h->exact, ms, n), b200rx_get_stage_times_ex(h->relaxed, ms, n));
}

int b200seed_run_vertices(b200seed_handle* h, uint32_t nSpacePoints, const float* x, const float* y, const float* z,
                          const float* r, const float* varZ, const float* varR, uint32_t nVertices,
                          const double* vertexZ, const double* vertexVarZ, b200seed_seeds* out) {
  B200SEED_FORWARD(h, b200ex_run_vertices(h->exact, nSpacePoints, x, y, z, r, varZ, varR, nVertices, vertexZ, vertexVarZ, out),
                   b200rx_run_vertices(h->relaxed, nSpacePoints, x, y, z, r, varZ, varR, nVertices, vertexZ, vertexVarZ, out));
}

int b200seed_run_strips(b200seed_handle* h, uint32_t nSpacePoints, const float* x, const float* y, const float* z,
                        const float* r, const float* varZ, const float* varR, const float* stripDetails,
                        float cotThetaDiffMax, b200seed_seeds* out) {
  B200SEED_FORWARD(h, b200ex_run_strips(h->exact, nSpacePoints, x, y, z, r, varZ, varR, stripDetails, cotThetaDiffMax, out),
                   b200rx_run_strips(h->relaxed, nSpacePoints, x, y, z, r, varZ, varR, stripDetails, cotThetaDiffMax, out));
}

int b200seed_vertex_windows(const b200seed_handle* h, uint32_t nVertices, const double* vertexZ, const double* vertexVarZ,
                            float* windowLo, float* windowHi) {
  B200SEED_FORWARD(h, b200ex_vertex_windows(h->exact, nVertices, vertexZ, vertexVarZ, windowLo, windowHi),
                   b200rx_vertex_windows(h->relaxed, nVertices, vertexZ, vertexVarZ, windowLo, windowHi));
}

int b200seed_run_batch_windows(b200seed_handle* h, uint32_t nEvents, const uint32_t* spOffsets, const float* x,
                               const float* y, const float* z, const float* r, const float* varZ, const float* varR,
                               const uint32_t* windowOffsets, const float* zWindowLo, const float* zWindowHi,
                               uint64_t* seedOffsets, b200seed_seeds* out) {
  B200SEED_FORWARD(h,
                   b200ex_run_batch_windows(h->exact, nEvents, spOffsets, x, y, z, r, varZ, varR, windowOffsets, zWindowLo,
                                            zWindowHi, seedOffsets, out),
                   b200rx_run_batch_windows(h->relaxed, nEvents, spOffsets, x, y, z, r, varZ, varR, windowOffsets,
                                            zWindowLo, zWindowHi, seedOffsets, out));
}

int b200seed_estimate_params(b200seed_handle* h, uint64_t nSeeds, const uint32_t* bottom, const uint32_t* middle,
                             const uint32_t* top, uint32_t nSpacePoints, const float* x, const float* y,
                             const float* z, const double* bField, double* freeParams) {
  B200SEED_FORWARD(
      h, b200ex_estimate_params(h->exact, nSeeds, bottom, middle, top, nSpacePoints, x, y, z, bField, freeParams),
      b200rx_estimate_params(h->relaxed, nSeeds, bottom, middle, top, nSpacePoints, x, y, z, bField, freeParams));
}

int b200seed_make_pixel_spacepoints(b200seed_handle* h, uint32_t n, const uint32_t* surface, const double* loc0,
                                    const double* loc1, const double* cov00, const double* cov01, const double* cov11,
                                    uint32_t nSurfaces, const double* transforms, float* x, float* y, float* z,
                                    float* r, float* varZ, float* varR) {
  B200SEED_FORWARD(h,
                   b200ex_make_pixel_spacepoints(h->exact, n, surface, loc0, loc1, cov00, cov01, cov11, nSurfaces,
                                                 transforms, x, y, z, r, varZ, varR),
                   b200rx_make_pixel_spacepoints(h->relaxed, n, surface, loc0, loc1, cov00, cov01, cov11, nSurfaces,
                                                 transforms, x, y, z, r, varZ, varR));
}

int b200seed_run_measurements(b200seed_handle* h, uint32_t n, const uint32_t* surface, const double* loc0,
                              const double* loc1, const double* cov00, const double* cov01, const double* cov11,
                              uint32_t nSurfaces, const double* transforms, uint32_t nZWindows, const float* zWindowLo,
                              const float* zWindowHi, float* x, float* y, float* z, float* r, float* varZ, float* varR,
                              b200seed_seeds* out) {
  B200SEED_FORWARD(h,
                   b200ex_run_measurements(h->exact, n, surface, loc0, loc1, cov00, cov01, cov11, nSurfaces, transforms,
                                           nZWindows, zWindowLo, zWindowHi, x, y, z, r, varZ, varR, out),
                   b200rx_run_measurements(h->relaxed, n, surface, loc0, loc1, cov00, cov01, cov11, nSurfaces, transforms,
                                           nZWindows, zWindowLo, zWindowHi, x, y, z, r, varZ, varR, out));
}

int b200seed_debug_grid(b200seed_handle* h, uint64_t capacity, uint32_t* copiedFromIndex, float* x, float* y,
                        float* z, float* r, float* varZ, float* varR, uint64_t binCapacity, uint32_t* binBegin,
                        uint32_t* binEnd) {
  B200SEED_FORWARD(
      h, b200ex_debug_grid(h->exact, capacity, copiedFromIndex, x, y, z, r, varZ, varR, binCapacity, binBegin, binEnd),
      b200rx_debug_grid(h->relaxed, capacity, copiedFromIndex, x, y, z, r, varZ, varR, binCapacity, binBegin, binEnd));
}

int b200seed_debug_doublets(b200seed_handle* h, b200seed_doublets* out) {
  B200SEED_FORWARD(h, b200ex_debug_doublets(h->exact, out), b200rx_debug_doublets(h->relaxed, out));
}

int b200seed_debug_atan2f(b200seed_handle* h, uint64_t n, const float* y, const float* x, float* phi) {
  B200SEED_FORWARD(h, b200ex_debug_atan2f(h->exact, n, y, x, phi), b200rx_debug_atan2f(h->relaxed, n, y, x, phi));
}

}  // extern "C"
