// engine_symbols.h -- pre-included (nvcc -include) into the translation units of ONE engine.
//
// The library holds two engines compiled from the same sources (seeding_plugin.cu +
// host_plan.cpp): the exact binary32 one and the relaxedFloat fast path.  Each is built
// with -DB200SEED_ENGINE_PREFIX=<prefix> so that its entry points (and its handle type)
// get private names; seeding_abi.cpp exports the public C ABI of
// include/acts_b200_seeding.h and forwards every call to the engine a handle was created with.
#pragma once

#ifdef B200SEED_ENGINE_PREFIX
#define B200SEED_PASTE2(a, b) a##b
#define B200SEED_PASTE(a, b) B200SEED_PASTE2(a, b)
#define B200SEED_E(name) B200SEED_PASTE(B200SEED_ENGINE_PREFIX, name)

#define b200seed_handle B200SEED_E(handle)
#define b200seed_config_init B200SEED_E(config_init)
#define b200seed_plan_info B200SEED_E(plan_info)
#define b200seed_plan_tables B200SEED_E(plan_tables)
#define b200seed_create B200SEED_E(create)
#define b200seed_create_orthogonal B200SEED_E(create_orthogonal)
#define b200seed_orthogonal_config_init B200SEED_E(orthogonal_config_init)
#define b200seed_destroy B200SEED_E(destroy)
#define b200seed_last_error B200SEED_E(last_error)
#define b200seed_alloc_pinned B200SEED_E(alloc_pinned)
#define b200seed_free_pinned B200SEED_E(free_pinned)
#define b200seed_get_info B200SEED_E(get_info)
#define b200seed_get_counters B200SEED_E(get_counters)
#define b200seed_run B200SEED_E(run)
#define b200seed_run_with_phi B200SEED_E(run_with_phi)
#define b200seed_run_batch B200SEED_E(run_batch)
#define b200seed_run_batch_device B200SEED_E(run_batch_device)
#define b200seed_sync B200SEED_E(sync)
#define b200seed_set_phi_sector B200SEED_E(set_phi_sector)
#define b200seed_get_stage_times B200SEED_E(get_stage_times)
#define b200seed_get_stage_times_ex B200SEED_E(get_stage_times_ex)
#define b200seed_run_vertices B200SEED_E(run_vertices)
#define b200seed_run_strips B200SEED_E(run_strips)
#define b200seed_vertex_windows B200SEED_E(vertex_windows)
#define b200seed_run_batch_windows B200SEED_E(run_batch_windows)
#define b200seed_estimate_params B200SEED_E(estimate_params)
#define b200seed_make_pixel_spacepoints B200SEED_E(make_pixel_spacepoints)
#define b200seed_run_measurements B200SEED_E(run_measurements)
#define b200seed_debug_grid B200SEED_E(debug_grid)
#define b200seed_debug_doublets B200SEED_E(debug_doublets)
#define b200seed_debug_atan2f B200SEED_E(debug_atan2f)
#endif
