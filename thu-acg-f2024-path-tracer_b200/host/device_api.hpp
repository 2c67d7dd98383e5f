// Late binding of the CUDA library (lib/libptb200.so, include/pt_b200.h) from the host mirror.
//
// The host library builds and flattens scenes without any device code; only Camera::render and the optional device SAH
// sweep need the GPU.  libptb200.so is therefore NOT a link-time dependency: it is opened from this library's own directory
// the first time a device entry point is called.  A process that only builds scene descriptions (bench.py's CPU reference
// arm, the oracle tests) never maps the CUDA library.
#pragma once
#include "../../include/pt_b200.h"

namespace pt {
struct DeviceApi {
    int (*ctx_create)(int, pt_ctx**);
    void (*ctx_destroy)(pt_ctx*);
    const char* (*last_error)(void);
    int (*scene_create)(pt_ctx*, const pt_scene_desc*, pt_scene**);
    void (*scene_destroy)(pt_scene*);
    int (*scene_build_env_sampler)(pt_scene*, uint32_t, uint32_t, uint32_t);
    int (*render)(pt_ctx*, const pt_scene*, const pt_camera*, const pt_render_params*, float*, pt_stats*);
    int (*render_multi)(int, const int*, const pt_scene_desc*, const pt_camera*, const pt_render_params*, float*, pt_stats*);
    int (*sah_sweep)(pt_ctx*, uint32_t, const double*, const double*, double*);
};
// Throws std::runtime_error when the library or one of its symbols is missing (there is no CPU fallback).
const DeviceApi& device_api();
}  // namespace pt
