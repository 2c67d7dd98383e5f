// Translation unit of ray generation, tonemap and the parity / test entry kernels (misc_kernels.cuh).
#include <algorithm>

#include "launch.h"
#include "misc_kernels.cuh"

namespace ptd {
static unsigned grid128(size_t n) { return (unsigned)((n + 127) / 128); }
void run_k_generate(cudaStream_t st, PathBuf out, uint32_t slot0, uint32_t n_new, uint64_t g0, uint32_t n_pixels, const DCameraEx& cam, const RenderConst& rc) {
    k_generate<<<(n_new + kBlock - 1) / kBlock, kBlock, 0, st>>>(out, slot0, n_new, g0, n_pixels, cam, rc);
}
void run_k_reduce_peers(cudaStream_t st, const float* const* src, uint32_t n_src, double scale, size_t n_values, float* out) {
    const unsigned grid = (unsigned)std::min<size_t>((n_values / 4 + 255) / 256 + 1, 148u * 8u);
    k_reduce_peers<<<grid, 256, 0, st>>>(src, n_src, scale, n_values, out);
}
void run_k_scale(cudaStream_t st, const float* accum, float scale, uint32_t n_values, float* out) { k_scale<<<(n_values + 255) / 256, 256, 0, st>>>(accum, scale, n_values, out); }
void run_k_tonemap(cudaStream_t st, const float* accum, double scale, uint32_t n_values, uint8_t* out) { k_tonemap<<<(n_values + 255) / 256, 256, 0, st>>>(accum, scale, n_values, out); }
void run_k_bsdf_eval(cudaStream_t st, uint32_t material, size_t n, const pt_bsdf_query* q, pt_bsdf_result* out, const DScene& S) { k_bsdf_eval<<<grid128(n), 128, 0, st>>>(material, n, q, out, S); }
void run_k_bsdf_sample(cudaStream_t st, uint32_t material, size_t n, const pt_bsdf_query* q, const double* uniforms8, pt_bsdf_sample_result* out, const DScene& S) {
    k_bsdf_sample<<<grid128(n), 128, 0, st>>>(material, n, q, uniforms8, out, S);
}
void run_k_camera_rays(cudaStream_t st, const DCameraEx& cam, uint64_t seed, size_t n, const uint32_t* row, const uint32_t* col, const uint32_t* sample, pt_ray* out) {
    k_camera_rays<<<grid128(n), 128, 0, st>>>(cam, seed, n, row, col, sample, out);
}
void run_k_lights(cudaStream_t st, size_t n, const pt_vec3* origin, const double* time, const double* uniforms4, pt_vec3* dir, uint32_t* valid, double* pdf, const DScene& S) {
    k_lights<<<grid128(n), 128, 0, st>>>(n, origin, time, uniforms4, dir, valid, pdf, S);
}
void run_k_sah_sweep(cudaStream_t st, uint32_t n, const SahBox* boxes, const SahBox& parent, double* cost) {
    k_sah_sweep<<<(3u * n + kSahTile - 1) / kSahTile, kSahTile, 0, st>>>(n, boxes, parent, cost);
}
void run_k_env(cudaStream_t st, size_t n, const double* uniforms2, pt_vec3* dir, double* pdf, const DEnvDist& E) { k_env<<<grid128(n), 128, 0, st>>>(n, uniforms2, dir, pdf, E); }
void run_k_div_check(cudaStream_t st, uint64_t n, uint64_t seed, unsigned long long* mismatches) { k_div_check<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, seed, mismatches); }
}  // namespace ptd
