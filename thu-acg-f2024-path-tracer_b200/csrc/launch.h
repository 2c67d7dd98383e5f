// Host-callable launchers of the sm_100a kernels.  The kernels live in four translation units that are compiled in
// parallel (trace.cu, shade_ref.cu, shade_nol.cu, shade_var.cu, shade_nee.cu, tail_wide.cu, tail_bin.cu, tail_flat.cu, misc.cu); api.cu holds no device code.
#pragma once
#include "wavefront.cuh"

namespace ptd {

// ---- trace.cu: World::intersect_all (world.rs:47-62) for the path pool and for host ray batches
struct TraceFlavour { int min_blocks; bool wide, count, vol, defer; };
void run_k_trace(const TraceFlavour& f, unsigned grid, cudaStream_t st, PathBuf in, uint32_t n, HitRec* hits, Queues q, const DScene& S,
                 unsigned long long* work, uint64_t seed, const uint32_t* n_dev, BlasQueues bq, uint2* ties, double t_min);
void run_k_trace_blas(bool refill, bool count, unsigned grid, cudaStream_t st, PathBuf in, uint32_t round, BlasQueues bq, HitRec* hits, uint2* ties,
                      Queues q, const DScene& S, unsigned long long* work, double t_min);
// flat top level + mesh rounds (trace_kernels.cuh)
void run_k_top(bool primary, bool count, cudaStream_t st, PathBuf pool, uint32_t slot0, uint32_t n, HitRec* hits, Queues q, const DScene& S, const TopList& top,
               MeshQueues mq, uint2* ties, double t_min, const GenArgs& gen, const uint32_t* n_dev, unsigned long long* work);
void run_k_mesh_enter(bool count, unsigned grid, cudaStream_t st, PathBuf pool, uint32_t round, MeshQueues mq, const HitRec* hits, const uint2* ties, Queues q,
                      const DScene& S, const TopList& top, double t_min, unsigned long long* work);
void run_k_mesh_walk(bool count, unsigned grid, cudaStream_t st, uint32_t round, MeshQueues mq, HitRec* hits, uint2* ties, Queues q, const DScene& S, double t_min,
                     unsigned long long* work);
void run_k_mesh_multi(bool count, unsigned grid, cudaStream_t st, PathBuf pool, MeshQueues mq, HitRec* hits, const uint2* ties, Queues q, const DScene& S, double t_min,
                      unsigned long long* work);
constexpr bool kMeshMulti = PT_MESH_MULTI != 0;
unsigned mesh_walk_resident_warps();  // persistent grid of k_mesh_walk: one warp per resident slot
void run_k_trace_batch(bool wide, cudaStream_t st, const pt_ray* rays, size_t n, double t_min, pt_hit* out, const DScene& S);
void run_k_trace_any_batch(bool wide, cudaStream_t st, const pt_ray* rays, size_t n, double t_min, const double* t_max, uint8_t* out, const DScene& S);
void run_k_rays_to_pool(cudaStream_t st, const pt_ray* rays, uint32_t n, uint32_t first, PathBuf out);
void run_k_hits_to_abi(cudaStream_t st, const pt_ray* rays, uint32_t n, const HitRec* hits, pt_hit* out, const DScene& S);
void run_k_pool_to_abi(cudaStream_t st, PathBuf pool, uint32_t n, const HitRec* hits, pt_ray* out_rays, pt_hit* out_hits, const DScene& S);
void run_k_check_queues(cudaStream_t st, Queues q, const HitRec* hits, uint32_t n, uint32_t* seen, uint32_t* errors, const DScene& S);
cudaError_t debug_histograms(unsigned long long* out512, bool reset);

// ---- shade_*.cu: one loop iteration of Camera::trace after intersect_all (camera.rs:180-225) per shade class
// var: 0 = reference mixture with quad / sphere lights, 2 = + cuboid / mesh / instance lights, 3 = + environment sampler
struct ShadeArgs {
    PathBuf in; Queues q; const HitRec* hits; PathBuf out; uint32_t* out_count; float* accum; unsigned long long* nonfinite;
    DScene S; DCameraEx cam; RenderConst rc;
};
void run_k_shade_ref(int cls, unsigned grid, cudaStream_t st, const ShadeArgs& a);            // shade_ref.cu (var 0)
void run_k_shade_nolights(int cls, unsigned grid, cudaStream_t st, const ShadeArgs& a);       // shade_nol.cu (var 4: World.lights is empty)
void run_k_shade_var(int cls, int var, unsigned grid, cudaStream_t st, const ShadeArgs& a);   // shade_var.cu (var 2, 3)
void run_k_shade_nee(int cls, unsigned grid, cudaStream_t st, const ShadeArgs& a);            // shade_nee.cu (PT_RENDER_NEE)

// ---- tail_wide.cu / tail_bin.cu: the tail megakernel (tail_kernels.cuh): every path of a small wavefront runs to its end in one launch
// counters: [0] += segments traced, [1] = max segments of one path (= the wavefront iterations the launch replaces)
struct TailArgs {
    PathBuf in; uint32_t n; float* accum; unsigned long long* nonfinite; DScene S; DCameraEx cam; RenderConst rc; double t_min; uint32_t* counters;
};
void run_k_tail_wide(cudaStream_t st, const TailArgs& a);
void run_k_tail_bin(cudaStream_t st, const TailArgs& a);
void run_k_tail_flat(cudaStream_t st, const TailArgs& a, const TopList& top);  // tail_flat.cu: flat scenes
void preload_k_tail_wide();
void preload_k_tail_bin();
void preload_k_tail_flat();

// ---- misc.cu: ray generation, tonemap, parity entry kernels
void run_k_generate(cudaStream_t st, PathBuf out, uint32_t slot0, uint32_t n_new, uint64_t g0, uint32_t n_pixels, const DCameraEx& cam, const RenderConst& rc);
void run_k_reduce_peers(cudaStream_t st, const float* const* src, uint32_t n_src, double scale, size_t n_values, float* out);
void run_k_scale(cudaStream_t st, const float* accum, float scale, uint32_t n_values, float* out);
void run_k_tonemap(cudaStream_t st, const float* accum, double scale, uint32_t n_values, uint8_t* out);
void run_k_bsdf_eval(cudaStream_t st, uint32_t material, size_t n, const pt_bsdf_query* q, pt_bsdf_result* out, const DScene& S);
void run_k_bsdf_sample(cudaStream_t st, uint32_t material, size_t n, const pt_bsdf_query* q, const double* uniforms8, pt_bsdf_sample_result* out, const DScene& S);
void run_k_camera_rays(cudaStream_t st, const DCameraEx& cam, uint64_t seed, size_t n, const uint32_t* row, const uint32_t* col, const uint32_t* sample, pt_ray* out);
void run_k_lights(cudaStream_t st, size_t n, const pt_vec3* origin, const double* time, const double* uniforms4, pt_vec3* dir, uint32_t* valid, double* pdf, const DScene& S);
void run_k_sah_sweep(cudaStream_t st, uint32_t n, const SahBox* boxes, const SahBox& parent, double* cost);
void run_k_div_check(cudaStream_t st, uint64_t n, uint64_t seed, unsigned long long* mismatches);
void run_k_env(cudaStream_t st, size_t n, const double* uniforms2, pt_vec3* dir, double* pdf, const DEnvDist& E);

}  // namespace ptd
