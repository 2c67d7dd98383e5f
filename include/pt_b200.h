/* pt_b200.h — C ABI of the B200-native wavefront integrator.
 *
 * Drop-in boundary for the reference's `Camera::render(&self, world: &World, filename: &str)`
 * (reference src/camera.rs:79), which is the single call every scene function makes
 * (src/main.rs:81,131,235,273,368,531,617).  The reference has no FFI of its own; a host
 * (Rust via a -sys crate, or the C++ mirror in thu-acg-f2024-path-tracer_b200/host) walks
 * its `World` once, fills the closed, flattened `pt_scene_desc` below and calls
 * `pt_scene_create` + `pt_render`.  The host keeps scene construction, asset decoding and
 * the SAH BVH build (src/hittable/bvh.rs:24-120); everything under `Camera::trace`
 * (src/camera.rs:170-228) runs on the GPU.
 *
 * Conventions: plain pointers and sizes, no exceptions across the boundary, every call
 * returns 0 on success or a negative pt_status; pt_last_error() gives the message of the last
 * failure on the calling thread.  All arrays are copied by pt_scene_create (caller keeps
 * ownership).  All geometry is f64, matching the reference (src/vec3.rs:3-6).
 */
#ifndef PT_B200_H
#define PT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_NONE 0xFFFFFFFFu
#define PT_ABI_VERSION 2

typedef enum {
    PT_OK = 0,
    PT_ERR_INVALID = -1,      /* bad argument / malformed scene description */
    PT_ERR_CUDA = -2,         /* CUDA runtime error (message in pt_last_error) */
    PT_ERR_UNSUPPORTED = -3,  /* construct outside the closed device subset */
    PT_ERR_NO_DEVICE = -4     /* no CUDA device: there is no CPU fallback */
} pt_status;

typedef struct { double x, y, z; } pt_vec3;

/* ---- textures: src/texture.rs:11-92 ------------------------------------------------ */
enum { PT_TEX_SOLID = 0, PT_TEX_CHECKER = 1, PT_TEX_IMAGE = 2 };
typedef struct {
    uint32_t kind;
    uint32_t tex1, tex2;   /* CHECKER children (texture indices), texture.rs:27-31 */
    uint32_t image;        /* IMAGE: index into pt_scene_desc.images */
    double   inv_scale;    /* CHECKER: 1/scale, texture.rs:36 */
    pt_vec3  value;        /* SOLID: colour; Texture<f64> keeps the scalar in .x */
} pt_texture;

/* RGB8, row-major, top row first — what image::to_rgb8() yields (texture.rs:62-69). */
typedef struct { const uint8_t* rgb; uint32_t width, height; } pt_image;

/* ---- materials: src/bsdf/*.rs, src/material.rs:150-191 ------------------------------ */
enum {
    PT_MAT_DIFFUSE = 0,    /* bsdf/diffuse.rs   */
    PT_MAT_METAL = 1,      /* bsdf/metal.rs     */
    PT_MAT_GLASS = 2,      /* bsdf/glass.rs     */
    PT_MAT_PRINCIPLED = 3, /* bsdf/principled.rs*/
    PT_MAT_LIGHT = 4,      /* material.rs DiffuseLight */
    PT_MAT_SHEEN = 5,      /* bsdf/sheen.rs     */
    PT_MAT_CLEARCOAT = 6,  /* bsdf/clearcoat.rs */
    PT_MAT_MIX = 7,        /* bsdf/mix.rs       */
    PT_MAT_ISOTROPIC = 8   /* volume.rs:18 phase_function (IsotropicMaterial, stub only): uniform sphere, albedo = base colour */
};
/* indices into pt_material.p */
enum {
    PT_P_METALLIC = 0, PT_P_ROUGHNESS = 1, PT_P_SUBSURFACE = 2, PT_P_SPECULAR = 3,
    PT_P_SPECULAR_TINT = 4, PT_P_IOR = 5, PT_P_SPEC_TRANS = 6, PT_P_SHEEN = 7,
    PT_P_SHEEN_TINT = 8, PT_P_CLEARCOAT = 9, PT_P_CLEARCOAT_GLOSS = 10,
    PT_P_ALPHA_G = 0,      /* CLEARCOAT: alpha_g (clearcoat.rs:16-18) */
    PT_P_MIX_T = 0,        /* MIX: t, already clamped (mix.rs:17) */
    PT_P_COLOR_R = 0, PT_P_COLOR_G = 1, PT_P_COLOR_B = 2 /* SHEEN base colour (sheen.rs:12) */
};
typedef struct {
    uint32_t kind;
    uint32_t base_color_tex; /* DIFFUSE/METAL/GLASS/PRINCIPLED base colour; LIGHT: emission */
    uint32_t roughness_tex;  /* METAL, GLASS: Texture<f64> */
    uint32_t normal_map;     /* DIFFUSE: image index or PT_NONE (diffuse.rs:16) */
    uint32_t mix_a, mix_b;   /* MIX: bxdf1, bxdf2 (material indices) */
    double   p[12];
} pt_material;

/* ---- hittables: src/hittable/*.rs ---------------------------------------------------- */
enum {
    PT_PRIM_SPHERE = 0, PT_PRIM_QUAD = 1, PT_PRIM_TRIANGLE = 2,
    PT_OBJ_CUBOID = 3, PT_OBJ_MESH = 4, PT_OBJ_INSTANCE = 5, PT_OBJ_VOLUME = 6
};
typedef struct { uint32_t kind, index; } pt_ref;

/* Constant-density medium — the reference's src/volume.rs:15-41 is a commented-out stub (`HomogeneousVolume { boundary,
 * negative_inv_density, phase_function }` with `intersects` = todo!()), so the behaviour below is OURS (SURVEY §8(f)-4;
 * oracle-only parity).  For a ray (unit direction) and interval [t_min, inf):
 *   h1 = boundary.intersects(ray, [t_min, inf));  none -> no hit
 *   h1.front_face (entering):  t_in = h1.t;  from just inside, ray' = (ray.at(t_in + 1e-4), same direction):
 *                              h2 = boundary.intersects(ray', [0, inf)),  none -> no hit,  t_out = t_in + 1e-4 + h2.t
 *                              (sphere.rs:80 returns only the near root to an outside origin, hence the re-origin)
 *   otherwise (origin inside): t_in = t_min, t_out = h1.t
 *   s = -ln(U) / density;  s > t_out - t_in -> no hit;  else hit at t = t_in + s, normal (1,0,0), u = v = 0.
 * U is NOT drawn from the path's sequential stream (traversal order must not matter): it is uniform #0 of
 * philox4x32-10(key = seed, counter = (bounce, pixel, sample, 1 + volume index)); pt_trace_closest / pt_trace_any use
 * seed 0, pixel = ray index, sample = bounce = 0.  The phase function is a PT_MAT_ISOTROPIC material: direction uniform
 * on the sphere from 2 uniforms (z = 1 - 2 u1, phi = 2 pi u2), pdf = 1/(4 pi), eval = albedo/(4 pi). */
typedef struct {
    pt_ref boundary;        /* PT_PRIM_SPHERE or PT_OBJ_CUBOID: closed and convex */
    double density;
    uint32_t material;      /* PT_MAT_ISOTROPIC */
    uint32_t _pad;
} pt_volume;

typedef struct {            /* sphere.rs:13-19 */
    pt_vec3 position1, position2;
    double radius;          /* as passed to Sphere::new_*; intersection uses max(0,r) (sphere.rs:26) */
    uint32_t material;
    uint32_t is_moving;     /* built by new_moving (sphere.rs:34): affects only the host-side bbox */
} pt_sphere;

typedef struct {            /* quad.rs:5-36; derived fields computed by the host */
    pt_vec3 q, u, v, w, normal;
    double d;
    uint32_t material, _pad;
} pt_quad;

typedef struct { pt_vec3 v0, v1, v2; } pt_triangle;   /* mesh.rs:13-19, vertices only */

typedef struct {            /* cuboid.rs:5-9: six quads, linear (BVH-less) list */
    uint32_t first_quad;    /* quads[first_quad .. first_quad+6) in cuboid.rs:18-53 order */
    uint32_t material;
    pt_vec3 a, b;           /* arguments of Cuboid::new (cuboid.rs:11), informational */
} pt_cuboid;

typedef struct {            /* mesh.rs:144-198 */
    uint32_t first_triangle, n_triangles;
    uint32_t material;
    uint32_t bvh_root;      /* node index, or PT_NONE => linear scan (list.rs:57-66) */
    uint32_t has_normals;   /* tri_normals[3*t .. 3*t+3) valid (mesh.rs:85-86) */
    uint32_t has_uvs;       /* tri_uvs[6*t .. 6*t+6) valid (mesh.rs:91-98) */
} pt_mesh;

typedef struct {            /* instance.rs:12-31; column-major 4x4 like glam::DMat4 */
    pt_ref child;           /* SPHERE, QUAD, CUBOID, MESH or VOLUME (no nested instances) */
    pt_vec3 axis;           /* arguments of Instance::new (instance.rs:20), informational: */
    double angle;           /*   the device consumes only the three matrices below        */
    pt_vec3 translation;
    double transform[16];
    double inverse[16];     /* transform.inverse(), instance.rs:36 */
    double normal_matrix[16]; /* Mat4::from_quat(rot).inverse().transpose(), instance.rs:45 */
} pt_instance;

/* Host-built BVH (bvh.rs:6-16), consumed as-is: tie-breaks depend on its DFS order. */
typedef struct {
    double bmin[3], bmax[3];
    uint32_t left, right;       /* internal: child node indices; leaf: PT_NONE */
    uint32_t first_ref, n_refs; /* leaf: leaf_refs[first_ref .. first_ref+n_refs) */
} pt_bvh_node;

typedef struct {
    uint32_t abi_version;       /* PT_ABI_VERSION */
    uint32_t n_textures, n_images, n_materials;
    uint32_t n_spheres, n_quads, n_triangles, n_cuboids, n_meshes, n_instances;
    uint32_t n_nodes, n_leaf_refs, n_objects, n_lights;
    const pt_texture*  textures;
    const pt_image*    images;
    const pt_material* materials;
    const pt_sphere*   spheres;
    const pt_quad*     quads;
    const pt_triangle* triangles;
    const pt_vec3*     tri_normals;  /* 3 per triangle, or NULL */
    const double*      tri_uvs;      /* 6 per triangle (u0,v0,u1,v1,u2,v2), or NULL */
    const pt_cuboid*   cuboids;
    const pt_mesh*     meshes;
    const pt_instance* instances;
    const pt_bvh_node* nodes;
    const pt_ref*      leaf_refs;
    const pt_ref*      objects;      /* World.objects in insertion order (world.rs:6) */
    const pt_ref*      lights;       /* World.lights in insertion order (world.rs:7) */
    uint32_t objects_bvh_root;       /* PT_NONE => linear scan */
    uint32_t lights_bvh_root;        /* PT_NONE => linear scan / empty */
    uint32_t n_volumes, _pad;        /* ABI version 2 */
    const pt_volume* volumes;
} pt_scene_desc;

/* ---- camera: public fields of src/camera.rs:23-36; init() (camera.rs:51-77) is
 *      re-derived inside the library ------------------------------------------------- */
typedef struct {
    double aspect_ratio;
    uint32_t image_width;
    uint32_t samples_per_pixel;
    uint32_t max_depth;
    uint32_t env_is_map;        /* 0: EnvironmentType::Color, 1: ::Map (camera.rs:16-19) */
    double vfov;
    pt_vec3 look_from, look_at, vup;
    double blur_strength, focal_length, defocus_angle;
    pt_vec3 env_color;
    uint32_t env_image;         /* image index when env_is_map */
    uint32_t _pad;
} pt_camera;

/* ---- render control --------------------------------------------------------------- */
/* flags bits 4-6: k_trace register-cap variant for 4-wide scenes (0 = default 72 regs; 4: 120, 5: 96, 6: 80, 7: 64) — tuning knob */
/* flags bit 13 (0x2000): run the per-class shade kernels of an iteration one after the other on the render stream instead
 * of forking them onto side streams (tuning / debugging knob; forking is the default and measured 2.5 % faster) */
/* flags bit 14 (0x4000): one host round trip per wavefront iteration even in the tail of a render (default: once nothing
 * is left to generate, 8 iterations are enqueued per round trip) — tuning / debugging knob, same image either way */
/* Traversal flavours (tuning / A-B knobs and test pins; same image either way).  Default: scenes whose World holds at most 32
 * objects + lights run the flat top level (k_top: reference list instead of a BVH, camera rays generated in the same kernel)
 * with mesh visits queued for k_mesh_enter + k_mesh_walk rounds; other scenes run the BVH kernel k_trace, which on scenes with
 * meshes queues mesh visits for k_trace_blas_refill.  flags bit 22 (0x400000): no flat top level (BVH kernels instead);
 * bit 20 (0x100000): no multi-pass traversal at all (one fused BVH kernel); bit 21 (0x200000): with the BVH kernels, walk the
 * queued meshes with plain grid-stride rounds instead of persistent lanes with refill. */
/* flags bit 15 (0x8000): take the octant mask of the survivor grouping from bits 16-18 (bit 16 = y, 17 = x, 18 = z sign of
 * the next ray direction; default 7, 0 = plain compaction) — tuning knob, same image either way */
/* PT_RENDER_ENV_IMPORTANCE (NOT reference behaviour): when the camera's environment is a map, the
 * direction mixture of camera.rs:199-215 gains a third sampler that draws from the map's luminance (built by
 * pt_scene_build_env_sampler): p_bsdf = 0.5, p_env = 0.5 without lights, p_light = p_env = 0.25 with lights; the
 * combined pdf gains p_env * pdf_env.  Same expectation as the reference's estimator, lower variance under
 * concentrated skylight.  Per-sample results differ from the reference's, so it is off by default. */
#define PT_RENDER_ENV_IMPORTANCE 0x1u
/* PT_RENDER_NEE (NOT reference behaviour): next-event estimation with multiple importance sampling instead
 * of the one-sample mixture of camera.rs:199-215.  At every non-emissive hit a direction from World.lights.sample is
 * tested with a shadow ray (a closest-hit ray: it contributes throughput * f * emitted / (pdf_light + pdf_bsdf) if the
 * first thing it meets is an emitter), and the path continues by BSDF sampling alone; emission found by the continued
 * path is weighted pdf_bsdf / (pdf_bsdf + pdf_light) (balance heuristic; 1 for camera rays and for emitters outside
 * World.lights).  Environment hits keep weight 1.  Shadow rays count as segments in pt_stats.  Same expectation as the
 * reference's estimator up to its quirks Q8/Q11; cannot be combined with PT_RENDER_ENV_IMPORTANCE yet. */
#define PT_RENDER_NEE 0x2u
enum { PT_NAN_REFERENCE = 0, /* non-finite samples poison the pixel like camera.rs:129 */
       PT_NAN_DROP = 1 };    /* documented divergence: a non-finite contribution is skipped and a non-finite
                                throughput ends the path (each counted once in pt_stats.nonfinite); finite
                                contributions that sample made earlier — e.g. a directly seen light — stay */
typedef struct {
    uint64_t seed;
    uint32_t sample_begin;   /* first sample index of this call (per pixel) */
    uint32_t sample_count;   /* samples this call renders per pixel */
    uint32_t sample_stride;  /* sample index = sample_begin + i*stride (multi-GPU spp split) */
    uint32_t nan_policy;
    uint32_t pool_paths;     /* in-flight path pool size; 0 = default */
    uint32_t flags;          /* PT_FLAG_* */
} pt_render_params;

typedef struct {
    uint64_t paths;          /* Camera::trace calls */
    uint64_t segments;       /* World::intersect_all calls (camera.rs:179) = "rays" */
    uint64_t nonfinite;      /* samples dropped/poisoned */
    uint64_t kernel_launches;
    uint32_t iterations;     /* wavefront iterations */
    uint32_t width, height;
    float    device_ms;      /* CUDA-event time of the render loop */
    float    trace_ms, shade_ms, raygen_ms; /* per-stage totals when profiling enabled, else 0 */
    /* traversal work summed over all segments (profiling mode only, else 0): 64-byte node pairs fetched,
     * 32-byte reference boxes fetched, f64 primitive tests */
    uint64_t node_pairs, ref_boxes, prim_tests;
    uint32_t two_pass_iterations; /* wavefront iterations whose traversal stage took the two-pass path (top-level pass +
                                   * mesh rounds) instead of the fused kernel: tests assert on it so that an ID comparison
                                   * can never silently run on the other flavour */
    uint32_t queue_errors;        /* pt_trace_*_wavefront only: rays that did not join exactly one shade-class queue, or joined the
                                   * queue of another class than their hit's material (an invariant of the traversal stage; 0) */
    uint32_t p2p_shares;          /* pt_render_multi only: shares whose radiance sums the reduce kernel read in place over peer access */
    uint32_t tail_paths;          /* paths the tail megakernel ran to their end in one launch (0: the render never took that path) */
} pt_stats;

typedef struct pt_ctx pt_ctx;
typedef struct pt_scene pt_scene;

/* ---- lifecycle -------------------------------------------------------------------- */
int  pt_ctx_create(int device, pt_ctx** out);
void pt_ctx_destroy(pt_ctx* ctx);
/* Use an external stream (e.g. torch's current stream handle); NULL = the ctx's own. */
int  pt_ctx_set_stream(pt_ctx* ctx, void* cuda_stream);
/* 0: off.  1: per-stage CUDA events around every launch (pt_stats.trace_ms / shade_ms / raygen_ms; the shade kernels then run
 * unforked and the tail unbatched).  2: additionally the traversal work counters in pt_stats; slower kernel variants. */
int  pt_ctx_set_profiling(pt_ctx* ctx, int level);
const char* pt_last_error(void);
int  pt_device_count(void);

int  pt_scene_create(pt_ctx* ctx, const pt_scene_desc* desc, pt_scene** out);
void pt_scene_destroy(pt_scene* scene);
/* bytes uploaded host->device by pt_scene_create */
uint64_t pt_scene_device_bytes(const pt_scene* scene);
/* Builds the importance sampler of environment image `image` (a lat-long map read as camera.rs:140-151 reads it):
 * a piecewise-constant density over min(height, max_rows) x min(width, max_cols) cells (0 = 512 x 1024), cell weight =
 * sum of texel luminance * sin(theta) + a 5 % uniform floor.  Required before rendering with PT_RENDER_ENV_IMPORTANCE. */
int  pt_scene_build_env_sampler(pt_scene* scene, uint32_t image, uint32_t max_rows, uint32_t max_cols);

/* ---- Camera::render replacement (camera.rs:79-126 minus the PNG encode) ------------ */
/* image height as Camera::init computes it (camera.rs:52) */
uint32_t pt_camera_image_height(const pt_camera* cam);
/* Adds the SUM of radiance samples into d_accum (device pointer, W*H*3 fp32, caller zeroes).
 * This is the multi-GPU building block: each rank renders its sample subset, then one
 * reduce(sum) of d_accum over NCCL. */
int  pt_render_accumulate(pt_ctx* ctx, const pt_scene* scene, const pt_camera* cam,
                          const pt_render_params* params, float* d_accum, pt_stats* stats);
/* Host-buffer convenience = the e2e call: renders params->sample_count samples and writes the
 * mean radiance (W*H*3 fp32, row-major, top row first) to host memory. */
int  pt_render(pt_ctx* ctx, const pt_scene* scene, const pt_camera* cam,
               const pt_render_params* params, float* h_mean_rgb, pt_stats* stats);
/* Camera::render on several GPUs from ONE process — what a single-process host like the reference's binary calls
 * (camera.rs:79; the reference parallelises over pixels with rayon inside that call, camera.rs:102).  One host thread
 * per entry of `devices` creates its own context, uploads `desc`, and renders the samples g, g + n, g + 2n, ... of the
 * call (spp split, independent counter-based streams).  The reduce(sum) runs ON THE DEVICE: one kernel on
 * devices[0] reads every share's fp32 sums in place — over NVLink peer access for the other GPUs (staged by one
 * cudaMemcpyPeer where two devices cannot address each other) — adds them in share order in f64, divides by sample_count
 * and the mean image leaves the GPU in a single D2H copy.  Same image as pt_render on one device up to fp32 summation
 * order.  The same device may be listed more than once.  stats: counters summed, times = the slowest device.  bench.py
 * and the tests' multi-GPU path use one process per GPU + NCCL instead (pt_render_accumulate). */
int  pt_render_multi(int n_devices, const int* devices, const pt_scene_desc* desc, const pt_camera* cam,
                     const pt_render_params* params, float* h_mean_rgb, pt_stats* stats);
/* pt_render_multi keeps its per-device contexts (stream, events, the path pool) between calls; this frees them. */
void pt_render_multi_release(void);
/* sqrt-gamma, clamp(0,0.999)*256 as u8 (camera.rs:109-114,128-130). Device in, host out. */
int  pt_tonemap_rgb8(pt_ctx* ctx, const float* d_accum, double scale, uint32_t n_pixels,
                     uint8_t* h_rgb8);

/* ---- parity/test entry points ------------------------------------------------------ */
typedef struct { pt_vec3 origin, direction; double time; } pt_ray; /* as built by Ray::new (ray.rs:23-29) */
typedef struct {
    double t, u, v;
    pt_vec3 point, geometric_normal, shading_normal;
    uint32_t hit;          /* 0 = miss */
    uint32_t prim_kind;    /* PT_PRIM_* */
    uint32_t prim_index;   /* index into spheres/quads/triangles */
    uint32_t instance;     /* instance index or PT_NONE */
    uint32_t material;
    uint32_t front_face;
    uint32_t is_light;     /* came from World.lights (world.rs:47-62) */
    uint32_t work;         /* diagnostics: node pairs fetched | primitive tests << 16 (saturating); oracle: 0 */
} pt_hit;
/* World::intersect_all(ray, [t_min, inf)) for a batch of host rays. */
int  pt_trace_closest(pt_ctx* ctx, const pt_scene* scene, size_t n, const pt_ray* rays,
                      double t_min, pt_hit* hits);
/* The same query through the RENDER's traversal stage: the batch is loaded into the wavefront path pool and traced by
 * exactly the kernels, grids and queues pt_render_accumulate launches for an iteration of that many live paths (two-pass
 * traversal on scenes with meshes from 65 536 rays up, the fused kernel below that or with flags bit 20), in chunks of
 * 4 Mi rays.  `flags`: the traversal knobs of pt_render_params.flags (bits 4-6, 20, 21).  Same result as pt_trace_closest
 * (`work` = 0); stats->two_pass_iterations tells which flavour ran.  This is the entry the ID-exactness tests use, so the
 * benchmarked traversal is the tested one. */
int  pt_trace_closest_wavefront(pt_ctx* ctx, const pt_scene* scene, size_t n, const pt_ray* rays,
                                double t_min, uint32_t flags, pt_hit* hits, pt_stats* stats);
/* The first wavefront iteration of a render, as a query: sample `sample` of every pixel — the camera ray
 * (Camera::generate_ray, camera.rs:153-168, from the library's counter-based RNG with `seed`) and its closest hit on
 * [1e-3, inf) — produced by the render's own start-of-path traversal stage (on scenes with a small World the top-level kernel
 * generates the ray in registers and traces it in the same launch; this entry is how that fused kernel is ID-compared).
 * rays / hits: W*H entries, row-major by pixel.  `flags` as in pt_trace_closest_wavefront. */
int  pt_trace_camera_wavefront(pt_ctx* ctx, const pt_scene* scene, const pt_camera* cam, uint64_t seed, uint32_t sample,
                               uint32_t flags, pt_ray* rays, pt_hit* hits, pt_stats* stats);
/* World::shadow_ray-style any-hit against World.objects on [t_min, t_max] (world.rs:31-36). */
int  pt_trace_any(pt_ctx* ctx, const pt_scene* scene, size_t n, const pt_ray* rays,
                  double t_min, const double* t_max, uint8_t* occluded);

typedef struct {
    pt_vec3 view_dir, light_dir;         /* world space; view_dir = -ray.direction */
    pt_vec3 point, geometric_normal, shading_normal;
    double u, v;
    uint32_t front_face, _pad;
} pt_bsdf_query;
typedef struct { pt_vec3 eval; double pdf; pt_vec3 emitted; double _pad; } pt_bsdf_result;
/* BxDFMaterial::{eval,pdf,emitted} (bsdf/mod.rs:21-57) */
int  pt_bsdf_eval_pdf(pt_ctx* ctx, const pt_scene* scene, uint32_t material, size_t n,
                      const pt_bsdf_query* q, pt_bsdf_result* out);
/* BxDFMaterial::sample with explicit uniforms (8 per query, consumed in reference order). */
typedef struct { pt_vec3 dir; uint32_t valid; uint32_t n_uniforms; } pt_bsdf_sample_result;
int  pt_bsdf_sample(pt_ctx* ctx, const pt_scene* scene, uint32_t material, size_t n,
                    const pt_bsdf_query* q, const double* uniforms8, pt_bsdf_sample_result* out);
/* Camera::generate_ray for explicit (pixel row, col, sample) triples with the library's
 * counter-based RNG (Philox4x32-10). */
int  pt_camera_rays(pt_ctx* ctx, const pt_camera* cam, uint64_t seed, size_t n,
                    const uint32_t* row, const uint32_t* col, const uint32_t* sample,
                    pt_ray* out);
/* World.lights.{sample,pdf} (list.rs:78-96): sample uses up to 4 uniforms per query (light pick, side/triangle pick, 2 on the surface). */
int  pt_lights_sample_pdf(pt_ctx* ctx, const pt_scene* scene, size_t n, const pt_vec3* origin,
                          const double* time, const double* uniforms4, pt_vec3* dir,
                          uint32_t* valid, double* pdf);
/* The step before the path: the exact SAH sweep of bvh.rs:54-120 for ONE node on the device.  For the
 * n items of the node in list order (boxes6 = lo xyz, hi xyz per item, padded as AABB::new leaves them) and every axis a,
 * cost3n[a * n + k] = BVH::evaluate_sah at item k's centroid[a] — the same sequential padded unions as the reference
 * (aabb.rs:16-25), so costs are bit-identical to the host's — or +inf where the reference rejects the split.  The host
 * keeps the recursion, the strict-'<' choice and the partition (bvh.rs:28-84); this only removes the O(n^2) part. */
int  pt_sah_sweep(pt_ctx* ctx, uint32_t n, const double* boxes6, const double* parent6, double* cost3n);
/* The environment sampler of pt_scene_build_env_sampler: direction from 2 uniforms per query and its solid-angle pdf. */
int  pt_env_sample_pdf(pt_ctx* ctx, const pt_scene* scene, size_t n, const double* uniforms2, pt_vec3* dir, double* pdf);

/* Self-check of the device's shared-reciprocal vector division (csrc/device_scene.cuh: div3_shared, what `DVec3 / f64` compiles to in
 * the shade kernels) against the plain IEEE `/` operator on n pseudo-random operand triples, bit for bit; *mismatches = differing quotients. */
int  pt_debug_div_check(pt_ctx* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches);
/* Diagnostics (profiling level 2 only): eight 64-bin histograms of per-ray / per-mesh-visit traversal work gathered by the
 * counting kernel variants since the last reset (layout: csrc/kernels.cuh, g_hist).  out512 may be NULL. */
int  pt_debug_histograms(pt_ctx* ctx, uint64_t* out512, int reset);
/* Diagnostics (profiling level >= 1): CUDA-event time per kernel family of the traversal stage since the last reset, ms:
 * [0] k_top on surviving paths, [1] k_top on new paths (camera rays generated in the kernel), [2] k_mesh_enter,
 * [3] k_mesh_walk, [4] BVH trace kernels (k_trace, k_trace_blas*), [5] k_generate, [6] k_tail (the tail megakernel).  out16 may be NULL. */
int  pt_debug_stage_ms(pt_ctx* ctx, double* out16, int reset);

#ifdef __cplusplus
}
#endif
#endif /* PT_B200_H */
