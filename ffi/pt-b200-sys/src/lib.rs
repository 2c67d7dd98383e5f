//! Raw FFI declarations for `include/pt_b200.h` (PT_ABI_VERSION 2).  Field order and sizes mirror the C structs one for one
//! (`tests/test_abi_and_host.py` pins the C side's sizes; the `size_of` assertions at the bottom pin this side).
//! NOT compiled in the build image of this repository (no Rust toolchain there): this is the source a maintainer of the
//! reference adds as a path dependency.  The executable versions of the same boundary are the C++ host mirror
//! (`thu-acg-f2024-path-tracer_b200/host`) and the ctypes layer (`thu-acg-f2024-path-tracer_b200/__init__.py`).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const PT_NONE: u32 = 0xFFFF_FFFF;
pub const PT_ABI_VERSION: u32 = 2;
pub const PT_OK: c_int = 0;
pub const PT_ERR_INVALID: c_int = -1;
pub const PT_ERR_CUDA: c_int = -2;
pub const PT_ERR_UNSUPPORTED: c_int = -3;
pub const PT_ERR_NO_DEVICE: c_int = -4;

// texture kinds, material kinds, parameter slots, hittable kinds (see the header for the reference lines)
pub const PT_TEX_SOLID: u32 = 0; pub const PT_TEX_CHECKER: u32 = 1; pub const PT_TEX_IMAGE: u32 = 2;
pub const PT_MAT_DIFFUSE: u32 = 0; pub const PT_MAT_METAL: u32 = 1; pub const PT_MAT_GLASS: u32 = 2; pub const PT_MAT_PRINCIPLED: u32 = 3;
pub const PT_MAT_LIGHT: u32 = 4; pub const PT_MAT_SHEEN: u32 = 5; pub const PT_MAT_CLEARCOAT: u32 = 6; pub const PT_MAT_MIX: u32 = 7;
pub const PT_MAT_ISOTROPIC: u32 = 8;
pub const PT_P_METALLIC: usize = 0; pub const PT_P_ROUGHNESS: usize = 1; pub const PT_P_SUBSURFACE: usize = 2; pub const PT_P_SPECULAR: usize = 3;
pub const PT_P_SPECULAR_TINT: usize = 4; pub const PT_P_IOR: usize = 5; pub const PT_P_SPEC_TRANS: usize = 6; pub const PT_P_SHEEN: usize = 7;
pub const PT_P_SHEEN_TINT: usize = 8; pub const PT_P_CLEARCOAT: usize = 9; pub const PT_P_CLEARCOAT_GLOSS: usize = 10;
pub const PT_P_ALPHA_G: usize = 0; pub const PT_P_MIX_T: usize = 0;
pub const PT_PRIM_SPHERE: u32 = 0; pub const PT_PRIM_QUAD: u32 = 1; pub const PT_PRIM_TRIANGLE: u32 = 2; pub const PT_OBJ_CUBOID: u32 = 3;
pub const PT_OBJ_MESH: u32 = 4; pub const PT_OBJ_INSTANCE: u32 = 5; pub const PT_OBJ_VOLUME: u32 = 6;
pub const PT_NAN_REFERENCE: u32 = 0; pub const PT_NAN_DROP: u32 = 1;
pub const PT_RENDER_ENV_IMPORTANCE: u32 = 0x1; pub const PT_RENDER_NEE: u32 = 0x2;

#[repr(C)] #[derive(Clone, Copy, Default, Debug, PartialEq)] pub struct pt_vec3 { pub x: f64, pub y: f64, pub z: f64 }
#[repr(C)] #[derive(Clone, Copy, Debug, PartialEq, Eq)] pub struct pt_ref { pub kind: u32, pub index: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_texture { pub kind: u32, pub tex1: u32, pub tex2: u32, pub image: u32, pub inv_scale: f64, pub value: pt_vec3 }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_image { pub rgb: *const u8, pub width: u32, pub height: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_material { pub kind: u32, pub base_color_tex: u32, pub roughness_tex: u32, pub normal_map: u32,
                                                          pub mix_a: u32, pub mix_b: u32, pub p: [f64; 12] }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_volume { pub boundary: pt_ref, pub density: f64, pub material: u32, pub _pad: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_sphere { pub position1: pt_vec3, pub position2: pt_vec3, pub radius: f64, pub material: u32, pub is_moving: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_quad { pub q: pt_vec3, pub u: pt_vec3, pub v: pt_vec3, pub w: pt_vec3, pub normal: pt_vec3, pub d: f64,
                                                      pub material: u32, pub _pad: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_triangle { pub v0: pt_vec3, pub v1: pt_vec3, pub v2: pt_vec3 }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_cuboid { pub first_quad: u32, pub material: u32, pub a: pt_vec3, pub b: pt_vec3 }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_mesh { pub first_triangle: u32, pub n_triangles: u32, pub material: u32, pub bvh_root: u32,
                                                      pub has_normals: u32, pub has_uvs: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_instance { pub child: pt_ref, pub axis: pt_vec3, pub angle: f64, pub translation: pt_vec3,
                                                          pub transform: [f64; 16], pub inverse: [f64; 16], pub normal_matrix: [f64; 16] }
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_bvh_node { pub bmin: [f64; 3], pub bmax: [f64; 3], pub left: u32, pub right: u32, pub first_ref: u32, pub n_refs: u32 }
#[repr(C)] pub struct pt_scene_desc {
    pub abi_version: u32, pub n_textures: u32, pub n_images: u32, pub n_materials: u32, pub n_spheres: u32, pub n_quads: u32,
    pub n_triangles: u32, pub n_cuboids: u32, pub n_meshes: u32, pub n_instances: u32, pub n_nodes: u32, pub n_leaf_refs: u32,
    pub n_objects: u32, pub n_lights: u32,
    pub textures: *const pt_texture, pub images: *const pt_image, pub materials: *const pt_material, pub spheres: *const pt_sphere,
    pub quads: *const pt_quad, pub triangles: *const pt_triangle, pub tri_normals: *const pt_vec3, pub tri_uvs: *const f64,
    pub cuboids: *const pt_cuboid, pub meshes: *const pt_mesh, pub instances: *const pt_instance, pub nodes: *const pt_bvh_node,
    pub leaf_refs: *const pt_ref, pub objects: *const pt_ref, pub lights: *const pt_ref,
    pub objects_bvh_root: u32, pub lights_bvh_root: u32,
    pub n_volumes: u32, pub _pad: u32, pub volumes: *const pt_volume,
}
#[repr(C)] #[derive(Clone, Copy)] pub struct pt_camera { pub aspect_ratio: f64, pub image_width: u32, pub samples_per_pixel: u32, pub max_depth: u32,
    pub env_is_map: u32, pub vfov: f64, pub look_from: pt_vec3, pub look_at: pt_vec3, pub vup: pt_vec3, pub blur_strength: f64,
    pub focal_length: f64, pub defocus_angle: f64, pub env_color: pt_vec3, pub env_image: u32, pub _pad: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct pt_render_params { pub seed: u64, pub sample_begin: u32, pub sample_count: u32, pub sample_stride: u32,
    pub nan_policy: u32, pub pool_paths: u32, pub flags: u32 }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)] pub struct pt_stats { pub paths: u64, pub segments: u64, pub nonfinite: u64, pub kernel_launches: u64,
    pub iterations: u32, pub width: u32, pub height: u32, pub device_ms: f32, pub trace_ms: f32, pub shade_ms: f32, pub raygen_ms: f32,
    pub node_pairs: u64, pub ref_boxes: u64, pub prim_tests: u64,
    pub two_pass_iterations: u32, pub queue_errors: u32, pub p2p_shares: u32, pub tail_paths: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct pt_ray { pub origin: pt_vec3, pub direction: pt_vec3, pub time: f64 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct pt_hit { pub t: f64, pub u: f64, pub v: f64, pub point: pt_vec3, pub geometric_normal: pt_vec3,
    pub shading_normal: pt_vec3, pub hit: u32, pub prim_kind: u32, pub prim_index: u32, pub instance: u32, pub material: u32,
    pub front_face: u32, pub is_light: u32, pub work: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct pt_bsdf_query { pub view_dir: pt_vec3, pub light_dir: pt_vec3, pub point: pt_vec3,
    pub geometric_normal: pt_vec3, pub shading_normal: pt_vec3, pub u: f64, pub v: f64, pub front_face: u32, pub _pad: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct pt_bsdf_result { pub eval: pt_vec3, pub pdf: f64, pub emitted: pt_vec3, pub _pad: f64 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct pt_bsdf_sample_result { pub dir: pt_vec3, pub valid: u32, pub n_uniforms: u32 }
pub enum pt_ctx {}
pub enum pt_scene {}

extern "C" {
    // lifecycle
    pub fn pt_ctx_create(device: c_int, out: *mut *mut pt_ctx) -> c_int;
    pub fn pt_ctx_destroy(ctx: *mut pt_ctx);
    pub fn pt_ctx_set_stream(ctx: *mut pt_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn pt_ctx_set_profiling(ctx: *mut pt_ctx, level: c_int) -> c_int;
    pub fn pt_last_error() -> *const c_char;
    pub fn pt_device_count() -> c_int;
    pub fn pt_scene_create(ctx: *mut pt_ctx, desc: *const pt_scene_desc, out: *mut *mut pt_scene) -> c_int;
    pub fn pt_scene_destroy(scene: *mut pt_scene);
    pub fn pt_scene_device_bytes(scene: *const pt_scene) -> u64;
    pub fn pt_scene_build_env_sampler(scene: *mut pt_scene, image: u32, max_rows: u32, max_cols: u32) -> c_int;
    // Camera::render (camera.rs:79-126 minus the PNG encode)
    pub fn pt_camera_image_height(cam: *const pt_camera) -> u32;
    pub fn pt_render(ctx: *mut pt_ctx, scene: *const pt_scene, cam: *const pt_camera, params: *const pt_render_params,
                     h_mean_rgb: *mut f32, stats: *mut pt_stats) -> c_int;
    pub fn pt_render_accumulate(ctx: *mut pt_ctx, scene: *const pt_scene, cam: *const pt_camera, params: *const pt_render_params,
                                d_accum: *mut f32, stats: *mut pt_stats) -> c_int;
    pub fn pt_render_multi(n_devices: c_int, devices: *const c_int, desc: *const pt_scene_desc, cam: *const pt_camera,
                           params: *const pt_render_params, h_mean_rgb: *mut f32, stats: *mut pt_stats) -> c_int;
    pub fn pt_render_multi_release();
    pub fn pt_tonemap_rgb8(ctx: *mut pt_ctx, d_accum: *const f32, scale: f64, n_pixels: u32, h_rgb8: *mut u8) -> c_int;
    // parity / test entry points
    pub fn pt_trace_closest(ctx: *mut pt_ctx, scene: *const pt_scene, n: usize, rays: *const pt_ray, t_min: f64, hits: *mut pt_hit) -> c_int;
    pub fn pt_trace_closest_wavefront(ctx: *mut pt_ctx, scene: *const pt_scene, n: usize, rays: *const pt_ray, t_min: f64, flags: u32,
                                      hits: *mut pt_hit, stats: *mut pt_stats) -> c_int;
    pub fn pt_trace_camera_wavefront(ctx: *mut pt_ctx, scene: *const pt_scene, cam: *const pt_camera, seed: u64, sample: u32, flags: u32,
                                     rays: *mut pt_ray, hits: *mut pt_hit, stats: *mut pt_stats) -> c_int;
    pub fn pt_trace_any(ctx: *mut pt_ctx, scene: *const pt_scene, n: usize, rays: *const pt_ray, t_min: f64, t_max: *const f64, occluded: *mut u8) -> c_int;
    pub fn pt_bsdf_eval_pdf(ctx: *mut pt_ctx, scene: *const pt_scene, material: u32, n: usize, q: *const pt_bsdf_query, out: *mut pt_bsdf_result) -> c_int;
    pub fn pt_bsdf_sample(ctx: *mut pt_ctx, scene: *const pt_scene, material: u32, n: usize, q: *const pt_bsdf_query, uniforms8: *const f64,
                          out: *mut pt_bsdf_sample_result) -> c_int;
    pub fn pt_camera_rays(ctx: *mut pt_ctx, cam: *const pt_camera, seed: u64, n: usize, row: *const u32, col: *const u32, sample: *const u32,
                          out: *mut pt_ray) -> c_int;
    pub fn pt_lights_sample_pdf(ctx: *mut pt_ctx, scene: *const pt_scene, n: usize, origin: *const pt_vec3, time: *const f64, uniforms4: *const f64,
                                dir: *mut pt_vec3, valid: *mut u32, pdf: *mut f64) -> c_int;
    pub fn pt_sah_sweep(ctx: *mut pt_ctx, n: u32, boxes6: *const f64, parent6: *const f64, cost3n: *mut f64) -> c_int;
    pub fn pt_env_sample_pdf(ctx: *mut pt_ctx, scene: *const pt_scene, n: usize, uniforms2: *const f64, dir: *mut pt_vec3, pdf: *mut f64) -> c_int;
    // diagnostics
    pub fn pt_debug_histograms(ctx: *mut pt_ctx, out512: *mut u64, reset: c_int) -> c_int;
    pub fn pt_debug_stage_ms(ctx: *mut pt_ctx, out16: *mut f64, reset: c_int) -> c_int;
    pub fn pt_debug_div_check(ctx: *mut pt_ctx, n: u64, seed: u64, mismatches: *mut u64) -> c_int;
}

/// The message of the last failure on the calling thread.
pub fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(pt_last_error()).to_string_lossy().into_owned() }
}

// the C side's sizes (tests/test_abi_and_host.py checks the same numbers against the compiled library)
const _: () = {
    assert!(std::mem::size_of::<pt_vec3>() == 24);
    assert!(std::mem::size_of::<pt_texture>() == 48);
    assert!(std::mem::size_of::<pt_material>() == 120);
    assert!(std::mem::size_of::<pt_sphere>() == 64);
    assert!(std::mem::size_of::<pt_quad>() == 136);
    assert!(std::mem::size_of::<pt_triangle>() == 72);
    assert!(std::mem::size_of::<pt_cuboid>() == 56);
    assert!(std::mem::size_of::<pt_mesh>() == 24);
    assert!(std::mem::size_of::<pt_instance>() == 448);
    assert!(std::mem::size_of::<pt_bvh_node>() == 64);
    assert!(std::mem::size_of::<pt_scene_desc>() == 200);
    assert!(std::mem::size_of::<pt_camera>() == 160);
    assert!(std::mem::size_of::<pt_render_params>() == 32);
    assert!(std::mem::size_of::<pt_stats>() == 104);
    assert!(std::mem::size_of::<pt_ray>() == 56);
    assert!(std::mem::size_of::<pt_hit>() == 128);
};
