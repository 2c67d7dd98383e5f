//! Reference-side patch, part 1 (goes into src/hittable/mod.rs, src/bsdf/mod.rs, src/texture.rs of the reference).
//!
//! The reference's scene is a graph of trait objects (`Arc<dyn Hittable>`, `Arc<dyn BxDFMaterial>`, `Arc<dyn Texture<T>>`).
//! A GPU cannot call back into them, so every trait gains ONE method that says what the object is; concrete types
//! implement it by handing out references to their own (already private) fields.  Nothing else in the types changes.
//! `flatten.rs` consumes these descriptions.
use std::sync::Arc;

use crate::{bsdf::BxDFMaterial, hittable::{BVHNode, Hittable, HittableList}, texture::ImageTexture, vec3::{Mat4, Vec3}};

pub type MatPtr = Arc<dyn BxDFMaterial>;

/// src/hittable/*.rs
pub enum HittableDesc<'a> {
    /// sphere.rs:13-19 (`moving` = built by `new_moving`, sphere.rs:34)
    Sphere { radius: f64, position1: Vec3, position2: Vec3, moving: bool, material: &'a MatPtr },
    /// quad.rs:5-14 — derived fields as `Quad::new` computed them (quad.rs:17-36)
    Quad { q: Vec3, u: Vec3, v: Vec3, w: Vec3, normal: Vec3, d: f64, material: &'a MatPtr },
    /// cuboid.rs:5-9 — `sides` holds the six quads in cuboid.rs:18-53 order
    Cuboid { a: Vec3, b: Vec3, sides: &'a HittableList, material: &'a MatPtr },
    /// mesh.rs:13-19
    Triangle { vertices: [Vec3; 3], normals: Option<[Vec3; 3]>, uvs: Option<[(f64, f64); 3]>, material: &'a MatPtr },
    /// mesh.rs:144-146 — `triangles` is a HittableList of Triangle with its own BVH (mesh.rs:195)
    Mesh { triangles: &'a HittableList },
    /// instance.rs:12-31
    Instance { object: &'a Arc<dyn Hittable>, axis: Vec3, angle: f64, translation: Vec3, transform: Mat4, normal_matrix: Mat4 },
}

/// src/bsdf/*.rs, src/material.rs:150-191
pub enum MaterialDesc<'a> {
    Diffuse { base_color: &'a Arc<dyn crate::texture::Texture<Vec3>>, normal_map: Option<&'a Arc<ImageTexture>> },  // diffuse.rs:14-17
    Metal { base_color: &'a Arc<dyn crate::texture::Texture<Vec3>>, roughness: &'a Arc<dyn crate::texture::Texture<f64>> },  // metal.rs:17-20
    Glass { base_color: &'a Arc<dyn crate::texture::Texture<Vec3>>, roughness: &'a Arc<dyn crate::texture::Texture<f64>>, ior: f64 },  // glass.rs:20-25
    /// principled.rs:23-41; `p` in the order of PT_P_METALLIC .. PT_P_CLEARCOAT_GLOSS
    Principled { base_color: &'a Arc<dyn crate::texture::Texture<Vec3>>, p: [f64; 11] },
    Light { emission: &'a Arc<dyn crate::texture::Texture<Vec3>> },  // material.rs:150-153
    Sheen { base_color: Vec3, sheen_tint: f64 },                     // sheen.rs:11-14
    Clearcoat { alpha_g: f64 },                                      // clearcoat.rs:10-12
    Mix { t: f64, bxdf1: &'a MatPtr, bxdf2: &'a MatPtr },            // mix.rs:8-12
}

/// src/texture.rs:11-92; a `Texture<f64>` puts its scalar in `.x`
pub enum TextureDesc<'a, T> {
    Solid(T),
    Checker { inv_scale: f64, tex1: &'a Arc<dyn crate::texture::Texture<T>>, tex2: &'a Arc<dyn crate::texture::Texture<T>> },
    Image(&'a image::RgbImage),
}

// Additions to the traits (each with a default so that third-party implementors keep compiling):
//
//   pub trait Hittable      { …existing…  fn describe(&self) -> Option<HittableDesc<'_>> { None } }
//   pub trait BxDFMaterial  { …existing…  fn describe(&self) -> Option<MaterialDesc<'_>> { None } }
//   pub trait Texture<T>    { …existing…  fn describe(&self) -> Option<TextureDesc<'_, T>> { None } }
//
// and two accessors on HittableList (list.rs:9-13 keeps `objects` and `bvh` private):
//
//   impl HittableList { pub fn objects(&self) -> &[Arc<dyn Hittable>] { &self.objects }
//                       pub fn bvh(&self) -> Option<&BVHNode> { self.bvh.as_ref() } }
//
// BVHNode (bvh.rs:6-16) is already a public enum with public variants, so the tree can be walked as is.
#[allow(dead_code)]
fn _types_used(_: &BVHNode) {}
