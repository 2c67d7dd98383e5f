//! Reference-side patch, part 2: `src/flatten.rs` — walks a `World` ONCE and produces the closed, flattened
//! `pt_scene_desc` of include/pt_b200.h.  This is what replaces the trait-object graph at the FFI boundary; derived
//! quantities (quad w / normal / d, instance matrices, every BVH) are COPIED from what the reference's own constructors
//! computed (quad.rs:17-36, instance.rs:20-31, bvh.rs:24-120) — the library never re-derives them, and exact-tie winners
//! depend on the host tree's DFS order, so the trees are copied verbatim.
//!
//! Rules (the C++ mirror `thu-acg-f2024-path-tracer_b200/host/scene.cpp: FlatScene` is the executable version of them):
//!  * primitive ids = first-visit order; an `Arc` that is shared (a mesh behind two instances, a material on many
//!    spheres) is emitted once (keyed by the Arc's data pointer);
//!  * a `Cuboid` contributes its six `Quad`s in cuboid.rs:18-53 order as `quads[first_quad .. first_quad + 6)`;
//!  * a mesh contributes its triangles in `triangles.objects()` order (= `mesh.indices` order, mesh.rs:174-193), its
//!    per-corner normals / uvs as 3 x pt_vec3 / 6 x f64 per triangle, and its own BVH;
//!  * BVH nodes: `bbox` as f64, children for `Internal`, the leaf's hittables as `pt_ref`s IN LEAF ORDER;
//!    `*_bvh_root = PT_NONE` when the list was never built (linear scan, list.rs:57-66);
//!  * `MixBxDf` children precede the mix in `materials`; images are the `RgbImage` bytes (`to_rgb8()`, texture.rs:62-69).
use std::{collections::HashMap, sync::Arc};

use pt_b200_sys::*;

use crate::{
    bsdf::BxDFMaterial,
    camera::EnvironmentType,
    describe::{HittableDesc, MaterialDesc, TextureDesc},
    hittable::{BVHNode, Hittable, HittableList, World},
    texture::Texture,
    vec3::{Mat4, Vec3},
};

fn v(a: Vec3) -> pt_vec3 { pt_vec3 { x: a.x, y: a.y, z: a.z } }
fn key<T: ?Sized>(a: &Arc<T>) -> usize { Arc::as_ptr(a) as *const u8 as usize }

/// Owns every array `pt_scene_desc` points into; keep it alive until `pt_scene_create` has returned (the library copies).
#[derive(Default)]
pub struct FlatScene {
    pub textures: Vec<pt_texture>, pub images: Vec<pt_image>, image_bytes: Vec<Vec<u8>>, pub materials: Vec<pt_material>,
    pub spheres: Vec<pt_sphere>, pub quads: Vec<pt_quad>, pub triangles: Vec<pt_triangle>, pub tri_normals: Vec<pt_vec3>, pub tri_uvs: Vec<f64>,
    pub cuboids: Vec<pt_cuboid>, pub meshes: Vec<pt_mesh>, pub instances: Vec<pt_instance>, pub nodes: Vec<pt_bvh_node>, pub leaf_refs: Vec<pt_ref>,
    pub objects: Vec<pt_ref>, pub lights: Vec<pt_ref>, pub objects_bvh_root: u32, pub lights_bvh_root: u32,
    pub env_color: pt_vec3, pub env_image: Option<u32>,
    any_normals: bool, any_uvs: bool,
    seen_hittable: HashMap<usize, pt_ref>, seen_material: HashMap<usize, u32>, seen_tex3: HashMap<usize, u32>, seen_tex1: HashMap<usize, u32>,
    seen_image: HashMap<usize, u32>,
}

impl FlatScene {
    fn image(&mut self, img: &image::RgbImage) -> u32 {
        let k = img.as_raw().as_ptr() as usize;
        if let Some(&i) = self.seen_image.get(&k) { return i; }
        self.image_bytes.push(img.as_raw().clone());
        let bytes = self.image_bytes.last().unwrap();
        self.images.push(pt_image { rgb: bytes.as_ptr(), width: img.width(), height: img.height() });
        let i = (self.images.len() - 1) as u32;
        self.seen_image.insert(k, i);
        i
    }
    fn texture3(&mut self, t: &Arc<dyn Texture<Vec3>>) -> u32 {
        if let Some(&i) = self.seen_tex3.get(&key(t)) { return i; }
        let rec = match t.describe().expect("texture kind outside the closed device subset") {
            TextureDesc::Solid(c) => pt_texture { kind: PT_TEX_SOLID, tex1: 0, tex2: 0, image: 0, inv_scale: 0.0, value: v(c) },
            TextureDesc::Checker { inv_scale, tex1, tex2 } => {
                let (a, b) = (self.texture3(tex1), self.texture3(tex2));
                pt_texture { kind: PT_TEX_CHECKER, tex1: a, tex2: b, image: 0, inv_scale, value: pt_vec3::default() }
            }
            TextureDesc::Image(img) => pt_texture { kind: PT_TEX_IMAGE, tex1: 0, tex2: 0, image: self.image(img), inv_scale: 0.0, value: pt_vec3::default() },
        };
        self.textures.push(rec);
        let i = (self.textures.len() - 1) as u32;
        self.seen_tex3.insert(key(t), i);
        i
    }
    fn texture1(&mut self, t: &Arc<dyn Texture<f64>>) -> u32 {  // Texture<f64>: the scalar travels in value.x
        if let Some(&i) = self.seen_tex1.get(&key(t)) { return i; }
        let rec = match t.describe().expect("texture kind outside the closed device subset") {
            TextureDesc::Solid(s) => pt_texture { kind: PT_TEX_SOLID, tex1: 0, tex2: 0, image: 0, inv_scale: 0.0, value: pt_vec3 { x: s, y: s, z: s } },
            TextureDesc::Checker { inv_scale, tex1, tex2 } => {
                let (a, b) = (self.texture1(tex1), self.texture1(tex2));
                pt_texture { kind: PT_TEX_CHECKER, tex1: a, tex2: b, image: 0, inv_scale, value: pt_vec3::default() }
            }
            TextureDesc::Image(_) => panic!("the reference has no ImageTexture of f64"),
        };
        self.textures.push(rec);
        let i = (self.textures.len() - 1) as u32;
        self.seen_tex1.insert(key(t), i);
        i
    }
    fn material(&mut self, m: &Arc<dyn BxDFMaterial>) -> u32 {
        if let Some(&i) = self.seen_material.get(&key(m)) { return i; }
        let mut rec = pt_material { kind: 0, base_color_tex: PT_NONE, roughness_tex: PT_NONE, normal_map: PT_NONE, mix_a: PT_NONE, mix_b: PT_NONE, p: [0.0; 12] };
        match m.describe().expect("material kind outside the closed device subset") {
            MaterialDesc::Diffuse { base_color, normal_map } => {
                rec.kind = PT_MAT_DIFFUSE; rec.base_color_tex = self.texture3(base_color);
                if let Some(nm) = normal_map { rec.normal_map = self.image(nm.image()); }   // diffuse.rs:16, texture.rs:58-60
            }
            MaterialDesc::Metal { base_color, roughness } => { rec.kind = PT_MAT_METAL; rec.base_color_tex = self.texture3(base_color); rec.roughness_tex = self.texture1(roughness); }
            MaterialDesc::Glass { base_color, roughness, ior } => {
                rec.kind = PT_MAT_GLASS; rec.base_color_tex = self.texture3(base_color); rec.roughness_tex = self.texture1(roughness); rec.p[PT_P_IOR] = ior;
            }
            MaterialDesc::Principled { base_color, p } => { rec.kind = PT_MAT_PRINCIPLED; rec.base_color_tex = self.texture3(base_color); rec.p[..11].copy_from_slice(&p); }
            MaterialDesc::Light { emission } => { rec.kind = PT_MAT_LIGHT; rec.base_color_tex = self.texture3(emission); }
            MaterialDesc::Sheen { base_color, sheen_tint } => { rec.kind = PT_MAT_SHEEN; rec.p[0] = base_color.x; rec.p[1] = base_color.y; rec.p[2] = base_color.z; rec.p[PT_P_SHEEN_TINT] = sheen_tint; }
            MaterialDesc::Clearcoat { alpha_g } => { rec.kind = PT_MAT_CLEARCOAT; rec.p[PT_P_ALPHA_G] = alpha_g; }
            MaterialDesc::Mix { t, bxdf1, bxdf2 } => {  // children first: the library requires mix_a, mix_b < own index
                let (a, b) = (self.material(bxdf1), self.material(bxdf2));
                rec.kind = PT_MAT_MIX; rec.mix_a = a; rec.mix_b = b; rec.p[PT_P_MIX_T] = t;
            }
        }
        self.materials.push(rec);
        let i = (self.materials.len() - 1) as u32;
        self.seen_material.insert(key(m), i);
        i
    }
    fn quad(&mut self, h: &Arc<dyn Hittable>) -> u32 {
        match h.describe() {
            Some(HittableDesc::Quad { q, u, v: vv, w, normal, d, material }) => {
                let material = self.material(material);
                self.quads.push(pt_quad { q: v(q), u: v(u), v: v(vv), w: v(w), normal: v(normal), d, material, _pad: 0 });
                (self.quads.len() - 1) as u32
            }
            _ => panic!("expected a Quad"),
        }
    }
    /// One host-built tree (bvh.rs:6-16) in DFS order, left before right; returns the root's node index.
    fn bvh(&mut self, node: &BVHNode) -> u32 {
        let me = self.nodes.len() as u32;
        let b = node.bounding_box();
        self.nodes.push(pt_bvh_node { bmin: [b.min.x, b.min.y, b.min.z], bmax: [b.max.x, b.max.y, b.max.z], left: PT_NONE, right: PT_NONE, first_ref: 0, n_refs: 0 });
        match node {
            BVHNode::Leaf { hittables, .. } => {
                let first = self.leaf_refs.len() as u32;
                for h in hittables { let r = self.hittable(h); self.leaf_refs.push(r); }   // leaf order = tie order (bvh.rs:131-140)
                self.nodes[me as usize].first_ref = first;
                self.nodes[me as usize].n_refs = hittables.len() as u32;
            }
            BVHNode::Internal { left, right, .. } => {
                let l = self.bvh(left);
                let r = self.bvh(right);
                self.nodes[me as usize].left = l;
                self.nodes[me as usize].right = r;
            }
        }
        me
    }
    /// Any object that can sit in a HittableList; shared Arcs are emitted once.
    fn hittable(&mut self, h: &Arc<dyn Hittable>) -> pt_ref {
        if let Some(&r) = self.seen_hittable.get(&key(h)) { return r; }
        let r = match h.describe().expect("hittable kind outside the closed device subset") {
            HittableDesc::Sphere { radius, position1, position2, moving, material } => {
                let material = self.material(material);
                self.spheres.push(pt_sphere { position1: v(position1), position2: v(position2), radius, material, is_moving: moving as u32 });
                pt_ref { kind: PT_PRIM_SPHERE, index: (self.spheres.len() - 1) as u32 }
            }
            HittableDesc::Quad { .. } => pt_ref { kind: PT_PRIM_QUAD, index: self.quad(h) },
            HittableDesc::Cuboid { a, b, sides, material } => {
                let first_quad = self.quads.len() as u32;
                for side in sides.objects() { self.quad(side); }            // six quads, cuboid.rs:18-53 order
                let material = self.material(material);
                self.cuboids.push(pt_cuboid { first_quad, material, a: v(a), b: v(b) });
                pt_ref { kind: PT_OBJ_CUBOID, index: (self.cuboids.len() - 1) as u32 }
            }
            HittableDesc::Triangle { .. } => panic!("a bare Triangle cannot be a top-level object"),
            HittableDesc::Mesh { triangles } => {
                let first_triangle = self.triangles.len() as u32;
                let (mut has_n, mut has_uv, mut material) = (false, false, 0u32);
                let mut local = HashMap::new();
                for (k, t) in triangles.objects().iter().enumerate() {
                    if let Some(HittableDesc::Triangle { vertices, normals, uvs, material: m }) = t.describe() {
                        local.insert(key(t), pt_ref { kind: PT_PRIM_TRIANGLE, index: first_triangle + k as u32 });
                        self.triangles.push(pt_triangle { v0: v(vertices[0]), v1: v(vertices[1]), v2: v(vertices[2]) });
                        let n = normals.unwrap_or([Vec3::ZERO; 3]);
                        self.tri_normals.extend(n.iter().map(|&x| v(x)));
                        let uv = uvs.unwrap_or([(0.0, 0.0); 3]);
                        self.tri_uvs.extend(uv.iter().flat_map(|&(a, b)| [a, b]));
                        has_n |= normals.is_some(); has_uv |= uvs.is_some();
                        material = self.material(m);                        // one material per mesh (mesh.rs:148)
                    } else { panic!("a TriangleMesh holds Triangles only"); }
                }
                self.any_normals |= has_n; self.any_uvs |= has_uv;
                self.seen_hittable.extend(local);                           // the mesh BVH's leaves refer to these
                let bvh_root = match triangles.bvh() { Some(root) => self.bvh(root), None => PT_NONE };
                self.meshes.push(pt_mesh { first_triangle, n_triangles: triangles.len() as u32, material, bvh_root, has_normals: has_n as u32, has_uvs: has_uv as u32 });
                pt_ref { kind: PT_OBJ_MESH, index: (self.meshes.len() - 1) as u32 }
            }
            HittableDesc::Instance { object, axis, angle, translation, transform, normal_matrix } => {
                let child = self.hittable(object);
                assert!(child.kind != PT_OBJ_INSTANCE, "nested instances are outside the closed device subset");
                let m = |a: Mat4| a.to_cols_array();                        // column-major like glam::DMat4
                self.instances.push(pt_instance { child, axis: v(axis), angle, translation: v(translation), transform: m(transform),
                                                  inverse: m(transform.inverse()), normal_matrix: m(normal_matrix) });  // instance.rs:36,45
                pt_ref { kind: PT_OBJ_INSTANCE, index: (self.instances.len() - 1) as u32 }
            }
        };
        self.seen_hittable.insert(key(h), r);
        r
    }
    fn list(&mut self, l: &HittableList) -> (Vec<pt_ref>, u32) {
        let refs: Vec<pt_ref> = l.objects().iter().map(|h| self.hittable(h)).collect();   // insertion order (world.rs:6-7)
        let root = match l.bvh() { Some(root) => self.bvh(root), None => PT_NONE };
        (refs, root)
    }

    /// The description the library consumes.  Borrows `self`: keep the FlatScene alive across `pt_scene_create`.
    pub fn desc(&self) -> pt_scene_desc {
        fn p<T>(x: &Vec<T>) -> *const T { if x.is_empty() { std::ptr::null() } else { x.as_ptr() } }
        pt_scene_desc {
            abi_version: PT_ABI_VERSION, n_textures: self.textures.len() as u32, n_images: self.images.len() as u32, n_materials: self.materials.len() as u32,
            n_spheres: self.spheres.len() as u32, n_quads: self.quads.len() as u32, n_triangles: self.triangles.len() as u32,
            n_cuboids: self.cuboids.len() as u32, n_meshes: self.meshes.len() as u32, n_instances: self.instances.len() as u32,
            n_nodes: self.nodes.len() as u32, n_leaf_refs: self.leaf_refs.len() as u32, n_objects: self.objects.len() as u32, n_lights: self.lights.len() as u32,
            textures: p(&self.textures), images: p(&self.images), materials: p(&self.materials), spheres: p(&self.spheres), quads: p(&self.quads),
            triangles: p(&self.triangles),
            tri_normals: if self.any_normals { p(&self.tri_normals) } else { std::ptr::null() },
            tri_uvs: if self.any_uvs { p(&self.tri_uvs) } else { std::ptr::null() },
            cuboids: p(&self.cuboids), meshes: p(&self.meshes), instances: p(&self.instances), nodes: p(&self.nodes), leaf_refs: p(&self.leaf_refs),
            objects: p(&self.objects), lights: p(&self.lights), objects_bvh_root: self.objects_bvh_root, lights_bvh_root: self.lights_bvh_root,
            n_volumes: 0, _pad: 0, volumes: std::ptr::null(),
        }
    }
}

/// `World` (world.rs:5-8) + the camera's environment (camera.rs:16-19) -> FlatScene.
pub fn flatten(world: &World, environment: &EnvironmentType) -> FlatScene {
    let mut f = FlatScene { objects_bvh_root: PT_NONE, lights_bvh_root: PT_NONE, ..Default::default() };
    let (objects, objects_root) = f.list(&world.objects);
    let (lights, lights_root) = f.list(&world.lights);
    f.objects = objects; f.lights = lights; f.objects_bvh_root = objects_root; f.lights_bvh_root = lights_root;
    match environment {
        EnvironmentType::Color(c) => f.env_color = v(*c),
        EnvironmentType::Map(tex) => f.env_image = Some(f.image(tex.image())),
    }
    f
}
