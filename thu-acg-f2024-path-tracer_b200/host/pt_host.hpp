// Host-side mirror of the reference's scene API (what a Rust host keeps: src/main.rs scene builders,
// the Hittable/Material/Texture surface, and the SAH BVH build of src/hittable/bvh.rs:24-120).
// It holds NO intersection or shading code: it describes the scene, builds the BVHs exactly as the
// reference does (tree shape fixes exact-tie winners) and flattens everything into
// the closed pt_scene_desc that the CUDA library consumes (include/pt_b200.h).
//
// Names follow the reference: World, HittableList, Sphere::new_still/new_moving, Quad, Cuboid, Instance,
// TriangleMesh::from_obj, DiffuseBRDF, MetalBRDF, GlassBSDF, PrincipledBSDF, DiffuseLight, SheenBRDF,
// ClearcoatBRDF, MixBxDf, SolidTexture, CheckerTexture, ImageTexture, Camera{init,render}.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pt_b200.h"

namespace pt {

// ---- minimal f64 math (host-side derivations only; glam 0.29.2 scalar semantics) -----------------
struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    double operator[](int i) const { return i == 0 ? x : i == 1 ? y : z; }
    pt_vec3 c() const { return pt_vec3{x, y, z}; }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Vec3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(double s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator*(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline Vec3 operator/(Vec3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(Vec3 a, Vec3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline double length(Vec3 a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(Vec3 a) { return a * (1.0 / length(a)); }
inline Vec3 vmin(Vec3 a, Vec3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline Vec3 vmax(Vec3 a, Vec3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }

struct Box {  // src/hittable/aabb.rs: every construction pads by 1e-3 (Q1)
    Vec3 lo{std::numeric_limits<double>::infinity(), std::numeric_limits<double>::infinity(), std::numeric_limits<double>::infinity()};
    Vec3 hi{-std::numeric_limits<double>::infinity(), -std::numeric_limits<double>::infinity(), -std::numeric_limits<double>::infinity()};
    static Box of(Vec3 a, Vec3 b) { Box r; Vec3 d(1e-3, 1e-3, 1e-3); r.lo = vmin(a, b) - d; r.hi = vmax(a, b) + d; return r; }
    Box merged(const Box& o) const { return of(vmin(lo, o.lo), vmax(hi, o.hi)); }
    Vec3 centroid() const { return 0.5 * (lo + hi); }
    double half_area() const { Vec3 e = hi - lo; return e.x * e.y + e.x * e.z + e.y * e.z; }
    Box transformed(const double m[16]) const;
};

// ---- textures (src/texture.rs) --------------------------------------------------------------------
struct Image { std::vector<uint8_t> rgb; uint32_t width = 0, height = 0; };
using ImagePtr = std::shared_ptr<const Image>;
struct Texture {
    uint32_t kind = PT_TEX_SOLID; Vec3 value; double inv_scale = 0;
    std::shared_ptr<const Texture> tex1, tex2; ImagePtr image;
};
using TexPtr = std::shared_ptr<const Texture>;
struct SolidTexture { static TexPtr make(Vec3 v); static TexPtr scalar(double v); };
struct CheckerTexture { static TexPtr make(double scale, TexPtr a, TexPtr b); };
struct ImageTexture {
    static ImagePtr load(const std::string& path);        // .png / .jpg / .hdr / .rgb8 (see assets.cpp)
    static ImagePtr from_rgb8(const uint8_t* rgb, uint32_t w, uint32_t h);
    static TexPtr make(ImagePtr img);
};

// ---- materials (src/bsdf/*.rs, src/material.rs) ---------------------------------------------------
struct Material {
    uint32_t kind = PT_MAT_DIFFUSE; TexPtr base_color, roughness; ImagePtr normal_map;
    std::shared_ptr<const Material> mix_a, mix_b; double p[12] = {0};
};
using MatPtr = std::shared_ptr<const Material>;
struct DiffuseBRDF {
    static MatPtr make(TexPtr base_color);                          // DiffuseBRDF::new
    static MatPtr from_rgb(Vec3 c);
    static MatPtr from_textures(TexPtr color, ImagePtr normal_map); // normal_map may be null
};
struct MetalBRDF { static MatPtr make(TexPtr c, TexPtr rough); static MatPtr from_rgb(Vec3 c, double rough); };
struct GlassBSDF { static MatPtr make(TexPtr c, TexPtr rough, double anisotropic, double ior); static MatPtr basic(double ior); };
struct PrincipledBSDF {
    static MatPtr make(TexPtr base_color, double metallic, double roughness, double subsurface, double specular,
                       double specular_tint, double ior, double spec_trans, double sheen, double sheen_tint,
                       double clearcoat, double clearcoat_gloss);
};
struct DiffuseLight { static MatPtr make(TexPtr emission); static MatPtr from_rgb(Vec3 rgb); };
struct SheenBRDF { static MatPtr make(Vec3 base_color, double sheen_tint); };
struct ClearcoatBRDF { static MatPtr make(double clearcoat_gloss); };
struct MixBxDf { static MatPtr make(double t, MatPtr a, MatPtr b); };
struct IsotropicMaterial { static MatPtr from_texture(TexPtr albedo); static MatPtr from_albedo(Vec3 albedo); };  // volume.rs:18 (stub)

// ---- hittables (src/hittable/*.rs) ----------------------------------------------------------------
struct BvhTree {  // host form of bvh.rs:6-16 over the items of one list
    struct Node { Box box; int32_t left = -1, right = -1; std::vector<uint32_t> items; };
    std::vector<Node> nodes;  // nodes[0] = root, DFS pre-order
};
struct Hittable;
using HitPtr = std::shared_ptr<const Hittable>;
struct HittableList {  // list.rs
    std::vector<HitPtr> objects; Box bbox; std::shared_ptr<BvhTree> bvh;
    void add(HitPtr h);
    void build_bvh();
    bool is_empty() const { return objects.empty(); }
};
struct Hittable {
    uint32_t kind = PT_PRIM_SPHERE; Box bbox; MatPtr material;
    // sphere
    double radius = 0; Vec3 p1, p2; bool moving = false;
    // quad
    Vec3 q, u, v, w, normal; double d = 0;
    // cuboid (6 quads) / mesh (triangles + own BVH)
    Vec3 a, b; HittableList sides;
    std::vector<pt_triangle> triangles; std::vector<pt_vec3> tri_normals; std::vector<double> tri_uvs;
    std::vector<Box> tri_boxes; std::shared_ptr<BvhTree> mesh_bvh;
    // instance (child = the transformed object) / volume (child = the boundary)
    HitPtr child; Vec3 axis, translation; double angle = 0; double transform[16], inverse[16], normal_matrix[16];
    double density = 0;
};
// volume.rs:15-41 (a commented-out stub in the reference; semantics: include/pt_b200.h, pt_volume)
struct HomogeneousVolume {
    static HitPtr from_texture(HitPtr boundary, double density, TexPtr texture);
    static HitPtr from_albedo(HitPtr boundary, double density, Vec3 albedo);
};
struct Sphere {
    static HitPtr new_still(double radius, Vec3 position, MatPtr m);
    static HitPtr new_moving(double radius, Vec3 p1, Vec3 p2, MatPtr m);
};
struct Quad { static HitPtr make(Vec3 q, Vec3 u, Vec3 v, MatPtr m); };
struct Cuboid { static HitPtr make(Vec3 a, Vec3 b, MatPtr m); };
struct Instance { static HitPtr make(HitPtr object, Vec3 axis, double angle, Vec3 translation); };
struct ObjMesh {  // what tobj::load_obj(.., OFFLINE_RENDERING_LOAD_OPTIONS) yields for models[0].mesh
    std::vector<float> positions, texcoords, normals; std::vector<uint32_t> indices;
    static ObjMesh load(const std::string& path);  // .obj (parsed) or .mesh (baked, tools/bake_assets.py)
};
struct TriangleMesh { static HitPtr from_obj(double scale, const ObjMesh& mesh, MatPtr m); };

struct World {  // world.rs
    HittableList objects, lights;
    void add_object(HitPtr h) { objects.add(std::move(h)); }
    void add_light(HitPtr h) { lights.add(std::move(h)); }
    void build_bvh() { objects.build_bvh(); lights.build_bvh(); }
};

// SAH builder restating bvh.rs:24-120 (full sweep, strict '<', stable partition, Q3 fallbacks).
std::shared_ptr<BvhTree> build_bvh(const std::vector<Box>& boxes);
// Optional: with a device context set (per thread), nodes of 256+ items price their splits through pt_sah_sweep — the same
// tree, bit for bit, without the O(n^2) host loop.  nullptr (default) = host only, e.g. on a machine without a GPU.
void set_build_context(pt_ctx* ctx);

// ---- flattening to the C ABI ----------------------------------------------------------------------
struct FlatScene {
    std::vector<pt_texture> textures; std::vector<pt_image> images; std::vector<ImagePtr> image_owner;
    std::vector<pt_material> materials; std::vector<pt_sphere> spheres; std::vector<pt_quad> quads;
    std::vector<pt_triangle> triangles; std::vector<pt_vec3> tri_normals; std::vector<double> tri_uvs;
    std::vector<pt_cuboid> cuboids; std::vector<pt_mesh> meshes; std::vector<pt_instance> instances;
    std::vector<pt_bvh_node> nodes; std::vector<pt_ref> leaf_refs, objects, lights;
    std::vector<pt_volume> volumes;
    pt_scene_desc desc{};
    // identity maps (pointer -> index) so shared Arc<> objects flatten once
    std::vector<const Texture*> tex_keys; std::vector<const Image*> img_keys; std::vector<const Material*> mat_keys;
    std::vector<const Hittable*> mesh_keys, cuboid_keys;
    uint32_t add_image(const ImagePtr& im);
    uint32_t add_texture(const TexPtr& t);
    uint32_t add_material(const MatPtr& m);
    pt_ref add_hittable(const HitPtr& h, bool allow_instance);
    uint32_t emit_tree(const BvhTree& tree, const std::vector<pt_ref>& item_refs);
    void finish();
};
std::unique_ptr<FlatScene> flatten(const World& world);

// ---- camera (src/camera.rs) -----------------------------------------------------------------------
struct EnvironmentType { bool is_map = false; Vec3 color; ImagePtr map; };
struct RenderOptions {  // knobs the reference does not have (it is unseeded / single device)
    uint64_t seed = 1; int device = 0; uint32_t nan_policy = PT_NAN_REFERENCE; bool verbose = true;
    bool env_importance = false;  // PT_RENDER_ENV_IMPORTANCE: sample the environment map by luminance (not reference behaviour)
    bool nee = false;             // PT_RENDER_NEE: next-event estimation with MIS (not reference behaviour)
    int gpus = 1;                 // > 1: pt_render_multi on devices device .. device + gpus - 1 (spp split, one process)
};
struct Camera {
    double aspect_ratio = 1.0; uint32_t image_width = 0, samples_per_pixel = 0, max_depth = 0;
    double vfov = 0; Vec3 look_from, look_at, vup;
    double blur_strength = 0, focal_length = 0, defocus_angle = 0;
    EnvironmentType environment;
    uint32_t image_height = 0;
    void init();  // camera.rs:51-77 (only image_height is kept host-side; the library re-derives the rest)
    pt_camera to_abi(FlatScene& flat) const;
    // camera.rs:79-126: render through the CUDA library and save an 8-bit PNG.  Returns 0 or a pt_status.
    int render(const World& world, const std::string& filename, const RenderOptions& opt = RenderOptions(),
               pt_stats* stats_out = nullptr) const;
};

// Camera::render on an already-flattened scene (what the FFI crate's render_gpu ends up calling).
int render_flat(const pt_scene_desc& desc, const pt_camera& cam, const std::string& filename, const RenderOptions& opt, pt_stats* stats_out);

// ---- the seven scenes of src/main.rs:14-618 -------------------------------------------------------
struct SceneBundle { World world; Camera camera; std::string output_name; };
// `seed` drives scene 1's random layout (the reference uses an unseeded thread_rng, main.rs:38).
// Throws std::runtime_error on a missing asset.
std::unique_ptr<SceneBundle> build_scene(int scene, uint32_t width, uint32_t spp, uint64_t seed, const std::string& assets_dir);

// Baseline / progressive Huffman JPEG -> RGB8 (jpeg.cpp); throws std::runtime_error on anything else
ImagePtr decode_jpeg(const std::vector<uint8_t>& file);
// Radiance RGBE (.hdr) -> RGB8 with the reference's clamp-and-round `to_rgb8` (hdr.cpp); throws on anything else
ImagePtr decode_hdr(const std::vector<uint8_t>& file);
// PNG I/O (assets.cpp)
bool write_png_rgb8(const std::string& path, const uint8_t* rgb, uint32_t w, uint32_t h);

}  // namespace pt
