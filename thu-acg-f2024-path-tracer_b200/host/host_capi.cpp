// C wrappers over the host mirror so Python (ctypes) tests and bench.py can build scenes the way the
// reference's main.rs does.  Handles are opaque pointers to heap-allocated shared_ptrs.
#include <cstring>
#include <stdexcept>
#include <string>

#include "pt_host.hpp"

using namespace pt;
namespace pt { extern ImagePtr g_envmap_override; }

static thread_local std::string g_err;
template <class T> static void* box(std::shared_ptr<const T> p) { return new std::shared_ptr<const T>(std::move(p)); }
template <class T> static std::shared_ptr<const T> unbox(void* h) { return h ? *static_cast<std::shared_ptr<const T>*>(h) : nullptr; }
static Vec3 V(const double* p) { return Vec3(p[0], p[1], p[2]); }
#define PTH_TRY(expr)                                             \
    try { expr; } catch (const std::exception& e) { g_err = e.what(); return nullptr; }

struct pth_scene {
    std::unique_ptr<SceneBundle> bundle;
    std::unique_ptr<FlatScene> flat;
    pt_camera cam{};
    std::string output_name;
};

extern "C" {

const char* pth_last_error() { return g_err.c_str(); }

// ---- textures / images ----
void* pth_solid_texture(double r, double g, double b) { return box<Texture>(SolidTexture::make(Vec3(r, g, b))); }
void* pth_solid_scalar(double v) { return box<Texture>(SolidTexture::scalar(v)); }
void* pth_checker_texture(double scale, void* t1, void* t2) { return box<Texture>(CheckerTexture::make(scale, unbox<Texture>(t1), unbox<Texture>(t2))); }
void* pth_image(const uint8_t* rgb, uint32_t w, uint32_t h) { return box<Image>(ImageTexture::from_rgb8(rgb, w, h)); }
void* pth_image_load(const char* path) { PTH_TRY(return box<Image>(ImageTexture::load(path))); }
void* pth_image_texture(void* image) { return box<Texture>(ImageTexture::make(unbox<Image>(image))); }
// decoded pixels of an image handle: returns the RGB8 bytes (row-major, top row first) and its size
const uint8_t* pth_image_pixels(void* image, uint32_t* w, uint32_t* h) {
    auto im = unbox<Image>(image);
    *w = im->width; *h = im->height;
    return im->rgb.data();
}
// ---- materials ----
void* pth_diffuse(void* tex, void* normal_image) { return box<Material>(DiffuseBRDF::from_textures(unbox<Texture>(tex), unbox<Image>(normal_image))); }
void* pth_metal(void* tex, void* rough) { return box<Material>(MetalBRDF::make(unbox<Texture>(tex), unbox<Texture>(rough))); }
void* pth_glass(void* tex, void* rough, double ior) { return box<Material>(GlassBSDF::make(unbox<Texture>(tex), unbox<Texture>(rough), 0.0, ior)); }
void* pth_principled(void* tex, const double* p) {
    return box<Material>(PrincipledBSDF::make(unbox<Texture>(tex), p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10]));
}
void* pth_diffuse_light(void* tex) { return box<Material>(DiffuseLight::make(unbox<Texture>(tex))); }
void* pth_sheen(double r, double g, double b, double tint) { return box<Material>(SheenBRDF::make(Vec3(r, g, b), tint)); }
void* pth_clearcoat(double gloss) { return box<Material>(ClearcoatBRDF::make(gloss)); }
void* pth_mix(double t, void* a, void* b) { return box<Material>(MixBxDf::make(t, unbox<Material>(a), unbox<Material>(b))); }
// ---- hittables ----
void* pth_sphere_still(double r, const double* p, void* mat) { return box<Hittable>(Sphere::new_still(r, V(p), unbox<Material>(mat))); }
void* pth_sphere_moving(double r, const double* p1, const double* p2, void* mat) { return box<Hittable>(Sphere::new_moving(r, V(p1), V(p2), unbox<Material>(mat))); }
void* pth_quad(const double* q, const double* u, const double* v, void* mat) { return box<Hittable>(Quad::make(V(q), V(u), V(v), unbox<Material>(mat))); }
void* pth_cuboid(const double* a, const double* b, void* mat) { return box<Hittable>(Cuboid::make(V(a), V(b), unbox<Material>(mat))); }
void* pth_mesh_load(const char* path, double scale, void* mat) { PTH_TRY(return box<Hittable>(TriangleMesh::from_obj(scale, ObjMesh::load(path), unbox<Material>(mat)))); }
void* pth_mesh_from_arrays(double scale, const float* pos, uint32_t n_pos_floats, const uint32_t* idx, uint32_t n_idx, const float* tex,
                           uint32_t n_tex_floats, const float* nrm, uint32_t n_nrm_floats, void* mat) {
    ObjMesh m;
    m.positions.assign(pos, pos + n_pos_floats); m.indices.assign(idx, idx + n_idx);
    if (tex) m.texcoords.assign(tex, tex + n_tex_floats);
    if (nrm) m.normals.assign(nrm, nrm + n_nrm_floats);
    PTH_TRY(return box<Hittable>(TriangleMesh::from_obj(scale, m, unbox<Material>(mat))));
}
void* pth_instance(void* child, const double* axis, double angle, const double* translation) {
    PTH_TRY(return box<Hittable>(Instance::make(unbox<Hittable>(child), V(axis), angle, V(translation))));
}
void* pth_volume(void* boundary, double density, void* albedo_tex) {
    PTH_TRY(return box<Hittable>(HomogeneousVolume::from_texture(unbox<Hittable>(boundary), density, unbox<Texture>(albedo_tex))));
}
void pth_set_build_context(void* ctx) { set_build_context(static_cast<pt_ctx*>(ctx)); }
// ---- world ----
void* pth_world_new() { return new World(); }
void pth_world_free(void* w) { delete static_cast<World*>(w); }
void pth_world_add_object(void* w, void* h) { static_cast<World*>(w)->add_object(unbox<Hittable>(h)); }
void pth_world_add_light(void* w, void* h) { static_cast<World*>(w)->add_light(unbox<Hittable>(h)); }
void pth_world_build_bvh(void* w) { static_cast<World*>(w)->build_bvh(); }

// ---- scenes ----
// Flatten a hand-built world; `cam` holds the public Camera fields, env_image is a pth_image handle or NULL.
pth_scene* pth_scene_from_world(void* world, const pt_camera* cam, void* env_image) {
    try {
        auto s = std::make_unique<pth_scene>();
        s->flat = flatten(*static_cast<World*>(world));
        s->cam = *cam;
        s->cam.env_image = PT_NONE;
        if (cam->env_is_map) { s->cam.env_image = s->flat->add_image(unbox<Image>(env_image)); s->flat->finish(); }
        s->output_name = "render.png";
        return s.release();
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
// One of the reference's scenes (1..7; 70 = our mesh variant of scene 7).  env_rgb: optional decoded
// assets/envmap.jpg for scene 5 (skips the native decode when the caller already holds the pixels).
pth_scene* pth_scene_build(int scene, uint32_t width, uint32_t spp, uint64_t seed, const char* assets_dir, const uint8_t* env_rgb,
                           uint32_t env_w, uint32_t env_h) {
    try {
        g_envmap_override = env_rgb ? ImageTexture::from_rgb8(env_rgb, env_w, env_h) : nullptr;
        auto s = std::make_unique<pth_scene>();
        s->bundle = build_scene(scene, width, spp, seed, assets_dir ? assets_dir : "assets");
        g_envmap_override = nullptr;
        s->flat = flatten(s->bundle->world);
        s->cam = s->bundle->camera.to_abi(*s->flat);
        s->output_name = s->bundle->output_name;
        return s.release();
    } catch (const std::exception& e) { g_envmap_override = nullptr; g_err = e.what(); return nullptr; }
}
const pt_scene_desc* pth_scene_desc(const pth_scene* s) { return &s->flat->desc; }
const pt_camera* pth_scene_camera(const pth_scene* s) { return &s->cam; }
const char* pth_scene_output_name(const pth_scene* s) { return s->output_name.c_str(); }
void pth_scene_free(pth_scene* s) { delete s; }
// Camera::render (camera.rs:79): render through the CUDA library and write `filename` as PNG.
int pth_scene_render(const pth_scene* s, const char* filename, uint64_t seed, int device, uint32_t nan_policy, int verbose, pt_stats* stats) {
    RenderOptions o; o.seed = seed; o.device = device; o.nan_policy = nan_policy; o.verbose = verbose != 0;
    try { return render_flat(s->flat->desc, s->cam, filename, o, stats); }
    catch (const std::exception& e) { g_err = e.what(); return PT_ERR_NO_DEVICE; }  // the CUDA library could not be loaded
}
int pth_write_png(const char* path, const uint8_t* rgb, uint32_t w, uint32_t h) { return write_png_rgb8(path, rgb, w, h) ? 0 : -1; }

}  // extern "C"
