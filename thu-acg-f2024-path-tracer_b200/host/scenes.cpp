// The seven scene builders of the reference (src/main.rs:14-618), restated against the host mirror.
// Geometry, materials and camera values are the reference's; only scene 1's random layout differs in
// that it is seeded (the reference draws from an unseeded thread_rng, main.rs:38-59).
#include <cstdio>
#include <stdexcept>

#include "pt_host.hpp"

namespace pt {

ImagePtr g_envmap_override;  // scene 5: a pre-decoded assets/envmap.jpg handed in by the caller (host_capi.cpp)

// An assets directory may be the reference's own `assets/` (bunny.obj, earthmap.jpg, grace_probe_latlong.hdr,
// bricks/color.png: all decoded natively by host/assets.cpp, jpeg.cpp, hdr.cpp) or this repo's `assets/`, which also
// holds bakes made by tools/bake_assets.py (.mesh, .png).  The reference's file name wins when it is there.
static std::string asset(const std::string& dir, const char* original, const char* baked) {
    std::string p = dir + "/" + original;
    if (FILE* f = fopen(p.c_str(), "rb")) { fclose(f); return p; }
    return dir + "/" + baked;
}

namespace {
struct SplitMix {  // seeded stand-in for rand::thread_rng() during scene construction
    uint64_t s;
    double next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
    Vec3 vec() { double a = next(), b = next(), c = next(); return Vec3(a, b, c); }                          // vec3.rs:17-20
    Vec3 vec_range(double lo, double hi) { Vec3 v = vec(); return Vec3(lo, lo, lo) + v * (hi - lo); }       // vec3.rs:8-15
};
void set_camera(Camera& c, double aspect, uint32_t width, uint32_t spp, double vfov, Vec3 from, Vec3 at, double focal, double defocus) {
    c.aspect_ratio = aspect; c.image_width = width; c.samples_per_pixel = spp; c.max_depth = 50; c.vfov = vfov;
    c.look_from = from; c.look_at = at; c.vup = Vec3(0, 1, 0); c.blur_strength = 0.5; c.focal_length = focal; c.defocus_angle = defocus;
}
MatPtr principled(Vec3 color, double metallic, double roughness, double subsurface, double specular, double specular_tint, double ior,
                  double spec_trans, double sheen, double sheen_tint, double clearcoat, double clearcoat_gloss) {
    return PrincipledBSDF::make(SolidTexture::make(color), metallic, roughness, subsurface, specular, specular_tint, ior, spec_trans, sheen,
                                sheen_tint, clearcoat, clearcoat_gloss);
}

void balls_scene(SceneBundle& b, uint32_t width, uint32_t spp, uint64_t seed) {  // main.rs:14-82
    World& world = b.world;
    auto checker = CheckerTexture::make(0.32, SolidTexture::make(Vec3(0.2, 0.3, 0.1)), SolidTexture::make(Vec3(0.9, 0.9, 0.9)));
    world.add_object(Sphere::new_still(1000.0, Vec3(0.0, -1000.0, 0.0), DiffuseBRDF::make(checker)));
    world.add_object(Sphere::new_still(1.0, Vec3(0.0, 1.0, 0.0), GlassBSDF::basic(1.5)));
    world.add_object(Sphere::new_still(1.0, Vec3(-4.0, 1.0, 0.0), DiffuseBRDF::from_rgb(Vec3(0.4, 0.2, 0.1))));
    world.add_object(Sphere::new_still(1.0, Vec3(4.0, 1.0, 0.0), MetalBRDF::from_rgb(Vec3(0.7, 0.6, 0.5), 0.0)));
    SplitMix rng{seed};
    for (int ai = -11; ai < 11; ai++) {
        for (int bi = -11; bi < 11; bi++) {
            double a = ai, bb = bi;
            double choose_mat = rng.next();
            double cx = a + 0.9 * rng.next(); double cz = bb + 0.9 * rng.next();
            Vec3 center(cx, 0.2, cz);
            if (length(center - Vec3(4.0, 0.2, 0.0)) > 0.9) {
                if (choose_mat < 0.8) {
                    Vec3 albedo = rng.vec() * rng.vec();
                    Vec3 pos2 = center + Vec3(0.0, rng.next() * 0.5, 0.0);
                    world.add_object(Sphere::new_moving(0.2, center, pos2, DiffuseBRDF::from_rgb(albedo)));
                } else if (choose_mat < 0.95) {
                    world.add_object(Sphere::new_still(0.2, center, MetalBRDF::from_rgb(rng.vec_range(0.5, 1.0), 0.0)));
                } else {
                    world.add_object(Sphere::new_still(0.2, center, GlassBSDF::basic(1.5)));
                }
            }
        }
    }
    world.build_bvh();
    set_camera(b.camera, 16.0 / 9.0, width, spp, 20.0, Vec3(13.0, 2.0, 3.0), Vec3(0, 0, 0), 10.0, 0.6);
    b.camera.environment.color = Vec3(0.7, 0.8, 1.0);
    b.output_name = "balls.png";
}

void earth_scene(SceneBundle& b, uint32_t width, uint32_t spp, const std::string& assets) {  // main.rs:84-132
    World& world = b.world;
    auto earth = ImageTexture::make(ImageTexture::load(asset(assets, "earthmap.jpg", "earthmap.png")));
    world.add_object(Sphere::new_still(1.0, Vec3(4.9, 1.0, 3.0), DiffuseBRDF::make(earth)));
    world.add_object(Sphere::new_still(1.0, Vec3(0.0, 1.0, 0.0), DiffuseBRDF::from_rgb(Vec3(0.4, 0.2, 0.1))));
    world.add_object(Sphere::new_still(1.0, Vec3(4.0, 1.0, 0.0), MetalBRDF::from_rgb(Vec3(0.7, 0.6, 0.5), 0.1)));
    auto checker = CheckerTexture::make(0.62, SolidTexture::make(Vec3(0.9, 0.0, 0.1)), SolidTexture::make(Vec3(0.9, 0.9, 0.9)));
    world.add_object(Sphere::new_still(1000.0, Vec3(0.0, -1000.0, 0.0), DiffuseBRDF::make(checker)));
    world.build_bvh();
    set_camera(b.camera, 16.0 / 9.0, width, spp, 28.0, Vec3(8.8, 2.0, 3.0), Vec3(0, 0, 0), 2.869817807, 2.5);
    b.camera.environment.color = Vec3(0.85, 0.85, 1.0);
    b.output_name = "earth.png";
}

void cornell_walls(World& world, MatPtr right_wall, MatPtr left_wall, MatPtr white) {  // main.rs:140-169 / 545-574
    world.add_object(Quad::make(Vec3(555.0, 0.0, 0.0), Vec3(0.0, 555.0, 0.0), Vec3(0.0, 0.0, 555.0), right_wall));
    world.add_object(Quad::make(Vec3(0.0, 0.0, 0.0), Vec3(0.0, 555.0, 0.0), Vec3(0.0, 0.0, 555.0), left_wall));
    world.add_object(Quad::make(Vec3(0.0, 0.0, 0.0), Vec3(555.0, 0.0, 0.0), Vec3(0.0, 0.0, 555.0), white));
    world.add_object(Quad::make(Vec3(555.0, 555.0, 555.0), Vec3(-555.0, 0.0, 0.0), Vec3(0.0, 0.0, -555.0), white));
    world.add_object(Quad::make(Vec3(0.0, 0.0, 555.0), Vec3(555.0, 0.0, 0.0), Vec3(0.0, 555.0, 0.0), white));
}
void cornell_camera(Camera& c, uint32_t width, uint32_t spp) {  // main.rs:217-232 / 599-614
    set_camera(c, 1.0, width, spp, 40.0, Vec3(278.0, 278.0, -800.0), Vec3(278.0, 278.0, 0.0), 10.0, 0.0);
    c.environment.color = Vec3(0, 0, 0);
}
void cornell_box_scene(SceneBundle& b, uint32_t width, uint32_t spp) {  // main.rs:134-236
    World& world = b.world;
    auto red = DiffuseBRDF::from_rgb(Vec3(0.65, 0.05, 0.05));
    auto white = DiffuseBRDF::from_rgb(Vec3(0.73, 0.73, 0.73));
    auto green = DiffuseBRDF::from_rgb(Vec3(0.12, 0.45, 0.15));
    cornell_walls(world, green, red, white);
    world.add_light(Quad::make(Vec3(343.0, 554.0, 332.0), Vec3(-130.0, 0.0, 0.0), Vec3(0.0, 0.0, -105.0), DiffuseLight::from_rgb(Vec3(25.0, 25.0, 25.0))));
    world.add_object(Sphere::new_still(135.0, Vec3(113.0, 170.0, 372.0),
                                       principled(Vec3(1, 1, 1), 0.01, 0.01, 0.01, 0.91, 0.91, 1.5, 0.91, 0.91, 0.91, 0.91, 0.01)));
    auto box1 = Cuboid::make(Vec3(0, 0, 0), Vec3(165.0, 330.0, 165.0), MetalBRDF::from_rgb(Vec3(1, 1, 1), 0.1));
    world.add_object(Instance::make(box1, Vec3(0, 1, 0), 0.261799, Vec3(265.0, 0.0, 295.0)));
    auto box2 = Cuboid::make(Vec3(0, 0, 0), Vec3(165.0, 165.0, 165.0), white);
    world.add_object(Instance::make(box2, Vec3(0, 1, 0), -0.29, Vec3(130.0, 0.0, 65.0)));
    world.build_bvh();
    cornell_camera(b.camera, width, spp);
    b.output_name = "cornell.png";
}

void environment_map_scene(SceneBundle& b, uint32_t width, uint32_t spp, const std::string& assets) {  // main.rs:238-274
    World& world = b.world;
    world.add_object(Sphere::new_still(9.0, Vec3(4.0, 2.0, 0.0), MetalBRDF::from_rgb(Vec3(1, 1, 1), 0.001)));
    world.add_object(Quad::make(Vec3(-2.0, 6.5, 0.0), Vec3(4.0, 0.0, 0.0), Vec3(0.0, 0.0, 2.0), DiffuseLight::from_rgb(Vec3(10.0, 10.0, 10.0))));
    world.build_bvh();
    set_camera(b.camera, 16.0 / 9.0, width, spp, 90.0, Vec3(0.0, 3.0, 17.0), Vec3(0.0, 2.0, 0.0), 17.0, 1.5);
    b.camera.environment.is_map = true;
    b.camera.environment.map = ImageTexture::load(asset(assets, "grace_probe_latlong.hdr", "grace_probe_latlong.png"));
    b.output_name = "lights.png";
}

void bsdf_demo_scene(SceneBundle& b, uint32_t width, uint32_t spp, const std::string& assets) {  // main.rs:276-369
    World& world = b.world;
    for (int i = 0; i < 5; i++) {
        double roughness = 0.1 + 0.2 * (double)i;
        world.add_object(Sphere::new_still(0.5, Vec3(-4.0 + (double)i, 1.0, -5.0),
                                           principled(Vec3(0.65, 0.05, 0.05), 0.00, roughness, 0.01, 0.01, 0.01, 1.5, 0.01, 0.01, 0.01, 0.01, 0.01)));
    }
    for (int i = 0; i < 5; i++) {
        double roughness = 0.1 + 0.2 * (double)i;
        world.add_object(Sphere::new_still(0.5, Vec3(-4.0 + (double)i, 2.0, -5.0),
                                           principled(Vec3(0.05, 0.65, 0.05), 0.99, roughness, 0.01, 0.01, 0.01, 1.5, 0.01, 0.01, 0.01, 0.01, 0.01)));
    }
    for (int i = 0; i < 5; i++) {
        double roughness = (0.1 + 0.2 * (double)i) * 0.3;
        world.add_object(Sphere::new_still(0.5, Vec3(-4.0 + (double)i, 3.0, -5.0),
                                           principled(Vec3(0.25, 0.05, 0.65), 0.01, roughness, 0.01, 0.01, 0.01, 1.5, 0.99, 0.01, 0.01, 0.01, 0.01)));
    }
    world.build_bvh();
    Vec3 from(-2.0, 2.0, -1.0);
    set_camera(b.camera, 16.0 / 9.0, width, spp, 60.0, from, from + Vec3(0.0, 0.0, -1000.0), 5.0, 0.0);
    b.camera.environment.is_map = true;
    b.camera.environment.map = g_envmap_override ? g_envmap_override : ImageTexture::load(assets + "/envmap.jpg");  // main.rs:365
    b.output_name = "bsdf.png";
}

void everything_scene(SceneBundle& b, uint32_t width, uint32_t spp, const std::string& assets) {  // main.rs:371-532
    World& world = b.world;
    auto checker = CheckerTexture::make(0.92, SolidTexture::make(Vec3(0.2, 0.3, 0.1)), SolidTexture::make(Vec3(0.9, 0.9, 0.9)));
    world.add_object(Quad::make(Vec3(-1000.0, 0.0, -1000.0), Vec3(0.0, 0.0, 5000.0), Vec3(5000.0, 0.0, 0.0), DiffuseBRDF::from_textures(checker, nullptr)));
    world.add_object(Sphere::new_still(2.0, Vec3(-4.0, 2.0, 9.8), MetalBRDF::from_rgb(Vec3(1, 1, 1), 0.001)));
    world.add_object(Sphere::new_still(1.0, Vec3(4.0, 1.0, 6.0), GlassBSDF::basic(1.5)));
    auto box1 = Cuboid::make(Vec3(0, 0, 0), Vec3(1.0, 2.0, 1.0), DiffuseBRDF::from_rgb(Vec3(0.0, 0.5, 1.0)));
    world.add_object(Instance::make(box1, Vec3(0, 1, 0), 0.5, Vec3(1.2, 0.0, 6.0)));
    auto bunny = TriangleMesh::from_obj(10.0, ObjMesh::load(asset(assets, "bunny.obj", "bunny.mesh")),
                                        principled(Vec3(1, 1, 1), 0.91, 0.01, 0.01, 0.01, 0.91, 1.5, 0.01, 0.91, 0.91, 0.91, 0.01));
    world.add_object(Instance::make(bunny, Vec3(0, 1, 0), 3.14, Vec3(0.1, -0.327, 5.0)));
    auto spot = TriangleMesh::from_obj(0.65, ObjMesh::load(asset(assets, "spot.obj", "spot.mesh")),
                                       principled(Vec3(0.65, 0.05, 0.05), 0.01, 0.01, 0.91, 0.01, 0.01, 1.5, 0.01, 0.91, 0.91, 0.91, 0.01));
    world.add_object(Instance::make(spot, Vec3(0, 1, 0), 0.87, Vec3(-1.5, 2.8, 4.3)));
    auto cow = TriangleMesh::from_obj(0.75, ObjMesh::load(asset(assets, "cow.obj", "cow.mesh")),
                                      principled(Vec3(0.05, 0.65, 0.05), 0.91, 0.21, 0.91, 0.01, 0.01, 1.5, 0.01, 0.91, 0.91, 0.91, 0.01));
    world.add_object(Instance::make(cow, Vec3(0, 1, 0), 0.93, Vec3(2.5, 3.8, 12.0)));
    world.add_object(Sphere::new_still(0.1, Vec3(1.0, 0.1, 3.0), DiffuseLight::from_rgb(Vec3(20.0, 20.0, 10.0))));
    world.add_object(Sphere::new_still(0.2, Vec3(0.0, 0.2, 3.0), MetalBRDF::from_rgb(Vec3(0.6, 0.05, 0.05), 0.1)));
    world.add_object(Sphere::new_still(0.3, Vec3(1.2, 0.3, 3.4), GlassBSDF::make(SolidTexture::make(Vec3(0.7, 0.3, 0.3)), SolidTexture::scalar(0.3), 0.0, 1.5)));
    world.build_bvh();
    set_camera(b.camera, 16.0 / 9.0, width, spp, 60.0, Vec3(0.0, 1.5, 0.0), Vec3(0.0, 1.5, 100000.0), 6.0, 1.0);
    b.camera.environment.is_map = true;
    b.camera.environment.map = ImageTexture::load(asset(assets, "grace_probe_latlong.hdr", "grace_probe_latlong.png"));
    b.output_name = "scene6.png";
}

void normal_demo_scene(SceneBundle& b, uint32_t width, uint32_t spp, const std::string& assets) {  // main.rs:534-618
    World& world = b.world;
    auto albedo = ImageTexture::make(ImageTexture::load(asset(assets, "bricks/color.png", "bricks_color.png")));
    auto normal = ImageTexture::load(asset(assets, "bricks/normal.png", "bricks_normal.png"));
    auto with_normal = DiffuseBRDF::from_textures(albedo, normal);
    auto without_normal = DiffuseBRDF::from_textures(albedo, nullptr);
    auto white = DiffuseBRDF::from_rgb(Vec3(0.73, 0.73, 0.73));
    cornell_walls(world, without_normal, with_normal, white);
    world.add_light(Quad::make(Vec3(343.0, 554.0, 332.0), Vec3(-130.0, 0.0, 0.0), Vec3(0.0, 0.0, -105.0), DiffuseLight::from_rgb(Vec3(27.0, 28.0, 20.0))));
    auto box1 = Cuboid::make(Vec3(0, 0, 0), Vec3(165.0, 330.0, 165.0), MetalBRDF::from_rgb(Vec3(0.94, 0.94, 0.94), 0.1));
    world.add_object(Instance::make(box1, Vec3(0, 1, 0), 0.261799, Vec3(265.0, 0.0, 295.0)));
    world.add_object(Sphere::new_still(100.0, Vec3(130.0, 100.0, 65.0), GlassBSDF::basic(1.5)));
    world.build_bvh();
    cornell_camera(b.camera, width, spp);
    b.output_name = "normals.png";
}

// Scene "7m" (ours, measurement only — SURVEY §0.1): the scene-7 Cornell box plus bunny and teapot meshes,
// the traversal-bound case BASELINE.json's configs[3] describes but the reference never built.
void normal_demo_mesh_scene(SceneBundle& b, uint32_t width, uint32_t spp, const std::string& assets) {
    normal_demo_scene(b, width, spp, assets);
    World& world = b.world;
    auto bunny = TriangleMesh::from_obj(1400.0, ObjMesh::load(asset(assets, "bunny.obj", "bunny.mesh")),
                                        principled(Vec3(0.9, 0.7, 0.3), 0.91, 0.2, 0.01, 0.5, 0.01, 1.5, 0.01, 0.01, 0.01, 0.5, 0.5));
    world.add_object(Instance::make(bunny, Vec3(0, 1, 0), 3.3, Vec3(400.0, 280.0, 330.0)));
    auto teapot = TriangleMesh::from_obj(28.0, ObjMesh::load(asset(assets, "teapot.obj", "teapot.mesh")), DiffuseBRDF::from_rgb(Vec3(0.3, 0.4, 0.8)));
    world.add_object(Instance::make(teapot, Vec3(0, 1, 0), 0.6, Vec3(380.0, 0.0, 120.0)));
    world.build_bvh();
    b.output_name = "normals_mesh.png";
}
}  // namespace

std::unique_ptr<SceneBundle> build_scene(int scene, uint32_t width, uint32_t spp, uint64_t seed, const std::string& assets) {
    auto b = std::make_unique<SceneBundle>();
    switch (scene) {  // main.rs:635-644
        case 1: balls_scene(*b, width, spp, seed); break;
        case 2: earth_scene(*b, width, spp, assets); break;
        case 3: cornell_box_scene(*b, width, spp); break;
        case 4: environment_map_scene(*b, width, spp, assets); break;
        case 5: bsdf_demo_scene(*b, width, spp, assets); break;
        case 6: everything_scene(*b, width, spp, assets); break;
        case 7: normal_demo_scene(*b, width, spp, assets); break;
        case 70: normal_demo_mesh_scene(*b, width, spp, assets); break;
        default: throw std::runtime_error("unknown scene");
    }
    b->camera.init();
    return b;
}

}  // namespace pt
