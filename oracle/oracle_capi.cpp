// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).  PARITY UNPINNED except via demo/*.png.
//
// C entry points (ctypes-friendly) over the restated reference.  The oracle consumes the same
// pt_scene_desc the product does, but rebuilds every derived quantity itself from the constructor
// arguments (quad w/normal/d, cuboid sides, instance matrices, all BVHs via the restated bvh.rs), so a
// wrong host-side derivation shows up as a mismatch rather than being inherited.
#include <omp.h>

#include <chrono>
#include <cstring>
#include <string>

#include "../include/pt_b200.h"
#include "oracle_camera.hpp"

using namespace orc;

struct orc_scene {
    std::vector<std::unique_ptr<Texture>> textures;
    std::vector<std::unique_ptr<ImageTexture>> image_tex;  // one per pt_image (normal maps, env maps)
    std::vector<std::unique_ptr<BxDF>> materials;
    std::vector<std::shared_ptr<TriangleMesh>> meshes;
    std::vector<std::vector<uint8_t>> image_store;
    World world;
    std::string error;
    EnvDist env; uint32_t env_image = UINT32_MAX;  // orc_scene_build_env_sampler
};

static Vec3 V(const pt_vec3& v) { return Vec3(v.x, v.y, v.z); }
static pt_vec3 P(const Vec3& v) { return pt_vec3{v.x, v.y, v.z}; }
static thread_local std::string g_err;

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

static HitPtr build_ref(orc_scene* s, const pt_scene_desc* d, pt_ref r, bool allow_instance) {
    switch (r.kind) {
        case PT_PRIM_SPHERE: {
            const pt_sphere& sp = d->spheres[r.index];
            return std::make_shared<Sphere>(sp.radius, V(sp.position1), V(sp.position2), sp.is_moving != 0,
                                            s->materials[sp.material].get(), r.index);
        }
        case PT_PRIM_QUAD: {
            const pt_quad& q = d->quads[r.index];
            return std::make_shared<Quad>(V(q.q), V(q.u), V(q.v), s->materials[q.material].get(), r.index);
        }
        case PT_OBJ_CUBOID: {
            const pt_cuboid& c = d->cuboids[r.index];
            return std::make_shared<Cuboid>(V(c.a), V(c.b), s->materials[c.material].get(), c.first_quad);
        }
        case PT_OBJ_MESH:
            return s->meshes[r.index];
        case PT_OBJ_INSTANCE: {
            if (!allow_instance) return nullptr;
            const pt_instance& in = d->instances[r.index];
            HitPtr child = build_ref(s, d, in.child, false);
            if (!child) return nullptr;
            return std::make_shared<Instance>(child, V(in.axis), in.angle, V(in.translation), r.index);
        }
        case PT_OBJ_VOLUME: {
            if (d->abi_version < 2 || r.index >= d->n_volumes) return nullptr;
            const pt_volume& v = d->volumes[r.index];
            if (v.boundary.kind != PT_PRIM_SPHERE && v.boundary.kind != PT_OBJ_CUBOID) return nullptr;
            HitPtr boundary = build_ref(s, d, v.boundary, false);
            if (!boundary) return nullptr;
            return std::make_shared<HomogeneousVolume>(boundary, v.density, s->materials[v.material].get(), r.index);
        }
        default:
            return nullptr;
    }
}

int orc_scene_create(const pt_scene_desc* d, orc_scene** out) {
    auto s = std::make_unique<orc_scene>();
    // images are copied so the scene owns its data
    for (uint32_t i = 0; i < d->n_images; i++) {
        const pt_image& im = d->images[i];
        s->image_store.emplace_back(im.rgb, im.rgb + (size_t)3 * im.width * im.height);
        s->image_tex.push_back(std::make_unique<ImageTexture>(ImageData{s->image_store.back().data(), im.width, im.height}));
    }
    s->textures.resize(d->n_textures);
    // children may appear after parents: resolve in two passes
    for (int pass = 0; pass < 2; pass++) {
        for (uint32_t i = 0; i < d->n_textures; i++) {
            const pt_texture& t = d->textures[i];
            if (pass == 0 && t.kind == PT_TEX_SOLID) s->textures[i] = std::make_unique<SolidTexture>(V(t.value));
            if (pass == 0 && t.kind == PT_TEX_IMAGE)
                s->textures[i] = std::make_unique<ImageTexture>(ImageData{s->image_store[t.image].data(), d->images[t.image].width, d->images[t.image].height});
        }
    }
    // checkers can nest: iterate until all resolved
    for (uint32_t iter = 0; iter <= d->n_textures; iter++) {
        bool progress = false, pending = false;
        for (uint32_t i = 0; i < d->n_textures; i++) {
            const pt_texture& t = d->textures[i];
            if (t.kind != PT_TEX_CHECKER || s->textures[i]) continue;
            if (s->textures[t.tex1] && s->textures[t.tex2]) {
                s->textures[i] = std::make_unique<CheckerTexture>(t.inv_scale, s->textures[t.tex1].get(), s->textures[t.tex2].get());
                progress = true;
            } else pending = true;
        }
        if (!pending) break;
        if (!progress) { g_err = "checker texture cycle"; return -1; }
    }
    s->materials.resize(d->n_materials);
    for (int pass = 0; pass < 2; pass++) {
        for (uint32_t i = 0; i < d->n_materials; i++) {
            const pt_material& m = d->materials[i];
            auto tex = [&](uint32_t k) -> const Texture* { return k == PT_NONE ? nullptr : s->textures[k].get(); };
            std::unique_ptr<BxDF> b;
            if (pass == 0) {
                switch (m.kind) {
                    case PT_MAT_DIFFUSE:
                        b = std::make_unique<DiffuseBRDF>(tex(m.base_color_tex), m.normal_map == PT_NONE ? nullptr : s->image_tex[m.normal_map].get());
                        break;
                    case PT_MAT_METAL: b = std::make_unique<MetalBRDF>(tex(m.base_color_tex), tex(m.roughness_tex)); break;
                    case PT_MAT_GLASS: b = std::make_unique<GlassBSDF>(tex(m.base_color_tex), tex(m.roughness_tex), m.p[PT_P_IOR]); break;
                    case PT_MAT_PRINCIPLED: {
                        auto p = std::make_unique<PrincipledBSDF>();
                        p->base_color = tex(m.base_color_tex);
                        p->metallic = m.p[PT_P_METALLIC]; p->roughness = m.p[PT_P_ROUGHNESS]; p->subsurface = m.p[PT_P_SUBSURFACE];
                        p->specular = m.p[PT_P_SPECULAR]; p->specular_tint = m.p[PT_P_SPECULAR_TINT]; p->ior = m.p[PT_P_IOR];
                        p->spec_trans = m.p[PT_P_SPEC_TRANS]; p->sheen = m.p[PT_P_SHEEN]; p->sheen_tint = m.p[PT_P_SHEEN_TINT];
                        p->clearcoat = m.p[PT_P_CLEARCOAT]; p->clearcoat_gloss = m.p[PT_P_CLEARCOAT_GLOSS];
                        b = std::move(p);
                        break;
                    }
                    case PT_MAT_LIGHT: b = std::make_unique<DiffuseLight>(tex(m.base_color_tex)); break;
                    case PT_MAT_SHEEN:
                        b = std::make_unique<SheenBRDF>(Vec3(m.p[PT_P_COLOR_R], m.p[PT_P_COLOR_G], m.p[PT_P_COLOR_B]), m.p[PT_P_SHEEN_TINT]);
                        break;
                    case PT_MAT_CLEARCOAT: b = std::make_unique<ClearcoatBRDF>(m.p[PT_P_ALPHA_G]); break;
                    case PT_MAT_ISOTROPIC: b = std::make_unique<IsotropicMaterial>(tex(m.base_color_tex)); break;
                    case PT_MAT_MIX: break;
                    default: g_err = "unknown material kind"; return -1;
                }
            } else if (m.kind == PT_MAT_MIX) {
                // children must precede the mix (builders create them first)
                if (m.mix_a >= i || m.mix_b >= i || !s->materials[m.mix_a] || !s->materials[m.mix_b]) { g_err = "mix children must precede the mix material"; return -1; }
                b = std::make_unique<MixBxDf>(m.p[PT_P_MIX_T], s->materials[m.mix_a].get(), s->materials[m.mix_b].get());
            }
            if (b) { b->index = i; s->materials[i] = std::move(b); }
        }
    }
    // meshes: triangles list + own BVH (mesh.rs:172-196)
    for (uint32_t mi = 0; mi < d->n_meshes; mi++) {
        const pt_mesh& m = d->meshes[mi];
        auto mesh = std::make_shared<TriangleMesh>();
        for (uint32_t k = 0; k < m.n_triangles; k++) {
            uint32_t ti = m.first_triangle + k;
            const pt_triangle& t = d->triangles[ti];
            auto tri = std::make_shared<Triangle>(V(t.v0), V(t.v1), V(t.v2), s->materials[m.material].get(), ti);
            if (m.has_normals) { tri->has_n = true; for (int j = 0; j < 3; j++) tri->n[j] = V(d->tri_normals[3 * ti + j]); }
            if (m.has_uvs) { tri->has_uv = true; for (int j = 0; j < 3; j++) { tri->uv[j][0] = d->tri_uvs[6 * ti + 2 * j]; tri->uv[j][1] = d->tri_uvs[6 * ti + 2 * j + 1]; } }
            mesh->triangles.add(tri);
        }
        if (m.bvh_root != PT_NONE) mesh->triangles.build_bvh();
        s->meshes.push_back(mesh);
    }
    for (uint32_t i = 0; i < d->n_objects; i++) {
        HitPtr h = build_ref(s.get(), d, d->objects[i], true);
        if (!h) { g_err = "bad object ref"; return -1; }
        s->world.objects.add(h);
    }
    for (uint32_t i = 0; i < d->n_lights; i++) {
        HitPtr h = build_ref(s.get(), d, d->lights[i], true);
        if (!h) { g_err = "bad light ref"; return -1; }
        s->world.lights.add(h);
    }
    if (d->objects_bvh_root != PT_NONE) s->world.objects.build_bvh();
    if (d->lights_bvh_root != PT_NONE) s->world.lights.build_bvh();
    *out = s.release();
    return 0;
}
void orc_scene_destroy(orc_scene* s) { delete s; }

// ---- derived-quantity readback (to cross-check the host's restatement) -------------------------
// which: 0 = objects list, 1 = lights list
static const Hittable* top_item(const orc_scene* s, int which, uint32_t i) {
    const HittableList& l = which ? s->world.lights : s->world.objects;
    return i < l.objects.size() ? l.objects[i].get() : nullptr;
}
int orc_item_bbox(const orc_scene* s, int which, uint32_t i, double out6[6]) {
    const Hittable* h = top_item(s, which, i);
    if (!h) return -1;
    AABB b = h->bounding_box();
    out6[0] = b.min.x; out6[1] = b.min.y; out6[2] = b.min.z; out6[3] = b.max.x; out6[4] = b.max.y; out6[5] = b.max.z;
    return 0;
}
int orc_quad_derived(const orc_scene* s, int which, uint32_t i, double out7[7]) {
    auto q = dynamic_cast<const Quad*>(top_item(s, which, i));
    if (!q) return -1;
    out7[0] = q->w.x; out7[1] = q->w.y; out7[2] = q->w.z; out7[3] = q->normal.x; out7[4] = q->normal.y; out7[5] = q->normal.z; out7[6] = q->d;
    return 0;
}
int orc_instance_matrices(const orc_scene* s, int which, uint32_t i, double out48[48]) {
    auto in = dynamic_cast<const Instance*>(top_item(s, which, i));
    if (!in) return -1;
    memcpy(out48, in->transform.c, 128); memcpy(out48 + 16, in->inverse.c, 128); memcpy(out48 + 32, in->normal_mat.c, 128);
    return 0;
}
// BVH signature, DFS pre-order: internal -> -1; leaf -> n followed by the n list positions of its items.
static void sig_rec(const BVHNode* n, std::vector<int64_t>& out, std::vector<double>* boxes) {
    if (boxes) { boxes->push_back(n->bbox.min.x); boxes->push_back(n->bbox.min.y); boxes->push_back(n->bbox.min.z);
                 boxes->push_back(n->bbox.max.x); boxes->push_back(n->bbox.max.y); boxes->push_back(n->bbox.max.z); }
    if (n->is_leaf) { out.push_back((int64_t)n->item_ids.size()); for (auto id : n->item_ids) out.push_back(id); }
    else { out.push_back(-1); sig_rec(n->left.get(), out, boxes); sig_rec(n->right.get(), out, boxes); }
}
// which: 0 objects, 1 lights, 2+m mesh m.  Returns the number of int64 written (or needed if cap too small).
int64_t orc_bvh_signature(const orc_scene* s, int which, int64_t* out, int64_t cap, double* boxes6, int64_t boxes_cap) {
    const BVHNode* root = which == 0 ? s->world.objects.bvh.get() : which == 1 ? s->world.lights.bvh.get()
                          : (size_t)(which - 2) < s->meshes.size() ? s->meshes[which - 2]->triangles.bvh.get() : nullptr;
    if (!root) return 0;
    std::vector<int64_t> v; std::vector<double> b;
    sig_rec(root, v, boxes6 ? &b : nullptr);
    if ((int64_t)v.size() <= cap) memcpy(out, v.data(), v.size() * 8);
    if (boxes6 && (int64_t)b.size() <= boxes_cap) memcpy(boxes6, b.data(), b.size() * 8);
    return (int64_t)v.size();
}

// ---- camera ------------------------------------------------------------------------------------
static Camera make_camera(const orc_scene* s, const pt_camera* c) {
    Camera cam;
    cam.aspect_ratio = c->aspect_ratio; cam.image_width = c->image_width; cam.samples_per_pixel = c->samples_per_pixel;
    cam.max_depth = c->max_depth; cam.vfov = c->vfov; cam.look_from = V(c->look_from); cam.look_at = V(c->look_at);
    cam.vup = V(c->vup); cam.blur_strength = c->blur_strength; cam.focal_length = c->focal_length;
    cam.defocus_angle = c->defocus_angle; cam.env_is_map = c->env_is_map != 0; cam.env_color = V(c->env_color);
    if (cam.env_is_map && s) cam.env_map = s->image_tex[c->env_image].get();
    cam.init();
    return cam;
}
uint32_t orc_camera_image_height(const pt_camera* c) { return make_camera(nullptr, c).image_height; }

int orc_camera_rays(const pt_camera* c, uint64_t seed, size_t n, const uint32_t* row, const uint32_t* col,
                    const uint32_t* sample, pt_ray* out) {
    pt_camera cc = *c; cc.env_is_map = 0;
    Camera cam = make_camera(nullptr, &cc);
    for (size_t i = 0; i < n; i++) {
        Rng rng; rng.seed = seed; rng.pixel = row[i] * cam.image_width + col[i]; rng.sample = sample[i];
        Ray r = cam.generate_ray(row[i], col[i], rng);
        out[i].origin = P(r.origin); out[i].direction = P(r.direction); out[i].time = r.time;
    }
    return 0;
}

// ---- closest hit -------------------------------------------------------------------------------
static void fill_hit(const World& w, const std::optional<std::pair<HitInfo, bool>>& h, pt_hit* o) {
    memset(o, 0, sizeof(*o));
    o->instance = PT_NONE;
    if (!h) return;
    const HitInfo& i = h->first;
    o->hit = 1; o->t = i.dist; o->u = i.u; o->v = i.v; o->point = P(i.point); o->geometric_normal = P(i.geometric_normal);
    o->shading_normal = P(i.shading_normal); o->prim_kind = i.prim_kind; o->prim_index = i.prim_index; o->instance = i.instance;
    o->material = i.mat->index; o->front_face = i.front_face; o->is_light = h->second;
    (void)w;
}
int orc_trace_closest(const orc_scene* s, size_t n, const pt_ray* rays, double t_min, pt_hit* hits) {
#pragma omp parallel for schedule(dynamic, 256)
    for (size_t i = 0; i < n; i++) {
        Ray r{V(rays[i].origin), V(rays[i].direction), rays[i].time};
        g_path_key = PathKey{0, (uint32_t)i, 0, 0};  // keyed uniforms of ray batches (include/pt_b200.h, pt_volume)
        fill_hit(s->world, s->world.intersect_all(r, Interval{t_min, INF}), &hits[i]);
    }
    return 0;
}
int orc_trace_any(const orc_scene* s, size_t n, const pt_ray* rays, double t_min, const double* t_max, uint8_t* occluded) {
#pragma omp parallel for schedule(dynamic, 256)
    for (size_t i = 0; i < n; i++) {
        Ray r{V(rays[i].origin), V(rays[i].direction), rays[i].time};
        g_path_key = PathKey{0, (uint32_t)i, 0, 0};
        occluded[i] = s->world.occluded(r, Interval{t_min, t_max[i]}) ? 1 : 0;
    }
    return 0;
}

// ---- BSDF --------------------------------------------------------------------------------------
static HitInfo info_from_query(const orc_scene* s, uint32_t material, const pt_bsdf_query& q) {
    HitInfo h;
    h.point = V(q.point); h.geometric_normal = V(q.geometric_normal); h.shading_normal = V(q.shading_normal);
    h.u = q.u; h.v = q.v; h.front_face = q.front_face != 0; h.mat = s->materials[material].get();
    return h;
}
int orc_bsdf_eval_pdf(const orc_scene* s, uint32_t material, size_t n, const pt_bsdf_query* q, pt_bsdf_result* out) {
    if (material >= s->materials.size()) return -1;
    for (size_t i = 0; i < n; i++) {
        HitInfo h = info_from_query(s, material, q[i]);
        out[i].eval = P(h.mat->eval(V(q[i].view_dir), V(q[i].light_dir), h));
        out[i].pdf = h.mat->pdf(V(q[i].view_dir), V(q[i].light_dir), h);
        out[i].emitted = P(h.mat->emitted(h.u, h.v, h.point));
        out[i]._pad = 0;
    }
    return 0;
}
int orc_bsdf_sample(const orc_scene* s, uint32_t material, size_t n, const pt_bsdf_query* q, const double* uniforms8,
                    pt_bsdf_sample_result* out) {
    if (material >= s->materials.size()) return -1;
    for (size_t i = 0; i < n; i++) {
        HitInfo h = info_from_query(s, material, q[i]);
        Rng rng; rng.arr = uniforms8 + 8 * i; rng.arr_n = 8;
        Vec3 vd = V(q[i].view_dir);
        Ray ray{h.point, -vd, 0.0};  // BxDF::sample reads only ray.direction (= -view_dir)
        auto d = h.mat->sample(ray, h, rng);
        out[i].valid = d.has_value(); out[i].dir = d ? P(*d) : pt_vec3{0, 0, 0}; out[i].n_uniforms = rng.used;
    }
    return 0;
}
int orc_lights_sample_pdf(const orc_scene* s, size_t n, const pt_vec3* origin, const double* time, const double* uniforms4,
                          pt_vec3* dir, uint32_t* valid, double* pdf) {
    for (size_t i = 0; i < n; i++) {
        Rng rng; rng.arr = uniforms4 + 4 * i; rng.arr_n = 4;
        auto d = s->world.lights.sample(V(origin[i]), time[i], rng);
        valid[i] = d.has_value(); dir[i] = d ? P(*d) : pt_vec3{0, 0, 0};
        pdf[i] = d ? s->world.lights.pdf(V(origin[i]), *d, time[i]) : 0.0;
    }
    return 0;
}

// ---- environment importance sampler (ours, not the reference's: see EnvDist) --------------------------
int orc_scene_build_env_sampler(orc_scene* s, uint32_t image, uint32_t max_rows, uint32_t max_cols) {
    if (image >= s->image_tex.size() || s->image_tex[image]->img.width == 0 || s->image_tex[image]->img.height == 0) { g_err = "bad env image"; return 1; }
    s->env.build(s->image_tex[image]->img, max_rows, max_cols);
    s->env_image = image;
    return 0;
}
int orc_env_sample_pdf(const orc_scene* s, size_t n, const double* uniforms2, pt_vec3* dir, double* pdf) {
    if (s->env_image == UINT32_MAX) { g_err = "no env sampler"; return 1; }
    for (size_t i = 0; i < n; i++) {
        Vec3 d = s->env.sample(uniforms2[2 * i], uniforms2[2 * i + 1]);
        dir[i] = P(d); pdf[i] = s->env.pdf(d);
    }
    return 0;
}

// ---- render (Camera::render, camera.rs:79-126, minus PNG) ---------------------------------------
struct orc_stats { uint64_t paths, segments, boxes, spheres, quads, triangles, instances, nonfinite; double seconds; int threads; };

// out_mean: W*H*3 doubles = sum over the rendered samples / sample_count.
int orc_render(const orc_scene* s, const pt_camera* c, const pt_render_params* p, int threads, double* out_mean,
               orc_stats* st) {
    Camera cam = make_camera(s, c);
    if ((p->flags & PT_RENDER_ENV_IMPORTANCE) && c->env_is_map) {
        if (s->env_image != c->env_image) { g_err = "orc_render: build the env sampler first"; return 1; }
        cam.env_dist = &s->env;
    }
    const uint32_t W = cam.image_width, H = cam.image_height;
    if (threads <= 0) threads = omp_get_max_threads();
    Counters total; uint64_t nonfinite = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(threads)
    {
        g_cnt = Counters();
        uint64_t nf = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t px = 0; px < (int64_t)W * H; px++) {
            uint32_t r = (uint32_t)(px / W), col = (uint32_t)(px % W);
            Vec3 color(0, 0, 0);
            for (uint32_t k = 0; k < p->sample_count; k++) {
                Rng rng; rng.seed = p->seed; rng.pixel = (uint32_t)px; rng.sample = p->sample_begin + k * p->sample_stride;
                uint64_t* drop = p->nan_policy == PT_NAN_DROP ? &nf : nullptr;
                Vec3 rad = (p->flags & PT_RENDER_NEE) ? cam.trace_nee(r, col, s->world, rng, drop) : cam.trace(r, col, s->world, rng, nullptr, drop);
                bool fin = std::isfinite(rad.x) && std::isfinite(rad.y) && std::isfinite(rad.z);
                if (!fin) nf++;  // PT_NAN_REFERENCE only: the sample poisons its pixel (camera.rs:129)
                color += rad;
            }
            color *= 1.0 / (double)p->sample_count;
            out_mean[3 * px] = color.x; out_mean[3 * px + 1] = color.y; out_mean[3 * px + 2] = color.z;
        }
#pragma omp critical
        { total.add(g_cnt); nonfinite += nf; }
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (st) {
        st->paths = total.paths; st->segments = total.segments; st->boxes = total.boxes; st->spheres = total.spheres;
        st->quads = total.quads; st->triangles = total.triangles; st->instances = total.instances; st->nonfinite = nonfinite;
        st->seconds = sec; st->threads = threads;
    }
    return 0;
}
// Tonemap exactly like camera.rs:109-114,128-130: sqrt(max(x,0)), clamp(0,0.999)*256 as u8.
void orc_tonemap_rgb8(const double* mean, size_t n_values, uint8_t* out) {
    for (size_t i = 0; i < n_values; i++) {
        double g = std::sqrt(fmax_(mean[i], 0.0));
        double v = clamp_(g, 0.0, 0.999) * 256.0;
        out[i] = std::isnan(v) ? 0 : (v >= 255.0 ? 255 : (v <= 0.0 ? 0 : (uint8_t)v));
    }
}
// Dump the rays a set of paths hands to intersect_all (for incoherent-ray parity batches).
// Writes up to cap rays from paths (pixel stride `pixel_step`, samples 0..spp); returns count.
int64_t orc_dump_path_rays(const orc_scene* s, const pt_camera* c, uint64_t seed, uint32_t pixel_step, uint32_t spp,
                           uint32_t min_bounce, pt_ray* out, int64_t cap) {
    Camera cam = make_camera(s, c);
    int64_t n = 0;
    std::vector<Ray> rec;
    for (uint32_t px = 0; px < cam.image_width * cam.image_height && n < cap; px += pixel_step) {
        for (uint32_t k = 0; k < spp && n < cap; k++) {
            Rng rng; rng.seed = seed; rng.pixel = px; rng.sample = k;
            rec.clear();
            cam.trace(px / cam.image_width, px % cam.image_width, s->world, rng, &rec);
            for (size_t b = min_bounce; b < rec.size() && n < cap; b++) {
                out[n].origin = P(rec[b].origin); out[n].direction = P(rec[b].direction); out[n].time = rec[b].time; n++;
            }
        }
    }
    return n;
}
int orc_num_threads() { return omp_get_max_threads(); }
// RNG contract probes (known-answer tests)
void orc_philox_block(uint32_t k0, uint32_t k1, const uint32_t ctr[4], uint32_t out[4]) { Philox::block(k0, k1, ctr[0], ctr[1], ctr[2], ctr[3], out); }
void orc_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, double* out) {
    Rng rng; rng.seed = seed; rng.pixel = pixel; rng.sample = sample;
    for (uint32_t i = 0; i < n; i++) out[i] = rng.next();
}

}  // extern "C"
