// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).  PARITY UNPINNED except via demo/*.png.
//
// CPU restatement of the reference's geometry + texture layer, mirroring its trait-object structure:
//   src/hittable/{mod,hit_info,aabb,bvh,list,world,sphere,quad,cuboid,instance,mesh}.rs, src/texture.rs
// Every function cites the lines it follows.  Primitive/instance IDs are an addition (the reference's
// HitInfo has none): they index the pt_scene_desc arrays so oracle and device speak the same IDs.
#pragma once
#include <algorithm>
#include <memory>
#include <optional>
#include <vector>

#include "oracle_math.hpp"

namespace orc {

struct Counters {  // per-thread work counters (N_node, N_prim,k)
    uint64_t boxes = 0, spheres = 0, quads = 0, triangles = 0, instances = 0, segments = 0, paths = 0;
    void add(const Counters& o) {
        boxes += o.boxes; spheres += o.spheres; quads += o.quads; triangles += o.triangles;
        instances += o.instances; segments += o.segments; paths += o.paths;
    }
};
inline thread_local Counters g_cnt;

// ---------------------------------------------------------------- textures (src/texture.rs)
struct ImageData { const uint8_t* rgb; uint32_t width, height; };
struct Texture {  // Texture<Vec3>; Texture<f64> keeps its scalar in .x
    virtual ~Texture() = default;
    virtual Vec3 value(double u, double v, const Vec3& p) const = 0;
};
struct SolidTexture : Texture {  // texture.rs:11-25
    Vec3 val;
    explicit SolidTexture(Vec3 v) : val(v) {}
    Vec3 value(double, double, const Vec3&) const override { return val; }
};
struct CheckerTexture : Texture {  // texture.rs:27-54
    double inv_scale; const Texture *tex1, *tex2;
    CheckerTexture(double inv_scale_, const Texture* a, const Texture* b) : inv_scale(inv_scale_), tex1(a), tex2(b) {}
    static int32_t as_i32(double f) {  // Rust `as i32`: saturating, NaN -> 0
        if (std::isnan(f)) return 0;
        if (f >= 2147483647.0) return INT32_MAX;
        if (f <= -2147483648.0) return INT32_MIN;
        return (int32_t)f;
    }
    Vec3 value(double u, double v, const Vec3& p) const override {
        int32_t x = as_i32(std::floor(p.x * inv_scale));
        int32_t y = as_i32(std::floor(p.y * inv_scale));
        int32_t z = as_i32(std::floor(p.z * inv_scale));
        // release-mode i32 add wraps; `%` keeps the sign (so -1 selects tex2)
        int32_t s = (int32_t)((uint32_t)x + (uint32_t)y + (uint32_t)z);
        if (s % 2 == 0) return tex1->value(u, v, p);
        return tex2->value(u, v, p);
    }
};
struct ImageTexture : Texture {  // texture.rs:56-92
    ImageData img;
    explicit ImageTexture(ImageData d) : img(d) {}
    static uint32_t as_u32(double f) {  // Rust `as u32`: saturating, NaN -> 0
        if (!(f > 0.0)) return 0;
        if (f >= 4294967295.0) return UINT32_MAX;
        return (uint32_t)f;
    }
    Vec3 value(double u, double v, const Vec3&) const override {
        if (img.height == 0) return Vec3(0.0, 1.0, 1.0);
        u = clamp_(u, 0.0, 1.0);
        v = 1.0 - clamp_(v, 0.0, 1.0);
        uint32_t i = as_u32(u * (double)img.width);
        uint32_t j = as_u32(v * (double)img.height);
        // Q22: the reference panics when i == width or j == height (get_pixel OOB, measure-zero set);
        // the restatement clamps, as the device does (documented divergence).
        if (i >= img.width) i = img.width - 1;
        if (j >= img.height) j = img.height - 1;
        const uint8_t* px = img.rgb + 3 * ((size_t)j * img.width + i);
        double s = 1.0 / 255.0;
        return Vec3(s * (double)px[0], s * (double)px[1], s * (double)px[2]);
    }
};

// ---------------------------------------------------------------- forward decls
struct BxDF;
struct HitInfo {  // hit_info.rs:4-13 (+ IDs)
    Vec3 point, geometric_normal, shading_normal;
    double dist = 0;
    bool front_face = false;
    const BxDF* mat = nullptr;
    double u = 0, v = 0;
    uint32_t prim_kind = 0, prim_index = 0, instance = 0xFFFFFFFFu;
};

struct BxDF {  // bsdf/mod.rs:21-57
    uint32_t index = 0;  // material index in the scene description
    virtual ~BxDF() = default;
    virtual std::optional<Vec3> sample(const Ray& ray, const HitInfo& info, Rng& rng) const = 0;
    virtual double pdf(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const = 0;
    virtual Vec3 eval(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const = 0;
    virtual Vec3 emitted(double, double, Vec3) const { return Vec3(0, 0, 0); }
    virtual bool is_emitter() const { return false; }  // DiffuseLight only (used by the NEE integrator, ours)
    virtual const ImageTexture* normal_map() const { return nullptr; }
};

// hit_info.rs:58-67
inline void get_tangent_basis(Vec3 n, Vec3& tangent, Vec3& bitangent) {
    Vec3 a = std::fabs(n.x) > 0.9 ? Vec3(0.0, 1.0, 0.0) : Vec3(1.0, 0.0, 0.0);
    tangent = normalize(cross(n, a));
    bitangent = cross(n, tangent);
}
// hit_info.rs:16-55
inline HitInfo make_hit_info(const Ray& ray, Vec3 point, Vec3 geometric_normal, double dist, const BxDF* mat,
                             double u, double v) {
    bool front_face = dot(ray.direction, geometric_normal) < 0.0;
    Vec3 gn = front_face ? normalize(geometric_normal) : -normalize(geometric_normal);
    Vec3 sn;
    if (const ImageTexture* nm = mat->normal_map()) {
        Vec3 c = nm->value(u, v, point);
        Vec3 mapped = 2.0 * c - Vec3(1.0, 1.0, 1.0);
        Vec3 t, b;
        get_tangent_basis(gn, t, b);
        sn = normalize(mapped.x * t + mapped.y * b + mapped.z * gn);
    } else {
        sn = gn;
    }
    HitInfo h;
    h.point = point; h.geometric_normal = gn; h.shading_normal = sn; h.dist = dist;
    h.front_face = front_face; h.mat = mat; h.u = u; h.v = v;
    return h;
}

// ---------------------------------------------------------------- AABB (src/hittable/aabb.rs)
struct AABB {
    Vec3 min{INF, INF, INF}, max{-INF, -INF, -INF};  // Default: aabb.rs:81-88
    static AABB make(Vec3 a, Vec3 b) {               // aabb.rs:16-21 (pads on every construction, Q1)
        Vec3 delta = Vec3::splat(1e-3);
        AABB r; r.min = vmin(a, b) - delta; r.max = vmax(a, b) + delta; return r;
    }
    AABB unite(const AABB& o) const { return make(vmin(min, o.min), vmax(max, o.max)); }  // aabb.rs:23-25
    Vec3 centroid() const { return 0.5 * (min + max); }                                    // aabb.rs:27-29
    bool intersects(const Ray& ray, Interval ray_t) const {                                // aabb.rs:31-42
        g_cnt.boxes++;
        Vec3 m = recip(ray.direction);
        Vec3 t1 = (min - ray.origin) * m;
        Vec3 t2 = (max - ray.origin) * m;
        double t_near = max_element(vmin(t1, t2));
        double t_far = min_element(vmax(t1, t2));
        return t_near <= t_far && t_far >= ray_t.min && t_near <= ray_t.max;
    }
    double surface_area() const {  // aabb.rs:49-52 (half area)
        Vec3 e = max - min;
        return e.x * e.y + e.x * e.z + e.y * e.z;
    }
    AABB transform(const Mat4& mat) const {  // aabb.rs:54-78
        Vec3 corners[8] = {min, Vec3(min.x, min.y, max.z), Vec3(min.x, max.y, min.z), Vec3(min.x, max.y, max.z),
                           Vec3(max.x, min.y, min.z), Vec3(max.x, min.y, max.z), Vec3(max.x, max.y, min.z), max};
        Vec3 nmin = Vec3::splat(INF), nmax = Vec3::splat(-INF);
        for (auto& c : corners) {
            Vec3 t = transform_point3(mat, c);
            nmin = vmin(nmin, t); nmax = vmax(nmax, t);
        }
        return make(nmin, nmax);
    }
};

// ---------------------------------------------------------------- Hittable (src/hittable/mod.rs:38-48)
struct Hittable {
    virtual ~Hittable() = default;
    virtual std::optional<HitInfo> intersects(const Ray& ray, Interval ray_t) const = 0;
    virtual AABB bounding_box() const = 0;
    virtual std::optional<Vec3> sample(Vec3 origin, double time, Rng& rng) const = 0;
    virtual double pdf(Vec3 origin, Vec3 direction, double time) const = 0;
};
using HitPtr = std::shared_ptr<const Hittable>;

// ---------------------------------------------------------------- BVH (src/hittable/bvh.rs)
struct BVHNode : Hittable {
    bool is_leaf = true;
    AABB bbox;
    std::vector<HitPtr> hittables;           // Leaf
    std::vector<uint32_t> item_ids;          // list positions of the leaf items (for tree signatures)
    std::unique_ptr<BVHNode> left, right;    // Internal

    std::optional<HitInfo> intersects(const Ray& ray, Interval ray_t) const override {  // bvh.rs:124-164
        if (!bbox.intersects(ray, ray_t)) return std::nullopt;
        if (is_leaf) {
            std::optional<HitInfo> hit;
            double closest = ray_t.max;
            for (auto& p : hittables) {
                if (auto info = p->intersects(ray, Interval{ray_t.min, closest})) {
                    closest = info->dist;
                    hit = info;
                }
            }
            return hit;
        }
        bool lh = left->bbox.intersects(ray, ray_t);
        bool rh = right->bbox.intersects(ray, ray_t);
        if (!lh && !rh) return std::nullopt;
        if (!lh) return right->intersects(ray, ray_t);
        if (!rh) return left->intersects(ray, ray_t);
        auto l = left->intersects(ray, ray_t);
        auto r = right->intersects(ray, ray_t);
        if (!l && !r) return std::nullopt;
        if (!l) return r;
        if (!r) return l;
        if (l->dist < r->dist) return l;  // tie -> right (Q2)
        return r;
    }
    AABB bounding_box() const override { return bbox; }
    std::optional<Vec3> sample(Vec3, double, Rng&) const override { return std::nullopt; }
    double pdf(Vec3, Vec3, double) const override { return 0.0; }
};

struct BVH {
    static constexpr size_t MAX_HITTABLES_PER_LEAF = 4;  // bvh.rs:22
    struct Item { HitPtr h; uint32_t id; AABB box; Vec3 centroid; };

    static std::unique_ptr<BVHNode> build(const std::vector<HitPtr>& hs) {
        std::vector<Item> items;
        for (uint32_t i = 0; i < hs.size(); i++) {
            AABB b = hs[i]->bounding_box();
            items.push_back({hs[i], i, b, b.centroid()});
        }
        return build_recursive(items);
    }
    static std::unique_ptr<BVHNode> make_leaf(const std::vector<Item>& items) {
        auto n = std::make_unique<BVHNode>();
        AABB b;
        for (auto& it : items) b = b.unite(it.box);  // fold(AABB::default(), union): bvh.rs:30-32
        n->bbox = b;
        for (auto& it : items) { n->hittables.push_back(it.h); n->item_ids.push_back(it.id); }
        return n;
    }
    static std::unique_ptr<BVHNode> build_recursive(const std::vector<Item>& items) {  // bvh.rs:28-52
        if (items.size() <= MAX_HITTABLES_PER_LEAF) return make_leaf(items);
        std::vector<Item> l, r;
        find_best_split(items, l, r);
        if (l.empty() || r.empty()) return make_leaf(items);
        auto ln = build_recursive(l);
        auto rn = build_recursive(r);
        auto n = std::make_unique<BVHNode>();
        n->is_leaf = false;
        n->bbox = ln->bbox.unite(rn->bbox);
        n->left = std::move(ln); n->right = std::move(rn);
        return n;
    }
    static double evaluate_sah(int axis, double split_pos, const AABB& parent, const std::vector<Item>& items) {
        AABB lb, rb; size_t lc = 0, rc = 0;  // bvh.rs:86-120
        for (auto& it : items) {
            if (it.centroid[axis] < split_pos) { lb = lb.unite(it.box); lc++; }
            else { rb = rb.unite(it.box); rc++; }
        }
        if (lc == 0 || rc == 0) return INF;
        double cost = lb.surface_area() * (double)lc + rb.surface_area() * (double)rc;
        double parent_cost = parent.surface_area() * (double)items.size();
        if (cost > 0.0 && cost < parent_cost) return cost;
        return INF;
    }
    static void find_best_split(const std::vector<Item>& items, std::vector<Item>& l, std::vector<Item>& r) {
        AABB parent;  // bvh.rs:54-84
        for (auto& it : items) parent = parent.unite(it.box);
        double best_cost = INF; int best_axis = 0; double best_split = 0.0;
        for (int axis = 0; axis < 3; axis++) {
            std::vector<double> pos;
            for (auto& it : items) pos.push_back(it.centroid[axis]);
            std::stable_sort(pos.begin(), pos.end());
            // evaluate_sah is a pure function of (axis, split_pos): evaluate in parallel, then take the
            // reference's sequential strict-'<' argmin (bvh.rs:68-76) over the same ordered costs.
            std::vector<double> costs(pos.size());
            const long n = (long)pos.size();
#pragma omp parallel for schedule(dynamic, 16) if (n > 256)
            for (long k = 0; k < n; k++) costs[k] = evaluate_sah(axis, pos[k], parent, items);
            for (long k = 0; k < n; k++) {
                if (costs[k] < best_cost) { best_cost = costs[k]; best_axis = axis; best_split = pos[k]; }
            }
        }
        for (auto& it : items) {  // partition keeps list order
            if (it.centroid[best_axis] < best_split) l.push_back(it); else r.push_back(it);
        }
    }
};

// ---------------------------------------------------------------- HittableList (src/hittable/list.rs)
struct HittableList : Hittable {
    std::vector<HitPtr> objects;
    AABB bbox;
    std::unique_ptr<BVHNode> bvh;
    void add(HitPtr o) { bbox = bbox.unite(o->bounding_box()); objects.push_back(std::move(o)); }  // list.rs:24-27
    void build_bvh() { if (!objects.empty()) bvh = BVH::build(objects); }                          // list.rs:29-33
    bool is_empty() const { return objects.empty(); }
    std::optional<HitInfo> intersects(const Ray& ray, Interval ray_t) const override {  // list.rs:49-68
        if (bvh) return bvh->intersects(ray, ray_t);
        double closest = ray_t.max;
        std::optional<HitInfo> hit;
        for (auto& o : objects) {
            if (auto info = o->intersects(ray, Interval{ray_t.min, closest})) { closest = info->dist; hit = info; }
        }
        return hit;
    }
    AABB bounding_box() const override { return bbox; }
    std::optional<Vec3> sample(Vec3 origin, double time, Rng& rng) const override {  // list.rs:78-84
        if (is_empty()) return std::nullopt;
        // gen_range(0..len): one draw; RNG contract: index = min(floor(U*len), len-1)
        size_t n = objects.size();
        size_t i = (size_t)(rng.next() * (double)n);
        if (i >= n) i = n - 1;
        return objects[i]->sample(origin, time, rng);
    }
    double pdf(Vec3 origin, Vec3 direction, double time) const override {  // list.rs:86-96
        if (objects.empty()) return 0.0;
        double s = 0.0;
        for (auto& o : objects) s += o->pdf(origin, direction, time);  // Iterator::sum from 0.0
        return s / (double)objects.size();
    }
};

// ---------------------------------------------------------------- Sphere (src/hittable/sphere.rs)
struct Sphere : Hittable {
    double radius; Vec3 position1, position2; const BxDF* material; AABB bbox; uint32_t id;
    Sphere(double r, Vec3 p1, Vec3 p2, bool moving, const BxDF* m, uint32_t id_) : position1(p1), position2(p2), material(m), id(id_) {
        Vec3 rvec(r, r, r);  // bbox uses the un-clamped radius (sphere.rs:23-24,35-38)
        if (!moving) bbox = AABB::make(p1 - rvec, p1 + rvec);  // new_still
        else bbox = AABB::make(p1 - rvec, p1 + rvec).unite(AABB::make(p2 - rvec, p2 + rvec));  // new_moving
        radius = fmax_(r, 0.0);
    }
    Vec3 get_position(double t) const { return position1 + (position2 - position1) * t; }  // sphere.rs:58-60
    static void get_uv(const Vec3& p, double& u, double& v) {                               // sphere.rs:52-56
        double theta = std::acos(-p.y);
        double phi = std::atan2(-p.z, p.x) + PI;
        u = phi / (2.0 * PI); v = theta / PI;
    }
    std::optional<HitInfo> intersects(const Ray& ray, Interval ray_t) const override {  // sphere.rs:64-100
        g_cnt.spheres++;
        Vec3 c = get_position(ray.time);
        Vec3 l = c - ray.origin;
        double s = dot(l, ray.direction);
        double l2 = length_squared(l);
        double r2 = radius * radius;
        if (s < 0.0 && l2 > r2) return std::nullopt;
        double d2 = l2 - s * s;
        if (d2 > r2) return std::nullopt;
        double q = std::sqrt(r2 - d2);
        double t = l2 > r2 ? s - q : s + q;
        if (t <= ray_t.min || t >= ray_t.max) return std::nullopt;  // exclusive
        Vec3 point = ray.at(t);
        Vec3 normal = normalize(point - c);
        double u, v; get_uv(normal, u, v);
        HitInfo h = make_hit_info(ray, point, normal, t, material, u, v);
        h.prim_kind = 0; h.prim_index = id;
        return h;
    }
    AABB bounding_box() const override { return bbox; }
    std::optional<Vec3> sample(Vec3 origin, double time, Rng& rng) const override {  // sphere.rs:110-121
        double u = rng.next(), v = rng.next();
        double theta = 2.0 * PI * u;
        double phi = std::acos(2.0 * v - 1.0);
        double x = std::sin(phi) * std::cos(theta), y = std::sin(phi) * std::sin(theta), z = std::cos(phi);
        Vec3 point = get_position(time) + Vec3(x, y, z) * radius;
        return normalize(point - origin);
    }
    double pdf(Vec3 origin, Vec3 direction, double time) const override {  // sphere.rs:123-135
        if (intersects(Ray::make(origin, direction, time), Interval{0.0, INF})) {
            double r2 = radius * radius;
            double solid_angle = 2.0 * PI * std::sqrt(1.0 - r2 / length_squared(get_position(time) - origin));
            return 1.0 / solid_angle;
        }
        return 0.0;
    }
};

// ---------------------------------------------------------------- Quad (src/hittable/quad.rs)
struct Quad : Hittable {
    Vec3 q, u, v, w, normal; double d; AABB bbox; const BxDF* material; uint32_t id;
    Quad(Vec3 q_, Vec3 u_, Vec3 v_, const BxDF* m, uint32_t id_) : q(q_), u(u_), v(v_), material(m), id(id_) {  // quad.rs:17-36
        AABB b1 = AABB::make(q, q + u + v);
        AABB b2 = AABB::make(q + u, q + v);
        bbox = b1.unite(b2);
        Vec3 n = cross(u, v);
        normal = normalize(n);
        d = dot(normal, q);
        w = n / length_squared(n);
    }
    std::optional<HitInfo> intersects(const Ray& ray, Interval ray_t) const override {  // quad.rs:40-70
        g_cnt.quads++;
        double nd = dot(normal, ray.direction);
        if (std::fabs(nd) < 1e-8) return std::nullopt;
        double t = (d - dot(normal, ray.origin)) / nd;
        if (!ray_t.contains(t)) return std::nullopt;  // inclusive
        Vec3 p = ray.at(t) - q;
        double alpha = dot(w, cross(p, v));
        double beta = dot(w, cross(u, p));
        if (!(0.0 <= alpha && alpha <= 1.0) || !(0.0 <= beta && beta <= 1.0)) return std::nullopt;
        HitInfo h = make_hit_info(ray, ray.at(t), normal, t, material, alpha, beta);
        h.prim_kind = 1; h.prim_index = id;
        return h;
    }
    AABB bounding_box() const override { return bbox; }
    std::optional<Vec3> sample(Vec3 origin, double, Rng& rng) const override {  // quad.rs:80-86
        double a = rng.next(), b = rng.next();
        Vec3 point = q + u * a + v * b;
        return normalize(point - origin);
    }
    double pdf(Vec3 origin, Vec3 direction, double time) const override {  // quad.rs:88-98
        Ray ray = Ray::make(origin, direction, time);
        if (auto hit = intersects(ray, Interval{0.0, INF})) {
            double area = length(cross(u, v));
            double dist = hit->dist;
            double cos_theta = std::fabs(dot(ray.direction, hit->shading_normal));  // Q9
            return (dist * dist) / (cos_theta * area);
        }
        return 0.0;
    }
};

// ---------------------------------------------------------------- Cuboid (src/hittable/cuboid.rs)
struct Cuboid : Hittable {
    HittableList sides;  // BVH-less (Q30)
    // quad ids are first_quad..first_quad+6 in cuboid.rs:18-53 order
    Cuboid(Vec3 a, Vec3 b, const BxDF* mat, uint32_t first_quad) {
        Vec3 mn = vmin(a, b), mx = vmax(a, b);
        Vec3 dx(mx.x - mn.x, 0, 0), dy(0, mx.y - mn.y, 0), dz(0, 0, mx.z - mn.z);
        auto add = [&](Vec3 q, Vec3 u, Vec3 v, uint32_t k) { sides.add(std::make_shared<Quad>(q, u, v, mat, first_quad + k)); };
        add(Vec3(mn.x, mn.y, mx.z), dx, dy, 0);    // front
        add(Vec3(mx.x, mn.y, mx.z), -dz, dy, 1);   // right
        add(Vec3(mx.x, mn.y, mn.z), -dx, dy, 2);   // back
        add(Vec3(mn.x, mn.y, mn.z), dz, dy, 3);    // left
        add(Vec3(mn.x, mx.y, mx.z), dx, -dz, 4);   // top
        add(Vec3(mn.x, mn.y, mn.z), dx, dz, 5);    // bottom
    }
    std::optional<HitInfo> intersects(const Ray& r, Interval t) const override { return sides.intersects(r, t); }
    AABB bounding_box() const override { return sides.bounding_box(); }
    std::optional<Vec3> sample(Vec3 o, double t, Rng& rng) const override { return sides.sample(o, t, rng); }
    double pdf(Vec3 o, Vec3 d, double t) const override { return sides.pdf(o, d, t); }
};

// ---------------------------------------------------------------- Triangle / TriangleMesh (src/hittable/mesh.rs)
struct Triangle : Hittable {
    Vec3 v[3]; bool has_n = false, has_uv = false; Vec3 n[3]; double uv[3][2];
    const BxDF* material; AABB bbox; uint32_t id;
    Triangle(Vec3 v0, Vec3 v1, Vec3 v2, const BxDF* m, uint32_t id_) : material(m), id(id_) {  // mesh.rs:22-40
        v[0] = v0; v[1] = v1; v[2] = v2;
        bbox = AABB::make(vmin(vmin(v0, v1), v2), vmax(vmax(v0, v1), v2));
    }
    double area() const { return 0.5 * length(cross(v[1] - v[0], v[2] - v[0])); }  // mesh.rs:42-46
    std::optional<HitInfo> intersects(const Ray& ray, Interval ray_t) const override {  // mesh.rs:50-112
        g_cnt.triangles++;
        Vec3 e1 = v[1] - v[0], e2 = v[2] - v[0];
        Vec3 h = cross(ray.direction, e2);
        double a = dot(e1, h);
        if (std::fabs(a) < 1e-8) return std::nullopt;
        double f = 1.0 / a;
        Vec3 s = ray.origin - v[0];
        double uu = f * dot(s, h);
        if (!(0.0 <= uu && uu <= 1.0)) return std::nullopt;
        Vec3 q = cross(s, e1);
        double vv = f * dot(ray.direction, q);
        if (vv < 0.0 || uu + vv > 1.0) return std::nullopt;
        double t = f * dot(e2, q);
        if (!ray_t.contains(t)) return std::nullopt;
        double w = 1.0 - uu - vv;
        Vec3 normal = has_n ? normalize(n[0] * w + n[1] * uu + n[2] * vv) : normalize(cross(e1, e2));
        double ou = uu, ov = vv;
        if (has_uv) {
            ou = uv[0][0] * w + uv[1][0] * uu + uv[2][0] * vv;
            ov = uv[0][1] * w + uv[1][1] * uu + uv[2][1] * vv;
        }
        HitInfo hi = make_hit_info(ray, ray.at(t), normal, t, material, ou, ov);
        hi.prim_kind = 2; hi.prim_index = id;
        return hi;
    }
    AABB bounding_box() const override { return bbox; }
    std::optional<Vec3> sample(Vec3 origin, double, Rng& rng) const override {  // mesh.rs:122-129
        double a = rng.next(), b = rng.next();
        double w = 1.0 - a - b;
        Vec3 point = v[0] * w + v[1] * a + v[2] * b;
        return normalize(point - origin);
    }
    double pdf(Vec3 origin, Vec3 direction, double time) const override {  // mesh.rs:131-141
        Ray ray = Ray::make(origin, direction, time);
        if (auto hit = intersects(ray, Interval{0.0, INF})) {
            double dist = hit->dist;
            double cos_theta = std::fabs(dot(direction, hit->shading_normal));  // un-normalised `direction`
            return dist * dist / (cos_theta * area());
        }
        return 0.0;
    }
};
struct TriangleMesh : Hittable {  // mesh.rs:144-220
    HittableList triangles;
    std::optional<HitInfo> intersects(const Ray& r, Interval t) const override { return triangles.intersects(r, t); }
    AABB bounding_box() const override { return triangles.bounding_box(); }
    std::optional<Vec3> sample(Vec3 o, double t, Rng& rng) const override { return triangles.sample(o, t, rng); }
    double pdf(Vec3 o, Vec3 d, double t) const override { return triangles.pdf(o, d, t); }
};

// ---------------------------------------------------------------- Instance (src/hittable/instance.rs)
struct Instance : Hittable {
    HitPtr object; AABB bbox; Quat rotation; Mat4 transform; uint32_t id;
    Mat4 inverse, normal_mat;  // the reference recomputes these per call (Q5); values are identical
    Instance(HitPtr obj, Vec3 axis, double angle, Vec3 translation, uint32_t id_) : object(std::move(obj)), id(id_) {
        rotation = quat_from_axis_angle(axis, angle);                       // instance.rs:21
        transform = mat4_from_rotation_translation(rotation, translation);  // instance.rs:22
        bbox = object->bounding_box().transform(transform);                 // instance.rs:23
        inverse = mat4_inverse(transform);
        normal_mat = mat4_transpose(mat4_inverse(mat4_from_quat(rotation)));  // instance.rs:45
    }
    std::optional<HitInfo> intersects(const Ray& ray, Interval ray_t) const override {  // instance.rs:34-54
        g_cnt.instances++;
        Vec3 lo = transform_point3(inverse, ray.origin);
        Vec3 ld = transform_vector3(inverse, ray.direction);
        Ray local = Ray::make(lo, ld, ray.time);
        auto info = object->intersects(local, ray_t);
        if (!info) return std::nullopt;
        HitInfo h = *info;  // shading_normal, u, v, front_face stay object-space (Q4)
        h.point = transform_point3(transform, info->point);
        h.geometric_normal = normalize(transform_vector3(normal_mat, info->geometric_normal));
        h.instance = id;
        return h;
    }
    AABB bounding_box() const override { return bbox; }
    std::optional<Vec3> sample(Vec3 origin, double time, Rng& rng) const override {  // instance.rs:64-69
        Vec3 lo = transform_point3(inverse, origin);
        auto ld = object->sample(lo, time, rng);
        if (!ld) return std::nullopt;
        return transform_vector3(transform, *ld);
    }
    double pdf(Vec3 origin, Vec3 direction, double time) const override {  // instance.rs:71-75
        Vec3 lo = transform_point3(inverse, origin);
        Vec3 ld = transform_vector3(inverse, direction);
        return object->pdf(lo, ld, time);
    }
};

// ---------------------------------------------------------------- HomogeneousVolume (src/volume.rs:15-41 is a stub)
// NOT reference behaviour: the reference never finished this type (`intersects` is todo!()).  The semantics are the
// ones include/pt_b200.h documents for pt_volume; this is the CPU statement of them the device is checked against.
struct HomogeneousVolume : Hittable {
    HitPtr boundary; double negative_inv_density; const BxDF* phase_function; uint32_t id;
    HomogeneousVolume(HitPtr b, double density, const BxDF* phase, uint32_t id_)
        : boundary(std::move(b)), negative_inv_density(-1.0 / density), phase_function(phase), id(id_) {}
    std::optional<HitInfo> intersects(const Ray& ray, Interval ray_t) const override {
        auto h1 = boundary->intersects(ray, Interval{ray_t.min, INF});
        if (!h1) return std::nullopt;
        double t_in, t_out;
        if (h1->front_face) {
            t_in = h1->dist;
            // the exit is found from just inside: sphere.rs:80 only ever returns the near root to an outside origin
            double step = t_in + 1e-4;
            Ray inner{ray.at(step), ray.direction, ray.time};
            auto h2 = boundary->intersects(inner, Interval{0.0, INF});
            if (!h2) return std::nullopt;
            t_out = step + h2->dist;
        } else { t_in = ray_t.min; t_out = h1->dist; }
        double s = negative_inv_density * std::log(keyed_uniform(id));
        if (s > t_out - t_in) return std::nullopt;
        double t = t_in + s;
        if (!ray_t.contains(t)) return std::nullopt;
        HitInfo h = make_hit_info(ray, ray.at(t), Vec3(1, 0, 0), t, phase_function, 0.0, 0.0);
        h.prim_kind = 6; h.prim_index = id;
        return h;
    }
    AABB bounding_box() const override { return boundary->bounding_box(); }
    std::optional<Vec3> sample(Vec3, double, Rng&) const override { return std::nullopt; }
    double pdf(Vec3, Vec3, double) const override { return 0.0; }
};

// ---------------------------------------------------------------- World (src/hittable/world.rs)
struct World {
    HittableList objects, lights;
    void build_bvh() { objects.build_bvh(); lights.build_bvh(); }  // world.rs:26-29
    // world.rs:47-62
    std::optional<std::pair<HitInfo, bool>> intersect_all(const Ray& ray, Interval ray_t) const {
        g_cnt.segments++;
        auto lh = lights.intersects(ray, ray_t);
        auto oh = objects.intersects(ray, ray_t);
        if (!lh && !oh) return std::nullopt;
        if (!lh) return std::make_pair(*oh, false);
        if (!oh) return std::make_pair(*lh, true);
        if (lh->dist < oh->dist) return std::make_pair(*lh, true);
        return std::make_pair(*oh, false);  // tie -> object (Q31)
    }
    // world.rs:31-36 (unused by the reference's integrator)
    bool occluded(const Ray& ray, Interval ray_t) const { return objects.intersects(ray, ray_t).has_value(); }
};

}  // namespace orc
