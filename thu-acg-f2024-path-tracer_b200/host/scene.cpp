// Host scene objects, SAH BVH build and flattening (see pt_host.hpp).  Reference citations inline.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "pt_host.hpp"
#include "device_api.hpp"

namespace pt {

static const double kInf = std::numeric_limits<double>::infinity();

// ---- glam DMat4 helpers, column-major m[col*4+row] ------------------------------------------------
static void mat_from_rotation_translation(const double q[4], Vec3 t, double m[16]) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double x2 = x + x, y2 = y + y, z2 = z + z;
    double xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2;
    double wx = w * x2, wy = w * y2, wz = w * z2;
    double r[16] = {1.0 - (yy + zz), xy + wz, xz - wy, 0.0, xy - wz, 1.0 - (xx + zz), yz + wx, 0.0,
                    xz + wy, yz - wx, 1.0 - (xx + yy), 0.0, t.x, t.y, t.z, 1.0};
    memcpy(m, r, sizeof(r));
}
static void mat_inverse(const double m[16], double out[16]) {  // cofactor expansion, as glam's scalar DMat4::inverse
    auto M = [&](int c, int r) { return m[c * 4 + r]; };
    double c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3), c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    double c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3), c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    double c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2), c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    double c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3), c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    double c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2), c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    double c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1), c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    double f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    double f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    double v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
    double v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
    const double sa[4] = {1.0, -1.0, 1.0, -1.0}, sb[4] = {-1.0, 1.0, -1.0, 1.0};
    for (int i = 0; i < 4; i++) {
        out[0 * 4 + i] = ((v1[i] * f0[i] - v2[i] * f1[i]) + v3[i] * f2[i]) * sa[i];
        out[1 * 4 + i] = ((v0[i] * f0[i] - v2[i] * f3[i]) + v3[i] * f4[i]) * sb[i];
        out[2 * 4 + i] = ((v0[i] * f1[i] - v1[i] * f3[i]) + v3[i] * f5[i]) * sa[i];
        out[3 * 4 + i] = ((v0[i] * f2[i] - v1[i] * f4[i]) + v2[i] * f5[i]) * sb[i];
    }
    double det = m[0] * out[0] + m[1] * out[4] + m[2] * out[8] + m[3] * out[12];
    double rcp = 1.0 / det;
    for (int i = 0; i < 16; i++) out[i] = out[i] * rcp;
}
static Vec3 mat_point(const double m[16], Vec3 p) {  // DMat4::transform_point3
    double r[3];
    for (int i = 0; i < 3; i++) { double s = m[i] * p.x; s = m[4 + i] * p.y + s; s = m[8 + i] * p.z + s; r[i] = m[12 + i] + s; }
    return {r[0], r[1], r[2]};
}
Box Box::transformed(const double m[16]) const {  // aabb.rs:54-78
    Vec3 cs[8] = {lo, {lo.x, lo.y, hi.z}, {lo.x, hi.y, lo.z}, {lo.x, hi.y, hi.z}, {hi.x, lo.y, lo.z}, {hi.x, lo.y, hi.z}, {hi.x, hi.y, lo.z}, hi};
    Vec3 mn(kInf, kInf, kInf), mx(-kInf, -kInf, -kInf);
    for (auto& c : cs) { Vec3 t = mat_point(m, c); mn = vmin(mn, t); mx = vmax(mx, t); }
    return of(mn, mx);
}

// ---- textures / materials ------------------------------------------------------------------------
TexPtr SolidTexture::make(Vec3 v) { auto t = std::make_shared<Texture>(); t->kind = PT_TEX_SOLID; t->value = v; return t; }
TexPtr SolidTexture::scalar(double v) { return make(Vec3(v, 0, 0)); }
TexPtr CheckerTexture::make(double scale, TexPtr a, TexPtr b) {
    auto t = std::make_shared<Texture>(); t->kind = PT_TEX_CHECKER; t->inv_scale = 1.0 / scale;  // texture.rs:36
    t->tex1 = std::move(a); t->tex2 = std::move(b); return t;
}
ImagePtr ImageTexture::from_rgb8(const uint8_t* rgb, uint32_t w, uint32_t h) {
    auto im = std::make_shared<Image>(); im->width = w; im->height = h; im->rgb.assign(rgb, rgb + (size_t)3 * w * h); return im;
}
TexPtr ImageTexture::make(ImagePtr img) { auto t = std::make_shared<Texture>(); t->kind = PT_TEX_IMAGE; t->image = std::move(img); return t; }

static std::shared_ptr<Material> new_mat(uint32_t kind) { auto m = std::make_shared<Material>(); m->kind = kind; return m; }
MatPtr DiffuseBRDF::make(TexPtr c) { auto m = new_mat(PT_MAT_DIFFUSE); m->base_color = std::move(c); return m; }
MatPtr DiffuseBRDF::from_rgb(Vec3 c) { return make(SolidTexture::make(c)); }
MatPtr DiffuseBRDF::from_textures(TexPtr c, ImagePtr n) { auto m = new_mat(PT_MAT_DIFFUSE); m->base_color = std::move(c); m->normal_map = std::move(n); return m; }
MatPtr MetalBRDF::make(TexPtr c, TexPtr r) { auto m = new_mat(PT_MAT_METAL); m->base_color = std::move(c); m->roughness = std::move(r); return m; }
MatPtr MetalBRDF::from_rgb(Vec3 c, double r) { return make(SolidTexture::make(c), SolidTexture::scalar(r)); }
MatPtr GlassBSDF::make(TexPtr c, TexPtr r, double, double ior) {
    auto m = new_mat(PT_MAT_GLASS); m->base_color = std::move(c); m->roughness = std::move(r); m->p[PT_P_IOR] = ior; return m;
}
MatPtr GlassBSDF::basic(double ior) { return make(SolidTexture::make(Vec3(1, 1, 1)), SolidTexture::scalar(0.001), 0.0, ior); }  // glass.rs:42-49
MatPtr PrincipledBSDF::make(TexPtr c, double metallic, double roughness, double subsurface, double specular, double specular_tint,
                            double ior, double spec_trans, double sheen, double sheen_tint, double clearcoat, double clearcoat_gloss) {
    auto m = new_mat(PT_MAT_PRINCIPLED); m->base_color = std::move(c);
    m->p[PT_P_METALLIC] = metallic; m->p[PT_P_ROUGHNESS] = roughness; m->p[PT_P_SUBSURFACE] = subsurface; m->p[PT_P_SPECULAR] = specular;
    m->p[PT_P_SPECULAR_TINT] = specular_tint; m->p[PT_P_IOR] = ior; m->p[PT_P_SPEC_TRANS] = spec_trans; m->p[PT_P_SHEEN] = sheen;
    m->p[PT_P_SHEEN_TINT] = sheen_tint; m->p[PT_P_CLEARCOAT] = clearcoat; m->p[PT_P_CLEARCOAT_GLOSS] = clearcoat_gloss;
    return m;
}
MatPtr DiffuseLight::make(TexPtr e) { auto m = new_mat(PT_MAT_LIGHT); m->base_color = std::move(e); return m; }
MatPtr DiffuseLight::from_rgb(Vec3 rgb) { return make(SolidTexture::make(rgb)); }
MatPtr SheenBRDF::make(Vec3 c, double tint) {
    auto m = new_mat(PT_MAT_SHEEN); m->p[PT_P_COLOR_R] = c.x; m->p[PT_P_COLOR_G] = c.y; m->p[PT_P_COLOR_B] = c.z; m->p[PT_P_SHEEN_TINT] = tint; return m;
}
MatPtr ClearcoatBRDF::make(double gloss) { auto m = new_mat(PT_MAT_CLEARCOAT); m->p[PT_P_ALPHA_G] = (1.0 - gloss) * 0.1 + gloss * 0.001; return m; }  // clearcoat.rs:16-18
MatPtr IsotropicMaterial::from_texture(TexPtr albedo) { auto m = new_mat(PT_MAT_ISOTROPIC); m->base_color = std::move(albedo); return m; }
MatPtr IsotropicMaterial::from_albedo(Vec3 albedo) { return from_texture(SolidTexture::make(albedo)); }
MatPtr MixBxDf::make(double t, MatPtr a, MatPtr b) {
    auto m = new_mat(PT_MAT_MIX); m->p[PT_P_MIX_T] = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t); m->mix_a = std::move(a); m->mix_b = std::move(b); return m;  // mix.rs:17
}

// ---- hittables -----------------------------------------------------------------------------------
void HittableList::add(HitPtr h) { bbox = bbox.merged(h->bbox); objects.push_back(std::move(h)); }  // list.rs:24-27
void HittableList::build_bvh() {  // list.rs:29-33
    if (objects.empty()) return;
    std::vector<Box> boxes; for (auto& o : objects) boxes.push_back(o->bbox);
    bvh = pt::build_bvh(boxes);
}
HitPtr Sphere::new_still(double r, Vec3 p, MatPtr m) {  // sphere.rs:22-32
    auto h = std::make_shared<Hittable>(); h->kind = PT_PRIM_SPHERE; Vec3 rv(r, r, r);
    h->bbox = Box::of(p - rv, p + rv); h->radius = r; h->p1 = p; h->p2 = p; h->material = std::move(m); return h;
}
HitPtr Sphere::new_moving(double r, Vec3 p1, Vec3 p2, MatPtr m) {  // sphere.rs:34-46
    auto h = std::make_shared<Hittable>(); h->kind = PT_PRIM_SPHERE; Vec3 rv(r, r, r);
    h->bbox = Box::of(p1 - rv, p1 + rv).merged(Box::of(p2 - rv, p2 + rv));
    h->radius = r; h->p1 = p1; h->p2 = p2; h->moving = true; h->material = std::move(m); return h;
}
HitPtr Quad::make(Vec3 q, Vec3 u, Vec3 v, MatPtr m) {  // quad.rs:17-36
    auto h = std::make_shared<Hittable>(); h->kind = PT_PRIM_QUAD;
    h->bbox = Box::of(q, q + u + v).merged(Box::of(q + u, q + v));
    Vec3 n = cross(u, v);
    h->q = q; h->u = u; h->v = v; h->normal = normalize(n); h->d = dot(h->normal, q); h->w = n / dot(n, n);
    h->material = std::move(m); return h;
}
HitPtr Cuboid::make(Vec3 a, Vec3 b, MatPtr m) {  // cuboid.rs:11-58
    auto h = std::make_shared<Hittable>(); h->kind = PT_OBJ_CUBOID; h->a = a; h->b = b; h->material = m;
    Vec3 mn = vmin(a, b), mx = vmax(a, b);
    Vec3 dx(mx.x - mn.x, 0, 0), dy(0, mx.y - mn.y, 0), dz(0, 0, mx.z - mn.z);
    h->sides.add(Quad::make(Vec3(mn.x, mn.y, mx.z), dx, dy, m));   // front
    h->sides.add(Quad::make(Vec3(mx.x, mn.y, mx.z), -dz, dy, m));  // right
    h->sides.add(Quad::make(Vec3(mx.x, mn.y, mn.z), -dx, dy, m));  // back
    h->sides.add(Quad::make(Vec3(mn.x, mn.y, mn.z), dz, dy, m));   // left
    h->sides.add(Quad::make(Vec3(mn.x, mx.y, mx.z), dx, -dz, m));  // top
    h->sides.add(Quad::make(Vec3(mn.x, mn.y, mn.z), dx, dz, m));   // bottom
    h->bbox = h->sides.bbox;
    return h;
}
HitPtr Instance::make(HitPtr object, Vec3 axis, double angle, Vec3 translation) {  // instance.rs:20-31
    if (object->kind == PT_OBJ_INSTANCE) throw std::runtime_error("nested Instance is outside the device subset");
    auto h = std::make_shared<Hittable>(); h->kind = PT_OBJ_INSTANCE; h->child = object; h->axis = axis; h->angle = angle; h->translation = translation;
    double s = std::sin(angle * 0.5), c = std::cos(angle * 0.5);  // DQuat::from_axis_angle
    Vec3 v = axis * s; double q[4] = {v.x, v.y, v.z, c};
    mat_from_rotation_translation(q, translation, h->transform);
    mat_inverse(h->transform, h->inverse);
    double rot[16], rinv[16];
    mat_from_rotation_translation(q, Vec3(0, 0, 0), rot);  // Mat4::from_quat
    mat_inverse(rot, rinv);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) h->normal_matrix[i * 4 + j] = rinv[j * 4 + i];  // transpose
    h->bbox = object->bbox.transformed(h->transform);
    h->material = object->material;
    return h;
}
HitPtr HomogeneousVolume::from_texture(HitPtr boundary, double density, TexPtr texture) {  // volume.rs:22-34 (stub)
    if (boundary->kind != PT_PRIM_SPHERE && boundary->kind != PT_OBJ_CUBOID) throw std::runtime_error("volume boundary must be a sphere or a cuboid");
    if (!(density > 0.0)) throw std::runtime_error("volume density must be positive");
    auto h = std::make_shared<Hittable>(); h->kind = PT_OBJ_VOLUME; h->child = std::move(boundary); h->density = density;
    h->material = IsotropicMaterial::from_texture(std::move(texture));
    h->bbox = h->child->bbox;  // volume.rs:48-50: the boundary's box, as is
    return h;
}
HitPtr HomogeneousVolume::from_albedo(HitPtr boundary, double density, Vec3 albedo) { return from_texture(std::move(boundary), density, SolidTexture::make(albedo)); }
HitPtr TriangleMesh::from_obj(double scale, const ObjMesh& mesh, MatPtr m) {  // mesh.rs:149-197
    auto h = std::make_shared<Hittable>(); h->kind = PT_OBJ_MESH; h->material = std::move(m);
    size_t nv = mesh.positions.size() / 3, nn = mesh.normals.size() / 3, nt = mesh.texcoords.size() / 2;
    std::vector<Vec3> verts(nv);
    for (size_t i = 0; i < nv; i++)
        verts[i] = Vec3((double)mesh.positions[3 * i], (double)mesh.positions[3 * i + 1], (double)mesh.positions[3 * i + 2]) * scale;
    HittableList tris;  // only bbox folding + BVH build are needed host-side
    for (size_t f = 0; f + 2 < mesh.indices.size(); f += 3) {
        uint32_t i0 = mesh.indices[f], i1 = mesh.indices[f + 1], i2 = mesh.indices[f + 2];
        if (i0 >= nv || i1 >= nv || i2 >= nv) throw std::runtime_error("mesh index out of range");
        Vec3 v0 = verts[i0], v1 = verts[i1], v2 = verts[i2];
        h->triangles.push_back(pt_triangle{v0.c(), v1.c(), v2.c()});
        Box b = Box::of(vmin(vmin(v0, v1), v2), vmax(vmax(v0, v1), v2));  // mesh.rs:30-32
        h->tri_boxes.push_back(b);
        h->sides.bbox = h->sides.bbox.merged(b);  // HittableList::add folding
        if (nn) {  // indexed by POSITION index (mesh.rs:175-179)
            if (i0 >= nn || i1 >= nn || i2 >= nn) throw std::runtime_error("mesh normal index out of range");
            for (uint32_t i : {i0, i1, i2}) h->tri_normals.push_back(pt_vec3{(double)mesh.normals[3 * i], (double)mesh.normals[3 * i + 1], (double)mesh.normals[3 * i + 2]});
        }
        if (nt) {  // Q7: uvs indexed by position index (mesh.rs:180-184)
            if (i0 >= nt || i1 >= nt || i2 >= nt) throw std::runtime_error("mesh uv index out of range");
            for (uint32_t i : {i0, i1, i2}) { h->tri_uvs.push_back((double)mesh.texcoords[2 * i]); h->tri_uvs.push_back((double)mesh.texcoords[2 * i + 1]); }
        }
    }
    h->bbox = h->sides.bbox;
    if (!h->tri_boxes.empty()) h->mesh_bvh = pt::build_bvh(h->tri_boxes);  // triangles.build_bvh(), mesh.rs:195
    return h;
}

// ---- SAH build, bvh.rs:24-120 --------------------------------------------------------------------
static thread_local pt_ctx* g_build_ctx = nullptr;  // set_build_context: sweep large nodes on the device (same tree, faster)
void set_build_context(pt_ctx* ctx) { g_build_ctx = ctx; }
namespace {
struct Builder {
    const std::vector<Box>& boxes; std::vector<Vec3> cent; BvhTree& tree; pt_ctx* ctx = nullptr;
    Box fold(const std::vector<uint32_t>& items) const { Box b; for (uint32_t i : items) b = b.merged(boxes[i]); return b; }
    double sah(int axis, double split, const Box& parent, const std::vector<uint32_t>& items) const {  // bvh.rs:86-120
        Box lb, rb; size_t lc = 0, rc = 0;
        for (uint32_t i : items) {
            if (cent[i][axis] < split) { lb = lb.merged(boxes[i]); lc++; } else { rb = rb.merged(boxes[i]); rc++; }
        }
        if (lc == 0 || rc == 0) return kInf;
        double cost = lb.half_area() * (double)lc + rb.half_area() * (double)rc;
        double parent_cost = parent.half_area() * (double)items.size();
        return (cost > 0.0 && cost < parent_cost) ? cost : kInf;
    }
    int32_t leaf(const std::vector<uint32_t>& items) {
        BvhTree::Node n; n.box = fold(items); n.items = items; tree.nodes.push_back(std::move(n)); return (int32_t)tree.nodes.size() - 1;
    }
    int32_t build(const std::vector<uint32_t>& items) {  // bvh.rs:28-52
        if (items.size() <= 4) return leaf(items);
        Box parent = fold(items);
        double best_cost = kInf, best_split = 0.0; int best_axis = 0;
        std::vector<double> pos(items.size()), cost(items.size());
        if (ctx && items.size() >= 256) {  // the O(n^2) sweep on the device (pt_sah_sweep): bit-identical costs, same choice
            const size_t n = items.size();
            std::vector<Box> ib(n); for (size_t k = 0; k < n; k++) ib[k] = boxes[items[k]];
            std::vector<double> dc(3 * n);
            if (device_api().sah_sweep(ctx, (uint32_t)n, reinterpret_cast<const double*>(ib.data()), reinterpret_cast<const double*>(&parent), dc.data()) != PT_OK)
                throw std::runtime_error(std::string("pt_sah_sweep: ") + device_api().last_error());
            std::vector<std::pair<double, double>> pc(n);
            for (int axis = 0; axis < 3; axis++) {
                for (size_t k = 0; k < n; k++) pc[k] = {cent[items[k]][axis], dc[axis * n + k]};
                std::stable_sort(pc.begin(), pc.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
                for (size_t k = 0; k < n; k++) {
                    if (k > 0 && pc[k].first == pc[k - 1].first) continue;  // duplicate split position
                    if (pc[k].second < best_cost) { best_cost = pc[k].second; best_axis = axis; best_split = pc[k].first; }
                }
            }
        } else
        for (int axis = 0; axis < 3; axis++) {  // bvh.rs:62-77
            for (size_t k = 0; k < items.size(); k++) pos[k] = cent[items[k]][axis];
            std::stable_sort(pos.begin(), pos.end());
            const int64_t n = (int64_t)pos.size();
#pragma omp parallel for schedule(dynamic, 16) if (n > 256)
            for (int64_t k = 0; k < n; k++) cost[k] = (k > 0 && pos[k] == pos[k - 1]) ? -1.0 : sah(axis, pos[k], parent, items);
            for (int64_t k = 0; k < n; k++) {
                double c = cost[k];
                if (c < 0.0) continue;  // duplicate split position: same cost as its predecessor, strict '<' never takes it
                if (c < best_cost) { best_cost = c; best_axis = axis; best_split = pos[k]; }
            }
        }
        std::vector<uint32_t> l, r;  // bvh.rs:78-81: partition keeps list order
        for (uint32_t i : items) (cent[i][best_axis] < best_split ? l : r).push_back(i);
        if (l.empty() || r.empty()) return leaf(items);  // bvh.rs:37-42 (Q3)
        int32_t me = (int32_t)tree.nodes.size();
        tree.nodes.emplace_back();
        int32_t li = build(l), ri = build(r);
        tree.nodes[me].left = li; tree.nodes[me].right = ri;
        tree.nodes[me].box = tree.nodes[li].box.merged(tree.nodes[ri].box);  // bvh.rs:46
        return me;
    }
};
}  // namespace
std::shared_ptr<BvhTree> build_bvh(const std::vector<Box>& boxes) {
    auto tree = std::make_shared<BvhTree>();
    Builder b{boxes, {}, *tree, g_build_ctx};
    for (auto& bx : boxes) b.cent.push_back(bx.centroid());
    std::vector<uint32_t> all(boxes.size()); for (uint32_t i = 0; i < boxes.size(); i++) all[i] = i;
    b.build(all);
    return tree;
}

// ---- flatten -------------------------------------------------------------------------------------
template <class K> static int find_key(const std::vector<const K*>& keys, const K* k) {
    for (size_t i = 0; i < keys.size(); i++) if (keys[i] == k) return (int)i;
    return -1;
}
uint32_t FlatScene::add_image(const ImagePtr& im) {
    if (!im) return PT_NONE;
    int f = find_key(img_keys, im.get()); if (f >= 0) return (uint32_t)f;
    img_keys.push_back(im.get()); image_owner.push_back(im);
    images.push_back(pt_image{im->rgb.data(), im->width, im->height});
    return (uint32_t)images.size() - 1;
}
uint32_t FlatScene::add_texture(const TexPtr& t) {
    if (!t) return PT_NONE;
    int f = find_key(tex_keys, t.get()); if (f >= 0) return (uint32_t)f;
    pt_texture o{}; o.kind = t->kind; o.tex1 = o.tex2 = o.image = PT_NONE; o.inv_scale = t->inv_scale; o.value = t->value.c();
    if (t->kind == PT_TEX_CHECKER) { o.tex1 = add_texture(t->tex1); o.tex2 = add_texture(t->tex2); }
    if (t->kind == PT_TEX_IMAGE) o.image = add_image(t->image);
    tex_keys.push_back(t.get()); textures.push_back(o);
    return (uint32_t)textures.size() - 1;
}
uint32_t FlatScene::add_material(const MatPtr& m) {
    int f = find_key(mat_keys, m.get()); if (f >= 0) return (uint32_t)f;
    pt_material o{}; o.kind = m->kind; o.mix_a = o.mix_b = PT_NONE;
    if (m->kind == PT_MAT_MIX) { o.mix_a = add_material(m->mix_a); o.mix_b = add_material(m->mix_b); }  // children first
    o.base_color_tex = add_texture(m->base_color); o.roughness_tex = add_texture(m->roughness); o.normal_map = add_image(m->normal_map);
    memcpy(o.p, m->p, sizeof(o.p));
    mat_keys.push_back(m.get()); materials.push_back(o);
    return (uint32_t)materials.size() - 1;
}
static pt_quad quad_abi(const Hittable& h, uint32_t mat) { return pt_quad{h.q.c(), h.u.c(), h.v.c(), h.w.c(), h.normal.c(), h.d, mat, 0}; }
uint32_t FlatScene::emit_tree(const BvhTree& tree, const std::vector<pt_ref>& item_refs) {
    uint32_t base = (uint32_t)nodes.size();
    for (auto& n : tree.nodes) {
        pt_bvh_node o{};
        o.bmin[0] = n.box.lo.x; o.bmin[1] = n.box.lo.y; o.bmin[2] = n.box.lo.z; o.bmax[0] = n.box.hi.x; o.bmax[1] = n.box.hi.y; o.bmax[2] = n.box.hi.z;
        if (n.left >= 0) { o.left = base + (uint32_t)n.left; o.right = base + (uint32_t)n.right; o.first_ref = 0; o.n_refs = 0; }
        else {
            o.left = o.right = PT_NONE; o.first_ref = (uint32_t)leaf_refs.size(); o.n_refs = (uint32_t)n.items.size();
            for (uint32_t it : n.items) leaf_refs.push_back(item_refs[it]);
        }
        nodes.push_back(o);
    }
    return base;
}
pt_ref FlatScene::add_hittable(const HitPtr& h, bool allow_instance) {
    switch (h->kind) {
        case PT_PRIM_SPHERE:
            spheres.push_back(pt_sphere{h->p1.c(), h->p2.c(), h->radius, add_material(h->material), h->moving ? 1u : 0u});
            return pt_ref{PT_PRIM_SPHERE, (uint32_t)spheres.size() - 1};
        case PT_PRIM_QUAD:
            quads.push_back(quad_abi(*h, add_material(h->material)));
            return pt_ref{PT_PRIM_QUAD, (uint32_t)quads.size() - 1};
        case PT_OBJ_CUBOID: {
            int f = find_key(cuboid_keys, h.get()); if (f >= 0) return pt_ref{PT_OBJ_CUBOID, (uint32_t)f};
            uint32_t mat = add_material(h->material);
            pt_cuboid c{(uint32_t)quads.size(), mat, h->a.c(), h->b.c()};
            for (auto& s : h->sides.objects) quads.push_back(quad_abi(*s, mat));
            cuboid_keys.push_back(h.get()); cuboids.push_back(c);
            return pt_ref{PT_OBJ_CUBOID, (uint32_t)cuboids.size() - 1};
        }
        case PT_OBJ_MESH: {
            int f = find_key(mesh_keys, h.get()); if (f >= 0) return pt_ref{PT_OBJ_MESH, (uint32_t)f};
            pt_mesh m{}; m.first_triangle = (uint32_t)triangles.size(); m.n_triangles = (uint32_t)h->triangles.size();
            m.material = add_material(h->material); m.has_normals = !h->tri_normals.empty(); m.has_uvs = !h->tri_uvs.empty();
            triangles.insert(triangles.end(), h->triangles.begin(), h->triangles.end());
            tri_normals.resize(3 * (size_t)m.first_triangle, pt_vec3{0, 0, 0});
            tri_uvs.resize(6 * (size_t)m.first_triangle, 0.0);
            if (m.has_normals) tri_normals.insert(tri_normals.end(), h->tri_normals.begin(), h->tri_normals.end());
            if (m.has_uvs) tri_uvs.insert(tri_uvs.end(), h->tri_uvs.begin(), h->tri_uvs.end());
            std::vector<pt_ref> refs(m.n_triangles);
            for (uint32_t k = 0; k < m.n_triangles; k++) refs[k] = pt_ref{PT_PRIM_TRIANGLE, m.first_triangle + k};
            m.bvh_root = h->mesh_bvh ? emit_tree(*h->mesh_bvh, refs) : PT_NONE;
            mesh_keys.push_back(h.get()); meshes.push_back(m);
            return pt_ref{PT_OBJ_MESH, (uint32_t)meshes.size() - 1};
        }
        case PT_OBJ_INSTANCE: {
            if (!allow_instance) throw std::runtime_error("nested Instance is outside the device subset");
            pt_instance in{}; in.child = add_hittable(h->child, false); in.axis = h->axis.c(); in.angle = h->angle; in.translation = h->translation.c();
            memcpy(in.transform, h->transform, 128); memcpy(in.inverse, h->inverse, 128); memcpy(in.normal_matrix, h->normal_matrix, 128);
            instances.push_back(in);
            return pt_ref{PT_OBJ_INSTANCE, (uint32_t)instances.size() - 1};
        }
        case PT_OBJ_VOLUME: {
            pt_volume v{}; v.boundary = add_hittable(h->child, false); v.density = h->density; v.material = add_material(h->material);
            volumes.push_back(v);
            return pt_ref{PT_OBJ_VOLUME, (uint32_t)volumes.size() - 1};
        }
    }
    throw std::runtime_error("unknown hittable kind");
}
void FlatScene::finish() {
    tri_normals.resize(3 * triangles.size(), pt_vec3{0, 0, 0});
    tri_uvs.resize(6 * triangles.size(), 0.0);
    for (size_t i = 0; i < images.size(); i++) images[i].rgb = image_owner[i]->rgb.data();
    desc.abi_version = PT_ABI_VERSION;
    desc.n_textures = (uint32_t)textures.size(); desc.n_images = (uint32_t)images.size(); desc.n_materials = (uint32_t)materials.size();
    desc.n_spheres = (uint32_t)spheres.size(); desc.n_quads = (uint32_t)quads.size(); desc.n_triangles = (uint32_t)triangles.size();
    desc.n_cuboids = (uint32_t)cuboids.size(); desc.n_meshes = (uint32_t)meshes.size(); desc.n_instances = (uint32_t)instances.size();
    desc.n_nodes = (uint32_t)nodes.size(); desc.n_leaf_refs = (uint32_t)leaf_refs.size(); desc.n_objects = (uint32_t)objects.size(); desc.n_lights = (uint32_t)lights.size();
    desc.textures = textures.data(); desc.images = images.data(); desc.materials = materials.data(); desc.spheres = spheres.data();
    desc.quads = quads.data(); desc.triangles = triangles.data(); desc.tri_normals = tri_normals.data(); desc.tri_uvs = tri_uvs.data();
    desc.cuboids = cuboids.data(); desc.meshes = meshes.data(); desc.instances = instances.data(); desc.nodes = nodes.data();
    desc.leaf_refs = leaf_refs.data(); desc.objects = objects.data(); desc.lights = lights.data();
    desc.n_volumes = (uint32_t)volumes.size(); desc.volumes = volumes.data();
}
std::unique_ptr<FlatScene> flatten(const World& world) {
    auto f = std::make_unique<FlatScene>();
    for (auto& o : world.objects.objects) f->objects.push_back(f->add_hittable(o, true));
    for (auto& o : world.lights.objects) f->lights.push_back(f->add_hittable(o, true));
    f->desc.objects_bvh_root = world.objects.bvh ? f->emit_tree(*world.objects.bvh, f->objects) : PT_NONE;
    f->desc.lights_bvh_root = world.lights.bvh ? f->emit_tree(*world.lights.bvh, f->lights) : PT_NONE;
    f->finish();
    return f;
}

// ---- camera --------------------------------------------------------------------------------------
void Camera::init() { image_height = (uint32_t)((double)image_width / aspect_ratio); }  // camera.rs:52
pt_camera Camera::to_abi(FlatScene& flat) const {
    pt_camera c{};
    c.aspect_ratio = aspect_ratio; c.image_width = image_width; c.samples_per_pixel = samples_per_pixel; c.max_depth = max_depth;
    c.vfov = vfov; c.look_from = look_from.c(); c.look_at = look_at.c(); c.vup = vup.c(); c.blur_strength = blur_strength;
    c.focal_length = focal_length; c.defocus_angle = defocus_angle; c.env_is_map = environment.is_map; c.env_color = environment.color.c();
    c.env_image = PT_NONE;
    if (environment.is_map) { c.env_image = flat.add_image(environment.map); flat.finish(); }
    return c;
}
int Camera::render(const World& world, const std::string& filename, const RenderOptions& opt, pt_stats* stats_out) const {
    auto flat = flatten(world);
    pt_camera cam = to_abi(*flat);
    return render_flat(flat->desc, cam, filename, opt, stats_out);
}
// camera.rs:109-123: tonemap the mean radiance and save the PNG
static void save_image(const std::vector<float>& mean, uint32_t W, uint32_t H, const std::string& filename) {
    std::vector<uint8_t> rgb(mean.size());
    for (size_t i = 0; i < mean.size(); i++) {  // camera.rs:109-114,128-130
        double g = std::sqrt(std::fmax((double)mean[i], 0.0));
        double v = (g < 0.0 ? 0.0 : (g > 0.999 ? 0.999 : g)) * 256.0;
        rgb[i] = std::isnan(v) ? 0 : (uint8_t)v;
    }
    if (!write_png_rgb8(filename, rgb.data(), W, H)) fprintf(stderr, "Failed to save image %s\n", filename.c_str());  // camera.rs:118-123
}
static void report(const RenderOptions& opt, uint32_t W, uint32_t H, uint32_t spp, const pt_stats& st) {
    if (opt.verbose)
        fprintf(stderr, "[pt_b200] %ux%u spp=%u on %d GPU(s): %.3f s device, %.1f Mrays/s, %.3g samples/s, %llu segments, %llu non-finite\n", W, H, spp,
                std::max(opt.gpus, 1), st.device_ms * 1e-3, st.segments / (st.device_ms * 1e3), st.paths / (st.device_ms * 1e-3),
                (unsigned long long)st.segments, (unsigned long long)st.nonfinite);
}
int render_flat(const pt_scene_desc& desc, const pt_camera& cam, const std::string& filename, const RenderOptions& opt, pt_stats* stats_out) {
    const uint32_t samples_per_pixel = cam.samples_per_pixel;
    const DeviceApi& dev = device_api();  // throws without the CUDA library: there is no CPU fallback
    uint32_t H = (uint32_t)((double)cam.image_width / cam.aspect_ratio), W = cam.image_width;  // camera.rs:52
    std::vector<float> mean((size_t)W * H * 3);
    pt_render_params p{}; p.seed = opt.seed; p.sample_begin = 0; p.sample_count = samples_per_pixel; p.sample_stride = 1; p.nan_policy = opt.nan_policy;
    if (opt.nee) p.flags |= PT_RENDER_NEE;
    if (opt.gpus > 1) {  // one process, several GPUs: spp split inside the library
        if (opt.env_importance && cam.env_is_map) p.flags |= PT_RENDER_ENV_IMPORTANCE;
        std::vector<int> devices(opt.gpus);
        for (int g = 0; g < opt.gpus; g++) devices[g] = opt.device + g;
        pt_stats st{};
        if (opt.verbose) printf("rendering production\n");  // camera.rs:101
        int rc = dev.render_multi(opt.gpus, devices.data(), &desc, &cam, &p, mean.data(), &st);
        if (rc) fprintf(stderr, "pt_render_multi: %s\n", dev.last_error());
        else { save_image(mean, W, H, filename); report(opt, W, H, samples_per_pixel, st); }
        if (stats_out) *stats_out = st;
        return rc;
    }
    pt_ctx* ctx = nullptr; pt_scene* scene = nullptr;
    int rc = dev.ctx_create(opt.device, &ctx);
    if (rc) { fprintf(stderr, "pt_ctx_create: %s\n", dev.last_error()); return rc; }
    rc = dev.scene_create(ctx, &desc, &scene);
    if (rc) { fprintf(stderr, "pt_scene_create: %s\n", dev.last_error()); dev.ctx_destroy(ctx); return rc; }
    if (opt.env_importance && cam.env_is_map) {
        rc = dev.scene_build_env_sampler(scene, cam.env_image, 0, 0);
        if (rc) { fprintf(stderr, "pt_scene_build_env_sampler: %s\n", dev.last_error()); dev.scene_destroy(scene); dev.ctx_destroy(ctx); return rc; }
        p.flags |= PT_RENDER_ENV_IMPORTANCE;
    }
    pt_stats st{};
    if (opt.verbose) printf("rendering production\n");  // camera.rs:101
    rc = dev.render(ctx, scene, &cam, &p, mean.data(), &st);
    if (rc) fprintf(stderr, "pt_render: %s\n", dev.last_error());
    else { save_image(mean, W, H, filename); report(opt, W, H, samples_per_pixel, st); }
    if (stats_out) *stats_out = st;
    dev.scene_destroy(scene); dev.ctx_destroy(ctx);
    return rc;
}

}  // namespace pt
