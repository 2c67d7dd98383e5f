// Closest-hit traversal and hit reconstruction (device side of src/hittable/*.rs).
//
// Traversal: ordered, early-out, two-level walk over the 32-byte node-pair array with a per-thread
// stack; fp32 conservative slab tests (aabb.rs:31-42) and f64 primitive tests in the reference's exact
// operation order (sphere.rs:64-100, quad.rs:40-70, mesh.rs:50-112, instance.rs:34-54).  The reference's
// recursive, un-narrowed traversal (bvh.rs:124-164) is reproduced through the order-independent rule of
// minimum t wins; an exact tie goes to the larger precomputed tie rank.
#pragma once
#include "device_scene.cuh"

namespace ptd {

struct RayD { d3 o, d; double time; };
PT_D RayD make_ray(d3 o, d3 d, double time) { RayD r; r.o = o; r.d = normalize(d); r.time = time; return r; }  // ray.rs:23-29
PT_D d3 ray_at(const RayD& r, double t) { return r.o + r.d * t; }                                              // ray.rs:31-33

// ---------------------------------------------------------------- primitive tests: return t or NaN-free "no hit" (-1)
// Lower bounds follow each primitive's own rule; the upper bound is applied by the caller's candidate logic.
PT_D bool sphere_t(const DSphere& s, const RayD& r, double t_min, double& t_out) {  // sphere.rs:64-86
    d3 p1 = mk(s.p1[0], s.p1[1], s.p1[2]), p2 = mk(s.p2[0], s.p2[1], s.p2[2]);
    d3 c = p1 + (p2 - p1) * r.time;
    d3 l = c - r.o;
    double sdot = dot(l, r.d);
    double l2 = dot(l, l);
    double rad = fmax(s.radius, 0.0);
    double r2 = rad * rad;
    if (sdot < 0.0 && l2 > r2) return false;
    double d2 = l2 - sdot * sdot;
    if (d2 > r2) return false;
    double q = sqrt(r2 - d2);
    double t = l2 > r2 ? sdot - q : sdot + q;
    if (t <= t_min) return false;  // exclusive (sphere.rs:84)
    t_out = t;
    return true;
}
PT_D bool quad_t(const DQuad& qd, const RayD& r, double t_min, double& t_out, double& alpha, double& beta) {  // quad.rs:40-58
    d3 n = mk(qd.n[0], qd.n[1], qd.n[2]);
    double nd = dot(n, r.d);
    if (fabs(nd) < 1e-8) return false;
    double t = (qd.d - dot(n, r.o)) / nd;
    if (!(t_min <= t)) return false;  // inclusive lower bound (interval.rs:26-28)
    d3 p = ray_at(r, t) - mk(qd.q[0], qd.q[1], qd.q[2]);
    d3 w = mk(qd.w[0], qd.w[1], qd.w[2]);
    alpha = dot(w, cross(p, mk(qd.v[0], qd.v[1], qd.v[2])));
    beta = dot(w, cross(mk(qd.u[0], qd.u[1], qd.u[2]), p));
    if (!(0.0 <= alpha && alpha <= 1.0) || !(0.0 <= beta && beta <= 1.0)) return false;
    t_out = t;
    return true;
}
PT_D bool tri_t(const DTri& tr, const RayD& r, double t_min, double& t_out, double& u, double& v) {  // mesh.rs:50-82
    d3 e1 = mk(tr.e1[0], tr.e1[1], tr.e1[2]), e2 = mk(tr.e2[0], tr.e2[1], tr.e2[2]);
    d3 h = cross(r.d, e2);
    double a = dot(e1, h);
    if (fabs(a) < 1e-8) return false;
    double f = 1.0 / a;
    d3 s = r.o - mk(tr.v0[0], tr.v0[1], tr.v0[2]);
    u = f * dot(s, h);
    if (!(0.0 <= u && u <= 1.0)) return false;
    d3 q = cross(s, e1);
    v = f * dot(r.d, q);
    if (v < 0.0 || u + v > 1.0) return false;
    double t = f * dot(e2, q);
    if (!(t_min <= t)) return false;
    t_out = t;
    return true;
}
PT_D RayD instance_local_ray(const DInstance& in, const RayD& r) {  // instance.rs:36-38
    d3 lo = xform_point(in.inv, r.o);
    d3 ld = xform_vector(in.inv, r.d);
    return make_ray(lo, ld, r.time);
}

// ---------------------------------------------------------------- fp32 slab test state
// Conservative by construction: the device boxes are rounded outward and padded by 8 ulp of their largest
// coordinate on upload (covers the rounding of inv and of the products), and the per-ray slack `e` below
// covers the rounding of the origin (|o| * 2^-22 per axis, scaled into t).  A NaN plane distance (0 * inf)
// is dropped by fminf/fmaxf, which only ever widens the accepted interval.
struct BoxRay { float ix, iy, iz, nx, ny, nz, fx, fy, fz; };  // inv dir; -o*inv -/+ slack for the near / far planes
PT_D BoxRay make_boxray(const RayD& r) {
    BoxRay b;
    const float ox = __double2float_rn(r.o.x), oy = __double2float_rn(r.o.y), oz = __double2float_rn(r.o.z);
    // |1/d| is clamped to 2^100: with an infinite reciprocal the FMA form lo*inv - o*inv would be inf - inf = NaN and the
    // axis would stop culling (rays exactly parallel to a slab, e.g. light-to-light samples in the Cornell scenes).  With a
    // huge finite reciprocal the sign of (lo - o) still decides, and the slack below covers the cancellation error.
    const float big = 1.2676506e30f;
    b.ix = fminf(fmaxf(__frcp_rn(__double2float_rn(r.d.x)), -big), big);
    b.iy = fminf(fmaxf(__frcp_rn(__double2float_rn(r.d.y)), -big), big);
    b.iz = fminf(fmaxf(__frcp_rn(__double2float_rn(r.d.z)), -big), big);
    const float k = 2.384185791015625e-07f;  // 2^-22
    const float ex = fabsf(ox * b.ix) * k, ey = fabsf(oy * b.iy) * k, ez = fabsf(oz * b.iz) * k;
    b.nx = -ox * b.ix - ex; b.ny = -oy * b.iy - ey; b.nz = -oz * b.iz - ez;
    b.fx = -ox * b.ix + ex; b.fy = -oy * b.iy + ey; b.fz = -oz * b.iz + ez;
    return b;
}
// returns entry distance (clamped at t_min) or +inf on miss; t_max is the current closest hit rounded up
PT_D float slab(const DNode& n, const BoxRay& b, float t_min, float t_max) {
    const bool sx = b.ix < 0.f, sy = b.iy < 0.f, sz = b.iz < 0.f;
    float x0 = __fmaf_rn(sx ? n.hi[0] : n.lo[0], b.ix, b.nx), x1 = __fmaf_rn(sx ? n.lo[0] : n.hi[0], b.ix, b.fx);
    float y0 = __fmaf_rn(sy ? n.hi[1] : n.lo[1], b.iy, b.ny), y1 = __fmaf_rn(sy ? n.lo[1] : n.hi[1], b.iy, b.fy);
    float z0 = __fmaf_rn(sz ? n.hi[2] : n.lo[2], b.iz, b.nz), z1 = __fmaf_rn(sz ? n.lo[2] : n.hi[2], b.iz, b.fz);
    float tn = fmaxf(fmaxf(x0, y0), fmaxf(z0, t_min));
    float tf = fminf(fminf(x1, y1), fminf(z1, t_max));
    return tn <= tf ? tn : __int_as_float(0x7fc00000);  // NaN on a miss: fails every `<=` test, even against t_max = +inf
}

// the same test on a bare (lo xyz, hi xyz) box
PT_D float slab6(const float* __restrict__ b6, const BoxRay& b, float t_min, float t_max) {
    const bool sx = b.ix < 0.f, sy = b.iy < 0.f, sz = b.iz < 0.f;
    float x0 = __fmaf_rn(sx ? b6[3] : b6[0], b.ix, b.nx), x1 = __fmaf_rn(sx ? b6[0] : b6[3], b.ix, b.fx);
    float y0 = __fmaf_rn(sy ? b6[4] : b6[1], b.iy, b.ny), y1 = __fmaf_rn(sy ? b6[1] : b6[4], b.iy, b.fy);
    float z0 = __fmaf_rn(sz ? b6[5] : b6[2], b.iz, b.nz), z1 = __fmaf_rn(sz ? b6[2] : b6[5], b.iz, b.fz);
    float tn = fmaxf(fmaxf(x0, y0), fmaxf(z0, t_min));
    float tf = fminf(fminf(x1, y1), fminf(z1, t_max));
    return tn <= tf ? tn : __int_as_float(0x7fc00000);
}

// ---------------------------------------------------------------- 4-wide box step of the mesh-walk layout (DWide2)
// The same step with the near / far planes picked by ADDRESS: which of lo / hi is the near plane depends only on the sign of
// the ray direction, so three byte offsets replace the 24 per-child selects (20 % of the step's instructions in k_mesh_walk).
#ifndef PT_WALK_SORT
#define PT_WALK_SORT 1   // 1: entered children fully sorted by entry distance; 0: only the nearest is singled out
#endif
PT_D void wide2_step_addr(const DWide2* __restrict__ node, const BoxRay& br, float tmin_f, float tmax_f, uint32_t e[4], float t[4]) {
    const char* nb = reinterpret_cast<const char*>(node);
    const uint32_t ox = br.ix < 0.f ? 0x30u : 0x00u, oy = br.iy < 0.f ? 0x40u : 0x10u, oz = br.iz < 0.f ? 0x50u : 0x20u;
    const float4 nx = *reinterpret_cast<const float4*>(nb + ox), fx = *reinterpret_cast<const float4*>(nb + (0x30u - ox));
    const float4 ny = *reinterpret_cast<const float4*>(nb + oy), fy = *reinterpret_cast<const float4*>(nb + (0x50u - oy));
    const float4 nz = *reinterpret_cast<const float4*>(nb + oz), fz = *reinterpret_cast<const float4*>(nb + (0x70u - oz));
    const uint4 ch = *reinterpret_cast<const uint4*>(nb + 0x60u);
    const float kInf = __int_as_float(0x7f800000);
#define PT_SLAB4(I, K)                                                                                                   \
    {                                                                                                                    \
        const float x0 = __fmaf_rn(nx.I, br.ix, br.nx), x1 = __fmaf_rn(fx.I, br.ix, br.fx);                             \
        const float y0 = __fmaf_rn(ny.I, br.iy, br.ny), y1 = __fmaf_rn(fy.I, br.iy, br.fy);                             \
        const float z0 = __fmaf_rn(nz.I, br.iz, br.nz), z1 = __fmaf_rn(fz.I, br.iz, br.fz);                             \
        const float tn = fmaxf(fmaxf(x0, y0), fmaxf(z0, tmin_f)), tf = fminf(fminf(x1, y1), fminf(z1, tmax_f));         \
        t[K] = tn <= tf ? tn : kInf;                                                                                     \
    }
    PT_SLAB4(x, 0) PT_SLAB4(y, 1) PT_SLAB4(z, 2) PT_SLAB4(w, 3)
#undef PT_SLAB4
    e[0] = ch.x; e[1] = ch.y; e[2] = ch.z; e[3] = ch.w;
#define PT_CSWAP(A, B) { const bool s_ = t[B] < t[A]; const float tt = s_ ? t[A] : t[B]; t[A] = s_ ? t[B] : t[A]; t[B] = tt; \
                         const uint32_t ee = s_ ? e[A] : e[B]; e[A] = s_ ? e[B] : e[A]; e[B] = ee; }
#if PT_WALK_SORT
    PT_CSWAP(0, 1) PT_CSWAP(2, 3) PT_CSWAP(0, 2) PT_CSWAP(1, 3) PT_CSWAP(1, 2)
#else
    PT_CSWAP(0, 1) PT_CSWAP(2, 3) PT_CSWAP(0, 2)   // slot 0 = nearest; the rest in no particular order
#endif
#undef PT_CSWAP
}


// ---------------------------------------------------------------- traversal
constexpr int kStack = 64;
constexpr uint32_t kTagRef = 0x40000000u, kTagSentinel = 0x80000000u, kTagMask = 0xC0000000u;

struct Closest {
    double t; uint32_t ref, inst, tie_outer, tie_inner; bool is_light;
    uint32_t n_pairs, n_wide, n_refs, n_prims;  // work counters: binary pairs (64 B) / wide nodes (128 B) fetched, reference boxes (32 B), f64 primitive tests
};
PT_D void consider(Closest& c, double t, uint32_t ref, uint32_t inst, uint32_t tie_o, uint32_t tie_i) {
    if (t < c.t || (t == c.t && (tie_o > c.tie_outer || (tie_o == c.tie_outer && tie_i > c.tie_inner)))) {
        c.t = t; c.ref = ref; c.inst = inst; c.tie_outer = tie_o; c.tie_inner = tie_i;
    }
}
// A direct (non-BVH) primitive or a cuboid's six quads, in the space of `r`.
PT_D void test_simple(const DScene& S, uint32_t kind, uint32_t index, const RayD& r, double t_min, Closest& c, uint32_t inst,
                      uint32_t tie_o, uint32_t tie_i) {
    double t, a, b;
    PT_ASSERT(kind == PT_PRIM_SPHERE ? index < S.n_spheres : kind == PT_PRIM_QUAD ? index < S.n_quads : kind == PT_OBJ_CUBOID);
    if (kind == PT_PRIM_SPHERE) {
        // exclusive upper bound against the initial interval (sphere.rs:84); later ties go through the ranks
        if (sphere_t(S.spheres[index], r, t_min, t) && t <= c.t && !(c.ref == kNone && t == c.t)) consider(c, t, ref_pack(PT_PRIM_SPHERE, index), inst, tie_o, tie_i);
    } else if (kind == PT_PRIM_QUAD) {
        if (quad_t(S.quads[index], r, t_min, t, a, b) && t <= c.t) consider(c, t, ref_pack(PT_PRIM_QUAD, index), inst, tie_o, tie_i);
    } else if (kind == PT_OBJ_CUBOID) {
        uint32_t fq = S.cuboids[index].first_quad;
        PT_ASSERT(fq + 6 <= S.n_quads);
        // a ray that enters the box crosses two of the six faces: each face's own fp32 box (conservative like every other box
        // here) spares the f64 plane test, division included, of the other four
        const BoxRay fbr = make_boxray(r);
        const float tmin_f = __double2float_rd(t_min);
#pragma unroll 1
        for (uint32_t k = 0; k < 6; k++) {  // linear list: later quad wins ties -> inner rank = k (cuboid.rs, list.rs:57-66)
            const float tmax_f = __double2float_ru(c.t);
            if (!(slab6(S.quad_box + 6ull * (fq + k), fbr, tmin_f, tmax_f) <= tmax_f)) continue;
            if (quad_t(S.quads[fq + k], r, t_min, t, a, b) && t <= c.t) consider(c, t, ref_pack(PT_PRIM_QUAD, fq + k), inst, tie_o, tie_i + k);
        }
    }
}

// ---------------------------------------------------------------- constant-density media (include/pt_b200.h, pt_volume; ours)
// Nearest crossing of a volume boundary (sphere or cuboid) at t > / >= t_lo as the primitive's own rule has it, plus the
// HitInfo::new front_face of that crossing (hit_info.rs:27: direction . outward normal < 0).
PT_D bool boundary_hit(const DScene& S, uint32_t kind, uint32_t index, const RayD& r, double t_lo, double& t_out, bool& front_face) {
    if (kind == PT_PRIM_SPHERE) {
        const DSphere& s = S.spheres[index];
        double t;
        if (!sphere_t(s, r, t_lo, t) || !(t < __longlong_as_double(0x7ff0000000000000ll))) return false;
        d3 p1 = mk(s.p1[0], s.p1[1], s.p1[2]), p2 = mk(s.p2[0], s.p2[1], s.p2[2]);
        d3 normal = normalize(ray_at(r, t) - (p1 + (p2 - p1) * r.time));
        t_out = t; front_face = dot(r.d, normal) < 0.0;
        return true;
    }
    const uint32_t fq = S.cuboids[index].first_quad;  // six quads, linear list: a later quad wins an equal t (list.rs:57-66)
    bool any = false; double best = __longlong_as_double(0x7ff0000000000000ll);
#pragma unroll 1
    for (uint32_t k = 0; k < 6; k++) {
        double t, a, b;
        if (quad_t(S.quads[fq + k], r, t_lo, t, a, b) && t <= best) {
            best = t; any = true;
            front_face = dot(r.d, mk(S.quads[fq + k].n[0], S.quads[fq + k].n[1], S.quads[fq + k].n[2])) < 0.0;
        }
    }
    t_out = best;
    return any;
}
// Scatter distance inside volume `vi` for ray r on [t_min, inf), given the ray's keyed uniform u.
PT_D bool volume_t(const DScene& S, uint32_t vi, const RayD& r, double t_min, double u, double& t_out) {
    const DVolume v = S.volumes[vi];
    double t1; bool ff;
    if (!boundary_hit(S, v.child_kind, v.child_index, r, t_min, t1, ff)) return false;
    double t_in, t_exit;
    if (ff) {  // entering: find the exit from just inside (sphere.rs:80 never returns the far root to an outside origin)
        t_in = t1;
        const double step = t_in + 1e-4;
        RayD inner; inner.o = ray_at(r, step); inner.d = r.d; inner.time = r.time;
        double t2; bool ff2;
        if (!boundary_hit(S, v.child_kind, v.child_index, inner, 0.0, t2, ff2)) return false;
        t_exit = step + t2;
    } else { t_in = t_min; t_exit = t1; }
    const double s = v.neg_inv_density * log(u);
    if (s > t_exit - t_in) return false;
    const double t = t_in + s;
    if (!(t_min <= t)) return false;
    t_out = t;
    return true;
}
struct NoVol { static constexpr bool kEnabled = false; PT_D double operator()(uint32_t) const { return 1.0; } };

// World::intersect_all(ray, [t_min, inf)) — world.rs:47-62.  any_hit: stop at the first hit with t <= t_max on World.objects.
//
// "while-while" structure: phase 1 walks 4-wide internal nodes only (fp32 slab tests, four children per 128-byte fetch,
// sorted near to far); leaves, deferred mesh/instance references and the instance-exit sentinel are pushed as tagged
// stack entries and handled in phase 2, so the lanes of a warp run box tests together and f64 primitive tests together.
constexpr uint32_t kTagLeaf = 0xC0000000u;  // kTagRef = 0x4..., kTagSentinel = 0x8..., internal node = 0x0...
constexpr uint32_t kWideBit = 0x20000000u;  // internal entries: set = index of a 4-wide node, clear = index of a binary node pair
// `reload()` returns the world ray again (called when an instance is left), so it need not be kept in registers;
// COUNT enables the work counters reported by pt_trace_closest.
// `vol_u(volume index)` returns the ray's keyed uniform for that medium; VolU::kEnabled is false for scenes without media,
// whose kernels then carry none of the medium code.
// `defer(reference slot, entry distance)` (Defer::kEnabled): asked before a mesh BLAS is entered; true = the caller has queued
// the reference for a later pass (k_trace_blas -> trace_blas) and the traversal moves on.
struct NoDefer { static constexpr bool kEnabled = false; PT_D bool operator()(uint32_t, float) const { return false; } };
template <bool ANY_HIT, bool COUNT, bool WIDE, class Reload, class VolU = NoVol, class Defer = NoDefer>
PT_D bool trace_closest(const DScene& S, Reload reload, double t_min, double t_max_any, Closest& c, VolU vol_u = VolU(), Defer defer = Defer()) {
    // (entry, fp32 entry distance) in one 64-bit local-memory word: a pop is a single load (+2 % over two 32-bit arrays).
    // Keeping the shallow 8 / 16 slots of every lane in shared memory instead measured -4 % / -7 % (less L1 for the nodes).
    uint2 stack[kStack];
    float pending_t = 0.f;
#define PT_PUSH(E, T) { stack[sp] = make_uint2((E), __float_as_uint(T)); sp++; }
    int sp = 0;
    c.t = ANY_HIT ? t_max_any : __longlong_as_double(0x7ff0000000000000ll);
    c.ref = kNone; c.inst = kInstNone; c.tie_outer = 0; c.tie_inner = 0; c.is_light = false; c.n_pairs = 0; c.n_wide = 0; c.n_refs = 0; c.n_prims = 0;
    RayD r = reload();
    BoxRay br = make_boxray(r);
    const float tmin_f = __double2float_rd(t_min);
    float tmax_f = __double2float_ru(c.t);
    uint32_t cur_inst = kInstNone, cur_tie = 0;
    uint32_t cur = S.root_entry;  // internal entry (binary pair or wide node) to visit next, or kNone
    while (true) {
        uint32_t pending = kNone;
        // ---------------- phase 1: internal pairs until a tagged entry surfaces (or the stack runs dry)
        while (true) {
            if (cur == kNone) {
                while (sp > 0) {
                    --sp;
                    const uint2 top = stack[sp];
                    const uint32_t e = top.x;
                    if (e != kTagSentinel && !(__uint_as_float(top.y) <= tmax_f)) continue;  // beyond the current closest hit
                    if ((e & kTagMask) == 0) cur = e; else { pending = e; pending_t = __uint_as_float(top.y); }
                    break;
                }
                if (cur == kNone) break;  // pending entry, or nothing left
            }
            if (!WIDE) {  // compile-time: a scene is traversed entirely with binary pairs or entirely with wide nodes
                // ---- binary pair (small BVHs: top-level lists of a few objects; most rays leave after one or two fetches)
                PT_ASSERT(cur + 1 < S.n_nodes);
                const DNode n0 = S.nodes[cur], n1 = S.nodes[cur + 1];
                if (COUNT) c.n_pairs++;
                const float t0 = slab(n0, br, tmin_f, tmax_f), t1 = slab(n1, br, tmin_f, tmax_f);
                const uint32_t e0 = n0.b == kNone ? n0.a : (kTagLeaf | cur), e1 = n1.b == kNone ? n1.a : (kTagLeaf | (cur + 1));
                const bool h0 = t0 <= tmax_f, h1 = t1 <= tmax_f;  // NaN (miss) compares false
                const bool swap = h1 && (!h0 || t1 < t0);        // near child first
                const uint32_t en = swap ? e1 : e0, ef = swap ? e0 : e1;
                const float tn = swap ? t1 : t0, tf = swap ? t0 : t1;
                const bool hn = swap ? h1 : h0, hf = swap ? h0 : h1;
                cur = kNone;
                if (hf && can_push(sp, kStack)) PT_PUSH(ef, tf)
                if (hn) {
                    if ((en & kTagMask) == 0) cur = en;
                    else if (can_push(sp, kStack)) PT_PUSH(en, tn)
                }
                continue;
            }
            // ---- 4-wide node (large BVHs: meshes, big object lists): 7 x 128-bit loads, four slabs, sorted near to far
            PT_ASSERT((cur & ~kWideBit) < S.n_wide);
            const DWide& w = S.wide[cur & ~kWideBit];
            const float4 lx = *reinterpret_cast<const float4*>(w.lo[0]), ly = *reinterpret_cast<const float4*>(w.lo[1]),
                         lz = *reinterpret_cast<const float4*>(w.lo[2]), hx = *reinterpret_cast<const float4*>(w.hi[0]),
                         hy = *reinterpret_cast<const float4*>(w.hi[1]), hz = *reinterpret_cast<const float4*>(w.hi[2]);
            const uint4 ch = *reinterpret_cast<const uint4*>(w.child);
            if (COUNT) c.n_wide++;
            const float kInf = __int_as_float(0x7f800000);
            const bool sx = br.ix < 0.f, sy = br.iy < 0.f, sz = br.iz < 0.f;
#define PT_SLAB4(I)                                                                                                          \
            float t##I;                                                                                                      \
            {                                                                                                                \
                const float x0 = __fmaf_rn(sx ? hx.I : lx.I, br.ix, br.nx), x1 = __fmaf_rn(sx ? lx.I : hx.I, br.ix, br.fx); \
                const float y0 = __fmaf_rn(sy ? hy.I : ly.I, br.iy, br.ny), y1 = __fmaf_rn(sy ? ly.I : hy.I, br.iy, br.fy); \
                const float z0 = __fmaf_rn(sz ? hz.I : lz.I, br.iz, br.nz), z1 = __fmaf_rn(sz ? lz.I : hz.I, br.iz, br.fz); \
                const float tn = fmaxf(fmaxf(x0, y0), fmaxf(z0, tmin_f)), tf = fminf(fminf(x1, y1), fminf(z1, tmax_f));     \
                t##I = tn <= tf ? tn : kInf; /* NaN planes (0*inf) drop out of fmaxf/fminf: conservative */                  \
            }
            PT_SLAB4(x) PT_SLAB4(y) PT_SLAB4(z) PT_SLAB4(w)
#undef PT_SLAB4
            uint32_t e0 = ch.x, e1 = ch.y, e2 = ch.z, e3 = ch.w;
            float t0 = tx, t1 = ty, t2 = tz, t3 = tw;
#define PT_CSWAP(A, B) { const bool s_ = t##B < t##A; const float tt = s_ ? t##A : t##B; t##A = s_ ? t##B : t##A; t##B = tt; \
                         const uint32_t ee = s_ ? e##A : e##B; e##A = s_ ? e##B : e##A; e##B = ee; }
            PT_CSWAP(0, 1) PT_CSWAP(2, 3) PT_CSWAP(0, 2) PT_CSWAP(1, 3) PT_CSWAP(1, 2)  // ascending entry distance
#undef PT_CSWAP
            cur = kNone;
            if (t3 < kInf && can_push(sp, kStack)) PT_PUSH(e3, t3)  // far children first: nearest is popped first
            if (t2 < kInf && can_push(sp, kStack)) PT_PUSH(e2, t2)
            if (t1 < kInf && can_push(sp, kStack)) PT_PUSH(e1, t1)
            if (t0 < kInf) {
                if ((e0 & kTagMask) == 0) cur = e0;
                else if (can_push(sp, kStack)) PT_PUSH(e0, t0)
            }
        }
        if (pending == kNone) break;  // traversal finished
        // ---------------- phase 2: one tagged entry
        if (pending == kTagSentinel) {  // leave the instance / mesh: back to the world ray
            r = reload(); br = make_boxray(r); cur_inst = kInstNone; cur_tie = 0;
            continue;
        }
        if ((pending & kTagMask) == kTagLeaf) {
            PT_ASSERT((pending & ~kTagMask) < S.n_nodes);
            const DNode& n = S.nodes[pending & ~kTagMask];
            const uint32_t first = n.a, count = n.b;
            PT_ASSERT(first + count <= S.n_refs);
            const float leaf_t = pending_t;  // entry distance of this leaf
            for (uint32_t k = 0; k < count; k++) {
                const DNode rb = S.refs[first + k];  // per-reference fp32 box + (kind|index, tie rank); prefetching the next one: -5 %
                if (COUNT) c.n_refs++;
                if (!(slab(rb, br, tmin_f, tmax_f) <= tmax_f)) continue;  // most f64 tests would be rejections: cull them in fp32
                if (COUNT) c.n_prims++;
                const uint32_t kind = ref_kind(rb.a), index = ref_index(rb.a);
                if (kind == PT_PRIM_TRIANGLE) {
                    double t, u, v;
                    PT_ASSERT(index < S.n_tris);
                    if (tri_t(S.tris[index], r, t_min, t, u, v) && t <= c.t) { consider(c, t, rb.a, cur_inst, cur_tie, rb.b); tmax_f = __double2float_ru(c.t); }
                } else {
                    if (ANY_HIT && !(rb.b >> 31)) continue;  // shadow rays test World.objects only (world.rs:31-36)
                    if (kind <= PT_OBJ_CUBOID) { test_simple(S, kind, index, r, t_min, c, kInstNone, rb.b, 0); tmax_f = __double2float_ru(c.t); }
                    else if (can_push(sp, kStack)) {  // mesh / instance: defer (order does not matter, ties use ranks)
                        PT_PUSH(kTagRef | (first + k), leaf_t)
                    }
                }
                if (ANY_HIT && c.ref != kNone) return true;
            }
            continue;
        }
        // deferred mesh / instance reference
        const DRef rf{S.refs[pending & ~kTagMask].a, S.refs[pending & ~kTagMask].b};
        const uint32_t kind = ref_kind(rf.kind_index), index = ref_index(rf.kind_index);
        if (Defer::kEnabled) {
            const bool to_mesh = kind == PT_OBJ_MESH || (kind == PT_OBJ_INSTANCE && S.instances[index].child_kind == PT_OBJ_MESH);
            if (to_mesh && defer(pending & ~kTagMask, pending_t)) continue;
        }
        uint32_t mesh = kNone;
        if (kind == PT_OBJ_MESH) { mesh = index; cur_inst = kInstNone; }
        else if (VolU::kEnabled && kind == PT_OBJ_VOLUME) {
            double t;
            if (volume_t(S, index, r, t_min, vol_u(index), t) && t <= c.t) consider(c, t, rf.kind_index, kInstNone, rf.tie, 0);
            if (ANY_HIT && c.ref != kNone) return true;
            tmax_f = __double2float_ru(c.t);
            continue;
        } else {  // PT_OBJ_INSTANCE
            const DInstance& in = S.instances[index];
            const RayD lr = instance_local_ray(in, r);
            if (in.child_kind == PT_OBJ_MESH) { mesh = in.child_index; r = lr; br = make_boxray(r); cur_inst = index; }
            else if (VolU::kEnabled && in.child_kind == PT_OBJ_VOLUME) {
                double t;
                if (volume_t(S, in.child_index, lr, t_min, vol_u(in.child_index), t) && t <= c.t)
                    consider(c, t, ref_pack(PT_OBJ_VOLUME, in.child_index), index, rf.tie, 0);
                if (ANY_HIT && c.ref != kNone) return true;
                tmax_f = __double2float_ru(c.t);
                continue;
            } else {
                test_simple(S, in.child_kind, in.child_index, lr, t_min, c, index, rf.tie, 0);
                if (ANY_HIT && c.ref != kNone) return true;
                tmax_f = __double2float_ru(c.t);
                continue;
            }
        }
        cur_tie = rf.tie;
        if (can_push(sp, kStack)) PT_PUSH(kTagSentinel, 0.f)
        cur = S.meshes[mesh].root_entry;
    }
    c.is_light = c.ref != kNone && !(c.tie_outer >> 31);  // objects carry bit 31 in their outer rank (object beats light, Q31)
    return c.ref != kNone;
#undef PT_PUSH
}

// Closest hit inside ONE mesh BLAS (4-wide nodes, triangle leaves) for a ray already in the mesh's space: the second pass of
// the two-pass traversal (k_trace_blas*).  `c` comes in holding the best hit so far.  Same node step, leaf loop, culling and
// tie ranks as trace_closest, without the top-level cases (instances, simple primitives, media, sentinel).
// blas_round = one while-while round: wide nodes until a leaf surfaces, then that leaf; the traversal state between rounds is
// the stack alone, so a caller may stop after any round (the persistent kernel does, to hand idle lanes a new ray).
// Returns true when the stack has run dry (traversal finished).
template <bool COUNT>
PT_D bool blas_round(const DScene& S, const RayD& r, const BoxRay& br, double t_min, float tmin_f, float& tmax_f, uint2* stack, int& sp, Closest& c,
                     uint32_t cur_inst, uint32_t cur_tie) {
    uint32_t pending = kNone, cur = kNone;
    while (true) {  // phase 1: wide nodes
        if (cur == kNone) {
            while (sp > 0) {
                --sp;
                const uint2 top = stack[sp];
                if (!(__uint_as_float(top.y) <= tmax_f)) continue;
                if ((top.x & kTagMask) == 0) cur = top.x; else pending = top.x;
                break;
            }
            if (cur == kNone) break;
        }
        PT_ASSERT((cur & ~kWideBit) < S.n_wide);
        const DWide& w = S.wide[cur & ~kWideBit];
        const float4 lx = *reinterpret_cast<const float4*>(w.lo[0]), ly = *reinterpret_cast<const float4*>(w.lo[1]),
                     lz = *reinterpret_cast<const float4*>(w.lo[2]), hx = *reinterpret_cast<const float4*>(w.hi[0]),
                     hy = *reinterpret_cast<const float4*>(w.hi[1]), hz = *reinterpret_cast<const float4*>(w.hi[2]);
        const uint4 ch = *reinterpret_cast<const uint4*>(w.child);
        if (COUNT) c.n_wide++;
        const float kInf = __int_as_float(0x7f800000);
        const bool sx = br.ix < 0.f, sy = br.iy < 0.f, sz = br.iz < 0.f;
#define PT_SLAB4(I)                                                                                                      \
        float t##I;                                                                                                      \
        {                                                                                                                \
            const float x0 = __fmaf_rn(sx ? hx.I : lx.I, br.ix, br.nx), x1 = __fmaf_rn(sx ? lx.I : hx.I, br.ix, br.fx); \
            const float y0 = __fmaf_rn(sy ? hy.I : ly.I, br.iy, br.ny), y1 = __fmaf_rn(sy ? ly.I : hy.I, br.iy, br.fy); \
            const float z0 = __fmaf_rn(sz ? hz.I : lz.I, br.iz, br.nz), z1 = __fmaf_rn(sz ? lz.I : hz.I, br.iz, br.fz); \
            const float tn = fmaxf(fmaxf(x0, y0), fmaxf(z0, tmin_f)), tf = fminf(fminf(x1, y1), fminf(z1, tmax_f));     \
            t##I = tn <= tf ? tn : kInf;                                                                                 \
        }
        PT_SLAB4(x) PT_SLAB4(y) PT_SLAB4(z) PT_SLAB4(w)
#undef PT_SLAB4
        uint32_t e0 = ch.x, e1 = ch.y, e2 = ch.z, e3 = ch.w;
        float t0 = tx, t1 = ty, t2 = tz, t3 = tw;
#define PT_CSWAP(A, B) { const bool s_ = t##B < t##A; const float tt = s_ ? t##A : t##B; t##A = s_ ? t##B : t##A; t##B = tt; \
                         const uint32_t ee = s_ ? e##A : e##B; e##A = s_ ? e##B : e##A; e##B = ee; }
        PT_CSWAP(0, 1) PT_CSWAP(2, 3) PT_CSWAP(0, 2) PT_CSWAP(1, 3) PT_CSWAP(1, 2)
#undef PT_CSWAP
        cur = kNone;
        if (t3 < kInf && can_push(sp, kStack)) { stack[sp] = make_uint2(e3, __float_as_uint(t3)); sp++; }
        if (t2 < kInf && can_push(sp, kStack)) { stack[sp] = make_uint2(e2, __float_as_uint(t2)); sp++; }
        if (t1 < kInf && can_push(sp, kStack)) { stack[sp] = make_uint2(e1, __float_as_uint(t1)); sp++; }
        if (t0 < kInf) {
            if ((e0 & kTagMask) == 0) cur = e0;
            else if (can_push(sp, kStack)) { stack[sp] = make_uint2(e0, __float_as_uint(t0)); sp++; }
        }
    }
    if (pending == kNone) return true;
    PT_ASSERT((pending & ~kTagMask) < S.n_nodes);
    const DNode& n = S.nodes[pending & ~kTagMask];  // phase 2: one triangle leaf
    const uint32_t first = n.a, count = n.b;
    PT_ASSERT(first + count <= S.n_refs);
    for (uint32_t k = 0; k < count; k++) {
        const DNode rb = S.refs[first + k];
        if (COUNT) c.n_refs++;
        if (!(slab(rb, br, tmin_f, tmax_f) <= tmax_f)) continue;
        if (COUNT) c.n_prims++;
        double t, u, v;
        PT_ASSERT(ref_kind(rb.a) == PT_PRIM_TRIANGLE && ref_index(rb.a) < S.n_tris);
        if (tri_t(S.tris[ref_index(rb.a)], r, t_min, t, u, v) && t <= c.t) { consider(c, t, rb.a, cur_inst, cur_tie, rb.b); tmax_f = __double2float_ru(c.t); }
    }
    return false;
}
template <bool COUNT>
PT_D void trace_blas(const DScene& S, uint32_t root_entry, const RayD& r, double t_min, Closest& c, uint32_t cur_inst, uint32_t cur_tie) {
    uint2 stack[kStack];
    int sp = 1;
    stack[0] = make_uint2(root_entry, 0u);  // entry distance 0: never culled
    const BoxRay br = make_boxray(r);
    const float tmin_f = __double2float_rd(t_min);  // the render passes Interval::new(eps, INFINITY), camera.rs:171,179
    float tmax_f = __double2float_ru(c.t);
    while (!blas_round<COUNT>(S, r, br, t_min, tmin_f, tmax_f, stack, sp, c, cur_inst, cur_tie)) {}
    c.is_light = c.ref != kNone && !(c.tie_outer >> 31);
}

// ---------------------------------------------------------------- hit reconstruction (HitInfo::new, hit_info.rs:16-55)
struct HitInfoD {
    d3 point, gn, sn; double t, u, v; bool front_face; uint32_t material;
};
PT_D d3 image_value(const DScene& S, uint32_t image, double u, double v) {  // texture.rs:72-91
    const DImage im = S.images[image];
    if (im.height == 0) return mk(0.0, 1.0, 1.0);
    u = clampd(u, 0.0, 1.0);
    v = 1.0 - clampd(v, 0.0, 1.0);
    double fi = u * (double)im.width, fj = v * (double)im.height;
    uint32_t i = fi > 0.0 ? (fi >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)fi) : 0u;  // `as u32` saturates
    uint32_t j = fj > 0.0 ? (fj >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)fj) : 0u;
    if (i >= im.width) i = im.width - 1;   // Q22: reference panics here; clamp (documented divergence)
    if (j >= im.height) j = im.height - 1;
    const uint8_t* px = S.image_data + im.offset + 3ull * ((uint64_t)j * im.width + i);
    const double s = 1.0 / 255.0;
    return mk(s * (double)px[0], s * (double)px[1], s * (double)px[2]);
}
// `unit`: normalize(normal) when the caller holds it precomputed (quads: DQuad::un)
PT_D void finish_hit(const DScene& S, const RayD& r, d3 point, d3 normal, double t, uint32_t material, double u, double v, HitInfoD& h, const d3* unit = nullptr) {
    bool ff = dot(r.d, normal) < 0.0;
    const d3 nn = unit ? *unit : normalize(normal);
    d3 gn = ff ? nn : -nn;
    d3 sn = gn;
    const DMaterial& m = S.materials[material];
    if (m.kind == PT_MAT_DIFFUSE && m.normal_map != kNone) {  // only DiffuseBRDF overrides normal_map() (diffuse.rs:81-83)
        d3 c = image_value(S, m.normal_map, u, v);
        d3 mapped = 2.0 * c - mk(1.0, 1.0, 1.0);
        d3 a = fabs(gn.x) > 0.9 ? mk(0.0, 1.0, 0.0) : mk(1.0, 0.0, 0.0);  // hit_info.rs:58-67
        d3 tangent = normalize(cross(gn, a));
        d3 bitangent = cross(gn, tangent);
        sn = normalize(mapped.x * tangent + mapped.y * bitangent + mapped.z * gn);
    }
    h.point = point; h.gn = gn; h.sn = sn; h.t = t; h.u = u; h.v = v; h.front_face = ff; h.material = material;
}
// Recompute the winning primitive's full HitInfo from (ref, instance, t): same formulas, same rounding as the oracle.
// FULL_UV = false (shading): a sphere's (u, v) — an f64 acos + atan2 — is skipped when no texture of the hit material
// reads it (DMaterial::uses_uv, resolved on upload); the parity entry points always compute it.
template <bool FULL_UV = true>
PT_D void reconstruct_hit(const DScene& S, const RayD& world_ray, uint32_t ref, uint32_t inst, double t, HitInfoD& h) {
    RayD r = world_ray;
    if (inst != kInstNone) r = instance_local_ray(S.instances[inst], world_ray);
    const uint32_t kind = ref_kind(ref), index = ref_index(ref);
    if (kind == PT_PRIM_SPHERE) {  // sphere.rs:88-99
        const DSphere& s = S.spheres[index];
        d3 p1 = mk(s.p1[0], s.p1[1], s.p1[2]), p2 = mk(s.p2[0], s.p2[1], s.p2[2]);
        d3 c = p1 + (p2 - p1) * r.time;
        d3 point = ray_at(r, t);
        d3 normal = normalize(point - c);
        double su = 0.0, sv = 0.0;
        if (FULL_UV || S.materials[s.material].uses_uv) {  // sphere.rs:52-56
            double theta = acos(-normal.y);
            double phi = atan2(-normal.z, normal.x) + kPi;
            su = phi / (2.0 * kPi); sv = theta / kPi;
        }
        finish_hit(S, r, point, normal, t, s.material, su, sv, h);
    } else if (kind == PT_PRIM_QUAD) {  // quad.rs:53-69
        const DQuad& qd = S.quads[index];
        d3 p = ray_at(r, t) - mk(qd.q[0], qd.q[1], qd.q[2]);
        d3 w = mk(qd.w[0], qd.w[1], qd.w[2]);
        double alpha = dot(w, cross(p, mk(qd.v[0], qd.v[1], qd.v[2])));
        double beta = dot(w, cross(mk(qd.u[0], qd.u[1], qd.u[2]), p));
        const d3 un = mk(qd.un[0], qd.un[1], qd.un[2]);
        finish_hit(S, r, ray_at(r, t), mk(qd.n[0], qd.n[1], qd.n[2]), t, S.quad_material[index], alpha, beta, h, &un);
    } else if (kind == PT_OBJ_VOLUME) {  // scatter point inside a medium: arbitrary normal (1,0,0), u = v = 0 (pt_volume)
        finish_hit(S, r, ray_at(r, t), mk(1.0, 0.0, 0.0), t, S.volumes[index].material, 0.0, 0.0, h);
    } else {  // triangle, mesh.rs:84-111
        const DTri& tr = S.tris[index];
        const DMesh& m = S.meshes[S.tri_mesh[index]];
        d3 e1 = mk(tr.e1[0], tr.e1[1], tr.e1[2]), e2 = mk(tr.e2[0], tr.e2[1], tr.e2[2]);
        d3 hh = cross(r.d, e2);
        double a = dot(e1, hh);
        double f = 1.0 / a;
        d3 s = r.o - mk(tr.v0[0], tr.v0[1], tr.v0[2]);
        double u = f * dot(s, hh);
        d3 q = cross(s, e1);
        double v = f * dot(r.d, q);
        double w = 1.0 - u - v;
        d3 normal;
        if (m.has_normals) {
            const double* n = S.tri_normals + 9ull * index;
            normal = normalize(mk(n[0], n[1], n[2]) * w + mk(n[3], n[4], n[5]) * u + mk(n[6], n[7], n[8]) * v);
        } else normal = normalize(cross(e1, e2));
        double ou = u, ov = v;
        if (m.has_uvs) {
            const double* t6 = S.tri_uvs + 6ull * index;
            ou = t6[0] * w + t6[2] * u + t6[4] * v;
            ov = t6[1] * w + t6[3] * u + t6[5] * v;
        }
        finish_hit(S, r, ray_at(r, t), normal, t, m.material, ou, ov, h);
    }
    if (inst != kInstNone) {  // instance.rs:43-53: point and geometric normal to world; the rest stays local (Q4)
        const DInstance& in = S.instances[inst];
        h.point = xform_point(in.fwd, h.point);
        h.gn = normalize(xform_vector(in.nrm, h.gn));
    }
}

}  // namespace ptd
