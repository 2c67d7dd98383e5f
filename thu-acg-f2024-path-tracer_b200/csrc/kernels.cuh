// Wavefront integrator kernels for sm_100a (device side of Camera::trace, src/camera.rs:170-228).
//
// Pipeline per iteration over a pool of in-flight paths held as SoA arrays in HBM:
//   k_generate  — camera rays for freshly started paths (camera.rs:153-168), appended after the survivors
//   k_trace     — World::intersect_all for every live path (world.rs:47-62) -> 16-byte hit records
//   k_shade     — miss/environment, emission, Russian roulette, light/BSDF mixture sampling, next ray
//                 (camera.rs:178-225); survivors are written COMPACTED into the other SoA buffer through a
//                 warp-ballot + block-prefix + one atomicAdd per block, so every later kernel reads dense,
//                 coalesced arrays and warps stay full.
// Radiance contributions go straight to the fp32 accumulators with red.global.add.f32.
#pragma once
#include "bsdf.cuh"

namespace ptd {

constexpr int kBlock = 128;
#ifndef PT_TRACE_BLOCK
#define PT_TRACE_BLOCK 32
#endif
// k_trace has no block-level cooperation, so its block size only decides how soon the slots of finished warps are reused:
// one warp per block measured 2-3 % faster than 128 threads on the mesh scenes (scene 6 FHD trace 38.1 -> 37.0 ms).
// MIN_BLOCKS in its launch bounds is stated for 128-thread blocks and scaled.
constexpr int kTraceBlock = PT_TRACE_BLOCK;
#ifndef PT_SHADE_MIN_BLOCKS
#define PT_SHADE_MIN_BLOCKS 4  // resident blocks per SM the shade kernels must allow (register cap 128)
#endif

struct PathBuf {
    double* f[10];  // ox oy oz dx dy dz time thr_r thr_g thr_b
    uint4* ids;     // pixel, sample, rng_used | bounce << 16, spare
};
// Importance sampler of a lat-long environment map (PT_RENDER_ENV_IMPORTANCE; not reference behaviour):
// a piecewise-constant density over rows x cols cells of the (u, theta/pi) unit square, built on the host in f64 by
// pt_scene_build_env_sampler.  marginal[rows + 1] is the CDF over rows (row 0 = theta 0 = +y), cond[r * (cols + 1) ...]
// the CDF over the columns of row r; both start at 0 and end at 1.
struct DEnvDist { const double* marginal; const double* cond; uint32_t rows, cols; };
struct RenderConst {
    uint64_t seed; uint32_t sample_begin, sample_stride, nan_policy, env_importance;
    DEnvDist env;
    uint32_t sort_mask = 0;  // survivors of a shade block are grouped by (direction octant & sort_mask): bit 0 = y, 1 = x, 2 = z
};

// ---------------------------------------------------------------- camera.rs:133-168
PT_D void random_offsets(Rng& rng, double& x, double& y) {
    double radius = sqrt(rng.next());
    double angle = rng.next() * 2.0 * kPi;
    x = radius * cos(angle); y = radius * sin(angle);
}
struct DCameraEx { DCamera c; d3 dof_right, dof_up; };  // dof_* = right/up * lens radius (camera.rs:159-161), host-derived
PT_D RayD generate_ray(const DCameraEx& cam, uint32_t row, uint32_t col, Rng& rng) {
    double bx, by; random_offsets(rng, bx, by);
    bx = bx * cam.c.blur_strength; by = by * cam.c.blur_strength;
    d3 sample_location = cam.c.pixel00 + (cam.c.pixel_dv * ((double)row + bx)) + (cam.c.pixel_du * ((double)col + by));
    double px, py; random_offsets(rng, px, py);
    d3 origin = cam.c.center + (cam.dof_right * px) + (cam.dof_up * py);
    d3 direction = sample_location - origin;
    double time = rng.next();
    return make_ray(origin, direction, time);
}
PT_D d3 sample_environment(const DScene& S, const DCamera& cam, d3 dir) {  // camera.rs:140-151
    if (!cam.env_is_map) return cam.env_color;
    double theta = acos(dir.y);
    double phi = atan2(dir.z, dir.x);
    double u = (phi + kPi) / (2.0 * kPi);
    double v = 1.0 - theta / kPi;
    return image_value(S, cam.env_image, u, v);
}
// largest i in [0, n) with cdf[i] <= u  (cdf[0] = 0, cdf[n] = 1)
PT_D uint32_t cdf_find(const double* __restrict__ cdf, uint32_t n, double u) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid; else hi = mid; }
    return lo;
}
// direction ~ the cell density; inverse of sample_environment's mapping (theta = acos(d.y), phi = atan2(d.z, d.x))
PT_D d3 env_sample(const DEnvDist& E, double u1, double u2) {
    const uint32_t r = cdf_find(E.marginal, E.rows, u1);
    const double m0 = E.marginal[r], m1 = E.marginal[r + 1];
    const double* __restrict__ row = E.cond + (size_t)r * (E.cols + 1);
    const uint32_t c = cdf_find(row, E.cols, u2);
    const double c0 = row[c], c1 = row[c + 1];
    const double fr = m1 > m0 ? (u1 - m0) / (m1 - m0) : 0.5, fc = c1 > c0 ? (u2 - c0) / (c1 - c0) : 0.5;
    const double theta = (((double)r + fr) / (double)E.rows) * kPi;
    const double phi = (((double)c + fc) / (double)E.cols) * (2.0 * kPi) - kPi;
    const double st = sin(theta);
    return mk(st * cos(phi), cos(theta), st * sin(phi));
}
PT_D double env_pdf(const DEnvDist& E, d3 dir) {  // solid-angle density of env_sample
    const double theta = acos(dir.y), phi = atan2(dir.z, dir.x);
    const double st = sin(theta);
    if (!(st > 0.0)) return 0.0;
    const double fu = (phi + kPi) / (2.0 * kPi) * (double)E.cols, fv = theta / kPi * (double)E.rows;
    uint32_t c = fu > 0.0 ? (uint32_t)fu : 0u, r = fv > 0.0 ? (uint32_t)fv : 0u;
    if (c >= E.cols) c = E.cols - 1;
    if (r >= E.rows) r = E.rows - 1;
    const double* __restrict__ row = E.cond + (size_t)r * (E.cols + 1);
    const double cell = (E.marginal[r + 1] - E.marginal[r]) * (row[c + 1] - row[c]);
    return cell * (double)E.rows * (double)E.cols / (2.0 * kPi * kPi * st);
}

PT_D void store_path(const PathBuf& b, uint32_t i, const RayD& r, d3 thr, uint4 ids) {
    b.f[0][i] = r.o.x; b.f[1][i] = r.o.y; b.f[2][i] = r.o.z; b.f[3][i] = r.d.x; b.f[4][i] = r.d.y; b.f[5][i] = r.d.z; b.f[6][i] = r.time;
    b.f[7][i] = thr.x; b.f[8][i] = thr.y; b.f[9][i] = thr.z; b.ids[i] = ids;
}
PT_D RayD load_ray(const PathBuf& b, uint32_t i) {
    RayD r;
    r.o = mk(b.f[0][i], b.f[1][i], b.f[2][i]); r.d = mk(b.f[3][i], b.f[4][i], b.f[5][i]); r.time = b.f[6][i];
    return r;
}

// g = global index of the path within this render call: sample-major so that consecutive threads take
// neighbouring pixels of the same sample (coherent primary rays).
__global__ void __launch_bounds__(kBlock) k_generate(PathBuf out, uint32_t slot0, uint32_t n_new, uint64_t g0, uint32_t n_pixels,
                                                      DCameraEx cam, RenderConst rc) {
    uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n_new) return;
    uint64_t g = g0 + i;
    uint32_t s_local = (uint32_t)(g / n_pixels), pix = (uint32_t)(g % n_pixels);
    if ((cam.c.width & 7u) == 0 && (cam.c.height & 3u) == 0) {  // a warp covers an 8x4 pixel tile: tighter ray bundles than a 32x1 strip
        const uint32_t tile = pix >> 5, within = pix & 31u, tiles_x = cam.c.width >> 3;
        pix = ((tile / tiles_x) * 4u + (within >> 3)) * cam.c.width + (tile % tiles_x) * 8u + (within & 7u);
    }
    uint32_t sample = rc.sample_begin + s_local * rc.sample_stride;
    Rng rng; rng.init(rc.seed, pix, sample, 0);
    RayD r = generate_ray(cam, pix / cam.c.width, pix % cam.c.width, rng);
    store_path(out, slot0 + i, r, mk(1, 1, 1), make_uint4(pix, sample, rng.used, 0));
}

// Shade classes: one queue and one specialised shade kernel per class, so warps shade one material kind.
enum { CLS_MISS = 0, CLS_LIGHT, CLS_DIFFUSE, CLS_METAL, CLS_GLASS, CLS_PRINCIPLED, CLS_OTHER, N_CLS };
PT_D uint32_t hit_material(const DScene& S, uint32_t ref) {
    const uint32_t kind = ref_kind(ref), index = ref_index(ref);
    if (kind == PT_PRIM_SPHERE) return S.spheres[index].material;
    if (kind == PT_PRIM_QUAD) return S.quad_material[index];
    if (kind == PT_OBJ_VOLUME) return S.volumes[index].material;
    return S.meshes[S.tri_mesh[index]].material;
}
PT_D uint32_t class_of_kind(uint32_t k) {
    return k == PT_MAT_LIGHT ? CLS_LIGHT : k == PT_MAT_DIFFUSE ? CLS_DIFFUSE : k == PT_MAT_METAL ? CLS_METAL : k == PT_MAT_GLASS ? CLS_GLASS
           : k == PT_MAT_PRINCIPLED ? CLS_PRINCIPLED : CLS_OTHER;
}
struct Queues { uint32_t* items; uint32_t* count; uint32_t stride; };  // items[cls * stride + k] = path slot

// Appends a traced path to the queue of its shade class (warp-aggregated: lanes of the same class share one atomicAdd).
// Must be called by all 32 lanes; lanes with nothing to append pass cls = N_CLS.
PT_D void queue_append(const Queues& q, uint32_t cls, uint32_t i) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cls);
    if (cls != N_CLS) {
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(q.count + cls, __popc(peers));
        base = __shfl_sync(peers, base, leader);
        q.items[(size_t)cls * q.stride + base + __popc(peers & ((1u << lane) - 1u))] = i;
    }
}

// World::intersect_all for every live path (one ray per thread), then the path joins the queue of its shade class.
// COUNT (profiling mode only): also sums the traversal work of all rays into work[3] = {node pairs, reference boxes,
// f64 primitive tests}, from which bench.py derives the bytes the device actually requests per ray.
// VOL: the scene holds constant-density media; their free-flight uniforms are keyed by (seed, pixel, sample, bounce).
struct PathVol {
    static constexpr bool kEnabled = true;
    uint64_t seed; const uint4* __restrict__ ids; uint32_t i;
    PT_D double operator()(uint32_t v) const { const uint4 id = ids[i]; return keyed_uniform(seed, id.x, id.y, id.z >> 16, v); }
};
// Two-pass traversal for scenes with mesh BLASes (DEFER): a warp of k_trace mixes rays that leave the top-level BVH after a
// node or two with rays that walk a mesh for ten times as long, so the short ones idle (8.9 of 32 lanes active on scene 6).
// With DEFER, k_trace walks the top level only — simple primitives are tested, every mesh whose box the ray enters is
// queued (up to kDeferMax per ray; a further one is walked inline) — and k_trace_blas<round> then walks the r-th queued
// mesh of each such ray, compacted so that its warps hold only rays inside a BLAS.  Hit record and tie ranks travel through
// `hits` / `ties`; a ray joins its shade-class queue after its last round.  Same closest hit, same tie rules (ranks).
constexpr int kDeferMax = 3;
struct BlasQueues { uint4* items; uint32_t* count; uint32_t stride; };  // items[round * stride + k] = {path, ref slot | last << 31, entry t, -}
struct DeferList {
    static constexpr bool kEnabled = true;
    uint32_t* n; uint32_t* slot; float* t;
    PT_D bool operator()(uint32_t s, float tt) const {
        if (*n >= (uint32_t)kDeferMax) return false;
        slot[*n] = s; t[*n] = tt; (*n)++;
        return true;
    }
};

template <int MIN_BLOCKS, bool WIDE, bool COUNT = false, bool VOL = false, bool DEFER = false>
__global__ void __launch_bounds__(kTraceBlock, MIN_BLOCKS * kBlock / kTraceBlock) k_trace(PathBuf in, uint32_t n, HitRec* __restrict__ hits, Queues q, DScene S,
                                                              unsigned long long* __restrict__ work = nullptr, uint64_t seed = 0,
                                                              const uint32_t* __restrict__ n_dev = nullptr, BlasQueues bq = BlasQueues{nullptr, nullptr, 0},
                                                              uint2* __restrict__ ties = nullptr, double t_min = 1e-3) {
    // batched tail iterations (api.cu): the host only knows an upper bound of the live count, the survivors counter of the
    // previous iteration (still in device memory) is the real one
    if (n_dev) n = min(n, *n_dev);
    const uint32_t i = blockIdx.x * kTraceBlock + threadIdx.x;
    uint32_t cls = N_CLS;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    uint32_t nd = 0, dslot[kDeferMax]; float dt[kDeferMax];
    if (i < n) {
        Closest c;
        // t_min: the render passes eps = 1e-3 (Interval::new(eps, INFINITY), camera.rs:171,179); ray batches pass their own
        if constexpr (VOL) trace_closest<false, COUNT, WIDE>(S, [&]() { return load_ray(in, i); }, t_min, 0.0, c, PathVol{seed, in.ids, i});
        else if constexpr (DEFER) trace_closest<false, COUNT, WIDE>(S, [&]() { return load_ray(in, i); }, t_min, 0.0, c, NoVol(), DeferList{&nd, dslot, dt});
        else trace_closest<false, COUNT, WIDE>(S, [&]() { return load_ray(in, i); }, t_min, 0.0, c);
        if (COUNT) { w0 = c.n_pairs + 2 * c.n_wide; w1 = c.n_refs; w2 = c.n_prims; }  // in 64-byte units
        HitRec h; h.t = c.t; h.ref = c.ref; h.inst_light = (c.inst & 0x7FFFFFFFu) | (c.is_light ? 0x80000000u : 0u);
        hits[i] = h;
        if (DEFER && nd) ties[i] = make_uint2(c.tie_outer, c.tie_inner);
        else cls = c.ref == kNone ? (uint32_t)CLS_MISS : class_of_kind(S.materials[hit_material(S, c.ref)].kind);
    }
    __syncwarp();
    queue_append(q, cls, i);
    if (DEFER) {
        const uint32_t lane = threadIdx.x & 31;
#pragma unroll
        for (int r = 0; r < kDeferMax; r++) {
            const bool has = nd > (uint32_t)r;
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, has);
            if (!b) break;
            const int leader = __ffs(b) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(bq.count + r, __popc(b));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (has) bq.items[(size_t)r * bq.stride + base + __popc(b & ((1u << lane) - 1u))] =
                         make_uint4(i, dslot[r] | (nd == (uint32_t)r + 1 ? 0x80000000u : 0u), __float_as_uint(dt[r]), 0u);
        }
    }
    if (COUNT) {
#ifdef PT_DIAG_WARP  // diagnostic build: ref_boxes := sum of per-ray cost, prim_tests := sum of the warp's maximum cost per lane
        w1 = 3 * w0 + w1 + 4 * w2 + 1; w2 = __reduce_max_sync(0xFFFFFFFFu, w1);
#endif
        w0 = __reduce_add_sync(0xFFFFFFFFu, w0); w1 = __reduce_add_sync(0xFFFFFFFFu, w1); w2 = __reduce_add_sync(0xFFFFFFFFu, w2);
        if ((threadIdx.x & 31) == 0) { atomicAdd(work, (unsigned long long)w0); atomicAdd(work + 1, (unsigned long long)w1); atomicAdd(work + 2, (unsigned long long)w2); }
    }
}

// Round `round` of the two-pass traversal: the round-th queued mesh of every ray that queued at least round + 1 of them.
// Warp-granular grid-stride loop over the compacted queue; one ray per lane.
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock, 7 * kBlock / kTraceBlock) k_trace_blas(PathBuf in, uint32_t round, BlasQueues bq, HitRec* __restrict__ hits,
                                                              uint2* __restrict__ ties, Queues q, DScene S, unsigned long long* __restrict__ work, double t_min) {
    const uint32_t count = bq.count[round];
    const uint4* __restrict__ items = bq.items + (size_t)round * bq.stride;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    for (uint32_t j = blockIdx.x * kTraceBlock + threadIdx.x; j - (threadIdx.x & 31) < count; j += gridDim.x * kTraceBlock) {
        uint32_t cls = N_CLS, i = 0;
        if (j < count) {
            const uint4 e = items[j];
            i = e.x;
            const HitRec h0 = hits[i];
            const uint2 tie = ties[i];
            Closest c; c.t = h0.t; c.ref = h0.ref; c.inst = h0.inst_light & 0x7FFFFFFFu; c.tie_outer = tie.x; c.tie_inner = tie.y;
            c.is_light = (h0.inst_light >> 31) != 0;
            if (__uint_as_float(e.z) <= __double2float_ru(c.t)) {  // the mesh may since have fallen behind the closest hit
                const DNode rf = S.refs[e.y & 0x7FFFFFFFu];  // a = kind | index of the queued mesh or instance, b = its outer tie rank
                const uint32_t kind = ref_kind(rf.a), index = ref_index(rf.a);
                RayD r = load_ray(in, i);
                uint32_t mesh = index, inst = kInstNone;
                if (kind == PT_OBJ_INSTANCE) { const DInstance& ins = S.instances[index]; r = instance_local_ray(ins, r); mesh = ins.child_index; inst = index; }
                c.n_pairs = 0; c.n_wide = 0; c.n_refs = 0; c.n_prims = 0;
                trace_blas<COUNT>(S, S.meshes[mesh].root_entry, r, t_min, c, inst, rf.b);
                if (COUNT) { w0 += c.n_pairs + 2 * c.n_wide; w1 += c.n_refs; w2 += c.n_prims; }
                HitRec h; h.t = c.t; h.ref = c.ref; h.inst_light = (c.inst & 0x7FFFFFFFu) | (c.is_light ? 0x80000000u : 0u);
                hits[i] = h;
                if (!(e.y >> 31)) ties[i] = make_uint2(c.tie_outer, c.tie_inner);
            }
            if (e.y >> 31) cls = c.ref == kNone ? (uint32_t)CLS_MISS : class_of_kind(S.materials[hit_material(S, c.ref)].kind);
        }
        __syncwarp();
        queue_append(q, cls, i);
    }
    if (COUNT) {
        w0 = __reduce_add_sync(0xFFFFFFFFu, w0); w1 = __reduce_add_sync(0xFFFFFFFFu, w1); w2 = __reduce_add_sync(0xFFFFFFFFu, w2);
        if ((threadIdx.x & 31) == 0 && (w0 | w1 | w2)) { atomicAdd(work, (unsigned long long)w0); atomicAdd(work + 1, (unsigned long long)w1); atomicAdd(work + 2, (unsigned long long)w2); }
    }
}

// The same round as a persistent kernel with lane refill: rays inside a mesh differ widely in length (most leave after the
// root node, a few walk thirty), so k_trace_blas runs at 5-6 of 32 lanes.  Here a warp keeps pulling entries from the queue
// (one atomicAdd per fetch on the round's cursor): every kBlasBurst while-while rounds the lanes meet, finished rays join
// their shade queue, and once kBlasRefillMin lanes are idle they fetch new rays.  With only two kinds of work in the loop
// (node step, triangle leaf) a fresh ray next to an old one costs little, unlike in the fused kernel.
#ifndef PT_BLAS_BURST
#define PT_BLAS_BURST 2
#endif
#ifndef PT_BLAS_REFILL_MIN
#define PT_BLAS_REFILL_MIN 8
#endif
#ifndef PT_BLAS_MIN_BLOCKS
#define PT_BLAS_MIN_BLOCKS 7  // resident 128-thread-equivalents per SM the refill kernel must allow (7 -> 72 registers, 28 warps)
#endif
constexpr int kBlasMinBlocks = PT_BLAS_MIN_BLOCKS;
constexpr int kBlasBurst = PT_BLAS_BURST;
constexpr int kBlasRefillMin = PT_BLAS_REFILL_MIN;
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock, kBlasMinBlocks * kBlock / kTraceBlock) k_trace_blas_refill(PathBuf in, uint32_t round, BlasQueues bq, HitRec* __restrict__ hits,
                                                              uint2* __restrict__ ties, Queues q, DScene S, unsigned long long* __restrict__ work, double t_min) {
    const uint32_t count = bq.count[round];
    uint32_t* __restrict__ cursor = bq.count + 4 + round;
    const uint4* __restrict__ items = bq.items + (size_t)round * bq.stride;
    const uint32_t lane = threadIdx.x & 31;
    const float tmin_f = __double2float_rd(t_min);
    uint2 stack[kStack];
    int sp = 0;
    RayD r = make_ray(mk(0, 0, 0), mk(0, 0, 1), 0.0);
    BoxRay br = make_boxray(r);
    Closest c; c.t = 0.0; c.ref = kNone; c.inst = kInstNone; c.tie_outer = 0; c.tie_inner = 0; c.is_light = false; c.n_pairs = 0; c.n_wide = 0; c.n_refs = 0; c.n_prims = 0;
    uint32_t i = 0, cur_inst = kInstNone, cur_tie = 0, done_cls = N_CLS, done_i = 0;
    float tmax_f = 0.f;
    bool active = false, last = false, drained = false;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    while (true) {
        // ---- all 32 lanes meet here: finished rays join their shade-class queue, idle lanes fetch
        if (__any_sync(0xFFFFFFFFu, done_cls != N_CLS)) { queue_append(q, done_cls, done_i); done_cls = N_CLS; }
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (!drained && __popc(idle) >= kBlasRefillMin) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(cursor, __popc(idle));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            drained = base + __popc(idle) >= count;
            const uint32_t j = base + __popc(idle & ((1u << lane) - 1u));
            if (!active && j < count) {
                const uint4 e = items[j];
                i = e.x; last = (e.y >> 31) != 0;
                const HitRec h0 = hits[i];
                const uint2 tie = ties[i];
                c.t = h0.t; c.ref = h0.ref; c.inst = h0.inst_light & 0x7FFFFFFFu; c.tie_outer = tie.x; c.tie_inner = tie.y;
                if (__uint_as_float(e.z) <= __double2float_ru(c.t)) {  // else the mesh has since fallen behind the closest hit
                    const DNode rf = S.refs[e.y & 0x7FFFFFFFu];  // a = kind | index of the queued mesh or instance, b = its outer tie rank
                    const uint32_t kind = ref_kind(rf.a), index = ref_index(rf.a);
                    r = load_ray(in, i);
                    uint32_t mesh = index;
                    cur_inst = kInstNone;
                    if (kind == PT_OBJ_INSTANCE) { const DInstance& ins = S.instances[index]; r = instance_local_ray(ins, r); mesh = ins.child_index; cur_inst = index; }
                    cur_tie = rf.b;
                    br = make_boxray(r);
                    tmax_f = __double2float_ru(c.t);
                    stack[0] = make_uint2(S.meshes[mesh].root_entry, 0u); sp = 1;
                    c.n_wide = 0; c.n_refs = 0; c.n_prims = 0;
                    active = true;
                } else if (last) { done_cls = c.ref == kNone ? (uint32_t)CLS_MISS : class_of_kind(S.materials[hit_material(S, c.ref)].kind); done_i = i; }
            }
        }
        if (!__any_sync(0xFFFFFFFFu, active)) {
            if (!drained) continue;
            if (__any_sync(0xFFFFFFFFu, done_cls != N_CLS)) queue_append(q, done_cls, done_i);
            break;
        }
        // ---- a burst of while-while rounds
#pragma unroll 1
        for (int s = 0; s < kBlasBurst; s++) {
            if (active && blas_round<COUNT>(S, r, br, t_min, tmin_f, tmax_f, stack, sp, c, cur_inst, cur_tie)) {
                active = false;
                const bool is_light = c.ref != kNone && !(c.tie_outer >> 31);
                HitRec h; h.t = c.t; h.ref = c.ref; h.inst_light = (c.inst & 0x7FFFFFFFu) | (is_light ? 0x80000000u : 0u);
                hits[i] = h;
                if (!last) ties[i] = make_uint2(c.tie_outer, c.tie_inner);
                else { done_cls = c.ref == kNone ? (uint32_t)CLS_MISS : class_of_kind(S.materials[hit_material(S, c.ref)].kind); done_i = i; }
                if (COUNT) { w0 += 2 * c.n_wide; w1 += c.n_refs; w2 += c.n_prims; }
            }
        }
    }
    if (COUNT) {
        w0 = __reduce_add_sync(0xFFFFFFFFu, w0); w1 = __reduce_add_sync(0xFFFFFFFFu, w1); w2 = __reduce_add_sync(0xFFFFFFFFu, w2);
        if (lane == 0 && (w0 | w1 | w2)) { atomicAdd(work, (unsigned long long)w0); atomicAdd(work + 1, (unsigned long long)w1); atomicAdd(work + 2, (unsigned long long)w2); }
    }
}

PT_D void add_radiance(float* __restrict__ accum, uint32_t pix, d3 v, uint32_t nan_policy, unsigned long long* nonfinite, bool& dead) {
    if (!finite3(v)) {
        atomicAdd(nonfinite, 1ull);
        if (nan_policy == PT_NAN_DROP) { dead = true; return; }  // drop the contribution and end the path
    }
    if (v.x != 0.0) atomicAdd(accum + 3ull * pix, (float)v.x);
    if (v.y != 0.0) atomicAdd(accum + 3ull * pix + 1, (float)v.y);
    if (v.z != 0.0) atomicAdd(accum + 3ull * pix + 2, (float)v.z);
}

template <int CLS> struct ClassKind { static constexpr int value = -1; };
template <> struct ClassKind<CLS_LIGHT> { static constexpr int value = PT_MAT_LIGHT; };
template <> struct ClassKind<CLS_DIFFUSE> { static constexpr int value = PT_MAT_DIFFUSE; };
template <> struct ClassKind<CLS_METAL> { static constexpr int value = PT_MAT_METAL; };
template <> struct ClassKind<CLS_GLASS> { static constexpr int value = PT_MAT_GLASS; };
template <> struct ClassKind<CLS_PRINCIPLED> { static constexpr int value = PT_MAT_PRINCIPLED; };

// One loop iteration of Camera::trace after intersect_all (camera.rs:180-225) for the paths of ONE shade class.
// Grid-stride over the class queue; survivors are written compacted into `out` (ballot + block prefix + one atomic).
// VAR bit 0: environment importance sampling joins the mixture; bit 1: World.lights holds more than quads and spheres
// (cuboid / mesh / instance lights).  The reference's shipped scenes need neither, and their kernels carry none of that code.
template <int CLS, int VAR = 0>
__global__ void __launch_bounds__(kBlock, PT_SHADE_MIN_BLOCKS) k_shade(PathBuf in, Queues q, const HitRec* __restrict__ hits, PathBuf out,
                                                    uint32_t* __restrict__ out_count, float* __restrict__ accum,
                                                    unsigned long long* __restrict__ nonfinite, DScene S, DCameraEx cam, RenderConst rc) {
    constexpr int K = ClassKind<CLS>::value;
    const uint32_t count = q.count[CLS];
    const uint32_t* __restrict__ items = q.items + (size_t)CLS * q.stride;
    __shared__ uint32_t bin_count[8][kBlock / 32];  // survivors per (octant bin, warp), then their exclusive prefix
    __shared__ uint32_t block_base;
    for (uint32_t base = blockIdx.x * kBlock; base < count; base += gridDim.x * kBlock) {
        const uint32_t j = base + threadIdx.x;
        bool alive = false;
        RayD next; d3 thr = mk(0, 0, 0); uint4 ids = make_uint4(0, 0, 0, 0);
        if (j < count) {
            const uint32_t i = items[j];
            RayD ray = load_ray(in, i);
            thr = mk(in.f[7][i], in.f[8][i], in.f[9][i]);
            ids = in.ids[i];
            const uint32_t pix = ids.x, bounces = ids.z >> 16;
            bool dead = false;
            if (CLS == CLS_MISS) {  // camera.rs:180-183
                add_radiance(accum, pix, thr * sample_environment(S, cam.c, ray.d), rc.nan_policy, nonfinite, dead);
            } else {
                const HitRec hr = hits[i];
                Rng rng; rng.init(rc.seed, pix, ids.y, ids.z & 0xFFFFu);
                HitInfoD h;
                reconstruct_hit<false>(S, ray, hr.ref, hr.inst_light & 0x7FFFFFFFu, hr.t, h);
                const DMaterial& m = S.materials[h.material];
                // camera.rs:186-187: `radiance += throughput * emitted` runs for every hit; for non-emitters it only
                // matters when the throughput is already inf/NaN (inf * 0 = NaN poisons the pixel, Q32).
                if (CLS == CLS_LIGHT || !finite3(thr)) {
                    d3 em = CLS == CLS_LIGHT ? texture_value(S, m.base_color_tex, h.u, h.v, h.point) : mk(0, 0, 0);
                    add_radiance(accum, pix, thr * em, rc.nan_policy, nonfinite, dead);
                }
                bool go = !dead;
                if (go && bounces > 5) {  // Russian roulette, camera.rs:190-196
                    double p = clampd(luminance(thr), 0.01, 1.0);
                    if (rng.next() > p) go = false;
                    else thr = thr / p;
                }
                if (go) {
                    // camera.rs:199-200: p_light = 0.5 iff lights exist.  With PT_RENDER_ENV_IMPORTANCE (ours) the environment map
                    // joins the mixture as a third sampler: p_bsdf = 0.5, the other half is split between lights and environment.
                    constexpr bool env_is = (VAR & 1) != 0, GEN = (VAR & 2) != 0;
                    const double p_env = env_is ? (S.n_lights == 0 ? 0.5 : 0.25) : 0.0;
                    const double p_light = S.n_lights == 0 ? 0.0 : 0.5 - p_env, p_bsdf = env_is ? 0.5 : 1.0 - p_light;
                    const double rsel = rng.next();
                    d3 dir;
                    bool ok;
                    if (rsel < p_light) ok = lights_sample<GEN>(S, h.point, ray.time, rng, dir);
                    else if (env_is && rsel < p_light + p_env) { const double u1 = rng.next(), u2 = rng.next(); dir = env_sample(rc.env, u1, u2); ok = true; }
                    else ok = bsdf_sample<K>(S, h.material, ray.d, h, rng, dir);
                    if (ok) {  // camera.rs:212-225
                        d3 f; double bsdf_pdf;
                        bsdf_eval_pdf<K>(S, h.material, -ray.d, dir, h, f, bsdf_pdf);
                        double light_pdf = lights_pdf<GEN>(S, h.point, dir, ray.time);
                        double pdf = p_bsdf * bsdf_pdf + p_light * light_pdf;
                        if (env_is) pdf = pdf + p_env * env_pdf(rc.env, dir);
                        d3 attenuation = f / pdf;
                        double e = 1e-3 * signum(dot(dir, h.gn));
                        next = make_ray(h.point + e * h.gn, dir, ray.time);
                        thr = thr * attenuation;
                        alive = bounces + 1 < cam.c.max_depth;  // `for bounces in 0..max_depth`, camera.rs:177
                        if (rc.nan_policy == PT_NAN_DROP && !finite3(thr)) { atomicAdd(nonfinite, 1ull); alive = false; }
                        ids.z = (rng.used & 0xFFFFu) | ((bounces + 1) << 16);
                    }
                }
            }
        }
        if (CLS == CLS_MISS) continue;  // a miss ends the path: nothing to compact
        // ---- compaction: survivors grouped by direction octant within the block (warps of the next trace launch then hold
        //      rays that walk the BVH in the same order and tend to cost the same), one atomic per block
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t key = alive ? (((next.d.y < 0.0) | ((next.d.x < 0.0) << 1) | ((next.d.z < 0.0) << 2)) & rc.sort_mask) : 8u;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
        if (threadIdx.x < 8 * (kBlock / 32)) (&bin_count[0][0])[threadIdx.x] = 0;
        __syncthreads();
        if (alive && lane == (uint32_t)(__ffs(peers) - 1)) bin_count[key][warp] = __popc(peers);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int b = 0; b < 8; b++)
#pragma unroll
                for (int w = 0; w < kBlock / 32; w++) { uint32_t c = bin_count[b][w]; bin_count[b][w] = total; total += c; }
            block_base = total ? atomicAdd(out_count, total) : 0;
        }
        __syncthreads();
        if (alive) {
            uint32_t dst = block_base + bin_count[key][warp] + __popc(peers & ((1u << lane) - 1u));
            store_path(out, dst, next, thr, ids);
        }
        __syncthreads();  // bin_count / block_base are reused by the next grid-stride iteration
    }
}

// ---- PT_RENDER_NEE (ours): next-event estimation with MIS instead of the reference's one-sample mixture.
// At every non-emissive hit: (1) a light direction from World.lights.sample spawns a SHADOW PATH — an ordinary pool entry
// marked kShadowMark that carries W = throughput * f / (pdf_light + pdf_bsdf) (balance heuristic folded in); it is traced
// by k_trace like any ray and only ever adds W * emitted if its closest hit is an emitter; (2) the path continues by BSDF
// sampling alone and remembers w = pdf_bsdf / (pdf_bsdf + pdf_light(dir)) (as fp32 in ids.w) to weight the emission it may
// run into next.  Environment hits keep weight 1 (the environment is not a NEE light).  Two outputs per input at most:
// out_count[0] counts all outputs, out_count[1] the non-shadow ones (the host keeps those <= pool / 2).
constexpr uint32_t kShadowMark = 0xFFFFFFFFu;
#ifndef PT_NEE_MIN_BLOCKS
#define PT_NEE_MIN_BLOCKS 4  // 128 registers; 3 blocks (168 registers, far fewer spills) measured 2-10 % slower
#endif
template <int CLS>
__global__ void __launch_bounds__(kBlock, PT_NEE_MIN_BLOCKS) k_shade_nee(PathBuf in, Queues q, const HitRec* __restrict__ hits, PathBuf out,
                                                        uint32_t* __restrict__ out_count, float* __restrict__ accum,
                                                        unsigned long long* __restrict__ nonfinite, DScene S, DCameraEx cam, RenderConst rc) {
    constexpr int K = ClassKind<CLS>::value;
    const uint32_t count = q.count[CLS];
    const uint32_t* __restrict__ items = q.items + (size_t)CLS * q.stride;
    __shared__ uint32_t warp_count[kBlock / 32], warp_alive[kBlock / 32];
    __shared__ uint32_t block_base;
    for (uint32_t base = blockIdx.x * kBlock; base < count; base += gridDim.x * kBlock) {
        const uint32_t j = base + threadIdx.x;
        bool alive = false, shadow = false;
        RayD next, sray; d3 thr = mk(0, 0, 0), sthr = mk(0, 0, 0); uint4 ids = make_uint4(0, 0, 0, 0);
        if (j < count) {
            const uint32_t i = items[j];
            RayD ray = load_ray(in, i);
            thr = mk(in.f[7][i], in.f[8][i], in.f[9][i]);
            ids = in.ids[i];
            const uint32_t pix = ids.x, bounces = ids.z >> 16;
            const bool is_shadow = ids.w == kShadowMark;
            bool dead = false;
            if (CLS == CLS_MISS) {
                if (!is_shadow) add_radiance(accum, pix, thr * sample_environment(S, cam.c, ray.d), rc.nan_policy, nonfinite, dead);
            } else {
                const HitRec hr = hits[i];
                HitInfoD h;
                reconstruct_hit<false>(S, ray, hr.ref, hr.inst_light & 0x7FFFFFFFu, hr.t, h);
                const DMaterial& m = S.materials[h.material];
                if (CLS == CLS_LIGHT) {  // emitters end the path (DiffuseLight::sample -> None); MIS weight of the strategy that got here
                    const double w = (is_shadow || bounces == 0) ? 1.0 : (double)__uint_as_float(ids.w);
                    add_radiance(accum, pix, thr * texture_value(S, m.base_color_tex, h.u, h.v, h.point) * w, rc.nan_policy, nonfinite, dead);
                } else if (!is_shadow) {
                    Rng rng; rng.init(rc.seed, pix, ids.y, ids.z & 0xFFFFu);
                    if (!finite3(thr)) add_radiance(accum, pix, thr * 0.0, rc.nan_policy, nonfinite, dead);  // poisons like camera.rs:186-187 (Q32)
                    bool go = !dead;
                    if (go && bounces > 5) {  // Russian roulette, camera.rs:190-196
                        double p = clampd(luminance(thr), 0.01, 1.0);
                        if (rng.next() > p) go = false;
                        else thr = thr / p;
                    }
                    if (go) {
                        const bool deeper = bounces + 1 < cam.c.max_depth;
                        d3 dir;
                        if (S.n_lights != 0 && lights_sample<true>(S, h.point, ray.time, rng, dir)) {
                            const double pl = lights_pdf<true>(S, h.point, dir, ray.time);
                            d3 fl; double pb;
                            bsdf_eval_pdf<K>(S, h.material, -ray.d, dir, h, fl, pb);
                            const d3 W = thr * (fl / (pl + pb));
                            if (deeper && pl > 0.0 && (W.x != 0.0 || W.y != 0.0 || W.z != 0.0)) {
                                shadow = true; sthr = W;
                                sray = make_ray(h.point + (1e-3 * signum(dot(dir, h.gn))) * h.gn, dir, ray.time);
                            }
                        }
                        if (bsdf_sample<K>(S, h.material, ray.d, h, rng, dir)) {
                            d3 f; double pb;
                            bsdf_eval_pdf<K>(S, h.material, -ray.d, dir, h, f, pb);
                            const double pl = S.n_lights != 0 ? lights_pdf<true>(S, h.point, dir, ray.time) : 0.0;
                            const float w = pl > 0.0 ? (float)(pb / (pb + pl)) : 1.0f;
                            next = make_ray(h.point + (1e-3 * signum(dot(dir, h.gn))) * h.gn, dir, ray.time);
                            thr = thr * (f / pb);
                            alive = deeper;
                            if (rc.nan_policy == PT_NAN_DROP && !finite3(thr)) { atomicAdd(nonfinite, 1ull); alive = false; }
                            ids.w = __float_as_uint(w);
                        }
                        ids.z = (rng.used & 0xFFFFu) | ((bounces + 1) << 16);
                    }
                }
            }
        }
        if (CLS == CLS_MISS || CLS == CLS_LIGHT) continue;  // nothing survives a miss or an emitter
        // ---- compaction of up to two outputs per lane
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t b_alive = __ballot_sync(0xFFFFFFFFu, alive), b_shadow = __ballot_sync(0xFFFFFFFFu, shadow);
        if (lane == 0) { warp_count[warp] = __popc(b_alive) + __popc(b_shadow); warp_alive[warp] = __popc(b_alive); }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = 0, total_alive = 0;
#pragma unroll
            for (int w = 0; w < kBlock / 32; w++) { uint32_t c = warp_count[w]; warp_count[w] = total; total += c; total_alive += warp_alive[w]; }
            block_base = total ? atomicAdd(out_count, total) : 0;
            if (total_alive) atomicAdd(out_count + 1, total_alive);
        }
        __syncthreads();
        const uint32_t below = (1u << lane) - 1u;
        const uint32_t wbase = block_base + warp_count[warp];
        if (alive) store_path(out, wbase + __popc(b_alive & below), next, thr, ids);
        if (shadow) store_path(out, wbase + __popc(b_alive) + __popc(b_shadow & below), sray, sthr, make_uint4(ids.x, ids.y, ids.z, kShadowMark));
        __syncthreads();
    }
}

// sqrt-gamma + 8-bit quantisation of camera.rs:109-114,128-130 on `scale * accum`
__global__ void k_tonemap(const float* __restrict__ accum, double scale, uint32_t n_values, uint8_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_values) return;
    double x = (double)accum[i] * scale;
    double g = sqrt(fmax(x, 0.0));
    double v = clampd(g, 0.0, 0.999) * 256.0;
    out[i] = (v != v) ? 0 : (uint8_t)v;
}
__global__ void k_scale(const float* __restrict__ accum, float scale, uint32_t n_values, float* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_values) out[i] = accum[i] * scale;
}

// ---------------------------------------------------------------- parity / test entry kernels
PT_D pt_vec3 to_abi(d3 v) { pt_vec3 r; r.x = v.x; r.y = v.y; r.z = v.z; return r; }
PT_D d3 from_abi(pt_vec3 v) { return mk(v.x, v.y, v.z); }

// keyed uniforms of ray batches (pt_volume): seed 0, pixel = ray index, sample = bounce = 0
struct BatchVol { static constexpr bool kEnabled = true; uint32_t i; PT_D double operator()(uint32_t v) const { return keyed_uniform(0, i, 0, 0, v); } };
template <bool WIDE>
__global__ void k_trace_batch(const pt_ray* __restrict__ rays, size_t n, double t_min, pt_hit* __restrict__ out, DScene S) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayD r; r.o = from_abi(rays[i].origin); r.d = from_abi(rays[i].direction); r.time = rays[i].time;
    Closest c;
    pt_hit o; memset(&o, 0, sizeof(o)); o.instance = PT_NONE;
    const bool hit = trace_closest<false, true, WIDE>(S, [&]() { return r; }, t_min, 0.0, c, BatchVol{(uint32_t)i});
    o.work = min(c.n_pairs + c.n_wide, 0xFFFFu) | (min(c.n_prims, 0xFFFFu) << 16);
    if (hit) {
        HitInfoD h;
        reconstruct_hit(S, r, c.ref, c.inst, c.t, h);
        o.hit = 1; o.t = c.t; o.u = h.u; o.v = h.v; o.point = to_abi(h.point); o.geometric_normal = to_abi(h.gn); o.shading_normal = to_abi(h.sn);
        o.prim_kind = ref_kind(c.ref); o.prim_index = ref_index(c.ref); o.instance = c.inst == kInstNone ? PT_NONE : c.inst;
        o.material = h.material; o.front_face = h.front_face; o.is_light = c.is_light;
    }
    out[i] = o;
}
template <bool WIDE>
__global__ void k_trace_any_batch(const pt_ray* __restrict__ rays, size_t n, double t_min, const double* __restrict__ t_max,
                                  uint8_t* __restrict__ out, DScene S) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayD r; r.o = from_abi(rays[i].origin); r.d = from_abi(rays[i].direction); r.time = rays[i].time;
    Closest c;
    out[i] = trace_closest<true, false, WIDE>(S, [&]() { return r; }, t_min, t_max[i], c, BatchVol{(uint32_t)i}) ? 1 : 0;
}
// pt_trace_closest_wavefront: a host ray batch goes through the RENDER's traversal stage (launch_trace in api.cu: the same
// kernels, grids and queues as a wavefront iteration).  These two kernels are only the adapters around it: rays into the
// SoA path pool (pixel = ray index keys the media uniforms exactly as BatchVol does with seed 0), hit records back out
// through the shade stage's own reconstruct_hit.
__global__ void k_rays_to_pool(const pt_ray* __restrict__ rays, uint32_t n, uint32_t first, PathBuf out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayD r; r.o = from_abi(rays[i].origin); r.d = from_abi(rays[i].direction); r.time = rays[i].time;
    store_path(out, i, r, mk(1, 1, 1), make_uint4(first + i, 0u, 0u, 0u));
}
__global__ void k_hits_to_abi(const pt_ray* __restrict__ rays, uint32_t n, const HitRec* __restrict__ hits, pt_hit* __restrict__ out, DScene S) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayD r; r.o = from_abi(rays[i].origin); r.d = from_abi(rays[i].direction); r.time = rays[i].time;
    const HitRec hr = hits[i];
    pt_hit o; memset(&o, 0, sizeof(o)); o.instance = PT_NONE;
    if (hr.ref != kNone) {
        const uint32_t inst = hr.inst_light & 0x7FFFFFFFu;
        HitInfoD h;
        reconstruct_hit(S, r, hr.ref, inst, hr.t, h);
        o.hit = 1; o.t = hr.t; o.u = h.u; o.v = h.v; o.point = to_abi(h.point); o.geometric_normal = to_abi(h.gn); o.shading_normal = to_abi(h.sn);
        o.prim_kind = ref_kind(hr.ref); o.prim_index = ref_index(hr.ref); o.instance = inst == kInstNone ? PT_NONE : inst;
        o.material = h.material; o.front_face = h.front_face; o.is_light = hr.inst_light >> 31;
    }
    out[i] = o;
}
PT_D HitInfoD info_from_query(const pt_bsdf_query& q, uint32_t material) {
    HitInfoD h; h.point = from_abi(q.point); h.gn = from_abi(q.geometric_normal); h.sn = from_abi(q.shading_normal);
    h.t = 0; h.u = q.u; h.v = q.v; h.front_face = q.front_face != 0; h.material = material;
    return h;
}
__global__ void k_bsdf_eval(uint32_t material, size_t n, const pt_bsdf_query* __restrict__ q, pt_bsdf_result* __restrict__ out, DScene S) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    HitInfoD h = info_from_query(q[i], material);
    d3 f; double pdf;
    bsdf_eval_pdf(S, material, from_abi(q[i].view_dir), from_abi(q[i].light_dir), h, f, pdf);
    pt_bsdf_result r; r.eval = to_abi(f); r.pdf = pdf; r.emitted = to_abi(bsdf_emitted(S, material, h.u, h.v, h.point)); r._pad = 0;
    out[i] = r;
}
__global__ void k_bsdf_sample(uint32_t material, size_t n, const pt_bsdf_query* __restrict__ q, const double* __restrict__ uniforms8,
                              pt_bsdf_sample_result* __restrict__ out, DScene S) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    HitInfoD h = info_from_query(q[i], material);
    Rng rng; rng.init_array(uniforms8 + 8 * i, 8);
    d3 dir = mk(0, 0, 0);
    bool ok = bsdf_sample(S, material, -from_abi(q[i].view_dir), h, rng, dir);
    pt_bsdf_sample_result r; r.dir = ok ? to_abi(dir) : to_abi(mk(0, 0, 0)); r.valid = ok; r.n_uniforms = rng.used;
    out[i] = r;
}
__global__ void k_camera_rays(DCameraEx cam, uint64_t seed, size_t n, const uint32_t* __restrict__ row, const uint32_t* __restrict__ col,
                              const uint32_t* __restrict__ sample, pt_ray* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Rng rng; rng.init(seed, row[i] * cam.c.width + col[i], sample[i], 0);
    RayD r = generate_ray(cam, row[i], col[i], rng);
    pt_ray o; o.origin = to_abi(r.o); o.direction = to_abi(r.d); o.time = r.time;
    out[i] = o;
}
__global__ void k_lights(size_t n, const pt_vec3* __restrict__ origin, const double* __restrict__ time, const double* __restrict__ uniforms4,
                         pt_vec3* __restrict__ dir, uint32_t* __restrict__ valid, double* __restrict__ pdf, DScene S) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Rng rng; rng.init_array(uniforms4 + 4 * i, 4);
    d3 d = mk(0, 0, 0);
    bool ok = lights_sample(S, from_abi(origin[i]), time[i], rng, d);
    valid[i] = ok; dir[i] = to_abi(ok ? d : mk(0, 0, 0));
    pdf[i] = ok ? lights_pdf(S, from_abi(origin[i]), d, time[i]) : 0.0;
}

// ---- exact SAH sweep of bvh.rs:54-120 for one node (the step before the path).  Thread (axis, k) folds all
// n items in list order into a left / right box exactly as BVH::evaluate_sah does (AABB::union pads 1e-3 on EVERY union,
// aabb.rs:16-25, so the fold order is part of the result) and prices the split at item k's centroid.  O(n^2) work like the
// reference, but 3n folds run at once; items are staged through shared memory, every thread reads the same item (broadcast).
struct SahBox { double lo[3], hi[3]; };
PT_D void sah_merge(SahBox& b, const SahBox& o) {  // AABB::union -> AABB::new(min, max) with its padding
    for (int a = 0; a < 3; a++) {
        const double mn = fmin(b.lo[a], o.lo[a]), mx = fmax(b.hi[a], o.hi[a]);
        b.lo[a] = fmin(mn, mx) - 1e-3; b.hi[a] = fmax(mn, mx) + 1e-3;
    }
}
PT_D double sah_half_area(const SahBox& b) {
    const double ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
    return ex * ey + ex * ez + ey * ez;
}
constexpr int kSahTile = 128;
__global__ void __launch_bounds__(kSahTile) k_sah_sweep(uint32_t n, const SahBox* __restrict__ boxes, SahBox parent, double* __restrict__ cost) {
    __shared__ SahBox tile[kSahTile];
    const uint32_t t = blockIdx.x * kSahTile + threadIdx.x;
    const bool active = t < 3u * n;
    const uint32_t axis = active ? t / n : 0u, k = active ? t % n : 0u;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const double split = 0.5 * (boxes[k].lo[axis] + boxes[k].hi[axis]);  // AABB::centroid
    SahBox lb, rb;
    for (int a = 0; a < 3; a++) { lb.lo[a] = rb.lo[a] = inf; lb.hi[a] = rb.hi[a] = -inf; }
    uint32_t lc = 0, rc = 0;
    for (uint32_t base = 0; base < n; base += kSahTile) {
        const uint32_t m = min((uint32_t)kSahTile, n - base);
        __syncthreads();
        if (threadIdx.x < m) tile[threadIdx.x] = boxes[base + threadIdx.x];
        __syncthreads();
        for (uint32_t i = 0; i < m; i++) {
            const SahBox& o = tile[i];
            if (0.5 * (o.lo[axis] + o.hi[axis]) < split) { sah_merge(lb, o); lc++; } else { sah_merge(rb, o); rc++; }
        }
    }
    if (!active) return;
    double c = inf;
    if (lc != 0 && rc != 0) {
        const double v = sah_half_area(lb) * (double)lc + sah_half_area(rb) * (double)rc;
        const double parent_cost = sah_half_area(parent) * (double)n;
        if (v > 0.0 && v < parent_cost) c = v;
    }
    cost[t] = c;
}

__global__ void k_env(size_t n, const double* __restrict__ uniforms2, pt_vec3* __restrict__ dir, double* __restrict__ pdf, DEnvDist E) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const d3 d = env_sample(E, uniforms2[2 * i], uniforms2[2 * i + 1]);
    dir[i] = to_abi(d);
    pdf[i] = env_pdf(E, d);
}

}  // namespace ptd
