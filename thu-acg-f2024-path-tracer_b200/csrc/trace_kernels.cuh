// Traversal-stage kernels (World::intersect_all, world.rs:47-62) — compiled by trace.cu only.
#pragma once
#include "wavefront.cuh"

namespace ptd {

// Diagnostics of the COUNT kernel variants (pt_ctx_set_profiling(2) + pt_debug_histograms): eight histograms of 64 bins.
// 0-2: top-level pass, per ray: node fetches, reference boxes, f64 primitive tests; 3: mesh visits queued per ray;
// 4-6: mesh rounds, per visit: wide nodes, reference boxes, f64 triangle tests; 7: [0] visits dropped at entry (the mesh
// had fallen behind the closest hit), [1] visits walked, [2] visits that improved the hit.
__device__ unsigned long long g_hist[8 * 64];
PT_D void hist_add(int h, uint32_t v) { atomicAdd(&g_hist[h * 64 + min(v, 63u)], 1ull); }

// World::intersect_all for every live path (one ray per thread), then the path joins the queue of its shade class.
// COUNT (profiling mode only): also sums the traversal work of all rays into work[3] = {node pairs, reference boxes,
// f64 primitive tests}, from which bench.py derives the bytes the device actually requests per ray.
// VOL: the scene holds constant-density media; their free-flight uniforms are keyed by (seed, pixel, sample, bounce).
struct PathVol {
    static constexpr bool kEnabled = true;
    uint64_t seed; PathBuf in; uint32_t i;
    PT_D double operator()(uint32_t v) const {
        return keyed_uniform(seed, in.ray[i].pixel, in.ray[i].sample, in.state[i].rng_bounce >> 16, v);
    }
};
// Two-pass traversal for scenes with mesh BLASes (DEFER): a warp of k_trace mixes rays that leave the top-level BVH after a
// node or two with rays that walk a mesh for ten times as long, so the short ones idle (8.9 of 32 lanes active on scene 6).
// With DEFER, k_trace walks the top level only — simple primitives are tested, every mesh whose box the ray enters is
// queued (up to kDeferMax per ray; a further one is walked inline) — and k_trace_blas<round> then walks the r-th queued
// mesh of each such ray, compacted so that its warps hold only rays inside a BLAS.  Hit record and tie ranks travel through
// `hits` / `ties`; a ray joins its shade-class queue after its last round.  Same closest hit, same tie rules (ranks).
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
        if constexpr (VOL) trace_closest<false, COUNT, WIDE>(S, [&]() { return load_ray(in, i); }, t_min, 0.0, c, PathVol{seed, in, i});
        else if constexpr (DEFER) trace_closest<false, COUNT, WIDE>(S, [&]() { return load_ray(in, i); }, t_min, 0.0, c, NoVol(), DeferList{&nd, dslot, dt});
        else trace_closest<false, COUNT, WIDE>(S, [&]() { return load_ray(in, i); }, t_min, 0.0, c);
        if (COUNT) {
            w0 = c.n_pairs + 2 * c.n_wide; w1 = c.n_refs; w2 = c.n_prims;  // in 64-byte units
            hist_add(0, c.n_pairs + c.n_wide); hist_add(1, c.n_refs); hist_add(2, c.n_prims); if (DEFER) hist_add(3, nd);
        }
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
                const double t_before = c.t;
                trace_blas<COUNT>(S, S.meshes[mesh].root_entry, r, t_min, c, inst, rf.b);
                if (COUNT) {
                    w0 += c.n_pairs + 2 * c.n_wide; w1 += c.n_refs; w2 += c.n_prims;
                    hist_add(4, c.n_wide); hist_add(5, c.n_refs); hist_add(6, c.n_prims); hist_add(7, 1); if (c.t < t_before) hist_add(7, 2);
                }
                HitRec h; h.t = c.t; h.ref = c.ref; h.inst_light = (c.inst & 0x7FFFFFFFu) | (c.is_light ? 0x80000000u : 0u);
                hits[i] = h;
                if (!(e.y >> 31)) ties[i] = make_uint2(c.tie_outer, c.tie_inner);
            } else if (COUNT) hist_add(7, 0);
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
    double t_entry = 0.0;
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
                if (COUNT) t_entry = c.t;
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
                } else {
                    if (COUNT) hist_add(7, 0);
                    if (last) { done_cls = c.ref == kNone ? (uint32_t)CLS_MISS : class_of_kind(S.materials[hit_material(S, c.ref)].kind); done_i = i; }
                }
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
                if (COUNT) {
                    w0 += 2 * c.n_wide; w1 += c.n_refs; w2 += c.n_prims;
                    hist_add(4, c.n_wide); hist_add(5, c.n_refs); hist_add(6, c.n_prims); hist_add(7, 1); if (c.t < t_entry) hist_add(7, 2);
                }
            }
        }
    }
    if (COUNT) {
        w0 = __reduce_add_sync(0xFFFFFFFFu, w0); w1 = __reduce_add_sync(0xFFFFFFFFu, w1); w2 = __reduce_add_sync(0xFFFFFFFFu, w2);
        if (lane == 0 && (w0 | w1 | w2)) { atomicAdd(work, (unsigned long long)w0); atomicAdd(work + 1, (unsigned long long)w1); atomicAdd(work + 2, (unsigned long long)w2); }
    }
}

// ================================================================== flat top level + mesh rounds (the production traversal of
// every scene whose World is small: all of the reference's scenes but the 480-sphere one)
//
//   k_top         one ray per lane over the top-level reference LIST (wavefront.cuh: TopList).  PRIMARY: the ray is a camera ray
//                 generated in registers (camera.rs:153-168) and its path record is written here, once — there is no separate
//                 ray-generation kernel and no re-read.  Simple primitives are tested in f64 on the spot; a mesh whose world
//                 box the ray enters is only noted in a bit mask and queued for the rounds below.
//   k_mesh_enter  round r, one queued visit per lane, fully convergent: the ray goes into the mesh's space (instance.rs:36-38),
//                 the mesh's root node is tested, and only rays that enter a child get a 128-byte walk record.  (Measured
//                 before the split: 37 % of the visits ended at the root node, yet paid the gathers and the f64 transform
//                 inside the divergent walk kernel.)
//   k_mesh_walk   persistent warps with lane refill from the walk records.  A lane holds at most one node to open (`cur`) and
//                 one triangle to test (`pend`); the warp votes on which of the two kinds of work runs next, so a 4-wide box
//                 step or an f64 Moeller-Trumbore test is issued for many lanes at once instead of each lane switching kinds
//                 on its own.
// Same closest hit as World::intersect_all: minimum t, exact ties by the precomputed ranks, so neither
// the visit order nor the split into passes can change a result.
struct __align__(16) WalkRec {
    double o[3], d[3];                       // ray in the mesh's space
    double t;                                // closest hit so far (provisional hit of the top-level pass and earlier rounds)
    uint32_t path, flags;                    // flags: last visit of this ray << 31 | shade class of the mesh << 12 | of the provisional hit << 8
    uint32_t ref, inst_light, tie_o, tie_i;  // provisional hit
    uint32_t cur_inst, cur_tie, n_init, pad;
    uint2 init[4];                           // children of the root node the ray enters, nearest first: (child, entry distance)
};
static_assert(sizeof(WalkRec) == 16 * kWalkRecU4, "walk record size");

// 4-wide box step: entry distances of the four children, ascending, +inf = not entered
PT_D void wide2_step(const DWide2& w, const BoxRay& br, float tmin_f, float tmax_f, uint32_t e[4], float t[4]) {
    const float4 lx = *reinterpret_cast<const float4*>(w.lo[0]), ly = *reinterpret_cast<const float4*>(w.lo[1]),
                 lz = *reinterpret_cast<const float4*>(w.lo[2]), hx = *reinterpret_cast<const float4*>(w.hi[0]),
                 hy = *reinterpret_cast<const float4*>(w.hi[1]), hz = *reinterpret_cast<const float4*>(w.hi[2]);
    const uint4 ch = *reinterpret_cast<const uint4*>(w.child);
    const float kInf = __int_as_float(0x7f800000);
    const bool sx = br.ix < 0.f, sy = br.iy < 0.f, sz = br.iz < 0.f;
#define PT_SLAB4(I, K)                                                                                                   \
    {                                                                                                                    \
        const float x0 = __fmaf_rn(sx ? hx.I : lx.I, br.ix, br.nx), x1 = __fmaf_rn(sx ? lx.I : hx.I, br.ix, br.fx);     \
        const float y0 = __fmaf_rn(sy ? hy.I : ly.I, br.iy, br.ny), y1 = __fmaf_rn(sy ? ly.I : hy.I, br.iy, br.fy);     \
        const float z0 = __fmaf_rn(sz ? hz.I : lz.I, br.iz, br.nz), z1 = __fmaf_rn(sz ? lz.I : hz.I, br.iz, br.fz);     \
        const float tn = fmaxf(fmaxf(x0, y0), fmaxf(z0, tmin_f)), tf = fminf(fminf(x1, y1), fminf(z1, tmax_f));         \
        t[K] = tn <= tf ? tn : kInf; /* NaN planes (0*inf) drop out of fmaxf/fminf: conservative */                      \
    }
    PT_SLAB4(x, 0) PT_SLAB4(y, 1) PT_SLAB4(z, 2) PT_SLAB4(w, 3)
#undef PT_SLAB4
    e[0] = ch.x; e[1] = ch.y; e[2] = ch.z; e[3] = ch.w;
#define PT_CSWAP(A, B) { const bool s_ = t[B] < t[A]; const float tt = s_ ? t[A] : t[B]; t[A] = s_ ? t[B] : t[A]; t[B] = tt; \
                         const uint32_t ee = s_ ? e[A] : e[B]; e[A] = s_ ? e[B] : e[A]; e[B] = ee; }
    PT_CSWAP(0, 1) PT_CSWAP(2, 3) PT_CSWAP(0, 2) PT_CSWAP(1, 3) PT_CSWAP(1, 2)
#undef PT_CSWAP
}

// (wide2_step_addr — the same step with the near / far planes picked by address — lives in geom.cuh: the tail megakernel uses it too)
#ifndef PT_TOP_MIN_BLOCKS
#define PT_TOP_MIN_BLOCKS 5
#endif
template <bool PRIMARY, bool COUNT>
__global__ void __launch_bounds__(kBlock, PT_TOP_MIN_BLOCKS) k_top(PathBuf pool, uint32_t slot0, uint32_t n, HitRec* __restrict__ hits, Queues q, DScene S,
                                                                    TopList top, MeshQueues mq, uint2* __restrict__ ties, double t_min, GenArgs gen,
                                                                    const uint32_t* __restrict__ n_dev, unsigned long long* __restrict__ work) {
    if (!PRIMARY && n_dev) n = min(n, *n_dev);  // batched tail iterations: the survivors counter of the iteration before
    const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
    const uint32_t i = slot0 + j;
    uint32_t cls = N_CLS, mesh_mask = 0, prov_cls = CLS_MISS;
    uint32_t w1 = 0, w2 = 0;
    if (j < n) {
        RayD r;
        if (PRIMARY) {
            uint32_t s_local, pix, row, col;
            path_pixel(gen.g0 + j, gen.n_pixels, gen.cam.c, gen.rc, pix, s_local, row, col);
            const uint32_t sample = gen.rc.sample_begin + s_local * gen.rc.sample_stride;
            Rng rng; rng.init(gen.rc.seed, pix, sample, 0);
            r = generate_ray(gen.cam, row, col, rng);
            PT_ASSERT(rng.used == kFreshDraws);
            if (gen.rc.fresh_from != 0xFFFFFFFFu) store_ray(pool, i, r, pix, sample);  // the state of a fresh path is implicit (RenderConst::fresh_from)
            else store_path(pool, i, r, mk(1, 1, 1), make_uint4(pix, sample, rng.used, 0));
        } else r = load_ray(pool, i);
        const BoxRay br = make_boxray(r);
        const float tmin_f = __double2float_rd(t_min);
        float tmax_f = __int_as_float(0x7f800000);
        Closest c;
        c.t = __longlong_as_double(0x7ff0000000000000ll); c.ref = kNone; c.inst = kInstNone; c.tie_outer = 0; c.tie_inner = 0; c.is_light = false;
        uint32_t kbest = 0;
        // ---- A: fp32 boxes of every non-mesh reference, all lanes on the same reference (uniform loads, no divergence but the predicate)
        uint32_t cand = 0;
        const uint32_t all_bits = top.n >= 32u ? 0xFFFFFFFFu : (1u << top.n) - 1u;
#pragma unroll 1
        for (uint32_t todo = all_bits & ~top.mesh_bits; todo; todo &= todo - 1u) {  // warp-uniform: the set bits of a kernel parameter
            const uint32_t k = (uint32_t)__ffs((int)todo) - 1u;
            if (COUNT) w1++;
            if (slab6(top.box[k], br, tmin_f, tmax_f) <= tmax_f) cand |= 1u << k;
        }
        // ---- B: f64 tests, one primitive kind at a time; every lane walks ITS OWN candidates of that kind, so lanes that test
        //      different quads (spheres) still run the same instructions.  Quads first: the big occluders shrink t for the rest.
#define PT_TOP_CANDIDATES(MASK, ...)                                                                    \
        for (uint32_t todo = cand & (MASK); todo;) {                                                    \
            const uint32_t k = (uint32_t)__ffs((int)todo) - 1u; todo &= todo - 1u;                      \
            const DNode rb = S.refs[k];                                                                 \
            if (c.ref != kNone && !(slab(rb, br, tmin_f, tmax_f) <= tmax_f)) continue; /* fallen behind the closest hit */ \
            if (COUNT) w2++;                                                                            \
            const uint32_t index = ref_index(rb.a);                                                     \
            __VA_ARGS__                                                                                 \
            if (c.ref != kNone && c.tie_outer == rb.b) { kbest = k; tmax_f = __double2float_ru(c.t); } /* outer ranks are unique per reference */ \
        }
        PT_TOP_CANDIDATES(top.quad_bits, { double t, a, b; if (quad_t(S.quads[index], r, t_min, t, a, b) && t <= c.t) consider(c, t, rb.a, kInstNone, rb.b, 0); })
        PT_TOP_CANDIDATES(top.sphere_bits, test_simple(S, PT_PRIM_SPHERE, index, r, t_min, c, kInstNone, rb.b, 0);)
        PT_TOP_CANDIDATES(~(top.quad_bits | top.sphere_bits), {
            const uint32_t kind = ref_kind(rb.a);
            if (kind == PT_OBJ_CUBOID) test_simple(S, kind, index, r, t_min, c, kInstNone, rb.b, 0);
            else {  // instance of a simple primitive or cuboid (instance.rs:34-54)
                const DInstance& in = S.instances[index];
                test_simple(S, in.child_kind, in.child_index, instance_local_ray(in, r), t_min, c, index, rb.b, 0);
            }
        })
#undef PT_TOP_CANDIDATES
        // ---- C: boxes of the mesh references against the FINAL closest hit: only meshes that can still matter are queued
#pragma unroll 1
        for (uint32_t todo = top.mesh_bits; todo; todo &= todo - 1u) {
            const uint32_t k = (uint32_t)__ffs((int)todo) - 1u;
            if (COUNT) w1++;
            if (slab6(top.box[k], br, tmin_f, tmax_f) <= tmax_f) mesh_mask |= 1u << k;
        }
        c.is_light = c.ref != kNone && !(c.tie_outer >> 31);  // objects carry bit 31 in their outer rank (object beats light, Q31)
        HitRec h; h.t = c.t; h.ref = c.ref; h.inst_light = (c.inst & 0x7FFFFFFFu) | (c.is_light ? 0x80000000u : 0u);
        hits[i] = h;
        prov_cls = c.ref == kNone ? (uint32_t)CLS_MISS : (uint32_t)top.cls[kbest];
        if (mesh_mask) ties[i] = make_uint2(c.tie_outer, c.tie_inner); else cls = prov_cls;
        if (COUNT) { hist_add(1, w1); hist_add(2, w2); hist_add(3, __popc(mesh_mask)); }
    }
    __syncwarp();
    queue_append(q, cls, i);
    if (top.mesh_bits) {  // queue the mesh visits (warp-aggregated appends)
        const uint32_t lane = threadIdx.x & 31, nm = __popc(mesh_mask);
#if PT_MESH_MULTI
        // a ray that enters ONE mesh box goes to the entry pass + walk (queue 0); the 1-4 % that enter several go, with their whole mask,
        // to queue 1: k_mesh_multi walks all their meshes in one thread, concurrently with the walk of queue 0
#pragma unroll 1
        for (uint32_t r = 0; r < 2u; r++) {
            const bool has = r == 0 ? nm == 1u : nm > 1u;
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, has);
            if (!b) continue;
            const int leader = __ffs(b) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(mq.count + r, __popc(b));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            PT_ASSERT(base + __popc(b) <= mq.stride);
            if (has) mq.items[(size_t)r * mq.stride + base + __popc(b & ((1u << lane) - 1u))] =
                         r == 0 ? make_uint2(i, ((uint32_t)__ffs((int)mesh_mask) - 1u) | (prov_cls << 8) | 0x80000000u) : make_uint2(i, mesh_mask);
        }
#else
#pragma unroll 1
        for (uint32_t r = 0; r < (uint32_t)kMeshRounds; r++) {  // the r-th entered mesh of every ray for round r
            const bool has = nm > r;
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, has);
            if (!b) break;
            const int leader = __ffs(b) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(mq.count + r, __popc(b));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            PT_ASSERT(base + __popc(b) <= mq.stride);
            if (has) mq.items[(size_t)r * mq.stride + base + __popc(b & ((1u << lane) - 1u))] =
                         make_uint2(i, __fns(mesh_mask, 0, r + 1) | (prov_cls << 8) | (nm == r + 1 ? 0x80000000u : 0u));
        }
#endif
    }
    if (COUNT) {
        w1 = __reduce_add_sync(0xFFFFFFFFu, w1); w2 = __reduce_add_sync(0xFFFFFFFFu, w2);
        if ((threadIdx.x & 31) == 0) { atomicAdd(work + 1, (unsigned long long)w1); atomicAdd(work + 2, (unsigned long long)w2); }
    }
}

// Rays whose top-level pass entered SEVERAL mesh boxes (queue 1 of k_top: {path, mask of the entered mesh references}): one thread walks
// all of them against the provisional hit — world box against the current closest hit, ray into the mesh's space, trace_blas over the
// mesh's own BVH — and the ray joins its shade-class queue.  As rounds 1.. of k_mesh_enter + k_mesh_walk these few rays were four more
// launches per wavefront iteration whose duration was the latency of their longest walk (a third of the mesh time of a 2.5 M-ray
// iteration for 1-4 % of the visits); here they run on a side stream while queue 0 is walked.
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock, 7 * kBlock / kTraceBlock) k_mesh_multi(PathBuf pool, MeshQueues mq, HitRec* __restrict__ hits, const uint2* __restrict__ ties,
                                                                                        Queues q, DScene S, double t_min, unsigned long long* __restrict__ work) {
    const uint32_t count = mq.count[1];
    const uint2* __restrict__ items = mq.items + (size_t)mq.stride;
    const float tmin_f = __double2float_rd(t_min);
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    for (uint32_t j = blockIdx.x * kTraceBlock + threadIdx.x; j - (threadIdx.x & 31) < count; j += gridDim.x * kTraceBlock) {
        uint32_t cls = N_CLS, i = 0;
        if (j < count) {
            const uint2 e = items[j];
            i = e.x;
            const HitRec h0 = hits[i];
            const uint2 tie = ties[i];
            Closest c; c.t = h0.t; c.ref = h0.ref; c.inst = h0.inst_light & 0x7FFFFFFFu; c.tie_outer = tie.x; c.tie_inner = tie.y;
            c.is_light = (h0.inst_light >> 31) != 0;
            c.n_pairs = 0; c.n_wide = 0; c.n_refs = 0; c.n_prims = 0;
            const RayD rw = load_ray(pool, i);
            const BoxRay brw = make_boxray(rw);
#pragma unroll 1
            for (uint32_t todo = e.y; todo; todo &= todo - 1u) {
                const uint32_t k = (uint32_t)__ffs((int)todo) - 1u;
                const DNode rb = S.refs[k];  // a = kind | index of the mesh or instance, b = its outer tie rank
                const float tmax_f = __double2float_ru(c.t);
                if (!(slab(rb, brw, tmin_f, tmax_f) <= tmax_f)) { if (COUNT) hist_add(7, 0); continue; }  // fallen behind the closest hit
                const uint32_t kind = ref_kind(rb.a), index = ref_index(rb.a);
                RayD r = rw;
                uint32_t mesh = index, inst = kInstNone;
                if (kind == PT_OBJ_INSTANCE) { const DInstance& ins = S.instances[index]; r = instance_local_ray(ins, rw); mesh = ins.child_index; inst = index; }
                PT_ASSERT(mesh < S.n_meshes);
                trace_blas<COUNT>(S, S.meshes[mesh].root_entry, r, t_min, c, inst, rb.b);
                if (COUNT) hist_add(7, 1);
            }
            if (COUNT) { w0 += c.n_pairs + 2 * c.n_wide; w1 += c.n_refs; w2 += c.n_prims; }
            HitRec h; h.t = c.t; h.ref = c.ref; h.inst_light = (c.inst & 0x7FFFFFFFu) | (c.is_light ? 0x80000000u : 0u);
            hits[i] = h;
            cls = c.ref == kNone ? (uint32_t)CLS_MISS : class_of_kind(S.materials[hit_material(S, c.ref)].kind);
        }
        __syncwarp();
        queue_append(q, cls, i);
    }
    if (COUNT) {
        w0 = __reduce_add_sync(0xFFFFFFFFu, w0); w1 = __reduce_add_sync(0xFFFFFFFFu, w1); w2 = __reduce_add_sync(0xFFFFFFFFu, w2);
        if ((threadIdx.x & 31) == 0 && (w0 | w1 | w2)) { atomicAdd(work, (unsigned long long)w0); atomicAdd(work + 1, (unsigned long long)w1); atomicAdd(work + 2, (unsigned long long)w2); }
    }
}

#ifndef PT_ENTER_MIN_BLOCKS
#define PT_ENTER_MIN_BLOCKS 4
#endif
template <bool COUNT>
__global__ void __launch_bounds__(kBlock, PT_ENTER_MIN_BLOCKS) k_mesh_enter(PathBuf pool, uint32_t round, MeshQueues mq, const HitRec* __restrict__ hits, const uint2* __restrict__ ties,
                                                           Queues q, DScene S, TopList top, double t_min, unsigned long long* __restrict__ work) {
    const uint32_t count = mq.count[round];
    const uint2* __restrict__ items = mq.items + (size_t)round * mq.stride;
    const float tmin_f = __double2float_rd(t_min);
    const uint32_t lane = threadIdx.x & 31;
    uint32_t w0 = 0;
    for (uint32_t base = blockIdx.x * kBlock; base < count; base += gridDim.x * kBlock) {
        const uint32_t j = base + threadIdx.x;
        bool walk = false;
        uint32_t cls = N_CLS, i = 0;
        WalkRec rec;
        if (j < count) {
            const uint2 e = items[j];
            i = e.x;
            const uint32_t k = e.y & 0xFFu;
            const bool last = (e.y >> 31) != 0;
            PT_ASSERT(k < top.n && ((top.mesh_bits >> k) & 1u) && i < mq.stride);
            const DNode rb = S.refs[k];
            const HitRec h0 = hits[i];
            // shade class of the provisional hit: in round 0 still what k_top noted in the queue entry; later an earlier round
            // may have replaced the hit, so it is looked up again (rounds >= 1 hold about 1 % of the rays)
            const uint32_t prov_cls = round == 0 ? ((e.y >> 8) & 0xFu)
                                                 : (h0.ref == kNone ? (uint32_t)CLS_MISS : class_of_kind(S.materials[hit_material(S, h0.ref)].kind));
            RayD r = load_ray(pool, i);
            BoxRay br = make_boxray(r);
            const float tmax_f = __double2float_ru(h0.t);
            if (slab(rb, br, tmin_f, tmax_f) <= tmax_f) {  // else the mesh has fallen behind the closest hit since it was queued
                uint32_t mesh = ref_index(rb.a), inst = kInstNone;
                if (ref_kind(rb.a) == PT_OBJ_INSTANCE) {
                    const DInstance& ins = S.instances[mesh];
                    r = instance_local_ray(ins, r); br = make_boxray(r); inst = mesh; mesh = ins.child_index;
                }
                uint32_t ce[4]; float ct[4];
                PT_ASSERT(mesh < S.n_meshes && S.meshes[mesh].root2 < S.n_wide2 && (inst == kInstNone || inst < S.n_instances));
                wide2_step(S.wide2[S.meshes[mesh].root2], br, tmin_f, tmax_f, ce, ct);
                if (COUNT) w0 += 2;
                const float kInf = __int_as_float(0x7f800000);
                if (ct[0] < kInf) {
                    walk = true;
                    const uint2 tie = ties[i];
                    rec.o[0] = r.o.x; rec.o[1] = r.o.y; rec.o[2] = r.o.z; rec.d[0] = r.d.x; rec.d[1] = r.d.y; rec.d[2] = r.d.z;
                    rec.t = h0.t; rec.path = i;
                    rec.flags = (last ? 0x80000000u : 0u) | ((uint32_t)top.cls[k] << 12) | (prov_cls << 8);
                    rec.ref = h0.ref; rec.inst_light = h0.inst_light; rec.tie_o = tie.x; rec.tie_i = tie.y;
                    rec.cur_inst = inst; rec.cur_tie = rb.b; rec.pad = 0;
                    uint32_t ni = 0;
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) { rec.init[c4] = make_uint2(ce[c4], __float_as_uint(ct[c4])); ni += ct[c4] < kInf; }
                    rec.n_init = ni;
                }
            }
            if (!walk && last) cls = prov_cls;  // nothing to walk: the provisional hit stands
            if (COUNT && !walk) hist_add(7, 0);
        }
        __syncwarp();
        queue_append(q, cls, i);
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, walk);
        if (b) {
            const int leader = __ffs(b) - 1;
            uint32_t wbase = 0;
            if ((int)lane == leader) wbase = atomicAdd(mq.walk_count + round, __popc(b));
            wbase = __shfl_sync(0xFFFFFFFFu, wbase, leader);
            PT_ASSERT(wbase + __popc(b) <= mq.stride);
            if (walk) {
                char* dst = reinterpret_cast<char*>(mq.walk + (size_t)(wbase + __popc(b & ((1u << lane) - 1u))) * kWalkRecU4);
                const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&rec);
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++) st256(dst + 32 * k4, src[4 * k4], src[4 * k4 + 1], src[4 * k4 + 2], src[4 * k4 + 3]);
            }
        }
    }
    if (COUNT) { w0 = __reduce_add_sync(0xFFFFFFFFu, w0); if (lane == 0 && w0) atomicAdd(work, (unsigned long long)w0); }
}

#ifndef PT_WALK_MIN_BLOCKS
#define PT_WALK_MIN_BLOCKS 7   // resident 128-thread-equivalents per SM (7 -> 72 registers, 28 warps)
#endif
#ifndef PT_WALK_BURST
#define PT_WALK_BURST 8        // work steps between two meeting points of the warp (flush finished rays, refill idle lanes); 4: +3 % time
#endif
#ifndef PT_WALK_REFILL_MIN
#define PT_WALK_REFILL_MIN 8   // idle lanes that trigger a fetch
#endif
#ifndef PT_WALK_TRI_MIN
#define PT_WALK_TRI_MIN 10     // lanes holding a triangle that trigger a triangle step while other lanes still have nodes to open
#endif
#ifndef PT_WALK_TOS
#define PT_WALK_TOS 1          // the top stack entry lives in registers: a pop uses it at once and only PREFETCHES the entry below
#endif                         // (the dependent local-memory load of a pop was 13 % of the kernel's stall samples, profiles/r2_kernels_ncu.md)
#ifndef PT_WALK_PREFETCH
#define PT_WALK_PREFETCH 0     // bit 0: the node a lane will open next, bit 1: the triangle it will test next, prefetched into L1 as soon as known; measured: no gain (profiles/r2_ab/r2_q_walk_prefetch.log)
#endif
constexpr int kWalkMinBlocks = PT_WALK_MIN_BLOCKS;
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock, kWalkMinBlocks * kBlock / kTraceBlock) k_mesh_walk(uint32_t round, MeshQueues mq, HitRec* __restrict__ hits, uint2* __restrict__ ties,
                                                                                                     Queues q, DScene S, double t_min, unsigned long long* __restrict__ work) {
    const uint32_t count = mq.walk_count[round];
    uint32_t* __restrict__ cursor = mq.walk_count + kMeshRounds + round;
    const uint32_t lane = threadIdx.x & 31;
    const float tmin_f = __double2float_rd(t_min);
    uint2 stack[kStack2];
    uint2 tos = make_uint2(0u, 0u);  // PT_WALK_TOS: copy of stack[sp - 1] while sp > 0
    int sp = 0;
    RayD r = make_ray(mk(0, 0, 0), mk(0, 0, 1), 0.0);
    BoxRay br = make_boxray(r);
    Closest c; c.t = 0.0; c.ref = kNone; c.inst = kInstNone; c.tie_outer = 0; c.tie_inner = 0; c.is_light = false;
    uint32_t i = 0, flags = 0, cur_inst = kInstNone, cur_tie = 0, done_cls = N_CLS, done_i = 0;
    uint32_t cur = kNone, pend = kNone;  // node to open / triangle to test
    float tmax_f = 0.f, pend_t = 0.f;
    bool active = false, drained = false, improved = false;
    uint32_t w0 = 0, w2 = 0, n_nodes = 0, n_tris = 0;
    while (true) {
        // ---- all 32 lanes meet here: finished rays join their shade-class queue, idle lanes fetch walk records
        if (__any_sync(0xFFFFFFFFu, done_cls != N_CLS)) { queue_append(q, done_cls, done_i); done_cls = N_CLS; }
        const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (!drained && __popc(idle) >= PT_WALK_REFILL_MIN) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(cursor, __popc(idle));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            drained = base + __popc(idle) >= count;
            const uint32_t j = base + __popc(idle & ((1u << lane) - 1u));
            if (!active && j < count) {
                PT_ASSERT(count <= mq.stride);
                const char* src = reinterpret_cast<const char*>(mq.walk + (size_t)j * kWalkRecU4);
                unsigned long long q0, q1, q2, q3, q4, q5, q6, q7, q8, q9, q10, q11, q12, q13, q14, q15;  // the 128-byte record in four 256-bit loads
                ld256(src, q0, q1, q2, q3); ld256(src + 32, q4, q5, q6, q7); ld256(src + 64, q8, q9, q10, q11); ld256(src + 96, q12, q13, q14, q15);
                r.o = mk(__longlong_as_double(q0), __longlong_as_double(q1), __longlong_as_double(q2));
                r.d = mk(__longlong_as_double(q3), __longlong_as_double(q4), __longlong_as_double(q5));
                c.t = __longlong_as_double(q6); i = (uint32_t)q7; flags = (uint32_t)(q7 >> 32);
                const uint4 a5 = make_uint4((uint32_t)q10, (uint32_t)(q10 >> 32), (uint32_t)q11, 0u);
                const uint4 a6 = make_uint4((uint32_t)q12, (uint32_t)(q12 >> 32), (uint32_t)q13, (uint32_t)(q13 >> 32));
                const uint4 a7 = make_uint4((uint32_t)q14, (uint32_t)(q14 >> 32), (uint32_t)q15, (uint32_t)(q15 >> 32));
                PT_ASSERT(i < mq.stride && a5.z >= 1 && a5.z <= 4);
                c.ref = (uint32_t)q8; c.inst = (uint32_t)(q8 >> 32) & 0x7FFFFFFFu; c.tie_outer = (uint32_t)q9; c.tie_inner = (uint32_t)(q9 >> 32);
                cur_inst = a5.x; cur_tie = a5.y;
                br = make_boxray(r);
                tmax_f = __double2float_ru(c.t);
                // the root node was opened by k_mesh_enter: its entered children, far ones first so that the nearest is popped first
                const uint32_t ni = a5.z;
                sp = 0;
                if (ni > 3) { stack[sp] = make_uint2(a7.z, a7.w); sp++; }
                if (ni > 2) { stack[sp] = make_uint2(a7.x, a7.y); sp++; }
                if (ni > 1) { stack[sp] = make_uint2(a6.z, a6.w); sp++; }
                stack[sp] = make_uint2(a6.x, a6.y); sp++;
                tos = make_uint2(a6.x, a6.y);
                cur = kNone; pend = kNone; improved = false; active = true;
                if (COUNT) { n_nodes = 1; n_tris = 0; }
            }
        }
        if (!__any_sync(0xFFFFFFFFu, active)) {
            if (!drained) continue;
            if (__any_sync(0xFFFFFFFFu, done_cls != N_CLS)) queue_append(q, done_cls, done_i);
            break;
        }
#pragma unroll 1
        for (int s = 0; s < PT_WALK_BURST; s++) {
            // ---- every active lane takes the nearest entries off its stack until it holds a node, or a second triangle turns up
            if (active && cur == kNone) {
#if PT_WALK_TOS
                while (sp > 0) {
                    const uint2 top = tos;
                    const bool beyond = !(__uint_as_float(top.y) <= tmax_f);  // beyond the closest hit
                    if (!beyond && (top.x & kTriBit) && pend != kNone) break;   // a second triangle: stays on the stack
                    sp--;
                    if (sp > 0) tos = stack[sp - 1];                            // prefetch: needed by the NEXT pop only
                    if (beyond) continue;
                    if (top.x & kTriBit) {
                        pend = top.x; pend_t = __uint_as_float(top.y);
                        if (PT_WALK_PREFETCH & 2) prefetch_l1(S.tris + (pend & ~kTriBit));
                        continue;
                    }
                    cur = top.x;
                    if (PT_WALK_PREFETCH & 1) prefetch_l1(S.wide2 + cur);
                    break;
                }
#else
                while (sp > 0) {
                    const uint2 top = stack[sp - 1];
                    if (!(__uint_as_float(top.y) <= tmax_f)) { sp--; continue; }  // beyond the closest hit
                    if (top.x & kTriBit) { if (pend != kNone) break; pend = top.x; pend_t = __uint_as_float(top.y); sp--; continue; }
                    cur = top.x; sp--;
                    break;
                }
#endif
                if (cur == kNone && pend == kNone) {  // stack empty: this visit is finished
                    active = false;
                    const bool last = (flags >> 31) != 0;
                    if (improved) {
                        HitRec h; h.t = c.t; h.ref = c.ref; h.inst_light = (c.inst & 0x7FFFFFFFu) | ((c.tie_outer >> 31) ? 0u : 0x80000000u);
                        hits[i] = h;
                        if (!last) ties[i] = make_uint2(c.tie_outer, c.tie_inner);
                    }
                    if (last) { done_cls = improved ? ((flags >> 12) & 0xFu) : ((flags >> 8) & 0xFu); done_i = i; }
                    if (COUNT) { hist_add(4, n_nodes); hist_add(6, n_tris); hist_add(7, 1); if (improved) hist_add(7, 2); }
                }
            }
            const uint32_t nb = __ballot_sync(0xFFFFFFFFu, active && cur != kNone), tb = __ballot_sync(0xFFFFFFFFu, active && pend != kNone);
            if (!(nb | tb)) break;
            if (!nb || __popc(tb) >= PT_WALK_TRI_MIN) {
                // ---- triangle step (mesh.rs:50-82 in f64) for every lane that holds one
                if (active && pend != kNone) {
                    const uint32_t tri = pend & ~kTriBit;
                    PT_ASSERT(tri < S.n_tris);
                    pend = kNone;
                    double t, u, v;
                    if (COUNT && pend_t <= tmax_f) { w2++; n_tris++; }
                    if (pend_t <= tmax_f && tri_t(S.tris[tri], r, t_min, t, u, v) && t <= c.t) {
                        const uint32_t rank = S.tri_rank[tri];
                        if (t < c.t || cur_tie > c.tie_outer || (cur_tie == c.tie_outer && rank > c.tie_inner)) {  // `consider`: ties by rank
                            c.t = t; c.ref = ref_pack(PT_PRIM_TRIANGLE, tri); c.inst = cur_inst; c.tie_outer = cur_tie; c.tie_inner = rank;
                            improved = true; tmax_f = __double2float_ru(t);
                        }
                    }
                }
            } else if (active && cur != kNone) {
                // ---- box step: open one node, push the entered children far to near; the nearest stays in hand
                uint32_t ce[4]; float ct[4];
                PT_ASSERT(cur < S.n_wide2);
                wide2_step_addr(S.wide2 + cur, br, tmin_f, tmax_f, ce, ct);
                if (COUNT) { w0 += 2; n_nodes++; }
                const float kInf = __int_as_float(0x7f800000);
                cur = kNone;
#define PT_WALK_PUSH(K) { const uint2 e_ = make_uint2(ce[K], __float_as_uint(ct[K])); stack[sp] = e_; tos = e_; sp++; }
                if (ct[3] < kInf && can_push(sp, kStack2)) PT_WALK_PUSH(3)
                if (ct[2] < kInf && can_push(sp, kStack2)) PT_WALK_PUSH(2)
                if (ct[1] < kInf && can_push(sp, kStack2)) PT_WALK_PUSH(1)
                if (ct[0] < kInf) {
                    if (!(ce[0] & kTriBit)) { cur = ce[0]; if (PT_WALK_PREFETCH & 1) prefetch_l1(S.wide2 + cur); }
                    else if (pend == kNone) { pend = ce[0]; pend_t = ct[0]; if (PT_WALK_PREFETCH & 2) prefetch_l1(S.tris + (pend & ~kTriBit)); }
                    else if (can_push(sp, kStack2)) PT_WALK_PUSH(0)
                }
#undef PT_WALK_PUSH
            }
        }
    }
    if (COUNT) {
        w0 = __reduce_add_sync(0xFFFFFFFFu, w0); w2 = __reduce_add_sync(0xFFFFFFFFu, w2);
        if (lane == 0 && (w0 | w2)) { atomicAdd(work, (unsigned long long)w0); atomicAdd(work + 2, (unsigned long long)w2); }
    }
}

// Invariant of the traversal stage, checked by the parity entry points: every ray has joined exactly ONE shade-class queue,
// the one of the material its final hit record shades with (a miss: CLS_MISS).  errors += wrong class; seen[i] counts appends.
__global__ void k_check_queues(Queues q, const HitRec* __restrict__ hits, uint32_t n, uint32_t* __restrict__ seen, uint32_t* __restrict__ errors, DScene S) {
    const uint32_t cls = blockIdx.y, count = q.count[cls];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
        const uint32_t i = q.items[(size_t)cls * q.stride + j];
        if (i >= n) { atomicAdd(errors, 1u); continue; }
        const HitRec h = hits[i];
        const uint32_t want = h.ref == kNone ? (uint32_t)CLS_MISS : class_of_kind(S.materials[hit_material(S, h.ref)].kind);
        if (want != cls) atomicAdd(errors, 1u);
        atomicAdd(seen + i, 1u);
    }
}
__global__ void k_check_seen(const uint32_t* __restrict__ seen, uint32_t n, uint32_t* __restrict__ errors) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && seen[i] != 1u) atomicAdd(errors, 1u);
}
// pt_trace_camera_wavefront: the camera rays k_top<PRIMARY> generated into the pool and their hit records, by pixel
__global__ void k_pool_to_abi(PathBuf pool, uint32_t n, const HitRec* __restrict__ hits, pt_ray* __restrict__ out_rays, pt_hit* __restrict__ out_hits, DScene S) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t pixel, sample;
    const RayD r = load_ray(pool, i, &pixel, &sample);
    pt_ray ro; ro.origin.x = r.o.x; ro.origin.y = r.o.y; ro.origin.z = r.o.z; ro.direction.x = r.d.x; ro.direction.y = r.d.y; ro.direction.z = r.d.z; ro.time = r.time;
    out_rays[pixel] = ro;
    const HitRec hr = hits[i];
    pt_hit o; memset(&o, 0, sizeof(o)); o.instance = PT_NONE;
    if (hr.ref != kNone) {
        const uint32_t inst = hr.inst_light & 0x7FFFFFFFu;
        HitInfoD h;
        reconstruct_hit(S, r, hr.ref, inst, hr.t, h);
        o.hit = 1; o.t = hr.t; o.u = h.u; o.v = h.v;
        o.point.x = h.point.x; o.point.y = h.point.y; o.point.z = h.point.z;
        o.geometric_normal.x = h.gn.x; o.geometric_normal.y = h.gn.y; o.geometric_normal.z = h.gn.z;
        o.shading_normal.x = h.sn.x; o.shading_normal.y = h.sn.y; o.shading_normal.z = h.sn.z;
        o.prim_kind = ref_kind(hr.ref); o.prim_index = ref_index(hr.ref); o.instance = inst == kInstNone ? PT_NONE : inst;
        o.material = h.material; o.front_face = h.front_face; o.is_light = hr.inst_light >> 31;
    }
    out_hits[pixel] = o;
}
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
}  // namespace ptd
