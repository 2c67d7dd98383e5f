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
    uint64_t seed; const uint4* __restrict__ ids; uint32_t i;
    PT_D double operator()(uint32_t v) const { const uint4 id = ids[i]; return keyed_uniform(seed, id.x, id.y, id.z >> 16, v); }
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
        if constexpr (VOL) trace_closest<false, COUNT, WIDE>(S, [&]() { return load_ray(in, i); }, t_min, 0.0, c, PathVol{seed, in.ids, i});
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
