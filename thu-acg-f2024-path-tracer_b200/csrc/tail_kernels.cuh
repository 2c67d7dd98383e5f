// Tail megakernel — compiled by tail_wide.cu / tail_bin.cu.
//
// The reference lets a path run max_depth = 50 bounces (camera.rs:177) with Russian roulette only after the fifth, so every
// render ends with ~45 wavefront iterations of a few thousand, then a few hundred, then a handful of rays.  Each of those
// costs a traversal launch plus up to six shade launches whose duration is the latency of ONE ray (~0.2 ms per iteration,
// 5-10 ms per render call whatever its size: a fifth of a 600-px x 100-spp render, 8 % of a 128-spp step at 1920 x 1080).
// k_tail finishes such a wavefront in ONE launch: a thread owns a path and loops World::intersect_all (the fused BVH
// traversal, geom.cuh: trace_closest) and the same shade_one<class> the wavefront kernels run, until the path ends.  Lanes of
// a warp shade different materials and die at different bounces — poor SIMD efficiency, but on a wavefront that no longer
// fills the machine it is the latency of the longest path that counts, not throughput.
// Same RNG contract, same functions, same per-path arithmetic: the samples are bit-identical to the wavefront's; only the
// order of the fp32 atomic adds into the accumulator differs (as it already does between any two runs).
#pragma once
#include "shade_kernels.cuh"

namespace ptd {

// World::intersect_all for ONE ray of a flat scene (TopList, wavefront.cuh) without the generic two-level BVH walk: the top-level references
// as a list — fp32 boxes from the constant bank, then the f64 tests of the entered ones — and every entered mesh by a single-ray walk of
// the mesh-walk nodes (DWide2: children are nodes or single triangles with their own boxes).  Same closest hit as trace_closest: minimum
// t, exact ties by the precomputed ranks (`consider`), so the order of the tests does not matter.  (54 % of k_tail's stall samples were
// the generic traversal: tagged stack entries, leaf reference loops and per-reference box fetches that a list of ten objects does not need.)
PT_D void walk_wide2_single(const DScene& S, uint32_t root2, const RayD& r, double t_min, Closest& c, uint32_t cur_inst, uint32_t cur_tie) {
    uint2 stack[kStack2];
    int sp = 0;
    const BoxRay br = make_boxray(r);
    const float tmin_f = __double2float_rd(t_min);
    float tmax_f = __double2float_ru(c.t);
    const float kInf = __int_as_float(0x7f800000);
    auto test_tri = [&](uint32_t e) {  // mesh.rs:50-82 in f64, ties by (outer, inner) rank
        const uint32_t tri = e & ~kTriBit;
        PT_ASSERT(tri < S.n_tris);
        double t, u, v;
        if (tri_t(S.tris[tri], r, t_min, t, u, v) && t <= c.t) {
            const uint32_t rank = S.tri_rank[tri];
            if (t < c.t || cur_tie > c.tie_outer || (cur_tie == c.tie_outer && rank > c.tie_inner)) {
                c.t = t; c.ref = ref_pack(PT_PRIM_TRIANGLE, tri); c.inst = cur_inst; c.tie_outer = cur_tie; c.tie_inner = rank;
                tmax_f = __double2float_ru(t);
            }
        }
    };
    uint32_t cur = root2;
    while (true) {
        if (cur == kNone) {
            while (sp > 0) {
                const uint2 top = stack[--sp];
                if (!(__uint_as_float(top.y) <= tmax_f)) continue;  // beyond the closest hit
                if (top.x & kTriBit) { test_tri(top.x); continue; }
                cur = top.x;
                break;
            }
            if (cur == kNone) return;
        }
        uint32_t ce[4]; float ct[4];
        PT_ASSERT(cur < S.n_wide2);
        wide2_step_addr(S.wide2 + cur, br, tmin_f, tmax_f, ce, ct);
        cur = kNone;
        if (ct[3] < kInf && can_push(sp, kStack2)) { stack[sp] = make_uint2(ce[3], __float_as_uint(ct[3])); sp++; }
        if (ct[2] < kInf && can_push(sp, kStack2)) { stack[sp] = make_uint2(ce[2], __float_as_uint(ct[2])); sp++; }
        if (ct[1] < kInf && can_push(sp, kStack2)) { stack[sp] = make_uint2(ce[1], __float_as_uint(ct[1])); sp++; }
        if (ct[0] < kInf) { if (ce[0] & kTriBit) test_tri(ce[0]); else cur = ce[0]; }
    }
}
PT_D void trace_flat(const DScene& S, const TopList& top, const RayD& r, double t_min, Closest& c) {
    const BoxRay br = make_boxray(r);
    const float tmin_f = __double2float_rd(t_min);
    float tmax_f = __int_as_float(0x7f800000);
    c.t = __longlong_as_double(0x7ff0000000000000ll); c.ref = kNone; c.inst = kInstNone; c.tie_outer = 0; c.tie_inner = 0; c.is_light = false;
    c.n_pairs = 0; c.n_wide = 0; c.n_refs = 0; c.n_prims = 0;
    const uint32_t all_bits = top.n >= 32u ? 0xFFFFFFFFu : (1u << top.n) - 1u;
    // simple primitives, cuboids and instances of them: box, then the f64 test against the shrinking interval
#pragma unroll 1
    for (uint32_t todo = all_bits & ~top.mesh_bits; todo; todo &= todo - 1u) {
        const uint32_t k = (uint32_t)__ffs((int)todo) - 1u;
        if (!(slab6(top.box[k], br, tmin_f, tmax_f) <= tmax_f)) continue;
        const DNode rb = S.refs[k];
        const uint32_t kind = ref_kind(rb.a), index = ref_index(rb.a);
        if (kind == PT_OBJ_INSTANCE) {  // instance of a simple primitive or cuboid (instance.rs:34-54)
            const DInstance& in = S.instances[index];
            test_simple(S, in.child_kind, in.child_index, instance_local_ray(in, r), t_min, c, index, rb.b, 0);
        } else test_simple(S, kind, index, r, t_min, c, kInstNone, rb.b, 0);
        tmax_f = __double2float_ru(c.t);
    }
    // meshes and instances of meshes
#pragma unroll 1
    for (uint32_t todo = top.mesh_bits; todo; todo &= todo - 1u) {
        const uint32_t k = (uint32_t)__ffs((int)todo) - 1u;
        if (!(slab6(top.box[k], br, tmin_f, tmax_f) <= tmax_f)) continue;
        const DNode rb = S.refs[k];
        const uint32_t kind = ref_kind(rb.a), index = ref_index(rb.a);
        RayD rl = r;
        uint32_t mesh = index, inst = kInstNone;
        if (kind == PT_OBJ_INSTANCE) { const DInstance& in = S.instances[index]; rl = instance_local_ray(in, r); mesh = in.child_index; inst = index; }
        PT_ASSERT(mesh < S.n_meshes);
        walk_wide2_single(S, S.meshes[mesh].root2, rl, t_min, c, inst, rb.b);
        tmax_f = __double2float_ru(c.t);
    }
    c.is_light = c.ref != kNone && !(c.tie_outer >> 31);  // objects carry bit 31 in their outer rank (object beats light, Q31)
}

#ifndef PT_TAIL_BLOCK
#define PT_TAIL_BLOCK 64
#endif
constexpr int kTailBlock = PT_TAIL_BLOCK;
// MODE 0 / 1: the fused BVH traversal over binary pairs / 4-wide nodes (large Worlds); 2: flat scenes (trace_flat)
template <int MODE>
__global__ void __launch_bounds__(kTailBlock) k_tail(PathBuf in, uint32_t n, float* __restrict__ accum, unsigned long long* __restrict__ nonfinite,
                                                      DScene S, DCameraEx cam, RenderConst rc, double t_min, uint32_t* __restrict__ counters, TopList top) {
    const uint32_t i = blockIdx.x * kTailBlock + threadIdx.x;
    uint32_t n_seg = 0;
    if (i < n) {
        uint4 ids = make_uint4(0, 0, 0, 0);
        RayD ray = load_ray(in, i, &ids.x, &ids.y);
        d3 thr = load_state(in, i, ids.z, ids.w);
        bool alive = true;
        while (alive) {
            Closest c;
            if constexpr (MODE == 2) trace_flat(S, top, ray, t_min, c);
            else trace_closest<false, false, MODE == 1>(S, [&]() { return ray; }, t_min, 0.0, c);
            n_seg++;
            HitRec hr; hr.t = c.t; hr.ref = c.ref; hr.inst_light = (c.inst & 0x7FFFFFFFu) | (c.is_light ? 0x80000000u : 0u);
            const uint32_t cls = c.ref == kNone ? (uint32_t)CLS_MISS : class_of_kind(S.materials[hit_material(S, c.ref)].kind);
            RayD next = ray;
            switch (cls) {
#define PT_GO(C) case C: alive = shade_one<C, 0>(S, cam, rc, accum, nonfinite, ray, hr, thr, ids, next); break;
                PT_GO(CLS_MISS) PT_GO(CLS_LIGHT) PT_GO(CLS_DIFFUSE) PT_GO(CLS_METAL) PT_GO(CLS_GLASS) PT_GO(CLS_PRINCIPLED)
                default: alive = shade_one<CLS_OTHER, 0>(S, cam, rc, accum, nonfinite, ray, hr, thr, ids, next); break;
#undef PT_GO
            }
            ray = next;
        }
    }
    // counters[0] += segments traced (one per intersect_all call, like the wavefront's count of live rays per iteration);
    // counters[1] = max over paths = the number of wavefront iterations this launch stands for
    const uint32_t longest = __reduce_max_sync(0xFFFFFFFFu, n_seg), total = __reduce_add_sync(0xFFFFFFFFu, n_seg);
    if ((threadIdx.x & 31) == 0 && total) { atomicAdd(counters, total); atomicMax(counters + 1, longest); }
}

}  // namespace ptd
