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

#ifndef PT_TAIL_BLOCK
#define PT_TAIL_BLOCK 64
#endif
constexpr int kTailBlock = PT_TAIL_BLOCK;
template <bool WIDE>
__global__ void __launch_bounds__(kTailBlock) k_tail(PathBuf in, uint32_t n, float* __restrict__ accum, unsigned long long* __restrict__ nonfinite,
                                                      DScene S, DCameraEx cam, RenderConst rc, double t_min, uint32_t* __restrict__ counters) {
    const uint32_t i = blockIdx.x * kTailBlock + threadIdx.x;
    uint32_t n_seg = 0;
    if (i < n) {
        uint4 ids = make_uint4(0, 0, 0, 0);
        RayD ray = load_ray(in, i, &ids.x, &ids.y);
        d3 thr = load_state(in, i, ids.z, ids.w);
        bool alive = true;
        while (alive) {
            Closest c;
            trace_closest<false, false, WIDE>(S, [&]() { return ray; }, t_min, 0.0, c);
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
