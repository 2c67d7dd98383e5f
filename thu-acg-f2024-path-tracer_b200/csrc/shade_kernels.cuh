// Shade-stage kernels (the loop body of Camera::trace after intersect_all, camera.rs:180-225) — templates instantiated by shade_*.cu.
#pragma once
#include "wavefront.cuh"

namespace ptd {

PT_D void add_radiance(float* __restrict__ accum, uint32_t pix, d3 v, uint32_t nan_policy, unsigned long long* nonfinite, bool& dead) {
    if (!finite3(v)) {
        atomicAdd(nonfinite, 1ull);
        if (nan_policy == PT_NAN_DROP) { dead = true; return; }  // drop the contribution and end the path
    }
    if (v.x != 0.0) atomicAdd(accum + 3ull * pix, (float)v.x);
    if (v.y != 0.0) atomicAdd(accum + 3ull * pix + 1, (float)v.y);
    if (v.z != 0.0) atomicAdd(accum + 3ull * pix + 2, (float)v.z);
}

// The material and texture tables are small (scene 6: 14 materials, 20 textures) but sit at the end of chains of dependent
// loads (hit -> primitive -> material -> texture -> checker child ...); the shade kernels are latency-bound, so every block
// copies them into shared memory once and redirects the scene's pointers (generic addressing) when they fit.
#ifndef PT_SHADE_SMEM
#define PT_SHADE_SMEM 0   // measured: no gain (shade 15.2 ms without, 15.7 ms with, scene 6 FHD x 32 spp): the tables already hit L1
#endif
constexpr uint32_t kShadeTableBytes = 6144;
PT_D void stage_tables(DScene& S) {
#if PT_SHADE_SMEM
    __shared__ __align__(16) unsigned char tab[kShadeTableBytes];
    const uint32_t mb = S.n_materials * (uint32_t)sizeof(DMaterial), tb = S.n_textures * (uint32_t)sizeof(DTexture);
    if (mb + tb <= kShadeTableBytes && mb + tb > 0) {  // block-uniform
        const uint4* src_m = reinterpret_cast<const uint4*>(S.materials);
        const uint4* src_t = reinterpret_cast<const uint4*>(S.textures);
        uint4* dst = reinterpret_cast<uint4*>(tab);
        for (uint32_t k = threadIdx.x; k < mb / 16; k += blockDim.x) dst[k] = src_m[k];
        for (uint32_t k = threadIdx.x; k < tb / 16; k += blockDim.x) dst[mb / 16 + k] = src_t[k];
        __syncthreads();
        S.materials = reinterpret_cast<const DMaterial*>(tab);
        S.textures = reinterpret_cast<const DTexture*>(tab + mb);
    }
#endif
}

// PT_SHADE_PREFETCH: the path records of a thread's NEXT loop iteration are prefetched into L1 while the current path is shaded
// (the gathers queue -> path -> ray / state / hit are a chain of dependent loads: 26 % of the diffuse kernel's stall samples).
#ifndef PT_SHADE_PREFETCH
#define PT_SHADE_PREFETCH 1   // measured: scene 6 FHD +1.4 %, scene 7m +0.7 %, scene 3 +1.7 % (profiles/r2_ab/r2_n_ab1.log)
#endif
// PT_SHADE_WARP_COMPACT: survivors are compacted per WARP (grouped by octant inside the warp, one atomicAdd per warp) instead of per
// block: no shared memory and no __syncthreads in the loop (barrier stalls: 5-11 % of the shade kernels' samples).
#ifndef PT_SHADE_WARP_COMPACT
#define PT_SHADE_WARP_COMPACT 0   // measured: scene 6 +0.5 %, scenes 7m / 3 -2.2 % (the block-wide octant grouping is worth more than the barriers cost)
#endif
// PT_SHADE_WARP_SCAN: the block prefix of the survivor compaction is computed by warp 0 with shuffles (one lane per counter) instead of
// serially by thread 0, which every other thread of the block waits for at the second barrier.
#ifndef PT_SHADE_WARP_SCAN
#define PT_SHADE_WARP_SCAN 1
#endif
PT_D void prefetch_path(const PathBuf& b, const HitRec* __restrict__ hits, uint32_t i, bool with_hit, uint32_t fresh_from) {
    prefetch_l1(b.ray + i);
    if (i < fresh_from) prefetch_l1(b.state + i);
    if (with_hit) prefetch_l1(hits + i);
}

template <int CLS> struct ClassKind { static constexpr int value = -1; };
template <> struct ClassKind<CLS_LIGHT> { static constexpr int value = PT_MAT_LIGHT; };
template <> struct ClassKind<CLS_DIFFUSE> { static constexpr int value = PT_MAT_DIFFUSE; };
template <> struct ClassKind<CLS_METAL> { static constexpr int value = PT_MAT_METAL; };
template <> struct ClassKind<CLS_GLASS> { static constexpr int value = PT_MAT_GLASS; };
template <> struct ClassKind<CLS_PRINCIPLED> { static constexpr int value = PT_MAT_PRINCIPLED; };

// One loop iteration of Camera::trace after intersect_all (camera.rs:180-225) for ONE path whose closest hit shades with class CLS:
// miss / environment, emission, Russian roulette, the light / BSDF mixture sample, eval / pdf, the next ray.  Returns true when
// the path continues: `next`, `thr` and `ids` then describe it.  Shared by the per-class wavefront kernels (k_shade) and by the
// tail megakernel (tail_kernels.cuh).  `after_hit()` runs once the path's own records have been consumed (prefetch hook).
// VAR bit 0: environment importance sampling joins the mixture; bit 1: World.lights holds more than quads and spheres
// (cuboid / mesh / instance lights).  The reference's shipped scenes need neither, and their kernels carry none of that code.
struct NoHook { PT_D void operator()() const {} };
template <int CLS, int VAR, class Hook = NoHook>
PT_D bool shade_one(const DScene& S, const DCameraEx& cam, const RenderConst& rc, float* __restrict__ accum, unsigned long long* __restrict__ nonfinite,
                    const RayD& ray, const HitRec& hr, d3& thr, uint4& ids, RayD& next, Hook after_hit = Hook()) {
    constexpr int K = ClassKind<CLS>::value;
    const uint32_t pix = ids.x, bounces = ids.z >> 16;
    bool dead = false, alive = false;
    if (CLS == CLS_MISS) {  // camera.rs:180-183
        after_hit();
        add_radiance(accum, pix, thr * sample_environment(S, cam.c, ray.d), rc.nan_policy, nonfinite, dead);
        return false;
    }
    Rng rng; rng.init(rc.seed, pix, ids.y, ids.z & 0xFFFFu);
    HitInfoD h;
    reconstruct_hit<false>(S, ray, hr.ref, hr.inst_light & 0x7FFFFFFFu, hr.t, h);
    after_hit();
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
        // VAR bit 2: World.lights is empty (known on the host): the light sampler, the light pdf and the two-block RNG cache drop out of the kernel
        constexpr bool env_is = (VAR & 1) != 0, GEN = (VAR & 2) != 0, NOL = (VAR & 4) != 0;
        const uint32_t n_lights = NOL ? 0u : S.n_lights;
        const double p_env = env_is ? (n_lights == 0 ? 0.5 : 0.25) : 0.0;
        const double p_light = n_lights == 0 ? 0.0 : 0.5 - p_env, p_bsdf = env_is ? 0.5 : 1.0 - p_light;
        const double rsel = rng.next();
        // with lights in the scene half of the lanes go to the light sampler, half to the BSDF sampler: the Philox blocks both consume are
        // computed by all lanes together first (scene 3 +1.7 %; without lights the second block would mostly be wasted)
        if (!NOL && n_lights != 0) rng.prefetch();
        d3 dir;
        bool ok;
        // frame, view direction and roughness of this hit: derived once for sample AND eval / pdf (bsdf.cuh: BsdfCtx)
        constexpr bool kCtx = PT_BSDF_CTX && (K == PT_MAT_DIFFUSE || K == PT_MAT_METAL || K == PT_MAT_GLASS || K == PT_MAT_PRINCIPLED);
        BsdfCtx cxv;
        if (kCtx) cxv = bsdf_prepare<K>(S, m, -ray.d, h);
        const BsdfCtx* cx = kCtx ? &cxv : nullptr;
        if (!NOL && rsel < p_light) ok = lights_sample<GEN>(S, h.point, ray.time, rng, dir);
        else if (env_is && rsel < p_light + p_env) { const double u1 = rng.next(), u2 = rng.next(); dir = env_sample(rc.env, u1, u2); ok = true; }
        else ok = bsdf_sample<K>(S, h.material, ray.d, h, rng, dir, cx);
        if (ok) {  // camera.rs:212-225
            d3 f; double bsdf_pdf;
            bsdf_eval_pdf<K>(S, h.material, -ray.d, dir, h, f, bsdf_pdf, cx);
            double light_pdf = NOL ? 0.0 : lights_pdf<GEN>(S, h.point, dir, ray.time);
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
    return alive;
}

// The paths of ONE shade class: grid-stride over the class queue, shade_one per path; survivors are written compacted into `out`
// (ballot + block prefix + one atomic).
struct PrefetchHook {
    const PathBuf& b; const HitRec* __restrict__ hits; uint32_t i; bool go, with_hit; uint32_t fresh_from;
    PT_D void operator()() const { if (go) prefetch_path(b, hits, i, with_hit, fresh_from); }
};
template <int CLS, int VAR = 0>
__global__ void __launch_bounds__(kShadeBlock, PT_SHADE_MIN_BLOCKS * 128 / kShadeBlock) k_shade(PathBuf in, Queues q, const HitRec* __restrict__ hits, PathBuf out,
                                                    uint32_t* __restrict__ out_count, float* __restrict__ accum,
                                                    unsigned long long* __restrict__ nonfinite, DScene S, DCameraEx cam, RenderConst rc) {
    const uint32_t count = q.count[CLS];
    const uint32_t* __restrict__ items = q.items + (size_t)CLS * q.stride;
    __shared__ uint32_t bin_count[2][8][kShadeBlock / 32];  // survivors per (octant bin, warp), then their exclusive prefix; double-buffered
    __shared__ uint32_t block_base[2];
    stage_tables(S);
    uint32_t par = 0;
#if PT_SHADE_PREFETCH
    uint32_t i_next = blockIdx.x * kShadeBlock + threadIdx.x < count ? items[blockIdx.x * kShadeBlock + threadIdx.x] : 0u;
#endif
    for (uint32_t base = blockIdx.x * kShadeBlock; base < count; base += gridDim.x * kShadeBlock, par ^= 1u) {
        const uint32_t j = base + threadIdx.x;
        bool alive = false;
        RayD next; d3 thr = mk(0, 0, 0); uint4 ids = make_uint4(0, 0, 0, 0);
#if PT_SHADE_PREFETCH
        // the queue entry of this thread's NEXT iteration is read one iteration ahead; its path records are prefetched into L1
        // by shade_one's hook, once this path's own loads have landed
        const uint32_t i = i_next;
        const uint32_t jn = j + gridDim.x * kShadeBlock;
        const bool has_next = jn < count;
        if (has_next) i_next = items[jn];
#endif
        if (j < count) {
#if !PT_SHADE_PREFETCH
            const uint32_t i = items[j];
#endif
            PT_ASSERT(i < q.stride);
            const RayD ray = load_ray(in, i, &ids.x, &ids.y);
            if (i >= rc.fresh_from) { thr = mk(1, 1, 1); ids.z = kFreshDraws; ids.w = 0; }  // started this iteration: implicit state
            else thr = load_state(in, i, ids.z, ids.w);
            HitRec hr; hr.t = 0.0; hr.ref = kNone; hr.inst_light = 0;
            if (CLS != CLS_MISS) hr = hits[i];
#if PT_SHADE_PREFETCH
            alive = shade_one<CLS, VAR>(S, cam, rc, accum, nonfinite, ray, hr, thr, ids, next, PrefetchHook{in, hits, i_next, has_next, CLS != CLS_MISS, rc.fresh_from});
#else
            alive = shade_one<CLS, VAR>(S, cam, rc, accum, nonfinite, ray, hr, thr, ids, next);
#endif
        }
        if (CLS == CLS_MISS) continue;  // a miss ends the path: nothing to compact
        // ---- compaction: survivors grouped by direction octant within the block (warps of the next trace launch then hold
        //      rays that walk the BVH in the same order and tend to cost the same), one atomic per block
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t key = alive ? (((next.d.y < 0.0) | ((next.d.x < 0.0) << 1) | ((next.d.z < 0.0) << 2)) & rc.sort_mask) : 8u;
#if PT_SHADE_WARP_COMPACT
        {
            uint32_t mine_w = 0, before = 0;  // lanes of my bin; survivors of this warp in lower bins
#pragma unroll
            for (uint32_t b = 0; b < 8; b++) {
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, key == b);
                if (key == b) mine_w = m;
                if (b < key) before += __popc(m);
            }
            const uint32_t live = __ballot_sync(0xFFFFFFFFu, alive);
            uint32_t wbase = 0;
            if (live) {
                const int leader = __ffs(live) - 1;
                if ((int)lane == leader) wbase = atomicAdd(out_count, __popc(live));
                wbase = __shfl_sync(0xFFFFFFFFu, wbase, leader);
            }
            if (alive) {
                const uint32_t dst = wbase + before + __popc(mine_w & ((1u << lane) - 1u));
                PT_ASSERT(dst < q.stride);
                store_path(out, dst, next, thr, ids);
            }
            (void)warp;
            continue;
        }
#endif
        uint32_t mine = 0, cnt = 0;  // lanes of my bin; lane b < 8 also holds the size of bin b
#pragma unroll
        for (uint32_t b = 0; b < 8; b++) {
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, key == b);
            if (key == b) mine = m;
            if (lane == b) cnt = __popc(m);
        }
        if (lane < 8) bin_count[par][lane][warp] = cnt;
        __syncthreads();
#if PT_SHADE_WARP_SCAN
        if (warp == 0) {  // exclusive prefix over the 8 x (warps per block) counters, bin-major, by one warp: lane l <-> (bin l / nw, warp l % nw)
            constexpr uint32_t nw = kShadeBlock / 32;
            static_assert(8 * nw <= 32, "one lane per (bin, warp) counter");
            const uint32_t c = lane < 8 * nw ? bin_count[par][lane / nw][lane % nw] : 0u;
            uint32_t incl = c;
#pragma unroll
            for (uint32_t d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (lane < 8 * nw) bin_count[par][lane / nw][lane % nw] = incl - c;
            if (lane == 0) block_base[par] = total ? atomicAdd(out_count, total) : 0;
        }
#else
        if (threadIdx.x == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int b = 0; b < 8; b++)
#pragma unroll
                for (int w = 0; w < kShadeBlock / 32; w++) { uint32_t c = bin_count[par][b][w]; bin_count[par][b][w] = total; total += c; }
            block_base[par] = total ? atomicAdd(out_count, total) : 0;
        }
#endif
        __syncthreads();
        if (alive) {
            uint32_t dst = block_base[par] + bin_count[par][key][warp] + __popc(mine & ((1u << lane) - 1u));
            PT_ASSERT(dst < q.stride);
            store_path(out, dst, next, thr, ids);
        }
        // no third barrier: the next iteration works on the other half of bin_count / block_base, and nobody can reach the
        // iteration after that before every thread has passed the two barriers in between
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
    stage_tables(S);
    for (uint32_t base = blockIdx.x * kBlock; base < count; base += gridDim.x * kBlock) {
        const uint32_t j = base + threadIdx.x;
        bool alive = false, shadow = false;
        RayD next, sray; d3 thr = mk(0, 0, 0), sthr = mk(0, 0, 0); uint4 ids = make_uint4(0, 0, 0, 0);
        if (j < count) {
            const uint32_t i = items[j];
            PT_ASSERT(i < q.stride);
            RayD ray = load_ray(in, i, &ids.x, &ids.y);
            thr = load_state(in, i, ids.z, ids.w);
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

}  // namespace ptd
