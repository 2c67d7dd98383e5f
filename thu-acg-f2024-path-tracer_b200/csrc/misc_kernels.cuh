// Ray generation, tonemap and the parity / test entry kernels — compiled by misc.cu only.
#pragma once
#include "wavefront.cuh"

namespace ptd {

// g = global index of the path within this render call: sample-major so that consecutive threads take
// neighbouring pixels of the same sample (coherent primary rays).
__global__ void __launch_bounds__(kBlock) k_generate(PathBuf out, uint32_t slot0, uint32_t n_new, uint64_t g0, uint32_t n_pixels,
                                                      DCameraEx cam, RenderConst rc) {
    uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n_new) return;
    uint32_t s_local, pix, row, col;
    path_pixel(g0 + i, n_pixels, cam.c, rc, pix, s_local, row, col);
    uint32_t sample = rc.sample_begin + s_local * rc.sample_stride;
    Rng rng; rng.init(rc.seed, pix, sample, 0);
    RayD r = generate_ray(cam, row, col, rng);
    store_path(out, slot0 + i, r, mk(1, 1, 1), make_uint4(pix, sample, rng.used, 0));
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
// pt_render_multi: the reduce(sum) of SURVEY §8(e) on the device.  src[g] is the radiance-sum buffer of share g — device
// memory of ANOTHER GPU for g > 0, read straight over NVLink (peer access) by the GPU that owns `out`; partial sums are added
// in share order in f64 and scaled once, exactly like the host loop this replaces, so the image is bit-identical to it.
__global__ void k_reduce_peers(const float* const* __restrict__ src, uint32_t n_src, double scale, size_t n_values, float* __restrict__ out) {
    const size_t n4 = n_values / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (uint32_t g = 0; g < n_src; g++) {
            const float4 v = reinterpret_cast<const float4*>(src[g])[i];
            a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
        }
        reinterpret_cast<float4*>(out)[i] = make_float4((float)(a0 * scale), (float)(a1 * scale), (float)(a2 * scale), (float)(a3 * scale));
    }
    if (blockIdx.x == 0 && threadIdx.x < n_values - 4 * n4) {  // up to three trailing values
        const size_t i = 4 * n4 + threadIdx.x;
        double a = 0.0;
        for (uint32_t g = 0; g < n_src; g++) a += (double)src[g][i];
        out[i] = (float)(a * scale);
    }
}
__global__ void k_scale(const float* __restrict__ accum, float scale, uint32_t n_values, float* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_values) out[i] = accum[i] * scale;
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

// pt_debug_div_check: div3_shared (device_scene.cuh) against the `/` operator on n pseudo-random (numerator, divisor) pairs, bit for
// bit.  Bit patterns come from Philox, so every exponent — denormals, infinities, NaNs, zeros — turns up; a second stream restricts
// both operands to the ranges the shade kernels divide in (|x| in [1e-12, 1e12]).  mismatches += pairs whose bits differ.
__global__ void k_div_check(uint64_t n, uint64_t seed, unsigned long long* __restrict__ mismatches) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x0 = (uint32_t)i, x1 = (uint32_t)(i >> 32), x2 = 0x2545F491u, x3 = 0u, a = (uint32_t)seed, c = (uint32_t)(seed >> 32);
    uint32_t w[8];
#pragma unroll
    for (int blk = 0; blk < 2; blk++) {
        uint32_t y0 = x0, y1 = x1, y2 = x2, y3 = (uint32_t)blk, ka = a, kc = c;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            uint32_t hi0 = __umulhi(0xD2511F53u, y0), lo0 = 0xD2511F53u * y0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, y2), lo1 = 0xCD9E8D57u * y2;
            uint32_t n0 = hi1 ^ y1 ^ ka, n2 = hi0 ^ y3 ^ kc;
            y0 = n0; y1 = lo1; y2 = n2; y3 = lo0;
            ka += 0x9E3779B9u; kc += 0xBB67AE85u;
        }
        w[4 * blk] = y0; w[4 * blk + 1] = y1; w[4 * blk + 2] = y2; w[4 * blk + 3] = y3;
    }
    (void)x3;
    double v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t hi = w[2 * k], lo = w[2 * k + 1];
        if (i & 1) hi = (hi & 0x800FFFFFu) | ((0x3FFu - 40u + (hi >> 20) % 80u) << 20);  // odd pairs: exponent within 2^-40 .. 2^39
        v[k] = __hiloint2double((int)hi, (int)lo);
    }
    if ((i & 15) == 3) { v[0] = 0.0; v[1] = -0.0; }                                   // zero numerators: the shade kernels' commonest special case
    if ((i & 255) == 7) v[3] = (i & 256) ? 0.0 : __longlong_as_double(0x7ff0000000000000ll);  // ... over a zero / infinite divisor
    const double s = v[3];
    const d3 q = div3_shared(mk(v[0], v[1], v[2]), s);
    const double e[3] = {v[0] / s, v[1] / s, v[2] / s};
    const double g[3] = {q.x, q.y, q.z};
    uint32_t bad = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const long long bq = __double_as_longlong(g[k]), be = __double_as_longlong(e[k]);
        const bool both_nan = g[k] != g[k] && e[k] != e[k];  // NaN payloads may differ between the two routes; NaN-ness may not
        if (bq != be && !both_nan) bad++;
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

}  // namespace ptd
