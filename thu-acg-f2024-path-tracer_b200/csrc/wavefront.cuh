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
#ifndef PT_SHADE_BLOCK
#define PT_SHADE_BLOCK 128   // threads per block of the k_shade<class> kernels (the block is the unit of the survivor compaction)
#endif
constexpr int kShadeBlock = PT_SHADE_BLOCK;
#ifndef PT_SHADE_MIN_BLOCKS
#define PT_SHADE_MIN_BLOCKS 4  // resident blocks per SM the shade kernels must allow (register cap 128)
#endif

// Path pool: two arrays of records per buffer (ping-pong A/B), moved with 128-bit loads and stores.  The trace stage reads
// only the 64-byte ray record; the shade stage GATHERS both by path index through its class queue, and a gathered record
// is made of whole 32-byte sectors (2 + 1), so no fetched byte is wasted (ten separate f64 arrays cost ten sectors for
// 80 useful bytes).
struct __align__(32) RayRec { double o[3], d[3], time; uint32_t pixel, sample; };    // 64 B
struct __align__(32) StateRec { double thr[3]; uint32_t rng_bounce, spare; };        // 32 B: throughput, rng_used | bounce << 16, spare word
struct PathBuf { RayRec* ray; StateRec* state; };
// Importance sampler of a lat-long environment map (PT_RENDER_ENV_IMPORTANCE; not reference behaviour):
// a piecewise-constant density over rows x cols cells of the (u, theta/pi) unit square, built on the host in f64 by
// pt_scene_build_env_sampler.  marginal[rows + 1] is the CDF over rows (row 0 = theta 0 = +y), cond[r * (cols + 1) ...]
// the CDF over the columns of row r; both start at 0 and end at 1.
struct DEnvDist { const double* marginal; const double* cond; uint32_t rows, cols; };
struct RenderConst {
    uint64_t seed; uint32_t sample_begin, sample_stride, nan_policy, env_importance;
    DEnvDist env;
    uint32_t sort_mask = 0;  // survivors of a shade block are grouped by (direction octant & sort_mask): bit 0 = y, 1 = x, 2 = z
    // path index -> (sample, pixel tile, row, column) without integer divisions: floor(2^64 / pixels) + 1, floor(2^32 / tiles per row) + 1,
    // floor(2^32 / width) + 1 where the host has checked that the multiply-high is exact for every index of the render, else 0 (divide)
    uint64_t inv_pixels = 0; uint32_t inv_tiles_x = 0, inv_width = 0;
    // Path slots >= fresh_from hold paths that were STARTED in this wavefront iteration (k_top<PRIMARY>): their state record is implicit —
    // throughput (1, 1, 1), bounce 0, kFreshDraws uniforms consumed by generate_ray — and is neither written by k_top nor read by the
    // shade kernels (32 B less to store and 32 B less to gather for half of all segments).  0xFFFFFFFF: every state record is real.
    uint32_t fresh_from = 0xFFFFFFFFu;
};
constexpr uint32_t kFreshDraws = 5;  // generate_ray: two draws per random_offsets call (pixel jitter, lens), one for the time (camera.rs:153-168)

// ---------------------------------------------------------------- camera.rs:133-168
PT_D void random_offsets(Rng& rng, double& x, double& y) {
    double radius = sqrt(rng.next());
    double angle = rng.next() * 2.0 * kPi;
    double sn, cs; pt_sincos(angle, sn, cs);
    x = radius * cs; y = radius * sn;
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
// g = index of a path within one render call -> (pixel, local sample): sample-major, and within a sample a warp covers an
// 8x4 pixel tile (tighter ray bundles than a 32x1 strip) when the image size allows
PT_D void path_pixel(uint64_t g, uint32_t n_pixels, const DCamera& cam, const RenderConst& rc, uint32_t& pix, uint32_t& s_local, uint32_t& row, uint32_t& col) {
    if (rc.inv_pixels) { s_local = (uint32_t)__umul64hi(g, rc.inv_pixels); pix = (uint32_t)(g - (uint64_t)s_local * n_pixels); }
    else { s_local = (uint32_t)(g / n_pixels); pix = (uint32_t)(g % n_pixels); }
    if ((cam.width & 7u) == 0 && (cam.height & 3u) == 0) {
        const uint32_t tile = pix >> 5, within = pix & 31u, tiles_x = cam.width >> 3;
        const uint32_t ty = rc.inv_tiles_x ? __umulhi(tile, rc.inv_tiles_x) : tile / tiles_x, tx = tile - ty * tiles_x;
        row = ty * 4u + (within >> 3); col = tx * 8u + (within & 7u);
        pix = row * cam.width + col;
    } else {
        row = rc.inv_width ? __umulhi(pix, rc.inv_width) : pix / cam.width; col = pix - row * cam.width;
    }
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
    double st, ct, sp, cp; pt_sincos(theta, st, ct); pt_sincos(phi, sp, cp);
    return mk(st * cp, ct, st * sp);
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

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): one instruction moves a whole 32-byte sector per lane, so a
// 64-byte ray record is two loads and every access of the strided record arrays is made of full sectors.
PT_D void ld256(const void* p, unsigned long long& a, unsigned long long& b, unsigned long long& c, unsigned long long& d) {
    asm volatile("ld.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
PT_D void st256(void* p, unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d) {
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
PT_D void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
PT_D unsigned long long pack2(uint32_t lo, uint32_t hi) { return (unsigned long long)lo | ((unsigned long long)hi << 32); }
// ids = {pixel, sample, rng_used | bounce << 16, spare} as the shade kernels carry them
PT_D void store_path(const PathBuf& b, uint32_t i, const RayD& r, d3 thr, uint4 ids) {
    char* p = reinterpret_cast<char*>(b.ray + i);
    st256(p, __double_as_longlong(r.o.x), __double_as_longlong(r.o.y), __double_as_longlong(r.o.z), __double_as_longlong(r.d.x));
    st256(p + 32, __double_as_longlong(r.d.y), __double_as_longlong(r.d.z), __double_as_longlong(r.time), pack2(ids.x, ids.y));
    st256(b.state + i, __double_as_longlong(thr.x), __double_as_longlong(thr.y), __double_as_longlong(thr.z), pack2(ids.z, ids.w));
}
PT_D void store_ray(const PathBuf& b, uint32_t i, const RayD& r, uint32_t pixel, uint32_t sample) {  // the ray record alone (fresh paths)
    char* p = reinterpret_cast<char*>(b.ray + i);
    st256(p, __double_as_longlong(r.o.x), __double_as_longlong(r.o.y), __double_as_longlong(r.o.z), __double_as_longlong(r.d.x));
    st256(p + 32, __double_as_longlong(r.d.y), __double_as_longlong(r.d.z), __double_as_longlong(r.time), pack2(pixel, sample));
}
PT_D RayD load_ray(const PathBuf& b, uint32_t i, uint32_t* pixel = nullptr, uint32_t* sample = nullptr) {
    const char* p = reinterpret_cast<const char*>(b.ray + i);
    unsigned long long a0, a1, a2, a3, c0, c1, c2, c3;
    ld256(p, a0, a1, a2, a3);
    ld256(p + 32, c0, c1, c2, c3);
    RayD r;
    r.o = mk(__longlong_as_double(a0), __longlong_as_double(a1), __longlong_as_double(a2));
    r.d = mk(__longlong_as_double(a3), __longlong_as_double(c0), __longlong_as_double(c1));
    r.time = __longlong_as_double(c2);
    if (pixel) *pixel = (uint32_t)c3;
    if (sample) *sample = (uint32_t)(c3 >> 32);
    return r;
}
PT_D d3 load_state(const PathBuf& b, uint32_t i, uint32_t& rng_bounce, uint32_t& spare) {
    unsigned long long a0, a1, a2, a3;
    ld256(b.state + i, a0, a1, a2, a3);
    rng_bounce = (uint32_t)a3; spare = (uint32_t)(a3 >> 32);
    return mk(__longlong_as_double(a0), __longlong_as_double(a1), __longlong_as_double(a2));
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
        PT_ASSERT(base + __popc(peers) <= q.stride && i < q.stride);
        q.items[(size_t)cls * q.stride + base + __popc(peers & ((1u << lane) - 1u))] = i;
    }
}

// ---------------------------------------------------------------- parity / test entry kernels
PT_D pt_vec3 to_abi(d3 v) { pt_vec3 r; r.x = v.x; r.y = v.y; r.z = v.z; return r; }
PT_D d3 from_abi(pt_vec3 v) { return mk(v.x, v.y, v.z); }

// Flat top level (k_top): scenes whose World holds at most kTopMax objects + lights are traversed as a LIST — every lane
// of a warp tests reference k at the same time, so the loop control, the reference loads and the primitive kind are
// warp-uniform and only the box-hit predicate diverges.  refs[0 .. n) are the top-level references (lights, then objects);
// cls[k] = shade class of what a hit on reference k shades with; mesh_bit k set = a mesh or an instance of one, which is
// not walked here but queued for the mesh rounds (k_mesh_enter + k_mesh_walk), at most kMeshRounds per ray.
#ifndef PT_MESH_MULTI
#define PT_MESH_MULTI 1   // rays that enter several mesh boxes: one thread walks all their meshes (k_mesh_multi) instead of rounds 1.. of the entry pass + walk
#endif
constexpr int kTopMax = 32;
constexpr int kMeshRounds = 8;
struct TopList {
    uint32_t n, mesh_bits, quad_bits, sphere_bits;  // *_bits: references that are a mesh (or an instance of one) / a bare quad / a bare sphere
    uint8_t cls[kTopMax];
    float box[kTopMax][6];  // fp32 box of reference k (lo xyz, hi xyz) — a copy of refs[k]'s: as a kernel parameter it sits in the constant
                            // bank, so the warp-uniform box loops of k_top read it through the uniform datapath with no load latency
};
// mesh-visit queues of one wavefront iteration: items[round * stride + j] = {path, top reference | provisional class << 8 |
// last << 31}; walk records of the round that is being processed; counters: count[round], then the walk counter and cursor
struct MeshQueues { uint2* items; uint32_t* count; uint32_t stride; uint4* walk; uint32_t* walk_count; };
struct GenArgs { uint64_t g0; uint32_t n_pixels; DCameraEx cam; RenderConst rc; };  // k_top<PRIMARY>: paths g0 + j are generated in registers
constexpr int kWalkRecU4 = 8;   // a walk record is 128 bytes (trace_kernels.cuh: WalkRec)
constexpr int kStack2 = 32;     // traversal stack of the mesh walk (three pushes per level at most; checked on upload)

// two-pass traversal (trace_kernels.cuh): at most kDeferMax mesh visits are queued per ray and walked by later rounds
constexpr int kDeferMax = 3;
struct BlasQueues { uint4* items; uint32_t* count; uint32_t stride; };  // items[round * stride + k] = {path, ref slot | last << 31, entry t, -}
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
struct SahBox { double lo[3], hi[3]; };  // pt_sah_sweep (misc_kernels.cuh)
constexpr int kSahTile = 128;

}  // namespace ptd
