// C ABI implementation (include/pt_b200.h): context, scene upload, the wavefront render loop and the
// parity entry points.  There is NO CPU fallback: without a CUDA device every call fails loudly.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "launch.h"

using namespace ptd;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) return fail(PT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

// Tail of a render (nothing left to generate, the live count only shrinks): kTailBatch wavefront iterations are enqueued
// back to back, each reading its ray count from the survivors counter of the one before, and the host reads all the
// counters in one round trip.  Grids are sized by the live count at the start of the batch (an upper bound).
constexpr int kTailBatch = 8;
// Tail megakernel (tail_kernels.cuh): once nothing is left to generate and at most tail_max_paths() paths are alive, ONE launch of
// k_tail runs every one of them to its end instead of ~45 more wavefront iterations.  The threshold depends on what a small
// wavefront iteration costs on the scene (measured, profiles/r2_ab/r2_o_tail_threshold.log): scenes whose small iterations run
// the fused BVH kernel (a large World, or meshes) gain up to 64 Ki paths (scene 6 FHD x 32 spp: 36.4 -> 32.1 ms; scene 1:
// 15.0 -> 13.4 ms); flat scenes without meshes iterate in ~65 us (k_top + three shade kernels) and the megakernel — 196
// registers, 4 of 32 lanes, every instruction an I-cache miss — only wins below a few thousand paths (scene 3: 20 Ki paths
// +1.3 ms, 700 paths -0.4 ms).  PT_B200_TAIL_MAX overrides the threshold (developer knob; 0 switches the megakernel off).
static uint32_t tail_max_paths(bool cheap_iterations) {  // cheap_iterations: flat scene without meshes
    static const long v = [] { const char* e = getenv("PT_B200_TAIL_MAX"); return e ? (long)strtoul(e, nullptr, 10) : -1l; }();
    if (v >= 0) return (uint32_t)v;
    return cheap_iterations ? (1u << 14) : (1u << 16);  // re-swept with the flat tail traversal: scenes 3 / 7 flat from 4 Ki to 64 Ki, scene 5 best at 16 Ki (+3 %)
}
constexpr int kSlot = 64;  // uint32 counters per wavefront iteration (see pt_ctx::d_count)
constexpr uint32_t kTailBatchMaxLive = 1u << 22;

struct pt_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int profiling = 0;                      // 1: per-stage CUDA events; 2: also the traversal work counters (COUNT kernel variants)
    // path pool (ping-pong SoA) + hit records, grown on demand
    uint32_t pool = 0;
    RayRec* pool_ray[2] = {nullptr, nullptr};
    StateRec* pool_state[2] = {nullptr, nullptr};
    HitRec* hits = nullptr;
    uint32_t* d_count = nullptr;            // kTailBatch slots of [kSlot]: [0] survivors counter, [4 .. 4+N_CLS) shade-class queue lengths; flat scenes: [12 .. 20) mesh-visit
                                            // queue lengths per round, [20 .. 28) walk records per round, [28 .. 36) their fetch cursors; k_trace<DEFER>: [12 .. 15) queue lengths, [16 .. 19) cursors
    uint32_t* q_items = nullptr;            // N_CLS queues of `pool` path slots each
    uint4* bq_items = nullptr;              // two-pass traversal: kDeferMax queues of `pool` deferred mesh visits each
    uint2* ties = nullptr;                  //   and the tie ranks of the provisional hit of every path (both allocated on first use)
    uint2* mq_items = nullptr; uint32_t mq_rounds = 0;  // flat scenes with meshes: mq_rounds queues of `pool` mesh visits each (k_top -> k_mesh_enter)
    uint4* walk = nullptr;                  //   and `pool` 128-byte walk records (k_mesh_enter -> k_mesh_walk)
    // profiling level >= 1: event marks inside an iteration; the time since the previous mark goes to the mark's stage
    // (pt_debug_stage_ms: 0 k_top on survivors, 1 k_top on new paths (+ ray generation), 2 k_mesh_enter, 3 k_mesh_walk,
    //  4 BVH trace kernels, 5 k_generate, 6 shade kernels)
    std::vector<cudaEvent_t> marks; std::vector<int> mark_tags; size_t n_marks = 0; double stage_ms[16] = {0};
    void mark(int tag) {
        if (!profiling) return;
        if (n_marks == marks.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; marks.push_back(e); mark_tags.push_back(0); }
        cudaEventRecord(marks[n_marks], stream); mark_tags[n_marks] = tag; n_marks++;
    }
    void collect_marks() {  // after the stream has been synchronised
        for (size_t k = 1; k < n_marks; k++) { float ms = 0; if (cudaEventElapsedTime(&ms, marks[k - 1], marks[k]) == cudaSuccess && mark_tags[k] >= 0) stage_ms[mark_tags[k] & 15] += ms; }
        n_marks = 0;
    }
    void* scratch = nullptr; size_t scratch_bytes = 0;  // pt_render / pt_tonemap_rgb8 output staging, grown on demand (no cudaMalloc per call)
    unsigned long long* d_nonfinite = nullptr;
    uint32_t* h_count = nullptr;            // pinned, same shape as d_count
    void* h_stage = nullptr; size_t stage_bytes = 0;  // pinned staging buffer for scene uploads
    std::vector<std::pair<void*, size_t>> free_blocks;  // device blocks of destroyed scenes, reused by the next upload
    // cudaMalloc/cudaFree cost milliseconds each and synchronise the device; returns the block and its true size
    int alloc_block(size_t bytes, std::pair<void*, size_t>* out) {
        size_t best = free_blocks.size();
        for (size_t i = 0; i < free_blocks.size(); i++)
            if (free_blocks[i].second >= bytes && free_blocks[i].second <= 2 * bytes + (1 << 20) &&
                (best == free_blocks.size() || free_blocks[i].second < free_blocks[best].second)) best = i;
        if (best != free_blocks.size()) { *out = free_blocks[best]; free_blocks.erase(free_blocks.begin() + best); return 0; }
        out->second = bytes;
        return cudaMalloc(&out->first, bytes) == cudaSuccess ? 0 : -1;
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evs[6] = {nullptr};
    // the shade kernels of one iteration are independent (one queue each): forked onto side streams so that their launch
    // ramps and tails overlap instead of adding up
    cudaStream_t shade_stream[N_CLS] = {nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[N_CLS] = {nullptr};
};
struct pt_scene {
    pt_ctx* ctx = nullptr;
    std::vector<std::pair<void*, size_t>> allocs;  // device blocks (returned to the ctx's free list on destroy)
    DScene d{};
    uint64_t bytes = 0;
    uint32_t n_materials = 0, n_images = 0, max_stack = 0;
    uint32_t class_mask = 1u << CLS_MISS;   // shade classes that can occur in this scene
    bool wide = false;                      // traversed with 4-wide nodes (some BVH has >= kWideMinItems items) or binary pairs
    std::vector<DImage> images;
    bool general_lights = false;            // World.lights holds something other than quads and spheres (shade kernel variant)
    bool has_volumes = false;               // constant-density media present (trace kernel variant with keyed uniforms)
    bool defer_meshes = false;              // two-pass traversal (k_trace<DEFER> + k_trace_blas): wide scenes that hold meshes
    bool flat = false;                      // flat top level (k_top): at most kTopMax objects + lights, no media, meshes walkable by k_mesh_walk
    TopList top{};                          //   its reference list
    uint32_t mesh_rounds = 0;               //   mesh references in it = rounds of k_mesh_enter + k_mesh_walk per iteration
    DEnvDist env{};                         // pt_scene_build_env_sampler: importance sampler of image `env_image`
    uint32_t env_image = 0xFFFFFFFFu;
    void* env_block = nullptr;
    // every exit path (pt_scene_destroy and the error returns of pt_scene_create) hands the device blocks back to the context
    ~pt_scene() {
        if (!ctx) return;
        for (auto& b : allocs) ctx->free_blocks.push_back(b);
        while (ctx->free_blocks.size() > 8) { cudaFree(ctx->free_blocks.front().first); ctx->free_blocks.erase(ctx->free_blocks.begin()); }
        if (env_block) cudaFree(env_block);
    }
};

extern "C" {

const char* pt_last_error(void) { return g_err.c_str(); }
void pt_ctx_destroy(pt_ctx* c);
int pt_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

static int ctx_init(pt_ctx* c) {
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU(cudaMalloc(&c->d_count, kTailBatch * kSlot * sizeof(uint32_t)));
    CU(cudaMalloc(&c->d_nonfinite, 4 * sizeof(unsigned long long)));  // [0] non-finite samples, [1..3] traversal work counters
    CU(cudaMallocHost(&c->h_count, kTailBatch * kSlot * sizeof(uint32_t)));
    CU(cudaEventCreate(&c->ev0)); CU(cudaEventCreate(&c->ev1));
    for (auto& e : c->evs) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int k = 0; k < N_CLS; k++) {
        CU(cudaStreamCreateWithFlags(&c->shade_stream[k], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
    }
    return PT_OK;
}
int pt_ctx_create(int device, pt_ctx** out) {
    if (!out) return fail(PT_ERR_INVALID, "pt_ctx_create: null out");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return fail(PT_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(PT_ERR_INVALID, "pt_ctx_create: bad device index");
    CU(cudaSetDevice(device));
    auto* c = new pt_ctx(); c->device = device;
    const int rc = ctx_init(c);
    if (rc) { const std::string msg = g_err; pt_ctx_destroy(c); return fail(rc, msg); }  // nothing created so far may leak
    *out = c;
    return PT_OK;
}
static void free_pool(pt_ctx* c) {
    for (int i = 0; i < 2; i++) { cudaFree(c->pool_ray[i]); cudaFree(c->pool_state[i]); c->pool_ray[i] = nullptr; c->pool_state[i] = nullptr; }
    cudaFree(c->hits); c->hits = nullptr; cudaFree(c->q_items); c->q_items = nullptr; c->pool = 0;
    cudaFree(c->bq_items); c->bq_items = nullptr; cudaFree(c->ties); c->ties = nullptr;
    cudaFree(c->mq_items); c->mq_items = nullptr; c->mq_rounds = 0; cudaFree(c->walk); c->walk = nullptr;
}
void pt_ctx_destroy(pt_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    free_pool(c);
    cudaFree(c->d_count); cudaFree(c->d_nonfinite); cudaFree(c->scratch);
    if (c->h_count) cudaFreeHost(c->h_count);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (auto& b : c->free_blocks) cudaFree(b.first);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (auto& e : c->evs) if (e) cudaEventDestroy(e);
    for (auto& e : c->marks) cudaEventDestroy(e);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int k = 0; k < N_CLS; k++) { if (c->shade_stream[k]) cudaStreamDestroy(c->shade_stream[k]); if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]); }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    cudaGetLastError();  // a half-built context may have handed CUDA a null handle above
    delete c;
}
int pt_ctx_set_stream(pt_ctx* c, void* s) { if (!c) return fail(PT_ERR_INVALID, "null ctx"); c->stream = s ? (cudaStream_t)s : c->own_stream; return PT_OK; }
int pt_ctx_set_profiling(pt_ctx* c, int on) { if (!c) return fail(PT_ERR_INVALID, "null ctx"); c->profiling = on < 0 ? 0 : (on > 2 ? 2 : on); return PT_OK; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------ scene upload
namespace {
// One device arena per scene: every table is staged into one pinned host buffer (owned by the ctx, reused across
// scenes) and moved with a single H2D copy; 256-byte aligned sub-allocations.
struct Uploader {
    pt_scene* s;
    struct Item { const void* src; size_t bytes; const void** dst; size_t offset; };
    std::vector<Item> items;
    size_t total = 0;
    void add(const void* src, size_t bytes, const void** dst) {
        *dst = nullptr;
        if (!bytes) return;
        items.push_back(Item{src, bytes, dst, total});
        total += (bytes + 255) / 256 * 256;
    }
    template <class T> void up(const std::vector<T>& v, const T** out) { add(v.data(), v.size() * sizeof(T), (const void**)out); }
    int commit() {
        if (!total) return PT_OK;
        pt_ctx* c = s->ctx;
        if (c->stage_bytes < total) {
            if (c->h_stage) cudaFreeHost(c->h_stage);
            c->h_stage = nullptr; c->stage_bytes = 0;
            CU(cudaMallocHost(&c->h_stage, total));
            c->stage_bytes = total;
        }
        std::pair<void*, size_t> blk{nullptr, 0};
        if (c->alloc_block(total, &blk)) return fail(PT_ERR_CUDA, "cudaMalloc failed for the scene arena");
        s->allocs.push_back(blk);
        void* base = blk.first;
        for (auto& it : items) { memcpy((char*)c->h_stage + it.offset, it.src, it.bytes); *it.dst = (char*)base + it.offset; s->bytes += it.bytes; }
        CU(cudaMemcpyAsync(base, c->h_stage, total, cudaMemcpyHostToDevice, c->stream));
        return PT_OK;
    }
};
// the neighbouring fp32 value towards -inf / +inf, as std::nextafterf gives it, by bit arithmetic (the library call was half of the
// box conversion's time: three calls per box, 50 000 boxes on scene 6)
float next_down(float f) {
    if (!(f == f) || f == -INFINITY) return f;
    uint32_t u; memcpy(&u, &f, 4);
    if (f == 0.0f) u = 0x80000001u;           // +-0 -> the smallest negative denormal
    else if (f > 0.0f) u -= 1u; else u += 1u;
    memcpy(&f, &u, 4); return f;
}
float next_up(float f) { return -next_down(-f); }
float round_down(double v) { float f = (float)v; if ((double)f > v) f = next_down(f); return f; }
float round_up(double v) { float f = (float)v; if ((double)f < v) f = next_up(f); return f; }

// Reciprocals for path_pixel (wavefront.cuh).  With 2^K = d k + r (0 < r < d) and M = k + 1: x M / 2^K = q + (q (d - r) + s (k + 1)) / 2^K for
// x = d q + s, so mulhi(x, M) = q = x / d as long as q (d - 1) + (d - 1)(k + 1) < 2^K, i.e. q_max (d - 1) < k + 1 + r - d.  0 = "divide".
template <class U, class W> static U exact_reciprocal(U d, uint64_t x_end) {  // for every x < x_end; U = uint32_t / uint64_t, W = the double-width type
    if (d < 2 || (d & (d - 1)) == 0 || x_end == 0) return 0;                 // powers of two: the division is a shift anyway
    const U k = (U)(~(U)0 / d), r = (U)(0 - d * k);                           // floor(2^K / d) (d does not divide 2^K), 2^K - d k
    const uint64_t q_max = (x_end - 1) / d;
    if ((W)k + 1 + r <= (W)d || (W)q_max * (d - 1) >= (W)k + 1 + r - d) return 0;
    return (U)(k + 1);
}
static void set_index_reciprocals(RenderConst& rc, uint32_t n_pixels, uint64_t total_paths, uint32_t width) {
    rc.inv_pixels = exact_reciprocal<uint64_t, unsigned __int128>(n_pixels, total_paths);
    rc.inv_tiles_x = exact_reciprocal<uint32_t, uint64_t>(width >> 3, (uint64_t)(n_pixels >> 5) + 1);
    rc.inv_width = exact_reciprocal<uint32_t, uint64_t>(width, n_pixels);
}

constexpr uint32_t kWideMinItems = 64;  // BVHs over at least this many items are collapsed to 4-wide nodes

struct Converter {
    const pt_scene_desc* d;
    std::vector<DNode> nodes;
    std::vector<DNode> refs;  // per-reference fp32 box + (kind|index, tie rank)
    uint32_t tie_counter = 0;
    uint32_t max_depth = 0;
    std::string err;
    // which tree is being converted: -1 = a top-level list (objects / lights: anything but bare triangles), else the mesh whose
    // BLAS it is (leaves may hold only that mesh's own triangles: blas_round indexes S.tris without looking at the kind)
    int64_t leaf_mesh = -1;
    uint32_t max_deferred_in_leaf = 0;  // most mesh / instance / volume references in one top-level leaf (each takes a stack slot)

    static bool tie_is_sphere(const pt_scene_desc* d, pt_ref r) {
        if (r.kind == PT_PRIM_SPHERE) return true;
        if (r.kind == PT_OBJ_INSTANCE) return d->instances[r.index].child.kind == PT_PRIM_SPHERE;
        return false;
    }
    // Conservative f64 bounds of one reference (own restatement of the geometry, independent of the host's padded boxes).
    static void grow(double lo[3], double hi[3], const pt_vec3& p) {
        lo[0] = std::min(lo[0], p.x); lo[1] = std::min(lo[1], p.y); lo[2] = std::min(lo[2], p.z);
        hi[0] = std::max(hi[0], p.x); hi[1] = std::max(hi[1], p.y); hi[2] = std::max(hi[2], p.z);
    }
    bool ref_box(pt_ref r, double lo[3], double hi[3]) const {
        for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
        auto add = [](const pt_vec3& a, const pt_vec3& b) { return pt_vec3{a.x + b.x, a.y + b.y, a.z + b.z}; };
        switch (r.kind) {
            case PT_PRIM_SPHERE: {
                const pt_sphere& s = d->spheres[r.index];
                double rad = std::fabs(s.radius) * (1.0 + 1e-12);
                for (const pt_vec3* c : {&s.position1, &s.position2}) { grow(lo, hi, pt_vec3{c->x - rad, c->y - rad, c->z - rad}); grow(lo, hi, pt_vec3{c->x + rad, c->y + rad, c->z + rad}); }
                return true;
            }
            case PT_PRIM_QUAD: {
                const pt_quad& q = d->quads[r.index];
                grow(lo, hi, q.q); grow(lo, hi, add(q.q, q.u)); grow(lo, hi, add(q.q, q.v)); grow(lo, hi, add(add(q.q, q.u), q.v));
                return true;
            }
            case PT_PRIM_TRIANGLE: {
                const pt_triangle& t = d->triangles[r.index];
                grow(lo, hi, t.v0); grow(lo, hi, t.v1); grow(lo, hi, t.v2);
                return true;
            }
            case PT_OBJ_CUBOID: {
                for (uint32_t k = 0; k < 6; k++) {
                    double l2[3], h2[3];
                    ref_box(pt_ref{PT_PRIM_QUAD, d->cuboids[r.index].first_quad + k}, l2, h2);
                    grow(lo, hi, pt_vec3{l2[0], l2[1], l2[2]}); grow(lo, hi, pt_vec3{h2[0], h2[1], h2[2]});
                }
                return true;
            }
            case PT_OBJ_MESH: {
                const pt_mesh& m = d->meshes[r.index];
                for (uint32_t k = 0; k < m.n_triangles; k++) { const pt_triangle& t = d->triangles[m.first_triangle + k]; grow(lo, hi, t.v0); grow(lo, hi, t.v1); grow(lo, hi, t.v2); }
                return m.n_triangles > 0;
            }
            case PT_OBJ_INSTANCE: {
                const pt_instance& in = d->instances[r.index];
                double l2[3], h2[3];
                if (!ref_box(in.child, l2, h2)) return false;
                double ext = 0.0;
                for (int c = 0; c < 8; c++) {  // 8 corners through the rigid transform
                    double p[3] = {(c & 1) ? h2[0] : l2[0], (c & 2) ? h2[1] : l2[1], (c & 4) ? h2[2] : l2[2]};
                    pt_vec3 w{in.transform[0] * p[0] + in.transform[4] * p[1] + in.transform[8] * p[2] + in.transform[12],
                              in.transform[1] * p[0] + in.transform[5] * p[1] + in.transform[9] * p[2] + in.transform[13],
                              in.transform[2] * p[0] + in.transform[6] * p[1] + in.transform[10] * p[2] + in.transform[14]};
                    grow(lo, hi, w);
                    ext = std::max({ext, std::fabs(p[0]), std::fabs(p[1]), std::fabs(p[2])});
                }
                double slack = 1e-9 * (ext + 1.0);  // f64 rounding of the forward/inverse transforms
                for (int k = 0; k < 3; k++) { lo[k] -= slack; hi[k] += slack; }
                return true;
            }
            case PT_OBJ_VOLUME: return ref_box(d->volumes[r.index].boundary, lo, hi);  // a scatter point lies inside the boundary
        }
        return false;
    }
    void set_box(DNode& n, const double* lo, const double* hi) {
        double m = 0.0;
        for (int k = 0; k < 3; k++) { m = std::max(m, std::fabs(lo[k])); m = std::max(m, std::fabs(hi[k])); }
        if (!std::isfinite(m)) { for (int k = 0; k < 3; k++) { n.lo[k] = -INFINITY; n.hi[k] = INFINITY; } return; }
        double pad = m * 9.5367431640625e-07;  // 8 ulp(fp32) of the largest coordinate: inv/product rounding in the slab test
        for (int k = 0; k < 3; k++) { n.lo[k] = round_down(lo[k] - pad); n.hi[k] = round_up(hi[k] + pad); }
    }
    // leaf refs of host node `hn` -> device refs (appended), with exact-tie ranks in DFS order
    bool make_leaf(DNode& out, const pt_ref* items, uint32_t n, uint32_t top_bit, bool box_inf, const double* lo, const double* hi) {
        uint32_t deferred = 0;
        for (uint32_t p = 0; p < n; p++) {
            if (leaf_mesh >= 0) {
                const pt_mesh& m = d->meshes[leaf_mesh];
                if (items[p].kind != PT_PRIM_TRIANGLE || items[p].index < m.first_triangle || items[p].index - m.first_triangle >= m.n_triangles) {
                    err = "a mesh BVH leaf may hold only the mesh's own triangles"; return false;
                }
            } else {
                if (items[p].kind == PT_PRIM_TRIANGLE) { err = "a bare triangle cannot be a top-level object"; return false; }
                deferred += items[p].kind >= PT_OBJ_MESH;
            }
        }
        max_deferred_in_leaf = std::max(max_deferred_in_leaf, deferred);
        out.a = (uint32_t)refs.size(); out.b = n;
        if (box_inf) { for (int k = 0; k < 3; k++) { out.lo[k] = -INFINITY; out.hi[k] = INFINITY; } } else set_box(out, lo, hi);
        uint32_t base = tie_counter; tie_counter += 2 * n + 2;
        for (uint32_t p = 0; p < n; p++) {
            // later non-sphere wins a tie; a later sphere loses it
            uint32_t rank = tie_is_sphere(d, items[p]) ? base + n - 1 - p : base + n + p;
            DNode rn{};
            double lo[3], hi[3];
            if (ref_box(items[p], lo, hi)) set_box(rn, lo, hi); else { for (int k = 0; k < 3; k++) { rn.lo[k] = -INFINITY; rn.hi[k] = INFINITY; } }
            rn.a = ref_pack(items[p].kind, items[p].index); rn.b = rank | top_bit;
            refs.push_back(rn);
        }
        return true;
    }
    // writes host node hn into nodes[slot]; children become a fresh adjacent pair
    bool fill(uint32_t slot, uint32_t hn, uint32_t top_bit, uint32_t depth) {
        if (hn >= d->n_nodes) { err = "bvh node index out of range"; return false; }
        if (depth > 200) { err = "bvh too deep / cyclic"; return false; }
        max_depth = std::max(max_depth, depth);
        const pt_bvh_node& h = d->nodes[hn];
        if (h.left == PT_NONE) {
            if ((uint64_t)h.first_ref + h.n_refs > d->n_leaf_refs) { err = "leaf refs out of range"; return false; }
            DNode n{};
            if (!make_leaf(n, d->leaf_refs + h.first_ref, h.n_refs, top_bit, false, h.bmin, h.bmax)) return false;
            nodes[slot] = n;
            return true;
        }
        uint32_t pair = (uint32_t)nodes.size();
        nodes.resize(nodes.size() + 2);
        DNode n{}; set_box(n, h.bmin, h.bmax); n.a = pair; n.b = kNone;
        nodes[slot] = n;
        return fill(pair, h.left, top_bit, depth + 1) && fill(pair + 1, h.right, top_bit, depth + 1);  // DFS: left before right
    }
    void dummy(uint32_t slot) { DNode n{}; for (int k = 0; k < 3; k++) { n.lo[k] = INFINITY; n.hi[k] = -INFINITY; } n.a = 0; n.b = 0; nodes[slot] = n; }

    // ---- mesh-walk layout (device_scene.cuh: DWide2): 4-wide nodes whose children are nodes or single triangles
    std::vector<DWide2> wide2;
    std::vector<uint32_t> tri_rank;
    uint32_t wide2_depth = 0;
    struct Item { bool tri; uint32_t id; };  // tri: index into refs (a triangle reference); else a binary node slot
    const DNode& item_box(const Item& it) const { return it.tri ? refs[it.id] : nodes[it.id]; }
    // No heap traffic per node: an item list never grows beyond four while leaves and nodes are opened (the loop only opens what
    // fits), and the one case with more than four items — the triangles of an oversized leaf — is a contiguous run of refs.
    // (pt_scene_create runs inside the timed region of an end-to-end step: scene 6's 30 000 nodes took 4.3 ms with std::vector items.)
    static void empty_wide2(DWide2& w) {
        for (int i = 0; i < 4; i++) { for (int k = 0; k < 3; k++) { w.lo[k][i] = INFINITY; w.hi[k][i] = -INFINITY; } w.child[i] = kNone; w.pad[i] = 0; }
    }
    // a node over the triangle references refs[first .. first + count): up to four as direct children, more as four runs
    uint32_t make_wide2_tris(uint32_t first, uint32_t count, uint32_t depth) {
        wide2_depth = std::max(wide2_depth, depth);
        const uint32_t wi = (uint32_t)wide2.size();
        wide2.emplace_back();
        DWide2 w{}; empty_wide2(w);
        const uint32_t n_slots = std::min(count, 4u);
        for (uint32_t g = 0; g < n_slots; g++) {
            const uint32_t b0 = count <= 4 ? g : (uint32_t)((uint64_t)count * g / 4), b1 = count <= 4 ? g + 1 : (uint32_t)((uint64_t)count * (g + 1) / 4);
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (uint32_t k = b0; k < b1; k++) { const DNode& b = refs[first + k]; for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
            for (int a = 0; a < 3; a++) { w.lo[a][g] = lo[a]; w.hi[a][g] = hi[a]; }
            w.child[g] = b1 - b0 == 1 ? (kTriBit | ref_index(refs[first + b0].a)) : make_wide2_tris(first + b0, b1 - b0, depth + 1);
        }
        wide2[wi] = w;
        return wi;
    }
    uint32_t make_wide2(const Item* in, size_t n_in, uint32_t depth) {
        wide2_depth = std::max(wide2_depth, depth);
        Item items[4]; size_t n = 0;
        for (; n < n_in && n < 4; n++) items[n] = in[n];
        // open the largest openable item until four are held: an internal node into its two children, a leaf into its
        // triangles (when they fit)
        while (n < 4) {
            int best = -1; float best_area = -1.f;
            for (size_t i = 0; i < n; i++) {
                if (items[i].tri) continue;
                const DNode& nd = nodes[items[i].id];
                const size_t grow = nd.b == kNone ? 2 : nd.b;
                if (n - 1 + grow > 4 || (nd.b != kNone && nd.b == 0)) continue;
                if (half_area(nd) > best_area) { best = (int)i; best_area = half_area(nd); }
            }
            if (best < 0) break;
            const DNode nd = nodes[items[best].id];
            const size_t grow = nd.b == kNone ? 2 : nd.b;
            for (size_t i = n; i-- > (size_t)best + 1;) items[i + grow - 1] = items[i];  // make room in place, order kept
            for (size_t k = 0; k < grow; k++) items[best + k] = nd.b == kNone ? Item{false, nd.a + (uint32_t)k} : Item{true, nd.a + (uint32_t)k};
            n += grow - 1;
        }
        const uint32_t wi = (uint32_t)wide2.size();
        wide2.emplace_back();
        DWide2 w{}; empty_wide2(w);
        for (size_t i = 0; i < n; i++) {
            const DNode& b = item_box(items[i]);
            for (int k = 0; k < 3; k++) { w.lo[k][i] = b.lo[k]; w.hi[k][i] = b.hi[k]; }
            if (items[i].tri) { w.child[i] = kTriBit | ref_index(refs[items[i].id].a); continue; }
            const DNode nd = nodes[items[i].id];
            if (nd.b == kNone) { const Item ch[2] = {Item{false, nd.a}, Item{false, nd.a + 1}}; w.child[i] = make_wide2(ch, 2, depth + 1); }
            else if (nd.b == 0) { w.child[i] = kNone; for (int k = 0; k < 3; k++) { w.lo[k][i] = INFINITY; w.hi[k][i] = -INFINITY; } }
            else w.child[i] = make_wide2_tris(nd.a, nd.b, depth + 1);  // a leaf that did not fit: its triangles become a node of their own
        }
        wide2[wi] = w;
        return wi;
    }

    // ---- collapse of the binary tree into 4-wide nodes (device_scene.cuh: DWide)
    std::vector<DWide> wide;
    uint32_t wide_depth = 0;
    static float half_area(const DNode& n) {
        float ex = n.hi[0] - n.lo[0], ey = n.hi[1] - n.lo[1], ez = n.hi[2] - n.lo[2];
        return ex * ey + ex * ez + ey * ez;
    }
    // children: binary node slots.  Internal children are opened, largest box first, until four are held.
    uint32_t make_wide(const uint32_t* in, size_t n_in, uint32_t depth) {
        wide_depth = std::max(wide_depth, depth);
        uint32_t children[4]; size_t n = 0;
        for (size_t i = 0; i < n_in && n < 4; i++) if (!(nodes[in[i]].b != kNone && nodes[in[i]].b == 0)) children[n++] = in[i];  // drop empty dummy leaves
        while (n < 4) {
            int best = -1; float best_area = -1.f;
            for (size_t i = 0; i < n; i++)
                if (nodes[children[i]].b == kNone && half_area(nodes[children[i]]) > best_area) { best = (int)i; best_area = half_area(nodes[children[i]]); }
            if (best < 0) break;
            const uint32_t pair = nodes[children[best]].a;
            for (size_t i = n; i-- > (size_t)best + 1;) children[i + 1] = children[i];
            children[best] = pair; children[best + 1] = pair + 1;
            n++;
        }
        const uint32_t wi = (uint32_t)wide.size();
        wide.emplace_back();
        DWide w{};
        for (int i = 0; i < 4; i++) {
            for (int k = 0; k < 3; k++) { w.lo[k][i] = INFINITY; w.hi[k][i] = -INFINITY; }
            w.child[i] = kNone; w.pad[i] = 0;
        }
        for (size_t i = 0; i < n; i++) {
            const DNode nd = nodes[children[i]];
            for (int k = 0; k < 3; k++) { w.lo[k][i] = nd.lo[k]; w.hi[k][i] = nd.hi[k]; }
            const uint32_t ch[2] = {nd.a, nd.a + 1};
            w.child[i] = nd.b == kNone ? (0x20000000u | make_wide(ch, 2, depth + 1)) : (0xC0000000u | children[i]);
        }
        wide[wi] = w;
        return wi;
    }
};
}  // namespace

static int check_desc(const pt_scene_desc* d) {
    if (!d) return fail(PT_ERR_INVALID, "null scene description");
    if (d->abi_version != PT_ABI_VERSION) return fail(PT_ERR_INVALID, "pt_scene_desc.abi_version mismatch");
    auto need = [](uint32_t n, const void* p) { return n == 0 || p != nullptr; };
    if (!need(d->n_textures, d->textures) || !need(d->n_images, d->images) || !need(d->n_materials, d->materials) ||
        !need(d->n_spheres, d->spheres) || !need(d->n_quads, d->quads) || !need(d->n_triangles, d->triangles) ||
        !need(d->n_cuboids, d->cuboids) || !need(d->n_meshes, d->meshes) || !need(d->n_instances, d->instances) ||
        !need(d->n_nodes, d->nodes) || !need(d->n_leaf_refs, d->leaf_refs) || !need(d->n_objects, d->objects) || !need(d->n_lights, d->lights))
        return fail(PT_ERR_INVALID, "scene description has a null array with a non-zero count");
    if (d->n_triangles >= (1u << 29) || d->n_quads >= (1u << 29) || d->n_spheres >= (1u << 29)) return fail(PT_ERR_UNSUPPORTED, "too many primitives");
    auto ref_ok = [&](pt_ref r, bool top) {
        switch (r.kind) {
            case PT_PRIM_SPHERE: return r.index < d->n_spheres;
            case PT_PRIM_QUAD: return r.index < d->n_quads;
            case PT_PRIM_TRIANGLE: return !top && r.index < d->n_triangles;
            case PT_OBJ_CUBOID: return r.index < d->n_cuboids;
            case PT_OBJ_MESH: return r.index < d->n_meshes;
            case PT_OBJ_INSTANCE: return top && r.index < d->n_instances;
            case PT_OBJ_VOLUME: return r.index < d->n_volumes;
            default: return false;
        }
    };
    if (d->n_volumes && !d->volumes) return fail(PT_ERR_INVALID, "scene description has a null array with a non-zero count");
    // everything the uploader dereferences through an index is range-checked here, before any of it is touched
    for (uint32_t i = 0; i < d->n_spheres; i++) if (d->spheres[i].material >= d->n_materials) return fail(PT_ERR_INVALID, "bad sphere material");
    for (uint32_t i = 0; i < d->n_quads; i++) if (d->quads[i].material >= d->n_materials) return fail(PT_ERR_INVALID, "bad quad material");
    for (uint32_t i = 0; i < d->n_cuboids; i++)
        if ((uint64_t)d->cuboids[i].first_quad + 6 > d->n_quads) return fail(PT_ERR_INVALID, "bad cuboid: first_quad + 6 exceeds n_quads");
    for (uint32_t i = 0; i < d->n_meshes; i++) {
        const pt_mesh& m = d->meshes[i];
        if ((uint64_t)m.first_triangle + m.n_triangles > d->n_triangles || m.material >= d->n_materials) return fail(PT_ERR_INVALID, "bad mesh: triangle range or material");
        if ((m.has_normals && !d->tri_normals) || (m.has_uvs && !d->tri_uvs)) return fail(PT_ERR_INVALID, "mesh flags without arrays");
        if (m.bvh_root != PT_NONE && m.bvh_root >= d->n_nodes) return fail(PT_ERR_INVALID, "bad mesh: bvh_root out of range");
    }
    for (uint32_t i = 0; i < d->n_nodes; i++) {
        const pt_bvh_node& nd = d->nodes[i];
        if (nd.left == PT_NONE ? (uint64_t)nd.first_ref + nd.n_refs > d->n_leaf_refs : (nd.left >= d->n_nodes || nd.right >= d->n_nodes))
            return fail(PT_ERR_INVALID, "bad bvh node " + std::to_string(i));
    }
    for (uint32_t i = 0; i < d->n_volumes; i++) {
        const pt_volume& v = d->volumes[i];
        if ((v.boundary.kind != PT_PRIM_SPHERE && v.boundary.kind != PT_OBJ_CUBOID) || !ref_ok(v.boundary, false))
            return fail(PT_ERR_UNSUPPORTED, "volume boundary must be a sphere or a cuboid");
        if (!(v.density > 0.0) || v.material >= d->n_materials || d->materials[v.material].kind != PT_MAT_ISOTROPIC)
            return fail(PT_ERR_INVALID, "volume needs a positive density and a PT_MAT_ISOTROPIC phase function");
    }
    for (uint32_t i = 0; i < d->n_lights; i++) {
        const pt_ref r = d->lights[i].kind == PT_OBJ_INSTANCE && d->lights[i].index < d->n_instances ? d->instances[d->lights[i].index].child : d->lights[i];
        if (r.kind == PT_OBJ_VOLUME) return fail(PT_ERR_UNSUPPORTED, "a volume cannot be a light");
    }
    for (uint32_t i = 0; i < d->n_objects; i++) if (!ref_ok(d->objects[i], true)) return fail(PT_ERR_INVALID, "bad ref in objects");
    for (uint32_t i = 0; i < d->n_lights; i++) if (!ref_ok(d->lights[i], true)) return fail(PT_ERR_INVALID, "bad ref in lights");
    for (uint32_t i = 0; i < d->n_leaf_refs; i++) if (!ref_ok(d->leaf_refs[i], false) && !ref_ok(d->leaf_refs[i], true)) return fail(PT_ERR_INVALID, "bad leaf ref");
    for (uint32_t i = 0; i < d->n_instances; i++) {
        pt_ref c = d->instances[i].child;
        if (c.kind == PT_OBJ_INSTANCE || c.kind == PT_PRIM_TRIANGLE || !ref_ok(c, false)) return fail(PT_ERR_UNSUPPORTED, "instance child must be a sphere, quad, cuboid or mesh");
    }
    for (uint32_t i = 0; i < d->n_materials; i++) {
        const pt_material& m = d->materials[i];
        auto tex_ok = [&](uint32_t t) { return t < d->n_textures; };
        bool ok = true;
        switch (m.kind) {
            case PT_MAT_DIFFUSE: ok = tex_ok(m.base_color_tex) && (m.normal_map == PT_NONE || m.normal_map < d->n_images); break;
            case PT_MAT_METAL: case PT_MAT_GLASS: ok = tex_ok(m.base_color_tex) && tex_ok(m.roughness_tex); break;
            case PT_MAT_PRINCIPLED: case PT_MAT_LIGHT: case PT_MAT_ISOTROPIC: ok = tex_ok(m.base_color_tex); break;
            case PT_MAT_SHEEN: case PT_MAT_CLEARCOAT: break;
            case PT_MAT_MIX: ok = m.mix_a < i && m.mix_b < i; break;
            default: ok = false;
        }
        if (!ok) return fail(PT_ERR_INVALID, "bad material " + std::to_string(i));
    }
    for (uint32_t i = 0; i < d->n_textures; i++) {
        const pt_texture& t = d->textures[i];
        if (t.kind == PT_TEX_CHECKER && (t.tex1 >= d->n_textures || t.tex2 >= d->n_textures)) return fail(PT_ERR_INVALID, "bad checker child");
        if (t.kind == PT_TEX_IMAGE && t.image >= d->n_images) return fail(PT_ERR_INVALID, "bad image index");
        if (t.kind > PT_TEX_IMAGE) return fail(PT_ERR_INVALID, "bad texture kind");
    }
    return PT_OK;
}

extern "C" {

int pt_scene_create(pt_ctx* ctx, const pt_scene_desc* d, pt_scene** out) {
    if (!ctx || !out) return fail(PT_ERR_INVALID, "pt_scene_create: null argument");
    int rc = check_desc(d);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    auto s = std::unique_ptr<pt_scene>(new pt_scene());
    s->ctx = ctx;
    Uploader U{s.get()};
    Converter C; C.d = d;
    const bool timing = getenv("PT_B200_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t0 = now();
    auto lap = [&](const char* what) { if (timing) { auto t1 = now(); fprintf(stderr, "[pt_scene_create] %-28s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count()); t0 = t1; } };

    // ---- BVH: pair 0 = (objects root, lights root); every mesh gets its own (root, dummy) pair
    C.nodes.reserve(2ull * d->n_nodes + 2ull * d->n_meshes + 8); C.refs.reserve((size_t)d->n_leaf_refs + d->n_objects + d->n_lights + d->n_triangles);
    C.wide.reserve(d->n_nodes / 2 + 8); C.wide2.reserve((size_t)d->n_nodes + d->n_triangles / 2 + 8);
    C.nodes.resize(2);
    auto top_list = [&](uint32_t slot, uint32_t root, const pt_ref* items, uint32_t n, uint32_t top_bit) -> bool {
        if (n == 0) { C.dummy(slot); return true; }
        if (root == PT_NONE) { DNode leaf{}; if (!C.make_leaf(leaf, items, n, top_bit, true, nullptr, nullptr)) return false; C.nodes[slot] = leaf; return true; }  // list.rs:57-66
        return C.fill(slot, root, top_bit, 1);
    };
    // lights first so that their ranks are lower; objects additionally carry bit 31 (object beats light, world.rs:55-59)
    if (!top_list(1, d->lights_bvh_root, d->lights, d->n_lights, 0u)) return fail(PT_ERR_INVALID, C.err);
    if (!top_list(0, d->objects_bvh_root, d->objects, d->n_objects, 0x80000000u)) return fail(PT_ERR_INVALID, C.err);
    const uint32_t tlas_depth = C.max_depth;
    const uint32_t n_top = (uint32_t)C.refs.size();  // refs[0 .. n_top) are the top-level references: lights, then objects, in leaf order
    if (C.tie_counter >= 0x7FFFFFFFu) return fail(PT_ERR_UNSUPPORTED, "too many top-level objects");
    // one traversal flavour per scene (compile-time in the kernels): 4-wide nodes as soon as any BVH is large
    bool use_wide = d->n_objects + d->n_lights >= kWideMinItems;
    for (uint32_t mi = 0; mi < d->n_meshes; mi++) use_wide |= d->meshes[mi].n_triangles >= kWideMinItems;
    s->wide = use_wide;
    std::vector<DMesh> meshes(d->n_meshes);
    std::vector<uint32_t> tri_mesh(d->n_triangles, 0);
    uint32_t blas_depth = 0;
    for (uint32_t mi = 0; mi < d->n_meshes; mi++) {
        const pt_mesh& m = d->meshes[mi];
        for (uint32_t k = 0; k < m.n_triangles; k++) tri_mesh[m.first_triangle + k] = mi;
        C.leaf_mesh = mi;
        uint32_t pair = (uint32_t)C.nodes.size();
        C.nodes.resize(C.nodes.size() + 2);
        C.dummy(pair + 1);
        C.max_depth = 0; C.tie_counter = 0;
        if (m.bvh_root == PT_NONE) {
            std::vector<pt_ref> items(m.n_triangles);
            for (uint32_t k = 0; k < m.n_triangles; k++) items[k] = pt_ref{PT_PRIM_TRIANGLE, m.first_triangle + k};
            DNode leaf{};
            if (!C.make_leaf(leaf, items.data(), m.n_triangles, 0u, true, nullptr, nullptr)) return fail(PT_ERR_INVALID, C.err);
            C.nodes[pair] = leaf;
        } else if (!C.fill(pair, m.bvh_root, 0u, 1)) return fail(PT_ERR_INVALID, C.err);
        // large BVHs are collapsed to 4-wide nodes; small ones keep the binary pairs (cheaper when most rays leave at once)
        uint32_t root_entry = pair, depth_slots = C.max_depth;
        if (use_wide) { C.wide_depth = 0; root_entry = 0x20000000u | C.make_wide(&pair, 1, 1); depth_slots = 3 * C.wide_depth; }
        blas_depth = std::max(blas_depth, depth_slots);
        const Converter::Item root_item{false, pair};
        const uint32_t root2 = m.n_triangles ? C.make_wide2(&root_item, 1, 1) : 0u;
        meshes[mi] = DMesh{root_entry, m.first_triangle, m.n_triangles, m.material, m.has_normals, m.has_uvs, m.bvh_root == PT_NONE, root2};
    }
    uint32_t world_root = 0, tlas_slots = tlas_depth;
    if (use_wide) { C.wide_depth = 0; const uint32_t tops[2] = {0u, 1u}; world_root = 0x20000000u | C.make_wide(tops, 2, 1); tlas_slots = 3 * C.wide_depth; }
    // stack bound: one pending sibling per binary level / three per wide level + the mesh / instance / volume references one
    // top-level leaf can defer (counted while converting: a failed SAH split leaves a leaf of any size, bvh.rs:37-42) + sentinel
    s->max_stack = tlas_slots + blas_depth + 2 + std::max(C.max_deferred_in_leaf, 8u);
    if (s->max_stack > (uint32_t)kStack) return fail(PT_ERR_UNSUPPORTED, "BVH too deep for the device traversal stack (" + std::to_string(s->max_stack) + " > " + std::to_string(kStack) + ")");

    lap("BVH conversion");
    // ---- primitives
    std::vector<DSphere> spheres(d->n_spheres);
    for (uint32_t i = 0; i < d->n_spheres; i++) {
        const pt_sphere& p = d->spheres[i];
        DSphere o{}; o.p1[0] = p.position1.x; o.p1[1] = p.position1.y; o.p1[2] = p.position1.z; o.p2[0] = p.position2.x; o.p2[1] = p.position2.y; o.p2[2] = p.position2.z;
        o.radius = p.radius; o.material = p.material; spheres[i] = o;
    }
    std::vector<DQuad> quads(d->n_quads); std::vector<uint32_t> quad_mat(d->n_quads);
    for (uint32_t i = 0; i < d->n_quads; i++) {
        const pt_quad& p = d->quads[i];
        DQuad o{}; const pt_vec3* src[5] = {&p.q, &p.u, &p.v, &p.w, &p.normal}; double* dst[5] = {o.q, o.u, o.v, o.w, o.n};
        for (int k = 0; k < 5; k++) { dst[k][0] = src[k]->x; dst[k][1] = src[k]->y; dst[k][2] = src[k]->z; }
        const d3 un = normalize(mk(o.n[0], o.n[1], o.n[2]));
        o.un[0] = un.x; o.un[1] = un.y; o.un[2] = un.z;
        o.area = length(cross(mk(o.u[0], o.u[1], o.u[2]), mk(o.v[0], o.v[1], o.v[2])));
        o.d = p.d; quads[i] = o; quad_mat[i] = p.material;
    }
    std::vector<DTri> tris(d->n_triangles);
    for (uint32_t i = 0; i < d->n_triangles; i++) {
        const pt_triangle& p = d->triangles[i];
        DTri o{}; o.v0[0] = p.v0.x; o.v0[1] = p.v0.y; o.v0[2] = p.v0.z;
        o.e1[0] = p.v1.x - p.v0.x; o.e1[1] = p.v1.y - p.v0.y; o.e1[2] = p.v1.z - p.v0.z;  // mesh.rs:55-56, same IEEE subtraction
        o.e2[0] = p.v2.x - p.v0.x; o.e2[1] = p.v2.y - p.v0.y; o.e2[2] = p.v2.z - p.v0.z;
        tris[i] = o;
    }
    std::vector<double> tri_normals, tri_uvs;
    bool any_n = false, any_uv = false;
    for (auto& m : meshes) { any_n |= m.has_normals != 0; any_uv |= m.has_uvs != 0; }
    if (any_n) { tri_normals.resize(9ull * d->n_triangles); memcpy(tri_normals.data(), d->tri_normals, tri_normals.size() * 8); }
    if (any_uv) { tri_uvs.assign(d->tri_uvs, d->tri_uvs + 6ull * d->n_triangles); }
    std::vector<DCuboid> cuboids(d->n_cuboids);
    for (uint32_t i = 0; i < d->n_cuboids; i++) cuboids[i] = DCuboid{d->cuboids[i].first_quad, d->cuboids[i].material};  // ranges: check_desc
    std::vector<DInstance> instances(d->n_instances);
    for (uint32_t i = 0; i < d->n_instances; i++) {
        const pt_instance& p = d->instances[i];
        DInstance o{};
        for (int c = 0; c < 4; c++) for (int r = 0; r < 3; r++) { o.inv[c * 3 + r] = p.inverse[c * 4 + r]; o.fwd[c * 3 + r] = p.transform[c * 4 + r]; o.nrm[c * 3 + r] = p.normal_matrix[c * 4 + r]; }
        o.child_kind = p.child.kind; o.child_index = p.child.index; o.tie_is_sphere = p.child.kind == PT_PRIM_SPHERE;
        instances[i] = o;
    }
    std::vector<DVolume> volumes(d->n_volumes);
    for (uint32_t i = 0; i < d->n_volumes; i++)
        volumes[i] = DVolume{d->volumes[i].boundary.kind, d->volumes[i].boundary.index, d->volumes[i].material, 0, -1.0 / d->volumes[i].density, 0.0};
    s->has_volumes = d->n_volumes > 0;
    s->defer_meshes = s->wide && d->n_meshes > 0 && !s->has_volumes;
    lap("primitives");
    // ---- textures, images, materials
    std::vector<DTexture> textures(d->n_textures);
    for (uint32_t i = 0; i < d->n_textures; i++) {
        const pt_texture& t = d->textures[i];
        textures[i] = DTexture{t.kind, t.tex1, t.tex2, t.image, t.inv_scale, {t.value.x, t.value.y, t.value.z}};
    }
    std::vector<DImage> images(d->n_images);
    uint64_t img_bytes = 0;
    for (uint32_t i = 0; i < d->n_images; i++) {
        if (!d->images[i].rgb && d->images[i].width * d->images[i].height) return fail(PT_ERR_INVALID, "null image data");
        images[i] = DImage{img_bytes, d->images[i].width, d->images[i].height};
        img_bytes += ((3ull * d->images[i].width * d->images[i].height + 255) / 256) * 256;
    }
    uint8_t* d_img = nullptr;
    if (img_bytes) {
        std::pair<void*, size_t> blk{nullptr, 0};
        if (ctx->alloc_block(img_bytes, &blk)) return fail(PT_ERR_CUDA, "cudaMalloc failed for the image block");
        s->allocs.push_back(blk);
        d_img = (uint8_t*)blk.first;
        for (uint32_t i = 0; i < d->n_images; i++) {
            size_t nb = 3ull * d->images[i].width * d->images[i].height;
            if (nb) CU(cudaMemcpyAsync(d_img + images[i].offset, d->images[i].rgb, nb, cudaMemcpyHostToDevice, ctx->stream));
            s->bytes += nb;
        }
    }
    std::vector<DMaterial> materials(d->n_materials);
    for (uint32_t i = 0; i < d->n_materials; i++) {
        const pt_material& m = d->materials[i];
        DMaterial o{}; o.kind = m.kind; o.base_color_tex = m.base_color_tex; o.roughness_tex = m.roughness_tex; o.normal_map = m.normal_map;
        o.mix_a = m.mix_a; o.mix_b = m.mix_b; memcpy(o.p, m.p, sizeof(o.p));
        auto tex_uses_uv = [&](uint32_t t) {  // an image texture anywhere below (checker children included) reads (u, v)
            std::vector<uint32_t> todo; if (t != PT_NONE) todo.push_back(t);
            for (size_t guard = 0; !todo.empty() && guard < 4096; guard++) {
                const pt_texture& tx = d->textures[todo.back()]; todo.pop_back();
                if (tx.kind == PT_TEX_IMAGE) return true;
                if (tx.kind == PT_TEX_CHECKER) { todo.push_back(tx.tex1); todo.push_back(tx.tex2); }
            }
            return !todo.empty();  // guard tripped (cyclic checker): be conservative
        };
        const bool has_color = m.kind <= PT_MAT_LIGHT, has_rough = m.kind == PT_MAT_METAL || m.kind == PT_MAT_GLASS;
        o.uses_uv = (m.kind == PT_MAT_DIFFUSE && m.normal_map != PT_NONE) || (has_color && tex_uses_uv(m.base_color_tex)) ||
                    (has_rough && tex_uses_uv(m.roughness_tex)) ||
                    (m.kind == PT_MAT_MIX && (materials[m.mix_a].uses_uv || materials[m.mix_b].uses_uv));
        materials[i] = o;
        s->class_mask |= 1u << (m.kind == PT_MAT_LIGHT ? CLS_LIGHT : m.kind == PT_MAT_DIFFUSE ? CLS_DIFFUSE : m.kind == PT_MAT_METAL ? CLS_METAL
                                : m.kind == PT_MAT_GLASS ? CLS_GLASS : m.kind == PT_MAT_PRINCIPLED ? CLS_PRINCIPLED : CLS_OTHER);
    }
    // ---- flat top level (wavefront.cuh: TopList): what k_top walks instead of a BVH when World is small
    {
        auto cls_of_mat = [&](uint32_t m) -> uint8_t {
            const uint32_t k = d->materials[m].kind;
            return (uint8_t)(k == PT_MAT_LIGHT ? CLS_LIGHT : k == PT_MAT_DIFFUSE ? CLS_DIFFUSE : k == PT_MAT_METAL ? CLS_METAL : k == PT_MAT_GLASS ? CLS_GLASS
                             : k == PT_MAT_PRINCIPLED ? CLS_PRINCIPLED : CLS_OTHER);
        };
        auto mat_of = [&](uint32_t kind, uint32_t index) -> uint32_t {
            switch (kind) {
                case PT_PRIM_SPHERE: return d->spheres[index].material;
                case PT_PRIM_QUAD: return d->quads[index].material;
                case PT_OBJ_CUBOID: return d->quads[d->cuboids[index].first_quad].material;
                case PT_OBJ_MESH: return d->meshes[index].material;
                default: return 0;
            }
        };
        TopList T{};
        bool ok = !s->has_volumes && n_top >= 1 && n_top <= (uint32_t)kTopMax && 3 * C.wide2_depth + 4 <= (uint32_t)kStack2;
        for (uint32_t k = 0; ok && k < n_top; k++) {
            uint32_t kind = ref_kind(C.refs[k].a), index = ref_index(C.refs[k].a);
            if (kind == PT_OBJ_INSTANCE) { const pt_ref c = d->instances[index].child; kind = c.kind; index = c.index; }
            if (kind == PT_OBJ_VOLUME || kind == PT_PRIM_TRIANGLE) { ok = false; break; }
            if (kind == PT_OBJ_MESH) { T.mesh_bits |= 1u << k; if (d->meshes[index].n_triangles == 0) T.mesh_bits &= ~(1u << k); }
            if (ref_kind(C.refs[k].a) == PT_PRIM_QUAD) T.quad_bits |= 1u << k;
            if (ref_kind(C.refs[k].a) == PT_PRIM_SPHERE) T.sphere_bits |= 1u << k;
            T.cls[k] = cls_of_mat(mat_of(kind, index));
            for (int a = 0; a < 3; a++) { T.box[k][a] = C.refs[k].lo[a]; T.box[k][3 + a] = C.refs[k].hi[a]; }
        }
        T.n = n_top;
        s->mesh_rounds = (uint32_t)__builtin_popcount(T.mesh_bits);
        s->flat = ok && s->mesh_rounds <= (uint32_t)kMeshRounds;
        s->top = T;
    }
    std::vector<float> quad_box(6ull * d->n_quads);
    for (uint32_t i = 0; i < d->n_quads; i++) {
        double lo[3], hi[3]; DNode bn{};
        C.ref_box(pt_ref{PT_PRIM_QUAD, i}, lo, hi); C.set_box(bn, lo, hi);
        for (int k = 0; k < 3; k++) { quad_box[6ull * i + k] = bn.lo[k]; quad_box[6ull * i + 3 + k] = bn.hi[k]; }
    }
    std::vector<uint32_t> tri_rank(d->n_triangles, 0u);
    for (const DNode& rn : C.refs) if (ref_kind(rn.a) == PT_PRIM_TRIANGLE) tri_rank[ref_index(rn.a)] = rn.b;

    std::vector<DRef> lights(d->n_lights);
    bool mesh_light = false;  // Triangle::sample needs the original vertices (mesh.rs:126), which the traversal layout does not keep
    for (uint32_t i = 0; i < d->n_lights; i++) {
        lights[i] = DRef{ref_pack(d->lights[i].kind, d->lights[i].index), 0};
        const pt_ref r = d->lights[i].kind == PT_OBJ_INSTANCE ? d->instances[d->lights[i].index].child : d->lights[i];
        mesh_light |= r.kind == PT_OBJ_MESH;
        s->general_lights |= d->lights[i].kind != PT_PRIM_QUAD && d->lights[i].kind != PT_PRIM_SPHERE;
    }
    std::vector<double> tri_verts;
    if (mesh_light) { tri_verts.resize(9ull * d->n_triangles); memcpy(tri_verts.data(), d->triangles, tri_verts.size() * sizeof(double)); }

    DScene& D = s->d;
    U.up(C.wide, &D.wide); U.up(C.nodes, &D.nodes); U.up(C.refs, &D.refs); U.up(spheres, &D.spheres); U.up(quads, &D.quads); U.up(quad_mat, &D.quad_material);
    U.up(tris, &D.tris); U.up(tri_normals, &D.tri_normals); U.up(tri_uvs, &D.tri_uvs); U.up(tri_mesh, &D.tri_mesh); U.up(cuboids, &D.cuboids);
    U.up(meshes, &D.meshes); U.up(instances, &D.instances); U.up(textures, &D.textures); U.up(images, &D.images); U.up(materials, &D.materials);
    U.up(lights, &D.lights); U.up(tri_verts, &D.tri_verts); U.up(volumes, &D.volumes); U.up(C.wide2, &D.wide2); U.up(tri_rank, &D.tri_rank); U.up(quad_box, &D.quad_box);
    lap("textures, top list, misc");
    if ((rc = U.commit())) return rc;
    lap("staging copy + H2D enqueue");
    D.image_data = d_img; D.n_lights = d->n_lights; D.root_entry = world_root; D.n_materials = d->n_materials; D.n_textures = d->n_textures;
    D.n_nodes = (uint32_t)C.nodes.size(); D.n_refs = (uint32_t)C.refs.size(); D.n_wide = (uint32_t)C.wide.size(); D.n_wide2 = (uint32_t)C.wide2.size();
    D.n_tris = d->n_triangles; D.n_quads = d->n_quads; D.n_spheres = d->n_spheres; D.n_instances = d->n_instances; D.n_meshes = d->n_meshes;
    s->n_materials = d->n_materials; s->n_images = d->n_images; s->images = images;
    CU(cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
    lap("stream synchronize");
    // the tail megakernel this scene's renders end in: load its code now (tail_wide.cu), not inside the first render
    if (!s->has_volumes && !s->general_lights) { if (s->flat) preload_k_tail_flat(); else if (s->wide) preload_k_tail_wide(); else preload_k_tail_bin(); }
    *out = s.release();
    return PT_OK;
}
void pt_scene_destroy(pt_scene* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);  // no kernel may still read the tables
    delete s;
}
uint64_t pt_scene_device_bytes(const pt_scene* s) { return s ? s->bytes : 0; }

// Importance sampler of a lat-long environment image (NOT reference behaviour, used only with
// PT_RENDER_ENV_IMPORTANCE).  Cell weight = sum over the cell's texels of luminance * sin(theta of the texel row), plus
// a uniform floor of 5 % of the mean so that every direction keeps a positive density; rows and columns follow
// ImageTexture::value's texel mapping (texture.rs:72-91: row 0 = theta 0).  Built on the host in f64, sequential sums.
int pt_scene_build_env_sampler(pt_scene* s, uint32_t image, uint32_t max_rows, uint32_t max_cols) {
    if (!s || image >= s->n_images) return fail(PT_ERR_INVALID, "pt_scene_build_env_sampler: bad scene or image index");
    const DImage im = s->images[image];
    if (im.width == 0 || im.height == 0) return fail(PT_ERR_INVALID, "pt_scene_build_env_sampler: empty image");
    CU(cudaSetDevice(s->ctx->device));
    const uint32_t W = im.width, H = im.height;
    const uint32_t rows = std::min<uint32_t>(H, max_rows ? max_rows : 512u), cols = std::min<uint32_t>(W, max_cols ? max_cols : 1024u);
    std::vector<uint8_t> px((size_t)W * H * 3);
    CU(cudaMemcpyAsync(px.data(), s->d.image_data + im.offset, px.size(), cudaMemcpyDeviceToHost, s->ctx->stream));
    CU(cudaStreamSynchronize(s->ctx->stream));
    std::vector<double> w((size_t)rows * cols, 0.0);
    for (uint32_t j = 0; j < H; j++) {
        const double st = sin(((double)j + 0.5) / (double)H * kPi);
        double* wr = w.data() + (size_t)((uint64_t)j * rows / H) * cols;
        const uint8_t* p = px.data() + (size_t)j * W * 3;
        for (uint32_t i = 0; i < W; i++) {
            const double lum = luminance(mk((double)p[3 * i] / 255.0, (double)p[3 * i + 1] / 255.0, (double)p[3 * i + 2] / 255.0));
            wr[(uint64_t)i * cols / W] += lum * st;
        }
    }
    double total = 0.0;
    for (double v : w) total += v;
    const double floor_w = total > 0.0 ? 0.05 * total / ((double)rows * (double)cols) : 1.0;
    std::vector<double> table((size_t)rows + 1 + (size_t)rows * (cols + 1));
    double* marginal = table.data();
    double* cond = table.data() + rows + 1;
    std::vector<double> row_sum(rows);
    double all = 0.0;
    for (uint32_t r = 0; r < rows; r++) {
        double rs = 0.0;
        for (uint32_t c = 0; c < cols; c++) { w[(size_t)r * cols + c] += floor_w; rs += w[(size_t)r * cols + c]; }
        row_sum[r] = rs; all += rs;
        double* cr = cond + (size_t)r * (cols + 1);
        cr[0] = 0.0;
        for (uint32_t c = 0; c < cols; c++) cr[c + 1] = cr[c] + w[(size_t)r * cols + c] / rs;
        cr[cols] = 1.0;
    }
    marginal[0] = 0.0;
    for (uint32_t r = 0; r < rows; r++) marginal[r + 1] = marginal[r] + row_sum[r] / all;
    marginal[rows] = 1.0;
    if (s->env_block) { CU(cudaStreamSynchronize(s->ctx->stream)); cudaFree(s->env_block); s->env_block = nullptr; s->env_image = 0xFFFFFFFFu; }
    CU(cudaMalloc(&s->env_block, table.size() * sizeof(double)));
    CU(cudaMemcpyAsync(s->env_block, table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice, s->ctx->stream));
    CU(cudaStreamSynchronize(s->ctx->stream));
    s->env = DEnvDist{(const double*)s->env_block, (const double*)s->env_block + rows + 1, rows, cols};
    s->env_image = image;
    s->bytes += table.size() * sizeof(double);
    return PT_OK;
}

// ------------------------------------------------------------------------------------------------ camera
static d3 dv(const pt_vec3& v) { return mk(v.x, v.y, v.z); }
uint32_t pt_camera_image_height(const pt_camera* c) { return c ? (uint32_t)((double)c->image_width / c->aspect_ratio) : 0; }  // camera.rs:52
static int make_camera(const pt_scene* s, const pt_camera* c, DCameraEx* out) {  // Camera::init, camera.rs:51-77
    if (!c) return fail(PT_ERR_INVALID, "null camera");
    if (c->image_width == 0 || !(c->aspect_ratio > 0.0)) return fail(PT_ERR_INVALID, "bad camera size");
    DCameraEx e{}; DCamera& k = e.c;
    k.width = c->image_width; k.height = pt_camera_image_height(c); k.max_depth = c->max_depth;
    if (k.height == 0) return fail(PT_ERR_INVALID, "camera image height is zero");
    k.center = dv(c->look_from);
    double theta = c->vfov * (kPi / 180.0);  // f64::to_radians
    double h = std::tan(theta / 2.0);
    double viewport_height = 2.0 * h * c->focal_length;
    double viewport_width = viewport_height * ((double)k.width / (double)k.height);
    d3 forward = normalize(dv(c->look_from) - dv(c->look_at));
    k.right = normalize(cross(dv(c->vup), forward));
    k.up = cross(forward, k.right);
    d3 viewport_u = k.right * viewport_width, viewport_v = k.up * -viewport_height;
    k.pixel_du = viewport_u / (double)k.width; k.pixel_dv = viewport_v / (double)k.height;
    d3 upperleft = k.center - (forward * c->focal_length) - (viewport_u / 2.0) - (viewport_v / 2.0);
    k.pixel00 = upperleft + (k.pixel_du + k.pixel_dv) * 0.5;
    k.blur_strength = c->blur_strength; k.focal_length = c->focal_length; k.defocus_angle = c->defocus_angle;
    double radius = std::tan((c->defocus_angle / 2.0) * (kPi / 180.0)) * c->focal_length;  // camera.rs:159
    e.dof_right = k.right * radius; e.dof_up = k.up * radius;
    k.env_is_map = c->env_is_map; k.env_color = dv(c->env_color); k.env_image = c->env_image;
    if (c->env_is_map && (!s || c->env_image >= s->n_images)) return fail(PT_ERR_INVALID, "camera env_image out of range");
    *out = e;
    return PT_OK;
}

// ------------------------------------------------------------------------------------------------ render
static int ensure_pool(pt_ctx* c, uint32_t paths) {
    if (c->pool >= paths) return PT_OK;
    free_pool(c);
    for (int i = 0; i < 2; i++) {
        CU(cudaMalloc(&c->pool_ray[i], (size_t)paths * sizeof(RayRec)));
        CU(cudaMalloc(&c->pool_state[i], (size_t)paths * sizeof(StateRec)));
    }
    CU(cudaMalloc(&c->hits, (size_t)paths * sizeof(HitRec)));
    CU(cudaMalloc(&c->q_items, (size_t)paths * N_CLS * sizeof(uint32_t)));
    c->pool = paths;
    return PT_OK;
}
// Buffers of the multi-pass traversals, only for scenes that take them: tie ranks of the provisional hits (8 B per path), then
// either the visit queues of k_top (8 B per path and round) + the walk records (128 B per path), or the queues of k_trace<DEFER>.
static int ensure_two_pass(pt_ctx* c, const pt_scene* scene) {
    if (!c->ties) CU(cudaMalloc(&c->ties, (size_t)c->pool * sizeof(uint2)));
    if (scene->flat && scene->mesh_rounds) {
        if (c->mq_rounds < scene->mesh_rounds) {
            cudaFree(c->mq_items); c->mq_items = nullptr; c->mq_rounds = 0;
            CU(cudaMalloc(&c->mq_items, (size_t)c->pool * scene->mesh_rounds * sizeof(uint2)));
            c->mq_rounds = scene->mesh_rounds;
        }
        if (!c->walk) CU(cudaMalloc(&c->walk, (size_t)c->pool * kWalkRecU4 * sizeof(uint4)));
    }
    if (scene->defer_meshes && !c->bq_items) CU(cudaMalloc(&c->bq_items, (size_t)c->pool * kDeferMax * sizeof(uint4)));
    return PT_OK;
}
// Path pool of `want` paths (plus the two-pass buffers when the scene needs them); on cudaErrorMemoryAllocation the pool is
// halved and the allocation retried (a smaller pool only means more wavefront iterations), down to 64 Ki paths.
static int ensure_render_buffers(pt_ctx* c, uint32_t want, const pt_scene* scene, uint32_t* got) {
    for (uint32_t pool = want;; pool = std::max((pool / 2 + kBlock - 1) / kBlock * kBlock, 1u << 16)) {
        int rc = ensure_pool(c, pool);
        if (rc == PT_OK && (scene->defer_meshes || (scene->flat && scene->mesh_rounds))) rc = ensure_two_pass(c, scene);
        if (rc == PT_OK) { *got = std::min(pool, c->pool); return PT_OK; }
        const bool oom = cudaGetLastError() == cudaErrorMemoryAllocation || g_err.find("out of memory") != std::string::npos;
        free_pool(c);
        if (!oom || pool <= (1u << 16)) return rc;
    }
}
static int ensure_scratch(pt_ctx* c, size_t bytes) {
    if (c->scratch_bytes >= bytes) return PT_OK;
    cudaFree(c->scratch); c->scratch = nullptr; c->scratch_bytes = 0;
    CU(cudaMalloc(&c->scratch, bytes));
    c->scratch_bytes = bytes;
    return PT_OK;
}
static PathBuf path_buf(pt_ctx* c, int which) {
    return PathBuf{c->pool_ray[which], c->pool_state[which]};
}

// World::intersect_all for the n (or min(n, *n_dev)) paths of `in` — THE traversal stage of a wavefront iteration, shared by
// the render loop and by pt_trace_closest_wavefront (so ray batches can be ID-compared on exactly what the render runs).
// One compile-time flavour per scene / mode.
struct TraceStage {
    pt_ctx* ctx; const pt_scene* scene; uint32_t flags; uint64_t seed; double t_min; unsigned long long* wk; pt_stats* S;
};
// `in` holds n_old live paths in slots [0, n_old); n_new more are STARTED in slots [n_old, n_old + n_new) — camera rays
// (camera.rs:153-168) generated by the top-level kernel itself on flat scenes, by k_generate otherwise.
// flat top level (k_top) + mesh kernels: every scene with a small World; with meshes not for small iterations (two more launches);
// flags 0x100000 / 0x400000 opt out (A/B measurements, and the tests that pin this path to the BVH kernels)
static bool takes_flat_path(const pt_scene* scene, uint32_t flags, uint32_t n) {
    return scene->flat && !(flags & 0x500000u) && (scene->mesh_rounds == 0 || n >= (1u << 16));
}
static void launch_trace(const TraceStage& T, const PathBuf& in, uint32_t n_old, uint32_t n_new, const GenArgs* gen, const Queues& q, const uint32_t* n_dev) {
    pt_ctx* ctx = T.ctx; const pt_scene* scene = T.scene; pt_stats& S = *T.S;
    unsigned long long* const wk = T.wk;
    cudaStream_t st = ctx->stream;
    const uint32_t n = n_old + n_new;
    if (takes_flat_path(scene, T.flags, n)) {
        uint32_t* slot = q.count - 4;
        const MeshQueues mq{ctx->mq_items, slot + 12, ctx->pool, ctx->walk, slot + 20};
        static const GenArgs no_gen{};
        ctx->mark(-1);
        if (n_old) { run_k_top(false, wk != nullptr, st, in, 0, n_old, ctx->hits, q, scene->d, scene->top, mq, ctx->ties, T.t_min, no_gen, n_dev, wk); S.kernel_launches++; ctx->mark(0); }
        if (n_new) { run_k_top(true, wk != nullptr, st, in, n_old, n_new, ctx->hits, q, scene->d, scene->top, mq, ctx->ties, T.t_min, *gen, nullptr, wk); S.kernel_launches++; ctx->mark(1); }
        const unsigned eg = std::min<unsigned>((n + kBlock - 1) / kBlock, 148u * 16u);
        const unsigned wg = std::min<unsigned>((n + kTraceBlock - 1) / kTraceBlock, mesh_walk_resident_warps());
        // rays that entered several mesh boxes (queue 1): k_mesh_multi, on a side stream next to the entry pass + walk of queue 0
        // (on the main stream when per-stage events are wanted or the iteration is small)
        const bool multi = kMeshMulti && scene->mesh_rounds >= 2;
#ifndef PT_MULTI_SIDE_MIN
#define PT_MULTI_SIDE_MIN (1u << 16)   // measured flat from 64 Ki to 1 Mi rays (profiles/r2_ab/r2_x_multi_side_threshold.log)
#endif
        const bool side = multi && !ctx->profiling && n >= PT_MULTI_SIDE_MIN;
        if (multi) {
            cudaStream_t ms = side ? ctx->shade_stream[0] : st;
            if (side) { cudaEventRecord(ctx->ev_fork, st); cudaStreamWaitEvent(ms, ctx->ev_fork, 0); }
            run_k_mesh_multi(wk != nullptr, std::min<unsigned>((n + kTraceBlock - 1) / kTraceBlock, 148u * 28u), ms, in, mq, ctx->hits, ctx->ties, q, scene->d, T.t_min, wk);
            if (side) cudaEventRecord(ctx->ev_join[0], ms); else ctx->mark(7);
            S.kernel_launches++;
        }
        for (uint32_t r = 0; r < (multi ? 1u : scene->mesh_rounds); r++) {
            run_k_mesh_enter(wk != nullptr, eg, st, in, r, mq, ctx->hits, ctx->ties, q, scene->d, scene->top, T.t_min, wk);
            ctx->mark(2);
            run_k_mesh_walk(wk != nullptr, wg, st, r, mq, ctx->hits, ctx->ties, q, scene->d, T.t_min, wk);
            ctx->mark(3);
            S.kernel_launches += 2;
        }
        if (side) cudaStreamWaitEvent(st, ctx->ev_join[0], 0);
        if (scene->mesh_rounds) S.two_pass_iterations++;
        return;
    }
    ctx->mark(-1);
    if (n_new) { run_k_generate(st, in, n_old, n_new, gen->g0, gen->n_pixels, gen->cam, gen->rc); S.kernel_launches++; ctx->mark(5); }
    const unsigned tg = (n + kTraceBlock - 1) / kTraceBlock;
    // two-pass traversal; not for small iterations (three more launches each); flag 0x100000 opts out (A/B measurements)
    if (scene->defer_meshes && n >= (1u << 16) && !(T.flags & 0x100000u)) {
        const BlasQueues bq{ctx->bq_items, q.count + 8, ctx->pool};  // counters in the free tail of the iteration's slot
        run_k_trace(TraceFlavour{7, true, wk != nullptr, false, true}, tg, st, in, n, ctx->hits, q, scene->d, wk, 0, n_dev, bq, ctx->ties, T.t_min);
        const bool refill = !(T.flags & 0x200000u);  // flag 0x200000: plain grid-stride rounds instead of persistent lanes with refill
        const unsigned grid = refill ? std::min<unsigned>(tg, 148u * 4u * (unsigned)kBlasMinBlocks)  // persistent: one resident warp per slot
                                     : std::min<unsigned>(tg, 148u * 56u);
        for (uint32_t r = 0; r < (uint32_t)kDeferMax; r++)
            run_k_trace_blas(refill, wk != nullptr, grid, st, in, r, bq, ctx->hits, ctx->ties, q, scene->d, wk, T.t_min);
        S.kernel_launches += 1 + kDeferMax;
        S.two_pass_iterations++;
        ctx->mark(4);
        return;
    }
    // fused kernel; for 4-wide scenes flags bits 4-6 pick the register cap (experiment knob): 4: 120, 5: 96, 6: 80, 7: 64 registers
    const uint32_t knob = (T.flags >> 4) & 7u;
    const int min_blocks = !scene->wide || scene->has_volumes || wk ? 6 : knob == 7 ? 8 : knob >= 4 ? (int)knob : 7;
    run_k_trace(TraceFlavour{min_blocks, scene->wide, wk != nullptr, scene->has_volumes, false}, tg, st, in, n, ctx->hits, q, scene->d, wk, T.seed,
                n_dev, BlasQueues{nullptr, nullptr, 0}, nullptr, T.t_min);
    S.kernel_launches++;
    ctx->mark(4);
}

int pt_render_accumulate(pt_ctx* ctx, const pt_scene* scene, const pt_camera* cam, const pt_render_params* p, float* d_accum, pt_stats* stats) {
    if (!ctx || !scene || !p || !d_accum) return fail(PT_ERR_INVALID, "pt_render_accumulate: null argument");
    if (scene->ctx != ctx) return fail(PT_ERR_INVALID, "scene belongs to another context");
    CU(cudaSetDevice(ctx->device));
    DCameraEx dcam;
    int rc = make_camera(scene, cam, &dcam);
    if (rc) return rc;
    const uint32_t n_pixels = dcam.c.width * dcam.c.height;
    const uint64_t total = (uint64_t)n_pixels * p->sample_count;
    // default 32 Mi paths in flight (8.0 GB of state, 9.8 GB with the two-pass traversal's queues): each wavefront iteration costs one host round trip (~0.2 ms with its
    // small-launch tail), so large iterations amortise it (scene 6 FHD: 4 Mi 2681, 8 Mi 2827, 16 Mi 2913 Mrays/s at the
    // time; with the final kernels 16 Mi 3113, 32 Mi 3170, 64 Mi 3204); never more than the render needs
    uint32_t pool = p->pool_paths ? p->pool_paths : (32u << 20);
    if ((uint64_t)pool > total) pool = (uint32_t)std::max<uint64_t>(total, 1);
    pool = (pool + kBlock - 1) / kBlock * kBlock;
    if ((rc = ensure_render_buffers(ctx, pool, scene, &pool))) return rc;
    RenderConst rcst{p->seed, p->sample_begin, p->sample_stride ? p->sample_stride : 1u, p->nan_policy, 0, DEnvDist{nullptr, nullptr, 0, 0}};
    if ((p->flags & PT_RENDER_ENV_IMPORTANCE) && cam->env_is_map) {  // a constant-colour environment needs no importance sampling
        if (scene->env_image != cam->env_image) return fail(PT_ERR_INVALID, "PT_RENDER_ENV_IMPORTANCE: call pt_scene_build_env_sampler for the camera's env_image first");
        rcst.env_importance = 1; rcst.env = scene->env;
    }
    // experiment knob: flag 0x8000 overrides the octant mask of the survivor grouping with flag bits 16-18
    rcst.sort_mask = (p->flags & 0x8000u) ? ((p->flags >> 16) & 7u) : 7u;
    set_index_reciprocals(rcst, n_pixels, total, dcam.c.width);
    cudaStream_t st = ctx->stream;
    CU(cudaMemsetAsync(ctx->d_nonfinite, 0, 4 * sizeof(unsigned long long), st));
    pt_stats S{}; S.width = dcam.c.width; S.height = dcam.c.height;
    float ms_gen = 0, ms_trace = 0, ms_shade = 0;
    CU(cudaEventRecord(ctx->ev0, st));
    uint64_t generated = 0; uint32_t live = 0, live_spawning = 0; int cur = 0;
    const bool nee = (p->flags & PT_RENDER_NEE) != 0;
    if (nee && rcst.env_importance) return fail(PT_ERR_UNSUPPORTED, "PT_RENDER_NEE and PT_RENDER_ENV_IMPORTANCE cannot be combined yet");
    unsigned long long* const wk = ctx->profiling >= 2 ? ctx->d_nonfinite + 1 : nullptr;
    const TraceStage tstage{ctx, scene, p->flags, p->seed, 1e-3, wk, &S};  // Interval::new(eps, INFINITY), camera.rs:171,179
    // one specialised kernel per shade class present in the scene; each walks its queue grid-stride (n: upper bound of the
    // paths in all queues together).  fork: the class kernels run on side streams and join `st` again.
    auto launch_shades = [&](const PathBuf& in, const PathBuf& outb, uint32_t n, const Queues& q, uint32_t* out_count, bool fork) -> int {
#ifndef PT_SHADE_GRID_WAVES
#define PT_SHADE_GRID_WAVES 16   // grid of a shade kernel = at most this many 128-thread blocks per SM; its threads walk the class queue grid-stride
#endif
        const unsigned sg = std::min<unsigned>((n + kBlock - 1) / kBlock, 148u * PT_SHADE_GRID_WAVES);
        if (fork) CU(cudaEventRecord(ctx->ev_fork, st));
        const ShadeArgs sa{in, q, ctx->hits, outb, out_count, d_accum, ctx->d_nonfinite, scene->d, dcam, rcst};
        ctx->mark(-1);
#ifndef PT_SHADE_ORDER_LPT
#define PT_SHADE_ORDER_LPT 0   // measured: scene 3 +1.1 %, scene 6 -0.8 %, scene 7m +-0 (profiles/r2_ab/r2_r_shade_launch_order.log): class order kept
#endif
        // forked launches: the costliest classes first (longest-processing-time order), so that the stage does not end on a long kernel
        // that started last
        static const int order_lpt[N_CLS] = {CLS_PRINCIPLED, CLS_DIFFUSE, CLS_GLASS, CLS_OTHER, CLS_METAL, CLS_MISS, CLS_LIGHT};
        for (int k = 0; k < N_CLS; k++) {
            const int cls = PT_SHADE_ORDER_LPT && fork ? order_lpt[k] : k;
            if (!(scene->class_mask & (1u << cls))) continue;
            cudaStream_t ss = fork ? ctx->shade_stream[cls] : st;
            if (fork) CU(cudaStreamWaitEvent(ss, ctx->ev_fork, 0));
            if (nee) run_k_shade_nee(cls, sg, ss, sa);
            else if (rcst.env_importance) run_k_shade_var(cls, 3, sg, ss, sa);
            else if (scene->general_lights) run_k_shade_var(cls, 2, sg, ss, sa);
            else if (scene->d.n_lights == 0) run_k_shade_nolights(cls, sg, ss, sa);
            else run_k_shade_ref(cls, sg, ss, sa);
            if (fork) { CU(cudaEventRecord(ctx->ev_join[cls], ss)); CU(cudaStreamWaitEvent(st, ctx->ev_join[cls], 0)); }
            else ctx->mark(8 + cls);
            S.kernel_launches++;
        }
        return PT_OK;
    };
    // flag 0x4000: opt out of the batched tail (A/B measurements)
    const bool tail_batching = !nee && !ctx->profiling && !(p->flags & 0x4000u);
    // the megakernel shades with the reference mixture (k_shade<class, 0>) and the fused BVH traversal without media
    const bool tail_mega = !nee && !(p->flags & 0x4000u) && ctx->profiling < 2 && !scene->has_volumes && !scene->general_lights && !rcst.env_importance;
    while (dcam.c.max_depth > 0 && (live > 0 || generated < total)) {
        if (tail_mega && generated == total && live > 0 && live <= tail_max_paths(scene->flat && scene->mesh_rounds == 0)) {
            // ---- tail megakernel: every remaining path runs to its end in one launch
            rcst.fresh_from = 0xFFFFFFFFu;
            CU(cudaMemsetAsync(ctx->d_count, 0, 2 * sizeof(uint32_t), st));
            ctx->mark(-1);
            const TailArgs ta{path_buf(ctx, cur), live, d_accum, ctx->d_nonfinite, scene->d, dcam, rcst, tstage.t_min, ctx->d_count};
            if (scene->flat && !(p->flags & 0x500000u)) run_k_tail_flat(st, ta, scene->top);  // flags 0x100000 / 0x400000 pin the BVH kernels
            else if (scene->wide) run_k_tail_wide(st, ta); else run_k_tail_bin(st, ta);
            ctx->mark(6);
            CU(cudaMemcpyAsync(ctx->h_count, ctx->d_count, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (ctx->profiling) ctx->collect_marks();
            S.segments += ctx->h_count[0]; S.iterations += ctx->h_count[1]; S.kernel_launches++; S.tail_paths = live;
            live = 0;
            continue;
        }
        if (tail_batching && generated == total && live <= kTailBatchMaxLive) {
            // ---- batched tail: kTailBatch iterations per host round trip (see kTailBatch)
            rcst.fresh_from = 0xFFFFFFFFu;  // nothing is started any more: every state record is real
            const uint32_t n0 = live;
            CU(cudaMemsetAsync(ctx->d_count, 0, kTailBatch * kSlot * sizeof(uint32_t), st));
            for (int k = 0; k < kTailBatch; k++) {
                uint32_t* slot = ctx->d_count + kSlot * k;
                const Queues q{ctx->q_items, slot + 4, ctx->pool};
                launch_trace(tstage, path_buf(ctx, cur), n0, 0, nullptr, q, k ? slot - kSlot : nullptr);
                if ((rc = launch_shades(path_buf(ctx, cur), path_buf(ctx, cur ^ 1), n0, q, slot, false))) return rc;
                cur ^= 1;
            }
            CU(cudaMemcpyAsync(ctx->h_count, ctx->d_count, kTailBatch * kSlot * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            uint32_t nk = n0;
            for (int k = 0; k < kTailBatch && nk > 0; k++) { S.segments += nk; S.iterations++; nk = ctx->h_count[kSlot * k]; }
            live = ctx->h_count[kSlot * (kTailBatch - 1)];
            if (live > pool) return fail(PT_ERR_CUDA, "internal error: the shade stage produced more paths than the pool holds");
            continue;
        }
        PathBuf in = path_buf(ctx, cur), outb = path_buf(ctx, cur ^ 1);
        uint32_t n_new = (uint32_t)std::min<uint64_t>(pool - live, total - generated);
        // NEE: a path may spawn a shadow path, so at most pool / 2 spawning (non-shadow) paths enter an iteration
        if (nee) n_new = std::min<uint32_t>(n_new, pool / 2 > live_spawning ? pool / 2 - live_spawning : 0u);
        const uint32_t n = live + n_new;
        CU(cudaMemsetAsync(ctx->d_count, 0, kSlot * sizeof(uint32_t), st));
        const Queues q{ctx->q_items, ctx->d_count + 4, ctx->pool};
        if (ctx->profiling) CU(cudaEventRecord(ctx->evs[1], st));
        // flat scenes: k_top<PRIMARY> starts the n_new paths in slots [live, n) and leaves their state record implicit (wavefront.cuh:
        // RenderConst::fresh_from); the NEE shade kernels and the k_generate route keep real state records
        rcst.fresh_from = (!nee && n_new > 0 && takes_flat_path(scene, p->flags, n)) ? live : 0xFFFFFFFFu;
        const GenArgs gen{generated, n_pixels, dcam, rcst};  // the n_new camera rays of this iteration are generated inside the traversal stage
        launch_trace(tstage, in, live, n_new, &gen, q, nullptr);
        if (ctx->profiling) CU(cudaEventRecord(ctx->evs[2], st));
        // fork: the class kernels run on side streams (unless per-stage events are wanted, or the caller opted out with flag
        // 0x2000).  Measured (repeated runs): scene 6 FHD 168.3 -> 164.1 ms per 128 spp, scene 3 128.0 -> 124.8 ms (+2.5 % each).
        // Small iterations are latency-bound: the cross-stream events would cost more than the overlap returns.
        const bool fork = !ctx->profiling && !(p->flags & 0x2000u) && n >= (1u << 16);
        if ((rc = launch_shades(in, outb, n, q, ctx->d_count, fork))) return rc;
        if (ctx->profiling) CU(cudaEventRecord(ctx->evs[3], st));
        CU(cudaMemcpyAsync(ctx->h_count, ctx->d_count, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (ctx->profiling) {
            ctx->collect_marks();
            float a = 0, b = 0, c2 = 0;
            cudaEventElapsedTime(&b, ctx->evs[1], ctx->evs[2]); cudaEventElapsedTime(&c2, ctx->evs[2], ctx->evs[3]);  // ray generation is part of the traversal stage
            ms_gen += a; ms_trace += b; ms_shade += c2;
        }
        generated += n_new; S.segments += n; S.iterations++;
        live = ctx->h_count[0]; live_spawning = ctx->h_count[1];
        if (live > pool) return fail(PT_ERR_CUDA, "internal error: the shade stage produced more paths than the pool holds");
        cur ^= 1;
    }
    CU(cudaEventRecord(ctx->ev1, st));
    unsigned long long nf[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(nf, ctx->d_nonfinite, sizeof(nf), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    cudaEventElapsedTime(&S.device_ms, ctx->ev0, ctx->ev1);
    S.paths = generated; S.nonfinite = nf[0]; S.raygen_ms = ms_gen; S.trace_ms = ms_trace; S.shade_ms = ms_shade;
    S.node_pairs = nf[1]; S.ref_boxes = nf[2]; S.prim_tests = nf[3];
    if (stats) *stats = S;
    return PT_OK;
}

int pt_render(pt_ctx* ctx, const pt_scene* scene, const pt_camera* cam, const pt_render_params* p, float* h_mean, pt_stats* stats) {
    if (!ctx || !cam || !p || !h_mean) return fail(PT_ERR_INVALID, "pt_render: null argument");
    if (p->sample_count == 0) return fail(PT_ERR_INVALID, "pt_render: sample_count is zero");
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)cam->image_width * pt_camera_image_height(cam) * 3;
    int rcs = ensure_scratch(ctx, n * sizeof(float));  // context-owned, reused by every call
    if (rcs) return rcs;
    float* d_accum = (float*)ctx->scratch;
    cudaError_t e = cudaMemsetAsync(d_accum, 0, n * sizeof(float), ctx->stream);
    int rc = e == cudaSuccess ? pt_render_accumulate(ctx, scene, cam, p, d_accum, stats) : fail(PT_ERR_CUDA, cudaGetErrorString(e));
    if (rc == PT_OK) {
        run_k_scale(ctx->stream, d_accum, 1.0f / (float)p->sample_count, (uint32_t)n, d_accum);
        e = cudaMemcpyAsync(h_mean, d_accum, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(PT_ERR_CUDA, cudaGetErrorString(e));
        if (stats) stats->kernel_launches++;
    }
    return rc;
}

}  // extern "C"

// small RAII device buffer (parity entry points, pt_render_multi's reduce)
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { CU(cudaMalloc(&p, bytes ? bytes : 1)); return PT_OK; }
    int from_host(const void* h, size_t bytes, cudaStream_t st) { int rc = alloc(bytes); if (rc) return rc; if (bytes) CU(cudaMemcpyAsync(p, h, bytes, cudaMemcpyHostToDevice, st)); return PT_OK; }
    int to_host(void* h, size_t bytes, cudaStream_t st) { if (bytes) CU(cudaMemcpyAsync(h, p, bytes, cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st)); CU(cudaGetLastError()); return PT_OK; }
};
}  // namespace

// One device's share of pt_render_multi: own context, own copy of the scene, samples share, share + n_shares, ... of the call.
// The radiance sums stay on the device (d_accum) for the peer reduce.
// Contexts of pt_render_multi are kept for the life of the process (pt_render_multi_release frees them): creating one costs a
// stream, events and above all the path pool (gigabytes of cudaMalloc), which would otherwise be paid on every call.
static std::mutex g_multi_mutex;
static std::vector<pt_ctx*> g_multi_idle;
static int acquire_multi_ctx(int device, pt_ctx** out) {
    {
        std::lock_guard<std::mutex> lock(g_multi_mutex);
        for (size_t k = 0; k < g_multi_idle.size(); k++)
            if (g_multi_idle[k]->device == device) { *out = g_multi_idle[k]; g_multi_idle.erase(g_multi_idle.begin() + k); return PT_OK; }
    }
    return pt_ctx_create(device, out);
}
struct Share {
    int device = 0; pt_ctx* ctx = nullptr; pt_scene* scene = nullptr; float* d_accum = nullptr;
    int rc = PT_OK; std::string err; pt_stats st{};
    void release() {
        if (d_accum) { cudaSetDevice(device); cudaFree(d_accum); d_accum = nullptr; }
        if (scene) { pt_scene_destroy(scene); scene = nullptr; }
        if (ctx) { std::lock_guard<std::mutex> lock(g_multi_mutex); g_multi_idle.push_back(ctx); ctx = nullptr; }
    }
};
static int render_share(Share& S, const pt_scene_desc* desc, const pt_camera* cam, const pt_render_params* p, uint32_t share, uint32_t n_shares, size_t n) {
    int rc = acquire_multi_ctx(S.device, &S.ctx);
    if (rc == PT_OK) rc = pt_scene_create(S.ctx, desc, &S.scene);
    if (rc == PT_OK && (p->flags & PT_RENDER_ENV_IMPORTANCE) && cam->env_is_map) rc = pt_scene_build_env_sampler(S.scene, cam->env_image, 0, 0);
    if (rc == PT_OK) {
        cudaError_t e = cudaMalloc(&S.d_accum, n * sizeof(float));
        if (e == cudaSuccess) e = cudaMemsetAsync(S.d_accum, 0, n * sizeof(float), S.ctx->stream);
        if (e != cudaSuccess) rc = fail(PT_ERR_CUDA, cudaGetErrorString(e));
    }
    if (rc == PT_OK) {
        pt_render_params q = *p;
        const uint32_t stride = p->sample_stride ? p->sample_stride : 1u;
        q.sample_begin = p->sample_begin + share * stride;
        q.sample_stride = stride * n_shares;
        q.sample_count = (p->sample_count - share + n_shares - 1) / n_shares;
        rc = pt_render_accumulate(S.ctx, S.scene, cam, &q, S.d_accum, &S.st);  // returns with the stream synchronised
    }
    return rc;
}
// reduce(sum) of the shares' device buffers on the first share's GPU: peers are read in place over NVLink where the devices
// can address each other, else staged through one cudaMemcpyPeer each; one D2H of the mean image at the end.
static int reduce_shares(std::vector<Share>& shares, size_t n, double scale, float* h_mean, uint32_t* p2p_direct) {
    Share& root = shares[0];
    CU(cudaSetDevice(root.device));
    cudaStream_t st = root.ctx->stream;
    std::vector<const float*> src(shares.size());
    std::vector<DevBuf> staged(shares.size());
    *p2p_direct = 0;
    for (size_t g = 0; g < shares.size(); g++) {
        src[g] = shares[g].d_accum;
        if (shares[g].device == root.device) continue;
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, root.device, shares[g].device));
        if (can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(shares[g].device, 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); (*p2p_direct)++; continue; }
            cudaGetLastError();
        }
        int rc = staged[g].alloc(n * sizeof(float));  // no peer addressing between this pair: one device-to-device copy
        if (rc) return rc;
        CU(cudaMemcpyPeerAsync(staged[g].p, root.device, shares[g].d_accum, shares[g].device, n * sizeof(float), st));
        src[g] = (const float*)staged[g].p;
    }
    DevBuf d_src, d_out;
    int rc;
    if ((rc = d_src.from_host(src.data(), src.size() * sizeof(float*), st)) || (rc = d_out.alloc(n * sizeof(float)))) return rc;
    run_k_reduce_peers(st, (const float* const*)d_src.p, (uint32_t)src.size(), scale, n, (float*)d_out.p);
    return d_out.to_host(h_mean, n * sizeof(float), st);
}

extern "C" {

void pt_render_multi_release(void) {
    std::lock_guard<std::mutex> lock(g_multi_mutex);
    for (pt_ctx* c : g_multi_idle) pt_ctx_destroy(c);
    g_multi_idle.clear();
}
int pt_render_multi(int n_devices, const int* devices, const pt_scene_desc* desc, const pt_camera* cam, const pt_render_params* p,
                    float* h_mean, pt_stats* stats) {
    if (n_devices < 1 || !devices || !desc || !cam || !p || !h_mean) return fail(PT_ERR_INVALID, "pt_render_multi: bad argument");
    if (p->sample_count == 0) return fail(PT_ERR_INVALID, "pt_render_multi: sample_count is zero");
    const size_t n = (size_t)cam->image_width * pt_camera_image_height(cam) * 3;
    const uint32_t G = std::min<uint32_t>((uint32_t)n_devices, p->sample_count);  // a share is at least one sample per pixel
    std::vector<Share> shares(G);
    for (uint32_t g = 0; g < G; g++) shares[g].device = devices[g];
    auto work = [&](uint32_t g) {
        try {
            shares[g].rc = render_share(shares[g], desc, cam, p, g, G, n);
            if (shares[g].rc != PT_OK) shares[g].err = g_err;  // the message is thread-local: carry it to the caller's thread
        } catch (const std::exception& e) { shares[g].rc = PT_ERR_CUDA; shares[g].err = e.what(); }
    };
    std::vector<std::thread> threads;
    for (uint32_t g = 1; g < G; g++) threads.emplace_back(work, g);
    work(0);
    for (auto& t : threads) t.join();
    int rc = PT_OK; std::string msg;
    for (uint32_t g = 0; g < G && rc == PT_OK; g++)
        if (shares[g].rc != PT_OK) { rc = shares[g].rc; msg = "pt_render_multi: device " + std::to_string(devices[g]) + ": " + shares[g].err; }
    uint32_t p2p_direct = 0;
    if (rc == PT_OK && (rc = reduce_shares(shares, n, 1.0 / (double)p->sample_count, h_mean, &p2p_direct)) != PT_OK) msg = "pt_render_multi: reduce: " + g_err;
    if (rc == PT_OK && stats) {
        pt_stats S = shares[0].st;
        for (uint32_t g = 1; g < G; g++) {
            const pt_stats& t = shares[g].st;
            S.paths += t.paths; S.segments += t.segments; S.nonfinite += t.nonfinite; S.kernel_launches += t.kernel_launches;
            S.node_pairs += t.node_pairs; S.ref_boxes += t.ref_boxes; S.prim_tests += t.prim_tests; S.two_pass_iterations += t.two_pass_iterations;
            S.iterations = std::max(S.iterations, t.iterations); S.device_ms = std::max(S.device_ms, t.device_ms);
            S.trace_ms = std::max(S.trace_ms, t.trace_ms); S.shade_ms = std::max(S.shade_ms, t.shade_ms); S.raygen_ms = std::max(S.raygen_ms, t.raygen_ms);
        }
        S.kernel_launches++;          // the reduce kernel
        S.p2p_shares = p2p_direct;
        *stats = S;
    }
    for (auto& s : shares) s.release();
    if (rc != PT_OK) return fail(rc, msg);
    return PT_OK;
}

int pt_tonemap_rgb8(pt_ctx* ctx, const float* d_accum, double scale, uint32_t n_pixels, uint8_t* h_rgb8) {
    if (!ctx || !d_accum || !h_rgb8) return fail(PT_ERR_INVALID, "pt_tonemap_rgb8: null argument");
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = n_pixels * 3;
    if (d_accum == (const float*)ctx->scratch) return fail(PT_ERR_INVALID, "pt_tonemap_rgb8: d_accum must be a caller-owned buffer");
    int rcs = ensure_scratch(ctx, n);
    if (rcs) return rcs;
    uint8_t* d_out = (uint8_t*)ctx->scratch;
    run_k_tonemap(ctx->stream, d_accum, scale, n, d_out);
    cudaError_t e = cudaMemcpyAsync(h_rgb8, d_out, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, cudaGetErrorString(e));
    return PT_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ parity entry points

extern "C" {

int pt_trace_closest(pt_ctx* ctx, const pt_scene* scene, size_t n, const pt_ray* rays, double t_min, pt_hit* hits) {
    if (!ctx || !scene || (n && (!rays || !hits))) return fail(PT_ERR_INVALID, "pt_trace_closest: null argument");
    if (n == 0) return PT_OK;
    CU(cudaSetDevice(ctx->device));
    DevBuf in, out; int rc;
    if ((rc = in.from_host(rays, n * sizeof(pt_ray), ctx->stream)) || (rc = out.alloc(n * sizeof(pt_hit)))) return rc;
    run_k_trace_batch(scene->wide, ctx->stream, (const pt_ray*)in.p, n, t_min, (pt_hit*)out.p, scene->d);
    return out.to_host(hits, n * sizeof(pt_hit), ctx->stream);
}
// parity entry points only: the shade-class queues the traversal stage filled are checked against the hit records
static int check_queues(pt_ctx* ctx, const pt_scene* scene, const Queues& q, uint32_t n, pt_stats* S) {
    DevBuf seen;
    int rc = seen.alloc(((size_t)n + 1) * sizeof(uint32_t));
    if (rc) return rc;
    CU(cudaMemsetAsync(seen.p, 0, ((size_t)n + 1) * sizeof(uint32_t), ctx->stream));
    uint32_t* errors = (uint32_t*)seen.p + n;
    run_k_check_queues(ctx->stream, q, ctx->hits, n, (uint32_t*)seen.p, errors, scene->d);
    uint32_t h_err = 0;
    CU(cudaMemcpyAsync(&h_err, errors, sizeof(h_err), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    S->queue_errors += h_err;
    return PT_OK;
}
int pt_trace_closest_wavefront(pt_ctx* ctx, const pt_scene* scene, size_t n, const pt_ray* rays, double t_min, uint32_t flags, pt_hit* hits, pt_stats* stats) {
    if (!ctx || !scene || (n && (!rays || !hits))) return fail(PT_ERR_INVALID, "pt_trace_closest_wavefront: null argument");
    if (scene->ctx != ctx) return fail(PT_ERR_INVALID, "scene belongs to another context");
    pt_stats S{};
    if (stats) *stats = S;
    if (n == 0) return PT_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint32_t chunk = (uint32_t)std::min<size_t>(n, 4u << 20);  // like a wavefront iteration of that many live paths
    chunk = (chunk + kBlock - 1) / kBlock * kBlock;
    int rc = ensure_render_buffers(ctx, chunk, scene, &chunk);
    if (rc) return rc;
    DevBuf in, out;
    if ((rc = in.alloc((size_t)chunk * sizeof(pt_ray))) || (rc = out.alloc((size_t)chunk * sizeof(pt_hit)))) return rc;
    unsigned long long* const wk = ctx->profiling >= 2 ? ctx->d_nonfinite + 1 : nullptr;
    if (wk) CU(cudaMemsetAsync(ctx->d_nonfinite, 0, 4 * sizeof(unsigned long long), st));
    const TraceStage T{ctx, scene, flags, 0, t_min, wk, &S};  // seed 0: media uniforms keyed like pt_trace_closest's
    const PathBuf pool = path_buf(ctx, 0);
    for (size_t first = 0; first < n; first += chunk) {
        const uint32_t m = (uint32_t)std::min<size_t>(chunk, n - first);
        CU(cudaMemcpyAsync(in.p, rays + first, (size_t)m * sizeof(pt_ray), cudaMemcpyHostToDevice, st));
        run_k_rays_to_pool(st, (const pt_ray*)in.p, m, (uint32_t)first, pool);
        CU(cudaMemsetAsync(ctx->d_count, 0, kSlot * sizeof(uint32_t), st));
        const Queues q{ctx->q_items, ctx->d_count + 4, ctx->pool};
        launch_trace(T, pool, m, 0, nullptr, q, nullptr);
        if ((rc = check_queues(ctx, scene, q, m, &S))) return rc;
        run_k_hits_to_abi(st, (const pt_ray*)in.p, m, ctx->hits, (pt_hit*)out.p, scene->d);
        if ((rc = out.to_host(hits + first, (size_t)m * sizeof(pt_hit), st))) return rc;
        S.segments += m; S.iterations++; S.kernel_launches += 2;
    }
    if (wk) {
        unsigned long long nf[4] = {0, 0, 0, 0};
        CU(cudaMemcpyAsync(nf, ctx->d_nonfinite, sizeof(nf), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        S.node_pairs = nf[1]; S.ref_boxes = nf[2]; S.prim_tests = nf[3];
    }
    if (stats) *stats = S;
    return PT_OK;
}
int pt_trace_camera_wavefront(pt_ctx* ctx, const pt_scene* scene, const pt_camera* cam, uint64_t seed, uint32_t sample, uint32_t flags, pt_ray* rays, pt_hit* hits, pt_stats* stats) {
    if (!ctx || !scene || !cam || !rays || !hits) return fail(PT_ERR_INVALID, "pt_trace_camera_wavefront: null argument");
    if (scene->ctx != ctx) return fail(PT_ERR_INVALID, "scene belongs to another context");
    CU(cudaSetDevice(ctx->device));
    DCameraEx dcam;
    int rc = make_camera(scene, cam, &dcam);
    if (rc) return rc;
    const uint32_t n = dcam.c.width * dcam.c.height;
    uint32_t pool = (n + kBlock - 1) / kBlock * kBlock;
    if ((rc = ensure_render_buffers(ctx, pool, scene, &pool))) return rc;
    if (pool < n) return fail(PT_ERR_UNSUPPORTED, "pt_trace_camera_wavefront: the image does not fit the path pool");
    cudaStream_t st = ctx->stream;
    pt_stats S{}; S.width = dcam.c.width; S.height = dcam.c.height;
    DevBuf d_rays, d_hits;
    if ((rc = d_rays.alloc((size_t)n * sizeof(pt_ray))) || (rc = d_hits.alloc((size_t)n * sizeof(pt_hit)))) return rc;
    RenderConst rcst{seed, sample, 1u, PT_NAN_REFERENCE, 0, DEnvDist{nullptr, nullptr, 0, 0}};
    set_index_reciprocals(rcst, n, n, dcam.c.width);
    const GenArgs gen{0, n, dcam, rcst};
    const TraceStage T{ctx, scene, flags, seed, 1e-3, nullptr, &S};  // Interval::new(eps, INFINITY), camera.rs:171,179
    const PathBuf in = path_buf(ctx, 0);
    CU(cudaMemsetAsync(ctx->d_count, 0, kSlot * sizeof(uint32_t), st));
    const Queues q{ctx->q_items, ctx->d_count + 4, ctx->pool};
    launch_trace(T, in, 0, n, &gen, q, nullptr);
    if ((rc = check_queues(ctx, scene, q, n, &S))) return rc;
    run_k_pool_to_abi(st, in, n, ctx->hits, (pt_ray*)d_rays.p, (pt_hit*)d_hits.p, scene->d);
    if ((rc = d_rays.to_host(rays, (size_t)n * sizeof(pt_ray), st)) || (rc = d_hits.to_host(hits, (size_t)n * sizeof(pt_hit), st))) return rc;
    S.paths = n; S.segments = n; S.iterations = 1; S.kernel_launches++;
    if (stats) *stats = S;
    return PT_OK;
}
int pt_debug_stage_ms(pt_ctx* ctx, double* out16, int reset) {
    if (!ctx) return fail(PT_ERR_INVALID, "pt_debug_stage_ms: null ctx");
    if (out16) memcpy(out16, ctx->stage_ms, sizeof(ctx->stage_ms));
    if (reset) memset(ctx->stage_ms, 0, sizeof(ctx->stage_ms));
    return PT_OK;
}
int pt_debug_div_check(pt_ctx* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches) {
    if (!ctx || !mismatches) return fail(PT_ERR_INVALID, "pt_debug_div_check: null argument");
    if (n > (1ull << 31) * 256ull) return fail(PT_ERR_INVALID, "pt_debug_div_check: n too large");
    CU(cudaSetDevice(ctx->device));
    DevBuf d; int rc;
    if ((rc = d.alloc(sizeof(unsigned long long)))) return rc;
    CU(cudaMemsetAsync(d.p, 0, sizeof(unsigned long long), ctx->stream));
    if (n) run_k_div_check(ctx->stream, n, seed, (unsigned long long*)d.p);
    return d.to_host(mismatches, sizeof(unsigned long long), ctx->stream);
}
int pt_debug_histograms(pt_ctx* ctx, uint64_t* out512, int reset) {
    if (!ctx) return fail(PT_ERR_INVALID, "pt_debug_histograms: null ctx");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(debug_histograms((unsigned long long*)out512, reset != 0));
    return PT_OK;
}
int pt_trace_any(pt_ctx* ctx, const pt_scene* scene, size_t n, const pt_ray* rays, double t_min, const double* t_max, uint8_t* occluded) {
    if (!ctx || !scene || (n && (!rays || !t_max || !occluded))) return fail(PT_ERR_INVALID, "pt_trace_any: null argument");
    if (n == 0) return PT_OK;
    CU(cudaSetDevice(ctx->device));
    DevBuf in, tm, out; int rc;
    if ((rc = in.from_host(rays, n * sizeof(pt_ray), ctx->stream)) || (rc = tm.from_host(t_max, n * 8, ctx->stream)) || (rc = out.alloc(n))) return rc;
    run_k_trace_any_batch(scene->wide, ctx->stream, (const pt_ray*)in.p, n, t_min, (const double*)tm.p, (uint8_t*)out.p, scene->d);
    return out.to_host(occluded, n, ctx->stream);
}
int pt_bsdf_eval_pdf(pt_ctx* ctx, const pt_scene* scene, uint32_t material, size_t n, const pt_bsdf_query* q, pt_bsdf_result* o) {
    if (!ctx || !scene || (n && (!q || !o))) return fail(PT_ERR_INVALID, "pt_bsdf_eval_pdf: null argument");
    if (material >= scene->n_materials) return fail(PT_ERR_INVALID, "pt_bsdf_eval_pdf: bad material");
    if (n == 0) return PT_OK;
    CU(cudaSetDevice(ctx->device));
    DevBuf in, out; int rc;
    if ((rc = in.from_host(q, n * sizeof(pt_bsdf_query), ctx->stream)) || (rc = out.alloc(n * sizeof(pt_bsdf_result)))) return rc;
    run_k_bsdf_eval(ctx->stream, material, n, (const pt_bsdf_query*)in.p, (pt_bsdf_result*)out.p, scene->d);
    return out.to_host(o, n * sizeof(pt_bsdf_result), ctx->stream);
}
int pt_bsdf_sample(pt_ctx* ctx, const pt_scene* scene, uint32_t material, size_t n, const pt_bsdf_query* q, const double* uniforms8, pt_bsdf_sample_result* o) {
    if (!ctx || !scene || (n && (!q || !uniforms8 || !o))) return fail(PT_ERR_INVALID, "pt_bsdf_sample: null argument");
    if (material >= scene->n_materials) return fail(PT_ERR_INVALID, "pt_bsdf_sample: bad material");
    if (n == 0) return PT_OK;
    CU(cudaSetDevice(ctx->device));
    DevBuf in, un, out; int rc;
    if ((rc = in.from_host(q, n * sizeof(pt_bsdf_query), ctx->stream)) || (rc = un.from_host(uniforms8, n * 64, ctx->stream)) || (rc = out.alloc(n * sizeof(pt_bsdf_sample_result)))) return rc;
    run_k_bsdf_sample(ctx->stream, material, n, (const pt_bsdf_query*)in.p, (const double*)un.p, (pt_bsdf_sample_result*)out.p, scene->d);
    return out.to_host(o, n * sizeof(pt_bsdf_sample_result), ctx->stream);
}
int pt_camera_rays(pt_ctx* ctx, const pt_camera* cam, uint64_t seed, size_t n, const uint32_t* row, const uint32_t* col, const uint32_t* sample, pt_ray* o) {
    if (!ctx || !cam || (n && (!row || !col || !sample || !o))) return fail(PT_ERR_INVALID, "pt_camera_rays: null argument");
    if (n == 0) return PT_OK;
    CU(cudaSetDevice(ctx->device));
    pt_camera c = *cam; c.env_is_map = 0;
    DCameraEx dcam; int rc = make_camera(nullptr, &c, &dcam);
    if (rc) return rc;
    DevBuf r, cc, s, out;
    if ((rc = r.from_host(row, n * 4, ctx->stream)) || (rc = cc.from_host(col, n * 4, ctx->stream)) || (rc = s.from_host(sample, n * 4, ctx->stream)) || (rc = out.alloc(n * sizeof(pt_ray)))) return rc;
    run_k_camera_rays(ctx->stream, dcam, seed, n, (const uint32_t*)r.p, (const uint32_t*)cc.p, (const uint32_t*)s.p, (pt_ray*)out.p);
    return out.to_host(o, n * sizeof(pt_ray), ctx->stream);
}
int pt_lights_sample_pdf(pt_ctx* ctx, const pt_scene* scene, size_t n, const pt_vec3* origin, const double* time, const double* uniforms4, pt_vec3* dir, uint32_t* valid, double* pdf) {
    if (!ctx || !scene || (n && (!origin || !time || !uniforms4 || !dir || !valid || !pdf))) return fail(PT_ERR_INVALID, "pt_lights_sample_pdf: null argument");
    if (n == 0) return PT_OK;
    CU(cudaSetDevice(ctx->device));
    DevBuf o, t, u, dd, vv, pp; int rc;
    if ((rc = o.from_host(origin, n * 24, ctx->stream)) || (rc = t.from_host(time, n * 8, ctx->stream)) || (rc = u.from_host(uniforms4, n * 32, ctx->stream)) ||
        (rc = dd.alloc(n * 24)) || (rc = vv.alloc(n * 4)) || (rc = pp.alloc(n * 8))) return rc;
    run_k_lights(ctx->stream, n, (const pt_vec3*)o.p, (const double*)t.p, (const double*)u.p, (pt_vec3*)dd.p, (uint32_t*)vv.p, (double*)pp.p, scene->d);
    if ((rc = dd.to_host(dir, n * 24, ctx->stream)) || (rc = vv.to_host(valid, n * 4, ctx->stream)) || (rc = pp.to_host(pdf, n * 8, ctx->stream))) return rc;
    return PT_OK;
}
int pt_sah_sweep(pt_ctx* ctx, uint32_t n, const double* boxes6, const double* parent6, double* cost3n) {
    if (!ctx || (n && (!boxes6 || !parent6 || !cost3n))) return fail(PT_ERR_INVALID, "pt_sah_sweep: null argument");
    if (n == 0) return PT_OK;
    if (n > (1u << 26)) return fail(PT_ERR_UNSUPPORTED, "pt_sah_sweep: too many items");
    CU(cudaSetDevice(ctx->device));
    DevBuf b, c; int rc;
    if ((rc = b.from_host(boxes6, (size_t)n * sizeof(SahBox), ctx->stream)) || (rc = c.alloc((size_t)n * 3 * sizeof(double)))) return rc;
    SahBox parent; memcpy(&parent, parent6, sizeof(parent));
    run_k_sah_sweep(ctx->stream, n, (const SahBox*)b.p, parent, (double*)c.p);
    return c.to_host(cost3n, (size_t)n * 3 * sizeof(double), ctx->stream);
}
int pt_env_sample_pdf(pt_ctx* ctx, const pt_scene* scene, size_t n, const double* uniforms2, pt_vec3* dir, double* pdf) {
    if (!ctx || !scene || (n && (!uniforms2 || !dir || !pdf))) return fail(PT_ERR_INVALID, "pt_env_sample_pdf: null argument");
    if (scene->env_image == 0xFFFFFFFFu) return fail(PT_ERR_INVALID, "pt_env_sample_pdf: no sampler built (pt_scene_build_env_sampler)");
    if (n == 0) return PT_OK;
    CU(cudaSetDevice(ctx->device));
    DevBuf u, dd, pp; int rc;
    if ((rc = u.from_host(uniforms2, n * 16, ctx->stream)) || (rc = dd.alloc(n * 24)) || (rc = pp.alloc(n * 8))) return rc;
    run_k_env(ctx->stream, n, (const double*)u.p, (pt_vec3*)dd.p, (double*)pp.p, scene->env);
    if ((rc = dd.to_host(dir, n * 24, ctx->stream)) || (rc = pp.to_host(pdf, n * 8, ctx->stream))) return rc;
    return PT_OK;
}

}  // extern "C"
