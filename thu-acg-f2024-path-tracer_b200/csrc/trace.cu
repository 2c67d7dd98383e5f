// Translation unit of the traversal-stage kernels (trace_kernels.cuh) and their launchers (launch.h).
#include <algorithm>

#include "launch.h"
#include "trace_kernels.cuh"

namespace ptd {

void run_k_trace(const TraceFlavour& f, unsigned grid, cudaStream_t st, PathBuf in, uint32_t n, HitRec* hits, Queues q, const DScene& S,
                 unsigned long long* work, uint64_t seed, const uint32_t* n_dev, BlasQueues bq, uint2* ties, double t_min) {
#define PT_GO(...) k_trace<__VA_ARGS__><<<grid, kTraceBlock, 0, st>>>(in, n, hits, q, S, work, seed, n_dev, bq, ties, t_min)
    if (f.defer) { if (f.count) PT_GO(7, true, true, false, true); else PT_GO(7, true, false, false, true); return; }
    if (f.vol) {
        if (f.wide) { if (f.count) PT_GO(6, true, true, true); else PT_GO(6, true, false, true); }
        else { if (f.count) PT_GO(6, false, true, true); else PT_GO(6, false, false, true); }
        return;
    }
    if (f.count) { if (f.wide) PT_GO(6, true, true); else PT_GO(6, false, true); return; }
    if (!f.wide) { PT_GO(6, false); return; }  // 80 regs (72: -2 % .. +1.5 %)
    switch (f.min_blocks) {                    // register cap of the 4-wide flavour (experiment knob, pt_render_params.flags bits 4-6)
        case 4: PT_GO(4, true); break;         // 120 regs
        case 5: PT_GO(5, true); break;         // 96 regs
        case 6: PT_GO(6, true); break;         // 80 regs
        case 8: PT_GO(8, true); break;         // 64 regs, spills
        default: PT_GO(7, true);               // 72 regs, 28 warps/SM (measured best: +2 % over 80)
    }
#undef PT_GO
}
void run_k_trace_blas(bool refill, bool count, unsigned grid, cudaStream_t st, PathBuf in, uint32_t round, BlasQueues bq, HitRec* hits, uint2* ties,
                      Queues q, const DScene& S, unsigned long long* work, double t_min) {
    if (refill) {
        if (count) k_trace_blas_refill<true><<<grid, kTraceBlock, 0, st>>>(in, round, bq, hits, ties, q, S, work, t_min);
        else k_trace_blas_refill<false><<<grid, kTraceBlock, 0, st>>>(in, round, bq, hits, ties, q, S, work, t_min);
    } else if (count) k_trace_blas<true><<<grid, kTraceBlock, 0, st>>>(in, round, bq, hits, ties, q, S, work, t_min);
    else k_trace_blas<false><<<grid, kTraceBlock, 0, st>>>(in, round, bq, hits, ties, q, S, work, t_min);
}
void run_k_top(bool primary, bool count, cudaStream_t st, PathBuf pool, uint32_t slot0, uint32_t n, HitRec* hits, Queues q, const DScene& S, const TopList& top,
               MeshQueues mq, uint2* ties, double t_min, const GenArgs& gen, const uint32_t* n_dev, unsigned long long* work) {
    const unsigned grid = (n + kBlock - 1) / kBlock;
#define PT_GO(P, C) k_top<P, C><<<grid, kBlock, 0, st>>>(pool, slot0, n, hits, q, S, top, mq, ties, t_min, gen, n_dev, work)
    if (primary) { if (count) PT_GO(true, true); else PT_GO(true, false); }
    else { if (count) PT_GO(false, true); else PT_GO(false, false); }
#undef PT_GO
}
void run_k_mesh_enter(bool count, unsigned grid, cudaStream_t st, PathBuf pool, uint32_t round, MeshQueues mq, const HitRec* hits, const uint2* ties, Queues q,
                      const DScene& S, const TopList& top, double t_min, unsigned long long* work) {
    if (count) k_mesh_enter<true><<<grid, kBlock, 0, st>>>(pool, round, mq, hits, ties, q, S, top, t_min, work);
    else k_mesh_enter<false><<<grid, kBlock, 0, st>>>(pool, round, mq, hits, ties, q, S, top, t_min, work);
}
void run_k_mesh_walk(bool count, unsigned grid, cudaStream_t st, uint32_t round, MeshQueues mq, HitRec* hits, uint2* ties, Queues q, const DScene& S, double t_min,
                     unsigned long long* work) {
    if (count) k_mesh_walk<true><<<grid, kTraceBlock, 0, st>>>(round, mq, hits, ties, q, S, t_min, work);
    else k_mesh_walk<false><<<grid, kTraceBlock, 0, st>>>(round, mq, hits, ties, q, S, t_min, work);
}
void run_k_mesh_multi(bool count, unsigned grid, cudaStream_t st, PathBuf pool, MeshQueues mq, HitRec* hits, const uint2* ties, Queues q, const DScene& S, double t_min,
                      unsigned long long* work) {
    if (count) k_mesh_multi<true><<<grid, kTraceBlock, 0, st>>>(pool, mq, hits, ties, q, S, t_min, work);
    else k_mesh_multi<false><<<grid, kTraceBlock, 0, st>>>(pool, mq, hits, ties, q, S, t_min, work);
}
unsigned mesh_walk_resident_warps() { return 148u * 4u * (unsigned)kWalkMinBlocks; }
static unsigned grid128(size_t n) { return (unsigned)((n + 127) / 128); }
void run_k_trace_batch(bool wide, cudaStream_t st, const pt_ray* rays, size_t n, double t_min, pt_hit* out, const DScene& S) {
    if (wide) k_trace_batch<true><<<grid128(n), 128, 0, st>>>(rays, n, t_min, out, S);
    else k_trace_batch<false><<<grid128(n), 128, 0, st>>>(rays, n, t_min, out, S);
}
void run_k_trace_any_batch(bool wide, cudaStream_t st, const pt_ray* rays, size_t n, double t_min, const double* t_max, uint8_t* out, const DScene& S) {
    if (wide) k_trace_any_batch<true><<<grid128(n), 128, 0, st>>>(rays, n, t_min, t_max, out, S);
    else k_trace_any_batch<false><<<grid128(n), 128, 0, st>>>(rays, n, t_min, t_max, out, S);
}
void run_k_rays_to_pool(cudaStream_t st, const pt_ray* rays, uint32_t n, uint32_t first, PathBuf out) {
    k_rays_to_pool<<<grid128(n), 128, 0, st>>>(rays, n, first, out);
}
void run_k_hits_to_abi(cudaStream_t st, const pt_ray* rays, uint32_t n, const HitRec* hits, pt_hit* out, const DScene& S) {
    k_hits_to_abi<<<grid128(n), 128, 0, st>>>(rays, n, hits, out, S);
}
void run_k_pool_to_abi(cudaStream_t st, PathBuf pool, uint32_t n, const HitRec* hits, pt_ray* out_rays, pt_hit* out_hits, const DScene& S) {
    k_pool_to_abi<<<grid128(n), 128, 0, st>>>(pool, n, hits, out_rays, out_hits, S);
}
void run_k_check_queues(cudaStream_t st, Queues q, const HitRec* hits, uint32_t n, uint32_t* seen, uint32_t* errors, const DScene& S) {
    k_check_queues<<<dim3(std::max(1u, std::min(grid128(n), 2048u)), N_CLS), 128, 0, st>>>(q, hits, n, seen, errors, S);
    k_check_seen<<<grid128(n), 128, 0, st>>>(seen, n, errors);
}
cudaError_t debug_histograms(unsigned long long* out512, bool reset) {
    cudaError_t e = cudaSuccess;
    if (out512) e = cudaMemcpyFromSymbol(out512, g_hist, sizeof(g_hist));
    if (e == cudaSuccess && reset) { static const unsigned long long zero[8 * 64] = {0}; e = cudaMemcpyToSymbol(g_hist, zero, sizeof(zero)); }
    return e;
}

}  // namespace ptd
