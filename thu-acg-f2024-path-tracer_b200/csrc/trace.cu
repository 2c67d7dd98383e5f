// Translation unit of the traversal-stage kernels (trace_kernels.cuh) and their launchers (launch.h).
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
cudaError_t debug_histograms(unsigned long long* out512, bool reset) {
    cudaError_t e = cudaSuccess;
    if (out512) e = cudaMemcpyFromSymbol(out512, g_hist, sizeof(g_hist));
    if (e == cudaSuccess && reset) { static const unsigned long long zero[8 * 64] = {0}; e = cudaMemcpyToSymbol(g_hist, zero, sizeof(zero)); }
    return e;
}

}  // namespace ptd
