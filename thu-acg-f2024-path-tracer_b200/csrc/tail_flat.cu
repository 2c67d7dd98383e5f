// Translation unit of the tail megakernel for flat scenes (tail_kernels.cuh: trace_flat — top-level list + single-ray mesh walks).
#include "launch.h"
#include "tail_kernels.cuh"

namespace ptd {
void run_k_tail_flat(cudaStream_t st, const TailArgs& a, const TopList& top) {
    k_tail<2><<<(a.n + kTailBlock - 1) / kTailBlock, kTailBlock, 0, st>>>(a.in, a.n, a.accum, a.nonfinite, a.S, a.cam, a.rc, a.t_min, a.counters, top);
}
// first use of a kernel loads its code (see tail_wide.cu); pt_scene_create calls this for flat scenes
void preload_k_tail_flat() { cudaFuncAttributes a; (void)cudaFuncGetAttributes(&a, k_tail<2>); }
}  // namespace ptd
