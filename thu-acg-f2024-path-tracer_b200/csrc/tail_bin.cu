// Translation unit of the tail megakernel for scenes traversed with binary node pairs (tail_kernels.cuh).
#include "launch.h"
#include "tail_kernels.cuh"

namespace ptd {
void run_k_tail_bin(cudaStream_t st, const TailArgs& a) {
    k_tail<0><<<(a.n + kTailBlock - 1) / kTailBlock, kTailBlock, 0, st>>>(a.in, a.n, a.accum, a.nonfinite, a.S, a.cam, a.rc, a.t_min, a.counters, TopList{});
}
// first use of a kernel loads its code (CUDA loads lazily, and k_tail is the largest kernel of the library: ~40 ms); pt_scene_create
// calls this so that the cost is not paid inside the first render
void preload_k_tail_bin() { cudaFuncAttributes a; (void)cudaFuncGetAttributes(&a, k_tail<0>); }
}  // namespace ptd
