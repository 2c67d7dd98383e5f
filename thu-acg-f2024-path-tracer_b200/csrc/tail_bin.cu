// Translation unit of the tail megakernel for scenes traversed with binary node pairs (tail_kernels.cuh).
#include "launch.h"
#include "tail_kernels.cuh"

namespace ptd {
void run_k_tail_bin(cudaStream_t st, const TailArgs& a) {
    k_tail<false><<<(a.n + kTailBlock - 1) / kTailBlock, kTailBlock, 0, st>>>(a.in, a.n, a.accum, a.nonfinite, a.S, a.cam, a.rc, a.t_min, a.counters);
}
}  // namespace ptd
