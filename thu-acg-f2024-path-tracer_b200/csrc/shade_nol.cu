// Translation unit of the shade kernels for scenes whose World.lights is empty (k_shade<class, 4>: no light sampler, no light pdf).
#include "launch.h"
#include "shade_kernels.cuh"

namespace ptd {
void run_k_shade_nolights(int cls, unsigned grid, cudaStream_t st, const ShadeArgs& a) {
#define PT_GO(C) case C: k_shade<C, 4><<<grid * (kBlock / kShadeBlock), kShadeBlock, 0, st>>>(a.in, a.q, a.hits, a.out, a.out_count, a.accum, a.nonfinite, a.S, a.cam, a.rc); break;
    switch (cls) { PT_GO(CLS_MISS) PT_GO(CLS_LIGHT) PT_GO(CLS_DIFFUSE) PT_GO(CLS_METAL) PT_GO(CLS_GLASS) PT_GO(CLS_PRINCIPLED) PT_GO(CLS_OTHER) }
#undef PT_GO
}
}  // namespace ptd
