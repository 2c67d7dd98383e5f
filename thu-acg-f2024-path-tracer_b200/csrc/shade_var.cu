// Translation unit of the shade kernel variants beyond the reference's shipped scenes: k_shade<class, 2> (cuboid / mesh /
// instance lights in World.lights) and k_shade<class, 3> (+ environment importance sampling, PT_RENDER_ENV_IMPORTANCE).
#include "launch.h"
#include "shade_kernels.cuh"

namespace ptd {
void run_k_shade_var(int cls, int var, unsigned grid, cudaStream_t st, const ShadeArgs& a) {
#define PT_GO(C) case C: if (var == 3) k_shade<C, 3><<<grid * (kBlock / kShadeBlock), kShadeBlock, 0, st>>>(a.in, a.q, a.hits, a.out, a.out_count, a.accum, a.nonfinite, a.S, a.cam, a.rc); \
                         else k_shade<C, 2><<<grid * (kBlock / kShadeBlock), kShadeBlock, 0, st>>>(a.in, a.q, a.hits, a.out, a.out_count, a.accum, a.nonfinite, a.S, a.cam, a.rc); break;
    switch (cls) { PT_GO(CLS_MISS) PT_GO(CLS_LIGHT) PT_GO(CLS_DIFFUSE) PT_GO(CLS_METAL) PT_GO(CLS_GLASS) PT_GO(CLS_PRINCIPLED) PT_GO(CLS_OTHER) }
#undef PT_GO
}
}  // namespace ptd
