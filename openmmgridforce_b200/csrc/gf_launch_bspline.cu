// gf_eval_bspline_kernel launcher. See gf_launch.h.
#include "gf_eval_bspline.cuh"
#include "gf_launch.h"

namespace gfb {

static_assert(kBsBlock == kBsplineBlockThreads, "gf_launch.h out of date");

void launch_bspline(const EvalParams& p, cudaStream_t stream) {
    const unsigned blocks = (unsigned) ((p.total + kBsBlock - 1) / kBsBlock);
    if (p.n_replicas == 1 && p.slots == nullptr) gf_eval_bspline_kernel<true><<<blocks, kBsBlock, 0, stream>>>(p);
    else gf_eval_bspline_kernel<false><<<blocks, kBsBlock, 0, stream>>>(p);
}

}  // namespace gfb
