// gf_eval_bspline_kernel / gf_eval_bspline_f64_kernel launchers. See gf_launch.h.
#include "gf_eval_bspline.cuh"
#include "gf_eval_bspline_f64.cuh"
#include "gf_launch.h"

namespace gfb {

static_assert(kBsBlock == kBsplineBlockThreads && kBsF64Block == kBsplineF64BlockThreads, "gf_launch.h out of date");

void launch_bspline(const EvalParams& p, cudaStream_t stream) {
    const unsigned blocks = (unsigned) ((p.total + kBsBlock - 1) / kBsBlock);
    if (p.n_replicas == 1 && p.slots == nullptr) gf_eval_bspline_kernel<true, 1><<<blocks, kBsBlock, 0, stream>>>(p);
    else gf_eval_bspline_kernel<false, 1><<<blocks, kBsBlock, 0, stream>>>(p);
}

void launch_tricubic_records(const EvalParams& p, cudaStream_t stream) {
    const unsigned blocks = (unsigned) ((p.total + kBsBlock - 1) / kBsBlock);
    if (p.n_replicas == 1 && p.slots == nullptr) gf_eval_bspline_kernel<true, 2><<<blocks, kBsBlock, 0, stream>>>(p);
    else gf_eval_bspline_kernel<false, 2><<<blocks, kBsBlock, 0, stream>>>(p);
}

void launch_bspline_f64(const EvalParams& p, cudaStream_t stream) {
    const unsigned blocks = (unsigned) ((p.total + kBsF64Block - 1) / kBsF64Block);
    if (p.n_replicas == 1 && p.slots == nullptr) gf_eval_bspline_f64_kernel<true, 1><<<blocks, kBsF64Block, 0, stream>>>(p);
    else gf_eval_bspline_f64_kernel<false, 1><<<blocks, kBsF64Block, 0, stream>>>(p);
}

void launch_tricubic_records_f64(const EvalParams& p, cudaStream_t stream) {
    const unsigned blocks = (unsigned) ((p.total + kBsF64Block - 1) / kBsF64Block);
    if (p.n_replicas == 1 && p.slots == nullptr) gf_eval_bspline_f64_kernel<true, 2><<<blocks, kBsF64Block, 0, stream>>>(p);
    else gf_eval_bspline_f64_kernel<false, 2><<<blocks, kBsF64Block, 0, stream>>>(p);
}

}  // namespace gfb
