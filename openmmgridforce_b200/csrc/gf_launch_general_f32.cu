// gf_eval_kernel instantiations for MIXED precision (values stored FP32). See gf_launch.h.
#include "gf_kernels.cuh"
#include "gf_launch.h"

namespace gfb {

template <typename S, int LAYOUT, int NG, bool SAME>
static void launch_eval3(const EvalParams& p, cudaStream_t stream) {
    const unsigned blocks = (unsigned) ((p.total + kBlock - 1) / kBlock);
    if (p.n_replicas == 1 && p.slots == nullptr) gf_eval_kernel<S, LAYOUT, NG, SAME, true><<<blocks, kBlock, 0, stream>>>(p);
    else gf_eval_kernel<S, LAYOUT, NG, SAME, false><<<blocks, kBlock, 0, stream>>>(p);
}

template <typename S, int LAYOUT>
static void launch_eval2(const EvalParams& p, bool same, cudaStream_t stream) {
    if (p.n_grids == 1) launch_eval3<S, LAYOUT, 1, true>(p, stream);
    else if (p.n_grids == 3 && same) launch_eval3<S, LAYOUT, 3, true>(p, stream);
    else if (same) launch_eval3<S, LAYOUT, 0, true>(p, stream);
    else launch_eval3<S, LAYOUT, 0, false>(p, stream);
}

#ifndef GFB_GENERAL_F64
void launch_general_f32(const EvalParams& p, int layout, bool same, cudaStream_t stream) {
    if (layout == GFB_LAYOUT_BSPLINE) launch_eval3<float, GFB_LAYOUT_BSPLINE, 0, false>(p, stream);   // each grid classified on its own
    else if (layout == GFB_LAYOUT_POINTS) launch_eval3<float, GFB_LAYOUT_POINTS, 0, false>(p, stream);
    else if (layout == GFB_LAYOUT_HERMITE) launch_eval3<float, GFB_LAYOUT_HERMITE, 0, false>(p, stream);
    else if (layout == GFB_LAYOUT_BSPLINE_POINTS) launch_eval3<float, GFB_LAYOUT_BSPLINE_POINTS, 0, false>(p, stream);
    else if (layout == GFB_LAYOUT_CELLS) launch_eval2<float, GFB_LAYOUT_CELLS>(p, same, stream);
    else if (layout == GFB_LAYOUT_ROWS) launch_eval2<float, GFB_LAYOUT_ROWS>(p, same, stream);
    else launch_eval2<float, GFB_LAYOUT_PAIRS>(p, same, stream);
}
#else
void launch_general_f64(const EvalParams& p, int layout, bool same, cudaStream_t stream) {
    if (layout == GFB_LAYOUT_BSPLINE) launch_eval3<double, GFB_LAYOUT_BSPLINE, 0, false>(p, stream);
    else if (layout == GFB_LAYOUT_POINTS) launch_eval3<double, GFB_LAYOUT_POINTS, 0, false>(p, stream);
    else if (layout == GFB_LAYOUT_HERMITE) launch_eval3<double, GFB_LAYOUT_HERMITE, 0, false>(p, stream);
    else if (layout == GFB_LAYOUT_BSPLINE_POINTS) launch_eval3<double, GFB_LAYOUT_BSPLINE_POINTS, 0, false>(p, stream);
    else if (layout == GFB_LAYOUT_CELLS) launch_eval2<double, GFB_LAYOUT_CELLS>(p, same, stream);
    else launch_eval2<double, GFB_LAYOUT_ROWS>(p, same, stream);
}
#endif

}  // namespace gfb
