// Host-callable launchers of the evaluation kernels; each family is compiled in its own translation unit so that the
// library builds in parallel (gf_launch_general_f32.cu / _f64.cu, gf_launch_lines.cu, gf_launch_records_f64.cu,
// gf_launch_bspline.cu). They launch and return; the caller checks cudaGetLastError().
#ifndef GF_LAUNCH_H_
#define GF_LAUNCH_H_

#include <cuda_runtime.h>

#include "gf_params.h"

namespace gfb {

// gf_eval_kernel (gf_kernels.cuh): every layout, both precisions, inv-power, mixed geometries, > 4 grids.
void launch_general_f32(const EvalParams& p, int layout, bool same_geom, cudaStream_t stream);
void launch_general_f64(const EvalParams& p, int layout, bool same_geom, cudaStream_t stream);
constexpr int kGeneralBlock = 256;

// gf_eval_lines_kernel (gf_eval_lines.cuh): MIXED packed cells of one geometry, 1-4 grids (p.n_grids), no inv-power.
// fmode: gfb_force_mode; forces == NULL in p selects the energy-only instantiation. fpath: 0 RED, 1 RED + L2 prefetch.
// One explicit instantiation per grid count, each in its own object file (gf_launch_lines.cu with -DGFB_LINES_NG=k).
template <int NG>
void launch_lines_ng(const EvalParams& p, int fmode, int fpath, cudaStream_t stream);
#ifndef GFB_LINES_NG   // the per-grid-count objects (gf_launch_lines.cu with -DGFB_LINES_NG=k) must not see this dispatcher: its
                       // non-dependent calls would instantiate every grid count in every one of them
inline void launch_lines(const EvalParams& p, int fmode, int fpath, cudaStream_t stream) {
    switch (p.n_grids) {
        case 1: launch_lines_ng<1>(p, fmode, fpath, stream); break;
        case 2: launch_lines_ng<2>(p, fmode, fpath, stream); break;
        case 3: launch_lines_ng<3>(p, fmode, fpath, stream); break;
        default: launch_lines_ng<4>(p, fmode, fpath, stream); break;
    }
}
#endif
#ifndef GFB_LINES_BLOCK_MULTI
#define GFB_LINES_BLOCK_MULTI 64
#endif
constexpr int lines_block_threads(int n_grids) { return n_grids == 1 ? 256 : GFB_LINES_BLOCK_MULTI; }

// gf_eval_lines_f64_kernel (gf_eval_lines_f64.cuh): DOUBLE 256-byte records, 2-4 grids of one geometry, no inv-power.
void launch_lines_f64(const EvalParams& p, int fmode, cudaStream_t stream);
constexpr int kLinesF64BlockThreads = 128;

// gf_eval_bspline_kernel (gf_eval_bspline.cuh): MIXED B-spline records of one geometry.
void launch_bspline(const EvalParams& p, cudaStream_t stream);
// the same kernel with METHOD = 2: MIXED tricubic Hermite on HERMITE records of one geometry.
void launch_tricubic_records(const EvalParams& p, cudaStream_t stream);
constexpr int kBsplineBlockThreads = 128;

// gf_eval_bspline_f64_kernel (gf_eval_bspline_f64.cuh): DOUBLE B-spline records of one geometry.
void launch_bspline_f64(const EvalParams& p, cudaStream_t stream);
void launch_tricubic_records_f64(const EvalParams& p, cudaStream_t stream);   // METHOD = 2 on DOUBLE HERMITE records
constexpr int kBsplineF64BlockThreads = 64;

}  // namespace gfb
#endif
