// gf_eval_lines_f64_kernel launcher (DOUBLE 256-byte records). See gf_launch.h.
#include <cstring>

#include "gf_eval_lines_f64.cuh"
#include "gf_launch.h"

namespace gfb {

static_assert(kLinesF64Block == kLinesF64BlockThreads, "gf_launch.h out of date");

template <int NG, int FMODE, bool SINGLE>
static void launch4(const EvalParams& p, cudaStream_t stream) {
    const unsigned blocks = (unsigned) ((p.total + kLinesF64Block - 1) / kLinesF64Block);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(kLinesF64Block);
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = p.pdl ? 1 : 0;
    if (p.grid_energies) cudaLaunchKernelEx(&cfg, gf_eval_lines_f64_kernel<NG, FMODE, SINGLE, true>, p);
    else cudaLaunchKernelEx(&cfg, gf_eval_lines_f64_kernel<NG, FMODE, SINGLE, false>, p);
}

template <int NG, int FMODE>
static void launch3(const EvalParams& p, cudaStream_t stream) {
    if (p.n_replicas == 1 && p.slots == nullptr) launch4<NG, FMODE, true>(p, stream);
    else launch4<NG, FMODE, false>(p, stream);
}

template <int NG>
static void launch2(const EvalParams& p, int fmode, cudaStream_t stream) {
    if (!p.forces) launch3<NG, kForceNone>(p, stream);
    else if (fmode == GFB_FORCE_F64_STORE) launch3<NG, GFB_FORCE_F64_STORE>(p, stream);
    else if (fmode == GFB_FORCE_F32_STORE) launch3<NG, GFB_FORCE_F32_STORE>(p, stream);
    else if (fmode == GFB_FORCE_FIXED_ADD) launch3<NG, GFB_FORCE_FIXED_ADD>(p, stream);
    else launch3<NG, GFB_FORCE_F64_ADD>(p, stream);
}

void launch_lines_f64(const EvalParams& p, int fmode, cudaStream_t stream) {
    switch (p.n_grids) {
        case 2: launch2<2>(p, fmode, stream); break;
        case 3: launch2<3>(p, fmode, stream); break;
        default: launch2<4>(p, fmode, stream); break;
    }
}

}  // namespace gfb
