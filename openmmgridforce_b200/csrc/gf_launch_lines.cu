// gf_eval_lines_kernel launcher (MIXED packed cells / 128-byte records). See gf_launch.h.
// GFB_LINES_NG selects which grid counts this translation unit instantiates (the Makefile builds one object per count so
// that they compile in parallel); undefined = all four.
#include <cstring>

#include "gf_eval_lines.cuh"
#include "gf_launch.h"

namespace gfb {

template <int NG, int FMODE, int FPATH, bool SINGLE>
static void launch_lines4(const EvalParams& p, cudaStream_t stream) {
    constexpr int block = lines_block(NG);
    unsigned blocks = (unsigned) ((p.total + block - 1) / block);
    if (!SINGLE && p.defer && p.persist_blocks && p.persist_blocks < blocks) blocks = p.persist_blocks;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(block);
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = p.pdl ? 1 : 0;
    constexpr bool kPersistable = !SINGLE && (FMODE == GFB_FORCE_F64_ADD || FMODE == GFB_FORCE_FIXED_ADD || FMODE == kForceNone);
    if (p.grid_energies) {
        cudaLaunchKernelEx(&cfg, gf_eval_lines_kernel<NG, FMODE, FPATH, SINGLE, true>, p);
    } else if (kPersistable && p.defer) {
        cudaLaunchKernelEx(&cfg, gf_eval_lines_kernel<NG, FMODE, FPATH, SINGLE, false, kPersistable>, p);
    } else {
        cudaLaunchKernelEx(&cfg, gf_eval_lines_kernel<NG, FMODE, FPATH, SINGLE, false>, p);
    }
}

template <int NG, int FMODE, int FPATH>
static void launch_lines3(const EvalParams& p, cudaStream_t stream) {
    if (p.n_replicas == 1 && p.slots == nullptr) launch_lines4<NG, FMODE, FPATH, true>(p, stream);
    else launch_lines4<NG, FMODE, FPATH, false>(p, stream);
}

template <int NG>
void launch_lines_ng(const EvalParams& p, int fmode, int fpath, cudaStream_t stream) {
    if (!p.forces) {
        launch_lines3<NG, kForceNone, kForceRed>(p, stream);
    } else if (fmode == GFB_FORCE_F64_STORE) {
        launch_lines3<NG, GFB_FORCE_F64_STORE, kForceRed>(p, stream);
    } else if (fmode == GFB_FORCE_F32_STORE) {
        launch_lines3<NG, GFB_FORCE_F32_STORE, kForceRed>(p, stream);
    } else if (fmode == GFB_FORCE_FIXED_ADD) {
        if (fpath == kForcePrefetch) launch_lines3<NG, GFB_FORCE_FIXED_ADD, kForcePrefetch>(p, stream);
        else launch_lines3<NG, GFB_FORCE_FIXED_ADD, kForceRed>(p, stream);
    } else {
        if (fpath == kForcePrefetch) launch_lines3<NG, GFB_FORCE_F64_ADD, kForcePrefetch>(p, stream);
        else launch_lines3<NG, GFB_FORCE_F64_ADD, kForceRed>(p, stream);
    }
}

#ifdef GFB_LINES_NG
template void launch_lines_ng<GFB_LINES_NG>(const EvalParams&, int, int, cudaStream_t);
#else
template void launch_lines_ng<1>(const EvalParams&, int, int, cudaStream_t);
template void launch_lines_ng<2>(const EvalParams&, int, int, cudaStream_t);
template void launch_lines_ng<3>(const EvalParams&, int, int, cudaStream_t);
template void launch_lines_ng<4>(const EvalParams&, int, int, cudaStream_t);
#endif

static_assert(lines_block(1) == lines_block_threads(1) && lines_block(3) == lines_block_threads(3), "gf_launch.h out of date");

}  // namespace gfb
