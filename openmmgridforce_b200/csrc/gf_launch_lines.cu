// gf_eval_lines_kernel launcher (MIXED packed cells / 128-byte records). See gf_launch.h.
// GFB_LINES_NG selects which grid counts this translation unit instantiates (the Makefile builds one object per count so
// that they compile in parallel); undefined = all four.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "gf_eval_lines.cuh"
#include "gf_launch.h"

namespace gfb {

// Tunables of the tile-striding launch (A/B measurements; defaults measured on B200).
static double env_double(const char* name, double dflt) {
    const char* e = getenv(name);
    return e ? atof(e) : dflt;
}

template <int NG, int FMODE, int FPATH, bool SINGLE>
static void launch_lines4(const EvalParams& p_in, cudaStream_t stream) {
    constexpr int block = lines_block(NG);
    EvalParams p = p_in;
    unsigned blocks = (unsigned) ((p.total + block - 1) / block);
    constexpr bool kPersistable = !SINGLE && (FMODE == GFB_FORCE_F64_ADD || FMODE == GFB_FORCE_FIXED_ADD || FMODE == kForceNone);
    bool persist = false;
    if (kPersistable && p.defer && !p.grid_energies) {
        // Small launch under launch overlap: a grid that is resident all at once strides over the tiles (gf_eval_lines.cuh).
        // resident = SMs x blocks of THIS instantiation per SM (asked from the runtime: registers, parked-energy smem).
        // Depth D: the grid is resident / D blocks, so that D consecutive launches share the SMs and every block still
        // runs >= min_tiles tiles — a launch whose blocks live for one or two tiles is over before the launch behind it
        // has been scheduled, and the freed slots sit empty for that latency.
        static const int per_sm = [] {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gf_eval_lines_kernel<NG, FMODE, FPATH, SINGLE, false, kPersistable>, block, 0) != cudaSuccess || n < 1)
                n = 1;
            return n;
        }();
        // A block must be able to park ALL its tiles (GFB_PERSIST_DEFER): one that reaches its wait before the end stalls
        // there until the launch in front — which started at almost the same time — is over (measured: 11.1 -> 18.1 us).
        // Tiles per block up to which the variant pays (C5 grids, 47-atom replicas, us per launch, one block per tile ->
        // tile-striding): 4096 replicas 8.6 -> 6.2, 8192 (2.5 tiles) 13.4 -> 10.8, 12288 (3.8) 18.3 -> 15.9, 16384 (5.1)
        // 23.2 -> 21.7, 32768 (10.2) 41.4 -> 47.0, 65536 83.2 -> 91.5: it runs 32 warps per SM instead of 40.
        static const double max_waves = env_double("GFB_PERSIST_MAX_WAVES", 6.0);
        static const int max_depth = (int) env_double("GFB_PERSIST_MAX_DEPTH", 2.0);
        const double resident = (double) p.persist_blocks * per_sm;   // p.persist_blocks carries the SM count
        const double tiles = (double) blocks;
        if (tiles > resident / max_depth && tiles <= max_waves * resident) {
            int depth = (int) floor((double) GFB_PERSIST_DEFER * resident / tiles);
            depth = depth < 1 ? 1 : (depth > max_depth ? max_depth : depth);
            const unsigned grid = (unsigned) ceil(resident / depth);
            if (grid < blocks) {
                blocks = grid;
                persist = true;
            }
        }
    }
    p.defer = persist ? 1u : 0u;
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
    if (p.grid_energies) {
        cudaLaunchKernelEx(&cfg, gf_eval_lines_kernel<NG, FMODE, FPATH, SINGLE, true>, p);
    } else if (kPersistable && persist) {
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
