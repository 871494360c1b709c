// Cubic B-spline evaluation kernel for MIXED-precision grids in the BSPLINE tile layout that share one geometry
// (GridForce::setInterpolationMethod(1); reference platforms/reference/src/ReferenceGridForceKernels.cpp:727-795).
// Everything else B-spline (DOUBLE precision, grids of different geometry, an evaluation order) runs
// gf_eval_kernel<S, BSPLINE, ...> in gf_kernels.cuh, which holds the layout's description and the reference arithmetic.
//
// Why a second kernel. The general kernel reads a stencil with 16 LDG.E.256 per lane and grid. An SM retires about one
// gather LANE per clock whatever the load width, so 16 x 3 grids x 3.08 M atoms of C5's shape cost
// 0.53 ms in the L1 pipeline alone (measured: 1.07 ms at 63 % L1 throughput, 128 registers, 24 % of warp slots).
// Here a stencil's four 128-byte tiles are fetched as four full LINES: the eight lanes of an octet copy the eight
// 16-byte granules of one tile with cp.async (LDGSTS.128), one warp instruction = the 4 tiles of one atom = 4 L2
// requests instead of 16 x 32 B from one lane. The tiles land in an XOR-swizzled slice of shared memory (16 KB per warp,
// no bank conflicts on either side); the owning lane then reads its 64+64 values with LDS.128. Per warp and grid:
// 32 LDGSTS + 32 x 32 LDS.128 instead of 512 LDG.256-lanes.
#ifndef GF_EVAL_BSPLINE_CUH_
#define GF_EVAL_BSPLINE_CUH_

#include "gf_eval_lines.cuh"

namespace gfb {

constexpr int kBsBlock = 64;                 // 2 warps x 16 KB of tiles: 6 blocks (12 warps, 198 KB of smem) per SM
constexpr unsigned kBsWarpBytes = 32 * 512;  // 32 atoms x 4 tiles x 128 bytes

// Per-atom interpolation weights, computed once and used for every grid (all grids share the geometry).
struct BsWeights {
    double bx[4], by[4], wz[8];            // value path, FP64; z folded into 8 zero-padded weights (bspline_interpolate)
    float dbx[4], dby[4], dwz[8];          // gradient path, FP32
};
__device__ __forceinline__ void bspline_weights(double fx, double fy, double fz, int off, BsWeights& w) {
    double d[4], bz[4], dbz[4];
    bspline_basis(fx, w.bx, d);   // :741-748
#pragma unroll
    for (int k = 0; k < 4; k++) w.dbx[k] = (float) d[k];
    bspline_basis(fy, w.by, d);
#pragma unroll
    for (int k = 0; k < 4; k++) w.dby[k] = (float) d[k];
    bspline_basis(fz, bz, dbz);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int m = k - off;
        w.wz[k] = m == 0 ? bz[0] : m == 1 ? bz[1] : m == 2 ? bz[2] : m == 3 ? bz[3] : 0.0;
        w.dwz[k] = (float) (m == 0 ? dbz[0] : m == 1 ? dbz[1] : m == 2 ? dbz[2] : m == 3 ? dbz[3] : 0.0);
    }
}

// One lane's stencil out of its 512-byte smem region (4 tiles of 4 rows x 8 floats; granule c of a tile stored at
// c ^ (lane & 7)). Value FP64 from the FP32 points, gradient FP32.
__device__ __forceinline__ void bspline_from_smem(unsigned rbase, unsigned sw, const BsWeights& w, double& val, float& gx,
                                                  float& gy, float& gz) {
    val = 0.0;
    gx = gy = gz = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double pv = 0.0;
        float pdy = 0.f, pdz = 0.f;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            float v[8];
            lds128(rbase + 128u * i + (((2u * r) << 4) ^ sw), v);
            lds128(rbase + 128u * i + (((2u * r + 1u) << 4) ^ sw), v + 4);
            double rz = 0.0;
            float drz = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                rz = fma(w.wz[k], (double) v[k], rz);
                drz = fmaf(w.dwz[k], v[k], drz);
            }
            pv = fma(w.by[r], rz, pv);
            pdy = fmaf(w.dby[r], (float) rz, pdy);
            pdz = fmaf((float) w.by[r], drz, pdz);
        }
        val = fma(w.bx[i], pv, val);
        gx = fmaf(w.dbx[i], (float) pv, gx);
        gy = fmaf((float) w.bx[i], pdy, gy);
        gz = fmaf((float) w.bx[i], pdz, gz);
    }
}

//   FMODE  gfb_force_mode;  SINGLE  one replica and no energy slots (block-level energy reduction)
template <int FMODE, bool SINGLE>
__global__ void __launch_bounds__(kBsBlock, 6) gf_eval_bspline_kernel(const __grid_constant__ EvalParams p) {
    __shared__ __align__(128) unsigned char s_tiles[(kBsBlock / 32) * kBsWarpBytes];

    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    const unsigned t = blockIdx.x * kBsBlock + tid;
    const unsigned total = (unsigned) p.total;
    const bool active = t < total;
    if (p.energies_clear && t < (unsigned) (p.n_replicas * p.n_slots)) p.energies_clear[t] = 0.0;

    // ---- who am I (as gf_eval_lines_kernel) ----------------------------------------------------------------------
    unsigned rep = 0, ia = t;
    if (!SINGLE) {
        rep = __umulhi(t, p.div_magic);
        ia = t - rep * (unsigned) p.n_atoms;
        if (ia >= (unsigned) p.n_atoms) {
            ia -= (unsigned) p.n_atoms;
            rep++;
        }
    }
    if (!active) ia = 0;
    const bool plain = p.particles == nullptr && p.n_particles == p.n_atoms;   // uniform
    unsigned gidx = t;
    if (!plain) gidx = rep * (unsigned) p.n_particles + (p.particles ? (unsigned) p.particles[ia] : ia);
    int key = -1;
    if (active) key = p.slots ? (int) rep * p.n_slots + p.slots[ia] : (int) rep;

    // ---- position (three streaming loads; a warp's positions are 768 contiguous bytes when no indirection is in play)
    double x = 0.0, y = 0.0, z = 0.0;
    if (active) {
        const double* mine = p.pos + 3 * (size_t) gidx;
        x = load_stream(mine);
        y = load_stream(mine + 1);
        z = load_stream(mine + 2);
    }

    // ---- classification (:687-715), bit-exact; all grids share grid 0's geometry ------------------------------------
    const GridView& G = p.grid[0];
    const FastCell fc = classify_fast(G, p.near_int, x, y, z, active);
    const bool inside = fc.inside;
    const int tc = fc.iz / 5, off = fc.iz - 5 * tc;
    const unsigned tile0 = inside ? (unsigned) ((fc.ix * G.nc[1] + fc.iy) * G.row_chunks + tc) : 0xffffffffu;
    const unsigned plane_tiles = (unsigned) (G.nc[1] * G.row_chunks);   // tiles from one x-plane to the next

    const unsigned warp_base = (unsigned) __cvta_generic_to_shared(s_tiles) + (tid >> 5) * kBsWarpBytes;
    const unsigned gran = lane & 7u, octet = lane >> 3;
    const unsigned rbase = warp_base + lane * 512u;
    const unsigned sw = (lane & 7u) << 4;

    BsWeights wts;
    bspline_weights(fc.fx, fc.fy, fc.fz, off, wts);

    double e_total = 0.0;
    double Fx = 0.0, Fy = 0.0, Fz = 0.0;
    unsigned heads = 0;
    unsigned span = 0;
    bool head = false;
    if (!SINGLE) {
        span = run_span(key, lane, heads);
        head = key >= 0 && ((heads >> lane) & 1u);
    }

    for (int g = 0; g < p.n_grids; g++) {
        const GridView& Gg = p.grid[g];
        const double s = active ? Gg.scaling[ia] : 0.0;
        const bool interp = inside && s != 0.0;   // :706
        // ---- fetch: round i brings the four tiles (x-planes) of lane i's atom, one octet per tile -----------------
        const unsigned mytile = interp ? tile0 : 0xffffffffu;
        const char* lane_base = static_cast<const char*>(Gg.cells) + 16u * gran + 128ull * (unsigned long long) octet * plane_tiles;
        __syncwarp();   // the previous grid's tiles have been consumed
#pragma unroll 8
        for (int i = 0; i < 32; i++) {
            const unsigned tl = __shfl_sync(kFull, mytile, i);
            if (tl != 0xffffffffu)
                cp_async16(warp_base + (unsigned) i * 512u + octet * 128u + ((gran ^ ((unsigned) i & 7u)) << 4), lane_base + 128ull * tl);
        }
        cp_async_wait_all();
        __syncwarp();
        // ---- evaluate -----------------------------------------------------------------------------------------------
        double e_g = 0.0;
        if (interp) {
            double val;
            float dx, dy, dz;
            bspline_from_smem(rbase, sw, wts, val, dx, dy, dz);
            double gx = (double) (dx * (float) Gg.inv_spacing[0]);   // :790
            double gy = (double) (dy * (float) Gg.inv_spacing[1]);
            double gz = (double) (dz * (float) Gg.inv_spacing[2]);
            if (Gg.inv_power > 0.0) {   // :778-787
                const double base = val;
                val = pow(base, Gg.inv_power);
                const double pf = Gg.inv_power * pow(base, Gg.inv_power - 1.0);
                gx *= pf;
                gy *= pf;
                gz *= pf;
            }
            e_g = s * val;   // :793
            Fx -= s * gx;    // :794
            Fy -= s * gy;
            Fz -= s * gz;
        } else if (active) {   // :1093-1117
            accumulate_restraint(Gg, x, y, z, e_g, Fx, Fy, Fz);
        }
        e_total += e_g;
        if (p.grid_energies) {   // uniform branch
            double eg = e_g;
            if (SINGLE) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) eg += __shfl_xor_sync(kFull, eg, o);
                if (lane == 0) red_add_f64(p.grid_energies + g, eg);
            } else {
                run_sum(eg, span);
                if (head) red_add_f64(p.grid_energies + (size_t) key * p.n_grids + g, eg);
            }
        }
    }

    // ---- forces --------------------------------------------------------------------------------------------------------
    if (active && p.forces) {
        if (FMODE == GFB_FORCE_FIXED_ADD) {   // OpenMM's 2^32 fixed point, gridForce.cu:487-499
            unsigned long long* f = static_cast<unsigned long long*>(p.forces);
            const double scale = 4294967296.0;
            red_add_u64(f + gidx, (unsigned long long) (long long) (Fx * scale));
            red_add_u64(f + p.force_stride + gidx, (unsigned long long) (long long) (Fy * scale));
            red_add_u64(f + 2 * p.force_stride + gidx, (unsigned long long) (long long) (Fz * scale));
        } else {
            double* f = static_cast<double*>(p.forces) + 3 * (size_t) gidx;
            if (FMODE == GFB_FORCE_F64_STORE) {
                f[0] = Fx;
                f[1] = Fy;
                f[2] = Fz;
            } else {
                red_add_f64(f, Fx);
                red_add_f64(f + 1, Fy);
                red_add_f64(f + 2, Fz);
            }
        }
    }

    // ---- energies ------------------------------------------------------------------------------------------------------
    if (p.energies) {   // uniform branch
        if (SINGLE) {
            __shared__ double warp_sum[kBsBlock / 32];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e_total += __shfl_xor_sync(kFull, e_total, o);
            if (lane == 0) warp_sum[tid >> 5] = e_total;
            __syncthreads();
            if (tid == 0) {
                double b = 0.0;
#pragma unroll
                for (int w = 0; w < kBsBlock / 32; w++) b += warp_sum[w];
                if (p.energy_store) *p.energies = b;
                else red_add_f64(p.energies, b);
            }
        } else {
            run_sum(e_total, span);
            if (head) red_add_f64(p.energies + key, e_total);
        }
    }
}

}  // namespace gfb
#endif
