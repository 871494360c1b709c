// The evaluation kernel of the named configurations: MIXED precision, packed cells, grids of one geometry, no inv-power.
//
// What differs from the general kernel (gf_kernels.cuh, gf_eval_kernel), measured reasons in DESIGN.md §4:
//   * LINES. HBM and L2 move 128-byte lines, a stencil is 32 bytes. With 2-4 grids the packed cells of all grids are
//     woven into ONE 128-byte record per cell (gf_interleave_cells_kernel), so everything an atom needs is one line,
//     and a warp fetches the lines of its 32 atoms with 4 warp-wide LDG.E.256 in which the four lanes of a quad read
//     the four 32-byte sectors of the SAME line (8 lines per instruction: 8 L1 wavefronts, one L2 request per line)
//     instead of 3 x 32 single-sector requests to 96 different lines. The sectors travel to the lane that owns the
//     atom through a swizzled, conflict-free shared-memory transpose (8 STS.128 + 2*NG LDS.128 per lane).
//   * INSTRUCTIONS. The general kernel issues 693 warp instructions per 32 atoms (ncu, r1), 42 % of the issue slots
//     at its HBM-bound 137 us; once the line traffic is halved that would bind. Here: replica/atom split by a host
//     magic multiplier, one 32-bit cell index, one near-integer test on the fractions that are needed anyway, gradient
//     scaled by 1/spacing once per atom instead of once per grid, FP32 force accumulation, no pow code, restraint
//     and exact re-division out of line.
//   * FORCES. Optional plain read-modify-write of the fixed-point force words, with the read issued together with the
//     position fetch (FPATH 2), or an L2 prefetch of the force lines at that point (FPATH 1), instead of RED alone.
//
// Reference semantics implemented (platforms/reference/src/ReferenceGridForceKernels.cpp): inside test :687-696,
// cell index/fraction :708-715 (bit-exact), trilinear value z->y->x :1039-1053, gradient :1066-1072, scaling and
// accumulation :1061-1063/:1082, restraint :1093-1117.
#ifndef GF_EVAL_LINES_CUH_
#define GF_EVAL_LINES_CUH_

#include "gf_kernels.cuh"

namespace gfb {

// Number of lanes that follow `lane` inside its run of equal keys (runs = maximal stretches of consecutive lanes with
// the same key). `heads` gets the ballot of run heads. Equal keys that are not adjacent form separate runs, which is
// still correct: each run issues its own atomic.
__device__ __forceinline__ unsigned run_span(int key, unsigned lane, unsigned& heads) {
    const int kprev = __shfl_up_sync(kFull, key, 1);
    heads = __ballot_sync(kFull, lane == 0 || kprev != key);
    const unsigned above = (heads >> 1) >> lane;   // bit i: lane+1+i starts a new run
    return above ? (unsigned) __ffs((int) above) - 1u : 31u - lane;
}

// Segmented sum over a run; the total lands in the run's first lane.
__device__ __forceinline__ void run_sum(double& e, unsigned span) {
#pragma unroll
    for (unsigned off = 1; off < 32; off <<= 1) {
        const double ev = __shfl_down_sync(kFull, e, off);
        if (off <= span) e += ev;
    }
}

// Exact IEEE re-division of all three axes, taken when one fast quotient lies within rounding distance of an integer
// (probability ~1e-12 per atom) or on the upper face. Out of line: the FP64 division sequence must not be if-converted
// into the main path.
struct ExactCell {
    int ix, iy, iz;
    double fx, fy, fz;
};
__device__ __noinline__ ExactCell exact_cell(const GridView& G, double px, double py, double pz) {
    ExactCell c;
    const double qx = px / G.spacing[0], qy = py / G.spacing[1], qz = pz / G.spacing[2];
    c.ix = min(__double2int_rz(qx), G.nc[0] - 1);
    c.iy = min(__double2int_rz(qy), G.nc[1] - 1);
    c.iz = min(__double2int_rz(qz), G.nc[2] - 1);
    c.fx = qx - (double) c.ix;
    c.fy = qy - (double) c.iy;
    c.fz = qz - (double) c.iz;
    return c;
}

// Restraint of an atom outside the (shared) grid box for all NG GridForces (:1093-1117): every force adds its own
// harmonic wall with its own constant. Rare; out of line.
template <int NG>
struct RestraintAll {
    double e[NG];
    float fx, fy, fz;
};
template <int NG>
__device__ __noinline__ RestraintAll<NG> restraint_all(const EvalParams& p, double x, double y, double z) {
    RestraintAll<NG> r;
    double fx = 0.0, fy = 0.0, fz = 0.0;
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const Restraint t = restraint_terms(p.grid[g], x, y, z);
        r.e[g] = t.e;
        fx += t.fx;
        fy += t.fy;
        fz += t.fz;
    }
    r.fx = (float) fx;
    r.fy = (float) fy;
    r.fz = (float) fz;
    return r;
}

__device__ __forceinline__ void sts128(unsigned addr, const float* v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ void lds128(unsigned addr, float* v) {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr) : "memory");
}
// One 32-byte sector of a record that no other atom of the launch is likely to read again: no L1 allocation.
__device__ __forceinline__ void load_sector(const char* p, float v[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

constexpr int kForceRed = 0, kForcePrefetch = 1, kForceRmw = 2;

#ifndef GF_LINES_ASYNC_BLOCKS
#define GF_LINES_ASYNC_BLOCKS 5
#endif

__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

//   NG     grids evaluated per atom, 1..4 (1: the grid's own packed cells; 2..4: 128-byte records of 4 slots)
//   FMODE  gfb_force_mode
//   FPATH  kForceRed | kForcePrefetch | kForceRmw (ADD modes only)
//   SINGLE one replica and no energy slots: block-level energy reduction, one atomic per block
//   GE     per-grid energies wanted (p.grid_energies): a template flag so that the per-grid terms cost no registers
//          in the common case
//   ASYNC  NG > 1: records go global -> shared with cp.async (no register staging) and each grid's corners are read from
//          shared memory right before they are used, which is what lets 5-6 blocks share an SM
template <int NG, int FMODE, int FPATH, bool SINGLE, bool GE, bool ASYNC>
__global__ void __launch_bounds__(kBlock, NG == 1 ? 6 : (ASYNC ? GF_LINES_ASYNC_BLOCKS : 4))
    gf_eval_lines_kernel(const __grid_constant__ EvalParams p) {
    __shared__ __align__(16) double2 s_pos2[kBlock * 3 / 2];
    __shared__ __align__(128) float4 s_rec[NG == 1 ? 1 : kBlock * 8];   // 128 bytes per atom of the block

    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    const unsigned t0 = blockIdx.x * kBlock;
    const unsigned t = t0 + tid;
    const unsigned total = (unsigned) p.total;
    const bool active = t < total;
    if (p.energies_clear && t < (unsigned) (p.n_replicas * p.n_slots)) p.energies_clear[t] = 0.0;

    // ---- who am I: replica, atom ordinal, particle -------------------------------------------------------------
    unsigned rep = 0, ia = t;
    if (!SINGLE) {   // t / n_atoms by the host's magic multiplier floor(2^32 / n_atoms): estimate is q or q-1
        rep = __umulhi(t, p.div_magic);
        ia = t - rep * (unsigned) p.n_atoms;
        if (ia >= (unsigned) p.n_atoms) {
            ia -= (unsigned) p.n_atoms;
            rep++;
        }
    }
    if (!active) ia = 0;
    const bool plain = p.particles == nullptr && p.n_particles == p.n_atoms;   // uniform
    unsigned gidx = t;                                                         // particle slot in pos / forces
    if (!plain) gidx = rep * (unsigned) p.n_particles + (p.particles ? (unsigned) p.particles[ia] : ia);
    int key = -1;
    if (active) key = p.slots ? (int) rep * p.n_slots + p.slots[ia] : (int) rep;

    // ---- loads that depend on the atom ordinal only go out first -------------------------------------------------
    double sd[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) sd[g] = (active && !(ASYNC && NG > 1)) ? p.grid[g].scaling[ia] : 0.0;

    unsigned long long* const ffix = static_cast<unsigned long long*>(p.forces);
    double* const fdbl = static_cast<double*>(p.forces);
    unsigned long long old_fixed[3] = {0ull, 0ull, 0ull};
    double old_f64[3] = {0.0, 0.0, 0.0};
    if (p.forces && FMODE != GFB_FORCE_F64_STORE) {
        if (FPATH == kForcePrefetch) {
            if (FMODE == GFB_FORCE_FIXED_ADD) {
                if (active && (lane & 15u) == 0) {   // 16 lanes x 8 bytes = one line per plane
                    prefetch_l2(ffix + gidx);
                    prefetch_l2(ffix + p.force_stride + gidx);
                    prefetch_l2(ffix + 2 * p.force_stride + gidx);
                }
            } else if (active && (lane & 3u) == 0) {   // 4 lanes x 24 bytes < one line
                prefetch_l2(fdbl + 3 * (size_t) gidx);
            }
        } else if (FPATH == kForceRmw && active) {
            if (FMODE == GFB_FORCE_FIXED_ADD) {
                old_fixed[0] = ffix[gidx];
                old_fixed[1] = ffix[p.force_stride + gidx];
                old_fixed[2] = ffix[2 * p.force_stride + gidx];
            } else {
                old_f64[0] = fdbl[3 * (size_t) gidx];
                old_f64[1] = fdbl[3 * (size_t) gidx + 1];
                old_f64[2] = fdbl[3 * (size_t) gidx + 2];
            }
        }
    }

    // ---- positions: the block's 256 atoms are 6144 contiguous bytes -> coalesced 16-byte loads through smem --------
    double x = 0.0, y = 0.0, z = 0.0;
    const bool staged = plain && (reinterpret_cast<uintptr_t>(p.pos) & 15) == 0;   // uniform
    if (staged) {
        const double2* src = reinterpret_cast<const double2*>(p.pos + 3 * (size_t) t0);
        if (t0 + kBlock <= total) {   // every block but the last
            double2 a, b = make_double2(0.0, 0.0);
            asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(a.x), "=d"(a.y) : "l"(src + tid));
            if (tid < kBlock / 2) asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(b.x), "=d"(b.y) : "l"(src + kBlock + tid));
            s_pos2[tid] = a;
            if (tid < kBlock / 2) s_pos2[kBlock + tid] = b;
        } else {
            const unsigned left = 3u * (total - t0);   // doubles left in the array
            for (unsigned i = tid; i < kBlock * 3 / 2; i += kBlock) {
                double2 v = make_double2(0.0, 0.0);
                if (2 * i + 1 < left) asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(src + i));
                else if (2 * i < left) v.x = load_stream(reinterpret_cast<const double*>(src + i));
                s_pos2[i] = v;
            }
        }
        __syncthreads();
        const double* mine = reinterpret_cast<const double*>(s_pos2) + 3 * tid;
        x = mine[0];
        y = mine[1];
        z = mine[2];
    } else if (active) {
        const double* mine = p.pos + 3 * (size_t) gidx;
        x = load_stream(mine);
        y = load_stream(mine + 1);
        z = load_stream(mine + 2);
    }

    // ---- classification (:687-715), bit-exact ------------------------------------------------------------------------
    const GridView& G = p.grid[0];
    const double px = x - G.origin[0], py = y - G.origin[1], pz = z - G.origin[2];
    const bool inside = active && (px >= 0.0 && px <= G.hcorner[0]) && (py >= 0.0 && py <= G.hcorner[1]) &&
                        (pz >= 0.0 && pz <= G.hcorner[2]);
    unsigned cell = 0xffffffffu;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    double dfx = 0.0, dfy = 0.0, dfz = 0.0;
    if (inside) {
        const double qx = px * G.inv_spacing[0], qy = py * G.inv_spacing[1], qz = pz * G.inv_spacing[2];
        int ix = __double2int_rz(qx), iy = __double2int_rz(qy), iz = __double2int_rz(qz);
        dfx = qx - (double) ix;
        dfy = qy - (double) iy;
        dfz = qz - (double) iz;
        // The fast quotient is within 3.3e-16*q of the correctly rounded one, so the truncation can only differ when
        // the fraction is that close to 0 or 1; near_int[k] = 1.8e-15 * cells on the axis covers it with margin.
        // (ix == nc: the upper face, fraction 0 -> also taken.)
        const bool near = dfx <= p.near_int[0] || dfx >= 1.0 - p.near_int[0] || dfy <= p.near_int[1] ||
                          dfy >= 1.0 - p.near_int[1] || dfz <= p.near_int[2] || dfz >= 1.0 - p.near_int[2];
        if (near) {
            const ExactCell c = exact_cell(G, px, py, pz);
            ix = c.ix;
            iy = c.iy;
            iz = c.iz;
            dfx = c.fx;
            dfy = c.fy;
            dfz = c.fz;
        }
        cell = ((unsigned) ix * (unsigned) G.nc[1] + (unsigned) iy) * (unsigned) G.nc[2] + (unsigned) iz;
        fx = (float) dfx;
        fy = (float) dfy;
        fz = (float) dfz;
    }

    // ---- stencils + interpolation: gradient FP32, value FP64 (see trilinear_value_f64) -----------------------------------
    double e_g[GE ? NG : 1];
    double e_total = 0.0;
    float sx = 0.f, sy = 0.f, sz = 0.f;   // sum over grids of scaling * (corner-difference gradient), before 1/spacing
    if (GE) {
#pragma unroll
        for (int g = 0; g < (GE ? NG : 1); g++) e_g[g] = 0.0;
    }
    auto one_grid = [&](int g, const float* v, double s) {
        float val, dx, dy, dz;
        trilinear<float>(v, fx, fy, fz, val, dx, dy, dz);
        const float sf = (float) s;
        sx = fmaf(sf, dx, sx);
        sy = fmaf(sf, dy, sy);
        sz = fmaf(sf, dz, sz);
        const double e = s * trilinear_value_f64(v, dfx, dfy, dfz);   // :1061
        e_total += e;
        if (GE) e_g[GE ? g : 0] = e;
    };
    if (NG == 1) {
        if (inside && sd[0] != 0.0) {   // :706
            float v[8];
            load32(static_cast<const float*>(G.cells) + 8 * (size_t) cell, v);
            one_grid(0, v, sd[0]);
        }
    } else if (ASYNC) {
        // The record of atom A (lane A of this warp) lives at warp_base + 128*A, its 16-byte granule c (slot c/2, half
        // c%2) at position c ^ (A & 7). Round i: the eight lanes of an octet copy the eight granules of the record of the
        // atom owned by lane 4i + octet -> 4 full lines per instruction, each written to one 128-byte smem row.
        const unsigned warp_base = (unsigned) __cvta_generic_to_shared(s_rec) + (tid >> 5) * 4096u;
        const unsigned gran = lane & 7u, octet = lane >> 3;
        const char* lane_base = static_cast<const char*>(p.lines) + 16u * gran;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const unsigned A = 4u * i + octet;
            const unsigned c = __shfl_sync(kFull, cell, (int) A);
            if (c != 0xffffffffu && gran < 2u * NG) cp_async16(warp_base + A * 128u + ((gran ^ (A & 7u)) << 4), lane_base + 128ull * c);
        }
        cp_async_wait_all();
        __syncwarp();
        const unsigned rbase = (warp_base + lane * 128u) ^ ((lane & 7u) << 4);   // 128-byte aligned base: xor == or
        if (inside) {
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const double s = p.grid[g].scaling[ia];   // 47 x NG doubles: L1-resident
                if (s != 0.0) {                           // :706
                    float v[8];
                    lds128(rbase ^ (32u * g), v);
                    lds128(rbase ^ (32u * g + 16u), v + 4);
                    one_grid(g, v, s);
                }
            }
        }
    } else {
        const unsigned sub = lane & 3u, quad = lane >> 2;
        const char* lane_base = static_cast<const char*>(p.lines) + 32u * sub;
        float r[4][8];
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {   // round j: quad q reads the record of the atom owned by lane 8j+q
            const unsigned c = __shfl_sync(kFull, cell, 8 * j + (int) quad);
            ok[j] = c != 0xffffffffu && sub < (unsigned) NG;
            if (ok[j]) load_sector(lane_base + 128ull * c, r[j]);
        }
        // smem transpose, same record placement as above: both the quad-wise writes and the per-owner reads touch
        // 8 different granule positions per quarter-warp -> no bank conflicts.
        const unsigned warp_base = (unsigned) __cvta_generic_to_shared(s_rec) + (tid >> 5) * 4096u;
        const unsigned wbase = warp_base + quad * 128u;
        const unsigned w0 = wbase + (((2u * sub) ^ quad) << 4), w1 = wbase + (((2u * sub + 1u) ^ quad) << 4);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (ok[j]) {
                sts128(w0 + 1024u * j, r[j]);
                sts128(w1 + 1024u * j, r[j] + 4);
            }
        __syncwarp();
        const unsigned rbase = (warp_base + lane * 128u) ^ ((lane & 7u) << 4);
        if (inside) {
            float v[NG][8];
#pragma unroll
            for (int g = 0; g < NG; g++) {
                lds128(rbase ^ (32u * g), v[g]);
                lds128(rbase ^ (32u * g + 16u), v[g] + 4);
            }
#pragma unroll
            for (int g = 0; g < NG; g++)
                if (sd[g] != 0.0) one_grid(g, v[g], sd[g]);   // :706
        }
    }
    float Fx = -sx * (float) G.inv_spacing[0];   // :1072, :1082
    float Fy = -sy * (float) G.inv_spacing[1];
    float Fz = -sz * (float) G.inv_spacing[2];
    if (active && !inside) {   // :1093-1117 (inside atoms with scale 0 take that branch too and add exactly 0)
        const RestraintAll<NG> r = restraint_all<NG>(p, x, y, z);
#pragma unroll
        for (int g = 0; g < NG; g++) {
            e_total += r.e[g];
            if (GE) e_g[GE ? g : 0] = r.e[g];
        }
        Fx -= r.fx;
        Fy -= r.fy;
        Fz -= r.fz;
    }

    // ---- forces --------------------------------------------------------------------------------------------------------
    if (active && p.forces) {
        if (FMODE == GFB_FORCE_FIXED_ADD) {   // OpenMM's 2^32 fixed point, gridForce.cu:487-499
            const unsigned long long ax = (unsigned long long) __float2ll_rz(Fx * 4294967296.f);
            const unsigned long long ay = (unsigned long long) __float2ll_rz(Fy * 4294967296.f);
            const unsigned long long az = (unsigned long long) __float2ll_rz(Fz * 4294967296.f);
            if (FPATH == kForceRmw) {
                ffix[gidx] = old_fixed[0] + ax;
                ffix[p.force_stride + gidx] = old_fixed[1] + ay;
                ffix[2 * p.force_stride + gidx] = old_fixed[2] + az;
            } else {
                red_add_u64(ffix + gidx, ax);
                red_add_u64(ffix + p.force_stride + gidx, ay);
                red_add_u64(ffix + 2 * p.force_stride + gidx, az);
            }
        } else {
            double* f = fdbl + 3 * (size_t) gidx;
            if (FMODE == GFB_FORCE_F64_STORE) {
                f[0] = (double) Fx;
                f[1] = (double) Fy;
                f[2] = (double) Fz;
            } else if (FPATH == kForceRmw) {
                f[0] = old_f64[0] + (double) Fx;
                f[1] = old_f64[1] + (double) Fy;
                f[2] = old_f64[2] + (double) Fz;
            } else {
                red_add_f64(f, (double) Fx);
                red_add_f64(f + 1, (double) Fy);
                red_add_f64(f + 2, (double) Fz);
            }
        }
    }

    // ---- energies ------------------------------------------------------------------------------------------------------
    if (SINGLE) {
        __shared__ double warp_sum[kBlock / 32];
        if (GE) {
#pragma unroll
            for (int g = 0; g < (GE ? NG : 1); g++) {
                double eg = e_g[g];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) eg += __shfl_xor_sync(kFull, eg, off);
                if (lane == 0) red_add_f64(p.grid_energies + g, eg);
            }
        }
        if (p.energies) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) e_total += __shfl_xor_sync(kFull, e_total, off);
            if (lane == 0) warp_sum[tid >> 5] = e_total;
            __syncthreads();
            if (tid < 32) {
                double b = tid < kBlock / 32 ? warp_sum[tid] : 0.0;
#pragma unroll
                for (int off = kBlock / 64; off > 0; off >>= 1) b += __shfl_xor_sync(kFull, b, off);
                if (tid == 0) red_add_f64(p.energies, b);
            }
        }
    } else if (p.energies || GE) {
        unsigned heads;
        const unsigned span = run_span(key, lane, heads);
        const bool head = key >= 0 && ((heads >> lane) & 1u);
        if (GE) {
#pragma unroll
            for (int g = 0; g < (GE ? NG : 1); g++) {
                double eg = e_g[g];
                run_sum(eg, span);
                if (head) red_add_f64(p.grid_energies + (size_t) key * NG + g, eg);
            }
        }
        if (p.energies) {
            run_sum(e_total, span);
            if (head) red_add_f64(p.energies + key, e_total);
        }
    }
}

}  // namespace gfb
#endif
