// Kernels around the evaluation: repack/interleave (once per grid), inv-power transformation, grid generation,
// classification-only (parity tests), Morton keys, fixed-point conversion, gather microbenchmarks.
// Included by gf_misc.cu only (non-template __global__ functions: one definition per library).
#ifndef GF_MISC_KERNELS_CUH_
#define GF_MISC_KERNELS_CUH_

#include "gf_eval_lines.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// Repack: x-major doubles (GridData.h:96-98) -> cell-major packed corners. One thread per cell.
// Runs once per grid in gfb_grid_create (the analogue of the reference's float upload).
// ------------------------------------------------------------------------------------------------
template <typename S>
static __global__ void __launch_bounds__(256) gf_repack_kernel(const double* __restrict__ vals, S* __restrict__ cells,
                                                        int nx, int ny, int nz) {
    const int ncx = nx - 1, ncy = ny - 1, ncz = nz - 1;
    const size_t ncell = (size_t) ncx * ncy * ncz;
    for (size_t c = (size_t) blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (size_t) gridDim.x * blockDim.x) {
        const int iz = (int) (c % ncz);
        const size_t r = c / ncz;
        const int iy = (int) (r % ncy);
        const int ix = (int) (r / ncy);
        const size_t im = ((size_t) ix * ny + iy) * nz + iz;   // :1022
        const size_t nyz = (size_t) ny * nz;
        S* o = cells + 8 * c;
        o[0] = (S) vals[im];
        o[1] = (S) vals[im + 1];
        o[2] = (S) vals[im + nz];
        o[3] = (S) vals[im + nz + 1];
        o[4] = (S) vals[im + nyz];
        o[5] = (S) vals[im + nyz + 1];
        o[6] = (S) vals[im + nyz + nz];
        o[7] = (S) vals[im + nyz + nz + 1];
    }
}

// Interleave: the packed cells of 2-4 grids that share a geometry are woven into one record per cell of 4 slots
// (MIXED: 4 x 32 B = 128 B = one L2/HBM line; DOUBLE: 4 x 64 B = 256 B = two lines), so that the stencils one atom needs
// from all its grids come from ONE record instead of k lines in k arrays. One thread per 16 bytes:
// (cell, slot, part), parts = 2 (MIXED) or 4 (DOUBLE) per slot.
static __global__ void __launch_bounds__(256) gf_interleave_cells_kernel(const float4* __restrict__ s0, const float4* __restrict__ s1,
                                                                  const float4* __restrict__ s2, const float4* __restrict__ s3,
                                                                  float4* __restrict__ dst, size_t n_cells, int slots, int parts) {
    const size_t total = n_cells * slots * parts;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const int part = (int) (i % parts);
        const size_t r = i / parts;
        const int slot = (int) (r % slots);
        const size_t cell = r / slots;
        const float4* src = slot == 0 ? s0 : slot == 1 ? s1 : slot == 2 ? s2 : s3;
        dst[i] = src ? src[cell * parts + part] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ROWS: one thread per (row, chunk). Chunk j of row (ix,iy) = values z = j*(W-1) .. j*(W-1)+W-1 (zero past the row).
template <typename S>
static __global__ void __launch_bounds__(256) gf_repack_rows_kernel(const double* __restrict__ vals, S* __restrict__ out,
                                                             int nx, int ny, int nz, int row_chunks) {
    constexpr int W = 32 / (int) sizeof(S);
    const size_t total = (size_t) nx * ny * row_chunks;
    for (size_t c = (size_t) blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (size_t) gridDim.x * blockDim.x) {
        const int j = (int) (c % row_chunks);
        const size_t row = c / row_chunks;
        const double* src = vals + row * nz;
        S* o = out + c * W;
#pragma unroll
        for (int k = 0; k < W; k++) {
            const int z = j * (W - 1) + k;
            o[k] = z < nz ? (S) src[z] : (S) 0;
        }
    }
}

// PAIRS (float): one thread per (ix, iy < ny-1, j): {row iy: z=3j..3j+3, row iy+1: z=3j..3j+3}.
static __global__ void __launch_bounds__(256) gf_repack_pairs_kernel(const double* __restrict__ vals, float* __restrict__ out,
                                                              int nx, int ny, int nz, int row_chunks) {
    const size_t total = (size_t) nx * (ny - 1) * row_chunks;
    for (size_t c = (size_t) blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (size_t) gridDim.x * blockDim.x) {
        const int j = (int) (c % row_chunks);
        const size_t r = c / row_chunks;
        const int iy = (int) (r % (ny - 1));
        const int ix = (int) (r / (ny - 1));
        const double* src = vals + ((size_t) ix * ny + iy) * nz;
        float* o = out + c * 8;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int z = 3 * j + k;
            o[k] = z < nz ? (float) src[z] : 0.f;
            o[4 + k] = z < nz ? (float) src[nz + z] : 0.f;
        }
    }
}

// BSPLINE records (see bspline_interpolate): one thread per (record, half h, row r): the 4 values
// P[a+h][iy+r][iz+k] = V[clamp(a+h-1)][clamp(iy+r-1)][clamp(iz+k-1)], k = 0..3. a < nx+1, iy < ny-1, iz < nz-1.
// FLAT (HERMITE records, tricubic_interpolate): the same records with P[a][b][c] = the value at flat index
// ((a-1)*ny + (b-1))*nz + (c-1), 0 outside the array — how the reference's tricubic branch addresses its neighbours.
template <typename S, bool FLAT>
static __global__ void __launch_bounds__(256) gf_repack_bspline_kernel(const double* __restrict__ vals, S* __restrict__ out,
                                                                int nx, int ny, int nz) {
    const size_t total = (size_t) (nx + 1) * (ny - 1) * (nz - 1) * 8;
    for (size_t c = (size_t) blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (size_t) gridDim.x * blockDim.x) {
        const int r = (int) (c & 3);
        const int h = (int) ((c >> 2) & 1);
        size_t t = c >> 3;
        const int iz = (int) (t % (nz - 1));
        t /= (nz - 1);
        const int iy = (int) (t % (ny - 1));
        const int a = (int) (t / (ny - 1));
        S* o = out + c * 4;
        if (FLAT) {
            const long long n = (long long) nx * ny * nz;
            const long long base = ((long long) (a + h - 1) * ny + (iy + r - 1)) * nz + (iz - 1);
#pragma unroll
            for (int k = 0; k < 4; k++) o[k] = (base + k >= 0 && base + k < n) ? (S) vals[base + k] : (S) 0;
            continue;
        }
        const int gx = min(max(a + h - 1, 0), nx - 1);
        const int gy = min(max(iy + r - 1, 0), ny - 1);
        const double* src = vals + ((size_t) gx * ny + gy) * nz;
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = (S) src[min(max(iz + k - 1, 0), nz - 1)];
    }
}

// POINTS (see tricubic_interpolate): the values as they are, in S, then n_guard zeros.
template <typename S>
static __global__ void __launch_bounds__(256) gf_repack_points_kernel(const double* __restrict__ vals, S* __restrict__ out,
                                                               size_t n_points, size_t n_guard) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_points + n_guard; i += (size_t) gridDim.x * blockDim.x)
        out[i] = i < n_points ? (S) vals[i] : (S) 0;
}

// GridForce::applyInvPowerTransformation (openmmapi/src/GridForce.cpp:262-268; CachedGridData.cpp:50-57): the RUNTIME
// inv-power mode stores G -> sign(G) * |G|^(1/n) once, and the evaluation applies ^n. In place, FP64.
static __global__ void __launch_bounds__(256) gf_inv_power_transform_kernel(double* __restrict__ vals, size_t n, double inv_n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const double v = vals[i];
        if (v != 0.0) vals[i] = (v >= 0.0 ? 1.0 : -1.0) * pow(fabs(v), inv_n);
    }
}

// ------------------------------------------------------------------------------------------------
// Grid generation from receptor atoms (reference: ReferenceGridForceKernels.cpp:465-544; the reference's own GPU
// version, platforms/cuda/src/kernels/gridGeneration.cu:198-371, is FP32). One thread per grid point, FP64 throughout,
// atoms streamed through shared memory in tiles of 256 {x, y, z, coefficient} (the N-body pattern: every atom is read
// once per block, from smem, by all 256 threads). Per atom the host has folded the parameters into one coefficient c
// exactly as the reference's expression associates: charge c = 138.935456*q, ljr c = sqrt(eps)*(2 sigma)^6,
// lja c = -2*sqrt(eps)*(2 sigma)^3; the term is c / r^P (P = 1, 12, 6) with r clamped to >= 1e-6 nm (:520-522),
// formed from rsqrt(r2) and multiplications (no FP64 division or pow in the inner loop). Atoms are summed in index
// order, like the reference; the result is capped with U*tanh(v/U) (:540).
// Compute-bound: ~20 FP64 instructions per (point, atom) pair.
// ------------------------------------------------------------------------------------------------
template <int P>
static __global__ void __launch_bounds__(256) gf_generate_grid_kernel(const double4* __restrict__ atoms, int n_atoms, int nx, int ny, int nz,
                                                               double ox, double oy, double oz, double sx, double sy, double sz,
                                                               double cap, double* __restrict__ out) {
    __shared__ double4 tile[256];
    const size_t n_points = (size_t) nx * ny * nz;
    const size_t idx = (size_t) blockIdx.x * 256 + threadIdx.x;
    const bool live = idx < n_points;
    const size_t pt = live ? idx : n_points - 1;
    const int k = (int) (pt % nz);
    const size_t r = pt / nz;
    const int j = (int) (r % ny);
    const int i = (int) (r / ny);
    const double gx = ox + i * sx, gy = oy + j * sy, gz = oz + k * sz;      // :502-504
    double v = 0.0;
    for (int base = 0; base < n_atoms; base += 256) {
        const int m = min(256, n_atoms - base);
        if ((int) threadIdx.x < m) tile[threadIdx.x] = atoms[base + threadIdx.x];
        __syncthreads();
#pragma unroll 4
        for (int a = 0; a < m; a++) {
            const double4 at = tile[a];
            const double dx = gx - at.x, dy = gy - at.y, dz = gz - at.z;
            const double r2 = dx * dx + dy * dy + dz * dz;
            const double rinv = fmin(rsqrt(r2), 1.0e6);                     // r = max(sqrt(r2), 1e-6)
            double t;
            if (P == 1) {
                t = rinv;
            } else {
                const double i2 = rinv * rinv, i6 = i2 * i2 * i2;
                t = P == 6 ? i6 : i6 * i6;
            }
            v += at.w * t;
        }
        __syncthreads();
    }
    if (live) out[idx] = cap * tanh(v / cap);
}

// Classification only (parity tests): same device function as the evaluation.
template <bool EXACT>
static __global__ void __launch_bounds__(256) gf_classify_kernel(const __grid_constant__ ClassifyParams p) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.total) return;
    const int rep = (int) (t / p.n_atoms);
    const int ia = (int) (t - (long long) rep * p.n_atoms);
    const int particle = p.particles ? p.particles[ia] : ia;
    const double* pp = p.pos + 3 * ((long long) rep * p.n_particles + particle);
    const AtomCell c = classify<EXACT>(p.grid, pp[0], pp[1], pp[2]);
    const double s = p.grid.scaling[ia];
    gfb_class out;
    out.inside = c.inside ? 1 : 0;
    const bool interp = c.inside && s != 0.0;
    out.cell[0] = interp ? c.ix : -1;
    out.cell[1] = interp ? c.iy : -1;
    out.cell[2] = interp ? c.iz : -1;
    p.out[t] = out;
}

// ------------------------------------------------------------------------------------------------
// Atom sort (gfb_kernel_sort_atoms): a one-pass counting sort by BRICK key. A brick is a cube of 2^shift cells per
// axis of grid 0; the key is the Morton interleave of the brick coordinates (`bits` bits per axis), atoms outside the
// grid take the last bin. Three small kernels, no library: count (keys + histogram), scan (exclusive prefix over the
// bins, one block), scatter (slot = atomicAdd on the bin's cursor). Atoms of one brick end up adjacent in an arbitrary
// order — which is all the locality the evaluation can use (neighbouring lanes then read neighbouring lines).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned spread10(unsigned v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

static __global__ void __launch_bounds__(256) gf_sort_count_kernel(const __grid_constant__ ClassifyParams p, int shift, unsigned n_bins,
                                                            unsigned* __restrict__ keys, unsigned* __restrict__ hist) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.total) return;
    const int rep = (int) (t / p.n_atoms);
    const int ia = (int) (t - (long long) rep * p.n_atoms);
    const int particle = p.particles ? p.particles[ia] : ia;
    const double* pp = p.pos + 3 * ((long long) rep * p.n_particles + particle);
    const AtomCell c = classify<false>(p.grid, pp[0], pp[1], pp[2]);
    unsigned key = n_bins - 1;   // outside atoms sort last
    if (c.inside)
        key = (spread10((unsigned) c.ix >> shift) << 2) | (spread10((unsigned) c.iy >> shift) << 1) | spread10((unsigned) c.iz >> shift);
    keys[t] = key;
    atomicAdd(hist + key, 1u);
}

// Exclusive prefix sum over n bins in place, one block of 1024 threads: chunk sums, block scan of the 1024 partials,
// chunk prefixes. n <= 2^21 + 1 bins: a few microseconds.
static __global__ void __launch_bounds__(1024) gf_sort_scan_kernel(unsigned* __restrict__ hist, unsigned n) {
    __shared__ unsigned part[1024];
    const unsigned tid = threadIdx.x;
    const unsigned chunk = (n + 1023u) / 1024u;
    const unsigned lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    unsigned sum = 0;
    for (unsigned i = lo; i < hi; i++) sum += hist[i];
    part[tid] = sum;
    __syncthreads();
    for (unsigned off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive scan
        const unsigned v = tid >= off ? part[tid - off] : 0u;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    unsigned run = part[tid] - sum;   // exclusive prefix of this chunk
    for (unsigned i = lo; i < hi; i++) {
        const unsigned c = hist[i];
        hist[i] = run;
        run += c;
    }
}

static __global__ void __launch_bounds__(256) gf_sort_scatter_kernel(const unsigned* __restrict__ keys, unsigned* __restrict__ cursor,
                                                              long long total, int* __restrict__ order) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    order[atomicAdd(cursor + keys[t], 1u)] = (int) t;
}

// Fixed-point (OpenMM long force buffer) -> double [n][3].
static __global__ void __launch_bounds__(256) gf_fixed_to_f64_kernel(const long long* __restrict__ fixed, long long stride,
                                                              long long n, double* __restrict__ out) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double inv = 1.0 / 4294967296.0;
    out[3 * t] = (double) fixed[t] * inv;
    out[3 * t + 1] = (double) fixed[stride + t] * inv;
    out[3 * t + 2] = (double) fixed[2 * stride + t] * inv;
}

// ------------------------------------------------------------------------------------------------
// Roofline denominator: random 32-byte-sector gather. Every lane of every warp reads a different
// pseudo-random sector of `buf` (n_sectors of them) with the same LDG.E.256 the evaluation uses.
// ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) gf_sector_gather_kernel(const float* __restrict__ buf, unsigned long long n_sectors,
                                                               int loads_per_thread, float* __restrict__ sink) {
    unsigned long long h = ((unsigned long long) blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < loads_per_thread; i++) {
        h ^= h >> 29;
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 32;
        const unsigned long long sector = __umul64hi(h, n_sectors);  // uniform in [0, n_sectors)
        float v[8];
        load_cell(buf + 8 * sector, v);
        acc += v[0] + v[3] + v[5] + v[7];
    }
    if (acc == 123.456f) sink[0] = acc;  // keeps the loads alive; practically never taken
}

// Same, with 16-byte (LDG.E.128) loads: the unit of the row-chunked layouts.
static __global__ void __launch_bounds__(256) gf_chunk_gather_kernel(const float4* __restrict__ buf, unsigned long long n_chunks,
                                                              int loads_per_thread, float* __restrict__ sink) {
    unsigned long long h = ((unsigned long long) blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < loads_per_thread; i++) {
        h ^= h >> 29;
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 32;
        const float4 v = __ldg(buf + __umul64hi(h, n_chunks));
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) sink[0] = acc;
}


// Classification only, through classify_fast (parity tests of the lines kernel's index math).
static __global__ void __launch_bounds__(256) gf_classify_lines_kernel(const __grid_constant__ ClassifyParams p) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.total) return;
    const int rep = (int) (t / p.n_atoms);
    const int ia = (int) (t - (long long) rep * p.n_atoms);
    const int particle = p.particles ? p.particles[ia] : ia;
    const double* pp = p.pos + 3 * ((long long) rep * p.n_particles + particle);
    const FastCell c = classify_fast(p.grid, p.near_int, pp[0], pp[1], pp[2], true);
    const double s = p.grid.scaling[ia];
    gfb_class out;
    out.inside = c.inside ? 1 : 0;
    const bool interp = c.inside && s != 0.0;
    out.cell[0] = interp ? c.ix : -1;
    out.cell[1] = interp ? c.iy : -1;
    out.cell[2] = interp ? c.iz : -1;
    p.out[t] = out;
}

}  // namespace gfb
#endif
