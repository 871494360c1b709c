// Grids: upload + on-device repack into a layout, V3 OMGRID files, generation from receptor atoms, inv-power
// transformation (the gfb_grid_* / gfb_gridfile_* / gfb_inv_power_transform entry points of include/gridforce_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "gf_handles.h"
#include "gf_gridfile.h"
#include "gf_misc_kernels.cuh"

using namespace gfb;

static int grid_create_common(gfb_device* dev, const int counts[3], const double spacing[3], const double origin[3],
                              const double* vals, bool vals_on_device, size_t n_vals, int precision, int layout,
                              gfb_grid** out) {
    if (!dev || !counts || !spacing || !origin || !vals || !out) return fail(GFB_ERR_INVALID, "gfb_grid_create: NULL argument");
    *out = nullptr;
    if (precision != GFB_PRECISION_MIXED && precision != GFB_PRECISION_DOUBLE)
        return fail(GFB_ERR_INVALID, "gfb_grid_create: unknown precision %d", precision);
    if (layout < GFB_LAYOUT_AUTO || layout > GFB_LAYOUT_BSPLINE_POINTS) return fail(GFB_ERR_INVALID, "gfb_grid_create: unknown layout %d", layout);
    if (layout == GFB_LAYOUT_PAIRS && precision == GFB_PRECISION_DOUBLE)
        return fail(GFB_ERR_UNSUPPORTED, "gfb_grid_create: the PAIRS layout exists for MIXED precision only (use ROWS or CELLS)");
    for (int k = 0; k < 3; k++) {
        if (counts[k] < 2) return fail(GFB_ERR_INVALID, "gfb_grid_create: counts[%d]=%d, need >= 2 points per axis", k, counts[k]);
        if (!(spacing[k] > 0.0) || !std::isfinite(spacing[k]))
            return fail(GFB_ERR_INVALID, "gfb_grid_create: spacing[%d]=%g must be positive and finite", k, spacing[k]);
    }
    const size_t n_points = (size_t) counts[0] * counts[1] * counts[2];
    if (n_vals != n_points)
        return fail(GFB_ERR_INVALID, "gfb_grid_create: %zu values given for a %dx%dx%d grid (%zu points)", n_vals, counts[0],
                    counts[1], counts[2], n_points);
    CUDA_TRY(cudaSetDevice(dev->ordinal));

    gfb_grid* g = new (std::nothrow) gfb_grid();
    if (!g) return fail(GFB_ERR_NOMEM, "gfb_grid_create: out of host memory");
    g->dev = dev;
    g->precision = precision;
    for (int k = 0; k < 3; k++) {
        g->counts[k] = counts[k];
        g->spacing[k] = spacing[k];
        g->origin[k] = origin[k];
    }
    const size_t n_cells = (size_t) (counts[0] - 1) * (counts[1] - 1) * (counts[2] - 1);
    const size_t cell_bytes = precision == GFB_PRECISION_MIXED ? 32 : 64;
    // AUTO: the packed-cell copy whenever it is affordable (one 128-byte line per stencil is what HBM and L2 move;
    // measured fastest or within 7 % of fastest on every named configuration), else the 1.14x rows copy.
    if (layout == GFB_LAYOUT_AUTO)
        layout = n_cells * cell_bytes <= dev->prop.totalGlobalMem / 16 ? GFB_LAYOUT_CELLS : GFB_LAYOUT_ROWS;
    g->layout = layout;
    g->row_chunks = 0;
    size_t n_units = n_cells;      // threads' worth of work for the repack kernel
    if (layout == GFB_LAYOUT_CELLS) {
        g->bytes = n_cells * cell_bytes;
    } else if (layout == GFB_LAYOUT_ROWS) {
        const int w = precision == GFB_PRECISION_MIXED ? 8 : 4;            // values per 32-byte chunk
        g->row_chunks = (counts[2] - 2) / (w - 1) + 1;                     // covers every pair (iz, iz+1), iz <= nz-2
        n_units = (size_t) counts[0] * counts[1] * g->row_chunks;
        g->bytes = n_units * 32;
    } else if (layout == GFB_LAYOUT_PAIRS) {
        g->row_chunks = (counts[2] - 2) / 3 + 1;
        n_units = (size_t) counts[0] * (counts[1] - 1) * g->row_chunks;
        g->bytes = n_units * 32;
    } else if (layout == GFB_LAYOUT_BSPLINE_POINTS) {   // the points themselves (clamped indexing needs no guard)
        n_units = n_points;
        g->bytes = n_units * (precision == GFB_PRECISION_MIXED ? sizeof(float) : sizeof(double));
    } else if (layout == GFB_LAYOUT_POINTS) {   // the points themselves + one zero x-slab (tricubic_interpolate's flat-index reads)
        n_units = n_points + (size_t) counts[1] * counts[2];
        g->bytes = n_units * (precision == GFB_PRECISION_MIXED ? sizeof(float) : sizeof(double));
    } else {   // BSPLINE / HERMITE: records (a < nx+1, iy < ny-1, iz < nz-1) of 2 planes x 4 rows x 4 values, one thread per row
        g->row_chunks = counts[2] - 1;
        n_units = (size_t) (counts[0] + 1) * (counts[1] - 1) * (counts[2] - 1) * 8;
        g->bytes = n_units * 4 * (precision == GFB_PRECISION_MIXED ? sizeof(float) : sizeof(double));
    }
    g->cells = nullptr;
    g->released_bytes = 0;

    // Say what does not fit before cudaMalloc says "out of memory": the B-spline record layout is 32x the raw grid
    // (512^3: 17 GB MIXED, 34 GB DOUBLE), packed cells 8x.
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t need = g->bytes + (vals_on_device ? 0 : n_points * sizeof(double));
            if (need > free_b) {
                const int lay = layout;
                delete g;
                return fail(GFB_ERR_NOMEM, "gfb_grid_create: a %dx%dx%d grid in layout %d (%s) needs %.2f GB of device memory, %.2f GB are free%s",
                            counts[0], counts[1], counts[2], lay, lay == GFB_LAYOUT_BSPLINE ? "B-spline records, 32x the raw grid" : lay == GFB_LAYOUT_HERMITE ? "tricubic Hermite records, 32x the raw grid" : "see gfb_layout",
                            need / 1e9, free_b / 1e9, lay == GFB_LAYOUT_CELLS ? "; GFB_LAYOUT_ROWS needs 1.14x the raw grid" : "");
            }
        } else {
            cudaGetLastError();
        }
    }
    const double* d_vals = vals;
    void* d_tmp = nullptr;
    cudaError_t err = cudaMalloc(&g->cells, g->bytes);
    if (err == cudaSuccess && !vals_on_device) {
        err = cudaMalloc(&d_tmp, n_points * sizeof(double));
        if (err == cudaSuccess) err = cudaMemcpyAsync(d_tmp, vals, n_points * sizeof(double), cudaMemcpyHostToDevice, dev->stream);
        d_vals = static_cast<const double*>(d_tmp);
    }
    if (err == cudaSuccess) {
        const int blocks = (int) std::min<size_t>((n_units + 255) / 256, (size_t) dev->prop.multiProcessorCount * 32);
        const bool mixed = precision == GFB_PRECISION_MIXED;
        float* cf = static_cast<float*>(g->cells);
        double* cd = static_cast<double*>(g->cells);
        if (layout == GFB_LAYOUT_CELLS) {
            if (mixed) gf_repack_kernel<float><<<blocks, 256, 0, dev->stream>>>(d_vals, cf, counts[0], counts[1], counts[2]);
            else gf_repack_kernel<double><<<blocks, 256, 0, dev->stream>>>(d_vals, cd, counts[0], counts[1], counts[2]);
        } else if (layout == GFB_LAYOUT_ROWS) {
            if (mixed) gf_repack_rows_kernel<float><<<blocks, 256, 0, dev->stream>>>(d_vals, cf, counts[0], counts[1], counts[2], g->row_chunks);
            else gf_repack_rows_kernel<double><<<blocks, 256, 0, dev->stream>>>(d_vals, cd, counts[0], counts[1], counts[2], g->row_chunks);
        } else if (layout == GFB_LAYOUT_PAIRS) {
            gf_repack_pairs_kernel<<<blocks, 256, 0, dev->stream>>>(d_vals, cf, counts[0], counts[1], counts[2], g->row_chunks);
        } else if (layout == GFB_LAYOUT_POINTS || layout == GFB_LAYOUT_BSPLINE_POINTS) {
            const size_t guard = layout == GFB_LAYOUT_POINTS ? (size_t) counts[1] * counts[2] : 0;
            if (mixed) gf_repack_points_kernel<float><<<blocks, 256, 0, dev->stream>>>(d_vals, cf, n_points, guard);
            else gf_repack_points_kernel<double><<<blocks, 256, 0, dev->stream>>>(d_vals, cd, n_points, guard);
        } else {
            if (layout == GFB_LAYOUT_HERMITE && mixed) gf_repack_bspline_kernel<float, true><<<blocks, 256, 0, dev->stream>>>(d_vals, cf, counts[0], counts[1], counts[2]);
            else if (layout == GFB_LAYOUT_HERMITE) gf_repack_bspline_kernel<double, true><<<blocks, 256, 0, dev->stream>>>(d_vals, cd, counts[0], counts[1], counts[2]);
            else if (mixed) gf_repack_bspline_kernel<float, false><<<blocks, 256, 0, dev->stream>>>(d_vals, cf, counts[0], counts[1], counts[2]);
            else gf_repack_bspline_kernel<double, false><<<blocks, 256, 0, dev->stream>>>(d_vals, cd, counts[0], counts[1], counts[2]);
        }
        g_launches++;
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(dev->stream);
    if (d_tmp) cudaFree(d_tmp);
    if (err != cudaSuccess) {
        if (g->cells) cudaFree(g->cells);
        delete g;
        return fail(GFB_ERR_CUDA, "gfb_grid_create: %s", cudaGetErrorString(err));
    }
    *out = g;
    return GFB_OK;
}

extern "C" {

int gfb_grid_create(gfb_device* dev, const int counts[3], const double spacing[3], const double origin[3],
                    const double* vals, size_t n_vals, int precision, int layout, gfb_grid** out) {
    return grid_create_common(dev, counts, spacing, origin, vals, false, n_vals, precision, layout, out);
}

int gfb_grid_create_from_device(gfb_device* dev, const int counts[3], const double spacing[3], const double origin[3],
                                const double* d_vals, size_t n_vals, int precision, int layout, gfb_grid** out) {
    return grid_create_common(dev, counts, spacing, origin, d_vals, true, n_vals, precision, layout, out);
}

// ---- V3 grid files -------------------------------------------------------------------------------------------
int gfb_gridfile_read_header(const char* path, gfb_gridfile_header* header) {
    if (!path || !header) return fail(GFB_ERR_INVALID, "gfb_gridfile_read_header: NULL argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(GFB_ERR_INVALID, "GridForce: Cannot open file '%s'", path);
    const std::string err = gridfile_read_header(f, header);
    fclose(f);
    if (!err.empty()) return fail(GFB_ERR_INVALID, "GridForce: %s (%s)", err.c_str(), path);
    return GFB_OK;
}

int gfb_gridfile_read_values(const char* path, double* vals, size_t n_vals) {
    if (!path || !vals) return fail(GFB_ERR_INVALID, "gfb_gridfile_read_values: NULL argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(GFB_ERR_INVALID, "GridForce: Cannot open file '%s'", path);
    gfb_gridfile_header h;
    std::string err = gridfile_read_header(f, &h);
    const size_t n = (size_t) h.counts[0] * h.counts[1] * h.counts[2];
    if (err.empty() && n != n_vals) err = "buffer holds " + std::to_string(n_vals) + " values, file has " + std::to_string(n);
    if (err.empty() && fseek(f, (long) h.data_offset, SEEK_SET) != 0) err = "cannot seek to the data offset";
    if (err.empty() && fread(vals, sizeof(double), n, f) != n) err = "file ends before the last grid value";
    fclose(f);
    if (!err.empty()) return fail(GFB_ERR_INVALID, "GridForce: %s (%s)", err.c_str(), path);
    return GFB_OK;
}

int gfb_gridfile_write(const char* path, const gfb_gridfile_header* header, const double* vals, size_t n_vals, int with_trailer) {
    if (!path || !header || !vals) return fail(GFB_ERR_INVALID, "gfb_gridfile_write: NULL argument");
    const size_t n = (size_t) header->counts[0] * header->counts[1] * header->counts[2];
    if (n != n_vals) return fail(GFB_ERR_INVALID, "GridForce: Number of grid values doesn't match dimensions");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(GFB_ERR_INVALID, "GridForce: Cannot create file '%s'", path);
    PackedHeader h;
    gridfile_fill_header(*header, h);
    bool ok = fwrite(&h, 1, sizeof h, f) == sizeof h && fwrite(vals, sizeof(double), n, f) == n;
    if (ok && with_trailer) {   // GridData::saveToFile trailer (GridData.cpp:250-256)
        const int32_t n_scaling = 0;
        ok = fwrite(&n_scaling, sizeof n_scaling, 1, f) == 1 && fwrite(header->origin, sizeof(double), 3, f) == 3;
    }
    ok = fclose(f) == 0 && ok;
    if (!ok) return fail(GFB_ERR_INVALID, "GridForce: write to '%s' failed", path);
    return GFB_OK;
}

int gfb_grid_create_from_file(gfb_device* dev, const char* path, int precision, int layout, gfb_grid** out,
                              gfb_gridfile_header* header_out) {
    if (!dev || !path || !out) return fail(GFB_ERR_INVALID, "gfb_grid_create_from_file: NULL argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(GFB_ERR_INVALID, "GridForce: Cannot open file '%s'", path);
    gfb_gridfile_header h;
    std::string err = gridfile_read_header(f, &h);
    if (err.empty() && fseek(f, (long) h.data_offset, SEEK_SET) != 0) err = "cannot seek to the data offset";
    if (!err.empty()) {
        fclose(f);
        return fail(GFB_ERR_INVALID, "GridForce: %s (%s)", err.c_str(), path);
    }
    if (cudaSetDevice(dev->ordinal) != cudaSuccess) {
        fclose(f);
        return fail(GFB_ERR_CUDA, "gfb_grid_create_from_file: cudaSetDevice failed");
    }
    const size_t n = (size_t) h.counts[0] * h.counts[1] * h.counts[2];
    // disk -> two pinned 32 MB buffers (alternating) -> device doubles; then the normal on-device repack
    const size_t piece = (size_t) 4 << 20;   // doubles per piece (32 MB)
    double* d_vals = nullptr;
    double* stage[2] = {nullptr, nullptr};
    cudaEvent_t used[2] = {nullptr, nullptr};
    cudaError_t ce = cudaMalloc((void**) &d_vals, n * sizeof(double));
    for (int b = 0; b < 2 && ce == cudaSuccess; b++) {
        ce = cudaHostAlloc((void**) &stage[b], std::min(piece, n) * sizeof(double), cudaHostAllocDefault);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&used[b], cudaEventDisableTiming);
    }
    size_t done = 0;
    int b = 0;
    while (ce == cudaSuccess && err.empty() && done < n) {
        const size_t cnt = std::min(piece, n - done);
        ce = cudaEventSynchronize(used[b]);      // the copy that last read this buffer has finished
        if (ce != cudaSuccess) break;
        if (fread(stage[b], sizeof(double), cnt, f) != cnt) {
            err = "file ends before the last grid value";
            break;
        }
        ce = cudaMemcpyAsync(d_vals + done, stage[b], cnt * sizeof(double), cudaMemcpyHostToDevice, dev->stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(used[b], dev->stream);
        done += cnt;
        b ^= 1;
    }
    fclose(f);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(dev->stream);
    int rc = GFB_OK;
    if (ce != cudaSuccess) rc = fail(GFB_ERR_CUDA, "gfb_grid_create_from_file: %s", cudaGetErrorString(ce));
    else if (!err.empty()) rc = fail(GFB_ERR_INVALID, "GridForce: %s (%s)", err.c_str(), path);
    else rc = grid_create_common(dev, h.counts, h.spacing, h.origin, d_vals, true, n, precision, layout, out);
    for (int i = 0; i < 2; i++) {
        if (stage[i]) cudaFreeHost(stage[i]);
        if (used[i]) cudaEventDestroy(used[i]);
    }
    if (d_vals) cudaFree(d_vals);
    if (rc == GFB_OK && header_out) *header_out = h;
    return rc;
}

int gfb_grid_generate(gfb_device* dev, const int counts[3], const double spacing[3], const double origin[3], int grid_type,
                      int n_atoms, const double* pos, const double* charges, const double* sigmas, const double* epsilons,
                      double grid_cap, double* vals_out, int precision, int layout, gfb_grid** grid_out) {
    if (!dev || !counts || !spacing || !origin || (n_atoms > 0 && !pos)) return fail(GFB_ERR_INVALID, "gfb_grid_generate: NULL argument");
    if (grid_out) *grid_out = nullptr;
    if (grid_type < 1 || grid_type > 3)
        return fail(GFB_ERR_INVALID, "GridForce: Invalid grid type code %d. Must be 1 (charge), 2 (ljr) or 3 (lja)", grid_type);
    if ((grid_type == 1 && !charges) || (grid_type != 1 && (!sigmas || !epsilons)))
        return fail(GFB_ERR_INVALID, "gfb_grid_generate: the parameter array this grid type needs is NULL");
    if (n_atoms < 0 || !(grid_cap > 0.0)) return fail(GFB_ERR_INVALID, "gfb_grid_generate: n_atoms=%d grid_cap=%g", n_atoms, grid_cap);
    for (int k = 0; k < 3; k++)
        if (counts[k] < 1) return fail(GFB_ERR_INVALID, "gfb_grid_generate: counts[%d]=%d", k, counts[k]);
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    const size_t n_points = (size_t) counts[0] * counts[1] * counts[2];
    // fold the per-atom parameters into one coefficient, associating exactly as the reference's expressions do
    std::vector<double> packed((size_t) std::max(n_atoms, 1) * 4, 0.0);
    for (int a = 0; a < n_atoms; a++) {
        packed[4 * (size_t) a] = pos[3 * a];
        packed[4 * (size_t) a + 1] = pos[3 * a + 1];
        packed[4 * (size_t) a + 2] = pos[3 * a + 2];
        double c;
        if (grid_type == 1) c = 138.935456 * charges[a];                                     // :527
        else if (grid_type == 2) c = std::sqrt(epsilons[a]) * std::pow(2.0 * sigmas[a], 6.0);   // :530-531
        else c = -2.0 * std::sqrt(epsilons[a]) * std::pow(2.0 * sigmas[a], 3.0);                // :534-535
        packed[4 * (size_t) a + 3] = c;
    }
    double4* d_atoms = nullptr;
    double* d_vals = nullptr;
    cudaError_t ce = cudaMalloc((void**) &d_atoms, packed.size() * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc((void**) &d_vals, n_points * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_atoms, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice, dev->stream);
    if (ce == cudaSuccess) {
        const unsigned blocks = (unsigned) ((n_points + 255) / 256);
        const double ox = origin[0], oy = origin[1], oz = origin[2], sx = spacing[0], sy = spacing[1], sz = spacing[2];
        if (grid_type == 1)
            gf_generate_grid_kernel<1><<<blocks, 256, 0, dev->stream>>>(d_atoms, n_atoms, counts[0], counts[1], counts[2], ox, oy, oz, sx, sy, sz, grid_cap, d_vals);
        else if (grid_type == 2)
            gf_generate_grid_kernel<12><<<blocks, 256, 0, dev->stream>>>(d_atoms, n_atoms, counts[0], counts[1], counts[2], ox, oy, oz, sx, sy, sz, grid_cap, d_vals);
        else
            gf_generate_grid_kernel<6><<<blocks, 256, 0, dev->stream>>>(d_atoms, n_atoms, counts[0], counts[1], counts[2], ox, oy, oz, sx, sy, sz, grid_cap, d_vals);
        g_launches++;
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess && vals_out) ce = cudaMemcpyAsync(vals_out, d_vals, n_points * sizeof(double), cudaMemcpyDeviceToHost, dev->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(dev->stream);
    int rc = GFB_OK;
    if (ce != cudaSuccess) rc = fail(GFB_ERR_CUDA, "gfb_grid_generate: %s", cudaGetErrorString(ce));
    else if (grid_out) rc = grid_create_common(dev, counts, spacing, origin, d_vals, true, n_points, precision, layout, grid_out);
    if (d_atoms) cudaFree(d_atoms);
    if (d_vals) cudaFree(d_vals);
    return rc;
}

int gfb_inv_power_transform(gfb_device* dev, double* vals, size_t n_vals, double inv_power, int vals_on_device) {
    if (!dev || (!vals && n_vals)) return fail(GFB_ERR_INVALID, "gfb_inv_power_transform: NULL argument");
    if (inv_power == 0.0) return fail(GFB_ERR_INVALID, "GridForce: inv_power must be non-zero");
    if (n_vals == 0) return GFB_OK;
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    double* d = vals;
    if (!vals_on_device) {
        CUDA_TRY(cudaMalloc((void**) &d, n_vals * sizeof(double)));
        cudaError_t e = cudaMemcpyAsync(d, vals, n_vals * sizeof(double), cudaMemcpyHostToDevice, dev->stream);
        if (e != cudaSuccess) {
            cudaFree(d);
            return fail(GFB_ERR_CUDA, "gfb_inv_power_transform: H2D: %s", cudaGetErrorString(e));
        }
    }
    const int blocks = (int) std::min<size_t>((n_vals + 255) / 256, (size_t) dev->prop.multiProcessorCount * 32);
    gf_inv_power_transform_kernel<<<blocks, 256, 0, dev->stream>>>(d, n_vals, 1.0 / inv_power);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !vals_on_device) e = cudaMemcpyAsync(vals, d, n_vals * sizeof(double), cudaMemcpyDeviceToHost, dev->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dev->stream);
    if (!vals_on_device) cudaFree(d);
    if (e != cudaSuccess) return fail(GFB_ERR_CUDA, "gfb_inv_power_transform: %s", cudaGetErrorString(e));
    return GFB_OK;
}

int gfb_grid_destroy(gfb_grid* grid) {
    if (!grid) return GFB_OK;
    cudaSetDevice(grid->dev->ordinal);
    cudaFree(grid->cells);
    delete grid;
    return GFB_OK;
}

int gfb_grid_release_cells(gfb_grid* grid) {
    if (!grid) return fail(GFB_ERR_INVALID, "gfb_grid_release_cells: NULL grid");
    if (!grid->cells) return GFB_OK;
    CUDA_TRY(cudaSetDevice(grid->dev->ordinal));
    CUDA_TRY(cudaStreamSynchronize(grid->dev->stream));
    CUDA_TRY(cudaFree(grid->cells));
    grid->cells = nullptr;
    grid->released_bytes = grid->bytes;
    grid->bytes = 0;
    return GFB_OK;
}

size_t gfb_grid_device_bytes(const gfb_grid* grid) { return grid ? grid->bytes : 0; }
int gfb_grid_layout(const gfb_grid* grid) { return grid ? grid->layout : GFB_LAYOUT_AUTO; }

}  // extern "C"

namespace gfb {
// Weaves the packed cells of up to four grids into one record per cell (see gf_interleave_cells_kernel). Stream-ordered
// on dev->stream and synchronised. bytes_per_slot: 32 (MIXED) or 64 (DOUBLE).
cudaError_t interleave_cells(gfb_device* dev, const void* const src[4], void* dst, size_t n_cells, int slots, int bytes_per_slot) {
    const int parts = bytes_per_slot / 16;
    const int blocks = (int) std::min<size_t>((n_cells * slots * parts + 255) / 256, (size_t) dev->prop.multiProcessorCount * 32);
    gf_interleave_cells_kernel<<<blocks, 256, 0, dev->stream>>>(static_cast<const float4*>(src[0]), static_cast<const float4*>(src[1]),
                                                                static_cast<const float4*>(src[2]), static_cast<const float4*>(src[3]),
                                                                static_cast<float4*>(dst), n_cells, slots, parts);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(dev->stream);
    return e;
}
}  // namespace gfb
