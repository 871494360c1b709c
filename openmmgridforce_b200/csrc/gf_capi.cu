// C ABI of libgridforce_b200.so (include/gridforce_b200.h): handles, uploads, launches.
// Host logic only — every number is produced by the kernels in gf_kernels.cuh. No CPU fallback.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "gf_gridfile.h"
#include "gf_kernels.cuh"
#include "gf_eval_lines.cuh"
#include "gf_eval_bspline.cuh"
#include "gridforce_b200.h"

using namespace gfb;

// ---------------------------------------------------------------------------------------------
// Errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_error;
static std::atomic<unsigned long long> g_launches(0);

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t err__ = (expr);                                                                         \
        if (err__ != cudaSuccess)                                                                           \
            return fail(GFB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Handles
// ---------------------------------------------------------------------------------------------
struct DeviceBuffer {
    void* ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return GFB_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4;
        CUDA_TRY(cudaMalloc(&ptr, want));
        cap = want;
        return GFB_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

struct PinnedBuffer {
    void* ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return GFB_OK;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4;
        CUDA_TRY(cudaHostAlloc(&ptr, want, cudaHostAllocDefault));
        cap = want;
        return GFB_OK;
    }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

struct gfb_device {
    int ordinal;
    cudaStream_t stream;
    cudaStream_t copy_stream;   // D2H of chunk i overlaps the kernel of chunk i+1 ...
    cudaStream_t h2d_stream;    // ... and the H2D of chunk i+2: three streams chained by events (execute_host)
    std::vector<cudaEvent_t> events;   // pool for the chunk pipeline (2 per chunk), created once
    std::mutex host_mutex;             // host-path calls share the three streams and the event pool: one at a time per GPU
    cudaDeviceProp prop;
};

struct gfb_grid {
    gfb_device* dev;
    int counts[3];
    double spacing[3], origin[3];
    int precision;
    int layout;        // resolved gfb_layout (never AUTO)
    int row_chunks;    // ROWS / PAIRS: 32-byte units per row
    void* cells;
    size_t bytes;
};

struct gfb_kernel {
    gfb_device* dev;
    int n_grids, n_atoms, precision;
    bool same_geom;
    gfb_grid* grids[GFB_MAX_GRIDS];
    double inv_power[GFB_MAX_GRIDS], oob_k[GFB_MAX_GRIDS];
    void* d_scaling;      // [n_grids][n_atoms] float|double
    int* d_particles;     // [n_atoms] or null
    int max_particle;     // largest particle index referenced (+1 = minimum n_particles)
    int* d_slots;         // [n_atoms] energy slot per atom (particle groups) or null
    int n_slots;          // energy slots per replica
    float* d_interleaved; // MIXED + CELLS + shared geometry + 2..4 grids: one record per cell holding every grid's corners
    int il_slots;         // grids per record incl. padding (2 or 4); 0 = not interleaved
    // host-path scratch
    DeviceBuffer d_pos, d_forces, d_energy, d_cls, d_sort;
    PinnedBuffer h_stage, h_energy;
    // one-ligand-per-step path (execute_host_small): host-mapped staging the kernel reads and writes over PCIe
    PinnedBuffer h_small;
    bool launch_overlap;     // gfb_kernel_set_launch_overlap: device-path launches may overlap the previous launch's tail
    bool unique_particles;   // no particle index occurs twice: a plain store per evaluated particle is the whole force
};

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// ---------------------------------------------------------------------------------------------
// Library / device
// ---------------------------------------------------------------------------------------------
extern "C" {

int gfb_version(void) { return GFB_VERSION; }
const char* gfb_last_error(void) { return g_error.c_str(); }
unsigned long long gfb_launch_count(void) { return g_launches.load(); }

int gfb_device_count(int* count) {
    if (!count) return fail(GFB_ERR_INVALID, "gfb_device_count: count is NULL");
    *count = 0;
    CUDA_TRY(cudaGetDeviceCount(count));
    return GFB_OK;
}

int gfb_device_open(int ordinal, gfb_device** out) {
    if (!out) return fail(GFB_ERR_INVALID, "gfb_device_open: out is NULL");
    *out = nullptr;
    int n = 0;
    CUDA_TRY(cudaGetDeviceCount(&n));
    if (ordinal < 0 || ordinal >= n) return fail(GFB_ERR_CUDA, "gfb_device_open: no CUDA device %d (found %d)", ordinal, n);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, ordinal));
    if (prop.major != 10)
        return fail(GFB_ERR_CUDA, "gfb_device_open: device %d (%s, sm_%d%d) is not Blackwell sm_100; this library has no other code path",
                    ordinal, prop.name, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(ordinal));
    if (const char* fg = getenv("GFB_L2_FETCH_GRANULARITY")) {   // tuning probe: 32 | 64 | 128 bytes
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t) atoi(fg));
        if (e != cudaSuccess) cudaGetLastError();
    }
    gfb_device* d = new (std::nothrow) gfb_device();
    if (!d) return fail(GFB_ERR_NOMEM, "gfb_device_open: out of host memory");
    d->ordinal = ordinal;
    d->prop = prop;
    CUDA_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&d->h2d_stream, cudaStreamNonBlocking));
    *out = d;
    return GFB_OK;
}

int gfb_device_close(gfb_device* dev) {
    if (!dev) return GFB_OK;
    cudaSetDevice(dev->ordinal);
    cudaStreamDestroy(dev->stream);
    cudaStreamDestroy(dev->copy_stream);
    cudaStreamDestroy(dev->h2d_stream);
    for (size_t i = 0; i < dev->events.size(); i++) cudaEventDestroy(dev->events[i]);
    delete dev;
    return GFB_OK;
}

int gfb_device_get_props(gfb_device* dev, gfb_device_props* props) {
    if (!dev || !props) return fail(GFB_ERR_INVALID, "gfb_device_get_props: NULL argument");
    memset(props, 0, sizeof *props);
    strncpy(props->name, dev->prop.name, sizeof props->name - 1);
    props->cc_major = dev->prop.major;
    props->cc_minor = dev->prop.minor;
    props->sm_count = dev->prop.multiProcessorCount;
    props->l2_bytes = dev->prop.l2CacheSize;
    props->total_mem_bytes = dev->prop.totalGlobalMem;
    return GFB_OK;
}

int gfb_device_synchronize(gfb_device* dev) {
    if (!dev) return fail(GFB_ERR_INVALID, "gfb_device_synchronize: NULL device");
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    CUDA_TRY(cudaStreamSynchronize(dev->h2d_stream));
    CUDA_TRY(cudaStreamSynchronize(dev->stream));
    CUDA_TRY(cudaStreamSynchronize(dev->copy_stream));
    return GFB_OK;
}

// ---------------------------------------------------------------------------------------------
// Grids
// ---------------------------------------------------------------------------------------------
static int grid_create_common(gfb_device* dev, const int counts[3], const double spacing[3], const double origin[3],
                              const double* vals, bool vals_on_device, size_t n_vals, int precision, int layout,
                              gfb_grid** out) {
    if (!dev || !counts || !spacing || !origin || !vals || !out) return fail(GFB_ERR_INVALID, "gfb_grid_create: NULL argument");
    *out = nullptr;
    if (precision != GFB_PRECISION_MIXED && precision != GFB_PRECISION_DOUBLE)
        return fail(GFB_ERR_INVALID, "gfb_grid_create: unknown precision %d", precision);
    if (layout < GFB_LAYOUT_AUTO || layout > GFB_LAYOUT_BSPLINE) return fail(GFB_ERR_INVALID, "gfb_grid_create: unknown layout %d", layout);
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
    } else {   // BSPLINE: records (a < nx+1, iy < ny-1, iz < nz-1) of 2 planes x 4 rows x 4 values, one thread per row
        g->row_chunks = counts[2] - 1;
        n_units = (size_t) (counts[0] + 1) * (counts[1] - 1) * (counts[2] - 1) * 8;
        g->bytes = n_units * 4 * (precision == GFB_PRECISION_MIXED ? sizeof(float) : sizeof(double));
    }
    g->cells = nullptr;

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
        } else {
            if (mixed) gf_repack_bspline_kernel<float><<<blocks, 256, 0, dev->stream>>>(d_vals, cf, counts[0], counts[1], counts[2]);
            else gf_repack_bspline_kernel<double><<<blocks, 256, 0, dev->stream>>>(d_vals, cd, counts[0], counts[1], counts[2]);
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

size_t gfb_grid_device_bytes(const gfb_grid* grid) { return grid ? grid->bytes : 0; }
int gfb_grid_layout(const gfb_grid* grid) { return grid ? grid->layout : GFB_LAYOUT_AUTO; }

// ---------------------------------------------------------------------------------------------
// Kernel state (= CalcGridForceKernel after initialize)
// ---------------------------------------------------------------------------------------------
static int upload_scaling(gfb_kernel* k, const double* scaling) {
    // Scaling factors stay FP64 on the device in both precisions: the energy term s*V is formed in FP64, and the
    // reference's branch on scale != 0.0 (:706) is then reproduced exactly.
    const size_t n = (size_t) k->n_grids * k->n_atoms;
    CUDA_TRY(cudaMemcpyAsync(k->d_scaling, scaling, n * sizeof(double), cudaMemcpyHostToDevice, k->dev->stream));
    CUDA_TRY(cudaStreamSynchronize(k->dev->stream));
    return GFB_OK;
}

int gfb_kernel_create(gfb_device* dev, int n_grids, gfb_grid* const* grids, int n_atoms, const double* scaling,
                      const int* particles, const double* inv_power, const double* oob_k, gfb_kernel** out) {
    if (!dev || !grids || !out || !oob_k) return fail(GFB_ERR_INVALID, "gfb_kernel_create: NULL argument");
    *out = nullptr;
    if (n_grids < 1 || n_grids > GFB_MAX_GRIDS)
        return fail(GFB_ERR_INVALID, "gfb_kernel_create: n_grids=%d, supported 1..%d", n_grids, GFB_MAX_GRIDS);
    if (n_atoms < 0) return fail(GFB_ERR_INVALID, "gfb_kernel_create: n_atoms=%d", n_atoms);
    if (n_atoms > 0 && !scaling) return fail(GFB_ERR_INVALID, "gfb_kernel_create: scaling is NULL");
    for (int g = 0; g < n_grids; g++) {
        if (!grids[g]) return fail(GFB_ERR_INVALID, "gfb_kernel_create: grids[%d] is NULL", g);
        if (grids[g]->dev != dev) return fail(GFB_ERR_INVALID, "gfb_kernel_create: grids[%d] lives on another device", g);
        if (grids[g]->precision != grids[0]->precision)
            return fail(GFB_ERR_INVALID, "gfb_kernel_create: grids[%d] precision differs from grids[0]", g);
        if (grids[g]->layout != grids[0]->layout)
            return fail(GFB_ERR_INVALID, "gfb_kernel_create: grids[%d] layout differs from grids[0]", g);
        if (inv_power && inv_power[g] < 0.0) return fail(GFB_ERR_INVALID, "gfb_kernel_create: inv_power[%d] < 0", g);
    }
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    gfb_kernel* k = new (std::nothrow) gfb_kernel();
    if (!k) return fail(GFB_ERR_NOMEM, "gfb_kernel_create: out of host memory");
    k->dev = dev;
    k->n_grids = n_grids;
    k->n_atoms = n_atoms;
    k->precision = grids[0]->precision;
    k->same_geom = true;
    k->d_scaling = nullptr;
    k->d_particles = nullptr;
    k->d_interleaved = nullptr;
    k->il_slots = 0;
    k->d_slots = nullptr;
    k->n_slots = 1;
    k->max_particle = n_atoms - 1;
    k->unique_particles = true;
    k->launch_overlap = false;
    for (int g = 0; g < n_grids; g++) {
        k->grids[g] = grids[g];
        k->inv_power[g] = inv_power ? inv_power[g] : 0.0;
        k->oob_k[g] = oob_k[g];
        for (int a = 0; a < 3; a++)
            if (grids[g]->counts[a] != grids[0]->counts[a] || grids[g]->spacing[a] != grids[0]->spacing[a] ||
                grids[g]->origin[a] != grids[0]->origin[a])
                k->same_geom = false;
    }
    cudaError_t err = cudaMalloc(&k->d_scaling, std::max<size_t>((size_t) n_grids * n_atoms * sizeof(double), 16));
    if (err == cudaSuccess && particles && n_atoms > 0) {
        k->max_particle = -1;
        for (int i = 0; i < n_atoms; i++) {
            if (particles[i] < 0) {
                cudaFree(k->d_scaling);
                delete k;
                return fail(GFB_ERR_INVALID, "gfb_kernel_create: particles[%d]=%d is negative", i, particles[i]);
            }
            k->max_particle = std::max(k->max_particle, particles[i]);
        }
        std::vector<int> sorted(particles, particles + n_atoms);
        std::sort(sorted.begin(), sorted.end());
        k->unique_particles = std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end();
        err = cudaMalloc((void**) &k->d_particles, (size_t) n_atoms * sizeof(int));
        if (err == cudaSuccess)
            err = cudaMemcpy(k->d_particles, particles, (size_t) n_atoms * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (err != cudaSuccess) {
        if (k->d_scaling) cudaFree(k->d_scaling);
        if (k->d_particles) cudaFree(k->d_particles);
        delete k;
        return fail(GFB_ERR_CUDA, "gfb_kernel_create: %s", cudaGetErrorString(err));
    }
    if (n_atoms > 0) {
        int rc = upload_scaling(k, scaling);
        if (rc != GFB_OK) {
            gfb_kernel_destroy(k);
            return rc;
        }
    }
    // 2-4 MIXED packed-cell grids of one geometry: weave them into one 128-byte record per cell (4 slots of 32 bytes),
    // so that everything an atom needs is ONE line of HBM/L2 (gf_eval_lines.cuh reads it with quad-coalesced loads).
    // Costs a second copy of the grids (4 x 32 bytes per cell); GFB_LINES=0 keeps the per-grid arrays only.
    const char* il = getenv("GFB_LINES");
    const bool want_lines = !(il && il[0] == '0');
    const size_t n_cells0 = grids[0]->bytes / 32;
    if (want_lines && k->same_geom && n_grids >= 2 && n_grids <= 4 && k->precision == GFB_PRECISION_MIXED &&
        grids[0]->layout == GFB_LAYOUT_CELLS && n_cells0 < 0xffffffffull &&
        n_cells0 * 128 <= dev->prop.totalGlobalMem / 8) {
        const size_t n_cells = n_cells0;
        const int slots = 4;
        cudaError_t e = cudaMalloc((void**) &k->d_interleaved, n_cells * slots * 32);
        if (e == cudaSuccess) {
            const float4* src[4] = {nullptr, nullptr, nullptr, nullptr};
            for (int g = 0; g < n_grids; g++) src[g] = static_cast<const float4*>(grids[g]->cells);
            const int blocks = (int) std::min<size_t>((n_cells * slots * 2 + 255) / 256, (size_t) dev->prop.multiProcessorCount * 32);
            gf_interleave_cells_kernel<<<blocks, 256, 0, dev->stream>>>(src[0], src[1], src[2], src[3],
                                                                        reinterpret_cast<float4*>(k->d_interleaved), n_cells, slots);
            g_launches++;
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(dev->stream);
        }
        if (e != cudaSuccess) {      // not fatal: the general kernel reads the per-grid arrays
            cudaGetLastError();
            if (k->d_interleaved) cudaFree(k->d_interleaved);
            k->d_interleaved = nullptr;
        } else {
            k->il_slots = slots;
        }
    }
    *out = k;
    return GFB_OK;
}

int gfb_kernel_destroy(gfb_kernel* k) {
    if (!k) return GFB_OK;
    cudaSetDevice(k->dev->ordinal);
    cudaStreamSynchronize(k->dev->stream);
    cudaStreamSynchronize(k->dev->copy_stream);
    if (k->d_scaling) cudaFree(k->d_scaling);
    if (k->d_particles) cudaFree(k->d_particles);
    if (k->d_interleaved) cudaFree(k->d_interleaved);
    if (k->d_slots) cudaFree(k->d_slots);
    k->d_pos.release();
    k->d_forces.release();
    k->d_energy.release();
    k->d_cls.release();
    k->d_sort.release();
    k->h_stage.release();
    k->h_energy.release();
    k->h_small.release();
    delete k;
    return GFB_OK;
}

int gfb_kernel_set_energy_slots(gfb_kernel* k, const int* slots, int n_slots) {
    if (!k) return fail(GFB_ERR_INVALID, "gfb_kernel_set_energy_slots: NULL kernel");
    CUDA_TRY(cudaSetDevice(k->dev->ordinal));
    if (!slots) {
        if (k->d_slots) cudaFree(k->d_slots);
        k->d_slots = nullptr;
        k->n_slots = 1;
        return GFB_OK;
    }
    if (n_slots < 1) return fail(GFB_ERR_INVALID, "gfb_kernel_set_energy_slots: n_slots=%d", n_slots);
    for (int i = 0; i < k->n_atoms; i++)
        if (slots[i] < 0 || slots[i] >= n_slots)
            return fail(GFB_ERR_INVALID, "gfb_kernel_set_energy_slots: slots[%d]=%d outside [0,%d)", i, slots[i], n_slots);
    if (!k->d_slots) CUDA_TRY(cudaMalloc((void**) &k->d_slots, std::max<size_t>((size_t) k->n_atoms * sizeof(int), 16)));
    CUDA_TRY(cudaMemcpy(k->d_slots, slots, (size_t) k->n_atoms * sizeof(int), cudaMemcpyHostToDevice));
    k->n_slots = n_slots;
    return GFB_OK;
}

int gfb_kernel_set_launch_overlap(gfb_kernel* k, int enable) {
    if (!k) return fail(GFB_ERR_INVALID, "gfb_kernel_set_launch_overlap: NULL kernel");
    k->launch_overlap = enable != 0;
    return GFB_OK;
}

int gfb_kernel_update_parameters(gfb_kernel* k, const double* scaling, const double* inv_power) {
    if (!k) return fail(GFB_ERR_INVALID, "gfb_kernel_update_parameters: NULL kernel");
    CUDA_TRY(cudaSetDevice(k->dev->ordinal));
    if (inv_power)
        for (int g = 0; g < k->n_grids; g++) {
            if (inv_power[g] < 0.0) return fail(GFB_ERR_INVALID, "gfb_kernel_update_parameters: inv_power[%d] < 0", g);
            k->inv_power[g] = inv_power[g];
        }
    if (scaling && k->n_atoms > 0) return upload_scaling(k, scaling);
    return GFB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Launch
// ---------------------------------------------------------------------------------------------
static void fill_grid_view(const gfb_kernel* k, int g, GridView& v) {
    const gfb_grid* gr = k->grids[g];
    v.cells = gr->cells;
    v.cell_stride = 8;
    v.pad_ = 0;
    v.scaling = static_cast<const double*>(k->d_scaling) + (size_t) g * k->n_atoms;
    for (int a = 0; a < 3; a++) {
        v.origin[a] = gr->origin[a];
        v.spacing[a] = gr->spacing[a];
        v.inv_spacing[a] = 1.0 / gr->spacing[a];
        v.hcorner[a] = gr->spacing[a] * (gr->counts[a] - 1);   // ReferenceGridForceKernels.cpp:654-656
        v.nc[a] = gr->counts[a] - 1;
    }
    v.row_chunks = gr->row_chunks;
    v.inv_power = k->inv_power[g];
    v.oob_k = k->oob_k[g];
}

template <typename S, int LAYOUT, int NG, bool SAME, int FMODE>
static void launch_eval4(const EvalParams& p, cudaStream_t stream) {
    const unsigned blocks = (unsigned) ((p.total + kBlock - 1) / kBlock);
    if (p.n_replicas == 1 && p.slots == nullptr)
        gf_eval_kernel<S, LAYOUT, NG, SAME, FMODE, true><<<blocks, kBlock, 0, stream>>>(p);
    else
        gf_eval_kernel<S, LAYOUT, NG, SAME, FMODE, false><<<blocks, kBlock, 0, stream>>>(p);
}

template <typename S, int LAYOUT, int NG, bool SAME>
static void launch_eval3(const EvalParams& p, int fmode, cudaStream_t stream) {
    switch (fmode) {
        case GFB_FORCE_FIXED_ADD: launch_eval4<S, LAYOUT, NG, SAME, GFB_FORCE_FIXED_ADD>(p, stream); break;
        case GFB_FORCE_F64_ADD: launch_eval4<S, LAYOUT, NG, SAME, GFB_FORCE_F64_ADD>(p, stream); break;
        default: launch_eval4<S, LAYOUT, NG, SAME, GFB_FORCE_F64_STORE>(p, stream); break;
    }
}

template <typename S, int LAYOUT>
static void launch_eval2(const EvalParams& p, bool same, int fmode, cudaStream_t stream) {
    if (p.n_grids == 1) launch_eval3<S, LAYOUT, 1, true>(p, fmode, stream);
    else if (p.n_grids == 3 && same) launch_eval3<S, LAYOUT, 3, true>(p, fmode, stream);
    else if (same) launch_eval3<S, LAYOUT, 0, true>(p, fmode, stream);
    else launch_eval3<S, LAYOUT, 0, false>(p, fmode, stream);
}

static void launch_eval1(const EvalParams& p, int precision, int layout, bool same, int fmode, cudaStream_t stream) {
    if (layout == GFB_LAYOUT_BSPLINE) {   // cubic B-spline: run-time grid count, each grid classified on its own geometry
        if (precision == GFB_PRECISION_DOUBLE) launch_eval3<double, GFB_LAYOUT_BSPLINE, 0, false>(p, fmode, stream);
        else launch_eval3<float, GFB_LAYOUT_BSPLINE, 0, false>(p, fmode, stream);
    } else if (precision == GFB_PRECISION_DOUBLE) {
        if (layout == GFB_LAYOUT_CELLS) launch_eval2<double, GFB_LAYOUT_CELLS>(p, same, fmode, stream);
        else launch_eval2<double, GFB_LAYOUT_ROWS>(p, same, fmode, stream);
    } else {
        if (layout == GFB_LAYOUT_CELLS) launch_eval2<float, GFB_LAYOUT_CELLS>(p, same, fmode, stream);
        else if (layout == GFB_LAYOUT_ROWS) launch_eval2<float, GFB_LAYOUT_ROWS>(p, same, fmode, stream);
        else launch_eval2<float, GFB_LAYOUT_PAIRS>(p, same, fmode, stream);
    }
}

// ---- gf_eval_lines_kernel dispatch (gf_eval_lines.cuh) ---------------------------------------------------------
// How the ADD force modes reach memory: 0 RED atomics alone, 1 RED + L2 prefetch of the force lines at kernel start.
// Measured (B200): the prefetch takes C3 (one grid: the force lines are a third of the traffic) from 45.5
// to 35.1 us and changes nothing for three grids. Default: prefetch for one grid. GFB_FORCE_PATH=0|1 overrides (probe).
static int force_path_default(int n_grids) {
    static const int v = [] {
        const char* e = getenv("GFB_FORCE_PATH");
        return e ? std::max(0, std::min(1, atoi(e))) : -1;
    }();
    return v >= 0 ? v : (n_grids == 1 ? kForcePrefetch : kForceRed);
}

template <int NG, int FMODE, int FPATH, bool SINGLE>
static void launch_lines4(const EvalParams& p, cudaStream_t stream) {
    constexpr int block = lines_block(NG);
    const unsigned blocks = (unsigned) ((p.total + block - 1) / block);
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
    if (p.grid_energies) cudaLaunchKernelEx(&cfg, gf_eval_lines_kernel<NG, FMODE, FPATH, SINGLE, true>, p);
    else cudaLaunchKernelEx(&cfg, gf_eval_lines_kernel<NG, FMODE, FPATH, SINGLE, false>, p);
}

template <int NG, int FMODE, int FPATH>
static void launch_lines3(const EvalParams& p, cudaStream_t stream) {
    if (p.n_replicas == 1 && p.slots == nullptr) launch_lines4<NG, FMODE, FPATH, true>(p, stream);
    else launch_lines4<NG, FMODE, FPATH, false>(p, stream);
}

template <int NG>
static void launch_lines2(const EvalParams& p, int fmode, int fpath, cudaStream_t stream) {
    if (fmode == GFB_FORCE_F64_STORE || !p.forces) {
        launch_lines3<NG, GFB_FORCE_F64_STORE, kForceRed>(p, stream);
    } else if (fmode == GFB_FORCE_FIXED_ADD) {
        if (fpath == kForcePrefetch) launch_lines3<NG, GFB_FORCE_FIXED_ADD, kForcePrefetch>(p, stream);
        else launch_lines3<NG, GFB_FORCE_FIXED_ADD, kForceRed>(p, stream);
    } else {
        if (fpath == kForcePrefetch) launch_lines3<NG, GFB_FORCE_F64_ADD, kForcePrefetch>(p, stream);
        else launch_lines3<NG, GFB_FORCE_F64_ADD, kForceRed>(p, stream);
    }
}

// The lines kernel serves MIXED packed cells of one geometry without inv-power and without an evaluation order.
static bool lines_eligible(const gfb_kernel* k, const EvalParams& p) {
    static const bool off = [] {
        const char* e = getenv("GFB_LINES");
        return e && e[0] == '0';
    }();
    if (off || k->precision != GFB_PRECISION_MIXED || k->grids[0]->layout != GFB_LAYOUT_CELLS || !k->same_geom) return false;
    if (p.order != nullptr || k->n_grids > 4) return false;
    if (k->n_grids > 1 && !k->il_slots) return false;
    if (k->grids[0]->bytes / 32 >= 0xffffffffull) return false;
    for (int g = 0; g < k->n_grids; g++)
        if (k->inv_power[g] > 0.0) return false;
    return true;
}

// gf_eval_bspline_kernel (gf_eval_bspline.cuh): MIXED B-spline tiles of one geometry, no evaluation order.
static bool bspline_tiles_eligible(const gfb_kernel* k, const EvalParams& p) {
    static const bool off = [] {
        const char* e = getenv("GFB_BSPLINE_TILES");   // 0: always the general kernel (A/B measurements)
        return e && e[0] == '0';
    }();
    if (off || k->precision != GFB_PRECISION_MIXED || k->grids[0]->layout != GFB_LAYOUT_BSPLINE || !k->same_geom) return false;
    if (p.order != nullptr) return false;
    return k->grids[0]->bytes / 128 < 0x7fffffffull;   // 32-bit record index
}

template <int FMODE>
static void launch_bspline2(const EvalParams& p, cudaStream_t stream) {
    const unsigned blocks = (unsigned) ((p.total + kBsBlock - 1) / kBsBlock);
    if (p.n_replicas == 1 && p.slots == nullptr) gf_eval_bspline_kernel<FMODE, true><<<blocks, kBsBlock, 0, stream>>>(p);
    else gf_eval_bspline_kernel<FMODE, false><<<blocks, kBsBlock, 0, stream>>>(p);
}

static void launch_bspline(const EvalParams& p, int fmode, cudaStream_t stream) {
    if (fmode == GFB_FORCE_FIXED_ADD) launch_bspline2<GFB_FORCE_FIXED_ADD>(p, stream);
    else if (fmode == GFB_FORCE_F64_ADD) launch_bspline2<GFB_FORCE_F64_ADD>(p, stream);
    else launch_bspline2<GFB_FORCE_F64_STORE>(p, stream);
}

static int enqueue_eval(gfb_kernel* k, int n_replicas, int n_particles, const double* d_pos, double* d_energies,
                        double* d_grid_energies, void* d_forces, int force_mode, long long force_stride,
                        const int* d_order, double* d_energies_clear, cudaStream_t stream, bool energy_store = false,
                        int atom_begin = 0, int atom_count = -1, bool overlap = false) {
    // atom_begin/atom_count: evaluate only atoms [atom_begin, atom_begin + atom_count) of a state without particle
    // indirection (the host path cuts one large replica into atom ranges); d_pos/d_forces then point at the range.
    const int n_atoms = atom_count >= 0 ? atom_count : k->n_atoms;
    EvalParams p;
    memset(&p, 0, sizeof p);
    p.energy_store = energy_store ? 1 : 0;
    p.pdl = overlap ? 1u : 0u;
    for (int g = 0; g < k->n_grids; g++) {
        fill_grid_view(k, g, p.grid[g]);
        p.grid[g].scaling += atom_begin;
    }
    p.n_grids = k->n_grids;
    p.n_atoms = n_atoms;
    p.n_particles = n_particles;
    p.n_replicas = n_replicas;
    p.total = (long long) n_replicas * n_atoms;
    p.pos = d_pos;
    p.particles = k->d_particles;
    p.order = d_order;
    p.slots = k->d_slots;
    p.n_slots = k->n_slots;
    p.energies = d_energies;
    p.grid_energies = d_grid_energies;
    p.energies_clear = d_energies_clear;
    p.forces = d_forces;
    p.force_stride = force_stride;
    if (p.total == 0) return GFB_OK;
    p.div_magic = (unsigned) std::min<unsigned long long>(0x100000000ull / (unsigned long long) n_atoms, 0xffffffffull);
    for (int a = 0; a < 3; a++) p.near_int[a] = 1.8e-15 * (double) std::max(1, p.grid[0].nc[a]);
    if (bspline_tiles_eligible(k, p)) {
        launch_bspline(p, force_mode, stream);
    } else if (lines_eligible(k, p)) {
        p.lines = k->d_interleaved;
        const int fpath = force_path_default(k->n_grids);
        switch (k->n_grids) {
            case 1: launch_lines2<1>(p, force_mode, fpath, stream); break;
            case 2: launch_lines2<2>(p, force_mode, fpath, stream); break;
            case 3: launch_lines2<3>(p, force_mode, fpath, stream); break;
            default: launch_lines2<4>(p, force_mode, fpath, stream); break;
        }
    } else {
        launch_eval1(p, k->precision, k->grids[0]->layout, k->same_geom, force_mode, stream);
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

static int check_exec_args(const char* fn, gfb_kernel* k, int n_replicas, int n_particles, const void* pos, int force_mode) {
    if (!k) return fail(GFB_ERR_INVALID, "%s: NULL kernel", fn);
    if (n_replicas < 0) return fail(GFB_ERR_INVALID, "%s: n_replicas=%d", fn, n_replicas);
    if (n_particles <= k->max_particle)
        return fail(GFB_ERR_INVALID, "%s: n_particles=%d but the kernel evaluates particle index %d", fn, n_particles, k->max_particle);
    if (!pos && (long long) n_replicas * k->n_atoms > 0) return fail(GFB_ERR_INVALID, "%s: positions pointer is NULL", fn);
    if (force_mode < GFB_FORCE_F64_STORE || force_mode > GFB_FORCE_FIXED_ADD)
        return fail(GFB_ERR_INVALID, "%s: unknown force_mode %d", fn, force_mode);
    if ((long long) n_replicas * (long long) n_particles > 2000000000LL)
        return fail(GFB_ERR_INVALID, "%s: %lld particles in one call; split the batch (limit 2e9)", fn,
                    (long long) n_replicas * n_particles);
    return GFB_OK;
}


// Threads per block of the kernel enqueue_eval would launch for this state without an evaluation order.
static int eval_block_threads(const gfb_kernel* k) {
    EvalParams probe;
    memset(&probe, 0, sizeof probe);
    if (bspline_tiles_eligible(k, probe)) return kBsBlock;
    return lines_eligible(k, probe) ? lines_block(k->n_grids) : kBlock;
}

// One replica of at most a few thousand particles — a ligand evaluated once per MD step (BASELINE configs[1]; what
// B200CalcGridForceKernel::execute issues). Such a call is pure latency, so instead of H2D copy -> kernel -> D2H copies
// (6-7 driver calls, 3 trips through the copy engines) the kernel works on HOST-MAPPED pinned memory directly: it reads
// the positions over PCIe, stores the forces over PCIe and — when one block covers the ligand — stores the energy too.
// One launch + one synchronize per call. The caller's forces are combined on the host from the kernel's stores
// (STORE: staged copy of the caller's array with the evaluated entries overwritten; ADD: caller's value + kernel's).
static int execute_host_small(gfb_kernel* k, int n_particles, const double* pos, double* energies, double* grid_energies,
                              double* forces, int force_mode) {
    gfb_device* dev = k->dev;
    const size_t np3 = (size_t) n_particles * 3;
    const int ng = k->n_grids;
    const size_t e_count = 1 + (size_t) ng;
    const size_t e_off = (2 * np3 + 15) & ~(size_t) 15;                    // doubles: [pos | forces | pad | energies]
    int rc = k->h_small.ensure((e_off + e_count) * sizeof(double));
    if (rc != GFB_OK) return rc;
    double* h_pos = static_cast<double*>(k->h_small.ptr);
    double* h_f = h_pos + np3;
    double* h_e = h_pos + e_off;
    memcpy(h_pos, pos, np3 * sizeof(double));
    if (forces) {
        if (force_mode == GFB_FORCE_F64_ADD) memset(h_f, 0, np3 * sizeof(double));
        else memcpy(h_f, forces, np3 * sizeof(double));
    }
    EvalParams probe;
    memset(&probe, 0, sizeof probe);
    // one block covers the ligand: its energy (and, in the lines kernel, its per-grid energies) are plain stores
    const bool one_block = k->n_atoms <= eval_block_threads(k) && (!grid_energies || !bspline_tiles_eligible(k, probe));
    double* d_e = nullptr;
    if (!one_block) {   // several blocks (or per-grid energies): device accumulators, cleared here, fetched below
        if ((rc = k->d_energy.ensure(e_count * sizeof(double))) != GFB_OK) return rc;
        d_e = static_cast<double*>(k->d_energy.ptr);
        CUDA_TRY(cudaMemsetAsync(d_e, 0, e_count * sizeof(double), dev->stream));
    }
    // cudaHostAlloc memory is mapped into the device's address space at the same address (unified addressing)
    double* e_dst = one_block ? h_e : d_e;
    rc = enqueue_eval(k, 1, n_particles, h_pos, e_dst, grid_energies ? e_dst + 1 : nullptr,
                      forces ? h_f : nullptr, GFB_FORCE_F64_STORE, 0, nullptr, nullptr, dev->stream, one_block);
    if (rc != GFB_OK) return rc;
    if (!one_block) CUDA_TRY(cudaMemcpyAsync(h_e, d_e, e_count * sizeof(double), cudaMemcpyDeviceToHost, dev->stream));
    CUDA_TRY(cudaStreamSynchronize(dev->stream));
    if (energies) energies[0] = h_e[0];
    if (grid_energies) memcpy(grid_energies, h_e + 1, ng * sizeof(double));
    if (forces) {
        if (force_mode == GFB_FORCE_F64_ADD) {
            for (size_t i = 0; i < np3; i++) forces[i] += h_f[i];
        } else {
            memcpy(forces, h_f, np3 * sizeof(double));
        }
    }
    return GFB_OK;
}

extern "C" {

int gfb_kernel_eval_path(const gfb_kernel* k) {
    if (!k) return 0;
    EvalParams probe;
    memset(&probe, 0, sizeof probe);
    return lines_eligible(k, probe) ? 1 : 0;
}

int gfb_kernel_execute_device(gfb_kernel* k, int n_replicas, int n_particles, const double* d_pos, double* d_energies,
                              double* d_grid_energies, void* d_forces, int force_mode, long long force_stride,
                              const int* d_order, double* d_energies_clear, void* stream) {
    int rc = check_exec_args("gfb_kernel_execute_device", k, n_replicas, n_particles, d_pos, force_mode);
    if (rc != GFB_OK) return rc;
    if (force_mode == GFB_FORCE_FIXED_ADD && d_forces && force_stride < (long long) n_replicas * n_particles)
        return fail(GFB_ERR_INVALID, "gfb_kernel_execute_device: force_stride=%lld < %lld particles", force_stride,
                    (long long) n_replicas * n_particles);
    CUDA_TRY(cudaSetDevice(k->dev->ordinal));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : k->dev->stream;
    return enqueue_eval(k, n_replicas, n_particles, d_pos, d_energies, d_grid_energies, d_forces, force_mode, force_stride,
                        d_order, d_energies_clear, s, false, 0, -1, k->launch_overlap);
}

// Host path. The batch is cut into replica chunks so that the H2D of chunk i+1, the kernel of chunk i and
// the D2H of chunk i-1 overlap (two streams + events). Pinned user buffers are DMA'd directly; pageable
// ones go through the handle's pinned staging buffer.
int gfb_kernel_execute_host(gfb_kernel* k, int n_replicas, int n_particles, const double* pos, double* energies,
                            double* grid_energies, double* forces, int force_mode) {
    int rc = check_exec_args("gfb_kernel_execute_host", k, n_replicas, n_particles, pos, force_mode);
    if (rc != GFB_OK) return rc;
    if (force_mode == GFB_FORCE_FIXED_ADD)
        return fail(GFB_ERR_INVALID, "gfb_kernel_execute_host: FIXED_ADD is a device-buffer format; use F64_STORE or F64_ADD");
    if (n_replicas == 0) return GFB_OK;
    gfb_device* dev = k->dev;
    std::lock_guard<std::mutex> host_lock(dev->host_mutex);   // several Contexts/threads may drive one GPU
    CUDA_TRY(cudaSetDevice(dev->ordinal));

    static const bool small_off = [] {
        const char* e = getenv("GFB_SMALL_PATH");   // 0: always take the copy pipeline (A/B measurements)
        return e && e[0] == '0';
    }();
    if (!small_off && n_replicas == 1 && k->d_slots == nullptr && k->unique_particles && n_particles <= 4096 && k->n_atoms > 0)
        return execute_host_small(k, n_particles, pos, energies, grid_energies, forces, force_mode);

    const size_t np = (size_t) n_replicas * n_particles;
    const size_t pos_bytes = np * 3 * sizeof(double);
    const int ng = k->n_grids;
    const size_t n_e = (size_t) n_replicas * k->n_slots;          // energy entries: [replica][slot]
    const size_t e_count = n_e * (1 + ng);
    // Forces straight into host memory: when every particle is an evaluated atom (no indirection), the mode is STORE
    // and the lines kernel runs, its warps emit their forces as full contiguous 768-byte runs (16-byte stores), which
    // the GPU writes over PCIe at copy-engine speed (51 GB/s measured) — so the D2H copies and the device force buffer
    // drop out of the pipeline, and the downloads overlap the uploads inside the kernels themselves. C5 (74 MB each
    // way): 2.11 -> 1.86 ms per step; the two directions together top out at ~80 GB/s on this box.
    static const bool zc_off = [] {
        const char* e = getenv("GFB_ZEROCOPY_FORCES");   // 0: always download with the copy engine (A/B measurements)
        return e && e[0] == '0';
    }();
    bool zc_forces = false;
    if (forces && force_mode == GFB_FORCE_F64_STORE && !zc_off && !k->d_particles && k->n_atoms == n_particles) {
        EvalParams probe;
        memset(&probe, 0, sizeof probe);
        zc_forces = lines_eligible(k, probe);
    }
    if ((rc = k->d_pos.ensure(pos_bytes)) != GFB_OK) return rc;
    if (forces && !zc_forces && (rc = k->d_forces.ensure(pos_bytes)) != GFB_OK) return rc;
    if ((rc = k->d_energy.ensure(e_count * sizeof(double))) != GFB_OK) return rc;
    if ((rc = k->h_energy.ensure(e_count * sizeof(double))) != GFB_OK) return rc;

    const bool pos_pinned = is_pinned(pos);
    const bool f_pinned = forces && is_pinned(forces);
    // staging: [pos | forces] when the user's buffers are pageable
    const size_t stage_bytes = (pos_pinned ? 0 : pos_bytes) + ((forces && !f_pinned) ? pos_bytes : 0);
    if (stage_bytes && (rc = k->h_stage.ensure(stage_bytes)) != GFB_OK) return rc;
    char* stage_pos = static_cast<char*>(k->h_stage.ptr);
    char* stage_f = stage_pos + (pos_pinned ? 0 : pos_bytes);

    double* d_pos = static_cast<double*>(k->d_pos.ptr);
    // where the kernels write forces: the device buffer, or (zero-copy) the pinned host destination itself
    double* d_f = nullptr;
    if (forces) d_f = !zc_forces ? static_cast<double*>(k->d_forces.ptr) : (f_pinned ? forces : reinterpret_cast<double*>(stage_f));
    double* d_e = static_cast<double*>(k->d_energy.ptr);
    double* d_ge = d_e + n_e;
    CUDA_TRY(cudaMemsetAsync(d_e, 0, e_count * sizeof(double), dev->stream));

    // Chunk pipeline over three streams: H2D(c) on h2d_stream -> [up c] -> kernel(c) on stream -> [done c] -> D2H(c) on
    // copy_stream. Uploads never wait for kernels, downloads overlap the next uploads (PCIe is full duplex: measured
    // 55 GB/s one way, 45-49 GB/s each way when both run). >= 2 MB of positions per chunk, at most 8 chunks: each
    // chunk costs ~15 us of copy/event overhead (4/8/16/32/64 chunks of C5's 74 MB: 2.02/1.99/2.10/2.43/2.80 ms).
    // What a chunk is made of: replicas, or — one large replica without particle indirection (C3: 1 M atoms) — atoms.
    const bool by_atoms = n_replicas == 1 && !k->d_particles && !k->d_slots && k->n_atoms == n_particles;
    const int n_units = by_atoms ? n_particles : n_replicas;
    const size_t unit_doubles = by_atoms ? 3 : (size_t) n_particles * 3;
    int n_chunks = 1;
    if (n_units > 1) {
        const char* env = getenv("GFB_HOST_CHUNKS");
        n_chunks = env ? atoi(env) : (int) std::min<size_t>(8, std::max<size_t>(1, pos_bytes / (2u << 20)));
        n_chunks = std::max(1, std::min(n_chunks, n_units));
    }
    while ((int) dev->events.size() < 2 * n_chunks) {
        cudaEvent_t ev;
        CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        dev->events.push_back(ev);
    }
    cudaEvent_t* up = dev->events.data();
    cudaEvent_t* done = dev->events.data() + n_chunks;
    // the energy memset (on `stream`) must precede every kernel; the first kernel waits on up[0] on the same stream
    const bool need_f_upload = forces && (force_mode == GFB_FORCE_F64_ADD || k->d_particles || k->n_atoms < n_particles);

    int status = GFB_OK;
    for (int c = 0; c < n_chunks && status == GFB_OK; c++) {
        int r0 = (int) ((long long) n_units * c / n_chunks);         // first unit (replica or atom) of the chunk
        int r1 = (int) ((long long) n_units * (c + 1) / n_chunks);
        if (by_atoms) {   // whole warps per range: keeps every range's positions 16-byte aligned (warp-staged loads)
            r0 &= ~31;
            if (c + 1 < n_chunks) r1 &= ~31;
        } else if (n_units >= 2 * n_chunks) {   // even replica boundaries: an odd particle count leaves odd replicas 8 mod 16
            r0 &= ~1;
            if (c + 1 < n_chunks) r1 &= ~1;
        }
        const size_t off = (size_t) r0 * unit_doubles;               // doubles
        const size_t cnt = (size_t) (r1 - r0) * unit_doubles;
        const double* src = pos + off;
        if (!pos_pinned) {
            memcpy(stage_pos + off * sizeof(double), src, cnt * sizeof(double));
            src = reinterpret_cast<const double*>(stage_pos) + off;
        }
        // One chunk (a single ligand per MD step, configs[1]) is pure latency: everything goes on one stream, no events.
        cudaStream_t up_stream = n_chunks == 1 ? dev->stream : dev->h2d_stream;
        cudaStream_t down_stream = n_chunks == 1 ? dev->stream : dev->copy_stream;
        cudaError_t err = cudaMemcpyAsync(d_pos + off, src, cnt * sizeof(double), cudaMemcpyHostToDevice, up_stream);
        if (err == cudaSuccess && need_f_upload) {
            // ADD: the caller's current forces are the accumulator's initial value.
            // STORE with a particle subset: untouched entries must come back as they were.
            const double* fsrc = forces + off;
            if (!f_pinned) {
                memcpy(stage_f + off * sizeof(double), fsrc, cnt * sizeof(double));
                fsrc = reinterpret_cast<const double*>(stage_f) + off;
            }
            err = cudaMemcpyAsync(d_f + off, fsrc, cnt * sizeof(double), cudaMemcpyHostToDevice, up_stream);
        }
        if (n_chunks > 1) {
            if (err == cudaSuccess) err = cudaEventRecord(up[c], dev->h2d_stream);
            if (err == cudaSuccess) err = cudaStreamWaitEvent(dev->stream, up[c], 0);
        }
        if (err != cudaSuccess) {
            status = fail(GFB_ERR_CUDA, "gfb_kernel_execute_host: H2D: %s", cudaGetErrorString(err));
            break;
        }
        if (by_atoms)   // every range accumulates into the one replica's energy entries
            status = enqueue_eval(k, 1, r1 - r0, d_pos + off, d_e, grid_energies ? d_ge : nullptr, d_f ? d_f + off : nullptr,
                                  force_mode, 0, nullptr, nullptr, dev->stream, false, r0, r1 - r0);
        else
            status = enqueue_eval(k, r1 - r0, n_particles, d_pos + off, d_e + (size_t) r0 * k->n_slots,
                                  grid_energies ? d_ge + (size_t) r0 * k->n_slots * ng : nullptr,
                                  d_f ? d_f + off : nullptr, force_mode, 0, nullptr, nullptr, dev->stream);
        if (status != GFB_OK) break;
        err = cudaSuccess;
        if (forces && !zc_forces) {
            if (n_chunks > 1) {
                err = cudaEventRecord(done[c], dev->stream);
                if (err == cudaSuccess) err = cudaStreamWaitEvent(dev->copy_stream, done[c], 0);
            }
            double* dst = f_pinned ? forces + off : reinterpret_cast<double*>(stage_f) + off;
            if (err == cudaSuccess)
                err = cudaMemcpyAsync(dst, d_f + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, down_stream);
        }
        if (err != cudaSuccess) status = fail(GFB_ERR_CUDA, "gfb_kernel_execute_host: D2H: %s", cudaGetErrorString(err));
    }
    if (status == GFB_OK) {
        cudaError_t err = cudaMemcpyAsync(k->h_energy.ptr, d_e, e_count * sizeof(double), cudaMemcpyDeviceToHost, dev->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(dev->stream);
        if (err == cudaSuccess && n_chunks > 1) err = cudaStreamSynchronize(dev->copy_stream);
        if (err != cudaSuccess) status = fail(GFB_ERR_CUDA, "gfb_kernel_execute_host: %s", cudaGetErrorString(err));
    } else {
        cudaStreamSynchronize(dev->h2d_stream);
        cudaStreamSynchronize(dev->stream);
        cudaStreamSynchronize(dev->copy_stream);
    }
    if (status != GFB_OK) return status;

    const double* he = static_cast<const double*>(k->h_energy.ptr);
    if (energies) memcpy(energies, he, n_e * sizeof(double));
    if (grid_energies) memcpy(grid_energies, he + n_e, n_e * ng * sizeof(double));
    if (forces && !f_pinned) memcpy(forces, stage_f, pos_bytes);
    return GFB_OK;
}

int gfb_kernel_classify_host(gfb_kernel* k, int grid_index, int n_replicas, int n_particles, const double* pos, gfb_class* cls) {
    int rc = check_exec_args("gfb_kernel_classify_host", k, n_replicas, n_particles, pos, GFB_FORCE_F64_STORE);
    if (rc != GFB_OK) return rc;
    if (grid_index < 0 || grid_index >= k->n_grids) return fail(GFB_ERR_INVALID, "gfb_kernel_classify_host: grid_index=%d", grid_index);
    if (!cls) return fail(GFB_ERR_INVALID, "gfb_kernel_classify_host: cls is NULL");
    const long long total = (long long) n_replicas * k->n_atoms;
    if (total == 0) return GFB_OK;
    gfb_device* dev = k->dev;
    std::lock_guard<std::mutex> host_lock(dev->host_mutex);
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    const size_t pos_bytes = (size_t) n_replicas * n_particles * 3 * sizeof(double);
    if ((rc = k->d_pos.ensure(pos_bytes)) != GFB_OK) return rc;
    if ((rc = k->d_cls.ensure((size_t) total * sizeof(gfb_class))) != GFB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(k->d_pos.ptr, pos, pos_bytes, cudaMemcpyHostToDevice, dev->stream));
    ClassifyParams p;
    memset(&p, 0, sizeof p);
    fill_grid_view(k, grid_index, p.grid);
    p.n_atoms = k->n_atoms;
    p.n_particles = n_particles;
    p.total = total;
    p.pos = static_cast<const double*>(k->d_pos.ptr);
    p.particles = k->d_particles;
    p.out = static_cast<gfb_class*>(k->d_cls.ptr);
    const unsigned blocks = (unsigned) ((total + 255) / 256);
    // the classification code of the kernel that would evaluate this state: lines kernel (MIXED packed cells of one
    // geometry) or the general one
    EvalParams probe;
    memset(&probe, 0, sizeof probe);
    if (lines_eligible(k, probe)) {
        for (int a = 0; a < 3; a++) p.near_int[a] = 1.8e-15 * (double) std::max(1, p.grid.nc[a]);
        gf_classify_lines_kernel<<<blocks, 256, 0, dev->stream>>>(p);
    } else if (k->precision == GFB_PRECISION_DOUBLE) {
        gf_classify_kernel<true><<<blocks, 256, 0, dev->stream>>>(p);
    } else {
        gf_classify_kernel<false><<<blocks, 256, 0, dev->stream>>>(p);
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(cls, k->d_cls.ptr, (size_t) total * sizeof(gfb_class), cudaMemcpyDeviceToHost, dev->stream));
    CUDA_TRY(cudaStreamSynchronize(dev->stream));
    return GFB_OK;
}

int gfb_kernel_sort_atoms(gfb_kernel* k, int n_replicas, int n_particles, const double* d_pos, int* d_order, void* stream) {
    int rc = check_exec_args("gfb_kernel_sort_atoms", k, n_replicas, n_particles, d_pos, GFB_FORCE_F64_STORE);
    if (rc != GFB_OK) return rc;
    if (!d_order) return fail(GFB_ERR_INVALID, "gfb_kernel_sort_atoms: d_order is NULL");
    const long long total = (long long) n_replicas * k->n_atoms;
    if (total == 0) return GFB_OK;
    if (total > 0x7fffffffLL) return fail(GFB_ERR_INVALID, "gfb_kernel_sort_atoms: %lld atoms exceed the int32 order index", total);
    gfb_device* dev = k->dev;
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : dev->stream;
    const int n = (int) total;
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (unsigned*) nullptr, (unsigned*) nullptr, (int*) nullptr,
                                             (int*) nullptr, n, 0, 32, s));
    const size_t key_bytes = ((size_t) n * sizeof(unsigned) + 255) & ~(size_t) 255;
    const size_t idx_bytes = ((size_t) n * sizeof(int) + 255) & ~(size_t) 255;
    if ((rc = k->d_sort.ensure(2 * key_bytes + idx_bytes + tmp_bytes)) != GFB_OK) return rc;
    char* base = static_cast<char*>(k->d_sort.ptr);
    unsigned* keys_in = reinterpret_cast<unsigned*>(base);
    unsigned* keys_out = reinterpret_cast<unsigned*>(base + key_bytes);
    int* idx_in = reinterpret_cast<int*>(base + 2 * key_bytes);
    void* tmp = base + 2 * key_bytes + idx_bytes;
    ClassifyParams p;
    memset(&p, 0, sizeof p);
    fill_grid_view(k, 0, p.grid);
    p.n_atoms = k->n_atoms;
    p.n_particles = n_particles;
    p.total = total;
    p.pos = d_pos;
    p.particles = k->d_particles;
    gf_morton_key_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, s>>>(p, keys_in, idx_in);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, idx_in, d_order, n, 0, 32, s));
    return GFB_OK;
}

int gfb_forces_fixed_to_f64(gfb_device* dev, const void* d_fixed, long long force_stride, long long n, double* d_out, void* stream) {
    if (!dev || !d_fixed || !d_out) return fail(GFB_ERR_INVALID, "gfb_forces_fixed_to_f64: NULL argument");
    if (n <= 0) return GFB_OK;
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : dev->stream;
    gf_fixed_to_f64_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(static_cast<const long long*>(d_fixed), force_stride, n, d_out);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

int gfb_peer_put(gfb_device* dev, const void* d_src, void* const* peer_dst, int n_peers, size_t dst_offset, size_t bytes,
                 int first_peer, void* stream) {
    if (!dev || !d_src || !peer_dst || n_peers < 1) return fail(GFB_ERR_INVALID, "gfb_peer_put: NULL argument");
    if (bytes == 0) return GFB_OK;
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : dev->stream;
    for (int i = 0; i < n_peers; i++) {
        const int p = ((first_peer % n_peers) + n_peers + i) % n_peers;
        if (!peer_dst[p]) return fail(GFB_ERR_INVALID, "gfb_peer_put: peer_dst[%d] is NULL", p);
        CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(peer_dst[p]) + dst_offset, d_src, bytes, cudaMemcpyDefault, s));
    }
    return GFB_OK;
}

int gfb_bench_sector_gather(gfb_device* dev, size_t bytes, long long n_loads, int reps, double* gbs) {
    if (!dev || !gbs) return fail(GFB_ERR_INVALID, "gfb_bench_sector_gather: NULL argument");
    if (bytes < 32 || n_loads < 1 || reps == 0) return fail(GFB_ERR_INVALID, "gfb_bench_sector_gather: bad sizes");
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    float* buf = nullptr;
    float* sink = nullptr;
    CUDA_TRY(cudaMalloc((void**) &buf, bytes));
    CUDA_TRY(cudaMalloc((void**) &sink, 256));
    CUDA_TRY(cudaMemsetAsync(buf, 0, bytes, dev->stream));
    const int per_thread = 16;
    const long long threads = (n_loads + per_thread - 1) / per_thread;
    const unsigned blocks = (unsigned) ((threads + 255) / 256);
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    // reps < 0: 16-byte loads instead of 32-byte sectors (probe for the row-chunked layouts)
    const bool chunk16 = reps < 0;
    if (chunk16) reps = -reps;
    auto launch = [&]() {
        if (chunk16) gf_chunk_gather_kernel<<<blocks, 256, 0, dev->stream>>>(reinterpret_cast<const float4*>(buf), bytes / 16, per_thread, sink);
        else gf_sector_gather_kernel<<<blocks, 256, 0, dev->stream>>>(buf, bytes / 32, per_thread, sink);
    };
    for (int i = 0; i < 3; i++) launch();
    CUDA_TRY(cudaEventRecord(e0, dev->stream));
    for (int i = 0; i < reps; i++) launch();
    CUDA_TRY(cudaEventRecord(e1, dev->stream));
    g_launches += reps + 3;
    CUDA_TRY(cudaStreamSynchronize(dev->stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(sink);
    *gbs = (double) blocks * 256.0 * per_thread * (chunk16 ? 16.0 : 32.0) * reps / (ms * 1e-3) / 1e9;
    return GFB_OK;
}

}  // extern "C"
