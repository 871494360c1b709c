// C ABI of libgridforce_b200.so (include/gridforce_b200.h): handles, uploads, launches.
// Host logic only — every number is produced by the kernels in gf_kernels.cuh. No CPU fallback.
#include <cstdarg>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "gf_handles.h"
#include "gf_launch.h"

using namespace gfb;

namespace gfb {
cudaError_t interleave_cells(gfb_device* dev, const void* const src[4], void* dst, size_t n_cells, int slots, int bytes_per_slot);   // gf_grids.cu

// ---------------------------------------------------------------------------------------------
// Errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_error;
std::atomic<unsigned long long> g_launches(0);

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}
}  // namespace gfb

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
        if (!grids[g]->cells) return fail(GFB_ERR_INVALID, "gfb_kernel_create: grids[%d] has released its packed cells (gfb_grid_release_cells)", g);
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
    k->want_atom_energies = false;
    k->atom_e_count = 0;
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
    // 2-4 packed-cell grids of one geometry: weave them into one record per cell (4 slots of 32 bytes = one 128-byte
    // line in MIXED, 4 slots of 64 bytes = two lines in DOUBLE), so that everything an atom needs from all its grids is
    // ONE record (gf_eval_lines.cuh / gf_eval_lines_f64.cuh fetch it with cp.async, full lines per request).
    // Costs a second copy of the grids; GFB_LINES=0 keeps the per-grid arrays only.
    const char* il = getenv("GFB_LINES");
    const bool want_lines = !(il && il[0] == '0');
    const size_t slot_bytes = k->precision == GFB_PRECISION_MIXED ? 32 : 64;
    const size_t n_cells0 = grids[0]->bytes / slot_bytes;
    k->n_cells = grids[0]->layout == GFB_LAYOUT_CELLS ? n_cells0 : 0;
    if (want_lines && k->same_geom && n_grids >= 2 && n_grids <= 4 && grids[0]->layout == GFB_LAYOUT_CELLS &&
        n_cells0 < 0xffffffffull && n_cells0 * 4 * slot_bytes <= dev->prop.totalGlobalMem / 8) {
        const int slots = 4;
        cudaError_t e = cudaMalloc(&k->d_interleaved, n_cells0 * slots * slot_bytes);
        if (e == cudaSuccess) {
            const void* src[4] = {nullptr, nullptr, nullptr, nullptr};
            for (int g = 0; g < n_grids; g++) src[g] = grids[g]->cells;
            e = interleave_cells(dev, src, k->d_interleaved, n_cells0, slots, (int) slot_bytes);
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
    resident_destroy(k);
    cudaStreamSynchronize(k->dev->stream);
    cudaStreamSynchronize(k->dev->copy_stream);
    if (k->d_scaling) cudaFree(k->d_scaling);
    if (k->d_particles) cudaFree(k->d_particles);
    if (k->d_interleaved) cudaFree(k->d_interleaved);
    if (k->d_slots) cudaFree(k->d_slots);
    k->d_pos.release();
    k->d_forces.release();
    k->d_energy.release();
    k->d_small_e.release();
    k->d_cls.release();
    k->d_sort.release();
    k->d_atom_e.release();
    k->h_stage.release();
    k->h_energy.release();
    k->h_small.release();
    delete k;
    return GFB_OK;
}

int gfb_kernel_set_energy_slots(gfb_kernel* k, const int* slots, int n_slots) {
    if (!k) return fail(GFB_ERR_INVALID, "gfb_kernel_set_energy_slots: NULL kernel");
    CUDA_TRY(cudaSetDevice(k->dev->ordinal));
    int rs = resident_stop(k);
    if (rs != GFB_OK) return rs;
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
    int rs = resident_stop(k);      // a resident block holds inv-power in its arguments and may hold scaling factors in L1
    if (rs != GFB_OK) return rs;
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
namespace gfb {
void fill_grid_view(const gfb_kernel* k, int g, GridView& v) {
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
}  // namespace gfb

// ---- which kernel serves a state -----------------------------------------------------------------------------------
// How the ADD force modes of gf_eval_lines_kernel reach memory: 0 RED atomics alone, 1 RED + L2 prefetch of the force
// lines at kernel start. Measured (B200): the prefetch takes C3 (one grid: the force lines are a third of
// the traffic) from 45.5 to 35.1 us and changes nothing for three grids. Default: prefetch for one grid.
// GFB_FORCE_PATH=0|1 overrides (probe).
static int force_path_default(int n_grids) {
    static const int v = [] {
        const char* e = getenv("GFB_FORCE_PATH");
        return e ? std::max(0, std::min(1, atoi(e))) : -1;
    }();
    return v >= 0 ? v : (n_grids == 1 ? 1 : 0);
}

static bool env_off(const char* name) {
    const char* e = getenv(name);
    return e && e[0] == '0';
}

namespace gfb {
// gf_eval_lines_kernel: MIXED packed cells of one geometry without inv-power; 1 grid from its own cells, 2-4 from records.
bool lines_eligible(const gfb_kernel* k, const EvalParams&) {
    static const bool off = env_off("GFB_LINES");
    if (off || k->precision != GFB_PRECISION_MIXED || k->grids[0]->layout != GFB_LAYOUT_CELLS || !k->same_geom) return false;
    if (k->n_grids > 4) return false;
    if (k->n_grids > 1 && !k->il_slots) return false;
    if (k->n_cells >= 0xffffffffull) return false;
    for (int g = 0; g < k->n_grids; g++)
        if (k->inv_power[g] > 0.0) return false;
    return true;
}

// gf_eval_lines_f64_kernel: DOUBLE 256-byte records, 2-4 grids of one geometry without inv-power.
bool lines_f64_eligible(const gfb_kernel* k) {
    static const bool off = env_off("GFB_LINES") || env_off("GFB_LINES_F64");
    if (off || k->precision != GFB_PRECISION_DOUBLE || k->grids[0]->layout != GFB_LAYOUT_CELLS || !k->same_geom) return false;
    if (k->n_grids < 2 || k->n_grids > 4 || !k->il_slots) return false;
    if (k->n_cells >= 0xffffffffull) return false;
    for (int g = 0; g < k->n_grids; g++)
        if (k->inv_power[g] > 0.0) return false;
    return true;
}

// gf_eval_bspline_kernel (gf_eval_bspline.cuh): MIXED B-spline tiles of one geometry, no evaluation order.
// `layout` = GFB_LAYOUT_BSPLINE (method 1) or GFB_LAYOUT_HERMITE (method 2: the same kernel, other arithmetic).
static bool record_tiles_eligible(const gfb_kernel* k, const EvalParams& p, int layout) {
    static const bool off = env_off("GFB_BSPLINE_TILES");   // 0: always the general kernel (A/B measurements)
    if (off || k->precision != GFB_PRECISION_MIXED || k->grids[0]->layout != layout || !k->same_geom) return false;
    if (p.order != nullptr) return false;
    return k->grids[0]->bytes / 128 < 0x7fffffffull;   // 32-bit record index
}
static bool bspline_tiles_eligible(const gfb_kernel* k, const EvalParams& p) { return record_tiles_eligible(k, p, GFB_LAYOUT_BSPLINE); }
static bool tricubic_tiles_eligible(const gfb_kernel* k, const EvalParams& p) { return record_tiles_eligible(k, p, GFB_LAYOUT_HERMITE); }

// gf_eval_bspline_f64_kernel (gf_eval_bspline_f64.cuh): DOUBLE B-spline records of one geometry, no evaluation order.
static bool record_f64_eligible(const gfb_kernel* k, const EvalParams& p, int layout) {
    static const bool off = env_off("GFB_BSPLINE_TILES") || env_off("GFB_BSPLINE_F64");   // 0: the general kernel (A/B measurements)
    if (off || k->precision != GFB_PRECISION_DOUBLE || k->grids[0]->layout != layout || !k->same_geom) return false;
    if (p.order != nullptr) return false;
    return k->grids[0]->bytes / 256 < 0x7fffffffull;   // 32-bit record index
}
static bool bspline_f64_eligible(const gfb_kernel* k, const EvalParams& p) { return record_f64_eligible(k, p, GFB_LAYOUT_BSPLINE); }
static bool tricubic_f64_eligible(const gfb_kernel* k, const EvalParams& p) { return record_f64_eligible(k, p, GFB_LAYOUT_HERMITE); }

int enqueue_eval(gfb_kernel* k, int n_replicas, int n_particles, const double* d_pos, double* d_energies,
                 double* d_grid_energies, void* d_forces, int force_mode, long long force_stride, const int* d_order,
                 double* d_energies_clear, cudaStream_t stream, const EvalExtra& x) {
    // atom_begin/atom_count: evaluate only atoms [atom_begin, atom_begin + atom_count) of a state without particle
    // indirection (the host path cuts one large replica into atom ranges); d_pos/d_forces then point at the range.
    const int n_atoms = x.atom_count >= 0 ? x.atom_count : k->n_atoms;
    EvalParams p;
    memset(&p, 0, sizeof p);
    p.energy_store = x.energy_store ? 1 : 0;
    p.pdl = x.overlap ? 1u : 0u;
    for (int g = 0; g < k->n_grids; g++) {
        fill_grid_view(k, g, p.grid[g]);
        p.grid[g].scaling += x.atom_begin;
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
    p.atom_energies = x.atom_energies;
    p.forces = d_forces;
    p.force_mode = force_mode;
    p.force_stride = force_stride;
    p.gather = x.gather;
    p.gather_offset = x.gather_offset;
    if (p.total == 0 && !x.gather) return GFB_OK;
    p.div_magic = (unsigned) std::min<unsigned long long>(0x100000000ull / (unsigned long long) std::max(n_atoms, 1), 0xffffffffull);
    for (int a = 0; a < 3; a++) p.near_int[a] = 1.8e-15 * (double) std::max(1, p.grid[0].nc[a]);
    const bool lines = lines_eligible(k, p), lines64 = !lines && lines_f64_eligible(k);
    if (!((lines && k->n_grids > 1) || lines64))
        for (int g = 0; g < k->n_grids; g++)
            if (!k->grids[g]->cells)
                return fail(GFB_ERR_INVALID, "grid %d released its packed cells (gfb_grid_release_cells) but this evaluation needs them "
                                             "(one grid, inv-power, more than 4 grids or GFB_LINES=0)", g);
    if (x.gather) {
        if (!lines && !lines64)
            return fail(GFB_ERR_UNSUPPORTED, "fused energy gather needs the record kernels (packed cells, one geometry, no inv-power)");
        if (!d_energies) return fail(GFB_ERR_INVALID, "fused energy gather needs d_energies");
        if (p.total == 0) return fail(GFB_ERR_INVALID, "fused energy gather needs at least one replica on every rank");
    }
    if (bspline_tiles_eligible(k, p)) {
        launch_bspline(p, stream);
    } else if (tricubic_tiles_eligible(k, p)) {
        launch_tricubic_records(p, stream);
    } else if (bspline_f64_eligible(k, p)) {
        launch_bspline_f64(p, stream);
    } else if (tricubic_f64_eligible(k, p)) {
        launch_tricubic_records_f64(p, stream);
    } else if (lines) {
        p.lines = k->d_interleaved;
        static const bool ahead_off = env_off("GFB_POS_PREFETCH");   // A/B measurements
        const int block = lines_block_threads(k->n_grids);
        const int per_sm = k->n_grids == 1 ? 6 : 1280 / block;
        const unsigned resident = (unsigned) (k->dev->prop.multiProcessorCount * per_sm);
        if (!ahead_off) p.ahead_blocks = resident;
        // Launch overlap, small launches: the launcher may pick the tile-striding variant (a resident grid, energy atomics
        // parked until the block's end; gf_eval_lines.cuh / gf_launch_lines.cu). Needs commutative (ADD) or no force
        // writes and nothing else stored.
        static const bool defer_off = env_off("GFB_DEFER");
        const bool single = n_replicas == 1 && k->d_slots == nullptr;
        const bool add_or_none = d_forces == nullptr || force_mode == GFB_FORCE_FIXED_ADD || force_mode == GFB_FORCE_F64_ADD;
        const bool plain = k->d_particles == nullptr && n_particles == n_atoms && d_order == nullptr && k->d_slots == nullptr &&
                           d_energies != nullptr && (reinterpret_cast<uintptr_t>(d_pos) & 15) == 0;
        if (x.overlap && !defer_off && !single && plain && add_or_none && !x.gather && !x.atom_energies && !d_grid_energies) {
            p.defer = 1u;
            p.persist_blocks = (unsigned) k->dev->prop.multiProcessorCount;
        }
        launch_lines(p, force_mode, force_path_default(k->n_grids), stream);
    } else if (lines64) {
        p.lines = k->d_interleaved;
        launch_lines_f64(p, force_mode, stream);
    } else if (k->precision == GFB_PRECISION_DOUBLE) {
        launch_general_f64(p, k->grids[0]->layout, k->same_geom, stream);
    } else {
        launch_general_f32(p, k->grids[0]->layout, k->same_geom, stream);
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

int check_exec_args(const char* fn, gfb_kernel* k, int n_replicas, int n_particles, const void* pos, int force_mode) {
    if (!k) return fail(GFB_ERR_INVALID, "%s: NULL kernel", fn);
    if (n_replicas < 0) return fail(GFB_ERR_INVALID, "%s: n_replicas=%d", fn, n_replicas);
    if (n_particles <= k->max_particle)
        return fail(GFB_ERR_INVALID, "%s: n_particles=%d but the kernel evaluates particle index %d", fn, n_particles, k->max_particle);
    if (!pos && (long long) n_replicas * k->n_atoms > 0) return fail(GFB_ERR_INVALID, "%s: positions pointer is NULL", fn);
    if (force_mode < GFB_FORCE_F64_STORE || force_mode > GFB_FORCE_F32_STORE)
        return fail(GFB_ERR_INVALID, "%s: unknown force_mode %d", fn, force_mode);
    if ((long long) n_replicas * (long long) n_particles > 2000000000LL)
        return fail(GFB_ERR_INVALID, "%s: %lld particles in one call; split the batch (limit 2e9)", fn,
                    (long long) n_replicas * n_particles);
    return GFB_OK;
}
}  // namespace gfb

// Threads per block of the kernel enqueue_eval would launch for this state without an evaluation order.
static int eval_block_threads(const gfb_kernel* k) {
    EvalParams probe;
    memset(&probe, 0, sizeof probe);
    if (bspline_tiles_eligible(k, probe) || tricubic_tiles_eligible(k, probe)) return kBsplineBlockThreads;
    if (bspline_f64_eligible(k, probe) || tricubic_f64_eligible(k, probe)) return kBsplineF64BlockThreads;
    if (lines_eligible(k, probe)) return lines_block_threads(k->n_grids);
    return lines_f64_eligible(k) ? kLinesF64BlockThreads : kGeneralBlock;
}

// A few thousand particles in all — one ligand evaluated once per MD step (BASELINE configs[1]; what
// B200CalcGridForceKernel::execute issues), or a handful of replicas of it (example/input.json: nstate = 21 — the size of a
// sampler.py sweep through GridForceBatch). Such a call is pure latency, so instead of H2D copy -> kernel -> D2H copies
// (6-7 driver calls, 3 trips through the copy engines) the kernel works on HOST-MAPPED pinned memory directly: it reads
// the positions over PCIe, stores the forces over PCIe and — when one block covers the ligand — stores the energy too.
// One launch + one synchronize per call. The caller's forces are combined on the host from the kernel's stores
// (STORE: staged copy of the caller's array with the evaluated entries overwritten; ADD: caller's value + kernel's).
static int execute_host_small(gfb_kernel* k, int n_replicas, int n_particles, const double* pos, double* energies,
                              double* grid_energies, void* forces, int force_mode) {
    gfb_device* dev = k->dev;
    if (n_replicas == 1 && resident_enabled(k) && !k->want_atom_energies && force_mode != GFB_FORCE_F32_STORE)   // gfb_kernel_set_resident
        return resident_step(k, n_particles, pos, energies, grid_energies, static_cast<double*>(forces), force_mode == GFB_FORCE_F64_ADD);
    const size_t R = (size_t) n_replicas;
    const size_t np3 = R * (size_t) n_particles * 3;
    const int ng = k->n_grids;
    const size_t e_count = R * (1 + (size_t) ng);                           // [R totals | R x ng per grid], as the batch path lays them out
    const size_t e_off = (2 * np3 + 15) & ~(size_t) 15;                    // doubles: [pos | forces | pad | energies | atom energies]
    const size_t ae_count = k->want_atom_energies ? R * (size_t) k->n_atoms : 0;
    int rc = k->h_small.ensure((e_off + e_count + ae_count) * sizeof(double));
    if (rc != GFB_OK) return rc;
    double* h_pos = static_cast<double*>(k->h_small.ptr);
    double* h_f = h_pos + np3;                                              // doubles, or floats in F32_STORE
    double* h_e = h_pos + e_off;
    const bool f32 = force_mode == GFB_FORCE_F32_STORE;
    const size_t f_bytes = np3 * (f32 ? sizeof(float) : sizeof(double));
    memcpy(h_pos, pos, np3 * sizeof(double));
    if (forces) {
        if (force_mode == GFB_FORCE_F64_ADD) memset(h_f, 0, f_bytes);
        else memcpy(h_f, forces, f_bytes);
    }
    EvalParams probe;
    memset(&probe, 0, sizeof probe);
    // one replica that one block covers: its energy (and, in the record kernels, its per-grid energies) are plain stores
    const bool one_block = n_replicas == 1 && k->n_atoms <= eval_block_threads(k) &&
                           (!grid_energies || !(bspline_tiles_eligible(k, probe) || tricubic_tiles_eligible(k, probe) || bspline_f64_eligible(k, probe) || tricubic_f64_eligible(k, probe)));
    // Several blocks or replicas (or per-grid energies): the energies are sums over blocks, accumulated with red.add.f64.
    // For up to 16 accumulators (4 replicas x 3 grids) they are the HOST-MAPPED array itself (zeroed here by the host; the
    // GPU's atomics on mapped memory are atomic among its own threads, which is all that is needed while the host waits):
    // no memset and no copy on the stream, the call stays one launch + one synchronise (2 replicas: 34.9 -> 24.9 us).
    // Each such atomic is a PCIe round trip, though (~1.3 us, serialised per address): at 8 replicas it is a draw and at
    // 21 it loses (49 against 36 us; profiles/logs_r2/r2v17_small_batch_*), so larger arrays accumulate on the device
    // and come back with one D2H copy. GFB_SMALL_HOST_ATOMICS=0: always the device route (A/B measurements).
    static const bool host_atomics_off = env_off("GFB_SMALL_HOST_ATOMICS");
    const bool host_acc = !one_block && !host_atomics_off && e_count <= 16;
    double* d_e = nullptr;
    double* d_clear = nullptr;     // the accumulator of the NEXT call, zero-filled by this launch (saves a memset per call)
    if (host_acc) {
        memset(h_e, 0, e_count * sizeof(double));
    } else if (!one_block) {
        // Two accumulator arrays take turns. Without per-grid energies only the R totals are used, and the launch that
        // accumulates into one array clears the other's totals (EvalParams::energies_clear), so a steady sequence of such
        // calls needs no memset; anything else (first call, more replicas than last time, per-grid energies) memsets.
        if ((rc = k->d_small_e.ensure(2 * e_count * sizeof(double))) != GFB_OK) return rc;
        double* base = static_cast<double*>(k->d_small_e.ptr);
        if (k->small_stride != e_count || k->small_base != base) {      // the arrays moved or changed size: nothing is known to be zero
            k->small_base = base;
            k->small_stride = e_count;
            k->small_zeroed = 0;
            k->small_toggle = 0;
        }
        d_e = base + (size_t) k->small_toggle * e_count;
        const size_t need = grid_energies ? e_count : R;
        if (grid_energies || k->small_zeroed < need) CUDA_TRY(cudaMemsetAsync(d_e, 0, need * sizeof(double), dev->stream));
        if (!grid_energies) {
            d_clear = base + (size_t) (k->small_toggle ^ 1) * e_count;
            k->small_zeroed = R;               // after this launch the other array's first R entries are zero
        } else {
            k->small_zeroed = 0;
        }
        k->small_toggle ^= 1;
    }
    // cudaHostAlloc memory is mapped into the device's address space at the same address (unified addressing)
    double* e_dst = (one_block || host_acc) ? h_e : d_e;
    EvalExtra x;
    x.energy_store = one_block;
    x.atom_energies = ae_count ? h_e + e_count : nullptr;
    rc = enqueue_eval(k, n_replicas, n_particles, h_pos, e_dst, grid_energies ? e_dst + R : nullptr, forces ? h_f : nullptr,
                      f32 ? GFB_FORCE_F32_STORE : GFB_FORCE_F64_STORE, 0, nullptr, d_clear, dev->stream, x);
    if (rc != GFB_OK) {
        k->small_zeroed = 0;
        return rc;
    }
    if (d_e) CUDA_TRY(cudaMemcpyAsync(h_e, d_e, (grid_energies ? e_count : R) * sizeof(double), cudaMemcpyDeviceToHost, dev->stream));
    CUDA_TRY(cudaStreamSynchronize(dev->stream));
    if (energies) memcpy(energies, h_e, R * sizeof(double));
    if (grid_energies) memcpy(grid_energies, h_e + R, R * ng * sizeof(double));
    if (ae_count) {
        if ((rc = k->d_atom_e.ensure(ae_count * sizeof(double))) != GFB_OK) return rc;
        CUDA_TRY(cudaMemcpy(k->d_atom_e.ptr, h_e + e_count, ae_count * sizeof(double), cudaMemcpyHostToDevice));
        k->atom_e_count = (long long) ae_count;
    }
    if (forces) {
        if (force_mode == GFB_FORCE_F64_ADD) {
            double* f = static_cast<double*>(forces);
            for (size_t i = 0; i < np3; i++) f[i] += h_f[i];
        } else {
            memcpy(forces, h_f, f_bytes);
        }
    }
    return GFB_OK;
}

// Device-visible address of a page-locked host buffer (cudaHostAlloc or cudaHostRegister), or NULL.
static void* host_device_pointer(void* p) {
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return d;
}

extern "C" {

int gfb_kernel_eval_path(const gfb_kernel* k) {
    if (!k) return 0;
    EvalParams probe;
    memset(&probe, 0, sizeof probe);
    if (lines_eligible(k, probe)) return 1;
    if (lines_f64_eligible(k)) return 2;
    if (bspline_tiles_eligible(k, probe)) return 3;
    if (tricubic_tiles_eligible(k, probe)) return 5;
    if (tricubic_f64_eligible(k, probe)) return 6;
    return bspline_f64_eligible(k, probe) ? 4 : 0;
}

int gfb_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return fail(GFB_ERR_INVALID, "gfb_host_register: NULL or empty buffer");
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return GFB_OK;
    }
    if (e != cudaSuccess) return fail(GFB_ERR_CUDA, "gfb_host_register: %s", cudaGetErrorString(e));
    return GFB_OK;
}

int gfb_host_unregister(void* ptr) {
    if (!ptr) return GFB_OK;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(GFB_ERR_CUDA, "gfb_host_unregister: %s", cudaGetErrorString(e));
    }
    return GFB_OK;
}

int gfb_kernel_request_atom_energies(gfb_kernel* k, int enable) {
    if (!k) return fail(GFB_ERR_INVALID, "gfb_kernel_request_atom_energies: NULL kernel");
    k->want_atom_energies = enable != 0;
    if (!enable) k->atom_e_count = 0;
    return GFB_OK;
}

int gfb_kernel_get_atom_energies(gfb_kernel* k, double* out, size_t n) {
    if (!k || (!out && n)) return fail(GFB_ERR_INVALID, "gfb_kernel_get_atom_energies: NULL argument");
    if ((long long) n != k->atom_e_count)
        return fail(GFB_ERR_INVALID, "gfb_kernel_get_atom_energies: %zu entries asked, the last host call produced %lld", n, k->atom_e_count);
    if (n == 0) return GFB_OK;
    CUDA_TRY(cudaSetDevice(k->dev->ordinal));
    CUDA_TRY(cudaMemcpy(out, k->d_atom_e.ptr, n * sizeof(double), cudaMemcpyDeviceToHost));
    return GFB_OK;
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
    EvalExtra x;
    x.overlap = k->launch_overlap;
    return enqueue_eval(k, n_replicas, n_particles, d_pos, d_energies, d_grid_energies, d_forces, force_mode, force_stride,
                        d_order, d_energies_clear, s, x);
}

// Host path. The batch is cut into replica chunks so that the H2D of chunk i+1, the kernel of chunk i and
// the D2H of chunk i-1 overlap (two streams + events). Page-locked user buffers (cudaHostAlloc, or registered with
// gfb_host_register) are DMA'd directly; pageable ones go through the handle's pinned staging buffer.
int gfb_kernel_execute_host(gfb_kernel* k, int n_replicas, int n_particles, const double* pos, double* energies,
                            double* grid_energies, void* forces, int force_mode) {
    int rc = check_exec_args("gfb_kernel_execute_host", k, n_replicas, n_particles, pos, force_mode);
    if (rc != GFB_OK) return rc;
    if (force_mode == GFB_FORCE_FIXED_ADD)
        return fail(GFB_ERR_INVALID, "gfb_kernel_execute_host: FIXED_ADD is a device-buffer format; use F64_STORE, F32_STORE or F64_ADD");
    if (n_replicas == 0) return GFB_OK;
    gfb_device* dev = k->dev;
    std::lock_guard<std::mutex> host_lock(dev->host_mutex);   // several Contexts/threads may drive one GPU
    CUDA_TRY(cudaSetDevice(dev->ordinal));

    static const bool small_off = env_off("GFB_SMALL_PATH");   // 0: always take the copy pipeline (A/B measurements)
    static const bool small_batch_off = env_off("GFB_SMALL_BATCH");   // 0: the host-mapped path for single replicas only (A/B)
    if (!small_off && (n_replicas == 1 || !small_batch_off) && k->d_slots == nullptr && k->unique_particles &&
        (long long) n_replicas * n_particles <= 4096 && k->n_atoms > 0)
        return execute_host_small(k, n_replicas, n_particles, pos, energies, grid_energies, forces, force_mode);

    const bool f32 = force_mode == GFB_FORCE_F32_STORE;
    const size_t fsz = f32 ? sizeof(float) : sizeof(double);      // bytes per force component
    const size_t np = (size_t) n_replicas * n_particles;
    const size_t pos_bytes = np * 3 * sizeof(double);
    const size_t f_bytes = np * 3 * fsz;
    const int ng = k->n_grids;
    const size_t n_e = (size_t) n_replicas * k->n_slots;          // energy entries: [replica][slot]
    const size_t e_count = n_e * (1 + ng);
    const bool pos_pinned = is_pinned(pos);
    const bool f_pinned = forces && is_pinned(forces);
    // Forces straight into host memory: when every particle is an evaluated atom (no indirection), the mode is a STORE
    // and a record kernel runs, its warps emit their forces as full contiguous 768-byte (F64) or 384-byte (F32) runs of
    // 16-byte stores, which the GPU writes over PCIe at copy-engine speed (51 GB/s measured) — so the D2H copies and the
    // device force buffer drop out of the pipeline, and the downloads overlap the uploads inside the kernels themselves.
    // C5 (74 MB each way in F64): 2.11 -> 1.86 ms per step; the two directions together top out at ~80 GB/s on this box.
    static const bool zc_off = env_off("GFB_ZEROCOPY_FORCES");   // 0: always download with the copy engine (A/B measurements)
    bool zc_forces = false;
    char* zc_dst = nullptr;                                       // device-visible address of the caller's pinned force buffer
    if (forces && (force_mode == GFB_FORCE_F64_STORE || f32) && !zc_off && !k->d_particles && k->n_atoms == n_particles) {
        EvalParams probe;
        memset(&probe, 0, sizeof probe);
        zc_forces = lines_eligible(k, probe) || (lines_f64_eligible(k) && !f32);
        if (zc_forces && f_pinned) {
            zc_dst = static_cast<char*>(host_device_pointer(forces));
            if (!zc_dst) zc_forces = false;
        }
    }
    if ((rc = k->d_pos.ensure(pos_bytes)) != GFB_OK) return rc;
    if (forces && !zc_forces && (rc = k->d_forces.ensure(f_bytes)) != GFB_OK) return rc;
    if ((rc = k->d_energy.ensure(e_count * sizeof(double))) != GFB_OK) return rc;
    if ((rc = k->h_energy.ensure(e_count * sizeof(double))) != GFB_OK) return rc;
    double* d_ae = nullptr;
    k->atom_e_count = 0;
    if (k->want_atom_energies) {
        if ((rc = k->d_atom_e.ensure((size_t) n_replicas * k->n_atoms * sizeof(double))) != GFB_OK) return rc;
        d_ae = static_cast<double*>(k->d_atom_e.ptr);
    }

    // staging: [pos | forces] when the user's buffers are pageable
    const size_t stage_bytes = (pos_pinned ? 0 : pos_bytes) + ((forces && !f_pinned) ? f_bytes : 0);
    if (stage_bytes && (rc = k->h_stage.ensure(stage_bytes)) != GFB_OK) return rc;
    char* stage_pos = static_cast<char*>(k->h_stage.ptr);
    char* stage_f = stage_pos + (pos_pinned ? 0 : pos_bytes);

    double* d_pos = static_cast<double*>(k->d_pos.ptr);
    // where the kernels write forces (byte address): the device buffer, or (zero-copy) the pinned host destination itself
    char* d_f = nullptr;
    if (forces) d_f = !zc_forces ? static_cast<char*>(k->d_forces.ptr) : (f_pinned ? zc_dst : stage_f);
    const char* forces_b = static_cast<const char*>(forces);
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
    const size_t unit_elems = by_atoms ? 3 : (size_t) n_particles * 3;      // position/force components per unit
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
        } else if (n_units >= 4 * n_chunks) {   // replica boundaries at multiples of 4: 16-byte aligned for any particle count
            r0 &= ~3;                           // (F64 needs even replicas, F32 forces multiples of 4)
            if (c + 1 < n_chunks) r1 &= ~3;
        }
        const size_t off = (size_t) r0 * unit_elems;                 // components
        const size_t cnt = (size_t) (r1 - r0) * unit_elems;
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
            const char* fsrc = forces_b + off * fsz;
            if (!f_pinned) {
                memcpy(stage_f + off * fsz, fsrc, cnt * fsz);
                fsrc = stage_f + off * fsz;
            }
            err = cudaMemcpyAsync(d_f + off * fsz, fsrc, cnt * fsz, cudaMemcpyHostToDevice, up_stream);
        }
        if (n_chunks > 1) {
            if (err == cudaSuccess) err = cudaEventRecord(up[c], dev->h2d_stream);
            if (err == cudaSuccess) err = cudaStreamWaitEvent(dev->stream, up[c], 0);
        }
        if (err != cudaSuccess) {
            status = fail(GFB_ERR_CUDA, "gfb_kernel_execute_host: H2D: %s", cudaGetErrorString(err));
            break;
        }
        EvalExtra x;
        if (by_atoms) {   // every range accumulates into the one replica's energy entries
            x.atom_begin = r0;
            x.atom_count = r1 - r0;
            x.atom_energies = d_ae ? d_ae + r0 : nullptr;
            status = enqueue_eval(k, 1, r1 - r0, d_pos + off, d_e, grid_energies ? d_ge : nullptr, d_f ? d_f + off * fsz : nullptr,
                                  force_mode, 0, nullptr, nullptr, dev->stream, x);
        } else {
            x.atom_energies = d_ae ? d_ae + (size_t) r0 * k->n_atoms : nullptr;
            status = enqueue_eval(k, r1 - r0, n_particles, d_pos + off, d_e + (size_t) r0 * k->n_slots,
                                  grid_energies ? d_ge + (size_t) r0 * k->n_slots * ng : nullptr,
                                  d_f ? d_f + off * fsz : nullptr, force_mode, 0, nullptr, nullptr, dev->stream, x);
        }
        if (status != GFB_OK) break;
        err = cudaSuccess;
        if (forces && !zc_forces) {
            if (n_chunks > 1) {
                err = cudaEventRecord(done[c], dev->stream);
                if (err == cudaSuccess) err = cudaStreamWaitEvent(dev->copy_stream, done[c], 0);
            }
            char* dst = (f_pinned ? static_cast<char*>(forces) : stage_f) + off * fsz;
            if (err == cudaSuccess) err = cudaMemcpyAsync(dst, d_f + off * fsz, cnt * fsz, cudaMemcpyDeviceToHost, down_stream);
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
    if (forces && !f_pinned) memcpy(forces, stage_f, f_bytes);
    if (d_ae) k->atom_e_count = (long long) n_replicas * k->n_atoms;
    return GFB_OK;
}

}  // extern "C"
