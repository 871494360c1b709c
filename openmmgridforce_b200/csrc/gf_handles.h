// Host-side state behind the opaque handles of include/gridforce_b200.h, shared by the translation units of
// libgridforce_b200.so (gf_capi.cu: devices, kernel state, execute paths; gf_grids.cu: grid ingest/generation;
// gf_aux.cu: classification, atom sort, probes; gf_multi.cu: multi-GPU). Host code only.
#ifndef GF_HANDLES_H_
#define GF_HANDLES_H_

#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "gf_params.h"
#include "gridforce_b200.h"

namespace gfb {

int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));   // sets gfb_last_error(), returns code
extern std::atomic<unsigned long long> g_launches;                                // gfb_launch_count()

#define CUDA_TRY(expr)                                                                                           \
    do {                                                                                                         \
        cudaError_t err__ = (expr);                                                                              \
        if (err__ != cudaSuccess)                                                                                \
            return gfb::fail(GFB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

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

}  // namespace gfb

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
    void* cells;       // nullptr after gfb_grid_release_cells
    size_t bytes;
    size_t released_bytes;   // size of the copy gfb_grid_release_cells freed (the geometry-derived cell count stays valid)
};

struct gfb_kernel {
    gfb_device* dev;
    int n_grids, n_atoms, precision;
    bool same_geom;
    gfb_grid* grids[GFB_MAX_GRIDS];
    double inv_power[GFB_MAX_GRIDS], oob_k[GFB_MAX_GRIDS];
    void* d_scaling;      // [n_grids][n_atoms] double
    int* d_particles;     // [n_atoms] or null
    int max_particle;     // largest particle index referenced (+1 = minimum n_particles)
    int* d_slots;         // [n_atoms] energy slot per atom (particle groups) or null
    int n_slots;          // energy slots per replica
    void* d_interleaved;  // CELLS + shared geometry + 2..4 grids: one record per cell holding every grid's corners
                          // (MIXED: 4 slots x 32 B = 128 B; DOUBLE: 4 slots x 64 B = 256 B)
    int il_slots;         // grids per record incl. padding (4); 0 = not interleaved
    size_t n_cells;       // cells per grid (CELLS layout), fixed at creation
    // host-path scratch
    gfb::DeviceBuffer d_pos, d_forces, d_energy, d_cls, d_sort, d_atom_e;
    gfb::PinnedBuffer h_stage, h_energy;
    // one-ligand-per-step path (execute_host_small): host-mapped staging the kernel reads and writes over PCIe
    gfb::PinnedBuffer h_small;
    bool launch_overlap;     // gfb_kernel_set_launch_overlap: device-path launches may overlap the previous launch's tail
    bool unique_particles;   // no particle index occurs twice: a plain store per evaluated particle is the whole force
    bool want_atom_energies; // gfb_kernel_request_atom_energies
    long long atom_e_count;  // entries of d_atom_e written by the last host-path call
    // execute_host_small's two alternating device accumulator arrays (d_small_e): entries per array, which one is next,
    // and how many leading entries of the next one are known to be zero
    gfb::DeviceBuffer d_small_e;
    double* small_base = nullptr;
    size_t small_stride = 0, small_zeroed = 0;
    int small_toggle = 0;
    void* resident = nullptr;   // gfb::ResidentState (gf_resident.cu) once gfb_kernel_set_resident(enable) has been called
};

namespace gfb {
void fill_grid_view(const gfb_kernel* k, int g, GridView& v);
bool lines_eligible(const gfb_kernel* k, const EvalParams& p);
bool lines_f64_eligible(const gfb_kernel* k);
int check_exec_args(const char* fn, gfb_kernel* k, int n_replicas, int n_particles, const void* pos, int force_mode);

// What a launch may carry beyond the arguments of gfb_kernel_execute_device.
struct EvalExtra {
    bool energy_store = false;         // single block: energies are plain stores (host-mapped destination)
    int atom_begin = 0, atom_count = -1;   // atom range of one large replica (host path chunks)
    bool overlap = false;              // programmatic dependent launch
    double* atom_energies = nullptr;   // [n_replicas][n_atoms] per-atom energies, stored
    GatherTable* gather = nullptr;     // fused energy gather (gf_multi.cu)
    long long gather_offset = 0;
};
// Enqueues ONE evaluation kernel on `stream` (gf_capi.cu); no synchronisation.
int enqueue_eval(gfb_kernel* k, int n_replicas, int n_particles, const double* d_pos, double* d_energies,
                 double* d_grid_energies, void* d_forces, int force_mode, long long force_stride, const int* d_order,
                 double* d_energies_clear, cudaStream_t stream, const EvalExtra& x = EvalExtra());

// Resident evaluator (gf_resident.cu): one ligand per MD step served by a block that stays on the GPU.
bool resident_enabled(const gfb_kernel* k);     // switched on and this state qualifies
int resident_step(gfb_kernel* k, int n_particles, const double* pos, double* energies, double* grid_energies, double* forces, bool add);
int resident_stop(gfb_kernel* k);               // the block exits (parameters it caches are about to change); relaunched on demand
void resident_destroy(gfb_kernel* k);
}  // namespace gfb

#endif
