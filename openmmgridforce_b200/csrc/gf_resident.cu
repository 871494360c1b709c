// Resident evaluator: one ligand per MD step without a kernel launch or a stream synchronisation per step
// (BASELINE configs[1]: a single ~50-atom ligand in receptor grids, MD ns/day; what B200CalcGridForceKernel::execute issues
// once per step, replacing the call sequence of ReferenceCalcGridForceKernel::execute,
// platforms/reference/src/ReferenceGridForceKernels.cpp:646-1121).
//
// Such a step is pure latency. execute_host_small (gf_capi.cu) already works on host-mapped memory with ONE launch and ONE
// synchronise per step, and that launch + synchronise is what is left (~15 us). Here a single block stays resident on
// the GPU and is driven through a page-locked control block that both sides address directly:
//
//   host:   positions -> ctl.pos, then cmd = step number | what is wanted (release)     spins on done_seq == n
//   block:  thread 0 polls cmd over PCIe (ld.relaxed.sys, one fence when it changes); the block brings the positions in
//           as contiguous 16-byte loads through shared memory, every thread evaluates its atom on all grids (the same
//           device functions as gf_eval_kernel: classify, load_stencil, accumulate_inside, accumulate_restraint), the
//           forces go out as contiguous 16-byte stores, the block sums the energies in a fixed order, and thread 0
//           publishes done_seq = n with one st.release.sys.
//
// A step is then three PCIe trips (poll sees the command, positions come in, results go out) plus ~2 us of evaluation.
// The block never holds the GPU: after `idle` microseconds without a command it clears ctl.alive and exits, and the next
// step launches it again (so cudaFree / cudaDeviceSynchronize elsewhere in the process wait for at most that long);
// every host-side wait is bounded and fails with an error instead of hanging.
//
// Opt-in (gfb_kernel_set_resident; platform property "ResidentKernel"): one replica, at most 256 evaluated atoms, no
// energy slots, trilinear packed cells (per-grid or interleaved records), either precision, inv-power included.
#include <time.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "gf_handles.h"
#include "gf_kernels.cuh"

using namespace gfb;

namespace {

constexpr unsigned long long kStop = ~0ull;
constexpr int kMaxResidentAtoms = 256;    // one block; 255 registers per thread stay available (a ligand has ~50 atoms)

// Page-locked, host- and device-addressable. Lines written by the host and lines written by the device are kept apart.
struct ResidentCtl {
    volatile unsigned long long cmd;        // host -> device: (step number << 2) | want bits (1 forces, 2 per-grid energies);
                                            // kStop = exit now
    unsigned char pad0[128 - 8];
    volatile unsigned long long done_seq;   // device -> host: last step whose results are complete
    volatile unsigned long long alive;      // device -> host: cleared by the block right before it exits
    unsigned char pad1[128 - 16];
    double energies[1 + GFB_MAX_GRIDS];     // device -> host: total, then per grid
    unsigned long long stamps[4];           // device -> host: %globaltimer (ns) of the last step: command seen, positions in,
                                            // evaluated, results stored (gfb_kernel_resident_timeline)
    unsigned char pad2[128 - (1 + GFB_MAX_GRIDS) * 8 - 32];
    // followed by: double pos[3 * n_particles] (host -> device), padded to 128 bytes, double forces[3 * n_particles]
};
static_assert(sizeof(ResidentCtl) == 384, "three 128-byte lines");

struct ResidentParams {
    GridView grid[GFB_MAX_GRIDS];
    int n_grids, n_atoms, same_geom, n_particles;
    const int* particles;              // [n_atoms] or null
    ResidentCtl* ctl;
    const double* pos;                 // ctl.pos
    double* forces;                    // ctl.forces
    unsigned long long start_seq;      // last step already done when this launch starts
    unsigned long long idle_ns;
};

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const volatile unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void st_release_sys(volatile unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_sys_f64(const double* p) {   // host memory the host has just written: never from a cache
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_sys_f64x2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f64(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_sys_f64x2(double* p, double2 v) {
    asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename S>
__global__ void __launch_bounds__(kMaxResidentAtoms, 1) gf_resident_kernel(const __grid_constant__ ResidentParams p) {
    constexpr bool EXACT = sizeof(S) == 8;
    __shared__ unsigned long long s_cmd;
    __shared__ unsigned long long s_t[3];
    __shared__ double s_e[kMaxResidentAtoms / 32][1 + GFB_MAX_GRIDS];
    __shared__ __align__(16) double s_xyz[3 * kMaxResidentAtoms + 2];    // positions in, then forces out (plain states)
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned n_warps = (blockDim.x + 31u) >> 5;
    const bool active = tid < (unsigned) p.n_atoms;
    const int particle = active ? (p.particles ? p.particles[tid] : (int) tid) : 0;
    // Plain state (every particle is an evaluated atom, in order): the 24*n bytes of positions come in, and the forces go
    // out, as contiguous 16-byte accesses (full PCIe payloads) through shared memory instead of stride-24 doubles.
    const bool plain = p.particles == nullptr && p.n_atoms == p.n_particles;
    const unsigned n_pairs = (3u * (unsigned) p.n_atoms + 1u) >> 1;       // 16-byte units; the arrays are padded to 128 bytes
    ResidentCtl* ctl = p.ctl;
    unsigned long long expected = p.start_seq + 1;

    for (;;) {
        if (tid == 0) {
            const unsigned long long t0 = global_ns();
            unsigned long long c;
            for (;;) {
                c = ld_relaxed_sys(&ctl->cmd);
                if (c == kStop || (c >> 2) == expected) break;
                if (global_ns() - t0 > p.idle_ns) {
                    c = kStop;
                    break;
                }
            }
            fence_sys();             // acquire: the positions written before the command are what the loads below see
            s_cmd = c;
            s_t[0] = global_ns();
        }
        __syncthreads();
        const unsigned long long cmd = s_cmd;
        if (cmd == kStop) break;
        const unsigned long long want = cmd & 3ull;

        double e_total = 0.0, Fx = 0.0, Fy = 0.0, Fz = 0.0;
        double e_grid[GFB_MAX_GRIDS];
        double x = 0.0, y = 0.0, z = 0.0;
        if (plain) {
            for (unsigned i = tid; i < n_pairs; i += blockDim.x) reinterpret_cast<double2*>(s_xyz)[i] = ld_sys_f64x2(p.pos + 2 * i);
            __syncthreads();
            if (active) {
                x = s_xyz[3 * tid];
                y = s_xyz[3 * tid + 1];
                z = s_xyz[3 * tid + 2];
            }
            __syncthreads();         // s_xyz is reused for the forces
        } else if (active) {
            x = ld_sys_f64(p.pos + 3 * particle);
            y = ld_sys_f64(p.pos + 3 * particle + 1);
            z = ld_sys_f64(p.pos + 3 * particle + 2);
        }
        if (tid == 0) s_t[1] = global_ns();
        // Pass 1: classify and put every grid's stencil load in flight (the loads are independent round trips to L2/HBM);
        // pass 2: the arithmetic.
        AtomCell c[GFB_MAX_GRIDS];
        S v[GFB_MAX_GRIDS][8];
        double sd[GFB_MAX_GRIDS];
        bool interp[GFB_MAX_GRIDS];
#pragma unroll
        for (int g = 0; g < GFB_MAX_GRIDS; g++) {
            interp[g] = false;
            if (g < p.n_grids && active) {
                const GridView& G = p.grid[g];
                c[g] = (p.same_geom && g > 0) ? c[0] : classify<EXACT>(G, x, y, z);
                sd[g] = G.scaling[tid];
                interp[g] = c[g].inside && sd[g] != 0.0;     // :706
                if (interp[g]) load_stencil<S, GFB_LAYOUT_CELLS>(G, c[g].ix, c[g].iy, c[g].iz, v[g]);
            }
        }
#pragma unroll
        for (int g = 0; g < GFB_MAX_GRIDS; g++) {
            e_grid[g] = 0.0;
            if (g < p.n_grids && active) {
                const GridView& G = p.grid[g];
                double e_g = 0.0;
                if (interp[g]) accumulate_inside<S>(G, v[g], c[g], sd[g], e_g, Fx, Fy, Fz);
                else accumulate_restraint(G, x, y, z, e_g, Fx, Fy, Fz);      // :1093-1117
                e_grid[g] = e_g;
                e_total += e_g;
            }
        }
        if (tid == 0) s_t[2] = global_ns();
        if (want & 1ull) {
            if (plain) {
                if (active) {
                    s_xyz[3 * tid] = Fx;
                    s_xyz[3 * tid + 1] = Fy;
                    s_xyz[3 * tid + 2] = Fz;
                }
                __syncthreads();
                for (unsigned i = tid; i < n_pairs; i += blockDim.x) st_sys_f64x2(p.forces + 2 * i, reinterpret_cast<double2*>(s_xyz)[i]);
            } else if (active) {
                st_sys_f64(p.forces + 3 * particle, Fx);
                st_sys_f64(p.forces + 3 * particle + 1, Fy);
                st_sys_f64(p.forces + 3 * particle + 2, Fz);
            }
        }
        // energies: lanes -> warp -> block, always in the same order
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e_total += __shfl_xor_sync(kFull, e_total, o);
        if (lane == 0) s_e[warp][0] = e_total;
        if (want & 2ull) {
#pragma unroll
            for (int g = 0; g < GFB_MAX_GRIDS; g++) {
                if (g < p.n_grids) {
                    double eg = e_grid[g];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) eg += __shfl_xor_sync(kFull, eg, o);
                    if (lane == 0) s_e[warp][1 + g] = eg;
                }
            }
        }
        __syncthreads();             // every thread's force stores are issued and the warp sums are in shared memory
        if (tid == 0) {
            const int n_out = (want & 2ull) ? 1 + p.n_grids : 1;
            for (int j = 0; j < n_out; j++) {
                double b = 0.0;
                for (unsigned w = 0; w < n_warps; w++) b += s_e[w][j];
                st_sys_f64(&ctl->energies[j], b);
            }
            st_sys_u64(&ctl->stamps[0], s_t[0]);
            st_sys_u64(&ctl->stamps[1], s_t[1]);
            st_sys_u64(&ctl->stamps[2], s_t[2]);
            st_sys_u64(&ctl->stamps[3], global_ns());
            // One release at system scope by one thread: the barrier above orders the other threads' stores before it
            // (cumulativity), so the host that reads done_seq == n reads this step's forces and energies.
            st_release_sys(&ctl->done_seq, expected);
        }
        expected++;
    }
    if (tid == 0) st_release_sys(&ctl->alive, 0ull);
}

double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

}  // namespace

namespace gfb {

struct ResidentState {
    ResidentCtl* ctl = nullptr;      // cudaHostAlloc
    size_t ctl_bytes = 0;
    int n_particles = 0;
    double* pos = nullptr;           // inside ctl
    double* forces = nullptr;
    cudaStream_t stream = nullptr;   // non-blocking: the block must not serialise with the device's other streams
    unsigned long long seq = 0;      // last step requested
    unsigned long long idle_us = 100000;
    unsigned long long launches = 0;
    bool running = false;            // a launch has been made whose exit has not been observed yet
};

static bool resident_supported(const gfb_kernel* k) {
    static const bool blocking = [] {
        const char* e = getenv("CUDA_LAUNCH_BLOCKING");
        return e && e[0] && e[0] != '0';
    }();
    if (blocking || k->n_atoms < 1 || k->n_atoms > kMaxResidentAtoms || k->d_slots || !k->unique_particles) return false;
    for (int g = 0; g < k->n_grids; g++) {
        if (k->grids[g]->layout != GFB_LAYOUT_CELLS) return false;
        if (!k->grids[g]->cells && !k->d_interleaved) return false;
    }
    return true;
}

// Waits for the block of an earlier launch to be gone (it exits on its own: stop command or idle time-out).
static int resident_join(gfb_kernel* k, ResidentState* r) {
    if (!r->running) return GFB_OK;
    cudaError_t e = cudaStreamSynchronize(r->stream);
    r->running = false;
    if (e != cudaSuccess) return fail(GFB_ERR_CUDA, "resident evaluator: %s", cudaGetErrorString(e));
    return GFB_OK;
}

static int resident_launch(gfb_kernel* k, ResidentState* r) {
    int rc = resident_join(k, r);
    if (rc != GFB_OK) return rc;
    ResidentParams p;
    memset(&p, 0, sizeof p);
    for (int g = 0; g < k->n_grids; g++) {
        fill_grid_view(k, g, p.grid[g]);
        if (!k->grids[g]->cells) {   // gfb_grid_release_cells: the corners live in slot g of the interleaved records
            const size_t slot_bytes = k->precision == GFB_PRECISION_MIXED ? 32 : 64;
            p.grid[g].cells = static_cast<const char*>(k->d_interleaved) + (size_t) g * slot_bytes;
            p.grid[g].cell_stride = 8 * k->il_slots;
        }
    }
    p.n_grids = k->n_grids;
    p.n_atoms = k->n_atoms;
    p.n_particles = r->n_particles;
    p.same_geom = k->same_geom ? 1 : 0;
    p.particles = k->d_particles;
    p.ctl = r->ctl;
    p.pos = r->pos;
    p.forces = r->forces;
    p.start_seq = r->seq - 1;        // the step just requested is the first one this launch runs
    p.idle_ns = r->idle_us * 1000ull;
    r->ctl->alive = 1;
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    const unsigned threads = (unsigned) ((k->n_atoms + 31) / 32 * 32);
    if (k->precision == GFB_PRECISION_DOUBLE) gf_resident_kernel<double><<<1, threads, 0, r->stream>>>(p);
    else gf_resident_kernel<float><<<1, threads, 0, r->stream>>>(p);
    g_launches++;
    r->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        r->ctl->alive = 0;
        return fail(GFB_ERR_CUDA, "resident evaluator launch: %s", cudaGetErrorString(e));
    }
    r->running = true;
    return GFB_OK;
}

int resident_stop(gfb_kernel* k) {
    ResidentState* r = static_cast<ResidentState*>(k->resident);
    if (!r) return GFB_OK;
    int rc = GFB_OK;
    if (r->running) {
        __atomic_store_n(&r->ctl->cmd, kStop, __ATOMIC_RELEASE);
        rc = resident_join(k, r);
    }
    return rc;
}

void resident_destroy(gfb_kernel* k) {
    ResidentState* r = static_cast<ResidentState*>(k->resident);
    if (!r) return;
    resident_stop(k);
    if (r->stream) cudaStreamDestroy(r->stream);
    if (r->ctl) cudaFreeHost(r->ctl);
    delete r;
    k->resident = nullptr;
}

bool resident_enabled(const gfb_kernel* k) { return k->resident != nullptr && resident_supported(k); }

// One step. Same contract as execute_host_small: forces (if not NULL) are F64, STORE (evaluated entries overwritten, the
// others left alone) or ADD.
int resident_step(gfb_kernel* k, int n_particles, const double* pos, double* energies, double* grid_energies, double* forces,
                  bool add) {
    ResidentState* r = static_cast<ResidentState*>(k->resident);
    const size_t np3 = (size_t) n_particles * 3;
    const size_t arr_bytes = (np3 * sizeof(double) + 127) & ~(size_t) 127;
    if (!r->ctl || r->n_particles != n_particles) {
        int rc = resident_stop(k);
        if (rc != GFB_OK) return rc;
        if (r->ctl) cudaFreeHost(r->ctl);
        r->ctl = nullptr;
        r->ctl_bytes = sizeof(ResidentCtl) + 2 * arr_bytes;
        void* mem = nullptr;
        CUDA_TRY(cudaHostAlloc(&mem, r->ctl_bytes, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(mem, 0, r->ctl_bytes);
        r->ctl = static_cast<ResidentCtl*>(mem);
        r->pos = reinterpret_cast<double*>(static_cast<char*>(mem) + sizeof(ResidentCtl));
        r->forces = reinterpret_cast<double*>(static_cast<char*>(mem) + sizeof(ResidentCtl) + arr_bytes);
        r->n_particles = n_particles;
        r->seq = 0;
    }
    if (!r->stream) CUDA_TRY(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    ResidentCtl* ctl = r->ctl;
    memcpy(r->pos, pos, np3 * sizeof(double));
    if (np3 & 1) r->pos[np3] = 0.0;      // the block reads whole 16-byte units (the arrays are padded to 128 bytes)
    // The block stores the forces of the evaluated particles only. STORE: the other entries must come back as they were;
    // ADD: they must come back unchanged, i.e. the block's array contributes 0 there.
    const bool all_written = !k->d_particles && k->n_atoms == n_particles;
    if (forces && !all_written) {
        if (add) memset(r->forces, 0, np3 * sizeof(double));
        else memcpy(r->forces, forces, np3 * sizeof(double));
    }
    const unsigned long long n = ++r->seq;
    __atomic_store_n(&ctl->cmd, (n << 2) | (forces ? 1ull : 0ull) | (grid_energies ? 2ull : 0ull), __ATOMIC_RELEASE);
    if (!r->running || !__atomic_load_n(&ctl->alive, __ATOMIC_ACQUIRE)) {
        int rc = resident_launch(k, r);     // joins a block that has timed out, then starts one at this step
        if (rc != GFB_OK) return rc;
    }
    // The answer arrives within microseconds; the wall-clock checks only bound a failure.
    double t0 = 0.0;
    for (unsigned long long spins = 1;; spins++) {
        if (__atomic_load_n(&ctl->done_seq, __ATOMIC_ACQUIRE) == n) break;
        if (!__atomic_load_n(&ctl->alive, __ATOMIC_ACQUIRE)) {
            // The block gave up waiting in the instant the command was posted. done_seq is final once alive is clear.
            if (__atomic_load_n(&ctl->done_seq, __ATOMIC_ACQUIRE) == n) break;
            int rc = resident_launch(k, r);
            if (rc != GFB_OK) return rc;
            t0 = 0.0;
            continue;
        }
        if ((spins & 0x3ffu) == 0) {
            const double t = now_s();
            if (t0 == 0.0) t0 = t;
            if (t - t0 > 5.0) {
                const cudaError_t q = cudaStreamQuery(r->stream);
                __atomic_store_n(&ctl->cmd, kStop, __ATOMIC_RELEASE);
                return fail(GFB_ERR_CUDA, "resident evaluator did not answer step %llu within 5 s (%s)", n,
                            q == cudaErrorNotReady ? "block still running" : cudaGetErrorString(q));
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    if (energies) energies[0] = ctl->energies[0];
    if (grid_energies) memcpy(grid_energies, const_cast<double*>(ctl->energies) + 1, (size_t) k->n_grids * sizeof(double));
    if (forces) {
        if (add) for (size_t i = 0; i < np3; i++) forces[i] += r->forces[i];
        else memcpy(forces, r->forces, np3 * sizeof(double));
    }
    return GFB_OK;
}

}  // namespace gfb

extern "C" {

int gfb_kernel_set_resident(gfb_kernel* k, int enable, long long idle_us) {
    if (!k) return fail(GFB_ERR_INVALID, "gfb_kernel_set_resident: NULL kernel");
    CUDA_TRY(cudaSetDevice(k->dev->ordinal));
    if (!enable) {
        resident_destroy(k);
        return GFB_OK;
    }
    if (!resident_supported(k))
        return fail(GFB_ERR_UNSUPPORTED, "gfb_kernel_set_resident: needs trilinear packed cells, 1..%d evaluated atoms, distinct particles, "
                                         "no energy slots, and CUDA_LAUNCH_BLOCKING unset", kMaxResidentAtoms);
    if (idle_us <= 0) idle_us = 100000;
    ResidentState* r = static_cast<ResidentState*>(k->resident);
    if (!r) {
        r = new (std::nothrow) ResidentState();
        if (!r) return fail(GFB_ERR_NOMEM, "gfb_kernel_set_resident: out of host memory");
        k->resident = r;
    } else if (r->idle_us != (unsigned long long) idle_us) {
        int rc = resident_stop(k);       // the time-out is a launch argument
        if (rc != GFB_OK) return rc;
    }
    r->idle_us = (unsigned long long) idle_us;
    return GFB_OK;
}

int gfb_kernel_resident_stop(gfb_kernel* k) {
    if (!k) return fail(GFB_ERR_INVALID, "gfb_kernel_resident_stop: NULL kernel");
    CUDA_TRY(cudaSetDevice(k->dev->ordinal));
    return resident_stop(k);
}

int gfb_kernel_resident_timeline(const gfb_kernel* k, double us[3]) {
    if (!k || !us) return fail(GFB_ERR_INVALID, "gfb_kernel_resident_timeline: NULL argument");
    const ResidentState* r = static_cast<const ResidentState*>(k->resident);
    if (!r || !r->ctl || r->seq == 0) return fail(GFB_ERR_INVALID, "gfb_kernel_resident_timeline: no resident step has run");
    const unsigned long long* t = r->ctl->stamps;
    for (int i = 0; i < 3; i++) us[i] = 1e-3 * (double) (long long) (t[i + 1] - t[i]);
    return GFB_OK;
}

long long gfb_kernel_resident_launches(const gfb_kernel* k) {
    if (!k || !k->resident) return 0;
    return (long long) static_cast<const ResidentState*>(k->resident)->launches;
}

}  // extern "C"
