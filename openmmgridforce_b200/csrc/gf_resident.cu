// Resident evaluator: one ligand per MD step without a kernel launch or a stream synchronisation per step
// (BASELINE configs[1]: a single ~50-atom ligand in receptor grids, MD ns/day; what B200CalcGridForceKernel::execute issues
// once per step, replacing the call sequence of ReferenceCalcGridForceKernel::execute,
// platforms/reference/src/ReferenceGridForceKernels.cpp:646-1121).
//
// Such a step is pure latency. execute_host_small (gf_capi.cu) already works on host-mapped memory with ONE launch and ONE
// synchronise per step, and that launch + synchronise is what is left (~15 us). Here a single block stays resident on
// the GPU and talks to the host through page-locked memory both sides address directly, with the DATA AS ITS OWN FLAG
// (the idea of gf_gather_ll_kernel, applied to PCIe): every double travels as a 16-byte packet
//     { low 32 bits | tag << 32 ,  high 32 bits | tag << 32 },      tag = (step number << 2 | what is wanted) mod 2^32
// so a packet whose two 8-byte halves both carry this step's tag is complete, whatever order the halves became visible
// in. Nothing else is needed: no command word, no fence, no second round trip.
//
//   host:   writes the 3 x n_atoms position packets of step n                          then reads the result packets
//   block:  its threads spin on the position packets (ld.relaxed.sys, 16 bytes each, contiguous across a warp, each
//           thread's up to three reads in flight at once: the poll IS the position read), every thread evaluates its
//           atom on all grids (the same device functions as gf_eval_kernel: classify,
//           load_stencil, accumulate_inside, accumulate_restraint), forces and energies go back as packets
//           (st.relaxed.sys, 16 bytes each), energies summed by the block in a fixed order.
//
// A step is then one PCIe read that sees the positions, ~2.5 us of evaluation, and posted writes back. One extra thread
// watches the control word (stop) and the clock: after `idle` microseconds without a step the block clears ctl.alive and
// exits, and the next step launches it again (so cudaFree / cudaDeviceSynchronize / the driver loading another module
// elsewhere in the process wait for at most that long); every host-side wait is bounded and fails with an error instead
// of hanging.
//
// Opt-in (gfb_kernel_set_resident; platform property "ResidentKernel"): one replica, at most 224 evaluated atoms, no
// energy slots, trilinear packed cells (per-grid or interleaved records), either precision, inv-power included.
#include <time.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "gf_handles.h"
#include "gf_kernels.cuh"

using namespace gfb;

namespace {

constexpr unsigned long long kStop = ~0ull;
constexpr int kMaxResidentAtoms = 224;    // + the watcher's warp = 256 threads: 255 registers per thread stay available (a ligand has ~50 atoms)

struct Packet {                           // 16 bytes, 16-byte aligned
    unsigned long long h[2];
};

// Page-locked, host- and device-addressable. Lines written by the host and lines written by the device are kept apart.
struct ResidentCtl {
    volatile unsigned long long cmd;        // host -> device: kStop = exit now (anything else: keep going)
    unsigned char pad0[128 - 8];
    volatile unsigned long long alive;      // device -> host: cleared by the block right before it exits
    unsigned long long stamps[3];           // device -> host: %globaltimer (ns) of the last step: positions seen, evaluated,
                                            // results stored (gfb_kernel_resident_timeline)
    unsigned char pad1[128 - 32];
    // followed by: Packet in[3 * n_atoms] (host -> device), padded to 128 bytes;
    //              Packet out[3 * n_atoms + 1 + GFB_MAX_GRIDS] (device -> host): forces, total energy, per-grid energies
};
static_assert(sizeof(ResidentCtl) == 256, "two 128-byte lines");

struct ResidentParams {
    GridView grid[GFB_MAX_GRIDS];
    int n_grids, n_atoms, same_geom, pad_;
    ResidentCtl* ctl;
    const Packet* in;
    Packet* out;
    unsigned long long start_seq;      // last step already done when this launch starts
    unsigned long long idle_ns;
};

__host__ __device__ inline Packet pack(unsigned long long bits, unsigned tag) {
    Packet p;
    p.h[0] = (bits & 0xffffffffull) | ((unsigned long long) tag << 32);
    p.h[1] = (bits >> 32) | ((unsigned long long) tag << 32);
    return p;
}
__host__ __device__ inline bool packet_has(const Packet& p, unsigned tag) {
    return (unsigned) (p.h[0] >> 32) == tag && (unsigned) (p.h[1] >> 32) == tag;
}
__host__ __device__ inline unsigned long long packet_bits(const Packet& p) { return (p.h[0] & 0xffffffffull) | (p.h[1] << 32); }

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const volatile unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ Packet ld_packet_sys(const Packet* p) {   // host memory the host is writing: never from a cache
    Packet v;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.h[0]), "=l"(v.h[1]) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_packet_sys(Packet* p, const Packet& v) {
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(v.h[0]), "l"(v.h[1]) : "memory");
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(volatile unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// blockDim.x = n_atoms rounded up to a warp + one more warp, whose first lane is the watcher.
template <typename S>
__global__ void __launch_bounds__(kMaxResidentAtoms + 32, 1) gf_resident_kernel(const __grid_constant__ ResidentParams p) {
    constexpr bool EXACT = sizeof(S) == 8;
    __shared__ volatile int s_stop;
    __shared__ unsigned s_want;
    __shared__ unsigned long long s_t[2];
    __shared__ double s_e[kMaxResidentAtoms / 32 + 1][1 + GFB_MAX_GRIDS];
    __shared__ double s_pos[3 * kMaxResidentAtoms];      // positions as they come in
    __shared__ double s_xyz[3 * kMaxResidentAtoms];      // forces on their way out
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned n_workers = blockDim.x - 32u;          // threads that fetch, evaluate and send
    const unsigned n_warps = n_workers >> 5;
    const bool worker = tid < n_workers;
    const bool watcher = tid == n_workers;
    const bool active = tid < (unsigned) p.n_atoms;
    const unsigned n_in = 3u * (unsigned) p.n_atoms;
    ResidentCtl* ctl = p.ctl;
    unsigned long long expected = p.start_seq + 1;
    // Scaling factors are fixed while the block is up (gfb_kernel_update_parameters stops it): loaded once.
    double sd[GFB_MAX_GRIDS];
#pragma unroll
    for (int g = 0; g < GFB_MAX_GRIDS; g++) sd[g] = (g < p.n_grids && active) ? p.grid[g].scaling[tid] : 0.0;
    if (tid == 0) s_stop = 0;
    __syncthreads();

    for (;;) {
        // ---- wait for step `expected`: its position packets are the signal ---------------------------------------------
        // Worker t polls packets t, t + W, t + 2W (W workers >= atoms, so that covers all 3A): the three reads are in
        // flight at once and every warp instruction covers 512 contiguous bytes — the whole set is seen one PCIe read
        // after the host wrote it, in as few read requests as its 48·A bytes allow. (One thread polling the three
        // packets of its own atom — stride 48 — was measured slower: 11.5 against 10.0 us per step with the loads of a
        // thread issued one after the other; many small PCIe reads cost more than a round trip.)
        const unsigned tag_hi = (unsigned) (expected << 2);            // the low two bits say what is wanted
        if (worker) {
            bool have[3];
#pragma unroll
            for (int j = 0; j < 3; j++) have[j] = tid + j * n_workers >= n_in;
            const unsigned long long t0 = global_ns();
            for (;;) {
                Packet q[3];
#pragma unroll
                for (int j = 0; j < 3; j++)
                    if (!have[j]) q[j] = ld_packet_sys(p.in + tid + j * n_workers);
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    if (have[j]) continue;
                    const unsigned long long a = q[j].h[0] >> 32;
                    if (((unsigned) a & ~3u) == tag_hi && a == (q[j].h[1] >> 32)) {
                        s_pos[tid + j * n_workers] = __longlong_as_double((long long) packet_bits(q[j]));
                        if (j == 0 && tid == 0) s_want = (unsigned) a & 3u;
                        have[j] = true;
                    }
                }
                if (have[0] && have[1] && have[2]) break;
                if (s_stop) break;
                if (global_ns() - t0 > p.idle_ns) {      // also bounds a step whose packets stop coming half-way
                    s_stop = 1;
                    break;
                }
            }
        } else if (watcher) {
            const unsigned long long t0 = global_ns();
            for (;;) {
                const unsigned long long c = ld_relaxed_sys(&ctl->cmd);
                const Packet q = ld_packet_sys(p.in);
                if (c == kStop || global_ns() - t0 > p.idle_ns) {
                    s_stop = 1;
                    break;
                }
                if (((unsigned) (q.h[0] >> 32) & ~3u) == tag_hi && (q.h[0] >> 32) == (q.h[1] >> 32)) break;
            }
        }
        __syncthreads();
        if (s_stop) break;       // (a step whose packets arrived in the same instant is run by the next launch)
        const unsigned want = s_want;
        const unsigned tag = tag_hi | want;
        if (tid == 0) s_t[0] = global_ns();

        double e_total = 0.0, Fx = 0.0, Fy = 0.0, Fz = 0.0;
        double e_grid[GFB_MAX_GRIDS];
        double x = 0.0, y = 0.0, z = 0.0;
        if (active) {
            x = s_pos[3 * tid];
            y = s_pos[3 * tid + 1];
            z = s_pos[3 * tid + 2];
        }
        // Pass 1: classify and put every grid's stencil load in flight (the loads are independent round trips to L2/HBM);
        // pass 2: the arithmetic.
        AtomCell c[GFB_MAX_GRIDS];
        S v[GFB_MAX_GRIDS][8];
        bool interp[GFB_MAX_GRIDS];
#pragma unroll
        for (int g = 0; g < GFB_MAX_GRIDS; g++) {
            interp[g] = false;
            if (g < p.n_grids && active) {
                const GridView& G = p.grid[g];
                c[g] = (p.same_geom && g > 0) ? c[0] : classify<EXACT>(G, x, y, z);
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
        if (tid == 0) s_t[1] = global_ns();
        // ---- results: forces as packets (contiguous 16-byte stores), energies summed lanes -> warps -> block ----------
        if (active && (want & 1u)) {
            s_xyz[3 * tid] = Fx;
            s_xyz[3 * tid + 1] = Fy;
            s_xyz[3 * tid + 2] = Fz;
        }
        if (worker) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e_total += __shfl_xor_sync(kFull, e_total, o);
            if (lane == 0) s_e[warp][0] = e_total;
            if (want & 2u) {
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
        }
        __syncthreads();
        if (worker && (want & 1u))
            for (unsigned i = tid; i < n_in; i += n_workers)
                st_packet_sys(p.out + i, pack((unsigned long long) __double_as_longlong(s_xyz[i]), tag));
        if (watcher) {               // the spare thread sends the energies while the workers send the forces
            const int n_out = (want & 2u) ? 1 + p.n_grids : 1;
            for (int j = 0; j < n_out; j++) {
                double b = 0.0;
                for (unsigned w = 0; w < n_warps; w++) b += s_e[w][j];
                st_packet_sys(p.out + n_in + j, pack((unsigned long long) __double_as_longlong(b), tag));
            }
            st_sys_u64(&ctl->stamps[0], s_t[0]);
            st_sys_u64(&ctl->stamps[1], s_t[1]);
            st_sys_u64(&ctl->stamps[2], global_ns());
        }
        expected++;
        __syncthreads();             // s_xyz / s_e are free again
    }
    if (tid == 0) st_release_sys(&ctl->alive, 0ull);
}

double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

inline void store_packet(Packet* dst, const Packet& v) {   // two aligned 8-byte stores: each is atomic, the order is free
    __atomic_store_n(&dst->h[0], v.h[0], __ATOMIC_RELAXED);
    __atomic_store_n(&dst->h[1], v.h[1], __ATOMIC_RELAXED);
}
inline Packet load_packet(const Packet* src) {
    Packet v;
    v.h[0] = __atomic_load_n(&src->h[0], __ATOMIC_RELAXED);
    v.h[1] = __atomic_load_n(&src->h[1], __ATOMIC_RELAXED);
    return v;
}

}  // namespace

namespace gfb {

struct ResidentState {
    ResidentCtl* ctl = nullptr;      // cudaHostAlloc
    Packet* in = nullptr;            // inside the same allocation
    Packet* out = nullptr;
    int n_atoms = 0;
    std::vector<int> particles;      // host copy of the kernel's particle map (empty: atom a is particle a)
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
    p.same_geom = k->same_geom ? 1 : 0;
    p.ctl = r->ctl;
    p.in = r->in;
    p.out = r->out;
    p.start_seq = r->seq - 1;        // the step just requested is the first one this launch runs
    p.idle_ns = r->idle_us * 1000ull;
    r->ctl->cmd = 0;
    r->ctl->alive = 1;
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    const unsigned threads = (unsigned) ((k->n_atoms + 31) / 32 * 32) + 32u;
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
    const int na = k->n_atoms;
    const size_t n_in = 3 * (size_t) na;
    if (!r->ctl || r->n_atoms != na) {
        int rc = resident_stop(k);
        if (rc != GFB_OK) return rc;
        if (r->ctl) cudaFreeHost(r->ctl);
        r->ctl = nullptr;
        const size_t in_bytes = (n_in * sizeof(Packet) + 127) & ~(size_t) 127;
        const size_t out_bytes = ((n_in + 1 + GFB_MAX_GRIDS) * sizeof(Packet) + 127) & ~(size_t) 127;
        void* mem = nullptr;
        CUDA_TRY(cudaHostAlloc(&mem, sizeof(ResidentCtl) + in_bytes + out_bytes, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(mem, 0, sizeof(ResidentCtl) + in_bytes + out_bytes);
        r->ctl = static_cast<ResidentCtl*>(mem);
        r->in = reinterpret_cast<Packet*>(static_cast<char*>(mem) + sizeof(ResidentCtl));
        r->out = reinterpret_cast<Packet*>(static_cast<char*>(mem) + sizeof(ResidentCtl) + in_bytes);
        r->n_atoms = na;
        r->seq = 0;
        r->particles.clear();
        if (k->d_particles) {
            r->particles.resize((size_t) na);
            CUDA_TRY(cudaMemcpy(r->particles.data(), k->d_particles, (size_t) na * sizeof(int), cudaMemcpyDeviceToHost));
        }
    }
    if (!r->stream) CUDA_TRY(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    ResidentCtl* ctl = r->ctl;
    const int* hp = r->particles.empty() ? nullptr : r->particles.data();
    const unsigned long long n = ++r->seq;
    const unsigned tag = (unsigned) (n << 2) | (forces ? 1u : 0u) | (grid_energies ? 2u : 0u);
    for (int a = 0; a < na; a++) {
        const double* src = pos + 3 * (size_t) (hp ? hp[a] : a);
        for (int c = 0; c < 3; c++) {
            unsigned long long bits;
            memcpy(&bits, src + c, 8);
            store_packet(r->in + 3 * (size_t) a + c, pack(bits, tag));
        }
    }
    if (!r->running || !__atomic_load_n(&ctl->alive, __ATOMIC_ACQUIRE)) {
        int rc = resident_launch(k, r);     // joins a block that has timed out, then starts one at this step
        if (rc != GFB_OK) return rc;
    }
    // Result packets, in the order they are needed: [forces] | total energy | [per-grid energies]. Each is awaited on its
    // own; they arrive within microseconds, and the wall-clock checks only bound a failure.
    const size_t first = forces ? 0 : n_in;
    const size_t last = n_in + 1 + (grid_energies ? (size_t) k->n_grids : 0);
    double t0 = 0.0;
    unsigned long long spins = 0;
    for (size_t i = first; i < last;) {
        const Packet q = load_packet(r->out + i);
        if (packet_has(q, tag)) {
            double v;
            const unsigned long long bits = packet_bits(q);
            memcpy(&v, &bits, 8);
            if (i < n_in) {
                const size_t a = i / 3, c = i - 3 * a;
                double* dst = forces + 3 * (size_t) (hp ? hp[a] : (int) a) + c;
                *dst = add ? *dst + v : v;
            } else if (i == n_in) {
                if (energies) energies[0] = v;
            } else {
                grid_energies[i - n_in - 1] = v;
            }
            i++;
            continue;
        }
        if (!__atomic_load_n(&ctl->alive, __ATOMIC_ACQUIRE)) {
            // The block gave up waiting in the instant the step was posted (its exit is final once alive is clear, and it
            // ran no part of this step: a block that stops does so before it evaluates). Start over with a new block.
            if (packet_has(load_packet(r->out + i), tag)) continue;
            int rc = resident_launch(k, r);
            if (rc != GFB_OK) return rc;
            t0 = 0.0;
            continue;
        }
        if ((++spins & 0x3ffu) == 0) {
            const double t = now_s();
            if (t0 == 0.0) t0 = t;
            if (t - t0 > 5.0) {
                const cudaError_t qe = cudaStreamQuery(r->stream);
                __atomic_store_n(&ctl->cmd, kStop, __ATOMIC_RELEASE);
                return fail(GFB_ERR_CUDA, "resident evaluator did not answer step %llu within 5 s (%s)", n,
                            qe == cudaErrorNotReady ? "block still running" : cudaGetErrorString(qe));
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
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

int gfb_kernel_resident_timeline(const gfb_kernel* k, double us[2]) {
    if (!k || !us) return fail(GFB_ERR_INVALID, "gfb_kernel_resident_timeline: NULL argument");
    const ResidentState* r = static_cast<const ResidentState*>(k->resident);
    if (!r || !r->ctl || r->seq == 0) return fail(GFB_ERR_INVALID, "gfb_kernel_resident_timeline: no resident step has run");
    const unsigned long long* t = r->ctl->stamps;
    for (int i = 0; i < 2; i++) us[i] = 1e-3 * (double) (long long) (t[i + 1] - t[i]);
    return GFB_OK;
}

long long gfb_kernel_resident_launches(const gfb_kernel* k) {
    if (!k || !k->resident) return 0;
    return (long long) static_cast<const ResidentState*>(k->resident)->launches;
}

}  // extern "C"
