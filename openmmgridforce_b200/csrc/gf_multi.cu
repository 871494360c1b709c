// Replica-sharded multi-GPU runs and CUDA graphs (gfb_graph_*, gfb_comm_*, gfb_multi_* of include/gridforce_b200.h).
//
// The path has no exchange step: replicas are independent (ReferenceGridForceKernels.cpp:682-1118 iterates atoms
// independently; example/sampler.py:130-151 keeps one Context per replica), grids are replicated, and the only
// cross-GPU traffic is the gather of per-replica energies. That gather exists here in two forms:
//   * ncclAllGather (NCCL resolved with dlopen at first use — the library has no link-time dependency on it);
//   * fused into the evaluation kernel: the last block of the launch stores the energies into every peer's gathered
//     array through NVLink peer mappings and raises an arrival flag there (gather_tail in gf_eval_lines.cuh); the
//     consumer side is gf_gather_wait_kernel below.
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: every call goes through the dlopen'ed table below

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "gf_gather.cuh"
#include "gf_handles.h"

using namespace gfb;

// ---------------------------------------------------------------------------------------------
// NCCL, loaded on demand
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("GFB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.error = "cannot load libnccl.so.2 (set GFB_NCCL_LIB to its path)";
            return;
        }
#define GFB_NCCL_SYM(name)                                                          \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.lib, "nccl" #name));  \
    if (!api.name) api.error = "libnccl lacks nccl" #name;
        GFB_NCCL_SYM(GetUniqueId)
        GFB_NCCL_SYM(CommInitRank)
        GFB_NCCL_SYM(CommInitAll)
        GFB_NCCL_SYM(CommDestroy)
        GFB_NCCL_SYM(AllGather)
        GFB_NCCL_SYM(GroupStart)
        GFB_NCCL_SYM(GroupEnd)
        GFB_NCCL_SYM(GetErrorString)
#undef GFB_NCCL_SYM
    });
    return &api;
}

#define NCCL_TRY(api, expr)                                                                                     \
    do {                                                                                                        \
        ncclResult_t r__ = (expr);                                                                              \
        if (r__ != ncclSuccess)                                                                                 \
            return gfb::fail(GFB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

// ---- consumer side of the fused gather ------------------------------------------------------------------------------
// kWaitBlocks blocks of 256 threads (enough that a 65,536-entry array is one 16-byte copy per thread). In every block, lane r of warp 0 waits until rank r has published this gather's
// sequence number in this rank's flag array (written by the last block of rank r's evaluation launch with
// st.release.sys after its data stores and a system-scope fence); then the block copies its share of the gathered array
// into `out`, the caller's stable result buffer (the gathered array itself is double-buffered and will be overwritten
// two gathers later). The sequence number is read from the device-resident table, so a captured graph replays
// correctly; the last block to finish advances it. Bounded: after ~20 s a block raises *timed_out instead of spinning
// forever (a rank that died must not hang the GPU).
constexpr int kWaitBlocks = 64;
__global__ void __launch_bounds__(256) gf_gather_wait_kernel(GatherTable* gt, const unsigned long long* flags_base,
                                                             const double* data_base, double* out) {
    __shared__ int s_ok;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // see launch_overlapped
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(&gt->waited) + 1ull;
    const int parity = (int) (seq & 1ull);
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if ((int) threadIdx.x < gt->n_peers) {
        const unsigned long long* f = flags_base + (size_t) parity * kMaxPeers + threadIdx.x;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (unsigned spin = 0;; spin++) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= seq) break;
            if ((spin & 1023u) == 1023u) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) {
                    gt->timed_out = 1u;
                    s_ok = 0;
                    break;
                }
            }
            if (spin > 64) __nanosleep(100);
        }
    }
    __syncthreads();
    if (s_ok && out) {
        const double* src = data_base + (size_t) parity * gt->count_total;
        const long long n = gt->count_total;
        if (((n | (long long) parity * n) & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            const double2* src2 = reinterpret_cast<const double2*>(src);
            double2* out2 = reinterpret_cast<double2*>(out);
            for (long long i = (long long) blockIdx.x * 256 + threadIdx.x; i < n / 2; i += (long long) gridDim.x * 256) out2[i] = __ldcg(src2 + i);
        } else {
            for (long long i = (long long) blockIdx.x * 256 + threadIdx.x; i < n; i += (long long) gridDim.x * 256) out[i] = __ldcg(src + i);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&gt->wait_ticket, 1u) == gridDim.x - 1) {   // every block has read `waited` by now
            gt->waited = seq;
            gt->wait_ticket = 0;
        }
    }
}

// ---- stand-alone producer: the same peer stores + flags as the fused tail, as a kernel of its own -----------------------
// (kPushBlocks x 256 threads, enqueued right after the evaluation launch: it costs one more small launch but nothing
// inside the evaluation kernel)
constexpr int kPushBlocks = 32;
__global__ void __launch_bounds__(256) gf_gather_push_kernel(GatherTable* gt, const double* src, int n, long long offset) {
    // launched with programmatic stream serialization: the blocks are resident before the evaluation kernel ends and go
    // the moment it has completed and flushed (no launch latency on the critical path); the wait kernel behind may do
    // the same
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    gather_publish<256>(gt, src, n, offset, blockIdx.x, gridDim.x);
}

// Device-side rendezvous of all ranks on their streams: every rank publishes its arrival number in its slot of every
// peer's rendezvous row (third row of the flag array) and waits until all peers' numbers have reached its own row. With
// `go` != nullptr the rank first waits until its own HOST has raised *go to go_value (host-mapped pinned word,
// gfb_comm_rendezvous_release): the host enqueues the work that follows the rendezvous first and releases then, so that
// when the ranks leave the rendezvous — within one NVLink round trip of each other — none of them has to wait for its
// host to submit. One block; bounded like the gather wait (~20 s, then timed_out).
__global__ void __launch_bounds__(32) gf_rendezvous_kernel(GatherTable* gt, const volatile unsigned int* go, unsigned int go_value) {
    __shared__ int s_ok;
    const unsigned lane = threadIdx.x;
    if (lane == 0) s_ok = 1;
    __syncwarp();
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (go != nullptr && lane == 0) {
        for (unsigned spin = 0; *go < go_value; spin++) {
            if ((spin & 255u) == 255u) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) {
                    s_ok = 0;
                    break;
                }
            }
            __nanosleep(200);
        }
    }
    __syncwarp();
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(&gt->rendezvous) + 1ull;
    if (s_ok && (int) lane < gt->n_peers) {
        unsigned long long* theirs = gt->peer_flags[lane] + 2 * kMaxPeers + gt->my_rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(seq) : "memory");
        const unsigned long long* mine = gt->peer_flags[gt->my_rank] + 2 * kMaxPeers + lane;
        for (unsigned spin = 0;; spin++) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            if (v >= seq) break;
            if ((spin & 1023u) == 1023u) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) {
                    s_ok = 0;
                    break;
                }
            }
        }
    }
    __syncwarp();
    if (lane == 0) {
        if (!s_ok) gt->timed_out = 1u;
        gt->rendezvous = seq;
    }
}

// Flag-in-data gather: publish + wait + copy-out in ONE kernel with no fence, no ticket and no flag round on the critical
// path (the protocol NCCL calls LL). Every double travels as a 16-byte packet {lo32, seq, hi32, seq}: the 8-byte halves
// are written atomically, so a consumer that sees seq in both halves holds the whole value — the data is its own arrival
// flag. Each block first stores its share of this rank's slice into every peer's packet array (plain 16-byte stores over
// the NVLink peer mappings, staggered start), then polls its share of THIS rank's packet array — all ranks' slices, its
// own included — and writes the values to `out`. Critical path after the producing kernel: one L2 read, one NVLink
// one-way trip, one poll. Packet arrays are double-buffered by the parity of seq: a rank can start gather k+2 only after
// its gather k+1 has seen every peer's k+1 packets, which those peers sent after finishing their gather k (stream order),
// i.e. after they stopped reading buffer k%2. seq lives in device memory (graph replay). Bounded spin as everywhere else.
constexpr int kLLBlocks = 64;
__global__ void __launch_bounds__(256) gf_gather_ll_kernel(GatherTable* gt, const double* src, int n, long long offset, double* out) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");      // src is complete (the evaluation launch in front)
    const unsigned long long seq64 = *reinterpret_cast<volatile unsigned long long*>(&gt->ll_seq) + 1ull;
    const unsigned seq = (unsigned) seq64;
    const long long total = gt->count_total;
    const long long buf = (long long) (seq & 1u) * total;
    const int np = gt->n_peers, me = gt->my_rank;
    const unsigned gtid = blockIdx.x * 256u + threadIdx.x, gsize = gridDim.x * 256u;
    for (unsigned i = gtid; i < (unsigned) n; i += gsize) {
        const double v = __ldcg(src + i);
        const unsigned lo = (unsigned) __double2loint(v), hi = (unsigned) __double2hiint(v);
        for (int r = 0; r < np; r++) {
            char* base = reinterpret_cast<char*>(gt->peer_data[(me + 1 + r) % np]) + gt->ll_offset;
            uint4* dst = reinterpret_cast<uint4*>(base) + buf + offset + i;
            asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(seq), "r"(hi), "r"(seq) : "memory");
        }
    }
    const uint4* mine = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(gt->peer_data[me]) + gt->ll_offset) + buf;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    bool ok = true;
    for (long long j = gtid; j < total && ok; j += gsize) {
        unsigned a, b, c, d;
        for (unsigned spin = 0;; spin++) {
            asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(mine + j) : "memory");
            if (b == seq && d == seq) break;
            if ((spin & 4095u) == 4095u) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) {
                    gt->timed_out = 1u;
                    ok = false;
                    break;
                }
            }
        }
        if (ok && out) out[j] = __hiloint2double((int) c, (int) a);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&gt->ll_ticket, 1u) == gridDim.x - 1) {   // every block has read ll_seq by now
            gt->ll_seq = seq64;
            gt->ll_ticket = 0;
        }
    }
}

// Launch with the programmatic-stream-serialization attribute (both gather kernels start with griddepcontrol.wait, so
// the attribute only moves their block scheduling ahead of the previous kernel's end; the ordering is unchanged).
template <typename... Args>
cudaError_t launch_overlapped(void (*kernel)(Args...), int blocks, int threads, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// This rank's gather memory: [2][count_total] doubles, then [2][kMaxPeers] arrival flags.
struct GatherMem {
    void* base = nullptr;
    size_t count_total = 0, flags_offset = 0, ll_offset = 0, bytes = 0;
    int alloc(size_t count) {
        count_total = count;
        flags_offset = round_up(2 * count * sizeof(double), 256);
        ll_offset = round_up(flags_offset + 3 * kMaxPeers * sizeof(unsigned long long), 256);   // flag rows 0/1: gather parities, 2: rendezvous
        bytes = ll_offset + 2 * count * 16;                                                      // flag-in-data packets, [2][count]
        CUDA_TRY(cudaMalloc(&base, bytes));
        CUDA_TRY(cudaMemset(base, 0, bytes));
        return GFB_OK;
    }
    double* data(int parity) const { return static_cast<double*>(base) + (size_t) parity * count_total; }
    unsigned long long* flags(int parity) const {
        return reinterpret_cast<unsigned long long*>(static_cast<char*>(base) + flags_offset) + (size_t) parity * kMaxPeers;
    }
};

// Builds the device-side peer table from the bases of every rank's gather memory (valid on this device).
int upload_table(const GatherMem& mine, void* const* peer_base, int world, int rank, GatherTable** d_table) {
    GatherTable t;
    memset(&t, 0, sizeof t);
    for (int r = 0; r < world; r++) {
        t.peer_data[r] = static_cast<double*>(peer_base[r]);
        t.peer_flags[r] = reinterpret_cast<unsigned long long*>(static_cast<char*>(peer_base[r]) + mine.flags_offset);
    }
    t.count_total = (long long) mine.count_total;
    t.ll_offset = (long long) mine.ll_offset;
    t.n_peers = world;
    t.my_rank = rank;
    if (!*d_table) CUDA_TRY(cudaMalloc((void**) d_table, sizeof t));
    CUDA_TRY(cudaMemcpy(*d_table, &t, sizeof t, cudaMemcpyHostToDevice));
    return GFB_OK;
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// Handles
// ---------------------------------------------------------------------------------------------
struct gfb_graph {
    gfb_device* dev;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
};

struct gfb_comm {
    gfb_device* dev;
    int world, rank;
    ncclComm_t nccl;
    GatherMem mem;
    GatherTable* d_table;
    void* peer_base[kMaxPeers];   // cudaIpcOpenMemHandle mappings (self: mem.base)
    bool attached;
    unsigned int* h_go;           // host-mapped pinned word the held rendezvous kernel polls (gfb_comm_rendezvous hold = 1)
    unsigned int* d_go;           // its device address
    unsigned int go_armed;        // value the most recent held rendezvous waits for
};

struct MultiShard {
    int lo = 0, hi = 0;                   // replicas [lo, hi)
    double* d_pos = nullptr;
    unsigned long long* d_forces = nullptr;   // OpenMM fixed point, [3][stride]
    long long stride = 0;
    double* d_e[2] = {nullptr, nullptr};  // alternating accumulators: a launch adds into one and clears the other
    double* d_padded = nullptr;           // ncclAllGather result, [n][width]
    double* d_gathered = nullptr;         // fused gather result, [n_replicas]
    GatherMem mem;
    GatherTable* d_table = nullptr;
};

struct gfb_multi {
    int n;
    std::vector<gfb_device*> devs;
    std::vector<std::vector<gfb_grid*> > grids;   // [device][grid]
    std::vector<gfb_kernel*> kernels;
    std::vector<ncclComm_t> nccl;                  // created at the first gather == 1 step
    std::vector<MultiShard> shards;
    int n_atoms, n_replicas, width;                // width = largest shard
    unsigned long long steps;
    int last_gather;
};

static void shard_bounds(int n_units, int world, int rank, int& lo, int& hi) {   // rank g owns [g*n/N, (g+1)*n/N)
    lo = (int) ((long long) n_units * rank / world);
    hi = (int) ((long long) n_units * (rank + 1) / world);
}

extern "C" {

// ---------------------------------------------------------------------------------------------
// Graphs
// ---------------------------------------------------------------------------------------------
int gfb_graph_begin(gfb_device* dev, void* stream) {
    if (!dev) return fail(GFB_ERR_INVALID, "gfb_graph_begin: NULL device");
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : dev->stream;
    CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    return GFB_OK;
}

int gfb_graph_end(gfb_device* dev, void* stream, gfb_graph** out) {
    if (!dev || !out) return fail(GFB_ERR_INVALID, "gfb_graph_end: NULL argument");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : dev->stream;
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamEndCapture(s, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(graph);
        return fail(GFB_ERR_CUDA, "gfb_graph_end: cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    gfb_graph* g = new (std::nothrow) gfb_graph();
    if (!g) return fail(GFB_ERR_NOMEM, "gfb_graph_end: out of host memory");
    g->dev = dev;
    g->graph = graph;
    g->exec = exec;
    *out = g;
    return GFB_OK;
}

int gfb_graph_launch(gfb_graph* g, void* stream) {
    if (!g) return fail(GFB_ERR_INVALID, "gfb_graph_launch: NULL graph");
    CUDA_TRY(cudaSetDevice(g->dev->ordinal));
    CUDA_TRY(cudaGraphLaunch(g->exec, stream ? static_cast<cudaStream_t>(stream) : g->dev->stream));
    return GFB_OK;
}

int gfb_graph_destroy(gfb_graph* g) {
    if (!g) return GFB_OK;
    cudaSetDevice(g->dev->ordinal);
    cudaGraphExecDestroy(g->exec);
    cudaGraphDestroy(g->graph);
    delete g;
    return GFB_OK;
}

// ---------------------------------------------------------------------------------------------
// One process per GPU
// ---------------------------------------------------------------------------------------------
int gfb_comm_unique_id(unsigned char id[GFB_COMM_ID_BYTES]) {
    if (!id) return fail(GFB_ERR_INVALID, "gfb_comm_unique_id: NULL argument");
    NcclApi* api = nccl_api();
    if (!api->error.empty()) return fail(GFB_ERR_UNSUPPORTED, "gfb_comm_unique_id: %s", api->error.c_str());
    static_assert(sizeof(ncclUniqueId) == GFB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId uid;
    NCCL_TRY(api, api->GetUniqueId(&uid));
    memcpy(id, &uid, sizeof uid);
    return GFB_OK;
}

int gfb_comm_create(gfb_device* dev, int world_size, int rank, const unsigned char id[GFB_COMM_ID_BYTES], gfb_comm** out) {
    if (!dev || !out) return fail(GFB_ERR_INVALID, "gfb_comm_create: NULL argument");
    *out = nullptr;
    if (world_size < 1 || world_size > kMaxPeers || rank < 0 || rank >= world_size)
        return fail(GFB_ERR_INVALID, "gfb_comm_create: rank %d of %d (at most %d ranks)", rank, world_size, kMaxPeers);
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    gfb_comm* c = new (std::nothrow) gfb_comm();
    if (!c) return fail(GFB_ERR_NOMEM, "gfb_comm_create: out of host memory");
    c->dev = dev;
    c->world = world_size;
    c->rank = rank;
    c->nccl = nullptr;
    c->d_table = nullptr;
    c->attached = false;
    c->h_go = c->d_go = nullptr;
    c->go_armed = 0;
    memset(c->peer_base, 0, sizeof c->peer_base);
    if (id) {   // id == NULL: no NCCL communicator (fused gather only)
        NcclApi* api = nccl_api();
        if (!api->error.empty()) {
            delete c;
            return fail(GFB_ERR_UNSUPPORTED, "gfb_comm_create: %s", api->error.c_str());
        }
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof uid);
        ncclResult_t r = api->CommInitRank(&c->nccl, world_size, uid, rank);
        if (r != ncclSuccess) {
            delete c;
            return fail(GFB_ERR_CUDA, "gfb_comm_create: ncclCommInitRank: %s", api->GetErrorString(r));
        }
    }
    *out = c;
    return GFB_OK;
}

int gfb_comm_destroy(gfb_comm* c) {
    if (!c) return GFB_OK;
    cudaSetDevice(c->dev->ordinal);
    cudaDeviceSynchronize();
    if (c->nccl) nccl_api()->CommDestroy(c->nccl);
    for (int r = 0; r < c->world; r++)
        if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
    if (c->mem.base) cudaFree(c->mem.base);
    if (c->d_table) cudaFree(c->d_table);
    if (c->h_go) cudaFreeHost(c->h_go);
    delete c;
    return GFB_OK;
}

int gfb_comm_all_gather(gfb_comm* c, const double* d_send, double* d_recv, size_t count, void* stream) {
    if (!c || !d_send || !d_recv) return fail(GFB_ERR_INVALID, "gfb_comm_all_gather: NULL argument");
    if (!c->nccl) return fail(GFB_ERR_INVALID, "gfb_comm_all_gather: this communicator was created without an NCCL id");
    CUDA_TRY(cudaSetDevice(c->dev->ordinal));
    NcclApi* api = nccl_api();
    NCCL_TRY(api, api->AllGather(d_send, d_recv, count, ncclDouble, c->nccl, stream ? static_cast<cudaStream_t>(stream) : c->dev->stream));
    return GFB_OK;
}

int gfb_comm_gather_alloc(gfb_comm* c, size_t count_total, unsigned char handle_out[GFB_IPC_HANDLE_BYTES]) {
    if (!c || !handle_out || count_total == 0) return fail(GFB_ERR_INVALID, "gfb_comm_gather_alloc: bad argument");
    if (c->mem.base) return fail(GFB_ERR_INVALID, "gfb_comm_gather_alloc: already allocated");
    CUDA_TRY(cudaSetDevice(c->dev->ordinal));
    int rc = c->mem.alloc(count_total);
    if (rc != GFB_OK) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == GFB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->mem.base));
    memcpy(handle_out, &h, sizeof h);
    return GFB_OK;
}

int gfb_comm_gather_attach(gfb_comm* c, const unsigned char* handles) {
    if (!c || !handles) return fail(GFB_ERR_INVALID, "gfb_comm_gather_attach: NULL argument");
    if (!c->mem.base) return fail(GFB_ERR_INVALID, "gfb_comm_gather_attach: call gfb_comm_gather_alloc first");
    if (c->attached) return fail(GFB_ERR_INVALID, "gfb_comm_gather_attach: already attached");
    CUDA_TRY(cudaSetDevice(c->dev->ordinal));
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) {
            c->peer_base[r] = c->mem.base;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t) r * GFB_IPC_HANDLE_BYTES, sizeof h);
        CUDA_TRY(cudaIpcOpenMemHandle(&c->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    int rc = upload_table(c->mem, c->peer_base, c->world, c->rank, &c->d_table);
    if (rc != GFB_OK) return rc;
    c->attached = true;
    return GFB_OK;
}

int gfb_kernel_execute_device_gather(gfb_kernel* k, int n_replicas, int n_particles, const double* d_pos, double* d_energies,
                                     void* d_forces, int force_mode, long long force_stride, double* d_energies_clear,
                                     gfb_comm* c, size_t gather_offset, void* stream) {
    int rc = check_exec_args("gfb_kernel_execute_device_gather", k, n_replicas, n_particles, d_pos, force_mode);
    if (rc != GFB_OK) return rc;
    if (!c || !c->attached) return fail(GFB_ERR_INVALID, "gfb_kernel_execute_device_gather: communicator without attached gather memory");
    if (c->dev != k->dev) return fail(GFB_ERR_INVALID, "gfb_kernel_execute_device_gather: kernel and communicator live on different devices");
    if (gather_offset + (size_t) n_replicas * k->n_slots > c->mem.count_total)
        return fail(GFB_ERR_INVALID, "gfb_kernel_execute_device_gather: slice [%zu, +%zu) exceeds the gathered array (%zu)", gather_offset,
                    (size_t) n_replicas * k->n_slots, c->mem.count_total);
    if (force_mode == GFB_FORCE_FIXED_ADD && d_forces && force_stride < (long long) n_replicas * n_particles)
        return fail(GFB_ERR_INVALID, "gfb_kernel_execute_device_gather: force_stride=%lld < %lld particles", force_stride,
                    (long long) n_replicas * n_particles);
    CUDA_TRY(cudaSetDevice(k->dev->ordinal));
    EvalExtra x;
    x.overlap = k->launch_overlap;
    x.gather = c->d_table;
    x.gather_offset = (long long) gather_offset;
    return enqueue_eval(k, n_replicas, n_particles, d_pos, d_energies, nullptr, d_forces, force_mode, force_stride, nullptr,
                        d_energies_clear, stream ? static_cast<cudaStream_t>(stream) : k->dev->stream, x);
}

int gfb_comm_gather_push(gfb_comm* c, const double* d_energies, size_t count, size_t gather_offset, void* stream) {
    if (!c || !c->attached) return fail(GFB_ERR_INVALID, "gfb_comm_gather_push: communicator without attached gather memory");
    if (!d_energies || count == 0) return fail(GFB_ERR_INVALID, "gfb_comm_gather_push: nothing to push");
    if (gather_offset + count > c->mem.count_total || count > 0x7fffffffull)
        return fail(GFB_ERR_INVALID, "gfb_comm_gather_push: slice [%zu, +%zu) exceeds the gathered array (%zu)", gather_offset, count, c->mem.count_total);
    CUDA_TRY(cudaSetDevice(c->dev->ordinal));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->dev->stream;
    CUDA_TRY(launch_overlapped(gf_gather_push_kernel, kPushBlocks, 256, s, c->d_table, d_energies, (int) count, (long long) gather_offset));
    g_launches++;
    return GFB_OK;
}

int gfb_comm_gather(gfb_comm* c, const double* d_energies, size_t count, size_t gather_offset, double* d_out, void* stream) {
    if (!c || !c->attached) return fail(GFB_ERR_INVALID, "gfb_comm_gather: communicator without attached gather memory");
    if (!d_energies || count == 0 || !d_out) return fail(GFB_ERR_INVALID, "gfb_comm_gather: NULL or empty argument");
    if (gather_offset + count > c->mem.count_total || count > 0x7fffffffull)
        return fail(GFB_ERR_INVALID, "gfb_comm_gather: slice [%zu, +%zu) exceeds the gathered array (%zu)", gather_offset, count, c->mem.count_total);
    CUDA_TRY(cudaSetDevice(c->dev->ordinal));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->dev->stream;
    CUDA_TRY(launch_overlapped(gf_gather_ll_kernel, kLLBlocks, 256, s, c->d_table, d_energies, (int) count, (long long) gather_offset, d_out));
    g_launches++;
    return GFB_OK;
}

int gfb_comm_gather_wait(gfb_comm* c, double* d_out, void* stream) {
    if (!c || !c->attached) return fail(GFB_ERR_INVALID, "gfb_comm_gather_wait: communicator without attached gather memory");
    if (!d_out) return fail(GFB_ERR_INVALID, "gfb_comm_gather_wait: d_out is NULL");
    CUDA_TRY(cudaSetDevice(c->dev->ordinal));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->dev->stream;
    CUDA_TRY(launch_overlapped(gf_gather_wait_kernel, kWaitBlocks, 256, s, c->d_table, (const unsigned long long*) c->mem.flags(0),
                               (const double*) c->mem.data(0), d_out));
    g_launches++;
    return GFB_OK;
}

int gfb_comm_rendezvous(gfb_comm* c, int hold, void* stream) {
    if (!c || !c->attached) return fail(GFB_ERR_INVALID, "gfb_comm_rendezvous: communicator without attached gather memory");
    CUDA_TRY(cudaSetDevice(c->dev->ordinal));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->dev->stream;
    const unsigned int* go = nullptr;
    if (hold) {
        if (!c->h_go) {
            CUDA_TRY(cudaHostAlloc((void**) &c->h_go, sizeof(unsigned int), cudaHostAllocMapped));
            *c->h_go = 0u;
            CUDA_TRY(cudaHostGetDevicePointer((void**) &c->d_go, c->h_go, 0));
        }
        c->go_armed++;
        go = c->d_go;
    }
    gf_rendezvous_kernel<<<1, 32, 0, s>>>(c->d_table, go, c->go_armed);
    CUDA_TRY(cudaGetLastError());
    g_launches++;
    return GFB_OK;
}

int gfb_comm_rendezvous_release(gfb_comm* c) {
    if (!c || !c->h_go) return fail(GFB_ERR_INVALID, "gfb_comm_rendezvous_release: no held rendezvous was enqueued");
    __atomic_store_n(c->h_go, c->go_armed, __ATOMIC_RELEASE);
    return GFB_OK;
}

int gfb_comm_gather_status(gfb_comm* c) {
    if (!c || !c->d_table) return fail(GFB_ERR_INVALID, "gfb_comm_gather_status: communicator without attached gather memory");
    CUDA_TRY(cudaSetDevice(c->dev->ordinal));
    unsigned int flag = 0;
    CUDA_TRY(cudaMemcpy(&flag, &c->d_table->timed_out, sizeof flag, cudaMemcpyDeviceToHost));
    if (flag) return fail(GFB_ERR_CUDA, "fused energy gather: a peer's slice did not arrive within 20 s (rank %d of %d)", c->rank, c->world);
    return GFB_OK;
}

// ---------------------------------------------------------------------------------------------
// One process, all GPUs
// ---------------------------------------------------------------------------------------------
static void multi_free_shards(gfb_multi* m) {
    for (int d = 0; d < (int) m->shards.size(); d++) {
        MultiShard& s = m->shards[d];
        cudaSetDevice(m->devs[d]->ordinal);
        cudaDeviceSynchronize();
        if (s.d_pos) cudaFree(s.d_pos);
        if (s.d_forces) cudaFree(s.d_forces);
        if (s.d_e[0]) cudaFree(s.d_e[0]);
        if (s.d_e[1]) cudaFree(s.d_e[1]);
        if (s.d_padded) cudaFree(s.d_padded);
        if (s.d_gathered) cudaFree(s.d_gathered);
        if (s.mem.base) cudaFree(s.mem.base);
        if (s.d_table) cudaFree(s.d_table);
    }
    m->shards.clear();
    m->n_replicas = 0;
}

int gfb_multi_create(int n_devices, const int* ordinals, gfb_multi** out) {
    if (!out || n_devices < 1 || n_devices > kMaxPeers) return fail(GFB_ERR_INVALID, "gfb_multi_create: 1..%d devices", kMaxPeers);
    *out = nullptr;
    gfb_multi* m = new (std::nothrow) gfb_multi();
    if (!m) return fail(GFB_ERR_NOMEM, "gfb_multi_create: out of host memory");
    m->n = n_devices;
    m->n_atoms = m->n_replicas = m->width = 0;
    m->steps = 0;
    m->last_gather = 0;
    for (int d = 0; d < n_devices; d++) {
        const int ord = ordinals ? ordinals[d] : d;
        for (int e = 0; e < d; e++)
            if (m->devs[e]->ordinal == ord) {
                gfb_multi_destroy(m);
                return fail(GFB_ERR_INVALID, "gfb_multi_create: device %d listed twice", ord);
            }
        gfb_device* dev = nullptr;
        int rc = gfb_device_open(ord, &dev);
        if (rc != GFB_OK) {
            gfb_multi_destroy(m);
            return rc;
        }
        m->devs.push_back(dev);
    }
    m->grids.resize(n_devices);
    m->kernels.assign(n_devices, nullptr);
    // peer access both ways between every pair (the fused gather stores into peer memory)
    for (int a = 0; a < n_devices; a++)
        for (int b = 0; b < n_devices; b++) {
            if (a == b) continue;
            cudaSetDevice(m->devs[a]->ordinal);
            cudaError_t e = cudaDeviceEnablePeerAccess(m->devs[b]->ordinal, 0);
            if (e != cudaSuccess) cudaGetLastError();   // already enabled, or no peer path: the fused gather will say so
        }
    *out = m;
    return GFB_OK;
}

int gfb_multi_destroy(gfb_multi* m) {
    if (!m) return GFB_OK;
    multi_free_shards(m);
    for (size_t d = 0; d < m->nccl.size(); d++)
        if (m->nccl[d]) nccl_api()->CommDestroy(m->nccl[d]);
    for (size_t d = 0; d < m->kernels.size(); d++)
        if (m->kernels[d]) gfb_kernel_destroy(m->kernels[d]);
    for (size_t d = 0; d < m->grids.size(); d++)
        for (size_t g = 0; g < m->grids[d].size(); g++) gfb_grid_destroy(m->grids[d][g]);
    for (size_t d = 0; d < m->devs.size(); d++) gfb_device_close(m->devs[d]);
    delete m;
    return GFB_OK;
}

int gfb_multi_num_devices(const gfb_multi* m) { return m ? m->n : 0; }

int gfb_multi_add_grid(gfb_multi* m, const int counts[3], const double spacing[3], const double origin[3], const double* vals,
                       size_t n_vals, int precision, int layout) {
    if (!m) return fail(GFB_ERR_INVALID, "gfb_multi_add_grid: NULL handle");
    if (m->kernels[0]) return fail(GFB_ERR_INVALID, "gfb_multi_add_grid: the evaluation state is already built");
    if ((int) m->grids[0].size() >= GFB_MAX_GRIDS) return fail(GFB_ERR_INVALID, "gfb_multi_add_grid: at most %d grids", GFB_MAX_GRIDS);
    // one upload + repack per device, all devices at once (each has its own stream and copy engines)
    std::vector<gfb_grid*> made(m->n, nullptr);
    std::vector<int> rcs(m->n, GFB_OK);
    std::vector<std::string> errs(m->n);
    std::vector<std::thread> th;
    for (int d = 0; d < m->n; d++)
        th.emplace_back([&, d] {
            rcs[d] = gfb_grid_create(m->devs[d], counts, spacing, origin, vals, n_vals, precision, layout, &made[d]);
            if (rcs[d] != GFB_OK) errs[d] = gfb_last_error();
        });
    for (auto& t : th) t.join();
    for (int d = 0; d < m->n; d++)
        if (rcs[d] != GFB_OK) {
            for (int e = 0; e < m->n; e++)
                if (made[e]) gfb_grid_destroy(made[e]);
            return fail(rcs[d], "gfb_multi_add_grid (device %d): %s", m->devs[d]->ordinal, errs[d].c_str());
        }
    for (int d = 0; d < m->n; d++) m->grids[d].push_back(made[d]);
    return (int) m->grids[0].size() - 1;
}

int gfb_multi_build(gfb_multi* m, int n_atoms, const double* scaling, const double* inv_power, const double* oob_k) {
    if (!m) return fail(GFB_ERR_INVALID, "gfb_multi_build: NULL handle");
    if (m->grids[0].empty()) return fail(GFB_ERR_INVALID, "gfb_multi_build: add at least one grid first");
    if (m->kernels[0] && !m->grids[0][0]->cells)
        return fail(GFB_ERR_INVALID, "gfb_multi_build: the grids' packed cells were released by the first build; create a new handle to rebuild");
    for (int d = 0; d < m->n; d++) {
        if (m->kernels[d]) gfb_kernel_destroy(m->kernels[d]);
        m->kernels[d] = nullptr;
        int rc = gfb_kernel_create(m->devs[d], (int) m->grids[d].size(), m->grids[d].data(), n_atoms, scaling, nullptr, inv_power,
                                   oob_k, &m->kernels[d]);
        if (rc != GFB_OK) return rc;
    }
    m->n_atoms = n_atoms;
    // The grids belong to this handle alone: once every state reads them through its interleaved records, the per-grid
    // packed cells are dead weight (3 x 192^3: 638 MB per device).
    for (int d = 0; d < m->n; d++) {
        EvalParams probe;
        memset(&probe, 0, sizeof probe);
        const gfb_kernel* k = m->kernels[d];
        if (k->n_grids > 1 && (lines_eligible(k, probe) || lines_f64_eligible(k)))
            for (size_t g = 0; g < m->grids[d].size(); g++) gfb_grid_release_cells(m->grids[d][g]);
    }
    return GFB_OK;
}

int gfb_multi_execute_host(gfb_multi* m, int n_replicas, const double* pos, double* energies, void* forces, int force_mode) {
    if (!m || !m->kernels[0]) return fail(GFB_ERR_INVALID, "gfb_multi_execute_host: call gfb_multi_build first");
    if (n_replicas < 0 || (n_replicas > 0 && !pos)) return fail(GFB_ERR_INVALID, "gfb_multi_execute_host: bad arguments");
    if (force_mode != GFB_FORCE_F64_STORE && force_mode != GFB_FORCE_F32_STORE && force_mode != GFB_FORCE_F64_ADD)
        return fail(GFB_ERR_INVALID, "gfb_multi_execute_host: force_mode must be F64_STORE, F32_STORE or F64_ADD");
    const size_t fsz = force_mode == GFB_FORCE_F32_STORE ? sizeof(float) : sizeof(double);
    const size_t per_rep = (size_t) m->n_atoms * 3;
    std::vector<int> rcs(m->n, GFB_OK);
    std::vector<std::string> errs(m->n);
    std::vector<std::thread> th;
    for (int d = 0; d < m->n; d++) {
        int lo, hi;
        shard_bounds(n_replicas, m->n, d, lo, hi);
        if (hi == lo) continue;
        th.emplace_back([=, &rcs, &errs] {
            rcs[d] = gfb_kernel_execute_host(m->kernels[d], hi - lo, m->n_atoms, pos + (size_t) lo * per_rep,
                                             energies ? energies + lo : nullptr, nullptr,
                                             forces ? static_cast<char*>(forces) + (size_t) lo * per_rep * fsz : nullptr, force_mode);
            if (rcs[d] != GFB_OK) errs[d] = gfb_last_error();
        });
    }
    for (auto& t : th) t.join();
    for (int d = 0; d < m->n; d++)
        if (rcs[d] != GFB_OK) return fail(rcs[d], "gfb_multi_execute_host (device %d): %s", m->devs[d]->ordinal, errs[d].c_str());
    return GFB_OK;
}

int gfb_multi_upload(gfb_multi* m, int n_replicas, const double* pos) {
    if (!m || !m->kernels[0]) return fail(GFB_ERR_INVALID, "gfb_multi_upload: call gfb_multi_build first");
    if (n_replicas < m->n || !pos) return fail(GFB_ERR_INVALID, "gfb_multi_upload: need at least one replica per device");
    multi_free_shards(m);
    m->shards.resize(m->n);
    m->n_replicas = n_replicas;
    m->width = 0;
    m->steps = 0;
    m->last_gather = 0;
    const size_t per_rep = (size_t) m->n_atoms * 3;
    for (int d = 0; d < m->n; d++) {
        MultiShard& s = m->shards[d];
        shard_bounds(n_replicas, m->n, d, s.lo, s.hi);
        m->width = std::max(m->width, s.hi - s.lo);
    }
    for (int d = 0; d < m->n; d++) {
        MultiShard& s = m->shards[d];
        CUDA_TRY(cudaSetDevice(m->devs[d]->ordinal));
        const size_t reps = (size_t) (s.hi - s.lo);
        s.stride = (long long) ((reps * m->n_atoms + 31) / 32 * 32);
        CUDA_TRY(cudaMalloc((void**) &s.d_pos, reps * per_rep * sizeof(double)));
        CUDA_TRY(cudaMalloc((void**) &s.d_forces, (size_t) s.stride * 3 * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(s.d_forces, 0, (size_t) s.stride * 3 * sizeof(unsigned long long)));
        for (int b = 0; b < 2; b++) {
            CUDA_TRY(cudaMalloc((void**) &s.d_e[b], (size_t) m->width * sizeof(double)));
            CUDA_TRY(cudaMemset(s.d_e[b], 0, (size_t) m->width * sizeof(double)));
        }
        CUDA_TRY(cudaMalloc((void**) &s.d_padded, (size_t) m->n * m->width * sizeof(double)));
        CUDA_TRY(cudaMalloc((void**) &s.d_gathered, (size_t) n_replicas * sizeof(double)));
        CUDA_TRY(cudaMemcpy(s.d_pos, pos + (size_t) s.lo * per_rep, reps * per_rep * sizeof(double), cudaMemcpyHostToDevice));
        int rc = s.mem.alloc((size_t) n_replicas);
        if (rc != GFB_OK) return rc;
    }
    // peer tables: inside one process every device's allocation is directly addressable once peer access is on
    void* bases[kMaxPeers];
    for (int d = 0; d < m->n; d++) bases[d] = m->shards[d].mem.base;
    for (int d = 0; d < m->n; d++) {
        CUDA_TRY(cudaSetDevice(m->devs[d]->ordinal));
        int rc = upload_table(m->shards[d].mem, bases, m->n, d, &m->shards[d].d_table);
        if (rc != GFB_OK) return rc;
    }
    return GFB_OK;
}

int gfb_multi_step(gfb_multi* m, int gather) {
    if (!m || m->shards.empty()) return fail(GFB_ERR_INVALID, "gfb_multi_step: call gfb_multi_upload first");
    if (gather < 0 || gather > 4) return fail(GFB_ERR_INVALID, "gfb_multi_step: gather must be 0, 1, 2, 3 or 4");
    NcclApi* api = nullptr;
    if (gather == 1) {
        api = nccl_api();
        if (!api->error.empty()) return fail(GFB_ERR_UNSUPPORTED, "gfb_multi_step: %s", api->error.c_str());
        if (m->nccl.empty()) {
            std::vector<int> ords(m->n);
            for (int d = 0; d < m->n; d++) ords[d] = m->devs[d]->ordinal;
            m->nccl.assign(m->n, nullptr);
            NCCL_TRY(api, api->CommInitAll(m->nccl.data(), m->n, ords.data()));
        }
    }
    const int cur = (int) (m->steps & 1);
    for (int d = 0; d < m->n; d++) {
        MultiShard& s = m->shards[d];
        CUDA_TRY(cudaSetDevice(m->devs[d]->ordinal));
        EvalExtra x;
        if (gather == 2) {
            x.gather = s.d_table;
            x.gather_offset = s.lo;
        }
        int rc = enqueue_eval(m->kernels[d], s.hi - s.lo, m->n_atoms, s.d_pos, s.d_e[cur], nullptr, s.d_forces, GFB_FORCE_FIXED_ADD,
                              s.stride, nullptr, s.d_e[cur ^ 1], m->devs[d]->stream, x);
        if (rc != GFB_OK) return rc;
    }
    if (gather == 1) {
        NCCL_TRY(api, api->GroupStart());
        for (int d = 0; d < m->n; d++) {
            MultiShard& s = m->shards[d];
            ncclResult_t r = api->AllGather(s.d_e[cur], s.d_padded, (size_t) m->width, ncclDouble, m->nccl[d], m->devs[d]->stream);
            if (r != ncclSuccess) {
                api->GroupEnd();
                return fail(GFB_ERR_CUDA, "gfb_multi_step: ncclAllGather: %s", api->GetErrorString(r));
            }
        }
        NCCL_TRY(api, api->GroupEnd());
    } else if (gather >= 2) {
        for (int d = 0; d < m->n; d++) {
            MultiShard& s = m->shards[d];
            CUDA_TRY(cudaSetDevice(m->devs[d]->ordinal));
            if (gather == 4) {   // flag-in-data: publish + wait + copy-out in one kernel per device
                CUDA_TRY(launch_overlapped(gf_gather_ll_kernel, kLLBlocks, 256, m->devs[d]->stream, s.d_table, (const double*) s.d_e[cur],
                                           s.hi - s.lo, (long long) s.lo, s.d_gathered));
                g_launches++;
                continue;
            }
            if (gather == 3) {
                CUDA_TRY(launch_overlapped(gf_gather_push_kernel, kPushBlocks, 256, m->devs[d]->stream, s.d_table, (const double*) s.d_e[cur],
                                           s.hi - s.lo, (long long) s.lo));
                g_launches++;
            }
            CUDA_TRY(launch_overlapped(gf_gather_wait_kernel, kWaitBlocks, 256, m->devs[d]->stream, s.d_table,
                                       (const unsigned long long*) s.mem.flags(0), (const double*) s.mem.data(0), s.d_gathered));
            g_launches++;
        }
    }
    m->steps++;
    m->last_gather = gather;
    return GFB_OK;
}

int gfb_multi_download(gfb_multi* m, int from_device, double* energies, double* forces) {
    if (!m || m->shards.empty()) return fail(GFB_ERR_INVALID, "gfb_multi_download: call gfb_multi_upload first");
    if (from_device < 0 || from_device >= m->n) return fail(GFB_ERR_INVALID, "gfb_multi_download: from_device=%d", from_device);
    if (m->steps == 0) return fail(GFB_ERR_INVALID, "gfb_multi_download: no step has run");
    for (int d = 0; d < m->n; d++) {
        CUDA_TRY(cudaSetDevice(m->devs[d]->ordinal));
        CUDA_TRY(cudaStreamSynchronize(m->devs[d]->stream));
    }
    const int cur = (int) ((m->steps - 1) & 1);
    if (energies) {
        MultiShard& s = m->shards[from_device];
        CUDA_TRY(cudaSetDevice(m->devs[from_device]->ordinal));
        if (m->last_gather >= 2) {
            unsigned int flag = 0;
            CUDA_TRY(cudaMemcpy(&flag, &s.d_table->timed_out, sizeof flag, cudaMemcpyDeviceToHost));
            if (flag) return fail(GFB_ERR_CUDA, "gfb_multi_download: the fused gather timed out on device %d", m->devs[from_device]->ordinal);
            CUDA_TRY(cudaMemcpy(energies, s.d_gathered, (size_t) m->n_replicas * sizeof(double), cudaMemcpyDeviceToHost));
        } else if (m->last_gather == 1) {
            std::vector<double> padded((size_t) m->n * m->width);
            CUDA_TRY(cudaMemcpy(padded.data(), s.d_padded, padded.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (int d = 0; d < m->n; d++)
                memcpy(energies + m->shards[d].lo, padded.data() + (size_t) d * m->width,
                       (size_t) (m->shards[d].hi - m->shards[d].lo) * sizeof(double));
        } else {   // no gather ran: every device still holds only its own shard
            for (int d = 0; d < m->n; d++) {
                CUDA_TRY(cudaSetDevice(m->devs[d]->ordinal));
                CUDA_TRY(cudaMemcpy(energies + m->shards[d].lo, m->shards[d].d_e[cur],
                                    (size_t) (m->shards[d].hi - m->shards[d].lo) * sizeof(double), cudaMemcpyDeviceToHost));
            }
        }
    }
    if (forces) {
        for (int d = 0; d < m->n; d++) {
            MultiShard& s = m->shards[d];
            CUDA_TRY(cudaSetDevice(m->devs[d]->ordinal));
            const long long n = (long long) (s.hi - s.lo) * m->n_atoms;
            double* d_tmp = nullptr;
            CUDA_TRY(cudaMalloc((void**) &d_tmp, (size_t) n * 3 * sizeof(double)));
            int rc = gfb_forces_fixed_to_f64(m->devs[d], s.d_forces, s.stride, n, d_tmp, nullptr);
            cudaError_t e = cudaSuccess;
            if (rc == GFB_OK) e = cudaMemcpyAsync(forces + (size_t) s.lo * m->n_atoms * 3, d_tmp, (size_t) n * 3 * sizeof(double),
                                                  cudaMemcpyDeviceToHost, m->devs[d]->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(m->devs[d]->stream);
            cudaFree(d_tmp);
            if (rc != GFB_OK) return rc;
            if (e != cudaSuccess) return fail(GFB_ERR_CUDA, "gfb_multi_download: %s", cudaGetErrorString(e));
        }
    }
    return GFB_OK;
}

}  // extern "C"
