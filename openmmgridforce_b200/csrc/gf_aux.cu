// Around the evaluation: classification-only launches (parity tests), atom sort, fixed-point conversion, copy-engine
// peer put, roofline probes (gfb_kernel_classify_host / gfb_kernel_sort_atoms / gfb_forces_fixed_to_f64 / gfb_peer_put /
// gfb_bench_* of include/gridforce_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "gf_handles.h"
#include "gf_misc_kernels.cuh"

using namespace gfb;

extern "C" {

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
    // brick size and key width: 4-cell bricks, doubled until 7 bits per axis cover the grid (<= 2^21 bins)
    const gfb_grid* g0 = k->grids[0];
    const int max_nc = std::max(g0->counts[0], std::max(g0->counts[1], g0->counts[2])) - 1;
    int shift = 2;
    while ((max_nc >> shift) >= 128) shift++;
    int bits = 1;
    while ((max_nc >> shift) >= (1 << bits)) bits++;
    const unsigned n_bins = (1u << (3 * bits)) + 1u;
    const size_t key_bytes = ((size_t) total * sizeof(unsigned) + 255) & ~(size_t) 255;
    const size_t hist_bytes = ((size_t) n_bins * sizeof(unsigned) + 255) & ~(size_t) 255;
    if ((rc = k->d_sort.ensure(key_bytes + hist_bytes)) != GFB_OK) return rc;
    unsigned* keys = static_cast<unsigned*>(k->d_sort.ptr);
    unsigned* hist = reinterpret_cast<unsigned*>(static_cast<char*>(k->d_sort.ptr) + key_bytes);
    CUDA_TRY(cudaMemsetAsync(hist, 0, hist_bytes, s));
    ClassifyParams p;
    memset(&p, 0, sizeof p);
    fill_grid_view(k, 0, p.grid);
    p.n_atoms = k->n_atoms;
    p.n_particles = n_particles;
    p.total = total;
    p.pos = d_pos;
    p.particles = k->d_particles;
    const unsigned blocks = (unsigned) ((total + 255) / 256);
    gf_sort_count_kernel<<<blocks, 256, 0, s>>>(p, shift, n_bins, keys, hist);
    gf_sort_scan_kernel<<<1, 1024, 0, s>>>(hist, n_bins);
    gf_sort_scatter_kernel<<<blocks, 256, 0, s>>>(keys, hist, total, d_order);
    g_launches += 3;
    CUDA_TRY(cudaGetLastError());
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

// Host<->device copy bandwidth of THIS process's pinned memory on this GPU's PCIe link (bench.py prints it beside the
// end-to-end figure; under torchrun every rank runs it at the same time, so the figure is what the shared host path
// gives each GPU). gbs[0] = H2D alone, gbs[1] = D2H alone, gbs[2] = both directions at once (sum of the two).
int gfb_bench_host_copy(gfb_device* dev, size_t bytes, int reps, double gbs[3]) {
    if (!dev || !gbs || bytes < 4096 || reps < 1) return fail(GFB_ERR_INVALID, "gfb_bench_host_copy: bad argument");
    CUDA_TRY(cudaSetDevice(dev->ordinal));
    void *h_in = nullptr, *h_out = nullptr, *d_a = nullptr, *d_b = nullptr;
    cudaEvent_t e0, e1, e2;
    cudaError_t err = cudaHostAlloc(&h_in, bytes, cudaHostAllocDefault);
    if (err == cudaSuccess) err = cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault);
    if (err == cudaSuccess) err = cudaMalloc(&d_a, bytes);
    if (err == cudaSuccess) err = cudaMalloc(&d_b, bytes);
    if (err == cudaSuccess) {
        memset(h_in, 1, bytes);
        memset(h_out, 0, bytes);
        err = cudaMemset(d_b, 0, bytes);
    }
    if (err != cudaSuccess) {
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
        if (d_a) cudaFree(d_a);
        if (d_b) cudaFree(d_b);
        return fail(GFB_ERR_CUDA, "gfb_bench_host_copy: %s", cudaGetErrorString(err));
    }
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    float ms = 0.f;
    for (int mode = 0; mode < 3; mode++) {
        for (int i = -1; i < reps; i++) {   // i == -1: warm-up
            if (i == 0) {
                cudaStreamSynchronize(dev->stream);
                cudaStreamSynchronize(dev->copy_stream);
                cudaEventRecord(e0, dev->stream);
            }
            if (mode != 1) cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, dev->stream);
            if (mode != 0) cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, mode == 2 ? dev->copy_stream : dev->stream);
        }
        if (mode == 2) {   // the clock stops when both directions are done
            cudaEventRecord(e2, dev->copy_stream);
            cudaStreamWaitEvent(dev->stream, e2, 0);
        }
        cudaEventRecord(e1, dev->stream);
        err = cudaStreamSynchronize(dev->stream);
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
        if (err != cudaSuccess) break;
        gbs[mode] = (double) bytes * reps * (mode == 2 ? 2.0 : 1.0) / (ms * 1e-3) / 1e9;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    cudaFreeHost(h_in);
    cudaFreeHost(h_out);
    cudaFree(d_a);
    cudaFree(d_b);
    if (err != cudaSuccess) return fail(GFB_ERR_CUDA, "gfb_bench_host_copy: %s", cudaGetErrorString(err));
    return GFB_OK;
}

}  // extern "C"
