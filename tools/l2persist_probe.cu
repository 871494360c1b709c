// Probe: can a ~74 MB grid be held in B200's L2 with a persisting access-policy window while 72 MB of positions/forces
// stream through? Gathers 4 random 32-byte sectors per thread from `grid` and streams `stream_bytes` from another buffer.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld32(const float* p, float v[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__global__ void k(const float* grid, unsigned long long n_sectors, const double2* stream, size_t n_stream, double* sink, int phase) {
    size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long h = (t + 1) * 0x9E3779B97F4A7C15ull + phase * 0x1234567ull;
    float acc = 0.f;
    // stream part: each thread reads 72 bytes-ish (position + force traffic per atom)
    for (int i = 0; i < 4; i++) {
        size_t idx = t * 4 + i;
        if (idx < n_stream) { double2 v; asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(stream + idx)); acc += (float) v.x; }
    }
    for (int i = 0; i < 4; i++) {
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        float v[8]; ld32(grid + 8 * __umul64hi(h, n_sectors), v); acc += v[0] + v[7];
    }
    if (acc == 123.f) sink[0] = acc;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("L2 %d MB, persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB\n", p.l2CacheSize >> 20, p.persistingL2CacheMaxSize >> 20, p.accessPolicyMaxWindowSize >> 20);
    const size_t n_atoms = 1 << 20;
    for (size_t grid_mb : {48, 64, 74, 96}) {
        size_t gb = grid_mb << 20; float* grid; cudaMalloc(&grid, gb); cudaMemset(grid, 0, gb);
        size_t sb = (size_t) 8 * n_atoms * 64; double2* st; cudaMalloc(&st, sb); cudaMemset(st, 0, sb);   // 8 rotating 64 MB stream sets
        double* sink; cudaMalloc(&sink, 8);
        cudaStream_t s; cudaStreamCreate(&s);
        for (int persist = 0; persist < 2; persist++) {
            if (persist) {
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, p.persistingL2CacheMaxSize);
                cudaStreamAttrValue a = {};
                a.accessPolicyWindow.base_ptr = grid; a.accessPolicyWindow.num_bytes = gb < (size_t) p.accessPolicyMaxWindowSize ? gb : p.accessPolicyMaxWindowSize;
                a.accessPolicyWindow.hitRatio = 1.0f; a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cudaError_t e = cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &a);
                if (e != cudaSuccess) printf("policy error %s\n", cudaGetErrorString(e));
            }
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int i = 0; i < 8; i++) k<<<n_atoms / 256, 256, 0, s>>>(grid, gb / 32, st + (size_t) (i % 8) * n_atoms * 4, n_atoms * 4, sink, i);
            cudaEventRecord(e0, s);
            for (int i = 0; i < 40; i++) k<<<n_atoms / 256, 256, 0, s>>>(grid, gb / 32, st + (size_t) (i % 8) * n_atoms * 4, n_atoms * 4, sink, i);
            cudaEventRecord(e1, s); cudaStreamSynchronize(s);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("grid %3zu MB persist=%d: %.2f us per launch (1M threads x (4 gathers + 64 B stream))\n", grid_mb, persist, ms / 40 * 1e3);
        }
        cudaCtxResetPersistingL2Cache();
        cudaFree(grid); cudaFree(st); cudaFree(sink); cudaStreamDestroy(s);
    }
    return 0;
}
