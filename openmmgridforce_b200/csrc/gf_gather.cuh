// Producer side of the energy gather of a replica-sharded multi-GPU run: peer stores over NVLink + arrival flags.
// Used fused into the evaluation kernels (gf_eval_lines.cuh, gf_eval_lines_f64.cuh: gather_ticket / gather_copy at the
// end of the launch) and stand-alone (gf_gather_push_kernel in gf_multi.cu). The consumer is gf_gather_wait_kernel.
#ifndef GF_GATHER_CUH_
#define GF_GATHER_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "gf_params.h"

namespace gfb {

// Copier `me` of K copies its slice of src[0..n) into every peer's gathered array (plain 16-byte stores to peer-mapped
// memory, local memory for itself; 4 independent loads in flight per thread: one dependent load->store chain per thread
// is latency-bound), fences at system scope and takes a copy ticket; the last copier publishes the gather's sequence
// number in this rank's flag slot on every peer and advances the device-resident counters.
template <int BLOCK>
__device__ __forceinline__ void gather_publish(GatherTable* gt, const double* src, int n, long long offset, unsigned me, unsigned K) {
    __shared__ unsigned s_done;
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(&gt->issued) + 1ull;   // this gather's number
    const int parity = (int) (seq & 1ull);
    const long long base = (long long) parity * gt->count_total + offset;
    const int np = gt->n_peers;
    const bool vec = ((base | (long long) n) & 1) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    if (vec) {
        const int n2 = n / 2;
        const int lo = (int) ((long long) n2 * me / K), hi = (int) ((long long) n2 * (me + 1) / K);
        const double2* src2 = reinterpret_cast<const double2*>(src);
        for (int i0 = lo + (int) threadIdx.x; i0 < hi; i0 += 4 * BLOCK) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i0 + u * BLOCK < hi) v[u] = __ldcg(src2 + i0 + u * BLOCK);
            for (int r = 0; r < np; r++) {
                double2* dst2 = reinterpret_cast<double2*>(gt->peer_data[(gt->my_rank + 1 + r) % np] + base);   // staggered start
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (i0 + u * BLOCK < hi) dst2[i0 + u * BLOCK] = v[u];
            }
        }
    } else {
        const int lo = (int) ((long long) n * me / K), hi = (int) ((long long) n * (me + 1) / K);
        for (int i = lo + (int) threadIdx.x; i < hi; i += BLOCK) {
            const double v = __ldcg(src + i);
            for (int r = 0; r < np; r++) gt->peer_data[(gt->my_rank + 1 + r) % np][base + i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_done = atomicAdd(&gt->copy_ticket, 1u);
    __syncthreads();
    if (s_done != K - 1) return;
    // last copier: every copier's stores were fenced at system scope before its copy ticket
    __threadfence();
    if ((int) threadIdx.x < np) {
        unsigned long long* flag = gt->peer_flags[threadIdx.x] + parity * kMaxPeers + gt->my_rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
    }
    if (threadIdx.x == 0) {
        gt->issued = seq;
        gt->ticket = 0;
        gt->copy_ticket = 0;
    }
}

// ---- fused into an evaluation kernel -----------------------------------------------------------------------------------
// After its energy atomics (and before its force writes, so that the fence below has only a handful of L2-resident
// atomics to wait for) every block passes gather_ticket: one barrier, then ONE thread fences — the barrier makes the fence
// cumulative over the block's atomics — and bumps the launch's block counter with a non-returning RED (24,064 returning
// atomics on one address serialise into tens of microseconds; REDs pipeline). The LAST kGatherCopiers blocks of the grid
// (highest block indices: dispatched last, so everything they wait for is already running) then wait in gather_copy
// until the counter says every block has passed — all energy atomics performed — and publish. No NCCL launch, no extra
// kernel on the producing side.
constexpr unsigned kGatherCopiers = 32;
template <int BLOCK>
__device__ __forceinline__ int gather_ticket(const EvalParams& p) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(&p.gather->ticket) : "memory");
    }
    const unsigned G = gridDim.x, K = G < kGatherCopiers ? G : kGatherCopiers;
    return blockIdx.x < G - K ? -1 : (int) (blockIdx.x - (G - K));
}

template <int BLOCK>
__device__ __forceinline__ void gather_copy(const EvalParams& p, unsigned me) {
    GatherTable* const gt = p.gather;
    const unsigned G = gridDim.x, K = G < kGatherCopiers ? G : kGatherCopiers;
    if (threadIdx.x == 0) {
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&gt->ticket) : "memory");
            if (seen < G) __nanosleep(40);
        } while (seen < G);
    }
    __syncthreads();
    __threadfence();
    gather_publish<BLOCK>(gt, p.energies, p.n_replicas * p.n_slots, p.gather_offset, me, K);
}

}  // namespace gfb
#endif
