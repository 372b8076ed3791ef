// Device side of the cross-GPU BatchNorm-sum exchange over NVLink peer memory (protocol: see peer.cu).  Shared by the
// stand-alone exchange kernel (peer.cu) and the reducing BatchNorm kernels whose last CTA exchanges in place (norm.cu).
#pragma once
#include "common.cuh"
#include <cstdio>

namespace mmpde {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_FLAGS_OFF = 256, PEER_SLOTS_OFF = 1024;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {          // never from a stale L1 line
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}


// Exchange number counter+1 of this rank's buffer, in two halves; all 256 threads of the calling CTA take part.
// POST: thread c stores v (this rank's c-th sum) into the slot of every peer, then the flags are raised.  The counter is
// not advanced: the matching WAIT (same kernel or a later one on the same stream, with no other exchange of this buffer
// in between) derives the same sequence number from it.
__device__ __forceinline__ void peer_post_256(double v, const int64_t* __restrict__ peer_base, int rank, int world) {
    __shared__ uint32_t s_seq_post;
    const int c = threadIdx.x;
    unsigned char* mine = reinterpret_cast<unsigned char*>(peer_base[rank]);
    __syncthreads();
    if (c == 0) s_seq_post = *reinterpret_cast<volatile uint32_t*>(mine) + 1u;
    __syncthreads();
    const uint32_t seq = s_seq_post, slot = seq & 3u;
    for (int r = 0; r < world; ++r) {
        double* dst = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(peer_base[r]) + PEER_SLOTS_OFF) +
                      ((size_t)slot * PEER_MAX_WORLD + rank) * 256 + c;
        *reinterpret_cast<volatile double*>(dst) = v;
    }
    __threadfence_system();
    __syncthreads();
    if (c < world)
        st_release_sys(reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(peer_base[c]) + PEER_FLAGS_OFF) +
                           slot * PEER_MAX_WORLD + rank, seq);
}
// WAIT: until every rank's sums of this exchange have arrived; returns the c-th sum over all ranks, added in rank order
// (identical bits everywhere), and advances the counter.
__device__ __forceinline__ double peer_wait_256(const int64_t* __restrict__ peer_base, int rank, int world,
                                                unsigned long long timeout_ns) {
    __shared__ uint32_t s_seq_wait;
    const int c = threadIdx.x;
    unsigned char* mine = reinterpret_cast<unsigned char*>(peer_base[rank]);
    __syncthreads();
    if (c == 0) s_seq_wait = *reinterpret_cast<volatile uint32_t*>(mine) + 1u;
    __syncthreads();
    const uint32_t seq = s_seq_wait, slot = seq & 3u;
    if (c < world) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine + PEER_FLAGS_OFF) + slot * PEER_MAX_WORLD + c;
        // A peer may legitimately be late by many seconds (rank-0-only checkpoint save, a re-recorded step graph, a
        // data-loader stall), so the wait is bounded in WALL time (globaltimer, independent of the SM clock) by a
        // generous, configurable limit (default 10 min, like a collective watchdog): only a dead peer trips it.
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(flag) != seq) {
            if ((++spins & 1023u) == 0u && global_ns() - t0 > timeout_ns) {
                printf("mmpde_bn_exchange: rank %d timed out waiting for rank %d (exchange %u)\n", rank, c, seq);
                __trap();
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    const double* slots = reinterpret_cast<const double*>(mine + PEER_SLOTS_OFF) + (size_t)slot * PEER_MAX_WORLD * 256 + c;
    double acc = 0.0;
    for (int r = 0; r < world; ++r) acc += ld_relaxed_sys_f64(slots + r * 256);
    __syncthreads();                                   // every thread has read the counter-derived slot before it moves on
    if (c == 0) *reinterpret_cast<volatile uint32_t*>(mine) = seq;
    return acc;
}
__device__ __forceinline__ double peer_exchange_256(double v, const int64_t* __restrict__ peer_base, int rank, int world,
                                                    unsigned long long timeout_ns) {
    peer_post_256(v, peer_base, rank, world);
    return peer_wait_256(peer_base, rank, world, timeout_ns);
}

// wall-clock limit of the flag wait (seconds -> ns); env MMPDE_PEER_TIMEOUT_S or mmpde_bn_exchange_set_timeout()
unsigned long long peer_timeout_ns();

}  // namespace mmpde
