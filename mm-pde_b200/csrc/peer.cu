// Cross-GPU sum of the BatchNorm column sums over NVLink peer memory, in ONE kernel (sync-BatchNorm of the batch-sharded
// and graph-partitioned paths; the reference is single-device).  Replaces an NCCL all-reduce of 2 KB,
// whose cost is pure launch / protocol latency and which a training step pays 32 times, one after the other.
//
// Every rank owns one exchange buffer that all peers can address (peer-mapped by the caller, e.g. torch symmetric
// memory):
//     +0    uint32 counter        exchanges completed on this rank (its own, never written by peers)
//     +256  uint32 flag[4][16]    flag[s][r] = sequence number of the last exchange rank r delivered into slot s
//     +1024 double slot[4][16][256]
// Exchange number q (counter + 1, identical on all ranks because all ranks run the same sequence of exchanges):
//   1. fold the local accumulator copies and STORE the 256 sums into slot[q & 3][rank] of every peer (remote stores),
//   2. fence, then raise flag[q & 3][rank] = q on every peer,
//   3. wait until the local flag[q & 3][r] == q for all r, and add the slots in rank order (every rank gets the same
//      bits).  A rank can run at most one exchange ahead of a peer (it needs the peer's flag to finish), so four slots
//      are never overwritten while still being read.
// The sequence number lives on the device, which makes the kernel replayable from a CUDA graph.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>

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

__global__ void __launch_bounds__(256) bn_exchange_kernel(const double* __restrict__ sums, int n_rep,
                                                          const int64_t* __restrict__ peer_base, int rank, int world,
                                                          double* __restrict__ out, unsigned long long timeout_ns) {
    __shared__ uint32_t s_seq;
    const int c = threadIdx.x;
    unsigned char* mine = reinterpret_cast<unsigned char*>(peer_base[rank]);
    if (c == 0) s_seq = *reinterpret_cast<volatile uint32_t*>(mine) + 1u;
    double v = 0.0;
    for (int r = 0; r < n_rep; ++r) v += sums[r * 256 + c];
    __syncthreads();
    const uint32_t seq = s_seq, slot = seq & 3u;
    for (int r = 0; r < world; ++r) {
        double* dst = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(peer_base[r]) + PEER_SLOTS_OFF) +
                      ((size_t)slot * PEER_MAX_WORLD + rank) * 256 + c;
        *reinterpret_cast<volatile double*>(dst) = v;
    }
    __threadfence_system();
    __syncthreads();
    if (c < world) {
        st_release_sys(reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(peer_base[c]) + PEER_FLAGS_OFF) +
                           slot * PEER_MAX_WORLD + rank, seq);
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
    out[c] = acc;
    if (c == 0) *reinterpret_cast<volatile uint32_t*>(mine) = seq;
}

}  // namespace mmpde

using namespace mmpde;

static_assert(MMPDE_BN_EXCHANGE_BYTES == PEER_SLOTS_OFF + 4 * PEER_MAX_WORLD * 256 * sizeof(double), "header and kernel layout differ");

static double g_peer_timeout_s = -1.0;

extern "C" int mmpde_bn_exchange_set_timeout(double seconds) {
    if (!(seconds > 0.0)) return MMPDE_EINVAL;
    g_peer_timeout_s = seconds;
    return MMPDE_OK;
}

extern "C" int mmpde_bn_exchange(const double* sums, int n_rep, const int64_t* peer_base, int rank, int world, double* out,
                                 void* stream) {
    if (!sums || !peer_base || !out || n_rep < 1 || world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world)
        return MMPDE_EINVAL;
    if (g_peer_timeout_s < 0.0) {
        const char* e = getenv("MMPDE_PEER_TIMEOUT_S");
        g_peer_timeout_s = (e && atof(e) > 0.0) ? atof(e) : 600.0;
    }
    bn_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sums, n_rep, peer_base, rank, world, out,
                                                            (unsigned long long)(g_peer_timeout_s * 1e9));
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
