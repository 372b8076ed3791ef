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
#include "peer.cuh"
#include <cstdlib>

namespace mmpde {

__global__ void __launch_bounds__(256) bn_exchange_kernel(const double* __restrict__ sums, int n_rep,
                                                          const int64_t* __restrict__ peer_base, int rank, int world,
                                                          double* __restrict__ out, unsigned long long timeout_ns) {
    const int c = threadIdx.x;
    double v = 0.0;
    for (int r = 0; r < n_rep; ++r) v += sums[r * 256 + c];
    out[c] = peer_exchange_256(v, peer_base, rank, world, timeout_ns);
}

// second half of an exchange whose first half ran in the last CTA of an earlier kernel (mmpde_bn_bwd_reduce_post)
__global__ void __launch_bounds__(256) bn_exchange_wait_kernel(const int64_t* __restrict__ peer_base, int rank, int world,
                                                               double* __restrict__ out, unsigned long long timeout_ns) {
    out[threadIdx.x] = peer_wait_256(peer_base, rank, world, timeout_ns);
}

}  // namespace mmpde

using namespace mmpde;

static_assert(MMPDE_BN_EXCHANGE_BYTES == PEER_SLOTS_OFF + 4 * PEER_MAX_WORLD * 256 * sizeof(double), "header and kernel layout differ");

static double g_peer_timeout_s = -1.0;

unsigned long long mmpde::peer_timeout_ns() {
    if (g_peer_timeout_s < 0.0) {
        const char* e = getenv("MMPDE_PEER_TIMEOUT_S");
        g_peer_timeout_s = (e && atof(e) > 0.0) ? atof(e) : 600.0;
    }
    return (unsigned long long)(g_peer_timeout_s * 1e9);
}

extern "C" int mmpde_bn_exchange_set_timeout(double seconds) {
    if (!(seconds > 0.0)) return MMPDE_EINVAL;
    g_peer_timeout_s = seconds;
    return MMPDE_OK;
}

extern "C" int mmpde_bn_exchange(const double* sums, int n_rep, const int64_t* peer_base, int rank, int world, double* out,
                                 void* stream) {
    if (!sums || !peer_base || !out || n_rep < 1 || world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world)
        return MMPDE_EINVAL;
    bn_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sums, n_rep, peer_base, rank, world, out, peer_timeout_ns());
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_bn_exchange_wait(const int64_t* peer_base, int rank, int world, double* out, void* stream) {
    if (!peer_base || !out || world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world) return MMPDE_EINVAL;
    bn_exchange_wait_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(peer_base, rank, world, out, peer_timeout_ns());
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
