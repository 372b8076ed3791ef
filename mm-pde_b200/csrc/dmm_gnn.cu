// Message-passing layer of the DMM mesh mover's graph branch (/root/reference/mesh/dmm_model.py:94-142), forward only:
// the mover is frozen (its parameters are not in the optimizer, /root/reference/mmpde.py:269-271) and the moved mesh is
// used detached, so no backward exists.  Hidden width 4: per edge an 11 -> 4 -> 4 tanh MLP, mean over the incoming edges,
// then an 8 -> 4 -> 4 tanh MLP per node and the residual.  In PyTorch this materialises [E,11], [E,4], [E,4] tensors and a
// scatter per layer (E = 1.4 M for 16 x 2521 nodes) -- ~2 ms per layer of launch- and bandwidth-bound glue; fused it is
// one pass over the edge list: one warp per target node, one lane per incoming edge.
//   out[i] = x[i] + tanh(W4 tanh(W3 [x[i] | mean_j m_ij] + b3) + b4)        (BatchNorm stays in the caller)
//   m_ij   = tanh(W2 tanh(W1 [x_i | x_j | u_i-u_j | px_i-px_j | py_i-py_j] + b1) + b2)
#include "common.cuh"

namespace mmpde {

struct DmmGnnWeights {          // nn.Linear layouts, row-major [out][in]
    float w1[4][11], b1[4], w2[4][4], b2[4], w3[4][8], b3[4], w4[4][4], b4[4];
};

__global__ void __launch_bounds__(256) dmm_gnn_layer_kernel(const float4* __restrict__ x, const float4* __restrict__ upos,
                                                             const int* __restrict__ row_ptr, const int* __restrict__ src,
                                                             int64_t n_nodes, const float* __restrict__ weights,
                                                             float4* __restrict__ out) {
    __shared__ DmmGnnWeights W;                                          // 124 floats, read as warp-wide broadcasts
    for (int k = threadIdx.x; k < MMPDE_DMM_GNN_NPARAM; k += blockDim.x) reinterpret_cast<float*>(&W)[k] = __ldg(weights + k);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_nodes) return;
    const float4 xi = __ldg(x + i), si = __ldg(upos + i);              // upos = (u, px, py, unused)
    const int e0 = __ldg(row_ptr + i), e1 = __ldg(row_ptr + i + 1);
    // the target's own contribution to message_net_1 is the same for all of its edges
    float base[4];
#pragma unroll
    for (int o = 0; o < 4; ++o)
        base[o] = W.b1[o] + W.w1[o][0] * xi.x + W.w1[o][1] * xi.y + W.w1[o][2] * xi.z + W.w1[o][3] * xi.w +
                  W.w1[o][8] * si.x + W.w1[o][9] * si.y + W.w1[o][10] * si.z;
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    for (int e = e0 + lane; e < e1; e += 32) {
        const int j = __ldg(src + e);
        const float4 xj = __ldg(x + j), sj = __ldg(upos + j);
        float h[4];
#pragma unroll
        for (int o = 0; o < 4; ++o)
            h[o] = tanhf(base[o] + W.w1[o][4] * xj.x + W.w1[o][5] * xj.y + W.w1[o][6] * xj.z + W.w1[o][7] * xj.w -
                         W.w1[o][8] * sj.x - W.w1[o][9] * sj.y - W.w1[o][10] * sj.z);
#pragma unroll
        for (int o = 0; o < 4; ++o)
            m[o] += tanhf(W.b2[o] + W.w2[o][0] * h[0] + W.w2[o][1] * h[1] + W.w2[o][2] * h[2] + W.w2[o][3] * h[3]);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) m[o] = warp_sum(m[o]);
    if (lane != 0) return;
    const float inv = 1.f / (float)max(e1 - e0, 1);
    const float in8[8] = {xi.x, xi.y, xi.z, xi.w, m[0] * inv, m[1] * inv, m[2] * inv, m[3] * inv};
    float h[4], r[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float a = W.b3[o];
#pragma unroll
        for (int k = 0; k < 8; ++k) a += W.w3[o][k] * in8[k];
        h[o] = tanhf(a);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) r[o] = tanhf(W.b4[o] + W.w4[o][0] * h[0] + W.w4[o][1] * h[1] + W.w4[o][2] * h[2] + W.w4[o][3] * h[3]);
    out[i] = make_float4(xi.x + r[0], xi.y + r[1], xi.z + r[2], xi.w + r[3]);
}

}  // namespace mmpde

using namespace mmpde;

extern "C" int mmpde_dmm_gnn_layer(const float* x, const float* upos, const int32_t* row_ptr, const int32_t* edge_src,
                                   int64_t n_nodes, const float* weights, float* out, void* stream) {
    if (n_nodes < 0 || weights == nullptr) return MMPDE_EINVAL;
    if (n_nodes == 0) return MMPDE_OK;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(upos) | reinterpret_cast<uintptr_t>(out)) & 15) return MMPDE_EINVAL;
    static_assert(sizeof(DmmGnnWeights) == MMPDE_DMM_GNN_NPARAM * sizeof(float), "packed parameter count");
    const int warps = 8;
    dmm_gnn_layer_kernel<<<(unsigned)((n_nodes + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(upos), row_ptr, edge_src, n_nodes, weights,
        reinterpret_cast<float4*>(out));
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
