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


// ================================================================================================================
// Mesh displacement grad_xi phi(u, xi) of the DMM mover (the reference: two autograd.grad(create_graph=True) calls,
// /root/reference/data_creator_2d.py:106-107) as ONE pass -- forward-mode Jacobian of the two-layer tanh trunk and the
// two-layer tanh out_nn (mesh/dmm_model.py:145-219 of the reference; formulas in DMM.displacement):
//     a   = tanh(W1 xi + b1)                       [K <= 32]
//     z_j = cst[sample][j] + sum_k a_k M[j][k]     [J],   h = tanh(z),  gate_j = (1 - h_j^2) w_j
//     out[n][d] = sum_j gate_j * sum_k (1 - a_k^2) W1[k][d] M[j][k]
// In tensor ops this is a [3N,32] x [32,512] fp32 product on the CUDA cores plus eight [N,512] elementwise / reduction
// passes (~430 us of the moved-mesh branch's critical path at N = 36 864); here the [N,512] intermediates never exist.
// Eight points per warp, four lanes per point (lane = quarter * 8 + point): every 8-lane phase of a 128-bit shared-memory
// load reads ONE row of M (broadcast, conflict-free); the four quarters of j are summed with two shuffles.
// ================================================================================================================
constexpr int DISP_K = 32;                                               // trunk width, zero-padded
__global__ void __launch_bounds__(256, 1) dmm_displacement_kernel(const float2* __restrict__ xi, const float* __restrict__ W1,
                                                                const float* __restrict__ b1, int K, const float* __restrict__ M,
                                                                const float* __restrict__ cst, const float* __restrict__ w, int J,
                                                                int64_t N, int64_t per_sample, float2* __restrict__ out) {
    extern __shared__ float s_m[];                                       // M [J][32] (zero-padded rows) | w [J]
    float* s_w = s_m + (size_t)J * DISP_K;
    if (K == DISP_K && (reinterpret_cast<uintptr_t>(M) & 15) == 0) {     // rows already 32 wide: a straight 128-bit copy
#pragma unroll 4
        for (int i = threadIdx.x; i < J * (DISP_K / 4); i += blockDim.x) reinterpret_cast<float4*>(s_m)[i] = ldg4(M + 4 * (size_t)i);
    } else {
        for (int i = threadIdx.x; i < J * DISP_K; i += blockDim.x) {
            const int j = i / DISP_K, k = i % DISP_K;
            s_m[i] = (k < K) ? __ldg(M + (size_t)j * K + k) : 0.f;
        }
    }
    for (int j = threadIdx.x; j < J; j += blockDim.x) s_w[j] = __ldg(w + j);
    __syncthreads();
    const int lane = threadIdx.x & 31, pt = lane & 7, quarter = lane >> 3;
    const int jq = J / 4;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t base = warp0 * 8; base < N; base += n_warps * 8) {
        const int64_t n = base + pt;
        const bool live = n < N;
        const float2 x = live ? __ldg(xi + n) : make_float2(0.f, 0.f);
        const float* c_row = cst + (live ? n / per_sample : 0) * J;
        float a[DISP_K], u[DISP_K], v[DISP_K];
#pragma unroll
        for (int k = 0; k < DISP_K; ++k) {
            float wx = 0.f, wy = 0.f, bb = 0.f;
            if (k < K) { wx = __ldg(W1 + 2 * k); wy = __ldg(W1 + 2 * k + 1); bb = __ldg(b1 + k); }
            a[k] = tanhf(fmaf(wx, x.x, fmaf(wy, x.y, bb)));
            const float da = 1.f - a[k] * a[k];
            u[k] = da * wx; v[k] = da * wy;
        }
        float acc1 = 0.f, acc2 = 0.f;
#pragma unroll 4
        for (int jj = 0; jj < jq; ++jj) {
            const int j = quarter * jq + jj;
            const float4* row = reinterpret_cast<const float4*>(s_m + (size_t)j * DISP_K);
            float z = __ldg(c_row + j), s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int q = 0; q < DISP_K / 4; ++q) {
                const float4 m = row[q];
                z = fmaf(a[4 * q], m.x, z); z = fmaf(a[4 * q + 1], m.y, z); z = fmaf(a[4 * q + 2], m.z, z); z = fmaf(a[4 * q + 3], m.w, z);
                s1 = fmaf(u[4 * q], m.x, s1); s1 = fmaf(u[4 * q + 1], m.y, s1); s1 = fmaf(u[4 * q + 2], m.z, s1); s1 = fmaf(u[4 * q + 3], m.w, s1);
                s2 = fmaf(v[4 * q], m.x, s2); s2 = fmaf(v[4 * q + 1], m.y, s2); s2 = fmaf(v[4 * q + 2], m.z, s2); s2 = fmaf(v[4 * q + 3], m.w, s2);
            }
            const float h = tanhf(z);
            const float gate = (1.f - h * h) * s_w[j];
            acc1 = fmaf(gate, s1, acc1);
            acc2 = fmaf(gate, s2, acc2);
        }
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 8);  acc2 += __shfl_xor_sync(0xffffffffu, acc2, 8);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 16); acc2 += __shfl_xor_sync(0xffffffffu, acc2, 16);
        if (live && quarter == 0) out[n] = make_float2(acc1, acc2);
    }
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

extern "C" int mmpde_dmm_displacement(const float* xi, const float* W1, const float* b1, int K, const float* M, const float* cst,
                                      const float* w, int J, int64_t n_points, int64_t per_sample, float* out, void* stream) {
    if (n_points < 0 || per_sample <= 0 || K < 1 || K > DISP_K || J < 4 || J % 4 || J > 1024 || !xi || !W1 || !b1 || !M || !cst || !w || !out)
        return MMPDE_EINVAL;
    if (n_points == 0) return MMPDE_OK;
    if ((reinterpret_cast<uintptr_t>(xi) | reinterpret_cast<uintptr_t>(out)) & 7) return MMPDE_EINVAL;
    const size_t smem = ((size_t)J * DISP_K + J) * sizeof(float);
    MMPDE_ENSURE_SMEM(dmm_displacement_kernel, 1024 * (DISP_K + 1) * sizeof(float));
    const int64_t warps = (n_points + 7) / 8;
    const int grid = (int)imin64((warps + 7) / 8, sm_count());     // 158 registers x 256 threads: one CTA per SM, M staged once
    dmm_displacement_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(xi), W1, b1, K, M, cst, w, J,
                                                                       n_points, per_sample, reinterpret_cast<float2*>(out));
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
