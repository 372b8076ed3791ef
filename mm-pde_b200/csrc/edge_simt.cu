// Message passing over the target-sorted edge list, fp32 SIMT reference kernels.
// Replaces PyG propagate + message_net_1/2 + scatter-mean (/root/reference/gnn_2d.py:55,59-63) and their
// autograd, using the algebraic split z1 = P[i] + Q[j] + W1c*e_ij.
// One CTA works on tiles of 128 consecutive edges; a tile may start/end inside a target's segment, so
// per-target results are added atomically (each (target,channel) sees at most a few adds).
#include "common.cuh"

namespace mmpde {

constexpr int ET = 128;        // edges per tile
constexpr int ELD = 132;       // padded row length of the [edge][channel] staging tiles

struct EdgeArgs {
    const float* PQ; const float4* node4; const int* src; const int* dst; const float* inv_deg;
    int64_t n_edges; const float* w1c; const float* w2; const float* b2;
    float* agg; int64_t ld_agg; uint32_t* mask2;
    // backward only
    const float* g_agg; int64_t ld_gagg; float* dPQ; float* dW2; float* db2; float* dW1c; float* g_u; int64_t g_u_stride;
};

// h1[e][c] = relu(P[dst][c] + Q[src][c] + W1c[c].e_ij) for the tile's rows -> sH; also sE[e] = e_ij, sDst/sSrc.
// warp w builds rows w*16..w*16+15, lane owns channels 4*lane..4*lane+3.
__device__ __forceinline__ void build_h1(const EdgeArgs& p, int64_t e0, int rows, float* sH, float4* sE, int* sDst,
                                         int* sSrc, const float (&w1c)[4][4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp * 16; r < warp * 16 + 16; ++r) {
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) {
            int i = __ldg(p.dst + e0 + r), j = __ldg(p.src + e0 + r);
            float4 ni = __ldg(p.node4 + i), nj = __ldg(p.node4 + j);
            float4 e = make_float4(ni.x - nj.x, ni.y - nj.y, ni.z - nj.z, ni.w);   // (u_i-u_j, px_i-px_j, py_i-py_j, v_i)
            float4 P = ldg4(p.PQ + (int64_t)i * 256 + lane * 4);
            float4 Q = ldg4(p.PQ + (int64_t)j * 256 + 128 + lane * 4);
            float z[4] = {P.x + Q.x, P.y + Q.y, P.z + Q.z, P.w + Q.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                z[c] = fmaf(w1c[c][0], e.x, z[c]);
                z[c] = fmaf(w1c[c][1], e.y, z[c]);
                z[c] = fmaf(w1c[c][2], e.z, z[c]);
                z[c] = fmaf(w1c[c][3], e.w, z[c]);
            }
            h = make_float4(fmaxf(z[0], 0.f), fmaxf(z[1], 0.f), fmaxf(z[2], 0.f), fmaxf(z[3], 0.f));
            if (lane == 0) { sE[r] = e; sDst[r] = i; sSrc[r] = j; }
        } else if (lane == 0) {
            sE[r] = make_float4(0.f, 0.f, 0.f, 0.f); sDst[r] = -1; sSrc[r] = -1;
        }
        *reinterpret_cast<float4*>(sH + r * ELD + lane * 4) = h;
    }
}

// acc[i][j] = sum_k sA[row_i][k] * sB[k][col_j]; rows {ty*4+i, 64+ty*4+i}, cols {tx*4+j, 64+tx*4+j}
__device__ __forceinline__ void tile_mm_rowk(const float* sA, int lda, const float* sB, int ldb, int ty, int tx,
                                             float (&acc)[8][8]) {
#pragma unroll 2
    for (int k = 0; k < 128; k += 4) {
        float a[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int row = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
            float4 v = *reinterpret_cast<const float4*>(sA + row * lda + k);
            a[i][0] = v.x; a[i][1] = v.y; a[i][2] = v.z; a[i][3] = v.w;
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float4 b0 = *reinterpret_cast<const float4*>(sB + (k + kk) * ldb + tx * 4);
            float4 b1 = *reinterpret_cast<const float4*>(sB + (k + kk) * ldb + 64 + tx * 4);
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i][kk], b[j], acc[i][j]);
        }
    }
}

// acc[i][j] += sum_e sA[e][row_i] * sB[e][col_j]   (both operands indexed [e][.], contraction over tile rows)
__device__ __forceinline__ void tile_mm_tn(const float* sA, const float* sB, int ld, int ty, int tx, float (&acc)[8][8]) {
#pragma unroll 4
    for (int e = 0; e < ET; ++e) {
        float4 a0 = *reinterpret_cast<const float4*>(sA + e * ld + ty * 4);
        float4 a1 = *reinterpret_cast<const float4*>(sA + e * ld + 64 + ty * 4);
        float4 b0 = *reinterpret_cast<const float4*>(sB + e * ld + tx * 4);
        float4 b1 = *reinterpret_cast<const float4*>(sB + e * ld + 64 + tx * 4);
        float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

// Segmented column sums of sT[e][c] over runs of equal sDst, scaled and added to out[dst*ld + c].
// thread (c = tid&127, half = tid>>7) scans rows half*64 .. half*64+63.
__device__ __forceinline__ void segment_add(const float* sT, const int* sDst, int rows, const float* inv_deg, float* out,
                                            int64_t ld) {
    const int c = threadIdx.x & 127, r0 = (threadIdx.x >> 7) * 64;
    int cur = -1;
    float s = 0.f;
    for (int r = r0; r < min(r0 + 64, rows); ++r) {
        int d = sDst[r];
        if (d != cur) {
            if (cur >= 0) atomicAdd(out + (int64_t)cur * ld + c, inv_deg ? s * __ldg(inv_deg + cur) : s);
            cur = d; s = 0.f;
        }
        s += sT[r * ELD + c];
    }
    if (cur >= 0) atomicAdd(out + (int64_t)cur * ld + c, inv_deg ? s * __ldg(inv_deg + cur) : s);
}

__global__ void __launch_bounds__(256, 1) edge_fwd_kernel(EdgeArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sH = reinterpret_cast<float*>(smem_raw);              // [128][132] h1, later messages
    float* sW2t = sH + ET * ELD;                                  // [k=in][o=out] 128x128
    float4* sE = reinterpret_cast<float4*>(sW2t + 128 * 128);     // [128]
    int* sDst = reinterpret_cast<int*>(sE + ET);
    int* sSrc = sDst + ET;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31;

    for (int idx = tid; idx < 128 * 128; idx += 256) {          // W2[o][k] -> sW2t[k][o]
        int o = idx >> 7, k = idx & 127;
        sW2t[k * 128 + o] = __ldg(p.w2 + idx);
    }
    float w1c[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int f = 0; f < 4; ++f) w1c[c][f] = __ldg(p.w1c + (lane * 4 + c) * 4 + f);
    float b2v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b2v[j] = __ldg(p.b2 + ((j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4)));

    const int64_t n_tiles = (p.n_edges + ET - 1) / ET;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t e0 = t * ET;
        const int rows = (int)((p.n_edges - e0 < ET) ? (p.n_edges - e0) : ET);
        __syncthreads();                                          // previous tile's readers are done
        build_h1(p, e0, rows, sH, sE, sDst, sSrc, w1c);
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        tile_mm_rowk(sH, ELD, sW2t, 128, ty, tx, acc);
        __syncthreads();                                          // everyone finished reading h1
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int row = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
            float m[8];
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float z = acc[i][j] + b2v[j];
                m[j] = fmaxf(z, 0.f);
                bits |= (z > 0.f ? 1u : 0u) << j;
            }
            *reinterpret_cast<float4*>(sH + row * ELD + tx * 4) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4*>(sH + row * ELD + 64 + tx * 4) = make_float4(m[4], m[5], m[6], m[7]);
            // mask words: channel c -> word c>>5, bit c&31.  cols tx*4..+3 and 64+tx*4..+3
            if (row < rows) {
                uint32_t lo = (bits & 0xFu) << ((tx * 4) & 31), hi = (bits >> 4) << ((tx * 4) & 31);
                atomicOr(p.mask2 + (e0 + row) * 4 + ((tx * 4) >> 5), lo);
                atomicOr(p.mask2 + (e0 + row) * 4 + 2 + ((tx * 4) >> 5), hi);
            }
        }
        __syncthreads();
        segment_add(sH, sDst, rows, p.inv_deg, p.agg, p.ld_agg);
    }
}

__global__ void __launch_bounds__(256, 1) edge_bwd_kernel(EdgeArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sH = reinterpret_cast<float*>(smem_raw);              // [128][132] h1
    float* sG = sH + ET * ELD;                                    // [128][132] g_z2, later g_z1
    float* sW2 = sG + ET * ELD;                                   // [o][c] 128x128 (as stored)
    float4* sE = reinterpret_cast<float4*>(sW2 + 128 * 128);
    int* sDst = reinterpret_cast<int*>(sE + ET);
    int* sSrc = sDst + ET;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;

    for (int idx = tid; idx < 128 * 128; idx += 256) sW2[idx] = __ldg(p.w2 + idx);
    float w1c[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int f = 0; f < 4; ++f) w1c[c][f] = __ldg(p.w1c + (lane * 4 + c) * 4 + f);

    float dw2[8][8];                                              // dW2[o][c] partial, o rows / c cols
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dw2[i][j] = 0.f;
    float db2_acc = 0.f;                                          // channel tid&127, row half tid>>7
    float dw1c_acc[4] = {0.f, 0.f, 0.f, 0.f};

    const int64_t n_tiles = (p.n_edges + ET - 1) / ET;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t e0 = t * ET;
        const int rows = (int)((p.n_edges - e0 < ET) ? (p.n_edges - e0) : ET);
        __syncthreads();
        build_h1(p, e0, rows, sH, sE, sDst, sSrc, w1c);
        // g_z2[e][o] = g_agg[dst][o] * inv_deg[dst] * [z2 > 0]
        for (int r = warp * 16; r < warp * 16 + 16; ++r) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) {
                int i = __ldg(p.dst + e0 + r);
                float s = __ldg(p.inv_deg + i);
                float4 ga = ldg4(p.g_agg + (int64_t)i * p.ld_gagg + lane * 4);
                uint32_t w = __ldg(p.mask2 + (e0 + r) * 4 + (lane >> 3));
                uint32_t b = w >> ((lane & 7) * 4);
                g.x = (b & 1u) ? ga.x * s : 0.f;
                g.y = (b & 2u) ? ga.y * s : 0.f;
                g.z = (b & 4u) ? ga.z * s : 0.f;
                g.w = (b & 8u) ? ga.w * s : 0.f;
            }
            *reinterpret_cast<float4*>(sG + r * ELD + lane * 4) = g;
        }
        __syncthreads();
        // g_h1 = g_z2 . W2   (contract over o)
        float gh[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) gh[i][j] = 0.f;
        tile_mm_rowk(sG, ELD, sW2, 128, ty, tx, gh);
        // dW2[o][c] += sum_e g_z2[e][o] * h1[e][c]
        tile_mm_tn(sG, sH, ELD, ty, tx, dw2);
        {   // db2[o] += sum_e g_z2[e][o]
            const int c = tid & 127, r0 = (tid >> 7) * 64;
            for (int r = r0; r < r0 + 64; ++r) db2_acc += sG[r * ELD + c];
        }
        __syncthreads();                                          // all reads of g_z2 done
        // g_z1 = g_h1 * [h1 > 0]  -> sG
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int row = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
            float4 h0 = *reinterpret_cast<const float4*>(sH + row * ELD + tx * 4);
            float4 h1 = *reinterpret_cast<const float4*>(sH + row * ELD + 64 + tx * 4);
            float4 o0 = make_float4(h0.x > 0.f ? gh[i][0] : 0.f, h0.y > 0.f ? gh[i][1] : 0.f,
                                    h0.z > 0.f ? gh[i][2] : 0.f, h0.w > 0.f ? gh[i][3] : 0.f);
            float4 o1 = make_float4(h1.x > 0.f ? gh[i][4] : 0.f, h1.y > 0.f ? gh[i][5] : 0.f,
                                    h1.z > 0.f ? gh[i][6] : 0.f, h1.w > 0.f ? gh[i][7] : 0.f);
            *reinterpret_cast<float4*>(sG + row * ELD + tx * 4) = o0;
            *reinterpret_cast<float4*>(sG + row * ELD + 64 + tx * 4) = o1;
        }
        __syncthreads();
        // dP[dst] += segment sums of g_z1
        segment_add(sG, sDst, rows, nullptr, p.dPQ, 256);
        {   // dW1c[c][f] += sum_e g_z1[e][c] * e_f
            const int c = tid & 127, r0 = (tid >> 7) * 64;
            for (int r = r0; r < min(r0 + 64, rows); ++r) {
                float g = sG[r * ELD + c];
                float4 e = sE[r];
                dw1c_acc[0] = fmaf(g, e.x, dw1c_acc[0]);
                dw1c_acc[1] = fmaf(g, e.y, dw1c_acc[1]);
                dw1c_acc[2] = fmaf(g, e.z, dw1c_acc[2]);
                dw1c_acc[3] = fmaf(g, e.w, dw1c_acc[3]);
            }
        }
        // dQ[src] += g_z1 (row scatter);  g_u[dst] += g_z1.W1c[:,0], g_u[src] -= same
        for (int r = warp * 16; r < min(warp * 16 + 16, rows); ++r) {
            float4 g = *reinterpret_cast<const float4*>(sG + r * ELD + lane * 4);
            int j = sSrc[r];
            red_add_v4(p.dPQ + (int64_t)j * 256 + 128 + lane * 4, g);
            if (p.g_u) {
                float part = g.x * w1c[0][0] + g.y * w1c[1][0] + g.z * w1c[2][0] + g.w * w1c[3][0];
                part = warp_sum(part);
                if (lane == 0) {
                    atomicAdd(p.g_u + (int64_t)sDst[r] * p.g_u_stride, part);
                    atomicAdd(p.g_u + (int64_t)j * p.g_u_stride, -part);
                }
            }
        }
    }
    // flush per-CTA partial weight gradients
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int o = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int c = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
            atomicAdd(p.dW2 + o * 128 + c, dw2[i][j]);
        }
    }
    atomicAdd(p.db2 + (tid & 127), db2_acc);
#pragma unroll
    for (int f = 0; f < 4; ++f) atomicAdd(p.dW1c + (tid & 127) * 4 + f, dw1c_acc[f]);
}

constexpr size_t EDGE_FWD_SMEM = sizeof(float) * (ET * ELD + 128 * 128) + sizeof(float4) * ET + 2 * sizeof(int) * ET;
constexpr size_t EDGE_BWD_SMEM = sizeof(float) * (2 * ET * ELD + 128 * 128) + sizeof(float4) * ET + 2 * sizeof(int) * ET;

}  // namespace mmpde

using namespace mmpde;

extern "C" int mmpde_edge_fwd_simt(const float* PQ, const float* node4, const int32_t* edge_src, const int32_t* edge_dst,
                              const float* inv_deg, int64_t n_edges, const float* w1c, const float* w2, const float* b2,
                              float* agg, int64_t ld_agg, uint32_t* mask2, void* stream) {
    if (n_edges < 0 || ld_agg < 128) return MMPDE_EINVAL;
    if (n_edges == 0) return MMPDE_OK;
    auto st = (cudaStream_t)stream;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(edge_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EDGE_FWD_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    cudaMemsetAsync(mask2, 0, sizeof(uint32_t) * 4 * (size_t)n_edges, st);
    EdgeArgs p = {};
    p.PQ = PQ; p.node4 = (const float4*)node4; p.src = edge_src; p.dst = edge_dst; p.inv_deg = inv_deg;
    p.n_edges = n_edges; p.w1c = w1c; p.w2 = w2; p.b2 = b2; p.agg = agg; p.ld_agg = ld_agg; p.mask2 = mask2;
    int64_t n_tiles = (n_edges + ET - 1) / ET;
    int grid = (int)imin64(n_tiles, sm_count());
    edge_fwd_kernel<<<grid, 256, EDGE_FWD_SMEM, st>>>(p);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_edge_bwd_simt(const float* PQ, const float* node4, const int32_t* edge_src, const int32_t* edge_dst,
                              const float* inv_deg, int64_t n_edges, const float* w1c, const float* w2,
                              const uint32_t* mask2, const float* g_agg, int64_t ld_gagg, float* dPQ, float* dW2,
                              float* db2, float* dW1c, float* g_u, int64_t g_u_stride, void* stream) {
    if (n_edges < 0 || ld_gagg < 128) return MMPDE_EINVAL;
    if (n_edges == 0) return MMPDE_OK;
    auto st = (cudaStream_t)stream;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(edge_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EDGE_BWD_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    EdgeArgs p = {};
    p.PQ = PQ; p.node4 = (const float4*)node4; p.src = edge_src; p.dst = edge_dst; p.inv_deg = inv_deg;
    p.n_edges = n_edges; p.w1c = w1c; p.w2 = w2; p.mask2 = const_cast<uint32_t*>(mask2);
    p.g_agg = g_agg; p.ld_gagg = ld_gagg; p.dPQ = dPQ; p.dW2 = dW2; p.db2 = db2; p.dW1c = dW1c; p.g_u = g_u;
    p.g_u_stride = g_u_stride;
    int64_t n_tiles = (n_edges + ET - 1) / ET;
    int grid = (int)imin64(n_tiles, sm_count());
    edge_bwd_kernel<<<grid, 256, EDGE_BWD_SMEM, st>>>(p);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
