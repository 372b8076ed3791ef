// Conv1d decoder over the 128-wide feature axis, one warp per node.
// Replaces nn.Conv1d(1,4,16,s3) -> ReLU -> Conv1d(4,8,12,s3) -> ReLU -> Conv1d(8,1,8,s2) and the
// 0.1*dt scaling (/root/reference/gnn_2d.py:108-114,136-139) plus their autograd.
// Lengths: 128 -> [4][38] -> [8][9] -> [1][1]; the last conv reads positions 0..7 of the 9.
#include "common.cuh"

namespace mmpde {

constexpr int DEC_NP = 525;
constexpr int OFF_W1 = 0, OFF_B1 = 64, OFF_W2 = 68, OFF_B2 = 452, OFF_W3 = 460, OFF_B3 = 524;
constexpr int WARPS = 8;

struct __align__(16) DecScratch {
    float h[128];
    float a1[4 * 38];   // post-ReLU
    float a2[8 * 9];    // post-ReLU
    float g1[4 * 38];
    float g2[8 * 9];
};

__device__ __forceinline__ float dec_forward(const float* sp, DecScratch& s, int lane) {
    for (int idx = lane; idx < 152; idx += 32) {
        int c = idx / 38, p = idx - c * 38;
        float acc = sp[OFF_B1 + c];
#pragma unroll
        for (int t = 0; t < 16; ++t) acc = fmaf(sp[OFF_W1 + c * 16 + t], s.h[3 * p + t], acc);
        s.a1[idx] = fmaxf(acc, 0.f);
    }
    __syncwarp();
    for (int idx = lane; idx < 72; idx += 32) {
        int o = idx / 9, p = idx - o * 9;
        float acc = sp[OFF_B2 + o];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int t = 0; t < 12; ++t) acc = fmaf(sp[OFF_W2 + (o * 4 + c) * 12 + t], s.a1[c * 38 + 3 * p + t], acc);
        s.a2[idx] = fmaxf(acc, 0.f);
    }
    __syncwarp();
    float part = 0.f;
    for (int idx = lane; idx < 64; idx += 32) {
        int c = idx >> 3, t = idx & 7;
        part = fmaf(sp[OFF_W3 + idx], s.a2[c * 9 + t], part);
    }
    return warp_sum(part) + sp[OFF_B3];
}

// The same forward with the convolution weights of each lane in REGISTERS.  dec_forward reads a weight AND an activation
// from shared memory for every multiply-add (the kernel sat at 2-3 % of its HBM roofline on the shared-memory pipe:
// mio_throttle, profiles/r02_ncu_full_summary.txt).  Here a lane keeps ONE output channel per layer -- layer 1: channel
// lane / 8, positions (lane % 8) + 8 j; layer 2: channel lane / 4, positions (lane % 4) + 4 j -- so its 16 + 48 weights are
// loaded once per kernel and a multiply-add costs one shared-memory load.  Same accumulation order per output: same bits.
struct DecRegs { float w1[16], b1, w2[48], b2, w3[2]; };
__device__ __forceinline__ void dec_load_regs(const float* sp, int lane, DecRegs& r) {
    const int c = lane >> 3, o = lane >> 2;
#pragma unroll
    for (int t = 0; t < 16; ++t) r.w1[t] = sp[OFF_W1 + c * 16 + t];
    r.b1 = sp[OFF_B1 + c];
#pragma unroll
    for (int t = 0; t < 48; ++t) r.w2[t] = sp[OFF_W2 + o * 48 + t];
    r.b2 = sp[OFF_B2 + o];
    r.w3[0] = sp[OFF_W3 + lane]; r.w3[1] = sp[OFF_W3 + 32 + lane];
}
__device__ __forceinline__ float dec_forward_regs(const float* sp, const DecRegs& r, DecScratch& s, int lane) {
    const int c = lane >> 3, o = lane >> 2;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int p = (lane & 7) + 8 * j;
        if (p < 38) {
            float acc = r.b1;
#pragma unroll
            for (int t = 0; t < 16; ++t) acc = fmaf(r.w1[t], s.h[3 * p + t], acc);
            s.a1[c * 38 + p] = fmaxf(acc, 0.f);
        }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int p = (lane & 3) + 4 * j;
        if (p < 9) {
            float acc = r.b2;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                for (int t = 0; t < 12; ++t) acc = fmaf(r.w2[cc * 12 + t], s.a1[cc * 38 + 3 * p + t], acc);
            s.a2[o * 9 + p] = fmaxf(acc, 0.f);
        }
    }
    __syncwarp();
    float part = 0.f;
    part = fmaf(r.w3[0], s.a2[(lane >> 3) * 9 + (lane & 7)], part);
    part = fmaf(r.w3[1], s.a2[((lane + 32) >> 3) * 9 + (lane & 7)], part);
    return warp_sum(part) + sp[OFF_B3];
}

// act1 [M,256] / act2 [M,128] (optional): the post-ReLU activations of the two hidden blocks in the layout of the
// Toeplitz form (column c*38+i / o*9+j, zero padding behind) -- saved for a backward that runs as dense contractions
// on the tensor cores while every ReLU mask still comes from THIS fp32 evaluation.
__global__ void __launch_bounds__(WARPS * 32) decoder_fwd_kernel(const float* __restrict__ h, int64_t ldh, int64_t M,
                                                                 const float* __restrict__ params, float scale,
                                                                 float* __restrict__ out, float* __restrict__ act1,
                                                                 float* __restrict__ act2) {
    __shared__ float sp[DEC_NP];
    __shared__ __align__(16) DecScratch scr[WARPS];
    for (int i = threadIdx.x; i < DEC_NP; i += blockDim.x) sp[i] = __ldg(params + i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DecScratch& s = scr[warp];
    DecRegs regs;
    dec_load_regs(sp, lane, regs);
    for (int64_t n = (int64_t)blockIdx.x * WARPS + warp; n < M; n += (int64_t)gridDim.x * WARPS) {
        __syncwarp();
        *reinterpret_cast<float4*>(&s.h[lane * 4]) = ldg4(h + n * ldh + lane * 4);
        __syncwarp();
        float v = dec_forward_regs(sp, regs, s, lane);
        if (lane == 0) out[n] = scale * v;
        if (act1 != nullptr) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c0 = (j * 32 + lane) * 4;                   // 152 = 38 float4
                float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 < 152) v4 = *reinterpret_cast<const float4*>(&s.a1[c0]);
                *reinterpret_cast<float4*>(act1 + n * 256 + c0) = v4;
            }
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < 18) v4 = *reinterpret_cast<const float4*>(&s.a2[lane * 4]);     // 72 = 18 float4
            *reinterpret_cast<float4*>(act2 + n * 128 + lane * 4) = v4;
        }
    }
}

__global__ void __launch_bounds__(WARPS * 32) decoder_bwd_kernel(const float* __restrict__ h, int64_t ldh, int64_t M,
                                                                 const float* __restrict__ params, float scale,
                                                                 const float* __restrict__ g_out, float* __restrict__ g_h,
                                                                 int64_t ldg, float* __restrict__ g_params) {
    __shared__ float sp[DEC_NP];
    __shared__ __align__(16) DecScratch scr[WARPS];
    for (int i = threadIdx.x; i < DEC_NP; i += blockDim.x) sp[i] = __ldg(params + i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DecScratch& s = scr[warp];
    // lane-private accumulators for parameter index lane + 32*i (17 slots cover 525 parameters)
    float gp[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) gp[i] = 0.f;
    // Everything that does not depend on the node is decoded ONCE: the scratch offsets of each parameter slot ...
    //   w1[c][t] (pi <  64): sum_p g1[c*38+p] * h[3p+t]         -> offA = c*38, offB = t,        n = 38, stride 3
    //   w2[o][c][t]        : sum_p g2[o*9+p]  * a1[c*38+3p+t]   -> offA = o*9,  offB = c*38 + t, n = 9,  stride 3
    int offA[17], offB[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) {
        const int pi = lane + 32 * i;
        offA[i] = offB[i] = 0;
        if (pi < OFF_B1) { offA[i] = (pi >> 4) * 38; offB[i] = pi & 15; }
        else if (pi >= OFF_W2 && pi < OFF_B2) { const int r = pi - OFF_W2; offA[i] = (r / 48) * 9; offB[i] = ((r % 48) / 12) * 38 + r % 12; }
    }
    // ... and, for the two transposed convolutions, quotient / remainder by the stride 3 of each output this lane owns
    int q1d[5], q1m[5];                                  // layer-1 outputs idx = lane + 32*j < 152: position q = idx % 38
#pragma unroll
    for (int j = 0; j < 5; ++j) { const int idx = lane + 32 * j, q = idx % 38; q1d[j] = q / 3; q1m[j] = q - 3 * (q / 3); }
    int q0d[4], q0m[4];                                  // input positions q = 4*lane + i
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int q = lane * 4 + i; q0d[i] = q / 3; q0m[i] = q - 3 * (q / 3); }

    for (int64_t n = (int64_t)blockIdx.x * WARPS + warp; n < M; n += (int64_t)gridDim.x * WARPS) {
        __syncwarp();
        *reinterpret_cast<float4*>(&s.h[lane * 4]) = ldg4(h + n * ldh + lane * 4);
        __syncwarp();
        (void)dec_forward(sp, s, lane);
        const float g3 = scale * __ldg(g_out + n);
        // layer 3: g wrt pre-activation of layer 2
        for (int idx = lane; idx < 72; idx += 32) {
            int c = idx / 9, t = idx - c * 9;
            float g = (t < 8) ? g3 * sp[OFF_W3 + c * 8 + t] : 0.f;
            s.g2[idx] = s.a2[idx] > 0.f ? g : 0.f;
        }
        __syncwarp();
        // layer 2 -> g wrt pre-activation of layer 1: g1[c][q] = sum over (p, t) with 3p + t = q, t < 12, p < 9
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int idx = lane + 32 * j;
            if (idx < 152) {
                const int c = idx / 38;
                float accm[4] = {0.f, 0.f, 0.f, 0.f};      // one chain per m: 4 independent 8-term chains
#pragma unroll
                for (int m = 0; m < 4; ++m) {              // t = q%3 + 3m, p = q/3 - m
                    const int t = q1m[j] + 3 * m, p = q1d[j] - m;
                    if (p >= 0 && p < 9) {
#pragma unroll
                        for (int o = 0; o < 8; ++o) accm[m] = fmaf(s.g2[o * 9 + p], sp[OFF_W2 + (o * 4 + c) * 12 + t], accm[m]);
                    }
                }
                const float acc = (accm[0] + accm[1]) + (accm[2] + accm[3]);
                s.g1[idx] = s.a1[idx] > 0.f ? acc : 0.f;
            }
        }
        __syncwarp();
        // layer 1 -> g_h[q] = sum over (p, t) with 3p + t = q, t < 16, p < 38
        float gh[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float accm[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const int t = q0m[i] + 3 * m, p = q0d[i] - m;
                if (t < 16 && p >= 0 && p < 38) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) accm[m] = fmaf(s.g1[c * 38 + p], sp[OFF_W1 + c * 16 + t], accm[m]);
                }
            }
            gh[i] = ((accm[0] + accm[1]) + (accm[2] + accm[3])) + (accm[4] + accm[5]);
        }
        *reinterpret_cast<float4*>(g_h + n * ldg + lane * 4) = make_float4(gh[0], gh[1], gh[2], gh[3]);
        // parameter gradients
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const int pi = lane + 32 * i;
            if (pi >= DEC_NP) break;
            float acc = 0.f;
            // (several partial sums per chain: a single 38-term FMA chain is 38 x the FMA latency)
            if (pi < OFF_B1) {                       // w1[c][t]
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int p = 0; p < 36; p += 4) {
                    a0 = fmaf(s.g1[offA[i] + p], s.h[3 * p + offB[i]], a0);
                    a1 = fmaf(s.g1[offA[i] + p + 1], s.h[3 * p + 3 + offB[i]], a1);
                    a2 = fmaf(s.g1[offA[i] + p + 2], s.h[3 * p + 6 + offB[i]], a2);
                    a3 = fmaf(s.g1[offA[i] + p + 3], s.h[3 * p + 9 + offB[i]], a3);
                }
                a0 = fmaf(s.g1[offA[i] + 36], s.h[108 + offB[i]], a0);
                a1 = fmaf(s.g1[offA[i] + 37], s.h[111 + offB[i]], a1);
                acc = (a0 + a1) + (a2 + a3);
            } else if (pi < OFF_W2) {                // b1[c]
                int c = pi - OFF_B1;
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int p = 0; p < 38; p += 2) { a0 += s.g1[c * 38 + p]; a1 += s.g1[c * 38 + p + 1]; }
                acc = a0 + a1;
            } else if (pi < OFF_B2) {                // w2[o][c][t]
                float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int p = 0; p < 9; p += 3) {
                    a0 = fmaf(s.g2[offA[i] + p], s.a1[offB[i] + 3 * p], a0);
                    a1 = fmaf(s.g2[offA[i] + p + 1], s.a1[offB[i] + 3 * p + 3], a1);
                    a2 = fmaf(s.g2[offA[i] + p + 2], s.a1[offB[i] + 3 * p + 6], a2);
                }
                acc = (a0 + a1) + a2;
            } else if (pi < OFF_W3) {                // b2[o]
                int o = pi - OFF_B2;
                for (int p = 0; p < 9; ++p) acc += s.g2[o * 9 + p];
            } else if (pi < OFF_B3) {                // w3[c][t]
                int r = pi - OFF_W3;
                acc = g3 * s.a2[(r >> 3) * 9 + (r & 7)];
            } else {
                acc = g3;
            }
            gp[i] += acc;
        }
    }
#pragma unroll
    for (int i = 0; i < 17; ++i) {
        int pi = lane + 32 * i;
        if (pi < DEC_NP) atomicAdd(g_params + pi, gp[i]);
    }
}

}  // namespace mmpde

using namespace mmpde;

// out[m][c] = g[m] * w[c] * (act[m][c] > 0): dL/d(pre-activation) of the decoder's second block from dL/d(out) (the last
// convolution is a dot product with w) -- the first step of the Toeplitz-GEMM decoder backward (ops._decoder_backward)
namespace mmpde {
__global__ void outer_gate_kernel(const float* __restrict__ g, const float* __restrict__ w, const float* __restrict__ act, int64_t lda,
                                  float* __restrict__ out, int64_t ldo, int64_t M) {
    const int lane = threadIdx.x & 31;
    const float4 wv = ldg4(w + lane * 4);
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t m = warp; m < M; m += n_warps) {
        const float gm = __ldg(g + m);
        const float4 a = ldg4(act + m * lda + lane * 4);
        float4 o;
        o.x = a.x > 0.f ? gm * wv.x : 0.f; o.y = a.y > 0.f ? gm * wv.y : 0.f;
        o.z = a.z > 0.f ? gm * wv.z : 0.f; o.w = a.w > 0.f ? gm * wv.w : 0.f;
        *reinterpret_cast<float4*>(out + m * ldo + lane * 4) = o;
    }
}
}  // namespace mmpde

extern "C" int mmpde_outer_gate(const float* g, const float* w, const float* act, int64_t lda, float* out, int64_t ldo, int64_t M,
                                void* stream) {
    if (M < 0 || (lda & 3) || (ldo & 3)) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(out)) & 15) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    const int grid = (int)imin64((M + 7) / 8, (int64_t)sm_count() * 8);
    outer_gate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, w, act, lda, out, ldo, M);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_decoder_fwd(const float* h, int64_t ldh, int64_t M, const float* params, float scale, float* out,
                                 void* stream) {
    if (M < 0 || ldh % 4) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    int grid = (int)imin64((M + WARPS - 1) / WARPS, (int64_t)sm_count() * 2);      // 108 registers: two CTAs per SM, weights loaded once
    decoder_fwd_kernel<<<grid, WARPS * 32, 0, (cudaStream_t)stream>>>(h, ldh, M, params, scale, out, nullptr, nullptr);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_decoder_fwd_acts(const float* h, int64_t ldh, int64_t M, const float* params, float scale, float* out,
                                      float* act1, float* act2, void* stream) {
    if (M < 0 || ldh % 4 || act1 == nullptr || act2 == nullptr) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(act1) | reinterpret_cast<uintptr_t>(act2)) & 15) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    int grid = (int)imin64((M + WARPS - 1) / WARPS, (int64_t)sm_count() * 2);      // 108 registers: two CTAs per SM, weights loaded once
    decoder_fwd_kernel<<<grid, WARPS * 32, 0, (cudaStream_t)stream>>>(h, ldh, M, params, scale, out, act1, act2);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_decoder_bwd(const float* h, int64_t ldh, int64_t M, const float* params, float scale,
                                 const float* g_out, float* g_h, int64_t ldg, float* g_params, void* stream) {
    if (M < 0 || ldh % 4 || ldg % 4) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    int grid = (int)imin64((M + WARPS - 1) / WARPS, (int64_t)sm_count() * 4);
    decoder_bwd_kernel<<<grid, WARPS * 32, 0, (cudaStream_t)stream>>>(h, ldh, M, params, scale, g_out, g_h, ldg, g_params);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
