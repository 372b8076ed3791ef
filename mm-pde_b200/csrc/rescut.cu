// ItpNet 'res_cut' residual network on a regular grid (/root/reference/interpolate.py:54-63,95-97):
//   Conv2d(1,4,5,p2) tanh Conv2d(4,16,5,p2) tanh Conv2d(16,4,5,p2) tanh Conv2d(4,1,5,p2) tanh   on [B,1,H,W]
// 125 MFLOP per pass at the benchmark size (16 x 48 x 48): cuDNN runs it as 4 + 8 library launches with layout
// transposes around them (~0.7 ms of GPU time per training step, 6 % of the step).  Here the whole stack is ONE launch
// per direction: a CTA owns a 16 x 16 output tile of one sample and keeps every intermediate activation of the tile (with
// the halo the following layers need: 8, 6, 4, 2 pixels) in shared memory.  fp32 throughout (cuDNN's default is TF32).
//   forward : a1..a4 of the tile centre are also written to HBM (25 floats per pixel) for the backward
//   backward: delta_l = dL/d(pre-activation l) flows back through the tile with the mirrored halos (6, 4, 2, 0); weight
//             and bias gradients are summed over the tile centre per CTA, written as one partial vector per CTA and added
//             up by a second tiny kernel (deterministic: no atomics).  The input field carries no gradient.
#include "common.cuh"

namespace mmpde {
namespace rescut {

constexpr int C0 = 1, C1 = 4, C2 = 16, C3 = 4, C4 = 1;
constexpr int T = 16;                                       // tile edge
constexpr int THREADS = 512;                              // 125 registers per thread: 512 x 128 = the whole register file
constexpr int NW1 = C1 * C0 * 25, NW2 = C2 * C1 * 25, NW3 = C3 * C2 * 25, NW4 = C4 * C3 * 25;
constexpr int NPARAM = NW1 + C1 + NW2 + C2 + NW3 + C3 + NW4 + C4;       // 3425 = MMPDE_RESCUT_NPARAM
constexpr int O_W1 = 0, O_B1 = NW1, O_W2 = O_B1 + C1, O_B2 = O_W2 + NW2, O_W3 = O_B2 + C2, O_B3 = O_W3 + NW3,
              O_W4 = O_B3 + C3, O_B4 = O_W4 + NW4;

// torch Conv2d weight [CO][CI][5][5] -> shared memory [CI][25][CO] (forward: the CO outputs of a pixel are the inner loop)
template <int CI, int CO>
__device__ __forceinline__ void load_w_fwd(const float* __restrict__ w, float* __restrict__ s) {
    for (int i = threadIdx.x; i < CO * CI * 25; i += THREADS) {
        const int co = i / (CI * 25), ci = (i / 25) % CI, k = i % 25;
        s[(ci * 25 + k) * CO + co] = __ldg(w + i);
    }
}
// ... -> [CO][25][CI] (data gradient: the CI inputs of a pixel are the inner loop)
template <int CI, int CO>
__device__ __forceinline__ void load_w_bwd(const float* __restrict__ w, float* __restrict__ s) {
    for (int i = threadIdx.x; i < CO * CI * 25; i += THREADS) {
        const int co = i / (CI * 25), ci = (i / 25) % CI, k = i % 25;
        s[(co * 25 + k) * CI + ci] = __ldg(w + i);
    }
}

// acc[0..N) += v * w[0..N): the weights of one (input channel, tap) are contiguous and 16-byte aligned in shared memory
// (every thread reads the same address: a broadcast), so N = 4, 16 go out as 128-bit loads
template <int N>
__device__ __forceinline__ void fma_row(float (&acc)[N], float v, const float* __restrict__ w) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const float4 x = reinterpret_cast<const float4*>(w)[q];
            acc[4 * q] = fmaf(v, x.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, x.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, x.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, x.w, acc[4 * q + 3]);
        }
    } else {
#pragma unroll
        for (int n = 0; n < N; ++n) acc[n] = fmaf(v, w[n], acc[n]);
    }
}

// out[co][y][x] = tanh(b[co] + sum in[ci][y+ky][x+kx] w[ci][ky,kx][co]) inside the image, 0 outside (= the zero padding
// the next layer sees).  in: [CI][RIN][RIN], out: [CO][RIN-4][RIN-4]; (oy0, ox0) = image coordinates of out(0,0).
template <int CI, int CO, int RIN>
__device__ __forceinline__ void conv5_tanh(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w,
                                           const float* __restrict__ bias, int oy0, int ox0, int Hh, int Ww) {
    constexpr int ROUT = RIN - 4;
    for (int p = threadIdx.x; p < ROUT * ROUT; p += THREADS) {
        const int py = p / ROUT, px = p % ROUT;
        float acc[CO];
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[co] = bias[co];
        for (int ci = 0; ci < CI; ++ci) {
            const float* ip = in + (ci * RIN + py) * RIN + px;
            const float* wp = w + ci * 25 * CO;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
                    const float v = ip[ky * RIN + kx];
                    fma_row<CO>(acc, v, wp + (ky * 5 + kx) * CO);
                }
        }
        const bool inside = (unsigned)(oy0 + py) < (unsigned)Hh && (unsigned)(ox0 + px) < (unsigned)Ww;
#pragma unroll
        for (int co = 0; co < CO; ++co) out[(co * ROUT + py) * ROUT + px] = inside ? tanhf(acc[co]) : 0.f;
    }
}

// centre (T x T at offset OFF) of a [C][R][R] shared-memory region -> act[b][c][h][w] in HBM
template <int C, int R, int OFF>
__device__ __forceinline__ void store_centre(const float* __restrict__ s, float* __restrict__ g, int h0, int w0, int Hh, int Ww) {
    for (int i = threadIdx.x; i < C * T * T; i += THREADS) {
        const int c = i / (T * T), y = (i / T) % T, x = i % T;
        if (h0 + y < Hh && w0 + x < Ww) g[((int64_t)c * Hh + h0 + y) * Ww + w0 + x] = s[(c * R + OFF + y) * R + OFF + x];
    }
}
// [C][R][R] region whose (0,0) sits at image (h0, w0) <- act[b][c][h][w], zero outside the image
template <int C, int R>
__device__ __forceinline__ void load_region(float* __restrict__ s, const float* __restrict__ g, int h0, int w0, int Hh, int Ww) {
    for (int i = threadIdx.x; i < C * R * R; i += THREADS) {
        const int c = i / (R * R), y = (i / R) % R, x = i % R;
        const int h = h0 + y, w = w0 + x;
        s[i] = ((unsigned)h < (unsigned)Hh && (unsigned)w < (unsigned)Ww) ? __ldg(g + ((int64_t)c * Hh + h) * Ww + w) : 0.f;
    }
}

struct FwdSmem {
    static constexpr int A0 = 0, A1 = A0 + C0 * 32 * 32, A2 = A1 + C1 * 28 * 28, A3 = A2 + C2 * 24 * 24, A4 = A3 + C3 * 20 * 20,
                         W = A4 + C4 * 16 * 16, TOTAL = W + NPARAM;
};

__global__ void __launch_bounds__(THREADS) rescut_fwd_kernel(const float* __restrict__ x, int Hh, int Ww,
                                                             const float* __restrict__ params, float* __restrict__ out,
                                                             float* __restrict__ acts) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.z, h0 = blockIdx.y * T, w0 = blockIdx.x * T;
    const int64_t plane = (int64_t)Hh * Ww;
    float* sw = sm + FwdSmem::W;
    load_w_fwd<C0, C1>(params + O_W1, sw + O_W1);
    load_w_fwd<C1, C2>(params + O_W2, sw + O_W2);
    load_w_fwd<C2, C3>(params + O_W3, sw + O_W3);
    load_w_fwd<C3, C4>(params + O_W4, sw + O_W4);
    for (int i = threadIdx.x; i < C1; i += THREADS) sw[O_B1 + i] = __ldg(params + O_B1 + i);
    for (int i = threadIdx.x; i < C2; i += THREADS) sw[O_B2 + i] = __ldg(params + O_B2 + i);
    for (int i = threadIdx.x; i < C3; i += THREADS) sw[O_B3 + i] = __ldg(params + O_B3 + i);
    for (int i = threadIdx.x; i < C4; i += THREADS) sw[O_B4 + i] = __ldg(params + O_B4 + i);
    load_region<C0, 32>(sm + FwdSmem::A0, x + (int64_t)b * C0 * plane, h0 - 8, w0 - 8, Hh, Ww);
    __syncthreads();
    conv5_tanh<C0, C1, 32>(sm + FwdSmem::A0, sm + FwdSmem::A1, sw + O_W1, sw + O_B1, h0 - 6, w0 - 6, Hh, Ww);
    __syncthreads();
    conv5_tanh<C1, C2, 28>(sm + FwdSmem::A1, sm + FwdSmem::A2, sw + O_W2, sw + O_B2, h0 - 4, w0 - 4, Hh, Ww);
    __syncthreads();
    conv5_tanh<C2, C3, 24>(sm + FwdSmem::A2, sm + FwdSmem::A3, sw + O_W3, sw + O_B3, h0 - 2, w0 - 2, Hh, Ww);
    __syncthreads();
    conv5_tanh<C3, C4, 20>(sm + FwdSmem::A3, sm + FwdSmem::A4, sw + O_W4, sw + O_B4, h0, w0, Hh, Ww);
    __syncthreads();
    store_centre<C4, 16, 0>(sm + FwdSmem::A4, out + (int64_t)b * C4 * plane, h0, w0, Hh, Ww);
    if (acts != nullptr) {                                  // [B][C1 + C2 + C3][H][W]; a4 = out
        float* ab = acts + (int64_t)b * (C1 + C2 + C3) * plane;
        store_centre<C1, 28, 6>(sm + FwdSmem::A1, ab, h0, w0, Hh, Ww);
        store_centre<C2, 24, 4>(sm + FwdSmem::A2, ab + C1 * plane, h0, w0, Hh, Ww);
        store_centre<C3, 20, 2>(sm + FwdSmem::A3, ab + (C1 + C2) * plane, h0, w0, Hh, Ww);
    }
}

// G[ci][y][x] = sum_{co,ky,kx} delta[co][y+4-ky][x+4-kx] W[co][ci][ky][kx]   (delta: [CO][RD][RD], G: [CI][RD-4][RD-4]);
// then delta_prev = G * (1 - a^2) in place, a: [CI][RA][RA] with the G region at offset OA; zero outside the image.
template <int CO, int CI, int RD, int RA, int OA>
__device__ __forceinline__ void conv5_dgrad_tanh(const float* __restrict__ delta, float* __restrict__ G, const float* __restrict__ w,
                                                 const float* __restrict__ a, int oy0, int ox0, int Hh, int Ww) {
    constexpr int RG = RD - 4;
    for (int p = threadIdx.x; p < RG * RG; p += THREADS) {
        const int py = p / RG, px = p % RG;
        float acc[CI];
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) acc[ci] = 0.f;
        for (int co = 0; co < CO; ++co) {
            const float* dp = delta + (co * RD + py + 4) * RD + px + 4;
            const float* wp = w + co * 25 * CI;
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
                    const float v = dp[-ky * RD - kx];
                    fma_row<CI>(acc, v, wp + (ky * 5 + kx) * CI);
                }
        }
        const bool inside = (unsigned)(oy0 + py) < (unsigned)Hh && (unsigned)(ox0 + px) < (unsigned)Ww;
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
            const float av = a[(ci * RA + OA + py) * RA + OA + px];
            G[(ci * RG + py) * RG + px] = inside ? acc[ci] * (1.f - av * av) : 0.f;
        }
    }
}

// dW[co][ci][k] = sum over the tile centre of delta[co][p] a[ci][p + k - 2];  db[co] = sum delta[co][p]
// delta: [CO][RD][RD] centre at OD;  a: [CI][RA][RA] centre at OA (OA >= 2).  Written (not added) to this CTA's partial.
template <int CO, int CI, int RD, int OD, int RA, int OA>
__device__ __forceinline__ void conv5_wgrad(const float* __restrict__ delta, const float* __restrict__ a, float* __restrict__ dW,
                                            float* __restrict__ db) {
    for (int o = threadIdx.x; o < CO * CI * 25 + CO; o += THREADS) {
        float s = 0.f;
        if (o < CO * CI * 25) {
            const int co = o / (CI * 25), ci = (o / 25) % CI, k = o % 25, ky = k / 5, kx = k % 5;
            const float* dp = delta + (co * RD + OD) * RD + OD;
            const float* ap = a + (ci * RA + OA + ky - 2) * RA + OA + kx - 2;
            for (int y = 0; y < T; ++y)
#pragma unroll
                for (int x = 0; x < T; ++x) s = fmaf(dp[y * RD + x], ap[y * RA + x], s);
            dW[o] = s;
        } else {
            const int co = o - CO * CI * 25;
            const float* dp = delta + (co * RD + OD) * RD + OD;
            for (int y = 0; y < T; ++y)
#pragma unroll
                for (int x = 0; x < T; ++x) s += dp[y * RD + x];
            db[co] = s;
        }
    }
}

struct BwdSmem {
    static constexpr int D4 = 0, D3 = D4 + C4 * 28 * 28, A3 = D3 + C3 * 24 * 24, D2 = A3 + C3 * 24 * 24, A2 = D2 + C2 * 20 * 20,
                         D1 = A2 + C2 * 20 * 20, A1 = D1 + C1 * 16 * 16, A0 = A1 + C1 * 20 * 20, W = A0 + C0 * 20 * 20,
                         TOTAL = W + NW2 + NW3 + NW4;
};

__global__ void __launch_bounds__(THREADS) rescut_bwd_kernel(const float* __restrict__ x, int Hh, int Ww,
                                                             const float* __restrict__ params, const float* __restrict__ out,
                                                             const float* __restrict__ acts, const float* __restrict__ g_out,
                                                             float* __restrict__ partial) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.z, h0 = blockIdx.y * T, w0 = blockIdx.x * T;
    const int64_t plane = (int64_t)Hh * Ww;
    float* sw2 = sm + BwdSmem::W; float* sw3 = sw2 + NW2; float* sw4 = sw3 + NW3;
    load_w_bwd<C1, C2>(params + O_W2, sw2);
    load_w_bwd<C2, C3>(params + O_W3, sw3);
    load_w_bwd<C3, C4>(params + O_W4, sw4);
    const float* ab = acts + (int64_t)b * (C1 + C2 + C3) * plane;
    load_region<C3, 24>(sm + BwdSmem::A3, ab + (C1 + C2) * plane, h0 - 4, w0 - 4, Hh, Ww);
    load_region<C2, 20>(sm + BwdSmem::A2, ab + C1 * plane, h0 - 2, w0 - 2, Hh, Ww);
    load_region<C1, 20>(sm + BwdSmem::A1, ab, h0 - 2, w0 - 2, Hh, Ww);
    load_region<C0, 20>(sm + BwdSmem::A0, x + (int64_t)b * C0 * plane, h0 - 2, w0 - 2, Hh, Ww);
    // delta4 = g_out * (1 - a4^2) on the 28 x 28 region (a4 = the forward's output), zero outside the image
    for (int i = threadIdx.x; i < 28 * 28; i += THREADS) {
        const int h = h0 - 6 + i / 28, w = w0 - 6 + i % 28;
        float d = 0.f;
        if ((unsigned)h < (unsigned)Hh && (unsigned)w < (unsigned)Ww) {
            const float a4 = __ldg(out + (int64_t)b * plane + (int64_t)h * Ww + w);
            d = __ldg(g_out + (int64_t)b * plane + (int64_t)h * Ww + w) * (1.f - a4 * a4);
        }
        sm[BwdSmem::D4 + i] = d;
    }
    __syncthreads();
    float* mine = partial + ((int64_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * NPARAM;
    conv5_dgrad_tanh<C4, C3, 28, 24, 0>(sm + BwdSmem::D4, sm + BwdSmem::D3, sw4, sm + BwdSmem::A3, h0 - 4, w0 - 4, Hh, Ww);
    conv5_wgrad<C4, C3, 28, 6, 24, 4>(sm + BwdSmem::D4, sm + BwdSmem::A3, mine + O_W4, mine + O_B4);
    __syncthreads();
    conv5_dgrad_tanh<C3, C2, 24, 20, 0>(sm + BwdSmem::D3, sm + BwdSmem::D2, sw3, sm + BwdSmem::A2, h0 - 2, w0 - 2, Hh, Ww);
    conv5_wgrad<C3, C2, 24, 4, 20, 2>(sm + BwdSmem::D3, sm + BwdSmem::A2, mine + O_W3, mine + O_B3);
    __syncthreads();
    conv5_dgrad_tanh<C2, C1, 20, 20, 2>(sm + BwdSmem::D2, sm + BwdSmem::D1, sw2, sm + BwdSmem::A1, h0, w0, Hh, Ww);
    conv5_wgrad<C2, C1, 20, 2, 20, 2>(sm + BwdSmem::D2, sm + BwdSmem::A1, mine + O_W2, mine + O_B2);
    __syncthreads();
    conv5_wgrad<C1, C0, 16, 0, 20, 2>(sm + BwdSmem::D1, sm + BwdSmem::A0, mine + O_W1, mine + O_B1);
}

__global__ void __launch_bounds__(256) rescut_reduce_kernel(const float* __restrict__ partial, int n_partial, float* __restrict__ g_params) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= NPARAM) return;
    float s = 0.f;
    for (int c = 0; c < n_partial; ++c) s += __ldg(partial + (int64_t)c * NPARAM + j);
    g_params[j] = s;
}

}  // namespace rescut
}  // namespace mmpde

using namespace mmpde;
using namespace mmpde::rescut;

static_assert(NPARAM == MMPDE_RESCUT_NPARAM, "packed parameter count");
static_assert(C1 + C2 + C3 == MMPDE_RESCUT_ACT_CHANNELS, "saved activation channels");

extern "C" int mmpde_rescut_fwd(const float* x, int64_t batch, int height, int width, const float* params, float* out,
                                float* acts, void* stream) {
    if (batch < 0 || height <= 0 || width <= 0 || params == nullptr) return MMPDE_EINVAL;
    if (batch == 0) return MMPDE_OK;
    if (batch > 65535) return MMPDE_EINVAL;
    constexpr size_t smem = FwdSmem::TOTAL * sizeof(float);
    MMPDE_ENSURE_SMEM(rescut_fwd_kernel, smem);
    const dim3 grid((width + T - 1) / T, (height + T - 1) / T, (unsigned)batch);
    rescut_fwd_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(x, height, width, params, out, acts);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int64_t mmpde_rescut_bwd_workspace_floats(int64_t batch, int height, int width) {
    return batch * ((width + T - 1) / T) * ((height + T - 1) / T) * (int64_t)NPARAM;
}

extern "C" int mmpde_rescut_bwd(const float* x, int64_t batch, int height, int width, const float* params, const float* out,
                                const float* acts, const float* g_out, float* workspace, float* g_params, void* stream) {
    if (batch < 0 || height <= 0 || width <= 0 || params == nullptr || g_params == nullptr) return MMPDE_EINVAL;
    if (batch > 65535) return MMPDE_EINVAL;
    constexpr size_t smem = BwdSmem::TOTAL * sizeof(float);
    MMPDE_ENSURE_SMEM(rescut_bwd_kernel, smem);
    const dim3 grid((width + T - 1) / T, (height + T - 1) / T, (unsigned)batch);
    if (batch > 0) {
        if (workspace == nullptr) return MMPDE_EINVAL;
        rescut_bwd_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(x, height, width, params, out, acts, g_out, workspace);
        MMPDE_CHECK_LAUNCH();
    }
    rescut_reduce_kernel<<<(NPARAM + 255) / 256, 256, 0, (cudaStream_t)stream>>>(workspace, (int)(batch * grid.x * grid.y), g_params);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
