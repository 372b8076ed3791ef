// Generic fp32 node-level contraction  C (+)= act(opA(A) * opB(B) + bias + r1_row x r1_col).
// Replaces nn.Linear (+ReLU) on node tensors and their dgrad / wgrad
// (/root/reference/gnn_2d.py:44-49,67-68,99-106 and autograd at train_helper_2d.py:126).
// 128x128x16 tiles, 256 threads, 8x8 register micro-tiles, double-buffered shared memory.
#include "common.cuh"

namespace mmpde {

constexpr int GBM = 128, GBN = 128, GBK = 16, GLD = 132;

struct GemmArgs {
    const float* A; int64_t lda;
    const float* B; int64_t ldb;
    float* C; int64_t ldc;
    int64_t M; int N; int64_t K;
    int64_t k_chunk;
    const float* bias; const float* r1_row; int64_t r1_stride; const float* r1_col;
    int relu, accumulate, atomic;
};

// 8 consecutive floats along the contiguous axis starting at (row, col); zero outside [rows, cols).
__device__ __forceinline__ void load8(const float* __restrict__ base, int64_t ld, int64_t row, int64_t col,
                                      int64_t rows, int64_t cols, bool vec_ok, float (&v)[8]) {
    if (row < rows && vec_ok && col + 8 <= cols) {
        const float* p = base + row * ld + col;
        float4 a = ldg4(p), b = ldg4(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (row < rows && col + j < cols) ? __ldg(base + row * ld + col + j) : 0.f;
    }
}

template <bool AK, bool BK>
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs p) {
    __shared__ __align__(16) float As[2][GBK][GLD];
    __shared__ __align__(16) float Bs[2][GBK][GLD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * GBM;
    const int n0 = blockIdx.x * GBN;
    const int64_t k_begin = (int64_t)blockIdx.z * p.k_chunk;
    const int64_t k_end = (k_begin + p.k_chunk < p.K) ? k_begin + p.k_chunk : p.K;
    const bool a_vec = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    const bool b_vec = (p.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    auto fetch = [&](int64_t k0) {
        if (AK) load8(p.A, p.lda, m0 + (tid & 127), k0 + (tid >> 7) * 8, p.M, k_end, a_vec, ra);
        else    load8(p.A, p.lda, k0 + (tid >> 4), m0 + (tid & 15) * 8, k_end, p.M, a_vec, ra);
        if (BK) load8(p.B, p.ldb, n0 + (tid & 127), k0 + (tid >> 7) * 8, p.N, k_end, b_vec, rb);
        else    load8(p.B, p.ldb, k0 + (tid >> 4), n0 + (tid & 15) * 8, k_end, p.N, b_vec, rb);
    };
    auto stash = [&](int buf) {
        if (AK) {
#pragma unroll
            for (int j = 0; j < 8; ++j) As[buf][(tid >> 7) * 8 + j][tid & 127] = ra[j];
        } else {
            float* d = &As[buf][tid >> 4][(tid & 15) * 8];
            *reinterpret_cast<float4*>(d) = make_float4(ra[0], ra[1], ra[2], ra[3]);
            *reinterpret_cast<float4*>(d + 4) = make_float4(ra[4], ra[5], ra[6], ra[7]);
        }
        if (BK) {
#pragma unroll
            for (int j = 0; j < 8; ++j) Bs[buf][(tid >> 7) * 8 + j][tid & 127] = rb[j];
        } else {
            float* d = &Bs[buf][tid >> 4][(tid & 15) * 8];
            *reinterpret_cast<float4*>(d) = make_float4(rb[0], rb[1], rb[2], rb[3]);
            *reinterpret_cast<float4*>(d + 4) = make_float4(rb[4], rb[5], rb[6], rb[7]);
        }
    };

    if (k_begin < k_end) {
        fetch(k_begin);
        stash(0);
        __syncthreads();
        int buf = 0;
        for (int64_t k0 = k_begin; k0 < k_end; k0 += GBK) {
            bool more = k0 + GBK < k_end;
            if (more) fetch(k0 + GBK);
#pragma unroll
            for (int kk = 0; kk < GBK; ++kk) {
                float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
                float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
                float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            if (more) {
                stash(buf ^ 1);
                __syncthreads();
                buf ^= 1;
            }
        }
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
        float r1 = p.r1_row ? __ldg(p.r1_row + m * p.r1_stride) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= p.N) continue;
            float v = acc[i][j];
            float* c = p.C + m * p.ldc + n;
            if (p.atomic) {
                if (blockIdx.z == 0) {
                    if (p.bias) v += __ldg(p.bias + n);
                    if (p.r1_row) v = fmaf(r1, __ldg(p.r1_col + n), v);
                }
                atomicAdd(c, v);
            } else {
                if (p.bias) v += __ldg(p.bias + n);
                if (p.r1_row) v = fmaf(r1, __ldg(p.r1_col + n), v);
                if (p.relu) v = fmaxf(v, 0.f);
                *c = p.accumulate ? (*c + v) : v;
            }
        }
    }
}

}  // namespace mmpde

using namespace mmpde;

extern "C" int mmpde_gemm(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor, float* C,
                          int64_t ldc, int64_t M, int N, int64_t K, const float* bias, const float* r1_row,
                          int64_t r1_stride, const float* r1_col, int relu, int accumulate, int split_k, void* stream) {
    if (M < 0 || N <= 0 || K < 0 || split_k < 1) return MMPDE_EINVAL;
    if ((r1_row == nullptr) != (r1_col == nullptr)) return MMPDE_EINVAL;
    if (split_k > 1 && relu) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    GemmArgs p;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
    int64_t chunk = (K + split_k - 1) / split_k;
    chunk = ((chunk + GBK - 1) / GBK) * GBK;
    if (chunk == 0) chunk = GBK;
    int splits = (int)((K + chunk - 1) / chunk);
    if (splits < 1) splits = 1;
    p.k_chunk = chunk;
    p.bias = bias; p.r1_row = r1_row; p.r1_stride = r1_stride; p.r1_col = r1_col;
    p.relu = relu; p.accumulate = accumulate; p.atomic = (split_k > 1) ? 1 : 0;   // split_k > 1: results are ADDED atomically onto C
    dim3 grid((unsigned)((N + GBN - 1) / GBN), (unsigned)((M + GBM - 1) / GBM), (unsigned)splits);
    auto st = (cudaStream_t)stream;
    if (a_kmajor && b_kmajor) gemm_kernel<true, true><<<grid, 256, 0, st>>>(p);
    else if (a_kmajor && !b_kmajor) gemm_kernel<true, false><<<grid, 256, 0, st>>>(p);
    else if (!a_kmajor && b_kmajor) gemm_kernel<false, true><<<grid, 256, 0, st>>>(p);
    else gemm_kernel<false, false><<<grid, 256, 0, st>>>(p);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
