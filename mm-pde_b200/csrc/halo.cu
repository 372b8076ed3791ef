// Halo exchange helpers for the graph-partitioned processor (mm-pde_b200/partition.py): rows of a strided fp32
// matrix <-> one contiguous send / receive buffer.  Both are pure HBM streaming (one warp per 128-column row,
// 128-bit accesses); the reference has no counterpart (single device).
#include "common.cuh"

namespace mmpde {

__global__ void rows_gather_kernel(const float* __restrict__ src, int64_t ld_src, const int* __restrict__ idx, int64_t n_rows,
                                   int ncols, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const float* s = src + (int64_t)__ldg(idx + r) * ld_src;
        float* o = out + r * ncols;
        for (int c = lane * 4; c < ncols; c += 128) *reinterpret_cast<float4*>(o + c) = ldg4(s + c);
    }
}

__global__ void rows_scatter_add_kernel(const float* __restrict__ in, const int* __restrict__ idx, int64_t n_rows, int ncols,
                                        float* __restrict__ dst, int64_t ld_dst) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        float* d = dst + (int64_t)__ldg(idx + r) * ld_dst;
        const float* s = in + r * ncols;
        for (int c = lane * 4; c < ncols; c += 128) red_add_v4(d + c, ldg4(s + c));   // a row may be needed by several peers
    }
}

}  // namespace mmpde

using namespace mmpde;

static int rows_grid(int64_t n_rows) { return (int)imin64((n_rows + 7) / 8, (int64_t)sm_count() * 8); }

extern "C" int mmpde_rows_gather(const float* src, int64_t ld_src, const int32_t* idx, int64_t n_rows, int ncols, float* out,
                                 void* stream) {
    if (n_rows < 0 || ncols <= 0 || (ncols & 3) || (ld_src & 3)) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) & 15) return MMPDE_EINVAL;
    if (n_rows == 0) return MMPDE_OK;
    rows_gather_kernel<<<rows_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(src, ld_src, idx, n_rows, ncols, out);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_rows_scatter_add(const float* in, const int32_t* idx, int64_t n_rows, int ncols, float* dst, int64_t ld_dst,
                                      void* stream) {
    if (n_rows < 0 || ncols <= 0 || (ncols & 3) || (ld_dst & 3)) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(dst)) & 15) return MMPDE_EINVAL;
    if (n_rows == 0) return MMPDE_OK;
    rows_scatter_add_kernel<<<rows_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(in, idx, n_rows, ncols, dst, ld_dst);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

// ---- out[m*stride] (+)= dot(A[m, 0:ncols], w[0:ncols])  -- the N = 1 contractions of the backward (dL/du from dP', dQ')
namespace mmpde {
__global__ void __launch_bounds__(256) rows_dot_kernel(const float* __restrict__ A, int64_t lda, int ncols, const float* __restrict__ w,
                                                       float* __restrict__ out, int64_t out_stride, int64_t n_rows, int accumulate) {
    // four rows per warp pass: their loads are issued together (one row at a time left a single 1 KB request in flight
    // per warp: 2 TB/s on a kernel that only streams A)
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp * 4; r < n_rows; r += n_warps * 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = lane * 4; c < ncols; c += 128) {
            const float4 y = ldg4(w + c);
            float4 x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = (r + j < n_rows) ? ldg4(A + (r + j) * lda + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fmaf(x[j].x, y.x, fmaf(x[j].y, y.y, fmaf(x[j].z, y.z, fmaf(x[j].w, y.w, acc[j]))));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
        const float mine = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
        if (lane < 4 && r + lane < n_rows) {
            float* o = out + (r + lane) * out_stride;
            *o = accumulate ? *o + mine : mine;
        }
    }
}

// out[m][n] = bias[n] + sum_f node4[m][f] W[n][f]   (n < 128, f < 4): the K = 4 input layer of the encoder
// (gnn_2d.py:100) as an elementwise pass -- one float4 of outputs per thread, a row per warp
__global__ void __launch_bounds__(256) node4_linear_kernel(const float4* __restrict__ node4, const float4* __restrict__ W,
                                                           const float* __restrict__ bias, float* __restrict__ out, int64_t ldo,
                                                           int64_t n_rows) {
    const int lane = threadIdx.x & 31;
    float4 w[4], b = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = __ldg(W + lane * 4 + j);
    if (bias) b = ldg4(bias + lane * 4);
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const float4 x = __ldg(node4 + r);
        float4 o;
        o.x = fmaf(x.x, w[0].x, fmaf(x.y, w[0].y, fmaf(x.z, w[0].z, fmaf(x.w, w[0].w, b.x))));
        o.y = fmaf(x.x, w[1].x, fmaf(x.y, w[1].y, fmaf(x.z, w[1].z, fmaf(x.w, w[1].w, b.y))));
        o.z = fmaf(x.x, w[2].x, fmaf(x.y, w[2].y, fmaf(x.z, w[2].z, fmaf(x.w, w[2].w, b.z))));
        o.w = fmaf(x.x, w[3].x, fmaf(x.y, w[3].y, fmaf(x.z, w[3].z, fmaf(x.w, w[3].w, b.w))));
        *reinterpret_cast<float4*>(out + r * ldo + lane * 4) = o;
    }
}
}  // namespace mmpde

extern "C" int mmpde_node4_linear(const float* node4, const float* W, const float* bias, float* out, int64_t ldo, int64_t n_rows,
                                  void* stream) {
    if (n_rows < 0 || W == nullptr || (ldo & 3)) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(node4) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(bias) |
         reinterpret_cast<uintptr_t>(out)) & 15) return MMPDE_EINVAL;
    if (n_rows == 0) return MMPDE_OK;
    mmpde::node4_linear_kernel<<<rows_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(node4), reinterpret_cast<const float4*>(W), bias, out, ldo, n_rows);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_rows_dot(const float* A, int64_t lda, int ncols, const float* w, float* out, int64_t out_stride, int64_t n_rows,
                              int accumulate, void* stream) {
    if (n_rows < 0 || ncols <= 0 || (ncols & 3) || (lda & 3)) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(w)) & 15) return MMPDE_EINVAL;
    if (n_rows == 0) return MMPDE_OK;
    mmpde::rows_dot_kernel<<<rows_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(A, lda, ncols, w, out, out_stride, n_rows, accumulate);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
