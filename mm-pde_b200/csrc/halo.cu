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
__global__ void rows_dot_kernel(const float* __restrict__ A, int64_t lda, int ncols, const float* __restrict__ w, float* __restrict__ out,
                                int64_t out_stride, int64_t n_rows, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const float* a = A + r * lda;
        float acc = 0.f;
        for (int c = lane * 4; c < ncols; c += 128) {
            const float4 x = ldg4(a + c), y = ldg4(w + c);
            acc = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, acc))));
        }
        acc = warp_sum(acc);
        if (lane == 0) out[r * out_stride] = accumulate ? out[r * out_stride] + acc : acc;
    }
}
}  // namespace mmpde

extern "C" int mmpde_rows_dot(const float* A, int64_t lda, int ncols, const float* w, float* out, int64_t out_stride, int64_t n_rows,
                              int accumulate, void* stream) {
    if (n_rows < 0 || ncols <= 0 || (ncols & 3) || (lda & 3)) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(w)) & 15) return MMPDE_EINVAL;
    if (n_rows == 0) return MMPDE_OK;
    mmpde::rows_dot_kernel<<<rows_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(A, lda, ncols, w, out, out_stride, n_rows, accumulate);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
