// Host side of the TMA-fed operand paths: tensor maps (cuTensorMapEncodeTiled, reached through the runtime's driver
// entry point so that nothing links against libcuda) and the device wrappers of the bulk-tensor copies.
//
// What the kernels use (semantics measured on a B200 with profiles/experiments/tma_gather4_probe.cu, see
// profiles/r02_tma_gather4_probe.txt -- no public documentation is reachable from the build box):
//   cp.async.bulk.tensor.2d ... tile::gather4  with a 2-D map whose box is {cols, 1}: the FOUR rows named by the
//   instruction land back to back in shared memory ([4][cols], no swizzle), the mbarrier sees 4*cols*4 bytes, and a
//   row index outside [0, rows) is filled with zeros instead of faulting.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmpde {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Map over `cols` consecutive fp32 columns (starting at `base`) of a row-major matrix with `rows` rows and a leading
// dimension of `ld` floats, for row gathers: box = one row of `cols` values.  base 16-byte aligned, ld multiple of 4,
// cols <= 256.  Returns 0 or a negative status.
inline int make_row_gather_map(CUtensorMap* map, const float* base, int64_t rows, int64_t ld, int cols) {
    EncodeTiledFn enc = encode_tiled();
    if (enc == nullptr) return -2;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)(rows > 0 ? rows : 1)};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)cols, 1}, estr[2] = {1, 1};
    CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS ? 0 : -3;
}

// rows r0..r3 (columns col .. col+box-1) of the mapped matrix -> shared memory at `dst` ([4][box] fp32), completion
// (4 * box * 4 bytes) on the mbarrier `bar`.  One thread issues.
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
        "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace tma
}  // namespace mmpde
