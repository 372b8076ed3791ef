// Shared device helpers for libmmpde_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
#include "../../include/mmpde_b200.h"

#define MMPDE_CHECK_LAUNCH()                                  \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return (int)e__;              \
    } while (0)

namespace mmpde {

constexpr int H = MMPDE_H;

// Host-side caches are keyed by the CURRENT device: one process may drive several GPUs (mmpde.py --device cuda:1).
constexpr int MAX_DEVICES = 64;
inline int cur_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0;
}
inline int sm_count() {
    static int n[MAX_DEVICES] = {0};
    const int dev = cur_device();
    if (n[dev] == 0) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}
// Cap on the CTAs of the persistent one-CTA-per-SM kernels (edge, node GEMM, weight gradient), per device; 0 = one per SM.
// A training step runs its two solvers as parallel branches: at full width their persistent kernels take turns on the whole
// chip, at half width they run side by side (mmpde_set_persistent_ctas; train_helper_2d sets it for the overlapped step).
inline int& persistent_cap() {
    static int cap[MAX_DEVICES] = {0};
    return cap[cur_device()];
}
inline int persistent_ctas() {
    const int c = persistent_cap(), n = sm_count();
    return (c > 0 && c < n) ? c : n;
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel AND device
#define MMPDE_ENSURE_SMEM(kernel, bytes)                                                                              \
    do {                                                                                                              \
        static bool done__[mmpde::MAX_DEVICES] = {false};                                                             \
        const int dev__ = mmpde::cur_device();                                                                        \
        if (!done__[dev__]) {                                                                                         \
            cudaError_t e__ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
            if (e__ != cudaSuccess) return (int)e__;                                                                  \
            done__[dev__] = true;                                                                                     \
        }                                                                                                             \
    } while (0)

inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 128-bit vector reduction to global memory (sm_90+): one L2 atomic transaction for 4 floats.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace mmpde
