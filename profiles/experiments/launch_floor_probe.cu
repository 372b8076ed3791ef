// Probe (run on a B200): what does a back-to-back launch of a persistent one-CTA-per-SM kernel cost before it does any
// work?  Variants add the fixed parts of the tensor-core kernels one by one: big dynamic shared memory, TMEM allocation,
// setmaxnreg, mbarrier initialisation.  Timed with CUDA events over 400 launches in one stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mm-pde_b200/csrc -o launch_floor_probe profiles/experiments/launch_floor_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace mmpde::tc;

template <int MODE>
__global__ void __launch_bounds__(640, 1) k(float* out) {
    extern __shared__ unsigned char sm[];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (MODE >= 2) {
        if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (MODE >= 3) {
        if (warp < 8) reg_inc<104>();
        else if (warp < 16) reg_inc<112>();
        else reg_dec<40>();
    }
    if (threadIdx.x == 0 && out != nullptr) out[blockIdx.x] = (float)sm[threadIdx.x];
    if (MODE >= 2) {
        tc_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc(slot, 512);
    }
}

template <int MODE>
float run(int grid, int threads, size_t smem, int n = 400) {
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 20; ++i) k<MODE><<<grid, threads, smem>>>(nullptr);
    cudaEventRecord(e0);
    for (int i = 0; i < n; ++i) k<MODE><<<grid, threads, smem>>>(nullptr);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / n;
}

int main() {
    printf("back-to-back launches of an empty kernel, us per launch (400 launches, one stream)\n");
    printf("  148 CTAs x 640 threads, 1 KB smem                         : %.2f\n", run<0>(148, 640, 1024));
    printf("  148 CTAs x 640 threads, 194 KB smem                       : %.2f\n", run<1>(148, 640, 194 * 1024));
    printf("  ... + TMEM alloc 512 / dealloc                            : %.2f\n", run<2>(148, 640, 194 * 1024));
    printf("  ... + setmaxnreg inc/dec                                  : %.2f\n", run<3>(148, 640, 194 * 1024));
    printf("  36 CTAs x 640 threads, 194 KB smem, TMEM, setmaxnreg      : %.2f\n", run<3>(36, 640, 194 * 1024));
    printf("  148 CTAs x 256 threads, 1 KB smem                         : %.2f\n", run<0>(148, 256, 1024));
    printf("  1184 CTAs x 256 threads, 1 KB smem                        : %.2f\n", run<0>(1184, 256, 1024));
    // alternating shared-memory configurations (what a step does: big-smem tensor-core kernels between small elementwise ones)
    {
        cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 194 * 1024);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < 200; ++i) { k<1><<<148, 640, 194 * 1024>>>(nullptr); k<0><<<1184, 256, 1024>>>(nullptr); }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("  alternating 194 KB / 1 KB kernels, per PAIR               : %.2f\n", ms * 1e3f / 200);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
