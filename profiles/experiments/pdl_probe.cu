// Probe (run on a B200): does programmatic dependent launch (PDL) shorten the gap between two dependent persistent kernels,
// in a stream and -- what the product needs -- inside a captured and replayed CUDA graph?
// Kernel = the fixed parts of the tensor-core kernels (148 CTAs x 640 threads, 194 KB dynamic shared memory, TMEM allocation,
// setmaxnreg) + a dependent read-modify-write of one word per CTA (out[cta] += 1 by one thread: the value after n launches
// proves that every launch saw its predecessor's store) + an optional busy loop standing in for the work.
//   PDL form: griddepcontrol.launch_dependents at the top (the next grid may be scheduled as soon as every CTA of this one
//   has started), the on-chip prologue, then griddepcontrol.wait before the first global access.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mm-pde_b200/csrc -o pdl_probe profiles/experiments/pdl_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace mmpde::tc;

template <bool PDL>
__global__ void __launch_bounds__(640, 1) k(float* out, int spin) {
    extern __shared__ unsigned char sm[];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp < 8) reg_inc<104>();
    else if (warp < 16) reg_inc<112>();
    else reg_dec<40>();
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) {
        float v = out[blockIdx.x];
        const long long t0 = clock64();
        while (clock64() - t0 < spin) { }
        out[blockIdx.x] = v + 1.0f + (float)sm[0] * 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

template <bool PDL>
static void launch(cudaStream_t st, float* out, int spin) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = 194 * 1024; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k<PDL>, out, spin);
}

template <bool PDL>
static void run(const char* name, int spin, bool graph) {
    const int n = 400;
    cudaFuncSetAttribute(k<PDL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 194 * 1024);
    float* out; cudaMalloc(&out, 148 * 4); cudaMemset(out, 0, 148 * 4);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0.f; int launches = 0;
    if (!graph) {
        for (int i = 0; i < 20; ++i) launch<PDL>(st, out, spin);
        cudaEventRecord(e0, st);
        for (int i = 0; i < n; ++i) launch<PDL>(st, out, spin);
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        launches = n + 20;
    } else {
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        for (int i = 0; i < n; ++i) launch<PDL>(st, out, spin);
        cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        cudaGraphLaunch(ge, st);
        cudaEventRecord(e0, st);
        cudaGraphLaunch(ge, st);
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        launches = 2 * n;
    }
    cudaEventElapsedTime(&ms, e0, e1);
    float h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    bool ok = true;
    for (int i = 0; i < 148; ++i) ok = ok && h[i] == (float)launches;
    printf("  %-44s spin %6d clk : %6.2f us per launch   chain %s   %s\n", name, spin, ms * 1e3f / n, ok ? "intact" : "BROKEN",
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaStreamDestroy(st);
}

int main() {
    printf("dependent launches of a 148 x 640-thread kernel (194 KB smem, TMEM alloc, setmaxnreg), 400 launches\n");
    for (int spin : {0, 4000, 20000}) {
        run<false>("stream, ordinary launches", spin, false);
        run<true>("stream, programmatic dependent launch", spin, false);
        run<false>("graph replay, ordinary edges", spin, true);
        run<true>("graph replay, programmatic edges", spin, true);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
