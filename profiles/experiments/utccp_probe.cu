// Probe (run on a B200): does tcgen05.cp.128x256b read a SWIZZLE_128B K-major operand image the way the MMA does?
// smem image: [128 rows][128 B], 16-byte chunks XOR-swizzled with (row & 7) (tc::tile_off).  Logical 32-bit word w of
// row r holds r * 1000 + w.  Copies: K step ks (32 B of every row) -> TMEM columns 8*ks .. 8*ks+7.  Expectation:
// TMEM[lane r][column c] == r * 1000 + c for c < 32 -- i.e. the A-operand layout weight_to_tmem() builds with tcgen05.st.
// Also times 16 copies (one 128 x 128 hi/lo weight pair) from issue to mbarrier completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mm-pde_b200/csrc -o utccp_probe profiles/experiments/utccp_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace mmpde::tc;

__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out, long long* clk) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(sm);
    const uint32_t bar = sbase + 4 * 16384;
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 4 * 16384 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(smem_u32(slot), 128);
    if (tid == 32) { mbar_init(bar, 1); fence_mbar_init(); }
    // four images (64 KB) so the timing copy has distinct sources; image 0 carries the test pattern
    for (int i = tid; i < 4 * 128 * 32; i += 128) {
        const int img = i / (128 * 32), r = (i / 32) % 128, w = i % 32;
        const uint32_t off = (uint32_t)img * 16384u + (uint32_t)r * 128u + ((((uint32_t)w >> 2) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)w & 3u) * 4u;
        *reinterpret_cast<uint32_t*>(sm + off) = (uint32_t)(r * 1000 + w + img * 100);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const long long t0 = clock64();
        for (int ks = 0; ks < 4; ++ks)
            utccp_128x256b(tmem + ks * 8, smem_desc_sw128(sbase + ks * 32, 16, 1024));
        umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        // timing: 16 copies = a hi/lo pair of [128][128] bf16 (64 KB)
        for (int q = 0; q < 16; ++q)
            utccp_128x256b(tmem + 32 + (q & 7) * 8, smem_desc_sw128(sbase + (q >> 2) * 16384 + (q & 3) * 32, 16, 1024));
        umma_commit(bar);
        mbar_wait(bar, 1);
        const long long t2 = clock64();
        clk[0] = t1 - t0; clk[1] = t2 - t1;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = v[c];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
    uint32_t* out; long long* clk;
    cudaMalloc(&out, 128 * 32 * 4); cudaMalloc(&clk, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 2048);
    probe<<<1, 128, 4 * 16384 + 2048>>>(out, clk);
    cudaError_t e = cudaDeviceSynchronize();
    printf("probe: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    static uint32_t h[128 * 32]; long long hc[2];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, clk, 16, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 32; ++c) bad += h[r * 32 + c] != (uint32_t)(r * 1000 + c);
    printf("tcgen05.cp.128x256b from a SWIZZLE_128B image: %d of 4096 words differ from lane*1000+column\n", bad);
    for (int r : {0, 1, 7, 8, 9, 31, 32, 127}) {
        printf("  lane %3d:", r);
        for (int c = 0; c < 12; ++c) printf(" %6u", h[r * 32 + c]);
        printf(" ...\n");
    }
    printf("4 copies + commit + wait: %lld clk;  16 copies (64 KB) + commit + wait: %lld clk\n", hc[0], hc[1]);
    return bad != 0;
}
