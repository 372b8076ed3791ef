// How fast can one SM (and the chip) add 512-byte fp32 rows into randomly chosen rows of a 38 MB matrix (dQ'[src] of the edge
// backward), and does the WIDTH of the per-lane reduction matter?
//   a. red.global.add.v4.f32 : one warp instruction = one 512-byte row            (what the row phase of edge_bwd does today)
//   b. red.global.add.f32    : one warp instruction = 128 contiguous bytes, 4 per row (what a thread-per-channel epilogue,
//                              registers = edges, would issue straight from tensor memory)
//   c. red.global.add.v2.f32 : 256 contiguous bytes per warp instruction
// Same bytes, same rows, 148 CTAs x 256 threads; reports us per 1.29 M rows (one edge_bwd launch at C1 size).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_scatter_probe red_scatter_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_red(float* __restrict__ out, const int* __restrict__ rows, int rows_per_warp) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int* my = rows + (size_t)warp * rows_per_warp;
    const float v = 1.0f + lane;
    for (int r = 0; r < rows_per_warp; r += 8) {
        int idx[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) idx[k] = __ldg(my + r + k);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float* row = out + (size_t)idx[k] * 256 + 128;
            if (MODE == 0) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(row + lane * 4), "f"(v) : "memory");
            } else if (MODE == 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(row + q * 32 + lane), "f"(v) : "memory");
            } else {
#pragma unroll
                for (int q = 0; q < 2; ++q) asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(row + q * 64 + lane * 2), "f"(v) : "memory");
            }
        }
    }
}

int main() {
    const int N = 36864, E = 1290240, G = 148, W = 8;          // rows of the matrix, rows to add, CTAs, warps per CTA
    const int per_warp = ((E + G * W - 1) / (G * W) + 7) / 8 * 8;
    std::vector<int> h((size_t)G * W * per_warp);
    // source indices of a 35-NN graph on a 48x48 grid are near the target: emulate with a window of +-3 grid rows around a
    // target that advances every 35 entries (same locality as the real edge list), samples of 2304 nodes
    uint32_t s = 12345u;
    for (size_t e = 0; e < h.size(); ++e) {
        s = s * 1664525u + 1013904223u;
        const size_t tgt = (e / 35) % N;
        const int smp = (int)(tgt / 2304), loc = (int)(tgt % 2304);
        int nb = loc + (int)((s >> 8) % 7 - 3) * 48 + (int)((s >> 16) % 7 - 3);
        nb = nb < 0 ? 0 : (nb > 2303 ? 2303 : nb);
        h[e] = smp * 2304 + nb;
    }
    int* d_rows; float* d_out;
    cudaMalloc(&d_rows, h.size() * 4); cudaMalloc(&d_out, (size_t)N * 256 * 4);
    cudaMemcpy(d_rows, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0, (size_t)N * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[3] = {"red.v4 (512 B / warp instr)", "red.f32 (128 B / warp instr)", "red.v2 (256 B / warp instr)"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {                    // second repetition is the one reported
            cudaEventRecord(e0);
            for (int it = 0; it < 10; ++it) {
                if (mode == 0) k_red<0><<<G, W * 32>>>(d_out, d_rows, per_warp);
                if (mode == 1) k_red<1><<<G, W * 32>>>(d_out, d_rows, per_warp);
                if (mode == 2) k_red<2><<<G, W * 32>>>(d_out, d_rows, per_warp);
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("%-30s %8.1f us per %d rows  (%.0f GB/s of reduced bytes)\n", names[mode], ms * 100.f, G * W * per_warp,
                            (double)G * W * per_warp * 512 / (ms * 1e-4) / 1e9);
        }
    }
    // SM-side issue rate: 16 CTAs only (the L2 is then far from its limit), 4 / 8 warps per CTA, same rows per warp
    printf("SM-side rate, 16 CTAs: clocks per warp instruction per SM (1.9 GHz assumed)\n");
    for (int warps = 4; warps <= 8; warps *= 2)
        for (int mode = 0; mode < 2; ++mode) {
            const int rpw = 2048;
            float ms = 0.f;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k_red<0><<<16, warps * 32>>>(d_out, d_rows, rpw);
                else k_red<1><<<16, warps * 32>>>(d_out, d_rows, rpw);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double instr = (double)warps * rpw * (mode == 0 ? 1 : 4);
            printf("  %2d warps  %-8s %8.1f us   %6.1f clk / warp-RED / SM   %6.1f clk per 512-byte row / SM\n", warps, mode == 0 ? "red.v4" : "red.f32",
                   ms * 1e3, ms * 1e-3 * 1.9e9 / instr, ms * 1e-3 * 1.9e9 / ((double)warps * rpw));
        }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
