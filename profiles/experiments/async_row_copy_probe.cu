// Which asynchronous path can feed ~21 B/clk/SM of randomly gathered 512-byte rows (Q'[src], fp32) into shared memory?
// (tma_gather4_probe.cu measured 378 clk of TMA occupancy per gather4 message of 4 x 512 B = 5.4 B/clk/SM: too slow.)
//   a. gather4 with 128-byte and 1024-byte rows: is the cost per message, per row or per byte?
//   b. cp.async.bulk (1-D bulk copy, UBLKCP) of one 512-byte row per issuing lane
//   c. cp.async (LDGSTS) 16 bytes per lane, one row per warp instruction, double-buffered commit groups
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o async_row_copy_probe async_row_copy_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) exit(1);
    return (EncodeTiled)fn;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float lds_sum(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v.x + v.w;
}

constexpr int N_ROUNDS = 1024;             // rounds per CTA; a round = 4 messages (a) / 32 rows (b) / 16 rows per warp (c)

// ---- a. gather4, `cols` fp32 per row, 4 issuing lanes, 8 message slots
__global__ void __launch_bounds__(128, 1) k_gather4(const __grid_constant__ CUtensorMap map, const int* __restrict__ rows, int cols, long long* clocks, float* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const uint32_t ring = smem_u32(sm), bars = smem_u32(sm + 32768);
    int4* sidx = reinterpret_cast<int4*>(sm + 33792);
    const int* my = rows + (size_t)blockIdx.x * N_ROUNDS * 16;
    for (int i = threadIdx.x; i < N_ROUNDS * 4; i += blockDim.x) sidx[i] = *reinterpret_cast<const int4*>(my + 4 * i);
    if (threadIdx.x == 0) { for (int d = 0; d < 8; ++d) mbar_init(bars + 8 * d, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    float acc = 0.f;
    const uint32_t msg = 4u * cols * 4u;
    const long long t0 = clock64();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int r = 0; r < N_ROUNDS + 2; ++r) {
            if (r >= 2) {
                for (int j = 0; j < 4; ++j) {
                    const int m = (r - 2) * 4 + j, d = m & 7;
                    int guard = 0;
                    while (!mbar_try(bars + 8 * d, (m >> 3) & 1) && ++guard < 100000) {}
                    acc += lds_sum(ring + d * 4096 + (lane * 16) % msg);
                }
                __syncwarp();
            }
            if (r < N_ROUNDS && lane < 4) {
                const int m = r * 4 + lane, d = m & 7;
                const int4 q = sidx[m];
                mbar_expect(bars + 8 * d, msg);
                gather4(ring + d * 4096, &map, 0, q.x, q.y, q.z, q.w, bars + 8 * d);
            }
            __syncwarp();
        }
    }
    if (threadIdx.x == 0) clocks[blockIdx.x] = clock64() - t0;
    if (acc == 12345.678f) sink[0] = acc;
}

// ---- b. one 512-byte bulk copy per lane and round (32 rows per round), two stages of 32 rows
__global__ void __launch_bounds__(128, 1) k_bulk(const float* __restrict__ pq, const int* __restrict__ rows, int issuers, long long* clocks, float* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const uint32_t ring = smem_u32(sm), bars = smem_u32(sm + 32768);     // 2 stages x 32 rows x 512 B
    int* sidx = reinterpret_cast<int*>(sm + 33792);
    const int* my = rows + (size_t)blockIdx.x * N_ROUNDS * 32;
    for (int i = threadIdx.x; i < N_ROUNDS * 32; i += blockDim.x) sidx[i] = my[i];
    if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    float acc = 0.f;
    const long long t0 = clock64();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int r = 0; r < N_ROUNDS + 1; ++r) {
            if (r < N_ROUNDS) {
                const int st = r & 1;
                if (lane == 0) mbar_expect(bars + 8 * st, issuers * 512);
                __syncwarp();
                if (lane < issuers)
                    bulk_g2s(ring + st * 16384 + lane * 512, pq + (size_t)sidx[r * 32 + lane] * 256 + 128, 512, bars + 8 * st);
            }
            if (r >= 1) {
                const int st = (r - 1) & 1;
                int guard = 0;
                while (!mbar_try(bars + 8 * st, ((r - 1) >> 1) & 1) && ++guard < 100000) {}
                for (int k = 0; k < issuers; k += 8) acc += lds_sum(ring + st * 16384 + k * 512 + lane * 16);
                __syncwarp();
            }
        }
    }
    if (threadIdx.x == 0) clocks[blockIdx.x] = clock64() - t0;
    if (acc == 12345.678f) sink[0] = acc;
}

// ---- c. cp.async 16 B per lane: `warps` warps, each copies 16 rows per round into its own slot (2 stages), reads them back
__global__ void __launch_bounds__(512, 1) k_ldgsts(const float* __restrict__ pq, const int* __restrict__ rows, long long* clocks, float* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const uint32_t ring = smem_u32(sm) + warp * (2 * 16 * 512);
    const int* my = rows + ((size_t)blockIdx.x * warps + warp) * N_ROUNDS * 16;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    int idx = (lane < 16) ? my[lane] : 0;
    for (int r = 0; r < N_ROUNDS + 1; ++r) {
        if (r < N_ROUNDS) {
            const int idx_n = (lane < 16 && r + 1 < N_ROUNDS) ? __ldg(my + (r + 1) * 16 + lane) : 0;
            const uint32_t dst = ring + (r & 1) * 8192 + lane * 16;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int s = __shfl_sync(0xffffffffu, idx, k);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k * 512), "l"(pq + (size_t)s * 256 + 128 + lane * 4) : "memory");
            }
            idx = idx_n;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (r >= 1) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            const uint32_t src = ring + ((r - 1) & 1) * 8192 + lane * 16;
#pragma unroll
            for (int k = 0; k < 16; ++k) acc += lds_sum(src + k * 512);
            __syncwarp();
        }
    }
    if (threadIdx.x == 0) clocks[blockIdx.x] = clock64() - t0;
    if (acc == 12345.678f) sink[0] = acc;
}

int main() {
    const int n_rows = 36864, ld = 256, grid = 148;
    float* pq;
    cudaMalloc(&pq, (size_t)n_rows * ld * 4);
    cudaMemset(pq, 0, (size_t)n_rows * ld * 4);
    std::vector<int> hr((size_t)grid * 16 * N_ROUNDS * 16);
    srand(1);
    for (auto& r : hr) r = rand() % n_rows;
    int* dr; float* sink; long long* clk;
    cudaMalloc(&dr, hr.size() * 4); cudaMalloc(&sink, 4); cudaMalloc(&clk, grid * 8);
    cudaMemcpy(dr, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice);
    EncodeTiled enc = get_encode();
    auto report = [&](const char* what, double rows_per_cta, double bytes_per_row) {
        cudaError_t e = cudaDeviceSynchronize();
        long long c0 = 0;
        cudaMemcpy(&c0, clk, 8, cudaMemcpyDeviceToHost);
        printf("%-58s %s  %8.1f clk/row  %6.2f B/clk/SM  (%.0f GB/s chip at 1.9 GHz)\n", what, cudaGetErrorString(e), (double)c0 / rows_per_cta,
               rows_per_cta * bytes_per_row / (double)c0, rows_per_cta * bytes_per_row / (double)c0 * 148 * 1.9);
        if (e != cudaSuccess) exit(2);
    };
    const int colsv[3] = {32, 128, 256};
    for (int cols : colsv) {
        CUtensorMap map;
        cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)n_rows};
        cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {(cuuint32_t)cols, 1}, estr[2] = {1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, pq, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 3; }
        const int smem = 33792 + N_ROUNDS * 4 * 16;
        cudaFuncSetAttribute(k_gather4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_gather4<<<grid, 128, smem>>>(map, dr, cols, clk, sink);
        k_gather4<<<grid, 128, smem>>>(map, dr, cols, clk, sink);
        char buf[96];
        snprintf(buf, sizeof buf, "a. gather4, %4d-byte rows, 4 issuing lanes", cols * 4);
        report(buf, N_ROUNDS * 16.0, cols * 4.0);
    }
    {
        const int smem = 33792 + N_ROUNDS * 32 * 4;
        cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int issuers = 8; issuers <= 32; issuers *= 2) {
            k_bulk<<<grid, 128, smem>>>(pq, dr, issuers, clk, sink);
            k_bulk<<<grid, 128, smem>>>(pq, dr, issuers, clk, sink);
            char buf[96];
            snprintf(buf, sizeof buf, "b. cp.async.bulk 512-byte rows, %2d issuing lanes per round", issuers);
            report(buf, N_ROUNDS * (double)issuers, 512.0);
        }
    }
    for (int warps : {2, 4, 8, 12}) {
        const int smem = warps * 2 * 16 * 512;
        cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_ldgsts<<<grid, warps * 32, smem>>>(pq, dr, clk, sink);
        k_ldgsts<<<grid, warps * 32, smem>>>(pq, dr, clk, sink);
        char buf[96];
        snprintf(buf, sizeof buf, "c. cp.async 16 B/lane (LDGSTS), %2d warps x 16 rows per round", warps);
        report(buf, N_ROUNDS * 16.0 * warps, 512.0);
    }
    return 0;
}
