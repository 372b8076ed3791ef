// Probe of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a (no public doc is reachable from the build box):
//   1. which box the tensor map needs ({cols, 1} or {cols, 4}), what lands where in shared memory, how many bytes the
//      mbarrier sees, what an out-of-range row index does;
//   2. sustained throughput of random 512-byte row gathers per SM with N messages in flight (the edge kernels need
//      ~21 B/clk/SM of Q' rows).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4_probe tma_gather4_probe.cu
// Every wait is bounded (no hang): a variant that never completes is reported as such.
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
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        printf("cuTensorMapEncodeTiled not found\n");
        exit(1);
    }
    return (EncodeTiled)fn;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

// ---- 1. semantics: one gather4 of rows (r0..r3) at column `col`, expected bytes `expect`; dumps 4 KB of smem
__global__ void probe_kernel(const __grid_constant__ CUtensorMap map, int col, int4 rows, uint32_t expect, float* dump, int* status) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float* buf = reinterpret_cast<float*>(sm);
    const uint32_t bar = smem_u32(sm + 8192);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = -7.f;
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect(bar, expect);
        gather4(smem_u32(buf), &map, col, rows.x, rows.y, rows.z, rows.w, bar);
        int ok = 0;
        for (int it = 0; it < 2000 && !ok; ++it) ok = mbar_try(bar, 0) ? 1 : 0;
        status[0] = ok;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) dump[i] = buf[i];
}

// ---- 2. throughput: every CTA keeps `depth` gather4 messages (2 KB each) in flight over a ring, random rows.
// Row indices are staged in shared memory first (a dependent global load per message would be what is measured);
// `issuers` lanes (1, 2 or 4) issue one message each per round, the way a builder warp would.
__global__ void __launch_bounds__(128, 1) bw_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ rows, int n_msgs,
                                                   int depth, int issuers, float* sink, long long* clocks) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const uint32_t ring = smem_u32(sm);                 // depth x 2 KB (<= 64 KB)
    const uint32_t bars = smem_u32(sm + 65536);         // depth mbarriers
    int4* sidx = reinterpret_cast<int4*>(sm + 66560);   // n_msgs x int4 (<= 64 KB)
    const int* my = rows + (size_t)blockIdx.x * n_msgs * 4;
    for (int i = threadIdx.x; i < n_msgs; i += blockDim.x) sidx[i] = *reinterpret_cast<const int4*>(my + 4 * i);
    if (threadIdx.x == 0) {
        for (int d = 0; d < depth; ++d) mbar_init(bars + 8 * d, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    const long long t0 = clock64();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        // round r issues messages r*issuers .. r*issuers+issuers-1 into slots (msg % depth); depth is a multiple of issuers
        const int rounds = n_msgs / issuers, lag = depth / issuers;
        for (int r = 0; r < rounds + lag; ++r) {
            if (r >= lag) {                              // consume the messages of round r - lag
                for (int j = 0; j < issuers; ++j) {
                    const int m = (r - lag) * issuers + j, d = m % depth;
                    const uint32_t par = (m / depth) & 1;
                    int guard = 0;
                    while (!mbar_try(bars + 8 * d, par) && ++guard < 100000) {}
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                 : "r"(ring + d * 2048 + lane * 16));
                    acc += v.x + v.w;
                }
                __syncwarp();
            }
            if (r < rounds && lane < issuers) {
                const int m = r * issuers + lane, d = m % depth;
                const int4 q = sidx[m];
                mbar_expect(bars + 8 * d, 2048);
                gather4(ring + d * 2048, &map, 128, q.x, q.y, q.z, q.w, bars + 8 * d);
            }
            __syncwarp();
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
    if (acc == 12345.678f) sink[0] = acc;
}

int main() {
    const int n_rows = 300000, ld = 256;               // PQ [n_rows, 256] fp32, Q' = columns 128..255
    float* pq;
    cudaMalloc(&pq, (size_t)n_rows * ld * 4);
    std::vector<float> h((size_t)n_rows * ld);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i / ld) + 0.001f * (float)(i % ld);     // row + col/1000
    cudaMemcpy(pq, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    EncodeTiled enc = get_encode();
    float* dump; int* status;
    cudaMalloc(&dump, 8192); cudaMalloc(&status, 16);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    for (int box_rows = 1; box_rows <= 1; box_rows += 3) {   // {128,4} is what a plain tile load would use; gather4 takes ONE-row boxes
        CUtensorMap map;
        cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)n_rows};
        cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {128, (cuuint32_t)box_rows}, estr[2] = {1, 1};
        CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, pq, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("== box {128,%d}: encode rc=%d\n", box_rows, (int)rc);
        if (rc != CUDA_SUCCESS) continue;
        const int4 cases[3] = {{5, 1000, 7, 299999}, {3, 3, 2, 1}, {10, 400000, -1, 20}};     // last: out-of-range rows
        for (int c = 0; c < 3; ++c) {
            for (uint32_t expect = 2048; expect <= 8192; expect *= 4) {
                cudaMemset(status, 0, 16);
                probe_kernel<<<1, 128, 16384>>>(map, 128, cases[c], expect, dump, status);
                cudaError_t e = cudaDeviceSynchronize();
                int st = -1;
                std::vector<float> d(2048);
                if (e == cudaSuccess) { cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost); cudaMemcpy(d.data(), dump, 8192, cudaMemcpyDeviceToHost); }
                printf("   rows (%d,%d,%d,%d) expect_tx %u: cuda=%s barrier_done=%d |", cases[c].x, cases[c].y, cases[c].z, cases[c].w,
                       expect, cudaGetErrorString(e), st);
                for (int k = 0; k < 8; ++k) printf(" [%d]=%.3f", k * 128, d[k * 128]);
                printf(" [127]=%.3f [511]=%.3f [2047]=%.3f\n", d[127], d[511], d[2047]);
                if (e != cudaSuccess) { printf("   CUDA error: stopping\n"); return 2; }
            }
        }
    }
    // ---- throughput with the {128,1} map
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)n_rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {128, 1}, estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, pq, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 3;
    const int n_msgs = 4096, grid = 148;      // 4096 x 16 B of indices = 64 KB of shared memory
    std::vector<int> hr((size_t)grid * n_msgs * 4);
    srand(1);
    for (int pass = 0; pass < 2; ++pass) {               // pass 0: rows from a 36 864-row window (L2-resident, C1), pass 1: all rows
        const int window = pass == 0 ? 36864 : n_rows;
        for (auto& r : hr) r = rand() % window;
        int* dr; float* sink; long long* clk;
        cudaMalloc(&dr, hr.size() * 4); cudaMalloc(&sink, 4); cudaMalloc(&clk, grid * 8);
        cudaMemcpy(dr, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice);
        const int smem_bw = 66560 + n_msgs * 16;
        cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bw);
        for (int issuers = 1; issuers <= 4; issuers *= 2)
        for (int depth = 4; depth <= 32; depth *= 2) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            bw_kernel<<<grid, 128, smem_bw>>>(map, dr, n_msgs, depth, issuers, sink, clk);
            cudaEventRecord(e0);
            bw_kernel<<<grid, 128, smem_bw>>>(map, dr, n_msgs, depth, issuers, sink, clk);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            long long c0 = 0;
            cudaMemcpy(&c0, clk, 8, cudaMemcpyDeviceToHost);
            printf("gather4 throughput, window %d rows, %d issuing lanes, depth %2d: %s  %.3f ms  %.1f GB/s chip  %.1f B/clk/SM (CTA0 %lld clk, %.0f clk/msg)\n",
                   window, issuers, depth, cudaGetErrorString(e), ms, (double)grid * n_msgs * 2048 / (ms * 1e-3) / 1e9,
                   (double)n_msgs * 2048 / (double)c0, c0, (double)c0 / n_msgs);
            if (e != cudaSuccess) return 4;
        }
        cudaFree(dr); cudaFree(sink); cudaFree(clk);
    }
    return 0;
}
