// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, TMA bulk copy, TMEM
// allocation, tcgen05.mma / commit / ld, shared-memory matrix descriptors and the split-bf16 helpers.
// Bit layouts follow the PTX ISA tcgen05 descriptor tables (same fields as CUTLASS' cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace mmpde {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Wait on a phase parity.  try_wait suspends the thread in hardware for a bounded time, so the loop rarely
// iterates; the clock is only read every 256 failed polls.  A wait that outlives ~4 s of SM clocks can only be a
// protocol bug: trap (the launch fails with an error) instead of hanging the GPU.
// (A leaner loop -- one try_wait per iteration with an iteration-count watchdog -- measured 25 % SLOWER on the
// backward edge kernel: the wake-up from the longer hardware suspend is slower than this poll.)
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (true) {
#pragma unroll 1
        for (int k = 0; k < 256; ++k)
            if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}

// ---- TMA (bulk, 1-D): contiguous global -> shared, completion on an mbarrier ---------------------
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane of its warp's quadrant)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- TMEM <-> registers, asynchronous forms ---------------------------------------------------------
// tcgen05.ld without the wait: the destination registers are only valid after tmem_wait_ld(), which takes them
// as in/out operands so that no use can be scheduled above the wait.
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
// 8-column variants (small epilogue loop bodies that stay inside the instruction cache)
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
// 16-column variant
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns <- 16 registers per thread
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// the same 32-bit value into 32 consecutive columns of this thread's lane
__device__ __forceinline__ void tmem_fill32(uint32_t taddr, uint32_t v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
        "r"(v)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors -------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B (layout_type 2), descriptor version 1 (sm_100).
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
//   [46,48) version, [61,64) layout type.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, dense, no negate.
//   [4,6) D format (1 = f32), [7,10) A format (1 = bf16), [10,13) B format, [15] A major (0 = K, 1 = MN),
//   [16] B major, [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with the A operand in TENSOR MEMORY (lane = row of A, one 32-bit column = two consecutive K elements);
// A from TMEM is always K-major.  Halves the shared-memory operand traffic of the MMA.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same two instructions with the shared-memory descriptors given as (low word, high word): the MMA-issuing
// thread shares its scheduler with busy warps, so its instruction count per MMA is what paces the tensor pipe.  Only the
// start-address field (bits 0..13 of the low word, in 16-byte units) changes between the MMAs of a tile, so a
// descriptor is "low word of the tile base + compile-time constant".
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi_sw128(uint32_t sbo_bytes) {     // SBO, version 1, SWIZZLE_128B
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ void umma_bf16_ts_lh(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared memory -> tensor memory, 128 lanes x 32 bytes per copy (8 TMEM columns); the source is described like an MMA operand
__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint32_t s_lo, uint32_t s_hi) {
    asm volatile(
        "{\n\t.reg .b64 ds;\n\t"
        "mov.b64 ds, {%1, %2};\n\t"
        "tcgen05.cp.cta_group::1.128x256b [%0], ds;\n\t}" ::"r"(taddr), "r"(s_lo), "r"(s_hi)
        : "memory");
}
// All MMAs / copies issued so far by this thread -> arrive (count 1) on the mbarrier when they have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- operand tiles ----------------------------------------------------------------------------------
// A [ROWS][128] bf16 operand tile is stored as two column blocks of 64 columns; inside a block row r owns 128
// contiguous bytes whose 16-byte chunks are XOR-swizzled with (r & 7)  (the TMA / UMMA SWIZZLE_128B pattern).
// The same image is a K-major operand (rows = M or N, columns = K) and an MN-major operand (columns = M or N,
// rows = K).  Tiles must start on a 1024-byte boundary.
constexpr uint32_t KBLK_BYTES = 128 * 128;   // 128 rows x 128 B

// byte offset of bf16 element (row, col), col < 128, in an operand image of ROWS rows
template <int ROWS = 128>
__device__ __forceinline__ uint32_t tile_off(int row, int col) {
    return (uint32_t)(col >> 6) * (uint32_t)(ROWS * 128) + (uint32_t)row * 128u +
           ((((uint32_t)(col & 63) >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)(col & 7) << 1);
}

// split-bf16: x = hi + lo (+ O(2^-17 |x|));  products hi*hi + hi*lo + lo*hi give ~fp32-grade GEMMs on the bf16 pipe
__device__ __forceinline__ uint32_t cvt_bf16x2(float first, float second) {      // first -> low half
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(second), "f"(first));
    return d;
}
__device__ __forceinline__ void split4(float4 v, uint2& hi, uint2& lo) {
    hi.x = cvt_bf16x2(v.x, v.y);
    hi.y = cvt_bf16x2(v.z, v.w);
    lo.x = cvt_bf16x2(v.x - __uint_as_float(hi.x << 16), v.y - __uint_as_float(hi.x & 0xFFFF0000u));
    lo.y = cvt_bf16x2(v.z - __uint_as_float(hi.y << 16), v.w - __uint_as_float(hi.y & 0xFFFF0000u));
}

// ---- explicit shared-space accesses (32-bit shared addresses: STS/LDS instead of generic ST/LD) --------------
// volatile without a memory clobber: they stay ordered among themselves and against the fence / mbarrier asm.
__device__ __forceinline__ void sts_v2(uint32_t saddr, uint2 v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(v.x), "r"(v.y));
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ uint4 lds_v4(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts_b16(uint32_t saddr, unsigned short v) { asm volatile("st.shared.b16 [%0], %1;" ::"r"(saddr), "h"(v)); }
__device__ __forceinline__ void sts_b32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v)); }
__device__ __forceinline__ uint32_t lds_b32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float4 lds_v4f(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

// fp32 row fragment (4 consecutive columns owned by this lane) -> bf16 hi / lo operand images (ROWS rows each, the lo
// image follows the hi image) at shared address `img`
template <int ROWS>
__device__ __forceinline__ void store_split(uint32_t img, int row, int lane, const float4& v) {
    uint2 hi, lo;
    split4(v, hi, lo);
    const uint32_t a = img + tile_off<ROWS>(row, lane * 4);
    sts_v2(a, hi);
    sts_v2(a + 2 * ROWS * 128, lo);
}

// W (fp32, element (r, k) at w[r*rs + k*ks], 128 x 128) -> TMEM A operand of the TS-form MMA: lane r, 64 columns hi at
// t_hi, 64 columns lo at t_lo (one 32-bit column = two consecutive k).  Called by the warps owning the 128 lanes; the
// four 32-column groups g0 <= g < g1 let two warps of the same lane quadrant share the work.
__device__ __forceinline__ void weight_to_tmem(const float* __restrict__ w, int64_t rs, int64_t ks, int r, uint32_t t_hi, uint32_t t_lo,
                                               int g0 = 0, int g1 = 4, float scale = 1.f) {
    const bool vec = (ks == 1) && ((rs & 3) == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);   // 128-bit loads need alignment
    const float* row = w + r * rs;
#pragma unroll 1
    for (int g = g0; g < g1; g += 2) {                         // two groups (64 values) in flight per round
        float4 x[16];
        if (vec) {
#pragma unroll
            for (int v = 0; v < 16; ++v) x[v] = ldg4(row + g * 32 + v * 4);
        } else {
            const float* q = row + (int64_t)(g * 32) * ks;
#pragma unroll
            for (int v = 0; v < 16; ++v) {
                x[v] = make_float4(__ldg(q), __ldg(q + ks), __ldg(q + 2 * ks), __ldg(q + 3 * ks));
                q += 4 * ks;
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                uint2 a, b;
                const float4 xs = make_float4(x[h * 8 + v].x * scale, x[h * 8 + v].y * scale, x[h * 8 + v].z * scale, x[h * 8 + v].w * scale);
                split4(xs, a, b);
                hi[2 * v] = a.x; hi[2 * v + 1] = a.y; lo[2 * v] = b.x; lo[2 * v + 1] = b.y;
            }
            tmem_st16(t_hi + (g + h) * 16, hi);
            tmem_st16(t_lo + (g + h) * 16, lo);
        }
    }
}

// ---- warp-role plumbing -----------------------------------------------------------------------------------
// Register budget moved between warpgroups (4 consecutive warps each); the CTA's pool is what it was launched with.
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

}  // namespace tc

// Optional per-phase timestamps (debug build only: make timeline -> libmmpde_b200_tl.so, read by profiles/timeline.py).
// Slot layout: [role][tile iteration < TL_ITERS][8 stamps]; roles: 0 first builder warp, 1 last builder warp, 2 MMA
// thread, 3 epilogue warp 0.  CTA 0 only.
#ifdef MMPDE_TIMELINE
static __device__ long long* g_timeline = nullptr;     // per translation unit (no -rdc): set via mmpde_debug_timeline
constexpr int TL_ITERS = 48;
#define TL(role, it, slot)                                                                                   \
    do {                                                                                                     \
        if (g_timeline != nullptr && blockIdx.x == 0 && (it) < TL_ITERS && (threadIdx.x & 31) == 0)          \
            g_timeline[((role) * TL_ITERS + (it)) * 8 + (slot)] = clock64();                                 \
    } while (0)
#else
#define TL(role, it, slot) do { } while (0)
#endif

}  // namespace mmpde
