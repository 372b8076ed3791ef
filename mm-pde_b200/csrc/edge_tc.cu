// Message passing over the target-sorted edge list on the 5th-gen tensor cores (tcgen05 + TMEM).
// Replaces PyG propagate + message_net_1/2 + scatter-mean (/root/reference/gnn_2d.py:55,59-63) and its autograd.
//
// Formulation.  message_net_1 is split per node: with e_ij = (u_i-u_j, px_i-px_j, py_i-py_j, v_i)
//   z1_ij = W1a x_i + W1b x_j + W1c e_ij + b1 = P'[i] + Q'[j],
//   P'[i] = W1a x_i + W1c (u,px,py,v)_i + b1,      Q'[j] = W1b x_j - W1c[:, :3] (u,px,py)_j
// (both node-level GEMMs, done by the caller), so an edge only costs h1 = relu(P'[dst] + Q'[src]) before the one
// dense per-edge contraction  z2 = W2 h1 + b2.  That contraction runs on tcgen05 as three bf16 products
// (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM: ~2^-16 relative error per term).
//
// Both kernels are persistent (one CTA per SM) and warp-specialised:
//   epilogue warps : thread = TMEM lane = channel; tcgen05.ld, bias, ReLU / masks / per-target sums
//                    (forward: 16 warps, lane quadrant x 32-column quarter; backward: 4 warps)
//   builder warps  : the Q'[src] rows arrive through a cp.async ring in shared memory (requested half a tile / a tile
//     (8)            ahead; no register prefetch), P'[dst] is added, ReLU, split into bf16 hi/lo, written as the
//                    SWIZZLE_128B operand tile -- the neighbour gather IS the operand load of the GEMM, no [E,260] /
//                    [E,128] tensor ever exists; in the backward they also run the row phase
//   MMA warp       : one thread issues tcgen05.mma; W2 (hi, lo) lives in TENSOR MEMORY as the A operand for the whole
//                    kernel, so only the per-tile operand is read from shared memory
// connected by mbarrier pipelines (operand tiles and accumulators are double-buffered); registers are moved between the
// roles with setmaxnreg (forward 896 threads: 64 / 104 / 40 per thread; backward 512 threads: 104 / 184 / 40).
#include "tc_common.cuh"

namespace mmpde {
using namespace tc;

constexpr int EPI_WARPS = 4;
constexpr int BLD_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS + BLD_WARPS;
constexpr int EDGE_THREADS = 512;                          // 4 warpgroups: epilogue | builders | builders | MMA + 3 idle
// Register budget moved between the warpgroups with setmaxnreg (launch value 65536 / 512 = 128 per thread):
// 4*32*104 + 8*32*184 + 4*32*40 = 65536.
constexpr int EPI_REGS = 104, BLD_REGS = 184, MMA_REGS = 40;
constexpr uint32_t TMEM_COLS = 512;

// ---- rows a builder warp owns --------------------------------------------------------------------------------
// lane r (< ROWS) of a builder warp holds (dst, src) of row  warp_row0 + r  of a tile; -1 beyond the last edge.
struct RowIdx { int d, s; };
template <int ROWS>
__device__ __forceinline__ RowIdx load_row_idx(const int* __restrict__ dst, const int* __restrict__ src, int64_t e0,
                                               int row0, int64_t n_edges, bool valid_tile) {
    RowIdx r; r.d = -1; r.s = -1;
    const int lane = threadIdx.x & 31;
    const int64_t e = e0 + row0 + lane;
    if (valid_tile && lane < ROWS && e < n_edges) { r.d = __ldg(dst + e); r.s = __ldg(src + e); }
    return r;
}

__device__ __forceinline__ float4 sel4(bool c, const float4& a, const float4& b) {
    return make_float4(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z, c ? a.w : b.w);
}
// mean message of one finished target: agg[cur][o] += run / deg (partial runs of a target add up atomically)
__device__ __noinline__ void flush_mean(float* agg, int64_t ld, int cur, float inv, int o, float run) {
    if (cur >= 0) atomicAdd(agg + (int64_t)cur * ld + o, run * inv);
}
__device__ __forceinline__ int lds_i32(uint32_t saddr) { return (int)lds_b32(saddr); }

// ================================================================================================================
// Forward warp roles (896 threads): 16 epilogue warps (TMEM lane quadrant = warp & 3, 32-column quarter = warp >> 2),
// 8 builder warps, 1 MMA warp (+ 3 idle to fill the warpgroup).  Every role is an instruction-issue-bound stream, so
// what counts is the number of instructions per edge: the Q'[src] rows are gathered ASYNCHRONOUSLY into a shared-memory
// ring (cp.async, 16 bytes per lane = one row per warp instruction, half a tile ahead: no register prefetch state, no
// scoreboard stalls in the builders), element-wise math uses the packed fp32x2 forms, and the bias lives in the
// accumulator (tcgen05.st) instead of an add per element.
// Why cp.async and not the TMA: measured on a B200 (profiles/r02_tma_gather4_probe.txt, r02_async_row_copy_probe.txt)
// the TMA spends ~60 clk PER ROW on tile::gather4 (8.6 B/clk/SM for 512-byte rows; 68-100 clk per row for 1-D bulk
// copies), cp.async sustains a row every 10-12 clk (42-51 B/clk/SM) -- the edge kernels need ~25.
// Register budget per thread moved with setmaxnreg from the launch value 72 (the CTA's pool is what it was launched
// with: 896*72 = 64512): 16*32*64 + 8*32*104 + 4*32*40 = 64512 (the builders hold no prefetched rows any more).
constexpr int F_EPI_WARPS = 16, F_BLD_WARPS = 8, F_MMA_WARP = F_EPI_WARPS + F_BLD_WARPS, F_THREADS = 896;
constexpr int F_EPI_REGS = 64, F_BLD_REGS = 104, F_MMA_REGS = 40;

// ---- packed fp32x2 (sm_100: FADD2 -- one issue slot for two fp32 additions) ---------------------------------------------
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// One operand row: 2*h1 = 2*relu(p + q) = z + |z| of this lane's 4 channels, split into bf16 hi / lo and written into
// the two SWIZZLE_128B images (the lo image follows the hi image).  z + |z| is two adds on the FMA pipe instead of an add
// plus a max on the ALU pipe, the busiest unit of these kernels; the factor 2 is exact: the forward folds 1/2 into W2 when
// it loads it into tensor memory, the backward into the final dW2 reduction.  (Callers select p with sel4, never with a branch: a branch per row would fence the rows of a warp
// off from each other and serialise their dependency chains.)
template <int ROWS>
__device__ __forceinline__ void build_row(uint32_t img, int row, int lane, const float4& p, const float4& q) {
    float z0, z1, z2, z3, l0, l1, l2, l3;
    unpack2(add2(pack2(p.x, p.y), pack2(q.x, q.y)), z0, z1);
    unpack2(add2(pack2(p.z, p.w), pack2(q.z, q.w)), z2, z3);
    const float r0 = z0 + fabsf(z0), r1 = z1 + fabsf(z1), r2 = z2 + fabsf(z2), r3 = z3 + fabsf(z3);
    const uint32_t h01 = cvt_bf16x2(r0, r1), h23 = cvt_bf16x2(r2, r3);
    unpack2(sub2(pack2(r0, r1), pack2(__uint_as_float(h01 << 16), __uint_as_float(h01 & 0xFFFF0000u))), l0, l1);
    unpack2(sub2(pack2(r2, r3), pack2(__uint_as_float(h23 << 16), __uint_as_float(h23 & 0xFFFF0000u))), l2, l3);
    const uint32_t a = img + tile_off<ROWS>(row, lane * 4);
    sts_v2(a, make_uint2(h01, h23));
    sts_v2(a + 2 * ROWS * 128, make_uint2(cvt_bf16x2(l0, l1), cvt_bf16x2(l2, l3)));
}

// cp.async: 16 bytes global -> shared without passing through registers; src_bytes = 0 zero-fills (rows past the last edge)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// One cp.async group = the operands of 8 consecutive rows of a builder warp (lane r0+k holds target / source of row k):
// Q'[src] of the 8 rows -> `rows` (8 x 512 B) and P'[dst] of the first and the last row -> `prow` (2 x 512 B).  A target's
// rows are consecutive, so 8 rows span <= 2 targets unless some in-degree is < 4 (such rows are patched at build time).
// Every lane copies -- and later reads back -- ITS 16 bytes of each row, so the copies need no cross-lane synchronisation.
__device__ __forceinline__ void gather_half_async(uint32_t rows, uint32_t prow, const float* __restrict__ PQ, RowIdx idx, int r0, int lane) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int s = __shfl_sync(0xffffffffu, idx.s, r0 + k);
        cp_async16(rows + k * 512 + lane * 16, PQ + (int64_t)(s >= 0 ? s : 0) * 256 + 128 + lane * 4, s >= 0 ? 16u : 0u);
    }
    const int da = __shfl_sync(0xffffffffu, idx.d, r0), db = __shfl_sync(0xffffffffu, idx.d, r0 + 7);
    cp_async16(prow + lane * 16, PQ + (int64_t)(da >= 0 ? da : 0) * 256 + lane * 4, da >= 0 ? 16u : 0u);
    cp_async16(prow + 512 + lane * 16, PQ + (int64_t)(db >= 0 ? db : 0) * 256 + lane * 4, db >= 0 ? 16u : 0u);
    cp_async_commit();
}
// only the 8 Q' rows (backward: the per-target operands P', g_agg, 1/deg, mask words are prefetched in registers)
__device__ __forceinline__ void gather_q8_async(uint32_t rows, const float* __restrict__ PQ, int s_lane, int lane) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int s = __shfl_sync(0xffffffffu, s_lane, k);
        cp_async16(rows + k * 512 + lane * 16, PQ + (int64_t)(s >= 0 ? s : 0) * 256 + 128 + lane * 4, s >= 0 ? 16u : 0u);
    }
    cp_async_commit();
}
// 8 gathered Q' rows (fp32, 512 bytes each, in the shared-memory ring at `rows`) + P' of their (<= 2) targets -> operand
// image rows rowbase .. rowbase+7; then the rare patch loop (rows of a third target, listed in `odd`).
template <int ROWS>
__device__ __forceinline__ void build8_core(uint32_t img, int rowbase, uint32_t rows, const float4& pa, const float4& pb, int da,
                                            uint32_t odd, RowIdx idx, int r0, const float* __restrict__ PQ, int lane) {
    // two groups of four rows: the shared-memory loads / stores are volatile asm and keep their order, so at most four
    // rows (and their temporaries) are live at a time
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = lds_v4f(rows + (4 * h + k) * 512 + lane * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int d = __shfl_sync(0xffffffffu, idx.d, r0 + 4 * h + k);
            build_row<ROWS>(img, rowbase + 4 * h + k, lane, sel4(d == da, pa, pb), q[k]);
        }
    }
    for (uint32_t o = odd; o != 0u; o &= o - 1u) {
        const int k = __ffs(o) - 1;
        const int d = __shfl_sync(0xffffffffu, idx.d, r0 + k), sr = __shfl_sync(0xffffffffu, idx.s, r0 + k);
        build_row<ROWS>(img, rowbase + k, lane, ldg4(PQ + (int64_t)d * 256 + lane * 4), ldg4(PQ + (int64_t)sr * 256 + 128 + lane * 4));
    }
}
// the same with the P' rows taken from the shared-memory slot gather_half_async filled
template <int ROWS>
__device__ __forceinline__ void build8(uint32_t img, int rowbase, uint32_t rows, uint32_t prow, RowIdx idx, int r0,
                                       const float* __restrict__ PQ, int lane) {
    const int da = __shfl_sync(0xffffffffu, idx.d, r0), db = __shfl_sync(0xffffffffu, idx.d, r0 + 7);
    const uint32_t odd = (__ballot_sync(0xffffffffu, idx.d >= 0 && idx.d != da && idx.d != db) >> r0) & 0xFFu;
    const float4 pa = lds_v4f(prow + lane * 16), pb = lds_v4f(prow + 512 + lane * 16);
    build8_core<ROWS>(img, rowbase, rows, pa, pb, da, odd, idx, r0, PQ, lane);
}

// ================================================================================================================
// Forward:  agg[i] = mean_{e: dst=i} relu(W2 relu(P'[i] + Q'[src_e]) + b2);  mask2 = sign bits of z2.
// Tile = 128 edges.  D[o][e] = -b2[o] - sum_c W2[o][c] h1[e][c] = -z2  (M = 128 channels on TMEM lanes, N = 128 edges):
// the NEGATED pre-activation, so that the sign bit of an accumulator element IS the mask bit (z2 > 0; a sum that
// cancels to zero is +0 in the accumulator, so z2 = 0 gives 0 like relu'(0)) and no negation is needed per element.
// ================================================================================================================
constexpr int FTE = 128;
constexpr uint32_t F_IMG = 2 * FTE * 128;                  // one [128][128] bf16 image (2 column blocks) = 32 KB
struct EdgeFwdArgs {
    const float* PQ; const int* src; const int* dst; const float* inv_deg; int64_t n_edges;
    const float* w2; const float* b2; float* agg; int64_t ld_agg; uint32_t* mask2;
};
struct FwdSmem {
    // Targets of the tile's edges for the epilogue, slot = tile iteration & 7.  The epilogue still reads slot i after it
    // has released accumulator stage i (boundary targets of its last piece); the builders may then run up to tile i+4
    // (operand stage free <=> MMA of tile i+2 done <=> accumulator stage of tile i released), so 8 slots keep the
    // slot being read and the slot being written apart by construction, not by timing.
    static constexpr int SLOTS = 8;
    static constexpr uint32_t H = 0;                       // 2 stages x (hi, lo)
    static constexpr uint32_t RING = 2 * 2 * F_IMG;        // Q' rows of one tile: 8 builder warps x 16 rows x 512 B (cp.async destination)
    static constexpr uint32_t PROW = RING + FTE * 512;     // P' rows: 8 builder warps x 2 halves x 2 rows x 512 B
    static constexpr uint32_t DST = PROW + 8 * 2 * 1024;   // int dst[SLOTS][128]
    static constexpr uint32_t INV = DST + SLOTS * FTE * 4; // float inv_deg[dst][SLOTS][128]
    static constexpr uint32_t BAR = INV + SLOTS * FTE * 4; // h_full[2] h_empty[2] tm_full[2] tm_empty[2], tmem slot
    static constexpr uint32_t TOTAL = BAR + 128;
};

__global__ void __launch_bounds__(F_THREADS, 1) edge_fwd_tc_kernel(EdgeFwdArgs p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(sm);
    const uint32_t bar0 = sbase + FwdSmem::BAR;
    const uint32_t h_full = bar0, h_empty = bar0 + 16, tm_full = bar0 + 32, tm_empty = bar0 + 48;   // [b] at +8*b
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + FwdSmem::BAR + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    if (tid == 32) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(h_full + 8 * b, F_BLD_WARPS); mbar_init(h_empty + 8 * b, 1);
            mbar_init(tm_full + 8 * b, 1); mbar_init(tm_empty + 8 * b, F_EPI_WARPS);
        }
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d = tmem_base, tmem_w_hi = tmem_base + 256, tmem_w_lo = tmem_base + 320;
    const int64_t n_tiles = (p.n_edges + FTE - 1) / FTE;

    if (warp < F_EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue: thread = out-channel o
        reg_dec<F_EPI_REGS>();
        const int q = warp & 3, quarter = warp >> 2;                       // TMEM lane quadrant, 32-column quarter of the tile
        const int o = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t nbias = __float_as_uint(0.f - __ldg(p.b2 + o));     // accumulators start at -b2 (never -0)
        if (quarter < 2)                                                   // -W2/2 (the operand tile holds 2*h1) -> tensor memory
            weight_to_tmem(p.w2, 128, 1, o, tmem_w_hi + lane_addr, tmem_w_lo + lane_addr, 2 * quarter, 2 * quarter + 2, -0.5f);
        tmem_fill32(tmem_d + lane_addr + quarter * 32, nbias);
        tmem_fill32(tmem_d + lane_addr + FTE + quarter * 32, nbias);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(tm_empty); mbar_arrive(tm_empty + 8); }
        int i = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int b = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            const uint32_t dsts = sbase + FwdSmem::DST + (uint32_t)(i & (FwdSmem::SLOTS - 1)) * (FTE * 4);
            mbar_wait(tm_full + 8 * b, ph);
            if (warp == 0) TL(3, i, 0);
            tc_fence_after();
            const uint32_t d_addr = tmem_d + lane_addr + b * FTE + quarter * 32;
            int cur = -1;
            uint64_t run_s = 0ull;                                         // packed pair of partial sums of a = -z2
            float run_a = 0.f, cur_inv = 0.f;                              // sum |a| of the current run
            // The thread's 32 edges in 4 pieces of 8 (a rolled loop over small tcgen05.ld.x8 pieces keeps the body short).
            // Per edge: running per-target sum of relu(z2) and the sign mask, kept off the busy ALU pipe:
            //   sum relu(z2) = (sum |a| - sum a) / 2   -> 1.5 adds per edge (the sum of a as packed pairs);
            //   mask bit     = sign bit of a           -> one funnel shift.
            // Targets are contiguous runs of edges; `bm` marks the first edge of each run (warp-uniform), so groups of 4
            // edges without a boundary take the short path.  A warp's 32 columns always start a new run (partial runs
            // add up atomically).
            const int e_first = quarter * 32;
            const int d_me = lds_i32(dsts + 4 * (e_first + lane));
            const int d_pv = (lane > 0) ? lds_i32(dsts + 4 * (e_first + lane) - 4) : -2;
            const uint32_t bm = __ballot_sync(0xffffffffu, d_me != d_pv);
            uint32_t word = 0;                                             // filled MSB-first: bit 31-j <- edge j
            auto run_sum = [&]() { float s0, s1; unpack2(run_s, s0, s1); return 0.5f * (run_a - (s0 + s1)); };
            auto process8 = [&](uint32_t (&vc)[8], int pc) {              // piece pc: columns e_first + 8*pc .. +7
                const int e0 = e_first + pc * 8;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const float a0 = __uint_as_float(vc[4 * g]), a1 = __uint_as_float(vc[4 * g + 1]),
                                a2 = __uint_as_float(vc[4 * g + 2]), a3 = __uint_as_float(vc[4 * g + 3]);
                    const uint32_t nib = (bm >> (pc * 8 + 4 * g)) & 15u;
                    if (nib == 0u) {
                        run_s = add2(add2(run_s, pack2(a0, a1)), pack2(a2, a3));
                        run_a += (fabsf(a0) + fabsf(a1)) + (fabsf(a2) + fabsf(a3));
                    } else {
                        const float a[4] = {a0, a1, a2, a3};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (nib & (1u << k)) {
                                flush_mean(p.agg, p.ld_agg, cur, cur_inv, o, run_sum());
                                cur = lds_i32(dsts + 4 * (e0 + 4 * g + k));                   // used at the NEXT flush:
                                cur_inv = __uint_as_float(lds_b32(dsts + FwdSmem::SLOTS * FTE * 4 + 4 * (e0 + 4 * g + k)));
                                run_s = 0ull;                                                 // latency stays hidden
                                run_a = 0.f;
                            }
                            run_s = add2(run_s, pack2(a[k], 0.f));
                            run_a += fabsf(a[k]);
                        }
                    }
                    word = __funnelshift_l(vc[4 * g], word, 1);
                    word = __funnelshift_l(vc[4 * g + 1], word, 1);
                    word = __funnelshift_l(vc[4 * g + 2], word, 1);
                    word = __funnelshift_l(vc[4 * g + 3], word, 1);
                }
            };
            uint32_t va[8], vb[8];
            tmem_ld8_async(d_addr, va);
#pragma unroll 1
            for (int pc = 0; pc < 4; pc += 2) {
                tmem_wait_ld8(va);
                tmem_ld8_async(d_addr + (pc + 1) * 8, vb);
                process8(va, pc);
                tmem_wait_ld8(vb);
                if (pc + 2 < 4) {
                    tmem_ld8_async(d_addr + (pc + 2) * 8, va);
                } else {                                                   // all 32 columns are in registers: back to -b2, release
                    tmem_fill32(d_addr, nbias);
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tm_empty + 8 * b);
                    if (warp == 0) TL(3, i, 1);
                }
                process8(vb, pc + 1);
            }
            // mask2[chunk of 32 edges][channel]: bit j = (z2 > 0) of edge 32*chunk + j
            p.mask2[((t * 4 + quarter) * 128) + o] = __brev(word);
            flush_mean(p.agg, p.ld_agg, cur, cur_inv, o, run_sum());
            if (warp == 0) TL(3, i, 2);
        }
    } else if (warp < F_MMA_WARP) {
        // ------------------------------------------------------------------ builders: warp w -> rows 16w .. 16w+15
        reg_inc<F_BLD_REGS>();
        const int w = warp - F_EPI_WARPS;
        const int row0 = w * 16;
        const uint32_t ring = sbase + FwdSmem::RING + (uint32_t)w * (16 * 512);      // this warp's 16 Q' rows
        const uint32_t prow = sbase + FwdSmem::PROW + (uint32_t)w * 2048;            // and the P' rows of its two halves
        const int64_t G = gridDim.x;
        // Indices run two tiles ahead of the build; the operands of a half tile (one cp.async group) are requested as soon
        // as the same half of the previous tile has been consumed.  Nothing is prefetched into registers.
        RowIdx idx = load_row_idx<16>(p.dst, p.src, (int64_t)blockIdx.x * FTE, row0, p.n_edges, (int64_t)blockIdx.x < n_tiles);
        RowIdx idx_n = load_row_idx<16>(p.dst, p.src, (blockIdx.x + G) * FTE, row0, p.n_edges, blockIdx.x + G < n_tiles);
        float inv = (idx.d >= 0) ? __ldg(p.inv_deg + idx.d) : 0.f;
        gather_half_async(ring, prow, p.PQ, idx, 0, lane);
        gather_half_async(ring + 8 * 512, prow + 1024, p.PQ, idx, 8, lane);
        int i = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += G, ++i) {
            const int b = i & 1;
            if (w == 0 || w == F_BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i, 0);
            const RowIdx idx_nn = load_row_idx<16>(p.dst, p.src, (t + 2 * G) * FTE, row0, p.n_edges, t + 2 * G < n_tiles);
            const float inv_n = (idx_n.d >= 0) ? __ldg(p.inv_deg + idx_n.d) : 0.f;
            mbar_wait(h_empty + 8 * b, ((uint32_t)(i >> 1) & 1u) ^ 1u);    // MMA of tile i-2 has consumed this stage
            if (w == 0 || w == F_BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i, 1);
            const uint32_t img = sbase + FwdSmem::H + b * (2 * F_IMG);
            if (lane < 16) {
                const uint32_t slot = sbase + FwdSmem::DST + (uint32_t)((i & (FwdSmem::SLOTS - 1)) * FTE + row0 + lane) * 4;
                sts_b32(slot, (uint32_t)idx.d);
                sts_b32(slot + FwdSmem::SLOTS * FTE * 4, __float_as_uint(inv));
            }
            cp_async_wait<1>();                                            // rows 0..7 have landed (groups complete in order)
            if (w == 0 || w == F_BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i, 3);
            build8<FTE>(img, row0, ring, prow, idx, 0, p.PQ, lane);
            if (w == 0 || w == F_BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i, 4);
            gather_half_async(ring, prow, p.PQ, idx_n, 0, lane);           // the same rows of the NEXT tile (all -1 past the end)
            cp_async_wait<1>();                                            // rows 8..15
            if (w == 0 || w == F_BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i, 5);
            build8<FTE>(img, row0 + 8, ring + 8 * 512, prow + 1024, idx, 8, p.PQ, lane);
            gather_half_async(ring + 8 * 512, prow + 1024, p.PQ, idx_n, 8, lane);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(h_full + 8 * b);
            if (w == 0 || w == F_BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i, 2);
            idx = idx_n; idx_n = idx_nn; inv = inv_n;
        }
        cp_async_wait<0>();
    } else {
        // ------------------------------------------------------------------ MMA issuer (one thread of warp 24)
        reg_dec<F_MMA_REGS>();
        constexpr uint32_t idesc = idesc_bf16(128, FTE, 0, 0);
        int i = 0;
        if (warp == F_MMA_WARP && lane == 0)
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int b = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            mbar_wait(h_full + 8 * b, ph);
            TL(2, i, 0);
            mbar_wait(tm_empty + 8 * b, ph);
            TL(2, i, 1);
            tc_fence_after();
            const uint32_t b_lo = desc_lo_sw128(sbase + FwdSmem::H + b * (2 * F_IMG), 16);
            constexpr uint32_t b_hi = desc_hi_sw128(1024);
            const uint32_t d_tm = tmem_d + b * FTE;
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {                         // hi*hi + hi*lo + lo*hi, on top of -b2
                const uint32_t a = (prod == 2) ? tmem_w_lo : tmem_w_hi;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16_ts_lh(d_tm, a + ks * 8,
                                    b_lo + (((prod == 1 ? F_IMG : 0) + (ks >> 2) * (FTE * 128) + (ks & 3) * 32) >> 4), b_hi, idesc, 1u);
            }
            umma_commit(h_empty + 8 * b);
            umma_commit(tm_full + 8 * b);
            TL(2, i, 2);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ================================================================================================================
// Backward (h1 recomputed, z2 mask read back).  Tile = 64 edges.
//   G[e][o]  = g_agg[dst][o] * inv_deg[dst] * [z2 > 0]
//   MMA-A    : D1[c][e] = sum_o W2[o][c] G[e][o]      (A = W2^T in TMEM, B = G tile K-major, N = 64)
//   MMA-B    : D2[o][c] += sum_e G[e][o] h1[e][c]     (A = G, B = h1 tiles read MN-major; D2 stays in TMEM for the
//                                                      whole kernel = this CTA's dW2 partial)
//   epilogue : D1 -> fp32 staging tile [e][c] in shared memory (transpose through TMEM lanes)
//   row phase: the builder warp that built rows r.. re-reads them: g_z1 = D1 * [h1 > 0]; dQ'[src] += g_z1 row
//              (128-bit vector reductions), dP'[dst] += per-target sums of consecutive rows.
// ================================================================================================================
constexpr int BTE = 64;
constexpr uint32_t B_IMG = 2 * BTE * 128;                  // one [64][128] bf16 image = 16 KB
struct EdgeBwdArgs {
    const float* PQ; const int* src; const int* dst; const float* inv_deg; int64_t n_edges;
    const float* w2; const uint32_t* mask2; const float* g_agg; int64_t ld_gagg;
    float* dPQ; float* dW2; float* db2;
};
struct BwdSmem {
    static constexpr uint32_t HG = 0;                      // 2 stages x (H hi, H lo, G hi, G lo) = 2 x 64 KB
    static constexpr uint32_t ST = 2 * 4 * B_IMG;          // 2 x fp32 [64][128] staging = 2 x 32 KB
    static constexpr uint32_t RING = ST + 2 * BTE * 128 * 4;   // Q' rows of one tile: 8 builder warps x 8 rows x 512 B (cp.async)
    static constexpr uint32_t BAR = RING + BTE * 512;
    static constexpr uint32_t TOTAL = BAR + 128;
};

__global__ void __launch_bounds__(EDGE_THREADS, 1) edge_bwd_tc_kernel(EdgeBwdArgs p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(sm);
    const uint32_t bar0 = sbase + BwdSmem::BAR;
    // [b] at +8*b
    const uint32_t hg_full = bar0, hg_empty = bar0 + 16, d1_full = bar0 + 32, d1_empty = bar0 + 48, st_full = bar0 + 64,
                   st_empty = bar0 + 80, all_done = bar0 + 96;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + BwdSmem::BAR + 104);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    if (tid == 32) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(hg_full + 8 * b, BLD_WARPS); mbar_init(hg_empty + 8 * b, 1);
            mbar_init(d1_full + 8 * b, 1); mbar_init(d1_empty + 8 * b, EPI_WARPS);
            mbar_init(st_full + 8 * b, EPI_WARPS); mbar_init(st_empty + 8 * b, BLD_WARPS);
        }
        mbar_init(all_done, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + 128, tmem_w_hi = tmem_base + 256, tmem_w_lo = tmem_base + 320;
    const int64_t n_tiles = (p.n_edges + BTE - 1) / BTE;

    if (warp < EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue: thread = channel c
        reg_dec<EPI_REGS>();
        const int c = warp * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        weight_to_tmem(p.w2, 1, 128, c, tmem_w_hi + lane_addr, tmem_w_lo + lane_addr);     // W2^T: row c, k = o
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(d1_empty); mbar_arrive(d1_empty + 8); }               // D1 stages usable, W2^T in place
        int i = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int b = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            const uint32_t stage = sbase + BwdSmem::ST + b * (BTE * 128 * 4) + c * 4;
            mbar_wait(d1_full + 8 * b, ph);
            if (warp == 0) TL(3, i, 0);
            tc_fence_after();
            mbar_wait(st_empty + 8 * b, ph ^ 1u);                          // row phase of tile i-2 has drained the stage
            if (warp == 0) TL(3, i, 1);
            uint32_t v[2][32];
            tmem_ld32_async(tmem_d1 + lane_addr + b * BTE, v[0]);
            tmem_ld32_async(tmem_d1 + lane_addr + b * BTE + 32, v[1]);
            tmem_wait_ld(v[0]);
            tmem_wait_ld(v[1]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d1_empty + 8 * b);                  // accumulator stage free for tile i+2
            if (warp == 0) TL(3, i, 2);
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 32; ++j) sts_b32(stage + (h * 32 + j) * 512, v[h][j]);
            __syncwarp();
            if (lane == 0) mbar_arrive(st_full + 8 * b);
            if (warp == 0) TL(3, i, 3);
        }
        // ---- this CTA's dW2 partial: D2[o][c], lane = o, registers = 32 consecutive c = one contiguous piece of
        //      row o of dW2 -> 128-bit vector reductions straight from the registers
        if (i > 0) {
            mbar_wait(all_done, 0);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {                              // rotated per CTA: spreads the same-address atomics
                const int chunk = (cc + (int)blockIdx.x) & 3;
                uint32_t v[32];
                tmem_ld32_async(tmem_d2 + lane_addr + chunk * 32, v);
                tmem_wait_ld(v);
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    red_add_v4(p.dW2 + c * 128 + chunk * 32 + 4 * m,             // the h1 tile holds 2*h1: halve (exact)
                               make_float4(0.5f * __uint_as_float(v[4 * m]), 0.5f * __uint_as_float(v[4 * m + 1]),
                                           0.5f * __uint_as_float(v[4 * m + 2]), 0.5f * __uint_as_float(v[4 * m + 3])));
            }
        }
    } else if (warp < MMA_WARP) {
        // ------------------------------------------------------------------ builders + row phase: warp w -> rows 8w .. 8w+7
        reg_inc<BLD_REGS>();
        const int w = warp - EPI_WARPS;
        const int row0 = w * 8;
        float db2_acc[4] = {0.f, 0.f, 0.f, 0.f};
        // gathered operands of one tile's 8 rows; ga/gb = g_agg[dst] of the first / last row with sa/sb = inv_deg,
        // mw = z2 sign words of the rows' 32-edge chunk for channels 4*lane..4*lane+3.  Everything is kept raw:
        // nothing may depend on a load inside the prefetch, or the prefetch turns into a stall.
        // (the 8 Q'[src] rows themselves come through the cp.async ring, not through registers)
        struct TileG { float4 pa, pb; int da, db; uint32_t odd; };
        struct Tile { TileG g; float4 ga, gb; float sa, sb; uint4 mw; };
        Tile cur, nxt;
        int i = 0;
        const uint32_t ring = sbase + BwdSmem::RING + (uint32_t)w * (8 * 512);
        auto gather_tile = [&](Tile& T, RowIdx idx, int64_t t) {
            T.g.da = __shfl_sync(0xffffffffu, idx.d, 0);
            T.g.db = __shfl_sync(0xffffffffu, idx.d, 7);
            T.g.odd = __ballot_sync(0xffffffffu, idx.d >= 0 && idx.d != T.g.da && idx.d != T.g.db) & 0xFFu;
            T.g.pa = T.g.pb = T.ga = T.gb = make_float4(0.f, 0.f, 0.f, 0.f);
            T.sa = T.sb = 0.f;
            T.mw = make_uint4(0u, 0u, 0u, 0u);
            if (T.g.da >= 0) {
                T.g.pa = ldg4(p.PQ + (int64_t)T.g.da * 256 + lane * 4);
                T.sa = __ldg(p.inv_deg + T.g.da);
                T.ga = ldg4(p.g_agg + (int64_t)T.g.da * p.ld_gagg + lane * 4);
            }
            if (T.g.db >= 0) {
                T.g.pb = ldg4(p.PQ + (int64_t)T.g.db * 256 + lane * 4);
                T.sb = __ldg(p.inv_deg + T.g.db);
                T.gb = ldg4(p.g_agg + (int64_t)T.g.db * p.ld_gagg + lane * 4);
            }
            if (t < n_tiles) T.mw = __ldg(reinterpret_cast<const uint4*>(p.mask2 + ((t * BTE + row0) >> 5) * 128 + lane * 4));
        };
        // row phase of a finished tile: this warp's rows of D1 from the staging tile, g_z1 = D1 * [h1 > 0] with the
        // sign taken from the bf16 hi image of h1 this warp wrote itself (still intact: stage b is rebuilt only by
        // this warp, one step later).  dQ'[src] += row; dP'[dst] += sum of the rows of each target (first / last row's
        // target in straight-line code, rows of a third target patched one by one).
        auto masked_row = [&](uint32_t stage, uint32_t imgH, int k) {
            float4 g = lds_v4f(stage + k * 512);
            const uint2 hh = lds_v2(imgH + tile_off<BTE>(row0 + k, lane * 4));
            g.x = (hh.x & 0xFFFFu) ? g.x : 0.f; g.y = (hh.x >> 16) ? g.y : 0.f;
            g.z = (hh.y & 0xFFFFu) ? g.z : 0.f; g.w = (hh.y >> 16) ? g.w : 0.f;
            return g;
        };
        auto row_phase = [&](int b, uint32_t ph, RowIdx idx, int da, int db, uint32_t odd) {
            const uint32_t stage = sbase + BwdSmem::ST + b * (BTE * 128 * 4) + row0 * 512 + lane * 16;
            const uint32_t imgH = sbase + BwdSmem::HG + b * (4 * B_IMG);
            mbar_wait(st_full + 8 * b, ph);
            if (w == 0 || w == BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i - 1, 3);
            float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int d = __shfl_sync(0xffffffffu, idx.d, k);
                const int sr = __shfl_sync(0xffffffffu, idx.s, k);
                const float4 g = masked_row(stage, imgH, k);               // rows beyond the last edge are exactly zero
#ifndef MMPDE_EXPERIMENT_NO_DQ_RED
                if (sr >= 0) red_add_v4(p.dPQ + (int64_t)sr * 256 + 128 + lane * 4, g);    // dQ'[src]
#else
                if (sr == -12345) red_add_v4(p.dPQ + (int64_t)sr * 256 + 128 + lane * 4, g);
#endif
                const float fa = (d == da) ? 1.f : 0.f, fb = (d == db && d != da) ? 1.f : 0.f;
                acc_a.x = fmaf(fa, g.x, acc_a.x); acc_a.y = fmaf(fa, g.y, acc_a.y);
                acc_a.z = fmaf(fa, g.z, acc_a.z); acc_a.w = fmaf(fa, g.w, acc_a.w);
                acc_b.x = fmaf(fb, g.x, acc_b.x); acc_b.y = fmaf(fb, g.y, acc_b.y);
                acc_b.z = fmaf(fb, g.z, acc_b.z); acc_b.w = fmaf(fb, g.w, acc_b.w);
            }
            if (da >= 0) red_add_v4(p.dPQ + (int64_t)da * 256 + lane * 4, acc_a);          // dP'[dst]
            if (db >= 0 && db != da) red_add_v4(p.dPQ + (int64_t)db * 256 + lane * 4, acc_b);
            for (; odd != 0u; odd &= odd - 1u) {
                const int k = __ffs(odd) - 1;
                const int d = __shfl_sync(0xffffffffu, idx.d, k);
                red_add_v4(p.dPQ + (int64_t)d * 256 + lane * 4, masked_row(stage, imgH, k));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(st_empty + 8 * b);
            if (w == 0 || w == BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i - 1, 4);
        };

        const int64_t G = gridDim.x;
        RowIdx idx = load_row_idx<8>(p.dst, p.src, (int64_t)blockIdx.x * BTE, row0, p.n_edges, (int64_t)blockIdx.x < n_tiles);
        RowIdx idx_n = load_row_idx<8>(p.dst, p.src, (blockIdx.x + G) * BTE, row0, p.n_edges, blockIdx.x + G < n_tiles);
        RowIdx idx_prev; idx_prev.d = idx_prev.s = -1;
        int da_prev = -1, db_prev = -1;
        uint32_t odd_prev = 0u;
        // per tile: `cur` was gathered one step ago; gather `nxt` (tile t + G) now, build tile t, then finish tile t - G
        gather_q8_async(ring, p.PQ, idx.s, lane);
        gather_tile(cur, idx, blockIdx.x);
        for (int64_t t = blockIdx.x; t < n_tiles; t += G) {
            const int b = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            if (w == 0 || w == BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i, 0);
            const RowIdx idx_nn = load_row_idx<8>(p.dst, p.src, (t + 2 * G) * BTE, row0, p.n_edges, t + 2 * G < n_tiles);
            gather_tile(nxt, idx_n, t + G);
            // G rows are the per-target vector g_agg[dst]/deg gated by the z2 sign bits: split it to bf16 hi/lo once per
            // target (first / last row's target), per row only select
            const float4 gsa = make_float4(cur.ga.x * cur.sa, cur.ga.y * cur.sa, cur.ga.z * cur.sa, cur.ga.w * cur.sa);
            const float4 gsb = make_float4(cur.gb.x * cur.sb, cur.gb.y * cur.sb, cur.gb.z * cur.sb, cur.gb.w * cur.sb);
            uint2 ahi, alo, bhi, blo;
            split4(gsa, ahi, alo);
            split4(gsb, bhi, blo);
            const int sh = (int)((t * BTE + row0) & 31);
            const uint32_t mx = cur.mw.x >> sh, my = cur.mw.y >> sh, mz = cur.mw.z >> sh, mw = cur.mw.w >> sh;  // bit k = row k
            // db2[o] = sum_e G[e][o] = sum over targets of (g/deg)[o] * #(rows of that target with z2 > 0)
            const uint32_t rows_a = __ballot_sync(0xffffffffu, idx.d == cur.g.da && idx.d >= 0) & 0xFFu;
            const uint32_t rows_b = __ballot_sync(0xffffffffu, idx.d == cur.g.db && idx.d >= 0) & 0xFFu & ~rows_a;
            db2_acc[0] += gsa.x * (float)__popc(mx & rows_a) + gsb.x * (float)__popc(mx & rows_b);
            db2_acc[1] += gsa.y * (float)__popc(my & rows_a) + gsb.y * (float)__popc(my & rows_b);
            db2_acc[2] += gsa.z * (float)__popc(mz & rows_a) + gsb.z * (float)__popc(mz & rows_b);
            db2_acc[3] += gsa.w * (float)__popc(mw & rows_a) + gsb.w * (float)__popc(mw & rows_b);
            mbar_wait(hg_empty + 8 * b, ph ^ 1u);
            if (w == 0 || w == BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i, 1);
            const uint32_t imgH = sbase + BwdSmem::HG + b * (4 * B_IMG);
            const uint32_t imgG = imgH + 2 * B_IMG;
            cp_async_wait<0>();                                            // this tile's Q' rows have landed
            build8_core<BTE>(imgH, row0, ring, cur.g.pa, cur.g.pb, cur.g.da, cur.g.odd, idx, 0, p.PQ, lane);
            gather_q8_async(ring, p.PQ, idx_n.s, lane);                    // the ring is free again: rows of the NEXT tile
            auto store_g = [&](int k, uint2 ghi, uint2 glo) {              // keep the halves whose z2 sign bit is set
                const uint32_t k01 = (((mx >> k) & 1u) ? 0x0000FFFFu : 0u) | (((my >> k) & 1u) ? 0xFFFF0000u : 0u);
                const uint32_t k23 = (((mz >> k) & 1u) ? 0x0000FFFFu : 0u) | (((mw >> k) & 1u) ? 0xFFFF0000u : 0u);
                const uint32_t off = tile_off<BTE>(row0 + k, lane * 4);
                sts_v2(imgG + off, make_uint2(ghi.x & k01, ghi.y & k23));
                sts_v2(imgG + B_IMG + off, make_uint2(glo.x & k01, glo.y & k23));
            };
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool is_a = (__shfl_sync(0xffffffffu, idx.d, k) == cur.g.da);        // rows past the end: gsb = 0
                store_g(k, make_uint2(is_a ? ahi.x : bhi.x, is_a ? ahi.y : bhi.y), make_uint2(is_a ? alo.x : blo.x, is_a ? alo.y : blo.y));
            }
            for (uint32_t odd = cur.g.odd; odd != 0u; odd &= odd - 1u) {   // rows of a third target (in-degree < 4): rare
                const int k = __ffs(odd) - 1;
                const int d = __shfl_sync(0xffffffffu, idx.d, k);
                const float sc = __ldg(p.inv_deg + d);
                const float4 x = ldg4(p.g_agg + (int64_t)d * p.ld_gagg + lane * 4);
                const float4 gs = make_float4(x.x * sc, x.y * sc, x.z * sc, x.w * sc);
                uint2 ghi, glo;
                split4(gs, ghi, glo);
                store_g(k, ghi, glo);
                db2_acc[0] += ((mx >> k) & 1u) ? gs.x : 0.f; db2_acc[1] += ((my >> k) & 1u) ? gs.y : 0.f;
                db2_acc[2] += ((mz >> k) & 1u) ? gs.z : 0.f; db2_acc[3] += ((mw >> k) & 1u) ? gs.w : 0.f;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(hg_full + 8 * b);
            if (w == 0 || w == BLD_WARPS - 1) TL(w == 0 ? 0 : 1, i, 2);
            if (i > 0) row_phase(b ^ 1, (uint32_t)((i - 1) >> 1) & 1u, idx_prev, da_prev, db_prev, odd_prev);
            idx_prev = idx; da_prev = cur.g.da; db_prev = cur.g.db; odd_prev = cur.g.odd;
            idx = idx_n; idx_n = idx_nn;
            cur = nxt;                                                     // loads issued a whole step ago: no stall
            ++i;
        }
        cp_async_wait<0>();
        if (i > 0) row_phase((i - 1) & 1, (uint32_t)((i - 1) >> 1) & 1u, idx_prev, da_prev, db_prev, odd_prev);
#pragma unroll
        for (int f = 0; f < 4; ++f) atomicAdd(p.db2 + lane * 4 + f, db2_acc[f]);
    } else {
        // ------------------------------------------------------------------ MMA issuer (one thread of warp 12)
        reg_dec<MMA_REGS>();
        constexpr uint32_t idesc_a = idesc_bf16(128, BTE, 0, 0);           // D1: A (TMEM) K-major, B = G K-major
        constexpr uint32_t idesc_b = idesc_bf16(128, 128, 1, 1);           // D2: A = G, B = h1, both MN-major
        int i = 0;
        if (warp == MMA_WARP && lane == 0)
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int b = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            mbar_wait(hg_full + 8 * b, ph);
            TL(2, i, 0);
            mbar_wait(d1_empty + 8 * b, ph);
            TL(2, i, 1);
            tc_fence_after();
            const uint32_t h_addr = sbase + BwdSmem::HG + b * (4 * B_IMG), g_addr = h_addr + 2 * B_IMG;
            constexpr uint32_t d_hi = desc_hi_sw128(1024);
            const uint32_t gk_lo = desc_lo_sw128(g_addr, 16);              // G read K-major (MMA-A)
            const uint32_t gm_lo = desc_lo_sw128(g_addr, BTE * 128), hm_lo = desc_lo_sw128(h_addr, BTE * 128);   // MN-major (MMA-B)
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {                         // W2^T hi * G hi + hi * G lo + lo * G hi
                const uint32_t a = (prod == 2) ? tmem_w_lo : tmem_w_hi;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)                             // 16 values of o per step
                    umma_bf16_ts_lh(tmem_d1 + b * BTE, a + ks * 8,
                                    gk_lo + (((prod == 1 ? B_IMG : 0) + (ks >> 2) * (BTE * 128) + (ks & 3) * 32) >> 4), d_hi, idesc_a,
                                    (prod | ks) ? 1u : 0u);
            }
            umma_commit(d1_full + 8 * b);
            TL(2, i, 2);
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {                         // G hi * h hi + G hi * h lo + G lo * h hi
#pragma unroll
                for (int ks = 0; ks < BTE / 16; ++ks)                      // 16 edges per step
                    umma_bf16_lh(tmem_d2, gm_lo + (((prod == 2 ? B_IMG : 0) + ks * 16 * 128) >> 4), d_hi,
                                 hm_lo + (((prod == 1 ? B_IMG : 0) + ks * 16 * 128) >> 4), d_hi, idesc_b,
                                 (i > 0 || prod > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(hg_empty + 8 * b);
            TL(2, i, 3);
        }
        if (warp == MMA_WARP && lane == 0 && i > 0) umma_commit(all_done);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace mmpde

using namespace mmpde;

#ifdef MMPDE_TIMELINE
extern "C" int mmpde_debug_timeline(long long* buf) {
    return (int)cudaMemcpyToSymbol(g_timeline, &buf, sizeof(buf));
}
#endif

static int edge_grid(int64_t n_tiles) { return (int)imin64(n_tiles, persistent_ctas()); }

extern "C" int mmpde_edge_fwd(const float* PQ, const int32_t* edge_src, const int32_t* edge_dst, const float* inv_deg,
                              int64_t n_edges, const float* w2, const float* b2, float* agg, int64_t ld_agg,
                              uint32_t* mask2, void* stream) {
    if (n_edges < 0 || ld_agg < 128) return MMPDE_EINVAL;
    if (n_edges == 0) return MMPDE_OK;
    if (reinterpret_cast<uintptr_t>(PQ) & 15) return MMPDE_EINVAL;
    constexpr size_t smem = FwdSmem::TOTAL + 1024;
    MMPDE_ENSURE_SMEM(edge_fwd_tc_kernel, smem);
    EdgeFwdArgs p;
    p.PQ = PQ; p.src = edge_src; p.dst = edge_dst; p.inv_deg = inv_deg; p.n_edges = n_edges; p.w2 = w2; p.b2 = b2;
    p.agg = agg; p.ld_agg = ld_agg; p.mask2 = mask2;
    edge_fwd_tc_kernel<<<edge_grid((n_edges + FTE - 1) / FTE), F_THREADS, smem, (cudaStream_t)stream>>>(p);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_edge_bwd(const float* PQ, const int32_t* edge_src, const int32_t* edge_dst, const float* inv_deg,
                              int64_t n_edges, const float* w2, const uint32_t* mask2, const float* g_agg,
                              int64_t ld_gagg, float* dPQ, float* dW2, float* db2, void* stream) {
    if (n_edges < 0 || ld_gagg < 128) return MMPDE_EINVAL;
    if (n_edges == 0) return MMPDE_OK;
    EdgeBwdArgs p;
    p.PQ = PQ; p.src = edge_src; p.dst = edge_dst; p.inv_deg = inv_deg; p.n_edges = n_edges; p.w2 = w2; p.mask2 = mask2;
    p.g_agg = g_agg; p.ld_gagg = ld_gagg; p.dPQ = dPQ; p.dW2 = dW2; p.db2 = db2;
    constexpr size_t smem = BwdSmem::TOTAL + 1024;
    MMPDE_ENSURE_SMEM(edge_bwd_tc_kernel, smem);
    edge_bwd_tc_kernel<<<edge_grid((n_edges + BTE - 1) / BTE), EDGE_THREADS, smem, (cudaStream_t)stream>>>(p);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
