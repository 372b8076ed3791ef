// Message passing over the target-sorted edge list on the 5th-gen tensor cores (tcgen05 + TMEM).
// Replaces PyG propagate + message_net_1/2 + scatter-mean (/root/reference/gnn_2d.py:55,59-63).
//
// Per tile of 128 consecutive edges (rows e), with z1 = P[dst] + Q[src] + W1c.e_ij:
//   build   : all 8 warps gather P/Q rows (128-bit coalesced loads), add the 4 scalar edge features, ReLU and
//             write h1[e][c] as split-bf16 (hi, lo) into a SWIZZLE_128B operand tile in shared memory
//             -> the gather is fused into the operand load of the GEMM, no [E,260] / [E,128] tensor exists;
//   MMA     : D[o][e] = sum_c W2[o][c] * h1[e][c]   (M = 128 out-channels, N = 128 edges, K = 128), three
//             bf16 products (hi*hi + hi*lo + lo*hi) accumulated in fp32 in TMEM, issued by one thread;
//             W2 lives in shared memory for the whole kernel (one TMA bulk copy of a pre-swizzled image);
//   epilogue: thread = out-channel (TMEM lane), registers = 64 consecutive edges: +b2, ReLU, z2>0 bit mask via
//             ballot, running sum along the edge axis, flushed as mean at every change of target (coalesced
//             128-byte reductions; a target's segment may continue in the neighbouring tile / column half).
// Double-buffered operand tiles and TMEM accumulators: MMA(t) overlaps build(t+1) and epilogue(t-1).
#include "tc_common.cuh"

namespace mmpde {
using namespace tc;

constexpr int TE = 128;                        // edges per tile
constexpr uint32_t TILE_BYTES = 2 * KBLK_BYTES; // one [128][128] bf16 operand image (2 K-blocks)
constexpr uint32_t W2_IMG_BYTES = 2 * TILE_BYTES;   // hi + lo

// ---------------------------------------------------------------------------------------------------
// weight image: W [128][128] fp32 row-major -> (hi, lo) bf16 SWIZZLE_128B tiles, 65 536 bytes
__global__ void pack_w128_kernel(const float* __restrict__ w, unsigned char* __restrict__ img) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 4 consecutive columns
    if (idx >= 128 * 32) return;
    int row = idx >> 5, col = (idx & 31) * 4;
    uint2 hi, lo;
    split4(ldg4(w + row * 128 + col), hi, lo);
    uint32_t off = tile_off(row, col);
    *reinterpret_cast<uint2*>(img + off) = hi;
    *reinterpret_cast<uint2*>(img + TILE_BYTES + off) = lo;
}

struct EdgeTcArgs {
    const float* PQ; const float4* node4; const int* src; const int* dst; const float* inv_deg;
    int64_t n_edges; const float* w1c; const unsigned char* w2_img; const float* b2;
    float* agg; int64_t ld_agg; uint32_t* mask2;
};

struct FwdSmem {
    // offsets from the 1024-aligned base
    static constexpr uint32_t W2 = 0;                               // hi, lo
    static constexpr uint32_t H0 = W2_IMG_BYTES;                    // buffer 0: hi, lo ; buffer 1 follows
    static constexpr uint32_t DST = H0 + 2 * W2_IMG_BYTES;          // int dst[3][128] (slot = tile iteration % 3)
    static constexpr uint32_t BAR = DST + 3 * TE * 4;               // 3 mbarriers + tmem slot
    static constexpr uint32_t TOTAL = BAR + 64;
};

// Edge indices of the 16 rows a warp builds: lane r (< 16) holds row 16*warp + r.  Loaded one tile AHEAD so the
// index -> row-gather dependency never sits on the critical path.
struct RowIdx { int d, s; };
__device__ __forceinline__ RowIdx load_row_idx(const int* __restrict__ dst, const int* __restrict__ src, int64_t e0,
                                               int64_t n_edges) {
    RowIdx r; r.d = -1; r.s = -1;
    const int lane = threadIdx.x & 31;
    const int64_t e = e0 + (threadIdx.x >> 5) * 16 + (lane & 15);
    if (lane < 16 && e0 >= 0 && e < n_edges) { r.d = __ldg(dst + e); r.s = __ldg(src + e); }
    return r;
}

__device__ __forceinline__ float4 h1_row(const float4& P, const float4& Q, const float4& ni, const float4& nj,
                                         const float (&w1c)[4][4], float4& e) {
    e = make_float4(ni.x - nj.x, ni.y - nj.y, ni.z - nj.z, ni.w);     // (u_i-u_j, px_i-px_j, py_i-py_j, v_i)
    float z[4] = {P.x + Q.x, P.y + Q.y, P.z + Q.z, P.w + Q.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        z[c] = fmaf(w1c[c][0], e.x, z[c]);
        z[c] = fmaf(w1c[c][1], e.y, z[c]);
        z[c] = fmaf(w1c[c][2], e.z, z[c]);
        z[c] = fmaf(w1c[c][3], e.w, z[c]);
    }
    return make_float4(fmaxf(z[0], 0.f), fmaxf(z[1], 0.f), fmaxf(z[2], 0.f), fmaxf(z[3], 0.f));
}

// h1 rows of one tile -> split-bf16 operand tile `hbase` (hi at +0, lo at +TILE_BYTES); warp w builds rows 16w..16w+15,
// 8 rows (16 independent 512-byte row gathers) in flight at a time.
__device__ __forceinline__ void build_h1_tile(const EdgeTcArgs& p, RowIdx idx, unsigned char* hbase, int* sDst,
                                              const float (&w1c)[4][4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 16) sDst[warp * 16 + lane] = idx.d;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        float4 P[8], Q[8], ni[8], nj[8];
        int di[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            di[k] = __shfl_sync(0xffffffffu, idx.d, b * 8 + k);
            int j = __shfl_sync(0xffffffffu, idx.s, b * 8 + k);
            if (di[k] >= 0) {
                P[k] = ldg4(p.PQ + (int64_t)di[k] * 256 + lane * 4);
                Q[k] = ldg4(p.PQ + (int64_t)j * 256 + 128 + lane * 4);
                ni[k] = __ldg(p.node4 + di[k]);
                nj[k] = __ldg(p.node4 + j);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f), e;
            if (di[k] >= 0) h = h1_row(P[k], Q[k], ni[k], nj[k], w1c, e);
            uint2 hi, lo;
            split4(h, hi, lo);
            uint32_t off = tile_off(warp * 16 + b * 8 + k, lane * 4);
            *reinterpret_cast<uint2*>(hbase + off) = hi;
            *reinterpret_cast<uint2*>(hbase + TILE_BYTES + off) = lo;
        }
    }
}

// 24 MMAs: (W2hi,Hhi) + (W2hi,Hlo) + (W2lo,Hhi), each 2 K-blocks x 4 K-steps of 16
__device__ __forceinline__ void issue_w2_times_h(uint32_t w2_addr, uint32_t h_addr, uint32_t tmem_d) {
    constexpr uint32_t idesc = idesc_bf16(128, TE, 0, 0);
    uint32_t acc = 0;
#pragma unroll
    for (int prod = 0; prod < 3; ++prod) {
        uint32_t a = w2_addr + (prod == 2 ? TILE_BYTES : 0);
        uint32_t b = h_addr + (prod == 1 ? TILE_BYTES : 0);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t o = kb * KBLK_BYTES + ks * 32;
                umma_bf16(tmem_d, smem_desc_sw128(a + o, 16, 1024), smem_desc_sw128(b + o, 16, 1024), idesc, acc);
                acc = 1;
            }
    }
}

__global__ void __launch_bounds__(256, 1) edge_fwd_tc_kernel(EdgeTcArgs p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(sm);
    int* sDst = reinterpret_cast<int*>(sm + FwdSmem::DST);
    const uint32_t bar_w = sbase + FwdSmem::BAR, bar_mma0 = bar_w + 8, bar_mma1 = bar_w + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + FwdSmem::BAR + 24);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
    if (tid == 32) {
        mbar_init(bar_w, 1); mbar_init(bar_mma0, 1); mbar_init(bar_mma1, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_w, W2_IMG_BYTES);
        tma_bulk_g2s(sbase + FwdSmem::W2, p.w2_img, W2_IMG_BYTES, bar_w);
    }
    float w1c[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int f = 0; f < 4; ++f) w1c[c][f] = __ldg(p.w1c + (lane * 4 + c) * 4 + f);
    const int q = warp & 3, half = warp >> 2;
    const int o = q * 32 + lane;                       // this thread's out-channel in the epilogue
    const float bias = __ldg(p.b2 + o);

    const int64_t n_tiles = (p.n_edges + TE - 1) / TE;
    int it = 0;
    int64_t prev_e0 = 0;
    int prev_rows = 0;

    auto epilogue = [&](int buf, int slot, int64_t e0, int rows, uint32_t parity) {
        mbar_wait(buf ? bar_mma1 : bar_mma0, parity);
        tc_fence_after();
        const int* dsts = sDst + slot * TE;
        int cur = -1;
        float run = 0.f;
#pragma unroll
        for (int chunk = 0; chunk < 2; ++chunk) {
            const int col0 = half * 64 + chunk * 32;
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TE + col0), v);
            uint32_t my_word = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int e = col0 + j;
                float z = __uint_as_float(v[j]) + bias;
                uint32_t ball = __ballot_sync(0xffffffffu, z > 0.f);
                if (lane == j) my_word = ball;
                int d = dsts[e];                       // -1 beyond the last edge
                if (d != cur) {
                    if (cur >= 0) atomicAdd(p.agg + (int64_t)cur * p.ld_agg + o, run * __ldg(p.inv_deg + cur));
                    cur = d; run = 0.f;
                }
                run += fmaxf(z, 0.f);
            }
            // mask2[e][q] : bit (o & 31) of (z2[e][o] > 0); lane j holds the word of edge col0 + j
            if (col0 + lane < rows) p.mask2[(e0 + col0 + lane) * 4 + q] = my_word;
        }
        if (cur >= 0) atomicAdd(p.agg + (int64_t)cur * p.ld_agg + o, run * __ldg(p.inv_deg + cur));
        tc_fence_before();
    };

    RowIdx idx = load_row_idx(p.dst, p.src, (int64_t)blockIdx.x < n_tiles ? (int64_t)blockIdx.x * TE : -1, p.n_edges);
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const int64_t e0 = t * TE;
        const int rows = (int)((p.n_edges - e0 < TE) ? (p.n_edges - e0) : TE);
        const RowIdx nxt = load_row_idx(p.dst, p.src, (t + gridDim.x < n_tiles) ? (t + gridDim.x) * TE : -1, p.n_edges);
        build_h1_tile(p, idx, sm + FwdSmem::H0 + buf * W2_IMG_BYTES, sDst + (it % 3) * TE, w1c);
        idx = nxt;
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            if (it == 0) mbar_wait(bar_w, 0);
            tc_fence_after();
            issue_w2_times_h(sbase + FwdSmem::W2, sbase + FwdSmem::H0 + buf * W2_IMG_BYTES, tmem_base + buf * TE);
            umma_commit(buf ? bar_mma1 : bar_mma0);
        }
        if (it > 0) epilogue(buf ^ 1, (it - 1) % 3, prev_e0, prev_rows, (uint32_t)(((it - 1) >> 1) & 1));
        prev_e0 = e0; prev_rows = rows;
    }
    if (it > 0) epilogue((it - 1) & 1, (it - 1) % 3, prev_e0, prev_rows, (uint32_t)(((it - 1) >> 1) & 1));
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}


// ---------------------------------------------------------------------------------------------------
// Backward.  Per tile (h1 recomputed, z2 mask read back):
//   G[e][o]  = g_agg[dst][o] * inv_deg[dst] * [z2 > 0]                     (split-bf16 operand tile)
//   MMA-A    : D1[c][e] = sum_o W2[o][c] * G[e][o]        (A = W2 image read MN-major, B = G K-major)
//   MMA-B    : D2[o][c] += sum_e G[e][o] * h1[e][c]       (A = G, B = h1 tiles read MN-major; D2 stays in
//                                                          TMEM for the whole kernel = this CTA's dW2 partial)
//   epilogue : thread = channel c: g_z1 = D1 * [h1 > 0]; running sums along e -> dP[dst]; dW1c partials;
//              g_z1 staged as fp32 rows, then one 128-bit vector reduction per 4 channels -> dQ[src], and g_u.
struct EdgeBwdTcArgs {
    const float* PQ; const float4* node4; const int* src; const int* dst; const float* inv_deg;
    int64_t n_edges; const float* w1c; const unsigned char* w2_img; const uint32_t* mask2;
    const float* g_agg; int64_t ld_gagg;
    float* dPQ; float* dW2; float* db2; float* dW1c; float* g_u; int64_t g_u_stride;
};

struct BwdSmem {
    static constexpr uint32_t W2 = 0;
    static constexpr uint32_t H = W2_IMG_BYTES;
    static constexpr uint32_t G = 2 * W2_IMG_BYTES;                 // later: fp32 staging [128][128]
    static constexpr uint32_t E4 = 3 * W2_IMG_BYTES;                // float4 e_ij[128]
    static constexpr uint32_t DST = E4 + TE * 16;
    static constexpr uint32_t SRC = DST + TE * 4;
    static constexpr uint32_t BAR = SRC + TE * 4;
    static constexpr uint32_t TOTAL = BAR + 64;
};

__global__ void __launch_bounds__(256, 1) edge_bwd_tc_kernel(EdgeBwdTcArgs p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(sm);
    unsigned char* sH = sm + BwdSmem::H;
    unsigned char* sG = sm + BwdSmem::G;
    float* stage = reinterpret_cast<float*>(sG);
    float4* sE = reinterpret_cast<float4*>(sm + BwdSmem::E4);
    int* sDst = reinterpret_cast<int*>(sm + BwdSmem::DST);
    int* sSrc = reinterpret_cast<int*>(sm + BwdSmem::SRC);
    const uint32_t bar_w = sbase + BwdSmem::BAR, bar_mma = bar_w + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + BwdSmem::BAR + 24);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
    if (tid == 32) {
        mbar_init(bar_w, 1); mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d1 = *tmem_slot, tmem_d2 = *tmem_slot + 128;
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_w, W2_IMG_BYTES);
        tma_bulk_g2s(sbase + BwdSmem::W2, p.w2_img, W2_IMG_BYTES, bar_w);
    }
    float w1c[4][4];                                   // rows 4*lane .. 4*lane+3 (build / scatter mapping)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int f = 0; f < 4; ++f) w1c[c][f] = __ldg(p.w1c + (lane * 4 + c) * 4 + f);
    const int q = warp & 3, half = warp >> 2;
    const int ch = q * 32 + lane;                      // epilogue channel of this thread
    float db2_acc[4] = {0.f, 0.f, 0.f, 0.f};           // channels 4*lane.. of the rows this warp builds
    float dw1c_acc[4] = {0.f, 0.f, 0.f, 0.f};          // channel ch, e-half `half`

    const int64_t n_tiles = (p.n_edges + TE - 1) / TE;
    int it = 0;
    RowIdx idx = load_row_idx(p.dst, p.src, (int64_t)blockIdx.x < n_tiles ? (int64_t)blockIdx.x * TE : -1, p.n_edges);
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int64_t e0 = t * TE;
        const int rows = (int)((p.n_edges - e0 < TE) ? (p.n_edges - e0) : TE);
        // ---- build h1 and G operand tiles (warp w: rows 16w .. 16w+15), 4 rows of gathers in flight
        const RowIdx nxt = load_row_idx(p.dst, p.src, (t + gridDim.x < n_tiles) ? (t + gridDim.x) * TE : -1, p.n_edges);
        if (lane < 16) { sDst[warp * 16 + lane] = idx.d; sSrc[warp * 16 + lane] = idx.s; }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float4 P[4], Q[4], ni[4], nj[4], ga[4];
            float sc[4];
            uint32_t mw[4];
            int di[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = warp * 16 + b * 4 + k;
                di[k] = __shfl_sync(0xffffffffu, idx.d, b * 4 + k);
                int j = __shfl_sync(0xffffffffu, idx.s, b * 4 + k);
                if (di[k] >= 0) {
                    P[k] = ldg4(p.PQ + (int64_t)di[k] * 256 + lane * 4);
                    Q[k] = ldg4(p.PQ + (int64_t)j * 256 + 128 + lane * 4);
                    ni[k] = __ldg(p.node4 + di[k]);
                    nj[k] = __ldg(p.node4 + j);
                    ga[k] = ldg4(p.g_agg + (int64_t)di[k] * p.ld_gagg + lane * 4);
                    sc[k] = __ldg(p.inv_deg + di[k]);
                    mw[k] = __ldg(p.mask2 + (e0 + r) * 4 + (lane >> 3));
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = warp * 16 + b * 4 + k;
                float4 h = make_float4(0.f, 0.f, 0.f, 0.f), g = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                if (di[k] >= 0) {
                    h = h1_row(P[k], Q[k], ni[k], nj[k], w1c, e);
                    uint32_t bits = mw[k] >> ((lane & 7) * 4);
                    g.x = (bits & 1u) ? ga[k].x * sc[k] : 0.f;
                    g.y = (bits & 2u) ? ga[k].y * sc[k] : 0.f;
                    g.z = (bits & 4u) ? ga[k].z * sc[k] : 0.f;
                    g.w = (bits & 8u) ? ga[k].w * sc[k] : 0.f;
                    db2_acc[0] += g.x; db2_acc[1] += g.y; db2_acc[2] += g.z; db2_acc[3] += g.w;
                }
                if (lane == 0) sE[r] = e;
                uint2 hi, lo;
                uint32_t off = tile_off(r, lane * 4);
                split4(h, hi, lo);
                *reinterpret_cast<uint2*>(sH + off) = hi;
                *reinterpret_cast<uint2*>(sH + TILE_BYTES + off) = lo;
                split4(g, hi, lo);
                *reinterpret_cast<uint2*>(sG + off) = hi;
                *reinterpret_cast<uint2*>(sG + TILE_BYTES + off) = lo;
            }
        }
        idx = nxt;
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            if (it == 0) mbar_wait(bar_w, 0);
            tc_fence_after();
            const uint32_t w2a = sbase + BwdSmem::W2, ha = sbase + BwdSmem::H, ga = sbase + BwdSmem::G;
            // MMA-A: D1[c][e] = sum_o W2[o][c] G[e][o]
            {
                constexpr uint32_t idesc = idesc_bf16(128, TE, /*A MN-major*/ 1, /*B K-major*/ 0);
                uint32_t acc = 0;
#pragma unroll
                for (int prod = 0; prod < 3; ++prod) {
                    uint32_t a = w2a + (prod == 2 ? TILE_BYTES : 0);
                    uint32_t b = ga + (prod == 1 ? TILE_BYTES : 0);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {          // 16 values of o per step
                        uint64_t ad = smem_desc_sw128(a + ks * 16 * 128, KBLK_BYTES, 1024);
                        uint64_t bd = smem_desc_sw128(b + (ks >> 2) * KBLK_BYTES + (ks & 3) * 32, 16, 1024);
                        umma_bf16(tmem_d1, ad, bd, idesc, acc);
                        acc = 1;
                    }
                }
            }
            // MMA-B: D2[o][c] += sum_e G[e][o] h1[e][c]
            {
                constexpr uint32_t idesc = idesc_bf16(128, 128, 1, 1);
#pragma unroll
                for (int prod = 0; prod < 3; ++prod) {
                    uint32_t a = ga + (prod == 2 ? TILE_BYTES : 0);
                    uint32_t b = ha + (prod == 1 ? TILE_BYTES : 0);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {          // 16 edges per step
                        uint64_t ad = smem_desc_sw128(a + ks * 16 * 128, KBLK_BYTES, 1024);
                        uint64_t bd = smem_desc_sw128(b + ks * 16 * 128, KBLK_BYTES, 1024);
                        umma_bf16(tmem_d2, ad, bd, idesc, (it > 0 || prod > 0 || ks > 0) ? 1u : 0u);
                    }
                }
            }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, (uint32_t)(it & 1));
        tc_fence_after();
        // ---- epilogue: thread = channel ch, 64 edges of half `half`
        {
            int cur = -1;
            float run = 0.f;
#pragma unroll
            for (int chunk = 0; chunk < 2; ++chunk) {
                const int col0 = half * 64 + chunk * 32;
                uint32_t v[32];
                tmem_ld32(tmem_d1 + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const int e = col0 + jj;
                    uint16_t hbits = *reinterpret_cast<const uint16_t*>(sH + tile_off(e, ch));
                    float gz = ((hbits & 0x7FFFu) != 0 && (hbits & 0x8000u) == 0) ? __uint_as_float(v[jj]) : 0.f;
                    int d = sDst[e];
                    if (d != cur) {
                        if (cur >= 0) atomicAdd(p.dPQ + (int64_t)cur * 256 + ch, run);
                        cur = d; run = 0.f;
                    }
                    run += gz;
                    float4 ef = sE[e];
                    dw1c_acc[0] = fmaf(gz, ef.x, dw1c_acc[0]);
                    dw1c_acc[1] = fmaf(gz, ef.y, dw1c_acc[1]);
                    dw1c_acc[2] = fmaf(gz, ef.z, dw1c_acc[2]);
                    dw1c_acc[3] = fmaf(gz, ef.w, dw1c_acc[3]);
                    stage[e * 128 + ch] = gz;          // G tile is dead once the MMAs have completed
                }
            }
            if (cur >= 0) atomicAdd(p.dPQ + (int64_t)cur * 256 + ch, run);
        }
        tc_fence_before();
        __syncthreads();
        // ---- dQ[src] += g_z1 rows (128-bit vector reductions), g_u[dst] += g_z1.W1c[:,0], g_u[src] -= same
        for (int r = warp * 16; r < min(warp * 16 + 16, rows); ++r) {
            float4 g = *reinterpret_cast<const float4*>(stage + r * 128 + lane * 4);
            int j = sSrc[r];
            red_add_v4(p.dPQ + (int64_t)j * 256 + 128 + lane * 4, g);
            if (p.g_u) {
                float part = g.x * w1c[0][0] + g.y * w1c[1][0] + g.z * w1c[2][0] + g.w * w1c[3][0];
                part = warp_sum(part);
                if (lane == 0) {
                    atomicAdd(p.g_u + (int64_t)sDst[r] * p.g_u_stride, part);
                    atomicAdd(p.g_u + (int64_t)j * p.g_u_stride, -part);
                }
            }
        }
        __syncthreads();                               // staging / tiles free for the next build
    }
    // ---- flush per-CTA partials
    if (it > 0) {
        tc_fence_after();
#pragma unroll
        for (int chunk = 0; chunk < 2; ++chunk) {
            const int col0 = half * 64 + chunk * 32;
            uint32_t v[32];
            tmem_ld32(tmem_d2 + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) atomicAdd(p.dW2 + ch * 128 + col0 + jj, __uint_as_float(v[jj]));
        }
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            atomicAdd(p.dW1c + ch * 4 + f, dw1c_acc[f]);
            atomicAdd(p.db2 + lane * 4 + f, db2_acc[f]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d1, 256);
}

}  // namespace mmpde

using namespace mmpde;

extern "C" int mmpde_pack_w128(const float* w, void* img, void* stream) {
    pack_w128_kernel<<<16, 256, 0, (cudaStream_t)stream>>>(w, (unsigned char*)img);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_edge_fwd(const float* PQ, const float* node4, const int32_t* edge_src, const int32_t* edge_dst,
                              const float* inv_deg, int64_t n_edges, const float* w1c, const void* w2_img, const float* b2,
                              float* agg, int64_t ld_agg, uint32_t* mask2, void* stream) {
    if (n_edges < 0 || ld_agg < 128) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(w2_img) & 15) != 0) return MMPDE_EINVAL;
    if (n_edges == 0) return MMPDE_OK;
    constexpr size_t smem = FwdSmem::TOTAL + 1024;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(edge_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    EdgeTcArgs p;
    p.PQ = PQ; p.node4 = (const float4*)node4; p.src = edge_src; p.dst = edge_dst; p.inv_deg = inv_deg;
    p.n_edges = n_edges; p.w1c = w1c; p.w2_img = (const unsigned char*)w2_img; p.b2 = b2;
    p.agg = agg; p.ld_agg = ld_agg; p.mask2 = mask2;
    int64_t n_tiles = (n_edges + TE - 1) / TE;
    int grid = (int)imin64(n_tiles, sm_count());
    edge_fwd_tc_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_edge_bwd(const float* PQ, const float* node4, const int32_t* edge_src, const int32_t* edge_dst,
                              const float* inv_deg, int64_t n_edges, const float* w1c, const void* w2_img,
                              const uint32_t* mask2, const float* g_agg, int64_t ld_gagg, float* dPQ, float* dW2,
                              float* db2, float* dW1c, float* g_u, int64_t g_u_stride, void* stream) {
    if (n_edges < 0 || ld_gagg < 128) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(w2_img) & 15) != 0) return MMPDE_EINVAL;
    if (n_edges == 0) return MMPDE_OK;
    constexpr size_t smem = BwdSmem::TOTAL + 1024;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(edge_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    EdgeBwdTcArgs p;
    p.PQ = PQ; p.node4 = (const float4*)node4; p.src = edge_src; p.dst = edge_dst; p.inv_deg = inv_deg;
    p.n_edges = n_edges; p.w1c = w1c; p.w2_img = (const unsigned char*)w2_img; p.mask2 = mask2;
    p.g_agg = g_agg; p.ld_gagg = ld_gagg; p.dPQ = dPQ; p.dW2 = dW2; p.db2 = db2; p.dW1c = dW1c; p.g_u = g_u;
    p.g_u_stride = g_u_stride;
    int64_t n_tiles = (n_edges + TE - 1) / TE;
    int grid = (int)imin64(n_tiles, sm_count());
    edge_bwd_tc_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
