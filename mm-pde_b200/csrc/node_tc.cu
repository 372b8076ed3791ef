// Node-level dense contractions of the processor on the 5th-gen tensor cores (tcgen05 + TMEM): nn.Linear forward /
// dgrad (mmpde_node_gemm) and wgrad (mmpde_node_wgrad) of /root/reference/gnn_2d.py:44-49,61,67-68,99-106 and their
// autograd.  Same numerics as the edge kernels: every product runs as three bf16 MMAs (hi*hi + hi*lo + lo*hi) with fp32
// accumulation in TMEM; operands are fp32 in HBM and are split to bf16 hi/lo while they are staged into the
// SWIZZLE_128B shared-memory tiles (so no TMA: the staging IS the conversion).
//
// mmpde_node_gemm  : C[m][n] = act(sum_k A[m][k] W(n,k) + node4[m] . Wext[n] + bias[n]) + R1[m][n] + R2[m][n]
//   computed TRANSPOSED, D[n][m] = W A^T: the 128 outputs n sit on the TMEM lanes, a tile of 128 rows m on the columns, so
//   an epilogue warp stores 32 consecutive n of one row per instruction (coalesced) and W (hi, lo) is loaded ONCE per CTA
//   into tensor memory as the A operand of the TS-form MMA.  K = 128 per segment, 1 or 2 segments (K = 256) with their
//   own base pointers, plus an optional 4-column extension (the node scalars u, x, y, t) as one more K step.
// mmpde_node_wgrad : dW[i][j] += sum_m A[m][i] B[m][j]  (+ dWext[i][f] += sum_m A[m][i] node4[m][f], dbias[i] += sum_m A[m][i])
//   K = the node axis, streamed in tiles of 64 rows; both operand tiles are read MN-major (no transposition in memory);
//   the accumulator stays in TMEM for the whole CTA and is reduced into HBM once (128-bit vector reductions).
// Both kernels are persistent and warp-specialised like the edge kernels (epilogue | builders | one MMA thread).
#include "tc_common.cuh"

namespace mmpde {
using namespace tc;

constexpr uint32_t N_TMEM_COLS = 512;

// ================================================================================================================
// forward / dgrad
// ================================================================================================================
constexpr int G_EPI_WARPS = 8, G_BLD_WARPS = 8, G_MMA_WARP = 16, G_THREADS = 640;
constexpr int G_EPI_REGS = 104, G_BLD_REGS = 112, G_MMA_REGS = 40;         // 8*32*104 + 8*32*112 + 4*32*40 <= 640*96
constexpr int GT = 128;                                                    // rows per tile
constexpr uint32_t G_IMG = 2 * GT * 128;                                   // one [128][128] bf16 image = 32 KB
constexpr uint32_t G_XIMG = GT * 128;                                      // extension image [128 rows][128 B] = 16 KB

struct NodeGemmArgs {
    const float* A[2]; int64_t lda[2]; int nseg;
    const float* W[2]; int64_t w_ns[2], w_ks[2];
    const unsigned char* Wimg[2];                                          // pre-split operand images (mmpde_weight_images) or NULL
    const float* Aext; const float* Wext;
    const float* bias; int relu;
    const float* R1; int64_t ldr1; const float* R2; int64_t ldr2;
    float* C; int64_t ldc; int64_t M;
};
struct GemmSmem {
    static constexpr uint32_t SEG = 0;                                     // 2 stages x (hi, lo)
    static constexpr uint32_t EXT = 2 * 2 * G_IMG;                         // 2 x (hi, lo) extension tiles
    static constexpr uint32_t BAR = EXT + 2 * 2 * G_XIMG;
    static constexpr uint32_t TOTAL = BAR + 128;
};

__global__ void __launch_bounds__(G_THREADS, 1) node_gemm_tc_kernel(NodeGemmArgs p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(sm);
    const uint32_t bar0 = sbase + GemmSmem::BAR;
    const uint32_t h_full = bar0, h_empty = bar0 + 16, tm_full = bar0 + 32, tm_empty = bar0 + 48;   // [b] at +8*b
    const uint32_t w_full = bar0 + 72, w_cp = bar0 + 80;                   // weight image landed | copied to TMEM ([seg] at +8*seg)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + GemmSmem::BAR + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) { TL(3, 0, 6); tmem_alloc(smem_u32(tmem_slot), N_TMEM_COLS); }
    if (tid == 32) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(h_full + 8 * b, G_BLD_WARPS); mbar_init(h_empty + 8 * b, 1);
            mbar_init(tm_full + 8 * b, 1); mbar_init(tm_empty + 8 * b, G_EPI_WARPS);
            mbar_init(w_cp + 8 * b, 1);
        }
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nseg = p.nseg;
    const bool has_ext = p.Aext != nullptr;
    const bool use_img = p.Wimg[0] != nullptr;
    const int d_stages = (nseg == 1) ? 2 : 1;                              // TMEM: D stages | W0 hi lo | ext hi lo | W1 hi lo
    const uint32_t tmem_d = tmem_base, tmem_w = tmem_base + d_stages * GT;
    const uint32_t tmem_x_hi = tmem_w + 128, tmem_x_lo = tmem_w + 136;
    const int64_t n_tiles = (p.M + GT - 1) / GT;

    if (warp < G_EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue: thread = output n, columns = rows m
        if (warp == 0) TL(3, 0, 4);
        reg_inc<G_EPI_REGS>();                                             // 104 > the launch value 96
        if (warp == 0) TL(3, 0, 5);
        const int q = warp & 3, half = warp >> 2;
        const int n = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t bias_bits = p.bias ? __float_as_uint(__ldg(p.bias + n)) : 0u;
        const float floor_ = (p.relu == 1) ? 0.f : -INFINITY;              // relu as max(z, floor): no branch in the store loop
        const bool gate = p.relu == 2;                                     // ReLU backward: C = acc where R1 > 0, else 0
        const int nres = gate ? 1 : (p.R1 != nullptr) + (p.R1 != nullptr && p.R2 != nullptr);
        // W -> tensor memory, shared by the two warps of a lane quadrant (32-column groups 2*half, 2*half+1)
        // (with pre-split images the MMA thread brings W in: TMA bulk copy -> shared memory -> tcgen05.cp, see below)
        if (!use_img) weight_to_tmem(p.W[0], p.w_ns[0], p.w_ks[0], n, tmem_w + lane_addr, tmem_w + 64 + lane_addr, 2 * half, 2 * half + 2);
        if (!use_img && nseg == 2)
            weight_to_tmem(p.W[1], p.w_ns[1], p.w_ks[1], n, tmem_w + 144 + lane_addr, tmem_w + 208 + lane_addr, 2 * half, 2 * half + 2);
        if (has_ext && half == 1) {
            // extension weights: K step of 16 = (w0 w1)(w2 w3) then zeros; hi in columns +0..7, lo in +8..15
            uint2 h, l;
            split4(ldg4(p.Wext + n * 4), h, l);
            uint32_t xw[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) xw[c] = 0u;
            xw[0] = h.x; xw[1] = h.y; xw[8] = l.x; xw[9] = l.y;
            tmem_st16(tmem_x_hi + lane_addr, xw);
        }
        for (int b = 0; b < d_stages; ++b) {
            tmem_fill32(tmem_d + lane_addr + b * GT + half * 64, bias_bits);
            tmem_fill32(tmem_d + lane_addr + b * GT + half * 64 + 32, bias_bits);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(tm_empty); if (d_stages == 2) mbar_arrive(tm_empty + 8); }
        if (warp == 0) TL(3, 0, 0);
        int i = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int b = (d_stages == 2) ? (i & 1) : 0;
            const uint32_t ph = (uint32_t)((d_stages == 2) ? (i >> 1) : i) & 1u;
            mbar_wait(tm_full + 8 * b, ph);
            if (warp == 0) TL(3, i, 1);
            tc_fence_after();
            const uint32_t d_addr = tmem_d + lane_addr + b * GT + half * 64;
            uint32_t v[32];
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                tmem_ld32_async(d_addr + c * 32, v);
                const int64_t m0 = t * GT + half * 64 + c * 32;
                const int rows = (int)((p.M - m0 < 32) ? (p.M - m0 < 0 ? 0 : p.M - m0) : 32);     // warp-uniform
                float* cp = p.C + m0 * p.ldc + n;
                if (nres == 0) {
                    tmem_wait_ld(v);
                    if (rows == 32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { *cp = fmaxf(__uint_as_float(v[j]), floor_); cp += p.ldc; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { if (j < rows) *cp = fmaxf(__uint_as_float(v[j]), floor_); cp += p.ldc; }
                    }
                } else {
                    // residuals first, all loads in flight (a residual may alias C, but only element-wise: every thread
                    // reads exactly the elements it writes afterwards)
                    float res[32];
                    const float* r1 = p.R1 + m0 * p.ldr1 + n;
                    const float* r2 = (nres == 2) ? p.R2 + m0 * p.ldr2 + n : nullptr;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        res[j] = 0.f;
                        if (j < rows) { res[j] = *r1; if (nres == 2) res[j] += *r2; }
                        r1 += p.ldr1;
                        if (nres == 2) r2 += p.ldr2;
                    }
                    tmem_wait_ld(v);
                    if (gate) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { if (j < rows) *cp = res[j] > 0.f ? __uint_as_float(v[j]) : 0.f; cp += p.ldc; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { if (j < rows) *cp = fmaxf(__uint_as_float(v[j]), floor_) + res[j]; cp += p.ldc; }
                    }
                }
            }
            tmem_fill32(d_addr, bias_bits);
            tmem_fill32(d_addr + 32, bias_bits);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tm_empty + 8 * b);
            if (warp == 0) TL(3, i, 2);
        }
    } else if (warp < G_MMA_WARP) {
        // ------------------------------------------------------------------ builders: warp w -> rows 16w .. 16w+15 of a tile
        if (warp == G_EPI_WARPS) TL(0, 0, 4);
        reg_inc<G_BLD_REGS>();
        if (warp == G_EPI_WARPS) TL(0, 0, 5);
        const int w = warp - G_EPI_WARPS;
        const int row0 = w * 16;
        // stream of half-stages u = 0, 1, ...: stage j = u >> 1 (tile j / nseg, segment j % nseg), rows row0 + 8*(u&1) ..
        // (32-bit counters, shifts instead of divisions: 64-bit division costs ~1000 clocks a piece)
        const int sh = (nseg == 2) ? 1 : 0;
        const int my_tiles = ((int64_t)blockIdx.x < n_tiles) ? (int)(((uint32_t)(n_tiles - blockIdx.x) + gridDim.x - 1) / gridDim.x) : 0;
        const int n_half = my_tiles * nseg * 2;
        auto load8 = [&](float4 (&r)[8], int u) {
            const int j = u >> 1, ti = j >> sh, seg = j & sh;
            const int64_t m0 = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * GT + row0 + 8 * (u & 1);
            const float* base = p.A[seg] + m0 * p.lda[seg] + lane * 4;
            const int64_t ld = p.lda[seg];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ti < my_tiles && m0 + k < p.M) r[k] = ldg4(base + k * ld);
            }
        };
        float4 ra[8], rb[8];
        if (n_half > 0) load8(ra, 0);
        for (int u = 0; u < n_half; u += 2) {
            const int j = u >> 1;
            const int sb = j & 1;
            const int ti = j >> sh, seg = j & sh;
            if (w == 0) TL(0, j, 0);
            load8(rb, u + 1);
            mbar_wait(h_empty + 8 * sb, ((uint32_t)(j >> 1) & 1u) ^ 1u);
            if (w == 0) TL(0, j, 1);
            const uint32_t img = sbase + GemmSmem::SEG + sb * (2 * G_IMG);
#pragma unroll
            for (int k = 0; k < 8; ++k) store_split<GT>(img, row0 + k, lane, ra[k]);
            if (u + 2 < n_half) load8(ra, u + 2);
#pragma unroll
            for (int k = 0; k < 8; ++k) store_split<GT>(img, row0 + 8 + k, lane, rb[k]);
            if (has_ext && seg == nseg - 1) {
                // extension K step: columns 0..3 = node4[m], 4..15 = 0.  16 rows x 4 pieces of 8 bytes (hi image; lo follows)
                const int64_t t = (int64_t)blockIdx.x + (int64_t)ti * gridDim.x;
                if (use_img && ti == 0) mbar_wait(w_cp + 8 * (nseg - 1), 0);   // the weight images were staged in this region
                const uint32_t ximg = sbase + GemmSmem::EXT + (uint32_t)(ti & 1) * (2 * G_XIMG);
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int id = it * 32 + lane, r = row0 + (id >> 2), pc = id & 3;
                    uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
                    const int64_t m = t * GT + r;
                    if (pc == 0 && m < p.M) split4(ldg4(p.Aext + m * 4), hi, lo);
                    const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)pc >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)pc & 1u) * 8u;
                    sts_v2(ximg + off, hi);
                    sts_v2(ximg + G_XIMG + off, lo);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(h_full + 8 * sb);
            if (w == 0) TL(0, j, 2);
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        reg_dec<G_MMA_REGS>();
        constexpr uint32_t idesc = idesc_bf16(128, GT, 0, 0);
        if (warp == G_MMA_WARP && lane == 0) {
            if (use_img) {
                // W (hi | lo operand images, 64 KB per 128 x 128 block, already in the SWIZZLE_128B K-major tile format)
                // -> shared memory with TMA bulk copies, -> tensor memory with tcgen05.cp (32 bytes = 16 k of every row
                // per copy, the same start-address stepping the MMA uses).  Copies and MMAs execute in issue order.
                const uint32_t stage = sbase + GemmSmem::EXT;
                for (int seg = 0; seg < nseg; ++seg) {
                    if (seg > 0) mbar_wait(w_cp, 0);                       // staging area drained by the copies of segment 0
                    mbar_arrive_expect_tx(w_full, 2 * G_IMG);
#pragma unroll
                    for (int c = 0; c < 4; ++c) tma_bulk_g2s(stage + c * (G_IMG / 2), p.Wimg[seg] + c * (G_IMG / 2), G_IMG / 2, w_full);
                    mbar_wait(w_full, (uint32_t)seg & 1u);
                    tc_fence_after();
                    const uint32_t t_hi = tmem_w + (seg == 0 ? 0 : 144);
                    const uint32_t s_lo = desc_lo_sw128(stage, 16);
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        utccp_128x256b(t_hi + (q >> 3) * 64 + (q & 7) * 8,
                                       s_lo + (((q >> 3) * G_IMG + ((q & 7) >> 2) * (GT * 128) + (q & 3) * 32) >> 4), desc_hi_sw128(1024));
                    umma_commit(w_cp + 8 * seg);
                }
                TL(2, 0, 4);
            }
            int i = 0;
            int64_t j = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
                const int b = (d_stages == 2) ? (i & 1) : 0;
                const uint32_t ph = (uint32_t)((d_stages == 2) ? (i >> 1) : i) & 1u;
                mbar_wait(tm_empty + 8 * b, ph);
                TL(2, i, 0);
                for (int seg = 0; seg < nseg; ++seg, ++j) {
                    const int sb = (int)(j & 1);
                    mbar_wait(h_full + 8 * sb, (uint32_t)(j >> 1) & 1u);
                    tc_fence_after();
                    const uint32_t b_lo = desc_lo_sw128(sbase + GemmSmem::SEG + sb * (2 * G_IMG), 16);
                    constexpr uint32_t b_hi = desc_hi_sw128(1024);
                    const uint32_t w_hi = tmem_w + (seg == 0 ? 0 : 144), w_lo = w_hi + 64;
#pragma unroll
                    for (int prod = 0; prod < 3; ++prod) {                 // hi*hi + hi*lo + lo*hi
                        const uint32_t a = (prod == 2) ? w_lo : w_hi;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            umma_bf16_ts_lh(tmem_d + b * GT, a + ks * 8,
                                            b_lo + (((prod == 1 ? G_IMG : 0) + (ks >> 2) * (GT * 128) + (ks & 3) * 32) >> 4), b_hi, idesc, 1u);
                    }
                    if (has_ext && seg == nseg - 1) {
                        const uint32_t x_lo = desc_lo_sw128(sbase + GemmSmem::EXT + (uint32_t)(i & 1) * (2 * G_XIMG), 16);
#pragma unroll
                        for (int prod = 0; prod < 3; ++prod)
                            umma_bf16_ts_lh(tmem_d + b * GT, (prod == 2) ? tmem_x_lo : tmem_x_hi,
                                            x_lo + ((prod == 1 ? G_XIMG : 0) >> 4), b_hi, idesc, 1u);
                    }
                    umma_commit(h_empty + 8 * sb);
                }
                umma_commit(tm_full + 8 * b);
                TL(2, i, 2);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { TL(3, 0, 7); tmem_dealloc(tmem_base, N_TMEM_COLS); }
}

// ================================================================================================================
// weight images: W (fp32, element (n, k) at W[n*ns + k*ks], 128 x 128) * scale -> bf16 hi | lo operand images in the
// SWIZZLE_128B K-major tile format (tile_off<128>(n, k); 32 KB each), ONCE per optimizer step for every weight block
// the node contractions of a solver pass use.  The GEMM kernels then fetch an image with TMA bulk copies and move it
// to tensor memory with tcgen05.cp instead of 148 CTAs x 100 launches re-splitting the same fp32 matrix in registers.
// ================================================================================================================
constexpr int IMG_MAX_TASKS = 64;
struct WImgTask { const float* W; int64_t ns, ks; unsigned char* img; float scale; };
struct WImgGroup { WImgTask t[IMG_MAX_TASKS]; };

__global__ void __launch_bounds__(256) weight_images_kernel(const __grid_constant__ WImgGroup grp) {
    const WImgTask& T = grp.t[blockIdx.x];
    const int n0 = blockIdx.y * 8;                                         // 8 rows x 32 groups of 4 consecutive k per CTA
    const bool k_fast = T.ks == 1;                                         // lanes along the contiguous axis of W
    const int g = (int)threadIdx.x;
    const int n = k_fast ? n0 + (g >> 5) : (n0 & ~31) + (g & 31);
    const int k = 4 * (k_fast ? (g & 31) : ((n0 & 31) + (g >> 5)));
    const float* q = T.W + (int64_t)n * T.ns + (int64_t)k * T.ks;
    const float4 v = make_float4(__ldg(q) * T.scale, __ldg(q + T.ks) * T.scale, __ldg(q + 2 * T.ks) * T.scale, __ldg(q + 3 * T.ks) * T.scale);
    uint2 hi, lo;
    split4(v, hi, lo);
    unsigned char* dst = T.img + tile_off<128>(n, k);
    *reinterpret_cast<uint2*>(dst) = hi;
    *reinterpret_cast<uint2*>(dst + G_IMG) = lo;
}

// ================================================================================================================
// wgrad
// ================================================================================================================
constexpr int W_EPI_WARPS = 4, W_BLD_WARPS = 8, W_MMA_WARP = 12, W_THREADS = 512;
constexpr int W_EPI_REGS = 104, W_BLD_REGS = 184, W_MMA_REGS = 40;         // 4*32*104 + 8*32*184 + 4*32*40 = 65536 = 512*128
constexpr int WT = 64;                                                     // node rows per K tile
constexpr uint32_t W_IMG = 2 * WT * 128;                                   // one [64][128] bf16 image = 16 KB
constexpr uint32_t W_XIMG = 16 * 128;                                      // extension tile, TRANSPOSED: [16 features][64 rows] = 2 KB

struct NodeWgradArgs {
    const float* A; int64_t lda; const float* B; int64_t ldb; const float* Bext; int want_bias;
    float* dW; int64_t ldw; float* dWext; int64_t ldwext; float* dbias; int64_t M;
};
// Several independent contractions in ONE launch: the CTAs are divided between the tasks (task k owns CTAs
// cta_begin[k] .. cta_begin[k+1]-1), so an output tile is summed over grid/n_tasks partials instead of grid: the
// atomic flush of the 64 KB partial tiles -- half the time of a single-task launch at 36 k rows -- shrinks with it.
constexpr int W_MAX_TASKS = 8;
struct NodeWgradGroup {
    NodeWgradArgs t[W_MAX_TASKS];
    int cta_begin[W_MAX_TASKS + 1];
};
struct WgradSmem {
    static constexpr uint32_t STAGE = 4 * W_IMG + 2 * W_XIMG;             // A hi lo | B hi lo | ext hi lo  = 68 KB
    static constexpr uint32_t BAR = 2 * STAGE;
    static constexpr uint32_t TOTAL = BAR + 128;
};

__global__ void __launch_bounds__(W_THREADS, 1) node_wgrad_tc_kernel(const __grid_constant__ NodeWgradGroup grp) {
    int task = 0;
#pragma unroll
    for (int k = 1; k < W_MAX_TASKS; ++k) task += ((int)blockIdx.x >= grp.cta_begin[k]) ? 1 : 0;
    const NodeWgradArgs p = grp.t[task];
    const uint32_t bx = blockIdx.x - (uint32_t)grp.cta_begin[task];                 // this CTA within its task
    const uint32_t gx = (uint32_t)(grp.cta_begin[task + 1] - grp.cta_begin[task]);  // CTAs of the task
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(sm);
    const uint32_t bar0 = sbase + WgradSmem::BAR;
    const uint32_t s_full = bar0, s_empty = bar0 + 16, all_done = bar0 + 32;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + WgradSmem::BAR + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool has_b = p.B != nullptr, has_ext = (p.Bext != nullptr) || p.want_bias;

    if (warp == 0) { TL(3, 0, 6); tmem_alloc(smem_u32(tmem_slot), 256); }
    if (tid == 32) {
        for (int b = 0; b < 2; ++b) { mbar_init(s_full + 8 * b, W_BLD_WARPS); mbar_init(s_empty + 8 * b, 1); }
        mbar_init(all_done, 1);
        fence_mbar_init();
    }
    // the extension tiles only ever receive features 0..4: zero them once (rows 5..15 stay zero)
    for (uint32_t o = tid * 16; o < 2 * 2 * W_XIMG; o += W_THREADS * 16) {
        const uint32_t stage = o / (2 * W_XIMG), off = o % (2 * W_XIMG);
        *reinterpret_cast<uint4*>(sm + stage * WgradSmem::STAGE + 4 * W_IMG + off) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d = tmem_base, tmem_dx = tmem_base + 128;
    const int64_t n_tiles = (p.M + WT - 1) / WT;
    const int64_t my_tiles = ((int64_t)bx < n_tiles) ? (int64_t)(((uint32_t)(n_tiles - bx) + gx - 1) / gx) : 0;

    if (warp < W_EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue (once): lane = row i of dW
        reg_dec<W_EPI_REGS>();
        const int i_row = warp * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        if (my_tiles > 0) {
            mbar_wait(all_done, 0);
            if (warp == 0) TL(3, 0, 1);
            tc_fence_after();
            if (has_b) {
                const bool vec = (p.ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dW) & 15) == 0;
                // every CTA adds its 128x128 partial onto the same 64 KB: walk the chunks in a per-CTA rotated order so that
                // the CTAs do not all hammer the same addresses at the same moment (same-address atomics serialise in L2)
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int chunk = (cc + (int)bx) & 3;
                    uint32_t v[32];
                    tmem_ld32_async(tmem_d + lane_addr + chunk * 32, v);
                    tmem_wait_ld(v);
                    float* row = p.dW + (int64_t)i_row * p.ldw + chunk * 32;
                    if (vec) {
                        if ((bx >> 2) & 1) {
#pragma unroll
                            for (int m = 7; m >= 0; --m)
                                red_add_v4(row + 4 * m, make_float4(__uint_as_float(v[4 * m]), __uint_as_float(v[4 * m + 1]),
                                                                     __uint_as_float(v[4 * m + 2]), __uint_as_float(v[4 * m + 3])));
                        } else {
#pragma unroll
                            for (int m = 0; m < 8; ++m)
                                red_add_v4(row + 4 * m, make_float4(__uint_as_float(v[4 * m]), __uint_as_float(v[4 * m + 1]),
                                                                     __uint_as_float(v[4 * m + 2]), __uint_as_float(v[4 * m + 3])));
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < 32; ++m) atomicAdd(row + m, __uint_as_float(v[m]));
                    }
                }
            }
            if (has_ext) {
                uint32_t v[32];                                            // only the first 16 columns are the extension
                tmem_ld32_async(tmem_dx + lane_addr, v);
                tmem_wait_ld(v);
                if (p.dWext) {
#pragma unroll
                    for (int f = 0; f < 4; ++f) atomicAdd(p.dWext + (int64_t)i_row * p.ldwext + f, __uint_as_float(v[f]));
                }
                if (p.dbias) atomicAdd(p.dbias + i_row, __uint_as_float(v[4]));
            }
            if (warp == 0) TL(3, 0, 2);
        }
    } else if (warp < W_MMA_WARP) {
        // ------------------------------------------------------------------ builders: warp w -> rows 8w .. 8w+7 of the A and B tiles
        reg_inc<W_BLD_REGS>();
        const int w = warp - W_EPI_WARPS;
        const int row0 = w * 8;
        // per tile: 8 rows of A and 8 rows of B per warp; the NEXT tile's 16 rows are already in flight while this
        // tile is converted (two register sets swapped by copy: the loads were issued a whole tile ago)
        auto load16 = [&](float4 (&ra)[8], float4 (&rb)[8], int64_t ti) {
            const int64_t t = (int64_t)bx + ti * gx;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t m = t * WT + row0 + k;
                ra[k] = rb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ti < my_tiles && m < p.M) {
                    ra[k] = ldg4(p.A + m * p.lda + lane * 4);
                    if (has_b) rb[k] = ldg4(p.B + m * p.ldb + lane * 4);
                }
            }
        };
        // convert + stage one tile from a register set whose loads were issued a WHOLE tile earlier
        auto build_tile = [&](int64_t ti, const float4 (&ra)[8], const float4 (&rb)[8]) {
            const int sb = (int)(ti & 1);
            const int64_t t = (int64_t)bx + ti * gx;
            if (w == 0) TL(0, (int)ti, 0);
            mbar_wait(s_empty + 8 * sb, ((uint32_t)(ti >> 1) & 1u) ^ 1u);
            if (w == 0) TL(0, (int)ti, 1);
            const uint32_t stage = sbase + sb * WgradSmem::STAGE;
#pragma unroll
            for (int k = 0; k < 8; ++k) store_split<WT>(stage, row0 + k, lane, ra[k]);
            if (has_b) {
#pragma unroll
                for (int k = 0; k < 8; ++k) store_split<WT>(stage + 2 * W_IMG, row0 + k, lane, rb[k]);
            }
            if (has_ext && lane < 8) {
                // extension tile, K-major and transposed: feature f (row of the tile), node r (column): f 0..3 = node4, 4 = 1
                const int r = row0 + lane;
                const int64_t m = t * WT + r;
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                float one = 0.f;
                if (m < p.M) { one = 1.f; if (p.Bext) x = ldg4(p.Bext + m * 4); }
                const float f5[5] = {x.x, x.y, x.z, x.w, one};
                const uint32_t ximg = stage + 4 * W_IMG;
#pragma unroll
                for (int f = 0; f < 5; ++f) {
                    const __nv_bfloat16 h = __float2bfloat16_rn(f5[f]);
                    const __nv_bfloat16 l = __float2bfloat16_rn(f5[f] - __bfloat162float(h));
                    const uint32_t off = (uint32_t)f * 128u + ((((uint32_t)r >> 3) ^ (uint32_t)f) << 4) + ((uint32_t)r & 7u) * 2u;
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(ximg + off), "h"(*reinterpret_cast<const unsigned short*>(&h)));
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(ximg + W_XIMG + off), "h"(*reinterpret_cast<const unsigned short*>(&l)));
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_full + 8 * sb);
            if (w == 0) TL(0, (int)ti, 2);
        };
        // Two register sets used ALTERNATELY (the loop is unrolled by two): while one tile is converted, the 16 rows of the
        // next one are in flight into the other set and are first touched a whole tile later.  (Swapping the sets by copy
        // at the end of each trip touched the fresh loads after one tile's worth of stores, ~500 clk: long_scoreboard was
        // the top stall of this kernel, profiles/r02b_ncu_full_summary.txt.)
        float4 ra[8], rb[8], na[8], nb[8];
        load16(ra, rb, 0);
        for (int64_t ti = 0; ti < my_tiles; ti += 2) {
            load16(na, nb, ti + 1);
            build_tile(ti, ra, rb);
            if (ti + 1 < my_tiles) {
                load16(ra, rb, ti + 2);
                build_tile(ti + 1, na, nb);
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        reg_dec<W_MMA_REGS>();
        constexpr uint32_t idesc_main = idesc_bf16(128, 128, 1, 1);        // A^T B: both tiles MN-major
        constexpr uint32_t idesc_ext = idesc_bf16(128, 16, 1, 0);          // A^T ext: ext tile K-major (transposed storage)
        if (warp == W_MMA_WARP && lane == 0) {
            for (int64_t ti = 0; ti < my_tiles; ++ti) {
                const int sb = (int)(ti & 1);
                mbar_wait(s_full + 8 * sb, (uint32_t)(ti >> 1) & 1u);
                TL(2, (int)ti, 0);
                tc_fence_after();
                const uint32_t a_addr = sbase + sb * WgradSmem::STAGE, b_addr = a_addr + 2 * W_IMG, x_addr = a_addr + 4 * W_IMG;
#pragma unroll
                constexpr uint32_t d_hi = desc_hi_sw128(1024);
                const uint32_t a_lo = desc_lo_sw128(a_addr, WT * 128), bm_lo = desc_lo_sw128(b_addr, WT * 128), x_lo = desc_lo_sw128(x_addr, 16);
#pragma unroll
                for (int prod = 0; prod < 3; ++prod) {
                    if (has_b) {
#pragma unroll
                        for (int ks = 0; ks < WT / 16; ++ks)               // 16 node rows per step
                            umma_bf16_lh(tmem_d, a_lo + (((prod == 2 ? W_IMG : 0) + ks * 16 * 128) >> 4), d_hi,
                                         bm_lo + (((prod == 1 ? W_IMG : 0) + ks * 16 * 128) >> 4), d_hi, idesc_main,
                                         (ti > 0 || prod > 0 || ks > 0) ? 1u : 0u);
                    }
                    if (has_ext) {
#pragma unroll
                        for (int ks = 0; ks < WT / 16; ++ks)
                            umma_bf16_lh(tmem_dx, a_lo + (((prod == 2 ? W_IMG : 0) + ks * 16 * 128) >> 4), d_hi,
                                         x_lo + (((prod == 1 ? W_XIMG : 0) + ks * 32) >> 4), d_hi, idesc_ext,
                                         (ti > 0 || prod > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                umma_commit(s_empty + 8 * sb);
                TL(2, (int)ti, 2);
            }
            if (my_tiles > 0) umma_commit(all_done);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { TL(3, 0, 7); tmem_dealloc(tmem_base, 256); }
}

}  // namespace mmpde

using namespace mmpde;

#ifdef MMPDE_TIMELINE
extern "C" int mmpde_debug_timeline_node(long long* buf) { return (int)cudaMemcpyToSymbol(g_timeline, &buf, sizeof(buf)); }
#endif

extern "C" int mmpde_weight_images(const mmpde_wimg_task* tasks, int n_tasks, void* stream) {
    if (n_tasks < 0 || (n_tasks > 0 && tasks == nullptr)) return MMPDE_EINVAL;
    for (int k = 0; k < n_tasks; ++k)
        if (tasks[k].W == nullptr || tasks[k].image == nullptr || (reinterpret_cast<uintptr_t>(tasks[k].image) & 127)) return MMPDE_EINVAL;
    for (int k0 = 0; k0 < n_tasks; k0 += IMG_MAX_TASKS) {
        WImgGroup g;
        const int n = (int)imin64(IMG_MAX_TASKS, n_tasks - k0);
        for (int k = 0; k < n; ++k) {
            const mmpde_wimg_task& t = tasks[k0 + k];
            g.t[k].W = t.W; g.t[k].ns = t.w_ns; g.t[k].ks = t.w_ks; g.t[k].img = static_cast<unsigned char*>(t.image); g.t[k].scale = t.scale;
        }
        for (int k = n; k < IMG_MAX_TASKS; ++k) g.t[k] = g.t[0];
        weight_images_kernel<<<dim3(n, 16), 256, 0, (cudaStream_t)stream>>>(g);
        MMPDE_CHECK_LAUNCH();
    }
    return MMPDE_OK;
}

static int node_gemm_launch(const float* A0, int64_t lda0, const float* A1, int64_t lda1,
                            const float* W0, int64_t w0_ns, int64_t w0_ks, const float* W1, int64_t w1_ns, int64_t w1_ks,
                            const void* img0, const void* img1,
                            const float* Aext, const float* Wext, const float* bias, int relu,
                            const float* R1, int64_t ldr1, const float* R2, int64_t ldr2,
                            float* C, int64_t ldc, int64_t M, void* stream) {
    const bool img = img0 != nullptr;
    if (M < 0 || A0 == nullptr || (W0 == nullptr && !img) || C == nullptr) return MMPDE_EINVAL;
    if (img ? ((A1 == nullptr) != (img1 == nullptr)) : ((A1 == nullptr) != (W1 == nullptr))) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(img0) | reinterpret_cast<uintptr_t>(img1)) & 127) return MMPDE_EINVAL;
    if ((Aext == nullptr) != (Wext == nullptr)) return MMPDE_EINVAL;
    if ((lda0 & 3) || (A1 && (lda1 & 3))) return MMPDE_EINVAL;
    if (relu < 0 || relu > 2 || (relu == 2 && (R1 == nullptr || R2 != nullptr))) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(A0) | reinterpret_cast<uintptr_t>(A1) | reinterpret_cast<uintptr_t>(Aext) |
         reinterpret_cast<uintptr_t>(Wext)) & 15) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    constexpr size_t smem = GemmSmem::TOTAL + 1024;
    MMPDE_ENSURE_SMEM(node_gemm_tc_kernel, smem);
    NodeGemmArgs p;
    p.A[0] = A0; p.A[1] = A1; p.lda[0] = lda0; p.lda[1] = lda1; p.nseg = A1 ? 2 : 1;
    p.W[0] = W0; p.W[1] = W1; p.w_ns[0] = w0_ns; p.w_ks[0] = w0_ks; p.w_ns[1] = w1_ns; p.w_ks[1] = w1_ks;
    p.Wimg[0] = static_cast<const unsigned char*>(img0); p.Wimg[1] = static_cast<const unsigned char*>(img1);
    p.Aext = Aext; p.Wext = Wext; p.bias = bias; p.relu = relu; p.R1 = R1; p.ldr1 = ldr1; p.R2 = R2; p.ldr2 = ldr2;
    p.C = C; p.ldc = ldc; p.M = M;
    const int64_t n_tiles = (M + GT - 1) / GT;
    node_gemm_tc_kernel<<<(int)imin64(n_tiles, persistent_ctas()), G_THREADS, smem, (cudaStream_t)stream>>>(p);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_node_gemm(const float* A0, int64_t lda0, const float* A1, int64_t lda1,
                               const float* W0, int64_t w0_ns, int64_t w0_ks, const float* W1, int64_t w1_ns, int64_t w1_ks,
                               const float* Aext, const float* Wext, const float* bias, int relu,
                               const float* R1, int64_t ldr1, const float* R2, int64_t ldr2,
                               float* C, int64_t ldc, int64_t M, void* stream) {
    return node_gemm_launch(A0, lda0, A1, lda1, W0, w0_ns, w0_ks, W1, w1_ns, w1_ks, nullptr, nullptr, Aext, Wext, bias, relu,
                            R1, ldr1, R2, ldr2, C, ldc, M, stream);
}

extern "C" int mmpde_node_gemm_img(const float* A0, int64_t lda0, const float* A1, int64_t lda1,
                                   const void* image0, const void* image1,
                                   const float* Aext, const float* Wext, const float* bias, int relu,
                                   const float* R1, int64_t ldr1, const float* R2, int64_t ldr2,
                                   float* C, int64_t ldc, int64_t M, void* stream) {
    if (image0 == nullptr) return MMPDE_EINVAL;
    return node_gemm_launch(A0, lda0, A1, lda1, nullptr, 0, 0, nullptr, 0, 0, image0, image1, Aext, Wext, bias, relu,
                            R1, ldr1, R2, ldr2, C, ldc, M, stream);
}

static int check_wgrad_task(const mmpde_wgrad_task& t) {
    if (t.M < 0) return MMPDE_EINVAL;
    if (t.M == 0) return MMPDE_OK;                                         // empty part: nothing to read, pointers may be NULL
    if (t.A == nullptr || (t.lda & 3) || (t.B && (t.ldb & 3))) return MMPDE_EINVAL;
    if ((t.B == nullptr) != (t.dW == nullptr) || (t.Bext == nullptr) != (t.dWext == nullptr)) return MMPDE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(t.A) | reinterpret_cast<uintptr_t>(t.B) | reinterpret_cast<uintptr_t>(t.Bext)) & 15) return MMPDE_EINVAL;
    return MMPDE_OK;
}

extern "C" int mmpde_node_wgrad_grouped(const mmpde_wgrad_task* tasks, int n_tasks, void* stream) {
    if (n_tasks < 0 || (n_tasks > 0 && tasks == nullptr)) return MMPDE_EINVAL;
    for (int k = 0; k < n_tasks; ++k)
        if (int rc = check_wgrad_task(tasks[k])) return rc;
    constexpr size_t smem = WgradSmem::TOTAL + 1024;
    MMPDE_ENSURE_SMEM(node_wgrad_tc_kernel, smem);
    int k0 = 0;
    while (k0 < n_tasks) {
        // next launch: up to W_MAX_TASKS tasks that have work
        NodeWgradGroup g;
        int64_t tiles[W_MAX_TASKS], total = 0;
        int n = 0;
        for (; k0 < n_tasks && n < W_MAX_TASKS; ++k0) {
            const mmpde_wgrad_task& t = tasks[k0];
            if (t.M == 0 || (t.B == nullptr && t.Bext == nullptr && t.dbias == nullptr)) continue;
            NodeWgradArgs& a = g.t[n];
            a.A = t.A; a.lda = t.lda; a.B = t.B; a.ldb = t.ldb; a.Bext = t.Bext; a.want_bias = t.dbias != nullptr;
            a.dW = t.dW; a.ldw = t.ldw; a.dWext = t.dWext; a.ldwext = t.ldwext; a.dbias = t.dbias; a.M = t.M;
            tiles[n] = (t.M + WT - 1) / WT;
            total += tiles[n];
            ++n;
        }
        if (n == 0) break;
        // CTAs per task in proportion to its K tiles (every task at least one, none more than it has tiles)
        const int64_t grid = imin64(total, imax64((int64_t)persistent_ctas(), (int64_t)n));
        int64_t given = 0;
        g.cta_begin[0] = 0;
        for (int k = 0; k < n; ++k) {
            int64_t share = (k == n - 1) ? grid - given : (tiles[k] * grid + total / 2) / total;
            const int64_t left_for_rest = (int64_t)(n - 1 - k);
            if (share < 1) share = 1;
            if (share > tiles[k]) share = tiles[k];
            if (given + share > grid - left_for_rest) share = grid - left_for_rest - given;
            if (share < 1) share = 1;
            given += share;
            g.cta_begin[k + 1] = (int)given;
        }
        for (int k = n; k < W_MAX_TASKS; ++k) { g.t[k] = g.t[0]; g.cta_begin[k + 1] = 0x7fffffff; }
        node_wgrad_tc_kernel<<<(int)given, W_THREADS, smem, (cudaStream_t)stream>>>(g);
        MMPDE_CHECK_LAUNCH();
    }
    return MMPDE_OK;
}

extern "C" int mmpde_node_wgrad(const float* A, int64_t lda, const float* B, int64_t ldb, const float* Bext,
                                float* dW, int64_t ldw, float* dWext, int64_t ldwext, float* dbias, int64_t M, void* stream) {
    mmpde_wgrad_task t;
    t.A = A; t.lda = lda; t.B = B; t.ldb = ldb; t.Bext = Bext; t.dW = dW; t.ldw = ldw; t.dWext = dWext; t.ldwext = ldwext;
    t.dbias = dbias; t.M = M;
    return mmpde_node_wgrad_grouped(&t, 1, stream);
}
