// Fused k-NN interpolation between the moved mesh and the reference mesh.
// Replaces the gather points[indices]/labels[indices] + ItpNet modes '1'/'2' + weighted sum
// (/root/reference/data_creator_2d.py:77-83, /root/reference/interpolate.py:79-93) and their autograd.
// A warp carries 8 queries through the 62->128->64->30 tanh MLP with the weights resident in shared
// memory; the neighbour coordinates/values are gathered straight from the source arrays (no [Q,30,2] tensor).
#include "common.cuh"

namespace mmpde {

constexpr int KN = 30;                 // neighbours per query (interpolate.py:8)
constexpr int IN0 = 62, H1 = 128, H2 = 64;
constexpr int QW = 8;                  // queries per warp
constexpr int IW = 8;                  // warps per CTA
constexpr int P_WA = 0, P_BA = 7936, P_WB = 8064, P_BB = 16256, P_WC = 16320, P_BC = 18240, P_TOTAL = 18270;
constexpr int LDB = 65, LDC = 33;      // padded strides of the transposed Wb / Wc images

struct ItpSmem {
    float waT[IN0 * H1];               // [k][o]
    float wbT[H1 * LDB];               // [k][o], stride 65
    float wcT[H2 * LDC];               // [k][o], stride 33 (o padded to 32)
    float ba[H1], bb[H2], bc[32];
};

struct ItpWarp {                        // activations of one warp's 8 queries, [feature][query]
    float p[64 * QW];
    float ha[H1 * QW];
    float hb[H2 * QW];
    float w[32 * QW];                  // interpolation weights, later g_w
};
struct ItpWarpGrad {
    float gza[H1 * QW];
    float gzb[H2 * QW];
};

__device__ __forceinline__ void itp_load_params(ItpSmem& s, const float* __restrict__ params) {
    for (int i = threadIdx.x; i < H1 * IN0; i += blockDim.x) { int o = i / IN0, k = i - o * IN0; s.waT[k * H1 + o] = __ldg(params + P_WA + i); }
    for (int i = threadIdx.x; i < H2 * H1; i += blockDim.x) { int o = i / H1, k = i - o * H1; s.wbT[k * LDB + o] = __ldg(params + P_WB + i); }
    for (int i = threadIdx.x; i < H2 * LDC; i += blockDim.x) s.wcT[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < KN * H2; i += blockDim.x) { int o = i / H2, k = i - o * H2; s.wcT[k * LDC + o] = __ldg(params + P_WC + i); }
    for (int i = threadIdx.x; i < H1; i += blockDim.x) s.ba[i] = __ldg(params + P_BA + i);
    for (int i = threadIdx.x; i < H2; i += blockDim.x) s.bb[i] = __ldg(params + P_BB + i);
    for (int i = threadIdx.x; i < 32; i += blockDim.x) s.bc[i] = i < KN ? __ldg(params + P_BC + i) : 0.f;
    __syncthreads();
}

// Forward of the warp's 8 queries.  Returns in `val` the gathered source value for (lane = neighbour k) of each query.
__device__ __forceinline__ void itp_forward_warp(const ItpSmem& s, ItpWarp& a, const float2* __restrict__ src_xy,
                                                 const float* __restrict__ src_val, const float2* __restrict__ qry_xy,
                                                 const int* __restrict__ idx, int64_t q0, int64_t nq, int lane,
                                                 float (&val)[QW], int (&nbr)[QW]) {
#pragma unroll
    for (int qq = 0; qq < QW; ++qq) {
        int64_t q = q0 + qq;
        float2 xy = make_float2(0.f, 0.f);
        val[qq] = 0.f; nbr[qq] = -1;
        if (q < nq) {
            if (lane < KN) {
                int j = __ldg(idx + q * KN + lane);
                nbr[qq] = j;
                xy = __ldg(src_xy + j);
                val[qq] = __ldg(src_val + j);
            } else if (lane == KN) {
                xy = __ldg(qry_xy + q);
            }
        }
        a.p[(2 * lane) * QW + qq] = xy.x;          // lane 31 writes rows 62,63 = 0 (padding)
        a.p[(2 * lane + 1) * QW + qq] = xy.y;
    }
    __syncwarp();
    {   // layer a: 62 -> 128, lane owns outputs lane + 32*i
        float acc[4][QW];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int qq = 0; qq < QW; ++qq) acc[i][qq] = s.ba[lane + 32 * i];
        for (int k = 0; k < IN0; ++k) {
            float4 x0 = *reinterpret_cast<const float4*>(&a.p[k * QW]), x1 = *reinterpret_cast<const float4*>(&a.p[k * QW + 4]);
            float x[QW] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float wv = s.waT[k * H1 + lane + 32 * i];
#pragma unroll
                for (int qq = 0; qq < QW; ++qq) acc[i][qq] = fmaf(wv, x[qq], acc[i][qq]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int qq = 0; qq < QW; ++qq) a.ha[(lane + 32 * i) * QW + qq] = tanhf(acc[i][qq]);
    }
    __syncwarp();
    {   // layer b: 128 -> 64
        float acc[2][QW];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int qq = 0; qq < QW; ++qq) acc[i][qq] = s.bb[lane + 32 * i];
        for (int k = 0; k < H1; ++k) {
            float4 x0 = *reinterpret_cast<const float4*>(&a.ha[k * QW]), x1 = *reinterpret_cast<const float4*>(&a.ha[k * QW + 4]);
            float x[QW] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float wv = s.wbT[k * LDB + lane + 32 * i];
#pragma unroll
                for (int qq = 0; qq < QW; ++qq) acc[i][qq] = fmaf(wv, x[qq], acc[i][qq]);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int qq = 0; qq < QW; ++qq) a.hb[(lane + 32 * i) * QW + qq] = tanhf(acc[i][qq]);
    }
    __syncwarp();
    {   // layer c: 64 -> 30 (lanes 30,31 compute zeros)
        float acc[QW];
#pragma unroll
        for (int qq = 0; qq < QW; ++qq) acc[qq] = s.bc[lane];
        for (int k = 0; k < H2; ++k) {
            float4 x0 = *reinterpret_cast<const float4*>(&a.hb[k * QW]), x1 = *reinterpret_cast<const float4*>(&a.hb[k * QW + 4]);
            float x[QW] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            float wv = s.wcT[k * LDC + lane];
#pragma unroll
            for (int qq = 0; qq < QW; ++qq) acc[qq] = fmaf(wv, x[qq], acc[qq]);
        }
#pragma unroll
        for (int qq = 0; qq < QW; ++qq) a.w[lane * QW + qq] = acc[qq];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(IW * 32, 1) itp_fwd_kernel(const float2* __restrict__ src_xy, const float* __restrict__ src_val,
                                                             const float2* __restrict__ qry_xy, const int* __restrict__ idx,
                                                             int64_t nq, const float* __restrict__ params, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ItpSmem& s = *reinterpret_cast<ItpSmem*>(smem_raw);
    ItpWarp* wa = reinterpret_cast<ItpWarp*>(smem_raw + sizeof(ItpSmem));
    itp_load_params(s, params);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ItpWarp& a = wa[warp];
    for (int64_t q0 = ((int64_t)blockIdx.x * IW + warp) * QW; q0 < nq; q0 += (int64_t)gridDim.x * IW * QW) {
        float val[QW]; int nbr[QW];
        itp_forward_warp(s, a, src_xy, src_val, qry_xy, idx, q0, nq, lane, val, nbr);
#pragma unroll
        for (int qq = 0; qq < QW; ++qq) {
            float t = warp_sum(a.w[lane * QW + qq] * val[qq]);
            if (lane == 0 && q0 + qq < nq) out[q0 + qq] = t;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(IW * 32, 1) itp_bwd_kernel(const float2* __restrict__ src_xy, const float* __restrict__ src_val,
                                                             const float2* __restrict__ qry_xy, const int* __restrict__ idx,
                                                             int64_t nq, const float* __restrict__ params,
                                                             const float* __restrict__ g_out, float* __restrict__ g_params,
                                                             float* __restrict__ g_src_val) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ItpSmem& s = *reinterpret_cast<ItpSmem*>(smem_raw);
    ItpWarp* wa = reinterpret_cast<ItpWarp*>(smem_raw + sizeof(ItpSmem));
    ItpWarpGrad* wg = reinterpret_cast<ItpWarpGrad*>(smem_raw + sizeof(ItpSmem) + sizeof(ItpWarp) * IW);
    itp_load_params(s, params);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    ItpWarp& a = wa[warp];
    ItpWarpGrad& g = wg[warp];

    // thread-owned parameter-gradient tiles (accumulated over the whole kernel)
    float dWa[8][4], dWb[4][8], dWc[2][4], dbias = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dWa[i][j] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dWb[i][j] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dWc[i][j] = 0.f;
    // Tiles: outputs og*8.. (dWa) / og*4.. (dWb) / og*2.. (dWc) x inputs kl + 16 j.  The inputs of neighbouring lanes are
    // NEIGHBOURING feature rows (8 floats apart) and every row is read as float4 over 4 queries: a quarter-warp covers
    // 256 contiguous bytes (2-way bank conflict).  The first version gave each lane 4 / 8 consecutive rows and read
    // scalars: all 16 lanes on one bank, and this phase was ~60 % of the kernel.
    const int og = tid >> 4, kl = tid & 15;
    const int ta_o = og * 8, tb_o = og * 4, tc_o = og * 2;
    const int64_t stride = (int64_t)gridDim.x * IW * QW;
    const int64_t n_iter = (nq + stride - 1) / stride;           // uniform trip count: block-wide barriers inside
    for (int64_t it = 0; it < n_iter; ++it) {
        const int64_t q0 = it * stride + ((int64_t)blockIdx.x * IW + warp) * QW;
        float val[QW]; int nbr[QW];
        itp_forward_warp(s, a, src_xy, src_val, qry_xy, idx, q0, nq, lane, val, nbr);
        // g_w[k][q] = g_out[q] * val ; g_src_val[nbr] += g_out[q] * w[k][q]
#pragma unroll
        for (int qq = 0; qq < QW; ++qq) {
            float go = (q0 + qq < nq) ? __ldg(g_out + q0 + qq) : 0.f;
            float wv = a.w[lane * QW + qq];
            if (g_src_val && nbr[qq] >= 0) atomicAdd(g_src_val + nbr[qq], go * wv);
            a.w[lane * QW + qq] = (lane < KN) ? go * val[qq] : 0.f;
        }
        __syncwarp();
        {   // g_zb[k][q] = (1 - hb^2) * sum_o g_w[o][q] * Wc[o][k],  lane owns k = lane + 32*i
            float acc[2][QW];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int qq = 0; qq < QW; ++qq) acc[i][qq] = 0.f;
            for (int o = 0; o < KN; ++o) {
                float4 x0 = *reinterpret_cast<const float4*>(&a.w[o * QW]), x1 = *reinterpret_cast<const float4*>(&a.w[o * QW + 4]);
                float x[QW] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float wv = s.wcT[(lane + 32 * i) * LDC + o];
#pragma unroll
                    for (int qq = 0; qq < QW; ++qq) acc[i][qq] = fmaf(wv, x[qq], acc[i][qq]);
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int qq = 0; qq < QW; ++qq) {
                    float h = a.hb[(lane + 32 * i) * QW + qq];
                    g.gzb[(lane + 32 * i) * QW + qq] = acc[i][qq] * (1.f - h * h);
                }
        }
        __syncwarp();
        {   // g_za[k][q] = (1 - ha^2) * sum_o g_zb[o][q] * Wb[o][k]
            float acc[4][QW];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int qq = 0; qq < QW; ++qq) acc[i][qq] = 0.f;
            for (int o = 0; o < H2; ++o) {
                float4 x0 = *reinterpret_cast<const float4*>(&g.gzb[o * QW]), x1 = *reinterpret_cast<const float4*>(&g.gzb[o * QW + 4]);
                float x[QW] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float wv = s.wbT[(lane + 32 * i) * LDB + o];
#pragma unroll
                    for (int qq = 0; qq < QW; ++qq) acc[i][qq] = fmaf(wv, x[qq], acc[i][qq]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int qq = 0; qq < QW; ++qq) {
                    float h = a.ha[(lane + 32 * i) * QW + qq];
                    g.gza[(lane + 32 * i) * QW + qq] = acc[i][qq] * (1.f - h * h);
                }
        }
        __syncthreads();
        // CTA-wide weight gradients over the 64 queries staged by the 8 warps
        for (int w = 0; w < IW; ++w) {
            const ItpWarp& aw = wa[w];
            const ItpWarpGrad& gw = wg[w];
#pragma unroll 1
            for (int h = 0; h < QW; h += 4) {                              // 4 queries at a time
                auto ld4 = [&](const float* base, int row) { return *reinterpret_cast<const float4*>(base + row * QW + h); };
                {
                    float4 ga[8], xa[4];
#pragma unroll
                    for (int i = 0; i < 8; ++i) ga[i] = ld4(gw.gza, ta_o + i);
#pragma unroll
                    for (int j = 0; j < 4; ++j) xa[j] = ld4(aw.p, kl + 16 * j);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dWa[i][j] = fmaf(ga[i].x, xa[j].x, fmaf(ga[i].y, xa[j].y, fmaf(ga[i].z, xa[j].z, fmaf(ga[i].w, xa[j].w, dWa[i][j]))));
                }
                {
                    float4 gb[4], xb[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) gb[i] = ld4(gw.gzb, tb_o + i);
#pragma unroll
                    for (int j = 0; j < 8; ++j) xb[j] = ld4(aw.ha, kl + 16 * j);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            dWb[i][j] = fmaf(gb[i].x, xb[j].x, fmaf(gb[i].y, xb[j].y, fmaf(gb[i].z, xb[j].z, fmaf(gb[i].w, xb[j].w, dWb[i][j]))));
                }
                {
                    float4 gc[2], xc[4];
#pragma unroll
                    for (int i = 0; i < 2; ++i) gc[i] = ld4(aw.w, tc_o + i);
#pragma unroll
                    for (int j = 0; j < 4; ++j) xc[j] = ld4(aw.hb, kl + 16 * j);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dWc[i][j] = fmaf(gc[i].x, xc[j].x, fmaf(gc[i].y, xc[j].y, fmaf(gc[i].z, xc[j].z, fmaf(gc[i].w, xc[j].w, dWc[i][j]))));
                }
                // biases: thread t < 128 -> ba[t]; 128..191 -> bb; 192..221 -> bc
                float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tid < H1) bsum = ld4(gw.gza, tid);
                else if (tid < H1 + H2) bsum = ld4(gw.gzb, tid - H1);
                else if (tid < H1 + H2 + KN) bsum = ld4(aw.w, tid - H1 - H2);
                dbias += (bsum.x + bsum.y) + (bsum.z + bsum.w);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (kl + 16 * j < IN0) atomicAdd(g_params + P_WA + (ta_o + i) * IN0 + kl + 16 * j, dWa[i][j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(g_params + P_WB + (tb_o + i) * H1 + kl + 16 * j, dWb[i][j]);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (tc_o + i < KN) atomicAdd(g_params + P_WC + (tc_o + i) * H2 + kl + 16 * j, dWc[i][j]);
    if (tid < H1) atomicAdd(g_params + P_BA + tid, dbias);
    else if (tid < H1 + H2) atomicAdd(g_params + P_BB + tid - H1, dbias);
    else if (tid < H1 + H2 + KN) atomicAdd(g_params + P_BC + tid - H1 - H2, dbias);
}

constexpr size_t ITP_FWD_SMEM = sizeof(ItpSmem) + sizeof(ItpWarp) * IW;
constexpr size_t ITP_BWD_SMEM = ITP_FWD_SMEM + sizeof(ItpWarpGrad) * IW;

}  // namespace mmpde

using namespace mmpde;

extern "C" int mmpde_itp_fwd(const float* src_xy, const float* src_val, const float* qry_xy, const int32_t* idx,
                             int64_t n_queries, const float* params, float* out, void* stream) {
    if (n_queries < 0) return MMPDE_EINVAL;
    if (n_queries == 0) return MMPDE_OK;
    MMPDE_ENSURE_SMEM(itp_fwd_kernel, ITP_FWD_SMEM);
    int grid = (int)imin64((n_queries + IW * QW - 1) / (IW * QW), sm_count());
    itp_fwd_kernel<<<grid, IW * 32, ITP_FWD_SMEM, (cudaStream_t)stream>>>((const float2*)src_xy, src_val, (const float2*)qry_xy, idx, n_queries, params, out);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_itp_bwd(const float* src_xy, const float* src_val, const float* qry_xy, const int32_t* idx,
                             int64_t n_queries, const float* params, const float* g_out, float* g_params,
                             float* g_src_val, void* stream) {
    if (n_queries < 0) return MMPDE_EINVAL;
    if (n_queries == 0) return MMPDE_OK;
    MMPDE_ENSURE_SMEM(itp_bwd_kernel, ITP_BWD_SMEM);
    int grid = (int)imin64((n_queries + IW * QW - 1) / (IW * QW), sm_count());
    itp_bwd_kernel<<<grid, IW * 32, ITP_BWD_SMEM, (cudaStream_t)stream>>>((const float2*)src_xy, src_val, (const float2*)qry_xy, idx, n_queries, params, g_out, g_params, g_src_val);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
