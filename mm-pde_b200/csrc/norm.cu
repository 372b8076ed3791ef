// BatchNorm over all nodes of the batch + the elementwise helpers of the node path (all HBM-bound).
// Replaces PyG BatchNorm / nn.BatchNorm1d (/root/reference/gnn_2d.py:51,56,101,104) and autograd's ReLU masks.
// Row-major [M,128] fp32; every thread moves one float4 (4 channels), a warp one 512-byte row.
#include "peer.cuh"
#include <cstdlib>

namespace mmpde {

constexpr int ROWS_PER_WARP = 8;                 // rows whose loads one warp keeps in flight together
constexpr int ROWS_PER_CTA = 8 * ROWS_PER_WARP;  // rows reduced by one CTA before it touches the fp64 accumulators
constexpr int APPLY_ROWS = 4;                    // rows per warp of the elementwise kernels (CTA = 32 rows)

__device__ __forceinline__ float4 load_y(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t r, int c4) {
    float4 y = ldg4(A + r * lda + c4 * 4);
    if (B) {
        float4 b = ldg4(B + r * ldb + c4 * 4);
        y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
    }
    return y;
}

// Column sums are accumulated in fp64 FROM THE FIRST ADD (per thread, then per CTA, then atomically): the value of a sum of
// a few 10^4 fp32 terms then carries an error of ~1e-19 of its scale, far below half an ulp of the fp32 mean / rstd derived
// from it -- so the statistics, and with them every activation and every ReLU mask of the forward, are the SAME BITS however
// the rows are grouped: whole graph or partitioned, one rank or sharded over eight.  (With fp32 partial sums per CTA the
// grouping moved mean / rstd by an ulp, and the few ReLU masks that flipped showed up as 1e-3 of gradient difference.)
// block-level reduction of per-thread partials (8 row groups x 32 lanes x 4 channels) into fp64 atomics on dst[0..127]
__device__ __forceinline__ void block_reduce_to_double(const double (&v)[4], double* dst) {
    __shared__ double red[8][32][4];
    const int c4 = threadIdx.x & 31, rg = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[rg][c4][j] = v[j];
    __syncthreads();
    if (rg == 0) {
        double s[4] = {0, 0, 0, 0};
#pragma unroll
        for (int g = 0; g < 8; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] += red[g][c4][j];
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dst + c4 * 4 + j, s[j]);
    }
}

// What the LAST CTA of a column-reducing kernel does once the sums are complete (ticket == nullptr: nothing -- the caller
// folds / exchanges / finalises with separate launches).  Saves a launch for the fold (+ one for the cross-GPU exchange
// + one for the finalisation) per BatchNorm pass: a step runs 32 passes, each on its critical path.
struct BnTail {
    unsigned int* ticket;                  // zeroed with the sums; counts the CTAs that have delivered their partials
    const int64_t* peer_base;              // cross-rank exchange buffers (mmpde_bn_exchange) or nullptr = this rank only
    int rank, world;
    int post_only;                         // 1: deliver this rank's sums to the peers and return (mmpde_bn_exchange_wait finishes)
    unsigned long long timeout_ns;
    double* local_out;                     // [256] this rank's folded sums (backward: dbeta | dgamma), or nullptr
    double* glob_out;                      // [256] sums over all ranks, or nullptr
    double count;                          // forward: rows of the whole batch; <= 0: no finalisation
    float eps, momentum;
    float* mean_rstd; float* rmean; float* rvar;
};

__device__ __forceinline__ void bn_tail(const BnTail& t, const double* __restrict__ sums_all) {
    if (t.ticket == nullptr) return;
    __shared__ bool s_last;
    __shared__ double s_fold[256];
    __threadfence();                       // this CTA's atomics before its ticket
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(t.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int c = threadIdx.x;
    double v = 0.0;
#pragma unroll
    for (int r = 0; r < MMPDE_BN_REPLICAS; ++r) v += __ldcg(sums_all + r * 256 + c);
    if (t.local_out) t.local_out[c] = v;
    if (t.peer_base != nullptr && t.world > 1) {
        if (t.post_only) { peer_post_256(v, t.peer_base, t.rank, t.world); return; }
        v = peer_exchange_256(v, t.peer_base, t.rank, t.world, t.timeout_ns);
    }
    if (t.glob_out) t.glob_out[c] = v;
    if (t.count > 0) {
        s_fold[c] = v;
        __syncthreads();
        if (c < 128) {
            const double mean = s_fold[c] / t.count;
            double var = s_fold[128 + c] / t.count - mean * mean;
            if (var < 0) var = 0;
            t.mean_rstd[c] = (float)mean;
            t.mean_rstd[128 + c] = (float)(1.0 / sqrt(var + (double)t.eps));
            if (t.rmean) {
                const double unbiased = t.count > 1 ? var * t.count / (t.count - 1) : var;
                t.rmean[c] = (float)((1.0 - t.momentum) * t.rmean[c] + t.momentum * mean);
                t.rvar[c] = (float)((1.0 - t.momentum) * t.rvar[c] + t.momentum * unbiased);
            }
        }
    }
}

__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                       int64_t ldb, int64_t M, double* __restrict__ sums,
                                                       const __grid_constant__ BnTail tail) {
    const int c4 = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const double* sums_all = sums;
    // fp32 partials over the CTA's row blocks (a thread sees M / (8 * grid) rows), ONE fp64 reduction per CTA into
    // replica blockIdx % MMPDE_BN_REPLICAS: same-address atomics serialise in L2 (~30-50 ns each when every CTA
    // arrives at once), so the time of the tail is the number of CTAs per replica (profiles/r01_norm_bench.txt)
    sums += (blockIdx.x % MMPDE_BN_REPLICAS) * 256;
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    for (int64_t base = (int64_t)blockIdx.x * ROWS_PER_CTA; base < M; base += (int64_t)gridDim.x * ROWS_PER_CTA) {
        auto acc = [&](float4 y) {
            const double v[4] = {(double)y.x, (double)y.y, (double)y.z, (double)y.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[j] += v[j]; q[j] = fma(v[j], v[j], q[j]); }      // y*y is exact in fp64
        };
        if (base + ROWS_PER_CTA <= M) {                 // full block: all loads of the warp's rows in flight at once
            float4 y[ROWS_PER_WARP];
#pragma unroll
            for (int k = 0; k < ROWS_PER_WARP; ++k) y[k] = load_y(A, lda, B, ldb, base + rg + 8 * k, c4);
#pragma unroll
            for (int k = 0; k < ROWS_PER_WARP; ++k) acc(y[k]);
        } else {
            for (int64_t r = base + rg; r < M; r += 8) acc(load_y(A, lda, B, ldb, r, c4));
        }
    }
    block_reduce_to_double(s, sums);
    block_reduce_to_double(q, sums + 128);
    bn_tail(tail, sums_all);
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, int n_rep, double count, float eps, float momentum,
                                   float* __restrict__ mean_rstd, float* __restrict__ rmean, float* __restrict__ rvar) {
    int c = threadIdx.x;
    if (c >= 128) return;
    double s = 0.0, q = 0.0;
    for (int r = 0; r < n_rep; ++r) { s += sums[r * 256 + c]; q += sums[r * 256 + 128 + c]; }
    double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0) var = 0;
    mean_rstd[c] = (float)mean;
    mean_rstd[128 + c] = (float)(1.0 / sqrt(var + (double)eps));
    if (rmean) {
        double unbiased = count > 1 ? var * count / (count - 1) : var;
        rmean[c] = (float)((1.0 - momentum) * rmean[c] + momentum * mean);
        rvar[c] = (float)((1.0 - momentum) * rvar[c] + momentum * unbiased);
    }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                       int64_t ldb, int64_t M, const float* __restrict__ mean_rstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int relu, float* __restrict__ out, int64_t ldo) {
    const int c4 = threadIdx.x & 31;
    float4 mu = ldg4(mean_rstd + c4 * 4), rs = ldg4(mean_rstd + 128 + c4 * 4);
    float4 ga = ldg4(gamma + c4 * 4), be = ldg4(beta + c4 * 4);
    float4 sc = make_float4(ga.x * rs.x, ga.y * rs.y, ga.z * rs.z, ga.w * rs.w);
    const int rg = threadIdx.x >> 5;
    auto apply = [&](float4 y, int64_t r) {
        float4 o = make_float4(fmaf(y.x - mu.x, sc.x, be.x), fmaf(y.y - mu.y, sc.y, be.y),
                               fmaf(y.z - mu.z, sc.z, be.z), fmaf(y.w - mu.w, sc.w, be.w));
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(out + r * ldo + c4 * 4) = o;
    };
    for (int64_t base = (int64_t)blockIdx.x * (8 * APPLY_ROWS); base < M; base += (int64_t)gridDim.x * (8 * APPLY_ROWS)) {
        if (base + 8 * APPLY_ROWS <= M) {
            float4 y[APPLY_ROWS];
#pragma unroll
            for (int k = 0; k < APPLY_ROWS; ++k) y[k] = load_y(A, lda, B, ldb, base + rg + 8 * k, c4);
#pragma unroll
            for (int k = 0; k < APPLY_ROWS; ++k) apply(y[k], base + rg + 8 * k);
        } else {
            for (int64_t r = base + rg; r < M; r += 8) apply(load_y(A, lda, B, ldb, r, c4), r);
        }
    }
}

__device__ __forceinline__ float4 gated_grad(const float* g, int64_t ldg, const float* out, int64_t ldo, int relu,
                                             int64_t r, int c4) {
    float4 gv = ldg4(g + r * ldg + c4 * 4);
    if (relu) {
        float4 o = ldg4(out + r * ldo + c4 * 4);
        gv.x = o.x > 0.f ? gv.x : 0.f; gv.y = o.y > 0.f ? gv.y : 0.f;
        gv.z = o.z > 0.f ? gv.z : 0.f; gv.w = o.w > 0.f ? gv.w : 0.f;
    }
    return gv;
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ out,
                                                            int64_t ldo, int relu, const float* __restrict__ A, int64_t lda,
                                                            const float* __restrict__ B, int64_t ldb, int64_t M,
                                                            const float* __restrict__ mean_rstd, double* __restrict__ bsums,
                                                            const __grid_constant__ BnTail tail) {
    const int c4 = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const double* sums_all = bsums;
    float4 mu = ldg4(mean_rstd + c4 * 4), rs = ldg4(mean_rstd + 128 + c4 * 4);
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    bsums += (blockIdx.x % MMPDE_BN_REPLICAS) * 256;
    for (int64_t base = (int64_t)blockIdx.x * ROWS_PER_CTA; base < M; base += (int64_t)gridDim.x * ROWS_PER_CTA) {
        auto acc = [&](float4 gv, float4 y) {
            const double g4[4] = {(double)gv.x, (double)gv.y, (double)gv.z, (double)gv.w};
            const double xh[4] = {(double)((y.x - mu.x) * rs.x), (double)((y.y - mu.y) * rs.y), (double)((y.z - mu.z) * rs.z),
                                  (double)((y.w - mu.w) * rs.w)};                      // x-hat as the apply pass computes it (fp32)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[j] += g4[j]; q[j] = fma(g4[j], xh[j], q[j]); }
        };
        if (base + ROWS_PER_CTA <= M) {
#pragma unroll
            for (int h = 0; h < ROWS_PER_WARP; h += APPLY_ROWS) {
                float4 gv[APPLY_ROWS], y[APPLY_ROWS];
#pragma unroll
                for (int k = 0; k < APPLY_ROWS; ++k) {
                    gv[k] = gated_grad(g, ldg, out, ldo, relu, base + rg + 8 * (h + k), c4);
                    y[k] = load_y(A, lda, B, ldb, base + rg + 8 * (h + k), c4);
                }
#pragma unroll
                for (int k = 0; k < APPLY_ROWS; ++k) acc(gv[k], y[k]);
            }
        } else {
            for (int64_t r = base + rg; r < M; r += 8)
                acc(gated_grad(g, ldg, out, ldo, relu, r, c4), load_y(A, lda, B, ldb, r, c4));
        }
    }
    block_reduce_to_double(s, bsums);
    block_reduce_to_double(q, bsums + 128);
    bn_tail(tail, sums_all);
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ out,
                                                           int64_t ldo, int relu, const float* __restrict__ A, int64_t lda,
                                                           const float* __restrict__ B, int64_t ldb, int64_t M,
                                                           const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                                           const double* __restrict__ bsums, double count,
                                                           float* __restrict__ gy, int64_t ldgy, int accumulate,
                                                           float* __restrict__ gyg, int64_t ldgg) {
    const int c4 = threadIdx.x & 31;
    float4 mu = ldg4(mean_rstd + c4 * 4), rs = ldg4(mean_rstd + 128 + c4 * 4), ga = ldg4(gamma + c4 * 4);
    float mg[4], mgy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        mg[j] = (float)(bsums[c4 * 4 + j] / count);
        mgy[j] = (float)(bsums[128 + c4 * 4 + j] / count);
    }
    float4 sc = make_float4(ga.x * rs.x, ga.y * rs.y, ga.z * rs.z, ga.w * rs.w);
    const int rg = threadIdx.x >> 5;
    auto apply = [&](float4 gv, float4 y, float4 p, int64_t r) {
        float4 o;
        o.x = sc.x * (gv.x - mg[0] - (y.x - mu.x) * rs.x * mgy[0]) + p.x;
        o.y = sc.y * (gv.y - mg[1] - (y.y - mu.y) * rs.y * mgy[1]) + p.y;
        o.z = sc.z * (gv.z - mg[2] - (y.z - mu.z) * rs.z * mgy[2]) + p.z;
        o.w = sc.w * (gv.w - mg[3] - (y.w - mu.w) * rs.w * mgy[3]) + p.w;
        *reinterpret_cast<float4*>(gy + r * ldgy + c4 * 4) = o;
        if (gyg) {                                     // B is re-read from L1/L2: it was loaded for y a moment ago
            const float4 b = ldg4(B + r * ldb + c4 * 4);
            *reinterpret_cast<float4*>(gyg + r * ldgg + c4 * 4) =
                make_float4(b.x > 0.f ? o.x : 0.f, b.y > 0.f ? o.y : 0.f, b.z > 0.f ? o.z : 0.f, b.w > 0.f ? o.w : 0.f);
        }
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t base = (int64_t)blockIdx.x * (8 * APPLY_ROWS); base < M; base += (int64_t)gridDim.x * (8 * APPLY_ROWS)) {
        if (base + 8 * APPLY_ROWS <= M) {
            float4 gv[APPLY_ROWS], y[APPLY_ROWS], p[APPLY_ROWS];
#pragma unroll
            for (int k = 0; k < APPLY_ROWS; ++k) {
                const int64_t r = base + rg + 8 * k;
                gv[k] = gated_grad(g, ldg, out, ldo, relu, r, c4);
                y[k] = load_y(A, lda, B, ldb, r, c4);
                p[k] = accumulate ? *reinterpret_cast<const float4*>(gy + r * ldgy + c4 * 4) : zero;
            }
#pragma unroll
            for (int k = 0; k < APPLY_ROWS; ++k) apply(gv[k], y[k], p[k], base + rg + 8 * k);
        } else {
            for (int64_t r = base + rg; r < M; r += 8)
                apply(gated_grad(g, ldg, out, ldo, relu, r, c4), load_y(A, lda, B, ldb, r, c4),
                      accumulate ? *reinterpret_cast<const float4*>(gy + r * ldgy + c4 * 4) : zero, r);
        }
    }
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ act,
                                                       int64_t lda, int64_t M, float* __restrict__ out, int64_t ldo,
                                                       float* __restrict__ colsum) {
    __shared__ float4 red[8][32];
    const int c4 = threadIdx.x & 31, rg = threadIdx.x >> 5;
    float4 s = make_float4(0, 0, 0, 0);
    auto apply = [&](float4 gv, float4 a, int64_t r) {
        gv.x = a.x > 0.f ? gv.x : 0.f; gv.y = a.y > 0.f ? gv.y : 0.f;
        gv.z = a.z > 0.f ? gv.z : 0.f; gv.w = a.w > 0.f ? gv.w : 0.f;
        *reinterpret_cast<float4*>(out + r * ldo + c4 * 4) = gv;
        s.x += gv.x; s.y += gv.y; s.z += gv.z; s.w += gv.w;
    };
    for (int64_t base = (int64_t)blockIdx.x * ROWS_PER_CTA; base < M; base += (int64_t)gridDim.x * ROWS_PER_CTA) {
        if (base + ROWS_PER_CTA <= M) {
#pragma unroll
            for (int h = 0; h < ROWS_PER_WARP; h += APPLY_ROWS) {
                float4 gv[APPLY_ROWS], a[APPLY_ROWS];
#pragma unroll
                for (int k = 0; k < APPLY_ROWS; ++k) {
                    const int64_t r = base + rg + 8 * (h + k);
                    gv[k] = ldg4(g + r * ldg + c4 * 4);
                    a[k] = ldg4(act + r * lda + c4 * 4);
                }
#pragma unroll
                for (int k = 0; k < APPLY_ROWS; ++k) apply(gv[k], a[k], base + rg + 8 * (h + k));
            }
        } else {
            for (int64_t r = base + rg; r < M; r += 8) apply(ldg4(g + r * ldg + c4 * 4), ldg4(act + r * lda + c4 * 4), r);
        }
    }
    if (colsum) {
        red[rg][c4] = s;
        __syncthreads();
        if (rg == 0) {
            float4 t = red[0][c4];
#pragma unroll
            for (int k = 1; k < 8; ++k) { float4 u = red[k][c4]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
            atomicAdd(colsum + c4 * 4 + 0, t.x); atomicAdd(colsum + c4 * 4 + 1, t.y);
            atomicAdd(colsum + c4 * 4 + 2, t.z); atomicAdd(colsum + c4 * 4 + 3, t.w);
        }
    }
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int64_t lda, int64_t M, int N,
                                                     float* __restrict__ colsum) {
    // thread -> column (coalesced across the row), blockIdx.y strides over row chunks
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    float s = 0.f;
    for (int64_t r = blockIdx.y; r < M; r += gridDim.y) s += __ldg(A + r * lda + c);
    atomicAdd(colsum + c, s);
}

// grid of the column-reducing kernels: a few CTAs per SM, each looping over its row blocks
inline int reduce_grid(int64_t M) {
    static const int per_sm = [] { const char* e = getenv("MMPDE_REDUCE_CTAS_PER_SM"); return e ? atoi(e) : 2; }();
    int64_t g = (M + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
    return (int)imin64(g > 0 ? g : 1, (int64_t)sm_count() * per_sm);
}

inline int row_grid(int64_t M, int rows_per_cta) {
    int64_t g = (M + rows_per_cta - 1) / rows_per_cta;
    return (int)imin64(g > 0 ? g : 1, (int64_t)sm_count() * 8);
}

}  // namespace mmpde

using namespace mmpde;

extern "C" int mmpde_bn_stats(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, double* sums, void* stream) {
    if (M < 0 || lda % 4 || (B && ldb % 4)) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    bn_stats_kernel<<<reduce_grid(M), 256, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, M, sums, BnTail{});
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

static int check_peer(const int64_t* peer_base, int rank, int world) {
    if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world || (world > 1 && peer_base == nullptr)) return MMPDE_EINVAL;
    return MMPDE_OK;
}

extern "C" int mmpde_bn_stats_fused(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, double* sums,
                                    uint32_t* ticket, double count, float eps, float momentum, float* mean_rstd,
                                    float* running_mean, float* running_var, const int64_t* peer_base, int rank, int world,
                                    void* stream) {
    if (M <= 0 || lda % 4 || (B && ldb % 4) || !sums || !ticket || !mean_rstd || count <= 0) return MMPDE_EINVAL;
    if ((running_mean == nullptr) != (running_var == nullptr)) return MMPDE_EINVAL;
    if (int rc = check_peer(peer_base, rank, world)) return rc;
    BnTail t{};
    t.ticket = ticket; t.peer_base = world > 1 ? peer_base : nullptr; t.rank = rank; t.world = world;
    t.timeout_ns = peer_timeout_ns(); t.count = count; t.eps = eps; t.momentum = momentum;
    t.mean_rstd = mean_rstd; t.rmean = running_mean; t.rvar = running_var;
    bn_stats_kernel<<<reduce_grid(M), 256, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, M, sums, t);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_bn_finalize(const double* sums, int n_rep, double count, float eps, float momentum, float* mean_rstd,
                                 float* running_mean, float* running_var, void* stream) {
    if (count <= 0 || n_rep < 1 || ((running_mean == nullptr) != (running_var == nullptr))) return MMPDE_EINVAL;
    bn_finalize_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(sums, n_rep, count, eps, momentum, mean_rstd, running_mean, running_var);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_bn_apply(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, const float* mean_rstd,
                              const float* gamma, const float* beta, int relu, float* out, int64_t ldo, void* stream) {
    if (M < 0 || lda % 4 || (B && ldb % 4) || ldo % 4) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    bn_apply_kernel<<<row_grid(M, 8 * APPLY_ROWS), 256, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, M, mean_rstd, gamma, beta, relu, out, ldo);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_bn_bwd_reduce(const float* g, int64_t ldg, const float* out, int64_t ldo, int relu, const float* A,
                                   int64_t lda, const float* B, int64_t ldb, int64_t M, const float* mean_rstd,
                                   double* bsums, void* stream) {
    if (M < 0 || ldg % 4 || lda % 4 || (B && ldb % 4) || (relu && (!out || ldo % 4))) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    bn_bwd_reduce_kernel<<<reduce_grid(M), 256, 0, (cudaStream_t)stream>>>(g, ldg, out, ldo, relu, A, lda, B, ldb, M, mean_rstd, bsums, BnTail{});
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_bn_bwd_reduce_fused(const float* g, int64_t ldg, const float* out, int64_t ldo, int relu, const float* A,
                                         int64_t lda, const float* B, int64_t ldb, int64_t M, const float* mean_rstd,
                                         double* bsums, uint32_t* ticket, double* local_out, double* glob_out,
                                         const int64_t* peer_base, int rank, int world, void* stream) {
    if (M <= 0 || ldg % 4 || lda % 4 || (B && ldb % 4) || (relu && (!out || ldo % 4)) || !bsums || !ticket) return MMPDE_EINVAL;
    if (int rc = check_peer(peer_base, rank, world)) return rc;
    BnTail t{};
    t.ticket = ticket; t.peer_base = (world > 1 && glob_out) ? peer_base : nullptr; t.rank = rank; t.world = world;
    t.timeout_ns = peer_timeout_ns(); t.local_out = local_out; t.glob_out = glob_out; t.count = 0.0;
    bn_bwd_reduce_kernel<<<reduce_grid(M), 256, 0, (cudaStream_t)stream>>>(g, ldg, out, ldo, relu, A, lda, B, ldb, M, mean_rstd, bsums, t);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

// The same reduction, but the last CTA only POSTS this rank's sums to the peers (first half of the exchange): the caller
// queues independent work behind it -- the weight-gradient launch of the layer above -- and completes the exchange with
// mmpde_bn_exchange_wait right before mmpde_bn_bwd_apply, so neither the link latency nor a peer that is a few tens of
// microseconds behind stalls this rank's backward chain.
extern "C" int mmpde_bn_bwd_reduce_post(const float* g, int64_t ldg, const float* out, int64_t ldo, int relu, const float* A,
                                        int64_t lda, const float* B, int64_t ldb, int64_t M, const float* mean_rstd,
                                        double* bsums, uint32_t* ticket, double* local_out, const int64_t* peer_base, int rank,
                                        int world, void* stream) {
    if (M <= 0 || ldg % 4 || lda % 4 || (B && ldb % 4) || (relu && (!out || ldo % 4)) || !bsums || !ticket || !peer_base || world < 2)
        return MMPDE_EINVAL;
    if (int rc = check_peer(peer_base, rank, world)) return rc;
    BnTail t{};
    t.ticket = ticket; t.peer_base = peer_base; t.rank = rank; t.world = world; t.post_only = 1;
    t.timeout_ns = peer_timeout_ns(); t.local_out = local_out; t.glob_out = nullptr; t.count = 0.0;
    bn_bwd_reduce_kernel<<<reduce_grid(M), 256, 0, (cudaStream_t)stream>>>(g, ldg, out, ldo, relu, A, lda, B, ldb, M, mean_rstd, bsums, t);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_bn_bwd_apply(const float* g, int64_t ldg, const float* out, int64_t ldo, int relu, const float* A,
                                  int64_t lda, const float* B, int64_t ldb, int64_t M, const float* mean_rstd,
                                  const float* gamma, const double* bsums, double count, float* gy, int64_t ldgy,
                                  int accumulate, float* gy_gated, int64_t ldgg, void* stream) {
    if (M < 0 || count <= 0 || ldg % 4 || lda % 4 || (B && ldb % 4) || ldgy % 4 || (relu && (!out || ldo % 4))) return MMPDE_EINVAL;
    if (gy_gated && (!B || ldgg % 4)) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    bn_bwd_apply_kernel<<<row_grid(M, 8 * APPLY_ROWS), 256, 0, (cudaStream_t)stream>>>(g, ldg, out, ldo, relu, A, lda, B, ldb, M, mean_rstd, gamma, bsums, count, gy, ldgy, accumulate, gy_gated, ldgg);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_relu_bwd(const float* g, int64_t ldg, const float* act, int64_t lda, int64_t M, float* out,
                              int64_t ldo, float* colsum, void* stream) {
    if (M < 0 || ldg % 4 || lda % 4 || ldo % 4) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    relu_bwd_kernel<<<reduce_grid(M) * 2, 256, 0, (cudaStream_t)stream>>>(g, ldg, act, lda, M, out, ldo, colsum);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_colsum(const float* A, int64_t lda, int64_t M, int N, float* colsum, void* stream) {
    if (M < 0 || N <= 0) return MMPDE_EINVAL;
    if (M == 0) return MMPDE_OK;
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)imin64((M + 63) / 64, (int64_t)sm_count() * 4));
    colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, M, N, colsum);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
