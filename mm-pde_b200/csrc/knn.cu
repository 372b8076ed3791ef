// Exact ordered k-nearest-neighbour search (graph construction + interpolation neighbour lists).
// Replaces torch_cluster.knn_graph (/root/reference/data_creator_2d.py:260) and
// sklearn NearestNeighbors.kneighbors (/root/reference/data_creator_2d.py:66,75-76).
// Integer outputs must equal oracle/knn_oracle.c bit for bit: ascending (d2, original index).
#include "common.cuh"

namespace mmpde {

template <typename D> struct Dist;
template <> struct Dist<float> {   // rule 0: torch_cluster CUDA kernel arithmetic (fp32, FMA on the 2nd term)
    static __device__ __forceinline__ float d2(float qx, float qy, float px, float py) {
        float dx = qx - px, dy = qy - py;
        return __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
    }
};
template <> struct Dist<double> {  // rule 1: sklearn kd-tree arithmetic (fp64, no contraction)
    static __device__ __forceinline__ double d2(float qx, float qy, float px, float py) {
        double dx = __dsub_rn((double)qx, (double)px), dy = __dsub_rn((double)qy, (double)py);
        return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    }
};

// Sorted list of the K best (d, idx), ascending lexicographically.  Lives in local memory (L1-resident).
template <typename D, int KMAX>
struct TopK {
    D d[KMAX];
    int i[KMAX];
    int cnt, k;
    __device__ __forceinline__ void init(int k_) { cnt = 0; k = k_; }
    __device__ __forceinline__ bool before(D a, int ai, D b, int bi) const { return a < b || (a == b && ai < bi); }
    __device__ __forceinline__ void push(D dv, int iv) {
        if (cnt == k && !before(dv, iv, d[k - 1], i[k - 1])) return;
        int pos = (cnt < k) ? cnt : k - 1;
        while (pos > 0 && before(dv, iv, d[pos - 1], i[pos - 1])) { d[pos] = d[pos - 1]; i[pos] = i[pos - 1]; --pos; }
        d[pos] = dv; i[pos] = iv;
        if (cnt < k) ++cnt;
    }
};

// The same list in SHARED memory, entry j of thread t at [j * 128 + t] (conflict-free).  The cell-binned search keeps its
// lists here: in local memory they are ~400 B per thread, and with several searches resident per SM they fall out of
// L1 (three grouped searches ran barely faster than one after the other).
template <typename D>
struct TopKShared {
    D* d;
    int* i;
    int cnt, k;
    static constexpr int S = 128;                 // = blockDim.x of knn_grid_kernel
    __device__ __forceinline__ void init(int k_, void* d_base, int* i_base) {
        cnt = 0; k = k_;
        d = reinterpret_cast<D*>(d_base) + threadIdx.x;
        i = i_base + threadIdx.x;
    }
    __device__ __forceinline__ bool before(D a, int ai, D b, int bi) const { return a < b || (a == b && ai < bi); }
    __device__ __forceinline__ D worst() const { return d[(k - 1) * S]; }
    __device__ __forceinline__ void push(D dv, int iv) {
        if (cnt == k && !before(dv, iv, d[(k - 1) * S], i[(k - 1) * S])) return;
        int pos = (cnt < k) ? cnt : k - 1;
        while (pos > 0 && before(dv, iv, d[(pos - 1) * S], i[(pos - 1) * S])) {
            d[pos * S] = d[(pos - 1) * S]; i[pos * S] = i[(pos - 1) * S]; --pos;
        }
        d[pos * S] = dv; i[pos * S] = iv;
        if (cnt < k) ++cnt;
    }
};

constexpr int KNN_KMAX = 64;
constexpr int KNN_TILE = 512;

// One thread per query, the sample's points streamed through shared memory in index order.
template <typename D>
__global__ void __launch_bounds__(128) knn_brute_kernel(const float2* __restrict__ pts, const int* __restrict__ pts_off,
                                                        const float2* __restrict__ qry, const int* __restrict__ qry_off,
                                                        int n_samples, int k, int exclude_self, int* __restrict__ out) {
    __shared__ float2 tile[KNN_TILE];
    // blockIdx.y = sample, blockIdx.x = query block inside the sample
    int s = blockIdx.y;
    int q0 = qry_off[s], q1 = qry_off[s + 1], p0 = pts_off[s], p1 = pts_off[s + 1];
    for (int qb = q0 + blockIdx.x * blockDim.x; qb < q1; qb += gridDim.x * blockDim.x) {
        int q = qb + threadIdx.x;
        bool live = q < q1;
        float2 qq = live ? qry[q] : make_float2(0.f, 0.f);
        int self = (exclude_self && live) ? p0 + (q - q0) : -1;
        TopK<D, KNN_KMAX> best;
        best.init(k);
        for (int t0 = p0; t0 < p1; t0 += KNN_TILE) {
            int nt = min(KNN_TILE, p1 - t0);
            __syncthreads();
            for (int j = threadIdx.x; j < nt; j += blockDim.x) tile[j] = pts[t0 + j];
            __syncthreads();
            if (live) {
                for (int j = 0; j < nt; ++j) {
                    int p = t0 + j;
                    if (p == self) continue;
                    float2 pp = tile[j];
                    best.push(Dist<D>::d2(qq.x, qq.y, pp.x, pp.y), p);
                }
            }
        }
        if (live)
            for (int j = 0; j < k; ++j) out[(int64_t)q * k + j] = (j < best.cnt) ? best.i[j] : -1;
    }
}

// ---- cell-binned search for one large sample ----------------------------------------------------
__device__ __forceinline__ int cell_coord(float v, float v0, float inv_cell, int g) {
    int c = (int)floorf((v - v0) * inv_cell);
    return min(max(c, 0), g - 1);
}

// sample s with off[s] <= i < off[s+1]
__device__ __forceinline__ int sample_of(const int* __restrict__ off, int n_samples, int64_t i) {
    int lo = 0, hi = n_samples;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(off + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void grid_count_kernel(const float2* __restrict__ pts, int64_t n, const int* __restrict__ pts_off, int n_samples,
                                  float x0, float y0, float inv_cell, int gx, int gy, int* __restrict__ cell_of_pt,
                                  int* __restrict__ counts) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float2 p = pts[i];
        int c = sample_of(pts_off, n_samples, i) * gx * gy + cell_coord(p.y, y0, inv_cell, gy) * gx + cell_coord(p.x, x0, inv_cell, gx);
        cell_of_pt[i] = c;
        atomicAdd(&counts[c + 1], 1);
    }
}

// single-block exclusive scan over counts[1..ncell] in place (ncell up to a few million: chunked)
__global__ void grid_scan_kernel(int* __restrict__ cs, int ncell) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 1; base <= ncell; base += blockDim.x) {
        int idx = base + threadIdx.x;
        int v = (idx <= ncell) ? cs[idx] : 0;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[w] = x;
        __syncthreads();
        if (w == 0) {
            int t = (lane < (blockDim.x >> 5)) ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
            warp_tot[lane] = t;
        }
        __syncthreads();
        int incl = x + (w > 0 ? warp_tot[w - 1] : 0) + carry;
        if (idx <= ncell) cs[idx] = incl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = incl;
        __syncthreads();
    }
}

__global__ void grid_fill_kernel(const int* __restrict__ cell_of_pt, int64_t n, const int* __restrict__ cell_start,
                                 int* __restrict__ cursor, int* __restrict__ order) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = cell_of_pt[i];
        int slot = cell_start[c] + atomicAdd(&cursor[c], 1);
        order[slot] = (int)i;   // order inside a cell is arbitrary; the (d2, idx) comparator makes the result unique
    }
}

// One search = one task; a launch carries up to KNN_MAX_TASKS of them, the CTAs divided between the tasks.  A search
// is latency-bound with one thread per query (10 % of the warp slots at 36 k queries), so the three searches of a
// training step (graph on the moved mesh, interpolation to it and back) run side by side in one launch.
constexpr int KNN_MAX_TASKS = 4;
#ifndef MMPDE_KNN_BATCH
#define MMPDE_KNN_BATCH 4
#endif
constexpr int KNN_BATCH = MMPDE_KNN_BATCH;              // candidates whose loads are in flight together
struct KnnGroup {
    mmpde_knn_task t[KNN_MAX_TASKS];
    int cta_begin[KNN_MAX_TASKS + 1];
    int kmax;                                     // largest k of the group: the shared-memory lists are sized for it
};

template <typename D>
__device__ __forceinline__ void knn_grid_body(const mmpde_knn_task& a, int64_t bx, int64_t nb, void* d_base, int* i_base) {
    const float2* __restrict__ pts = reinterpret_cast<const float2*>(a.pts);
    const float2* __restrict__ qry = reinterpret_cast<const float2*>(a.qry);
    const int* __restrict__ cell_start = a.cell_start;
    const int* __restrict__ order = a.order;
    const int gx = a.gx, gy = a.gy, k = a.k;
    const float x0 = a.x0, y0 = a.y0, inv_cell = a.inv_cell;
    const float cell = 1.0f / inv_cell;
    for (int64_t q = bx * (int64_t)blockDim.x + threadIdx.x; q < a.n_queries; q += nb * blockDim.x) {
        float2 qq = qry[q];
        int cx = cell_coord(qq.x, x0, inv_cell, gx), cy = cell_coord(qq.y, y0, inv_cell, gy);
        const int smp = sample_of(a.qry_off, a.n_samples, q);
        const int cell0 = smp * gx * gy;
        int self = a.exclude_self ? (int)(__ldg(a.pts_off + smp) + (q - __ldg(a.qry_off + smp))) : -1;
        TopKShared<D> best;
        best.init(k, d_base, i_base);
        int rmax = max(max(cx, gx - 1 - cx), max(cy, gy - 1 - cy));
        for (int r = 0; r <= rmax; ++r) {
            if (best.cnt == k && r > 0) {
                // every unvisited point lies outside the (2r-1)^2 block of cells around (cx,cy):
                // its distance is at least the gap from the query to that block's border.
                float lo_x = x0 + (cx - (r - 1)) * cell, hi_x = x0 + (cx + r) * cell;
                float lo_y = y0 + (cy - (r - 1)) * cell, hi_y = y0 + (cy + r) * cell;
                // border cells are clamped (they extend to infinity): sides on the domain edge impose no bound
                float gap = 3.0e38f;
                if (cx - (r - 1) > 0) gap = fminf(gap, qq.x - lo_x);
                if (cx + r < gx) gap = fminf(gap, hi_x - qq.x);
                if (cy - (r - 1) > 0) gap = fminf(gap, qq.y - lo_y);
                if (cy + r < gy) gap = fminf(gap, hi_y - qq.y);
                gap = gap * (1.0f - 1e-5f) - 2e-6f * (gx + gy) * cell;   // conservative (cell-assignment rounding): never stop early
                if (gap > 0.f && (double)best.worst() < (double)gap * (double)gap) break;
            }
            // the ring of cells at Chebyshev distance r: whole rows of cells at its top and bottom, the two end cells of the
            // rows in between.  Cells of a row are neighbours in `order`, so a row piece is ONE contiguous candidate range,
            // walked KNN_BATCH candidates at a time: the index loads of a group go out together, then the coordinate gathers,
            // then the insertions (one point at a time the kernel sat on two dependent L2 round trips per candidate:
            // long_scoreboard was half of all stall cycles, profiles/r01_ncu_knn_summary.txt)
            auto scan = [&](int c_lo, int c_hi) {                           // cells c_lo..c_hi of the sample (same row)
                const int s1 = __ldg(cell_start + c_hi + 1);
                for (int s = __ldg(cell_start + c_lo); s < s1; s += KNN_BATCH) {
                    int p[KNN_BATCH];
                    float2 pp[KNN_BATCH];
#pragma unroll
                    for (int j = 0; j < KNN_BATCH; ++j) p[j] = (s + j < s1) ? __ldg(order + s + j) : -1;
#pragma unroll
                    for (int j = 0; j < KNN_BATCH; ++j) pp[j] = (p[j] >= 0) ? __ldg(pts + p[j]) : make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < KNN_BATCH; ++j)
                        if (p[j] >= 0 && p[j] != self) best.push(Dist<D>::d2(qq.x, qq.y, pp[j].x, pp[j].y), p[j]);
                }
            };
            const int ylo = cy - r, yhi = cy + r, xlo = cx - r, xhi = cx + r;
            const int xa = max(xlo, 0), xb = min(xhi, gx - 1);
            for (int yy = max(ylo, 0); yy <= min(yhi, gy - 1); ++yy) {
                const int row = cell0 + yy * gx;
                if (yy == ylo || yy == yhi) {
                    scan(row + xa, row + xb);
                } else {
                    if (xlo >= 0) scan(row + xlo, row + xlo);
                    if (xhi < gx) scan(row + xhi, row + xhi);
                }
            }
        }
        for (int j = 0; j < k; ++j) a.out_idx[q * k + j] = (j < best.cnt) ? best.i[j * TopKShared<D>::S] : -1;
    }
}

__global__ void __launch_bounds__(128) knn_grid_kernel(const __grid_constant__ KnnGroup g) {
    int task = 0;
#pragma unroll
    for (int j = 1; j < KNN_MAX_TASKS; ++j) task += ((int)blockIdx.x >= g.cta_begin[j]) ? 1 : 0;
    const mmpde_knn_task& a = g.t[task];
    const int64_t bx = (int64_t)blockIdx.x - g.cta_begin[task], nb = g.cta_begin[task + 1] - g.cta_begin[task];
    extern __shared__ __align__(16) unsigned char knn_smem[];      // [kmax][128] distances (8 B slots), then [kmax][128] indices
    int* i_base = reinterpret_cast<int*>(knn_smem + (size_t)g.kmax * 128 * sizeof(double));
    if (a.rule == 0) knn_grid_body<float>(a, bx, nb, knn_smem, i_base);
    else knn_grid_body<double>(a, bx, nb, knn_smem, i_base);
}

__global__ void radius_kernel(const float2* __restrict__ pts, const int* __restrict__ off, int n_samples, float r2,
                              int max_nb, int* __restrict__ out) {
    int s = blockIdx.y;
    int p0 = off[s], p1 = off[s + 1];
    for (int q = p0 + blockIdx.x * blockDim.x + threadIdx.x; q < p1; q += gridDim.x * blockDim.x) {
        float2 qq = pts[q];
        int cnt = 0;
        for (int p = p0; p < p1 && cnt < max_nb; ++p) {
            if (p == q) continue;
            float2 pp = pts[p];
            if (Dist<float>::d2(qq.x, qq.y, pp.x, pp.y) < r2) out[(int64_t)q * max_nb + cnt++] = p;
        }
        for (int j = cnt; j < max_nb; ++j) out[(int64_t)q * max_nb + j] = -1;
    }
}

}  // namespace mmpde

using namespace mmpde;

extern "C" int mmpde_knn(const float* pts, const int32_t* pts_off, const float* qry, const int32_t* qry_off,
                         int n_samples, int64_t n_queries, int k, int rule, int exclude_self, int32_t* out_idx,
                         void* stream) {
    if (k <= 0 || k > KNN_KMAX || n_samples < 0 || (rule != 0 && rule != 1)) return MMPDE_EINVAL;
    if (n_samples == 0 || n_queries == 0) return MMPDE_OK;
    int64_t per = (n_queries + n_samples - 1) / n_samples;
    dim3 grid((unsigned)((per + 127) / 128), (unsigned)n_samples);
    auto st = (cudaStream_t)stream;
    if (rule == 0)
        knn_brute_kernel<float><<<grid, 128, 0, st>>>((const float2*)pts, pts_off, (const float2*)qry, qry_off, n_samples, k, exclude_self, out_idx);
    else
        knn_brute_kernel<double><<<grid, 128, 0, st>>>((const float2*)pts, pts_off, (const float2*)qry, qry_off, n_samples, k, exclude_self, out_idx);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_knn_grid_build(const float* pts, const int32_t* pts_off, int n_samples, int64_t n_pts, float x0,
                                    float y0, float inv_cell, int gx, int gy, int32_t* cell_of_pt, int32_t* cell_start,
                                    int32_t* cursor, int32_t* order, void* stream) {
    if (n_pts < 0 || n_samples <= 0 || gx <= 0 || gy <= 0 || (int64_t)gx * gy * n_samples > (1 << 28)) return MMPDE_EINVAL;
    auto st = (cudaStream_t)stream;
    int ncell = gx * gy * n_samples;
    cudaMemsetAsync(cell_start, 0, sizeof(int) * (size_t)(ncell + 1), st);
    cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)ncell, st);
    if (n_pts == 0) return MMPDE_OK;
    int blocks = (int)imin64((n_pts + 255) / 256, (int64_t)sm_count() * 16);
    grid_count_kernel<<<blocks, 256, 0, st>>>((const float2*)pts, n_pts, pts_off, n_samples, x0, y0, inv_cell, gx, gy, cell_of_pt, cell_start);
    grid_scan_kernel<<<1, 1024, 0, st>>>(cell_start, ncell);
    grid_fill_kernel<<<blocks, 256, 0, st>>>(cell_of_pt, n_pts, cell_start, cursor, order);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}

extern "C" int mmpde_knn_grid_multi(const mmpde_knn_task* tasks, int n_tasks, void* stream) {
    if (n_tasks < 0 || (n_tasks > 0 && tasks == nullptr)) return MMPDE_EINVAL;
    for (int j = 0; j < n_tasks; ++j) {
        const mmpde_knn_task& t = tasks[j];
        if (t.k <= 0 || t.k > KNN_KMAX || (t.rule != 0 && t.rule != 1) || t.gx <= 0 || t.gy <= 0 || t.n_samples <= 0 || t.n_queries < 0)
            return MMPDE_EINVAL;
    }
    auto st = (cudaStream_t)stream;
    int j0 = 0;
    while (j0 < n_tasks) {
        KnnGroup g;
        int n = 0;
        int64_t begin = 0;
        g.cta_begin[0] = 0;
        g.kmax = 1;
        for (; j0 < n_tasks && n < KNN_MAX_TASKS; ++j0) {
            if (tasks[j0].n_queries == 0) continue;
            g.t[n] = tasks[j0];
            if (tasks[j0].k > g.kmax) g.kmax = tasks[j0].k;
            begin += imin64((tasks[j0].n_queries + 127) / 128, (int64_t)sm_count() * 32);
            g.cta_begin[++n] = (int)begin;
        }
        if (n == 0) break;
        for (int j = n; j < KNN_MAX_TASKS; ++j) { g.t[j] = g.t[0]; g.cta_begin[j + 1] = 0x7fffffff; }
        const size_t smem = (size_t)g.kmax * 128 * (sizeof(double) + sizeof(int));     // <= 96 KB at k = 64
        if (smem > 48 * 1024) MMPDE_ENSURE_SMEM(knn_grid_kernel, 64 * 128 * 12);
        knn_grid_kernel<<<(int)begin, 128, smem, st>>>(g);
        MMPDE_CHECK_LAUNCH();
    }
    return MMPDE_OK;
}

extern "C" int mmpde_knn_grid(const float* pts, const int32_t* pts_off, const float* qry, const int32_t* qry_off,
                              int n_samples, int64_t n_queries, float x0, float y0, float inv_cell, int gx, int gy,
                              const int32_t* cell_start, const int32_t* order, int k, int rule, int exclude_self,
                              int32_t* out_idx, void* stream) {
    mmpde_knn_task t;
    t.pts = pts; t.pts_off = pts_off; t.qry = qry; t.qry_off = qry_off; t.n_samples = n_samples; t.k = k;
    t.n_queries = n_queries; t.x0 = x0; t.y0 = y0; t.inv_cell = inv_cell; t.gx = gx; t.gy = gy;
    t.cell_start = cell_start; t.order = order; t.rule = rule; t.exclude_self = exclude_self; t.out_idx = out_idx;
    return mmpde_knn_grid_multi(&t, 1, stream);
}

extern "C" int mmpde_radius(const float* pts, const int32_t* off, int n_samples, int64_t n_pts, float r, int max_nb,
                            int32_t* out_idx, void* stream) {
    if (max_nb <= 0 || n_samples < 0) return MMPDE_EINVAL;
    if (n_samples == 0 || n_pts == 0) return MMPDE_OK;
    int64_t per = (n_pts + n_samples - 1) / n_samples;
    dim3 grid((unsigned)((per + 127) / 128), (unsigned)n_samples);
    radius_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const float2*)pts, off, n_samples, r * r, max_nb, out_idx);
    MMPDE_CHECK_LAUNCH();
    return MMPDE_OK;
}
