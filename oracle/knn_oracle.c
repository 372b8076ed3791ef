/*
 * ORACLE -- test infrastructure only (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).
 * Never linked, imported or called from the product path under mm-pde_b200/.
 *
 * CPU restatement of the two k-nearest-neighbour searches on the MM-PDE hot path.
 * The arithmetic lives in un-vendored third-party code, so the rules are frozen here
 *  and anchored on the reference call sites:
 *
 *  (1) graph rule  -- torch_cluster 1.5.9 knn_graph CUDA kernel, called at
 *      /root/reference/data_creator_2d.py:260 and /root/reference/mesh/dmm_model.py:228.
 *      d2 = fmaf(dy,dy, dx*dx) in fp32, candidates scanned in index order, ascending d2,
 *      ties -> lower index; self excluded by index; neighbours never cross samples.
 *  (2) interpolation rule -- sklearn 1.3.0 NearestNeighbors(kd_tree) called at
 *      /root/reference/data_creator_2d.py:66,75-76.  fp32 coordinates widened to fp64,
 *      d2 = dx*dx + dy*dy in fp64 without contraction, ascending, ties -> lower index.
 *
 * Parity status: UNPINNED by the reference (it ships no tests/golden vectors); pinned
 * against sklearn / scipy.cKDTree as independent implementations in tests/test_oracle_knn.py.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* insert (d,i) into ascending list of length *cnt (capacity k); strict '<' keeps lower index first */
#define DEFINE_INSERT(NAME, T)                                                         \
static inline void NAME(T* bd, int64_t* bi, int* cnt, int k, T d, int64_t i) {         \
    int c = *cnt;                                                                      \
    if (c == k && !(d < bd[k - 1])) return;                                            \
    int pos = (c < k) ? c : k - 1;                                                     \
    while (pos > 0 && d < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; } \
    bd[pos] = d; bi[pos] = i;                                                          \
    if (c < k) *cnt = c + 1;                                                           \
}
DEFINE_INSERT(insert_f32, float)
DEFINE_INSERT(insert_f64, double)

/*
 * pts [P,2], qry [Q,2] fp32 row-major; pts_off/qry_off [S+1] sample offsets.
 * out_idx [Q,k] GLOBAL point indices (row of pts), padded with -1 when a sample has < k candidates.
 * exclude_self: skip the candidate whose within-sample index equals the query's within-sample index
 *               (graph construction, where qry == pts).
 * rule: 0 = fp32 fmaf (graph), 1 = fp64 (interpolation).
 */
int mmpde_oracle_knn(const float* pts, const float* qry,
                     const int64_t* pts_off, const int64_t* qry_off, int S,
                     int k, int exclude_self, int rule,
                     int64_t* out_idx, double* out_d2)
{
    if (k <= 0 || S < 0) return -1;
    float*   bd32 = (float*)malloc(sizeof(float) * (size_t)k);
    double*  bd64 = (double*)malloc(sizeof(double) * (size_t)k);
    int64_t* bi   = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    if (!bd32 || !bd64 || !bi) return -2;
    for (int s = 0; s < S; ++s) {
        int64_t p0 = pts_off[s], p1 = pts_off[s + 1];
        for (int64_t q = qry_off[s]; q < qry_off[s + 1]; ++q) {
            int cnt = 0;
            float qx = qry[2 * q], qy = qry[2 * q + 1];
            int64_t self = exclude_self ? (p0 + (q - qry_off[s])) : -1;
            for (int64_t p = p0; p < p1; ++p) {
                if (p == self) continue;
                if (rule == 0) {
                    float dx = qx - pts[2 * p], dy = qy - pts[2 * p + 1];
                    float d = fmaf(dy, dy, dx * dx);
                    insert_f32(bd32, bi, &cnt, k, d, p);
                } else {
                    double dx = (double)qx - (double)pts[2 * p];
                    double dy = (double)qy - (double)pts[2 * p + 1];
                    double d = dx * dx;
                    d = d + dy * dy;
                    insert_f64(bd64, bi, &cnt, k, d, p);
                }
            }
            for (int j = 0; j < k; ++j) {
                out_idx[q * k + j] = (j < cnt) ? bi[j] : -1;
                if (out_d2) out_d2[q * k + j] = (j < cnt) ? (rule == 0 ? (double)bd32[j] : bd64[j]) : INFINITY;
            }
        }
    }
    free(bd32); free(bd64); free(bi);
    return 0;
}

/*
 * radius_graph restatement (torch_cluster 1.5.9, call site /root/reference/data_creator_2d.py:258):
 * per query the FIRST max_nb candidates in index order with fp32 d2 < r*r (strict), self excluded.
 * out_idx [Q,max_nb] padded with -1; returns 0.
 */
int mmpde_oracle_radius(const float* pts, const int64_t* off, int S, float r, int max_nb,
                        int64_t* out_idx)
{
    float r2 = r * r;
    for (int s = 0; s < S; ++s) {
        for (int64_t q = off[s]; q < off[s + 1]; ++q) {
            int cnt = 0;
            float qx = pts[2 * q], qy = pts[2 * q + 1];
            for (int64_t p = off[s]; p < off[s + 1] && cnt < max_nb; ++p) {
                if (p == q) continue;
                float dx = qx - pts[2 * p], dy = qy - pts[2 * p + 1];
                float d = fmaf(dy, dy, dx * dx);
                if (d < r2) out_idx[q * max_nb + cnt++] = p;
            }
            for (int j = cnt; j < max_nb; ++j) out_idx[q * max_nb + j] = -1;
        }
    }
    return 0;
}
