/*
 * mmpde_b200.h -- C ABI of libmmpde_b200.so, the sm_100a kernels behind the MM-PDE hot path.
 *
 * The reference (Peiyannn/MM-PDE) has no FFI of its own: its boundary is the Python module surface
 * (SURVEY.md section 8b).  Each entry point below replaces one third-party operator call that the
 * reference's Python reaches (file:line relative to the reference tree), and is what the host-side
 * mirror in mm-pde_b200/*.py binds with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), never allocate or
 *     free, and keep no global state;
 *   - return 0 on success, a negative MMPDE_E* code for argument errors, a positive cudaError_t
 *     for launch failures;
 *   - fp32 row-major matrices with an explicit leading dimension (elements); int32 indices;
 *   - hidden width H = 128, time window 1, one "variables" column (gnn_2d.py:76-78,96): the node
 *     scalars travel as one float4 per node, node4 = (u, pos_x/Lx, pos_y/Ly, t/tmax).
 */
#ifndef MMPDE_B200_H
#define MMPDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMPDE_H 128
#define MMPDE_OK 0
#define MMPDE_EINVAL (-1)
#define MMPDE_EUNSUPPORTED (-2)

/* library / device facts: abi version, SM count the kernels were sized for. */
int mmpde_abi_version(void);
int mmpde_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Cap on the CTAs of the persistent one-CTA-per-SM kernels (mmpde_edge_fwd / _bwd, mmpde_node_gemm*, mmpde_node_wgrad*) on the
 * current device; 0 = one per SM (default).  Host-side state read at launch time (a recorded CUDA graph keeps the grids it was
 * recorded with).  Two independent solver passes issued on two streams run side by side at half width instead of taking
 * turns on the whole chip -- same throughput on one GPU, and on several GPUs the ranks stop drifting apart between the
 * cross-GPU BatchNorm exchanges. */
int mmpde_set_persistent_ctas(int n);

/* ---- graph construction --------------------------------------------------------------------
 * Replaces torch_cluster.knn_graph (data_creator_2d.py:260, mesh/dmm_model.py:228) and sklearn
 * NearestNeighbors.kneighbors (data_creator_2d.py:66,75-76).
 * pts [P,2], qry [Q,2] fp32; pts_off/qry_off [S+1] int32 sample offsets (neighbours never cross
 * samples).  out_idx [Q,k] int32 GLOBAL rows of pts, ascending (d2, index); -1 pads short samples.
 * rule 0: d2 = fmaf(dy,dy,dx*dx) fp32 (torch_cluster CUDA kernel); rule 1: fp64 dx*dx+dy*dy (sklearn).
 * exclude_self: skip the point whose within-sample index equals the query's (qry == pts). */
int mmpde_knn(const float* pts, const int32_t* pts_off, const float* qry, const int32_t* qry_off,
              int n_samples, int64_t n_queries, int k, int rule, int exclude_self,
              int32_t* out_idx, void* stream);

/* Same contract and bit-identical results through a uniform-cell binned exact search (the path used for
 * samples of more than a few hundred points; brute force costs O(P) per query, this O(k)).
 * All samples share one cell grid gx x gy over the box starting at (x0,y0) with cell size 1/inv_cell; points
 * outside the box are clamped into border cells (still exact).  cell_start [S*gx*gy+1], cursor [S*gx*gy],
 * cell_of_pt [P] and order [P] (points sorted by (sample, cell), original index) are caller workspace filled
 * by mmpde_knn_grid_build. */
int mmpde_knn_grid_build(const float* pts, const int32_t* pts_off, int n_samples, int64_t n_pts,
                         float x0, float y0, float inv_cell, int gx, int gy,
                         int32_t* cell_of_pt, int32_t* cell_start, int32_t* cursor, int32_t* order, void* stream);
int mmpde_knn_grid(const float* pts, const int32_t* pts_off, const float* qry, const int32_t* qry_off,
                   int n_samples, int64_t n_queries, float x0, float y0, float inv_cell, int gx, int gy,
                   const int32_t* cell_start, const int32_t* order,
                   int k, int rule, int exclude_self, int32_t* out_idx, void* stream);
/* Several such searches in ONE launch (they are latency-bound with one thread per query and leave most warp slots
 * empty, so the three searches of a training step -- graph on the moved mesh, interpolation to it and back -- run side
 * by side).  Every task has its own points / queries / cell grid (from mmpde_knn_grid_build) and output. */
typedef struct mmpde_knn_task {
    const float* pts; const int32_t* pts_off; const float* qry; const int32_t* qry_off;
    int32_t n_samples; int32_t k; int64_t n_queries;
    float x0, y0, inv_cell; int32_t gx, gy;
    const int32_t* cell_start; const int32_t* order;
    int32_t rule, exclude_self;
    int32_t* out_idx;
} mmpde_knn_task;
int mmpde_knn_grid_multi(const mmpde_knn_task* tasks, int n_tasks, void* stream);

/* radius_graph (data_creator_2d.py:258): first <= max_nb points in index order with d2 < r*r. */
int mmpde_radius(const float* pts, const int32_t* off, int n_samples, int64_t n_pts, float r,
                 int max_nb, int32_t* out_idx, void* stream);

/* ---- dense node-level contraction -------------------------------------------------------------
 * Replaces nn.Linear + ReLU on node tensors (gnn_2d.py:44-49,67-68,99-106) and their autograd.
 * C[M,N] (+)= act( opA(A)[M,K] * opB(B)[K,N] + bias[N] + r1_row[M]*r1_col[N] )
 *   a_kmajor = 1: A stored [M,K] (lda >= K);  0: A stored [K,M] (lda >= M)   (transposed operand)
 *   b_kmajor = 1: B stored [N,K] (ldb >= K);  0: B stored [K,N] (ldb >= N)
 *   bias, r1_row (stride r1_stride), r1_col may be NULL;  relu: 0/1;  accumulate: C += result;
 *   split_k > 1 splits the K loop over grid.z and adds atomically (C must be pre-zeroed or hold the
 *   value to accumulate onto; relu/bias then apply only when split_k == 1). */
int mmpde_gemm(const float* A, int64_t lda, int a_kmajor, const float* B, int64_t ldb, int b_kmajor,
               float* C, int64_t ldc, int64_t M, int N, int64_t K,
               const float* bias, const float* r1_row, int64_t r1_stride, const float* r1_col,
               int relu, int accumulate, int split_k, void* stream);

/* The same node-level contractions on the tcgen05 tensor cores (three split-bf16 products, fp32 accumulation in
 * TMEM: ~2^-16 relative error per term), for the shapes the processor uses: 128 outputs per call, K = 128 per
 * segment.  mmpde_gemm above stays for the odd shapes (K = 4, N = 1).
 *
 * mmpde_node_gemm:   C[m][n] = act( sum_k A0[m][k] W0(n,k) (+ sum_k A1[m][k] W1(n,k)) + node4[m] . Wext[n] + bias[n] )
 *                              + R1[m][n] + R2[m][n],      m < M, n < 128, k < 128
 *   A0 / A1 [M, >=128] fp32 row-major (ld multiple of 4, 16-byte aligned); A1 / W1 NULL for K = 128;
 *   W element (n,k) at W[n*w_ns + k*w_ks]  (nn.Linear weight [out,in]: w_ns = ld, w_ks = 1; its transpose for the
 *   data gradient: w_ns = 1, w_ks = ld);  Aext [M,4] / Wext [128,4] (ld 4) optional extra K columns (the node scalars);
 *   bias [128], R1, R2 (residuals, may alias C) optional;  relu = 1 applies before the residuals.
 *   relu = 2 is the ReLU BACKWARD gate fused into a dgrad: C = acc where R1 > 0, else 0 (R1 = the saved forward
 *   activation, R2 must be NULL).
 * mmpde_node_wgrad:  dW[i][j] += sum_m A[m][i] B[m][j]   (i, j < 128);   dWext[i][f] += sum_m A[m][i] Bext[m][f] (f < 4);
 *                    dbias[i] += sum_m A[m][i].   Any of (B, dW), (Bext, dWext), dbias may be NULL.  Accumulates
 *   atomically: zero the outputs first (or let several calls add up).
 * mmpde_node_wgrad_grouped: n_tasks independent contractions of that form in one launch (the five weight gradients of
 *   one message-passing layer: dW4, dW3 | x, dW3 | agg, dW1a, dW1b).  The CTAs are divided between the tasks in
 *   proportion to their row counts, so every output tile is summed over fewer partial tiles: at 36 k rows the atomic
 *   flush is half the time of a single-task launch. */
typedef struct mmpde_wgrad_task {
    const float* A; int64_t lda;        /* [M, >=128] */
    const float* B; int64_t ldb;        /* [M, >=128] or NULL */
    const float* Bext;                  /* [M,4] or NULL */
    float* dW; int64_t ldw;             /* [128, ldw] (NULL iff B is) */
    float* dWext; int64_t ldwext;       /* [128, ldwext >= 4] (NULL iff Bext is) */
    float* dbias;                       /* [128] or NULL */
    int64_t M;
} mmpde_wgrad_task;
int mmpde_node_gemm(const float* A0, int64_t lda0, const float* A1, int64_t lda1,
                    const float* W0, int64_t w0_ns, int64_t w0_ks, const float* W1, int64_t w1_ns, int64_t w1_ks,
                    const float* Aext, const float* Wext, const float* bias, int relu,
                    const float* R1, int64_t ldr1, const float* R2, int64_t ldr2,
                    float* C, int64_t ldc, int64_t M, void* stream);
/* Pre-split weight operands ("weight images").  The node contractions need W as two bf16 matrices (hi, lo) in the
 * tensor cores' SWIZZLE_128B K-major tile format.  mmpde_node_gemm builds them in registers in every CTA of every launch;
 * a training step launches ~100 contractions on ~60 distinct 128 x 128 weight blocks that only change in the optimizer.
 * mmpde_weight_images converts any number of blocks in one launch (image = MMPDE_WIMG_BYTES bytes, 128-byte aligned:
 * scale * W(n,k) for n, k < 128, element (n,k) read at W[n*w_ns + k*w_ks]); mmpde_node_gemm_img is mmpde_node_gemm
 * with the blocks given as images: the kernel fetches them with TMA bulk copies and moves them into tensor memory with
 * tcgen05.cp.  Same arithmetic, bit-identical results. */
#define MMPDE_WIMG_BYTES 65536
typedef struct mmpde_wimg_task {
    const float* W; int64_t w_ns, w_ks;
    float scale;
    void* image;
} mmpde_wimg_task;
int mmpde_weight_images(const mmpde_wimg_task* tasks, int n_tasks, void* stream);
int mmpde_node_gemm_img(const float* A0, int64_t lda0, const float* A1, int64_t lda1,
                        const void* image0, const void* image1,
                        const float* Aext, const float* Wext, const float* bias, int relu,
                        const float* R1, int64_t ldr1, const float* R2, int64_t ldr2,
                        float* C, int64_t ldc, int64_t M, void* stream);
int mmpde_node_wgrad(const float* A, int64_t lda, const float* B, int64_t ldb, const float* Bext,
                     float* dW, int64_t ldw, float* dWext, int64_t ldwext, float* dbias, int64_t M, void* stream);
int mmpde_node_wgrad_grouped(const mmpde_wgrad_task* tasks, int n_tasks, void* stream);

/* ---- message passing over the target-sorted edge list -------------------------------------------
 * Replaces PyG propagate + message_net_1/2 + scatter-mean (gnn_2d.py:55,59-63).  message_net_1 is split per
 * node: with e_ij = (u_i-u_j, px_i-px_j, py_i-py_j, v_i),
 *   z1_ij = P'[i] + Q'[j],  P' = x*W1a^T + node4*W1c^T + b1,  Q' = x*W1b^T - node4[:, :3]*W1c[:, :3]^T
 * (two node-level contractions, mmpde_gemm), so an edge costs h1 = relu(P'[dst] + Q'[src]) and the one dense
 * contraction z2 = W2*h1 + b2, which runs on the tcgen05 tensor cores as three split-bf16 products
 * (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) with W2 resident in tensor memory; the neighbour
 * gather is fused into the operand build, the per-target mean into the accumulator read-out.
 *
 *   PQ [N_src,256]: cols 0..127 = P', 128..255 = Q';
 *   edge_src / edge_dst [E] int32, sorted by dst;  inv_deg [N_dst] = 1/max(deg,1);
 *   w2 [128,128] fp32 row-major ([out,in]);  b2 [128];
 *   agg [N_dst, ld_agg]: mean message, MUST be zero on entry (partial segments add atomically);
 *   mask2: uint32 [ceil(E/128)*4, 128]: word [e/32][c] bit (e%32) = (z2[e][c] > 0), saved for the backward. */
int mmpde_edge_fwd(const float* PQ, const int32_t* edge_src, const int32_t* edge_dst, const float* inv_deg,
                   int64_t n_edges, const float* w2, const float* b2, float* agg, int64_t ld_agg,
                   uint32_t* mask2, void* stream);

/* Backward of the above (autograd of gnn_2d.py:59-63 + scatter-mean), h1 recomputed, z2 mask read.
 *   g_agg [N_dst, ld_gagg]: dL/d(mean message).
 *   Outputs (all ACCUMULATED atomically, zero them first):
 *   dPQ [N_src,256] (dP' summed per target, dQ' scattered to the source), dW2 [128,128], db2 [128].
 *   The caller turns dPQ into dW1 (a, b, c blocks), db1, dL/dx and dL/dnode4 with node-level contractions. */
int mmpde_edge_bwd(const float* PQ, const int32_t* edge_src, const int32_t* edge_dst, const float* inv_deg,
                   int64_t n_edges, const float* w2, const uint32_t* mask2, const float* g_agg, int64_t ld_gagg,
                   float* dPQ, float* dW2, float* db2, void* stream);

/* ---- BatchNorm over all nodes (PyG BatchNorm / nn.BatchNorm1d, gnn_2d.py:51,56,101,104) ---------
 * y = A (+ B if non-NULL), [M,128] with leading dims lda/ldb.
 * stats: sums [MMPDE_BN_REPLICAS][2,128] fp64 += (sum_c, sum_c^2), zero first.  The column sums are spread over
 *        MMPDE_BN_REPLICAS copies (same-address atomics of all CTAs would serialise in L2); the true sums are the
 *        sum over the copies.  For sync-BN fold the copies, all-reduce [2,128] and finalize with n_rep = 1.
 * finalize: from n_rep copies of sums and the GLOBAL row count -> mean_rstd [2,128] fp32, and, if running_* given,
 *           running stats update with momentum and the unbiased variance.
 * apply : out = gamma*(y-mean)*rstd + beta, optional ReLU. */
#define MMPDE_BN_REPLICAS 16
int mmpde_bn_stats(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, double* sums, void* stream);
int mmpde_bn_finalize(const double* sums, int n_rep, double count, float eps, float momentum,
                      float* mean_rstd, float* running_mean, float* running_var, void* stream);
int mmpde_bn_apply(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M,
                   const float* mean_rstd, const float* gamma, const float* beta, int relu,
                   float* out, int64_t ldo, void* stream);
/* backward: g [M,128] (ldg) is dL/d(out).  If relu, `out` (the forward output) gates g first.
 * reduce: bsums [MMPDE_BN_REPLICAS][2,128] fp64 += (sum g, sum g*yhat), spread over copies like `sums`; folded
 *         they are dbeta, dgamma (all-reduce the folded [2,128] for sync-BN)
 * apply : takes the FOLDED bsums [2,128]: gy = gamma*rstd*( g - sum_g/count - yhat*sum_gyhat/count ); written to gy (ldgy);
 *         accumulate!=0 adds into gy instead.  gy_gated (ldgg), if non-NULL, additionally receives gy * (B > 0): with
 *         y = x + relu_branch (B = the ReLU output of the residual branch) that is dL/d(pre-activation) of the branch,
 *         so the separate ReLU-backward pass over [M,128] disappears. */
int mmpde_bn_bwd_reduce(const float* g, int64_t ldg, const float* out, int64_t ldo, int relu,
                        const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M,
                        const float* mean_rstd, double* bsums, void* stream);
int mmpde_bn_bwd_apply(const float* g, int64_t ldg, const float* out, int64_t ldo, int relu,
                       const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M,
                       const float* mean_rstd, const float* gamma, const double* bsums, double count,
                       float* gy, int64_t ldgy, int accumulate, float* gy_gated, int64_t ldgg, void* stream);

/* ---- sync-BatchNorm sums across GPUs over NVLink peer memory (no reference counterpart) -------------
 * One kernel replaces the NCCL all-reduce of the [2,128] fp64 sums: it folds the n_rep local accumulator copies,
 * stores them into a slot of EVERY peer's exchange buffer, raises a flag there, waits for all peers' flags and adds
 * the world slots in rank order (identical bits on every rank) into out [2,128].
 * peer_base: device array [world] of the addresses (valid on THIS device) of every rank's exchange buffer of
 * MMPDE_BN_EXCHANGE_BYTES bytes, zero-initialised once, peer-mapped by the caller (e.g. torch symmetric memory).
 * All ranks must issue the same sequence of exchanges; the sequence number lives in the buffer, so the call can be
 * replayed from a CUDA graph.  A peer that never arrives makes the kernel trap after a wall-clock limit instead of
 * hanging: 600 s by default (a peer may legitimately be late -- checkpoint save on one rank, a re-recorded step graph),
 * overridden by the environment variable MMPDE_PEER_TIMEOUT_S or mmpde_bn_exchange_set_timeout(seconds). */
#define MMPDE_BN_EXCHANGE_BYTES (1024 + 4 * 16 * 256 * 8)      /* counter, flags, 4 slots x 16 ranks x 256 doubles */
int mmpde_bn_exchange(const double* sums, int n_rep, const int64_t* peer_base, int rank, int world, double* out,
                      void* stream);
int mmpde_bn_exchange_set_timeout(double seconds);

/* The reducing BatchNorm kernels with the rest of the pass done by their LAST CTA (a ticket counter, zeroed together
 * with the sums, tells a CTA that all partial sums have been delivered): fold of the accumulator copies, the cross-GPU
 * exchange of mmpde_bn_exchange (peer_base == NULL or world == 1: this rank only) and, forward, the finalisation of
 * mmpde_bn_finalize -- one launch instead of three or four per BatchNorm pass.
 * stats_fused: `count` = rows of the whole batch over all ranks; writes mean_rstd and updates the running statistics.
 * bwd_reduce_fused: local_out[256] = this rank's (sum g | sum g*yhat) (= dbeta | dgamma), glob_out[256] = the sums over
 * all ranks for mmpde_bn_bwd_apply; either may be NULL (glob_out NULL also skips the exchange: eval mode). */
int mmpde_bn_stats_fused(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, double* sums,
                         uint32_t* ticket, double count, float eps, float momentum, float* mean_rstd,
                         float* running_mean, float* running_var, const int64_t* peer_base, int rank, int world,
                         void* stream);
int mmpde_bn_bwd_reduce_fused(const float* g, int64_t ldg, const float* out, int64_t ldo, int relu,
                              const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M,
                              const float* mean_rstd, double* bsums, uint32_t* ticket, double* local_out,
                              double* glob_out, const int64_t* peer_base, int rank, int world, void* stream);
/* The same exchange in two halves, so that independent work can be queued between them: mmpde_bn_bwd_reduce_post reduces like
 * mmpde_bn_bwd_reduce_fused but its last CTA only delivers this rank's sums to the peers (local_out as above, world >= 2);
 * mmpde_bn_exchange_wait (one small kernel, later on the SAME stream, no other exchange of this buffer in between) waits for
 * the peers' sums and writes the sums over all ranks [256] to out. */
int mmpde_bn_bwd_reduce_post(const float* g, int64_t ldg, const float* out, int64_t ldo, int relu, const float* A,
                             int64_t lda, const float* B, int64_t ldb, int64_t M, const float* mean_rstd,
                             double* bsums, uint32_t* ticket, double* local_out, const int64_t* peer_base, int rank,
                             int world, void* stream);
int mmpde_bn_exchange_wait(const int64_t* peer_base, int rank, int world, double* out, void* stream);

/* ---- small elementwise helpers of the node path -------------------------------------------------
 * relu_bwd: out = g * (act > 0); colsum[128] += column sums of out (NULL to skip).  [M,128] */
int mmpde_relu_bwd(const float* g, int64_t ldg, const float* act, int64_t lda, int64_t M,
                   float* out, int64_t ldo, float* colsum, void* stream);
/* colsum[N] += sum over rows of A[M,N] */
int mmpde_colsum(const float* A, int64_t lda, int64_t M, int N, float* colsum, void* stream);

/* ---- Conv1d decoder over the feature axis (gnn_2d.py:108-114,136-139) ---------------------------
 * h [M,128] (ldh) -> out[M] = scale * conv3(relu(conv2(relu(conv1(h)))));  params packed fp32:
 * w1[4*16] b1[4] w2[8*4*12] b2[8] w3[8*8] b3[1]  (525 floats, torch Conv1d weight order). */
int mmpde_decoder_fwd(const float* h, int64_t ldh, int64_t M, const float* params, float scale,
                      float* out, void* stream);
/* The same forward, additionally saving the post-ReLU activations of the two hidden blocks for the backward:
 * act1 [M,256] (column c*38+i = channel c, position i of the first block; columns 152.. zero) and act2 [M,128]
 * (column o*9+j; columns 72.. zero), both contiguous and 16-byte aligned.  With them the backward runs as dense
 * (Toeplitz) contractions on mmpde_node_gemm / mmpde_node_wgrad_grouped while every ReLU mask comes from this fp32
 * evaluation (a mask taken from a split-bf16 contraction flips for pre-activations within ~1e-6 of zero). */
int mmpde_decoder_fwd_acts(const float* h, int64_t ldh, int64_t M, const float* params, float scale, float* out,
                           float* act1, float* act2, void* stream);
/* g_out[M] -> g_h [M,128] (ldg, overwritten) and g_params[525] (accumulated). */
int mmpde_decoder_bwd(const float* h, int64_t ldh, int64_t M, const float* params, float scale,
                      const float* g_out, float* g_h, int64_t ldg, float* g_params, void* stream);

/* out[m][c] = g[m] * w[c] * (act[m][c] > 0), c < 128: the decoder's last convolution is a dot product with w, so this is
 * dL/d(pre-activation) of the block in front of it.  The product path runs the decoder BACKWARD as Toeplitz contractions on
 * mmpde_node_gemm / mmpde_node_wgrad_grouped (ops._decoder_backward); mmpde_decoder_bwd above is the direct
 * warp-per-node form, kept as a second implementation the tests compare against. */
int mmpde_outer_gate(const float* g, const float* w, const float* act, int64_t lda, float* out, int64_t ldo, int64_t M,
                     void* stream);

/* ---- ItpNet 'res_cut' residual network on a regular grid (interpolate.py:54-63,95-97) -------------
 * out = tanh(conv4(tanh(conv3(tanh(conv2(tanh(conv1(x)))))))), Conv2d 5x5 / padding 2 with channels 1 -> 4 -> 16 -> 4 -> 1
 * (the reference's configuration, mmpde.py:343), x / out [B,1,H,W] fp32, one launch per direction (tile-resident stack).
 * params: w1[4*1*25] b1[4] w2[16*4*25] b2[16] w3[4*16*25] b3[4] w4[1*4*25] b4[1] (torch Conv2d layouts),
 * MMPDE_RESCUT_NPARAM floats.  acts [B, MMPDE_RESCUT_ACT_CHANNELS, H, W]: the three hidden activations, written by the
 * forward when non-NULL and read by the backward.  The backward OVERWRITES g_params[MMPDE_RESCUT_NPARAM] (x carries no
 * gradient); workspace = mmpde_rescut_bwd_workspace_floats(...) floats.  Deterministic (no atomics). */
#define MMPDE_RESCUT_NPARAM 3425
#define MMPDE_RESCUT_ACT_CHANNELS 24
int mmpde_rescut_fwd(const float* x, int64_t batch, int height, int width, const float* params, float* out,
                     float* acts, void* stream);
int64_t mmpde_rescut_bwd_workspace_floats(int64_t batch, int height, int width);
int mmpde_rescut_bwd(const float* x, int64_t batch, int height, int width, const float* params, const float* out,
                     const float* acts, const float* g_out, float* workspace, float* g_params, void* stream);

/* ---- DMM mesh mover, graph branch (mesh/dmm_model.py:94-142), forward only (the mover is frozen) --------
 * One tanh message-passing layer of hidden width 4 over a target-sorted edge list in CSR form:
 *   m_ij = tanh(W2 tanh(W1 [x_i | x_j | u_i-u_j | px_i-px_j | py_i-py_j] + b1) + b2),
 *   out[i] = x[i] + tanh(W4 tanh(W3 [x[i] | mean_j m_ij] + b3) + b4)      (the BatchNorm that follows stays in the caller).
 * x, out [N,4]; upos [N,4] = (u, px, py, unused); row_ptr int32 [N+1]; edge_src int32 [E]; weights = DEVICE pointer to
 * W1[4*11] b1[4] W2[4*4] b2[4] W3[4*8] b3[4] W4[4*4] b4[4] (nn.Linear layouts), MMPDE_DMM_GNN_NPARAM floats. */
#define MMPDE_DMM_GNN_NPARAM 124
int mmpde_dmm_gnn_layer(const float* x, const float* upos, const int32_t* row_ptr, const int32_t* edge_src,
                        int64_t n_nodes, const float* weights, float* out, void* stream);

/* ---- DMM mesh mover: displacement grad_xi phi(u, xi) (data_creator_2d.py:98-113, two autograd.grad calls through
 * mesh/dmm_model.py:145-219) as one forward-mode pass over the points, for the two-layer tanh trunk / out_nn:
 *   a = tanh(W1 xi + b1);  z = cst[n / per_sample] + M a;  h = tanh(z);
 *   out[n][d] = sum_j (1 - h_j^2) w_j sum_k (1 - a_k^2) W1[k][d] M[j][k],   d = 0, 1.
 * xi, out [N,2]; W1 [K,2], b1 [K] (trunk.layers[0], K <= 32); M [J,K] = Wt W2 (out_nn.layers[0][:, latent:] times
 * trunk.layers[1].weight); cst [samples,J] = Wl latent + Wt b2 + bo; w [J] = out_nn.layers[1].weight; J % 4 == 0, J <= 1024.
 * All pointers are device pointers. */
int mmpde_dmm_displacement(const float* xi, const float* W1, const float* b1, int K, const float* M, const float* cst,
                           const float* w, int J, int64_t n_points, int64_t per_sample, float* out, void* stream);

/* ---- fused k-NN interpolation (data_creator_2d.py:77-83 + interpolate.py:79-93) ----------------
 * For query q of sample s: p = (x_1,y_1,...,x_30,y_30,x_q,y_q) from idx[q,0..29];
 * w = Wc*tanh(Wb*tanh(Wa*p+ba)+bb)+bc;  out[q] = sum_k w_k * src_val[idx[q,k]].
 * params packed fp32: Wa[128*62] ba[128] Wb[64*128] bb[64] Wc[30*64] bc[30]  (18270 floats). */
int mmpde_itp_fwd(const float* src_xy, const float* src_val, const float* qry_xy, const int32_t* idx,
                  int64_t n_queries, const float* params, float* out, void* stream);
/* g_out[Q] -> g_params[18270] (accumulated), g_src_val[P] (accumulated atomically; NULL to skip). */
int mmpde_itp_bwd(const float* src_xy, const float* src_val, const float* qry_xy, const int32_t* idx,
                  int64_t n_queries, const float* params, const float* g_out,
                  float* g_params, float* g_src_val, void* stream);

/* The same interpolation on the tensor cores (tcgen05, split-bf16 products, 128 queries per tile; csrc/itp_tc.cu) -- the
 * form the product path uses (ops.InterpolateFn); mmpde_itp_fwd / mmpde_itp_bwd above are the direct fp32 form, kept as a
 * second implementation the tests compare against.  src_xy / qry_xy 8-byte aligned; idx and qry_xy 16-byte aligned
 * lets the neighbour lists come in through the TMA (otherwise they are loaded by the threads).
 * Backward: g_src_val[P] is accumulated atomically (NULL to skip); the weight gradients are left as the operands of
 * three contractions over the query axis, all [Q,128] fp32, 16-byte aligned, fully overwritten:
 *   G1 = dL/dza   G2 = [dL/dzb (64) | dL/dw (30) 0 0 | 0 (32)]   X1 = [p (62) 0 0 | hb (64)]   X2 = ha
 * dWa = (G1^T X1)[:, :62], ba = colsum G1;  dWb = (G2^T X2)[:64], bb = colsum G2[:, :64];
 * dWc = (G2^T X1)[64:94, 64:], bc = colsum G2[:, 64:94]  -- one mmpde_node_wgrad_grouped launch with three tasks. */
int mmpde_itp_fwd_tc(const float* src_xy, const float* src_val, const float* qry_xy, const int32_t* idx,
                     int64_t n_queries, const float* params, float* out, void* stream);
int mmpde_itp_bwd_tc(const float* src_xy, const float* src_val, const float* qry_xy, const int32_t* idx,
                     int64_t n_queries, const float* params, const float* g_out, float* g_src_val,
                     float* G1, float* G2, float* X1, float* X2, void* stream);

/* ---- halo exchange of the graph-partitioned processor (no reference counterpart) ----
 * Rows of a strided fp32 matrix <-> a contiguous buffer; ncols and the leading dimension are multiples of 4,
 * pointers 16-byte aligned.
 * gather:      out[i, 0:ncols]               = src[idx[i]*ld_src + 0:ncols]     (pack what a peer needs)
 * scatter_add: dst[idx[i]*ld_dst + 0:ncols] += in[i, 0:ncols]                   (atomic: idx may repeat) */
int mmpde_rows_gather(const float* src, int64_t ld_src, const int32_t* idx, int64_t n_rows, int ncols,
                      float* out, void* stream);
int mmpde_rows_scatter_add(const float* in, const int32_t* idx, int64_t n_rows, int ncols,
                           float* dst, int64_t ld_dst, void* stream);

/* out[m][n] = bias[n] + sum_f node4[m][f] W[n][f], n < 128, f < 4: the encoder's input layer nn.Linear(4,128)
 * (gnn_2d.py:100) as an elementwise pass.  node4 [M,4], W [128,4], bias [128] or NULL, out [M, ldo >= 128]; 16-byte aligned. */
int mmpde_node4_linear(const float* node4, const float* W, const float* bias, float* out, int64_t ldo, int64_t n_rows,
                       void* stream);
/* out[m*out_stride] (+)= sum_c A[m][c] * w[c]  (c < ncols, multiple of 4): the N = 1 contractions of the backward,
 * e.g. dL/du = dP'.W1c[:,0] - dQ'.W1c[:,0] of gnn_2d.py:61 (autograd) as ONE pass over dPQ [M,256]. */
int mmpde_rows_dot(const float* A, int64_t lda, int ncols, const float* w, float* out, int64_t out_stride,
                   int64_t n_rows, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMPDE_B200_H */
