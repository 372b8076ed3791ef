"""Host-side launch logic for the sm_100a kernels: tensor checks, the edge-list container, and the two
autograd Functions (whole-solver forward/backward, fused interpolation) that the module mirrors in
gnn_2d.py / data_creator_2d.py call.  All device work goes through the C ABI (_cabi.call); torch is
used for memory, streams, tiny O(parameters) reshuffles and autograd bookkeeping only.

Reference behaviour being reproduced: /root/reference/gnn_2d.py:53-69,119-141 (processor),
/root/reference/data_creator_2d.py:46-85 + /root/reference/interpolate.py:79-93 (interpolation).
"""
import os

import torch

from . import _cabi

H = 128
KN = 30
ITP_NPARAM = 18270
DEC_NPARAM = 525
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
BN_REPLICAS = 16            # include/mmpde_b200.h: MMPDE_BN_REPLICAS


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t, off=0):
    if t is None:
        return None
    return t.data_ptr() + off * t.element_size()


def _on(t):
    """Context that makes ``t``'s GPU the current device: the C ABI launches on torch's current stream of the current
    device and keeps its host-side caches (SM count, shared-memory opt-ins) per current device."""
    if not t.is_cuda:
        raise _cabi.MMPDEError("expected a CUDA tensor: the MM-PDE hot path has no CPU fallback")
    return torch.cuda.device(t.device)


def _chk(t, dtype=torch.float32, name="tensor"):
    if not t.is_cuda:
        raise _cabi.MMPDEError(f"{name} must be a CUDA tensor: the MM-PDE hot path has no CPU fallback")
    if t.device.index != torch.cuda.current_device():
        raise _cabi.MMPDEError(f"{name} lives on {t.device} but the current device is cuda:{torch.cuda.current_device()}: "
                               "all tensors of one call must be on the GPU the call is launched on")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


class _Comm:
    """Cross-rank coupling of the batch-sharded path (sync-BatchNorm statistics).  Single-GPU default:
    identity.  mmpde_b200.dist installs the torch.distributed (NCCL) version.  ``total_rows`` (set while a
    partitioned mesh is being processed) is the node count of the whole mesh = the BatchNorm row count."""
    total_rows = None
    # Which of the step's two concurrent solver branches is being recorded (train_helper_2d._forward_gnn issues the two
    # solvers on two streams).  Every branch owns its own cross-rank exchange sequence: the branches interleave
    # differently on different ranks, the exchanges WITHIN a branch come in the same order everywhere.
    branch = 0
    n_branches = 1

    def allreduce_(self, t):
        return t

    def reduce_bn_sums(self, spread, branch=0):
        """[n_rep, 256] local accumulator copies -> [256] sums over all ranks (here: one rank)."""
        return spread.sum(0) if spread.dim() == 2 else spread

    def peer_args(self, branch=0):
        """(peer_base pointer, rank, world) for the reducing BatchNorm kernels that exchange in their last CTA
        (mmpde_bn_stats_fused / mmpde_bn_bwd_reduce_fused), or None when the sums have to go through reduce_bn_sums."""
        return (None, 0, 1)

    def global_rows(self, n):
        return float(n) if self.total_rows is None else float(self.total_rows)


COMM = _Comm()


# ------------------------------------------------------------------------------------------------
# graph containers + construction
# ------------------------------------------------------------------------------------------------
class EdgeList:
    """Target-sorted edge list: int32 src/dst [E], inv_deg [N] = 1/max(in-degree, 1)."""

    def __init__(self, src, dst, inv_deg, n_nodes):
        self.src, self.dst, self.inv_deg, self.n_nodes = src, dst, inv_deg, int(n_nodes)
        self.n_edges = int(src.shape[0])

    _dst_cache = {}

    @classmethod
    def from_knn(cls, nbr, has_pad):
        """nbr int32 [N,k] from knn_graph_indices; rows are targets, entries sources (-1 = pad)."""
        N, k = nbr.shape
        dev = nbr.device
        if not has_pad:
            key = (N, k, dev)
            if key not in cls._dst_cache:
                cls._dst_cache[key] = (torch.arange(N, device=dev, dtype=torch.int32).repeat_interleave(k),
                                       torch.full((N,), 1.0 / k, device=dev, dtype=torch.float32))
            dst, inv_deg = cls._dst_cache[key]
            return cls(nbr.reshape(-1), dst, inv_deg, N)
        valid = nbr >= 0
        rows = torch.arange(N, device=dev, dtype=torch.int32)[:, None].expand(N, k)
        deg = valid.sum(1).clamp(min=1).to(torch.float32)
        return cls(nbr[valid].contiguous(), rows[valid].contiguous(), 1.0 / deg, N)

    @classmethod
    def from_edge_index(cls, edge_index, n_nodes):
        """edge_index [2,E] (row 0 = source j, row 1 = target i), any integer dtype; sorted by target if needed."""
        src, dst = edge_index[0], edge_index[1]
        if dst.numel() > 1 and not bool((dst[1:] >= dst[:-1]).all()):
            dst, order = torch.sort(dst, stable=True)
            src = src[order]
        deg = torch.bincount(dst, minlength=n_nodes).clamp(min=1).to(torch.float32)
        return cls(src.to(torch.int32).contiguous(), dst.to(torch.int32).contiguous(), 1.0 / deg, n_nodes)

    def edge_index(self):
        return torch.stack((self.src.long(), self.dst.long()))


GRID_MIN_POINTS = 384          # samples at least this large use the cell-binned search


def knn_indices(pts, pts_off, qry, qry_off, k, rule, exclude_self, bbox=None, per_sample=None):
    """Ordered k nearest points of each query inside its sample -> int32 [Q,k] global rows of pts (-1 pads).
    pts/qry fp32 [.,2]; *_off int32 [S+1] on the device.  rule 0 = fp32 graph rule, 1 = fp64 interpolation rule.
    ``bbox`` = (x0, y0, x1, y1) enclosing (most of) the points and ``per_sample`` = points per sample (host
    ints, so no device sync) switch large samples to the cell-binned search; both paths are exact and
    return identical indices."""
    with _on(pts):
        _chk(pts, name="pts"); _chk(qry, name="qry")
        _chk(pts_off, torch.int32, "pts_off"); _chk(qry_off, torch.int32, "qry_off")
        if bbox is not None and per_sample is not None and per_sample >= GRID_MIN_POINTS:
            return _knn_grid(pts, pts_off, qry, qry_off, k, rule, exclude_self, bbox, per_sample)
        Q = qry.shape[0]
        out = torch.empty((Q, k), dtype=torch.int32, device=pts.device)
        _cabi.call("mmpde_knn", _ptr(pts), _ptr(pts_off), _ptr(qry), _ptr(qry_off), pts_off.numel() - 1, Q, k, rule,
                   int(exclude_self), _ptr(out), _stream())
        return out


class CellBins:
    """Points binned into the uniform cell grid of the exact search (mmpde_knn_grid_build): reusable for every search
    over the same points (and across steps for points that never move, e.g. the reference grid)."""

    def __init__(self, pts, pts_off, bbox, per_sample, pts_per_cell=6.0):
        x0, y0, x1, y1 = [float(v) for v in bbox]
        S = pts_off.numel() - 1
        P = pts.shape[0]
        area = max((x1 - x0) * (y1 - y0), 1e-30)
        cell = max((area * pts_per_cell / max(per_sample, 1)) ** 0.5, 1e-9)
        self.x0, self.y0, self.inv_cell = x0, y0, 1.0 / cell
        self.gx = max(int((x1 - x0) / cell) + 1, 1)
        self.gy = max(int((y1 - y0) / cell) + 1, 1)
        self.pts, self.pts_off, self.S = pts, pts_off, S
        dev = pts.device
        ncell = S * self.gx * self.gy
        cell_of = torch.empty(P, dtype=torch.int32, device=dev)
        self.cell_start = torch.empty(ncell + 1, dtype=torch.int32, device=dev)
        cursor = torch.empty(ncell, dtype=torch.int32, device=dev)
        self.order = torch.empty(P, dtype=torch.int32, device=dev)
        with _on(pts):
            _cabi.call("mmpde_knn_grid_build", _ptr(pts), _ptr(pts_off), S, P, x0, y0, self.inv_cell, self.gx, self.gy,
                       _ptr(cell_of), _ptr(self.cell_start), _ptr(cursor), _ptr(self.order), _stream())

    def task(self, qry, qry_off, k, rule, exclude_self, out):
        return _cabi.KnnTask(_ptr(self.pts), _ptr(self.pts_off), _ptr(qry), _ptr(qry_off), self.S, k, qry.shape[0], self.x0,
                             self.y0, self.inv_cell, self.gx, self.gy, _ptr(self.cell_start), _ptr(self.order), rule,
                             int(exclude_self), _ptr(out))


def knn_grid_multi(searches):
    """Several cell-binned searches in ONE launch.  searches: (bins, qry, qry_off, k, rule, exclude_self) each; returns the
    int32 [Q,k] neighbour lists.  One search alone fills ~10 % of the GPU's warp slots."""
    import ctypes
    outs, tasks = [], []
    with _on(searches[0][1]):
        for bins, qry, qry_off, k, rule, exclude_self in searches:
            _chk(qry, name="qry"); _chk(qry_off, torch.int32, "qry_off")
            out = torch.empty((qry.shape[0], k), dtype=torch.int32, device=qry.device)
            outs.append(out)
            tasks.append(bins.task(qry, qry_off, k, rule, exclude_self, out))
        arr = (_cabi.KnnTask * len(tasks))(*tasks)
        _cabi.call("mmpde_knn_grid_multi", ctypes.addressof(arr), len(tasks), _stream())
    return outs


def _knn_grid(pts, pts_off, qry, qry_off, k, rule, exclude_self, bbox, per_sample, pts_per_cell=6.0):
    bins = CellBins(pts, pts_off, bbox, per_sample, pts_per_cell)
    return knn_grid_multi([(bins, qry, qry_off, k, rule, exclude_self)])[0]


def knn_indices_grid(pts, qry, k, rule, exclude_self):
    """Cell-binned search for ONE large sample with the bounding box measured on the device (one host sync)."""
    lo = torch.minimum(pts.min(0).values, qry.min(0).values)
    hi = torch.maximum(pts.max(0).values, qry.max(0).values)
    bbox = torch.cat((lo, hi)).tolist()
    dev = pts.device
    po = torch.tensor([0, pts.shape[0]], dtype=torch.int32, device=dev)
    qo = torch.tensor([0, qry.shape[0]], dtype=torch.int32, device=dev)
    return _knn_grid(pts, po, qry, qo, k, rule, exclude_self, bbox, pts.shape[0])


def radius_indices(pts, off, r, max_nb=32):
    with _on(pts):
        _chk(pts, name="pts"); _chk(off, torch.int32, "off")
        out = torch.empty((pts.shape[0], max_nb), dtype=torch.int32, device=pts.device)
        _cabi.call("mmpde_radius", _ptr(pts), _ptr(off), off.numel() - 1, pts.shape[0], float(r), max_nb, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# thin kernel wrappers (pointer + leading-dimension style, mirroring the C ABI)
# ------------------------------------------------------------------------------------------------
def gemm(A, lda, a_k, B, ldb, b_k, C, ldc, M, N, K, bias=None, r1_row=None, r1_stride=0, r1_col=None,
         relu=0, acc=0, split_k=1, st=None):
    _cabi.call("mmpde_gemm", A, lda, a_k, B, ldb, b_k, C, ldc, M, N, K, bias, r1_row, r1_stride, r1_col,
               relu, acc, split_k, st if st is not None else _stream())


WIMG_BYTES = 65536         # include/mmpde_b200.h: MMPDE_WIMG_BYTES
USE_WEIGHT_IMAGES = __import__("os").environ.get("MMPDE_WEIGHT_IMAGES", "1") != "0"     # 0: every CTA splits the fp32 weights itself


def weight_images(blocks, device, st=None):
    """blocks: [(W pointer, w_ns, w_ks)] or [(W pointer, w_ns, w_ks, scale)] -> ([image pointer per block], the uint8
    buffer holding them).  One launch pre-splits all 128 x 128 weight blocks into the bf16 hi | lo operand images the
    tensor-core kernels fetch by TMA (mmpde_weight_images)."""
    import ctypes
    buf = torch.empty(max(len(blocks), 1), WIMG_BYTES, dtype=torch.uint8, device=device)
    base = buf.data_ptr()
    tasks = [_cabi.WimgTask(b[0], b[1], b[2], b[3] if len(b) > 3 else 1.0, base + i * WIMG_BYTES) for i, b in enumerate(blocks)]
    if tasks:
        arr = (_cabi.WimgTask * len(tasks))(*tasks)
        _cabi.call("mmpde_weight_images", ctypes.addressof(arr), len(tasks), st if st is not None else _stream())
    return [base + i * WIMG_BYTES for i in range(len(blocks))], buf


def node_gemm(A0, lda0, W0, w0_ns, w0_ks, C, ldc, M, A1=None, lda1=0, W1=None, w1_ns=0, w1_ks=0, ext=None, bias=None, relu=0,
              R1=None, ldr1=0, R2=None, ldr2=0, st=None, img0=None, img1=None):
    """tcgen05 node contraction (include/mmpde_b200.h: mmpde_node_gemm); ext = (node4 pointer, Wext pointer).  With
    ``img0`` (and ``img1`` for the second K segment) the weight blocks are taken from pre-split images
    (mmpde_node_gemm_img) and the W pointers / strides are ignored."""
    aext, wext = ext if ext is not None else (None, None)
    if img0 is not None:
        _cabi.call("mmpde_node_gemm_img", A0, lda0, A1, lda1, img0, img1, aext, wext, bias, relu, R1, ldr1, R2, ldr2, C, ldc, M,
                   st if st is not None else _stream())
        return
    _cabi.call("mmpde_node_gemm", A0, lda0, A1, lda1, W0, w0_ns, w0_ks, W1, w1_ns, w1_ks, aext, wext, bias, relu,
               R1, ldr1, R2, ldr2, C, ldc, M, st if st is not None else _stream())


def node_wgrad(A, lda, M, B=None, ldb=0, dW=None, ldw=0, Bext=None, dWext=None, dbias=None, st=None):
    """tcgen05 weight gradient (mmpde_node_wgrad): dW += A^T B, dWext += A^T node4, dbias += colsum(A)."""
    _cabi.call("mmpde_node_wgrad", A, lda, B, ldb, Bext, dW, ldw, dWext, 4 if dWext is not None else 0, dbias, M,
               st if st is not None else _stream())


def wgrad_task(A, lda, M, B=None, ldb=0, dW=None, ldw=0, Bext=None, dWext=None, dbias=None):
    """One entry of a grouped weight-gradient launch (same arguments as node_wgrad)."""
    return _cabi.WgradTask(A, lda, B, ldb, Bext, dW, ldw, dWext, 4 if dWext is not None else 0, dbias, M)


def node_wgrad_grouped(tasks, st=None):
    """All weight gradients of one layer in ONE launch (mmpde_node_wgrad_grouped)."""
    import ctypes
    arr = (_cabi.WgradTask * len(tasks))(*tasks)
    _cabi.call("mmpde_node_wgrad_grouped", ctypes.addressof(arr), len(tasks), st if st is not None else _stream())


def _split_for(rows):
    """split-K factor of the weight-gradient contractions (K = node count): ~one 256-row chunk per CTA so the
    1-2 output tiles still spread over the whole chip."""
    return max(2, min(512, (rows + 255) // 256))


# multi-rank backward: weight-gradient launches queued between the halves of the BatchNorm exchanges (see _bn_backward)
SPLIT_BN_EXCHANGE = os.environ.get("MMPDE_SPLIT_BN_EXCHANGE", "1") != "0"


# Width (CTAs per persistent kernel, 0 = one per SM) of the solver pass being built; train_helper_2d._forward_gnn sets it
# around each of the two solver calls of an overlapped training step.  SolverFn keeps it for its backward.
SOLVER_WIDTH = 0


class persistent_ctas:
    """``with persistent_ctas(n):`` -- the persistent tensor-core kernels launched inside use at most n CTAs (0 = one per SM)
    on ``device``; see mmpde_set_persistent_ctas.  Not a launch: it only changes the grid of the launches that follow."""

    def __init__(self, n, device=None):
        self.n, self.device = int(n), device

    def _set(self, n):
        with torch.cuda.device(self.device):
            rc = _cabi.lib().mmpde_set_persistent_ctas(n)
        if rc != 0:
            raise _cabi.MMPDEError(f"mmpde_set_persistent_ctas({n}) failed: {rc}")

    def __enter__(self):
        if self.n > 0:
            self._set(self.n)
        return self

    def __exit__(self, *exc):
        if self.n > 0:
            self._set(0)
        return False


class _BNState:
    """mean/rstd [2,128] of one BatchNorm application (saved for the backward)."""
    __slots__ = ("mean_rstd", "count", "rows", "training", "branch")


BN_ACC = BN_REPLICAS * 2 * H + 1       # fp64 accumulator copies of one BatchNorm pass + the ticket of its last-CTA tail


def bn_accumulators(n, device):
    """n zeroed accumulator blocks [n, BN_ACC] (one fill for all BatchNorm passes of a solver pass)."""
    return torch.zeros(n, BN_ACC, dtype=torch.float64, device=device)


def _bn_forward(items, gamma, beta, relu, training, rmean, rvar, nbt, st, sums=None):
    """BatchNorm over the rows of ALL local parts (and all ranks, through COMM).
    items: one (A, lda, B, ldb, M, out, ldo) per local part, y = A (+ B).  Training: the column sums of the last part's
    launch are folded, summed over the ranks and turned into mean / rstd / running statistics by that launch's last CTA
    (mmpde_bn_stats_fused); only without peer memory (NCCL) the three steps are separate launches."""
    state = _BNState()
    dev = gamma.device
    state.rows = sum(it[4] for it in items)
    state.training = bool(training)
    state.branch = COMM.branch                         # the backward of this BatchNorm exchanges on the same sequence
    if training:
        if sums is None:                               # else: a zeroed [BN_ACC] block handed in by the solver
            sums = bn_accumulators(1, dev)[0]
        state.count = COMM.global_rows(state.rows)
        state.mean_rstd = torch.empty(2 * H, dtype=torch.float32, device=dev)
        peer = COMM.peer_args(state.branch)
        live = [it for it in items if it[4] > 0]
        if peer is not None and live:
            for A, lda, B, ldb, M, _, _ in live[:-1]:
                _cabi.call("mmpde_bn_stats", A, lda, B, ldb, M, _ptr(sums), st)
            A, lda, B, ldb, M, _, _ = live[-1]
            _cabi.call("mmpde_bn_stats_fused", A, lda, B, ldb, M, _ptr(sums), _ptr(sums, BN_ACC - 1), state.count, BN_EPS,
                       BN_MOMENTUM, _ptr(state.mean_rstd), _ptr(rmean), _ptr(rvar), peer[0], peer[1], peer[2], st)
        else:
            for A, lda, B, ldb, M, _, _ in items:
                _cabi.call("mmpde_bn_stats", A, lda, B, ldb, M, _ptr(sums), st)
            red, n_rep = sums[:BN_ACC - 1].view(BN_REPLICAS, 2 * H), BN_REPLICAS
            if state.count != float(state.rows):      # other ranks hold rows too: fold + sum over the ranks -> [2,128]
                red, n_rep = COMM.reduce_bn_sums(red, state.branch), 1
            _cabi.call("mmpde_bn_finalize", _ptr(red), n_rep, state.count, BN_EPS, BN_MOMENTUM, _ptr(state.mean_rstd),
                       _ptr(rmean), _ptr(rvar), st)
        if nbt is not None:        # (the solver passes None and bumps all of its counters with one launch)
            nbt += 1
    else:
        state.count = float(state.rows)
        state.mean_rstd = torch.cat((rmean, torch.rsqrt(rvar + BN_EPS)))
    for A, lda, B, ldb, M, out, ldo in items:
        _cabi.call("mmpde_bn_apply", A, lda, B, ldb, M, _ptr(state.mean_rstd), _ptr(gamma), _ptr(beta), int(relu), out, ldo, st)
    return state


def _bn_backward(items, relu, state, gamma, st, spread=None, between=None):
    """items: one (g, ldg, out, ldo, A, lda, B, ldb, M, gy, ldgy[, gy_gated, ldgg]) per local part (gy_gated =
    gy * (B > 0), the ReLU backward of a residual branch B fused into this pass).  Returns this rank's
    (dgamma, dbeta) as fp64 views; writes dL/dy into gy.  With several ranks the two column sums are summed over the
    ranks for the normalisation term (sync-BN), while the parameter grads stay per-rank sums (the gradient all-reduce adds
    them up afterwards).  Fold and cross-rank sum happen in the last CTA of the last part's reducing launch.
    ``between``: independent work of the caller (the deferred weight-gradient launch of the layer above).  With several ranks
    it is queued BETWEEN the two halves of the exchange -- the reducing launch only delivers this rank's sums, a one-CTA
    kernel collects the peers' right before the apply pass -- so the link latency and a peer that is some tens of
    microseconds behind (each rank interleaves its two solver branches in its own order) do not stall this rank's chain;
    otherwise it simply runs first."""
    dev = gamma.device
    if spread is None:
        spread = bn_accumulators(1, dev)[0]
    peer = COMM.peer_args(state.branch)
    live = [it for it in items if it[8] > 0]
    multi = COMM.global_rows(state.rows) != float(state.rows)
    split = between is not None and peer is not None and bool(live) and multi and state.training and SPLIT_BN_EXCHANGE
    if between is not None and not split:
        between()
    if peer is not None and live:
        both = torch.empty(2, 2 * H, dtype=torch.float64, device=dev)
        local, glob = both[0], both[1]
        for g, ldg, out, ldo, A, lda, B, ldb, M, *_ in live[:-1]:
            _cabi.call("mmpde_bn_bwd_reduce", g, ldg, out, ldo, int(relu), A, lda, B, ldb, M, _ptr(state.mean_rstd), _ptr(spread), st)
        g, ldg, out, ldo, A, lda, B, ldb, M, *_ = live[-1]
        want_glob = state.training
        if split:
            _cabi.call("mmpde_bn_bwd_reduce_post", g, ldg, out, ldo, int(relu), A, lda, B, ldb, M, _ptr(state.mean_rstd),
                       _ptr(spread), _ptr(spread, BN_ACC - 1), _ptr(local), peer[0], peer[1], peer[2], st)
            between()
            _cabi.call("mmpde_bn_exchange_wait", peer[0], peer[1], peer[2], _ptr(glob), st)
        else:
            _cabi.call("mmpde_bn_bwd_reduce_fused", g, ldg, out, ldo, int(relu), A, lda, B, ldb, M, _ptr(state.mean_rstd),
                       _ptr(spread), _ptr(spread, BN_ACC - 1), _ptr(local), _ptr(glob) if want_glob else None,
                       peer[0] if multi else None, peer[1] if multi else 0, peer[2] if multi else 1, st)
        if not want_glob:
            glob = torch.zeros_like(local)
    else:
        for g, ldg, out, ldo, A, lda, B, ldb, M, *_ in items:
            _cabi.call("mmpde_bn_bwd_reduce", g, ldg, out, ldo, int(relu), A, lda, B, ldb, M, _ptr(state.mean_rstd), _ptr(spread), st)
        red = spread[:BN_ACC - 1].view(BN_REPLICAS, 2 * H)
        local = red.sum(0)
        glob = local
        if state.training and multi:
            glob = COMM.reduce_bn_sums(red, state.branch)
        if not state.training:
            glob = torch.zeros_like(local)
    # eval mode normalised with the running statistics: they do not depend on the batch, so the two batch-mean terms
    # vanish (glob = 0) and dL/dy = g * gamma * rstd (what nn.BatchNorm1d.eval() gives); dgamma / dbeta stay the sums.
    for g, ldg, out, ldo, A, lda, B, ldb, M, gy, ldgy, *gated in items:
        gyg, ldgg = gated if gated else (None, 0)
        _cabi.call("mmpde_bn_bwd_apply", g, ldg, out, ldo, int(relu), A, lda, B, ldb, M, _ptr(state.mean_rstd), _ptr(gamma),
                   _ptr(glob), state.count, gy, ldgy, 0, gyg, ldgg, st)
    return local[H:], local[:H]                        # fp64 views; callers convert all of a solver's at once


# ------------------------------------------------------------------------------------------------
# the processor: encoder -> L message-passing layers -> Conv1d decoder, as ONE autograd node.
# Written over a LIST of graph parts that advance in lock step: one part = the whole (batched) graph in the
# ordinary case; several parts = a partitioned mesh whose halo rows are exchanged once per layer
# (partition.py).  Parts of other ranks are reached through the exchange object.
# ------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# Conv1d decoder (gnn_2d.py:108-114,136-139) as three dense contractions on the tcgen05 node-GEMM kernels.
# A Conv1d over the 128-wide feature axis is a banded (Toeplitz) matrix: Conv1d(1,4,16,s3) = T1 [152,128] with
# T1[c*38+i, 3i+k] = w1[c,0,k];  Conv1d(4,8,12,s3) = T2 [72,152] with T2[o*9+j, c*38+3j+k] = w2[o,c,k];  Conv1d(8,1,8,s2)
# reads positions 0..7 of the 9 = a dot product with w3row[o*9+j] = w3[0,o,j].  Padded to the kernels' 128-wide blocks
# (T1 -> [256,128], T2 -> [128,256]) the BACKWARD of the stack is three data-gradient GEMMs (ReLU gates fused into the
# epilogues), ONE grouped weight-gradient launch, and an index_add that folds the dense matrix gradients back onto the
# 525 convolution parameters.  The forward stays the direct fp32 kernel (see _decoder_forward).
# ------------------------------------------------------------------------------------------------
_DEC = {}
DEC_OFF = dict(w1=0, b1=64, w2=68, b2=452, w3=460, b3=524)


def _decoder_maps(device):
    """Static maps: (gather) element of (T1 | b1e | T2 | b2e | w3row) -> index into the flat parameter vector padded with
    a zero at position 525; (fold) parameter -> the dense elements holding it."""
    key = str(device)
    if key not in _DEC:
        Z = DEC_NPARAM
        t1 = torch.full((2 * H, H), Z, dtype=torch.int64)
        b1 = torch.full((2 * H,), Z, dtype=torch.int64)
        for c in range(4):
            for i in range(38):
                b1[c * 38 + i] = DEC_OFF["b1"] + c
                for k in range(16):
                    t1[c * 38 + i, 3 * i + k] = DEC_OFF["w1"] + c * 16 + k
        t2 = torch.full((H, 2 * H), Z, dtype=torch.int64)
        b2 = torch.full((H,), Z, dtype=torch.int64)
        w3 = torch.full((H,), Z, dtype=torch.int64)
        for o in range(8):
            for j in range(9):
                b2[o * 9 + j] = DEC_OFF["b2"] + o
                if j < 8:
                    w3[o * 9 + j] = DEC_OFF["w3"] + o * 8 + j
                for c in range(4):
                    for k in range(12):
                        t2[o * 9 + j, c * 38 + 3 * j + k] = DEC_OFF["w2"] + (o * 4 + c) * 12 + k
        gather = torch.cat((t1.reshape(-1), b1, t2.reshape(-1), b2, w3))
        # inverse map for the gradient fold: parameter p <- the (<= 38) dense elements that hold it, padded with the index
        # of an appended zero; a gather + row sum is deterministic, an index_add_ with duplicates is not
        n_dense = gather.numel()
        order = torch.argsort(gather, stable=True)
        counts = torch.bincount(gather, minlength=Z + 1)[:Z]
        fold = torch.full((Z, int(counts.max())), n_dense, dtype=torch.int64)
        start = torch.cumsum(counts, 0) - counts
        for p_ in range(Z):
            fold[p_, :int(counts[p_])] = order[int(start[p_]):int(start[p_]) + int(counts[p_])]
        _DEC[key] = (gather.to(device), fold.to(device))
    return _DEC[key]


_DEC_SPLIT = [2 * H * H, 2 * H, H * 2 * H, H, H]          # T1 b1e T2 b2e w3row


def _decoder_forward(parts, hs, dec, scale, st):
    """hs[p] [n_own,128] -> out[p] [n_own]; returns (outs, saved).  The forward is the direct fp32 kernel (every ReLU mask
    must come from an fp32 evaluation: a mask taken from a split-bf16 contraction flips for pre-activations within ~1e-6
    of zero, ~1e-3 of the rows, and each flip moves that row's upstream gradient by several percent); it also writes the
    activations in the Toeplitz layout for the backward."""
    f32 = dict(dtype=torch.float32, device=dec.device)
    outs, saved = [], []
    for part, h in zip(parts, hs):
        N = part.n_own
        out = torch.empty(N, **f32)
        A1, C2 = torch.empty(N, 2 * H, **f32), torch.empty(N, H, **f32)
        _cabi.call("mmpde_decoder_fwd_acts", _ptr(h), H, N, _ptr(dec), float(scale), _ptr(out), _ptr(A1), _ptr(C2), st)
        outs.append(out)
        saved.append((A1, C2))
    return outs, saved


def _decoder_backward(parts, hs, dec, scale, dec_saved, g_outs, st):
    """-> ([dL/dh per part], dL/d(flat decoder parameters) [525])."""
    saved = dec_saved
    dev = dec.device
    f32 = dict(dtype=torch.float32, device=dev)
    gather_map, fold_map = _decoder_maps(dev)
    padded = torch.cat((dec, dec.new_zeros(1)))
    T1, _, T2, _, w3row = torch.split(padded[gather_map], _DEC_SPLIT)
    w3s = w3row * scale
    # dense gradients of (T1 | b1e | T2 | b2e | w3row), one zeroed buffer, folded onto the parameters at the end
    flat = torch.zeros(sum(_DEC_SPLIT) + 1, **f32)         # + one slot that stays zero (padding target of the fold)
    dT1, db1e, dT2, db2e, dw3, _ = torch.split(flat, _DEC_SPLIT + [1])
    dw3x = torch.zeros(H, 4, **f32)                   # column 0 = C2^T g_out (the extension slot of the weight-gradient kernel)
    g_b3 = torch.zeros((), **f32)
    g_hs, tasks, keep = [], [], []
    for part, h, (A1, C2), g_out in zip(parts, hs, saved, g_outs):
        N = part.n_own
        g = g_out.contiguous().view(-1)
        g_b3 = g_b3 + g.sum()
        g4 = torch.zeros(N, 4, **f32)
        g4[:, 0] = g
        # dL/dC2 = g_out (x) (scale * w3row), gated by the ReLU of C2
        g_C2 = torch.empty(N, H, **f32)
        _cabi.call("mmpde_outer_gate", _ptr(g), _ptr(w3s), _ptr(C2), H, _ptr(g_C2), H, N, st)
        g_A1 = torch.empty(N, 2 * H, **f32)
        node_gemm(_ptr(g_C2), H, _ptr(T2), 1, 2 * H, _ptr(g_A1), 2 * H, N, relu=2, R1=_ptr(A1), ldr1=2 * H, st=st)
        node_gemm(_ptr(g_C2), H, _ptr(T2, H), 1, 2 * H, _ptr(g_A1, H), 2 * H, N, relu=2, R1=_ptr(A1, H), ldr1=2 * H, st=st)
        g_h = torch.empty(N, H, **f32)
        node_gemm(_ptr(g_A1), 2 * H, _ptr(T1), 1, H, _ptr(g_h), H, N, A1=_ptr(g_A1, H), lda1=2 * H, W1=_ptr(T1, H * H), w1_ns=1,
                  w1_ks=H, st=st)
        g_hs.append(g_h)
        tasks += [wgrad_task(_ptr(g_C2), H, N, B=_ptr(A1), ldb=2 * H, dW=_ptr(dT2), ldw=2 * H, dbias=_ptr(db2e)),
                  wgrad_task(_ptr(g_C2), H, N, B=_ptr(A1, H), ldb=2 * H, dW=_ptr(dT2, H), ldw=2 * H),
                  wgrad_task(_ptr(g_A1), 2 * H, N, B=_ptr(h), ldb=H, dW=_ptr(dT1), ldw=H, dbias=_ptr(db1e)),
                  wgrad_task(_ptr(g_A1, H), 2 * H, N, B=_ptr(h), ldb=H, dW=_ptr(dT1, H * H), ldw=H, dbias=_ptr(db1e, H)),
                  wgrad_task(_ptr(C2), H, N, Bext=_ptr(g4), dWext=_ptr(dw3x))]
        keep.append((g4, g_C2, g_A1))
    node_wgrad_grouped(tasks, st)
    del keep
    dw3.copy_(dw3x[:, 0] * scale)
    g_dec = flat[fold_map].sum(1)
    g_dec[DEC_OFF["b3"]] = g_b3 * scale
    return g_hs, g_dec


N_ENC = 8          # We1 be1 g1 bt1 We2 be2 g2 bt2
N_LAYER = 10       # W1 b1 W2 b2 W3 b3 W4 b4 gamma beta


class GraphPart:
    """node4 [n_own,4] of the owned nodes, target-sorted edges in local numbering (sources may point into the
    halo rows n_own .. n_own+n_halo), and the partition plan (None = no halo)."""

    def __init__(self, node4, edges, plan=None):
        self.node4, self.edges, self.plan = node4, edges, plan
        self.n_own = int(node4.shape[0])
        self.n_src = self.n_own + (plan.n_halo if plan is not None else 0)


def mask_words(n_edges):
    """uint32 words of the z2 sign mask the edge kernels exchange: [ceil(E/128)*4, 128]."""
    return max((n_edges + 127) // 128, 1) * 512


def _layer_prep(W1s, W3s):
    """Derived weights of SEVERAL layers in a handful of launches (instead of ~9 tiny ones per layer and pass).
    W1c = W1[:, 256:260] acts on e_ij = (u_i-u_j, px_i-px_j, py_i-py_j, v_i) (gnn_2d.py:61); split per node it is W1c on
    the target side and -W1c with the v column removed on the source side.  Returns per-layer tuples
    (w1c [128,4], w1cq [128,4], w3x [128,4], wu [256])."""
    L = len(W1s)
    w1c = torch.stack([W[:, 2 * H:2 * H + 4] for W in W1s])              # [L,128,4]
    w1cq = -w1c
    w1cq[:, :, 3] = 0.0
    w3x = torch.zeros(L, H, 4, dtype=torch.float32, device=w1c.device)   # update_net_1 sees [x, agg, v]: v = node4[:, 3]
    w3x[:, :, 3] = torch.stack([W[:, 2 * H] for W in W3s])
    wu = torch.cat((w1c[:, :, 0], w1cq[:, :, 0]), dim=1)                 # [L,256]: dL/du = [dP' | dQ'] . wu
    return [(w1c[l], w1cq[l], w3x[l], wu[l]) for l in range(L)]


def _layer_image_blocks(lp, backward):
    """The 128 x 128 weight blocks one layer's node contractions read: forward p q u1a u1b u2, backward (transposed
    reads) u2T u1aT u1bT pT qT."""
    W1, W3, W4 = lp[0], lp[4], lp[6]
    blocks = [(_ptr(W1), 260, 1), (_ptr(W1, H), 260, 1), (_ptr(W3), 257, 1), (_ptr(W3, H), 257, 1), (_ptr(W4), H, 1)]
    if backward:
        blocks += [(_ptr(W4), 1, H), (_ptr(W3), 1, 257), (_ptr(W3, H), 1, 257), (_ptr(W1), 1, 260), (_ptr(W1, H), 1, 260)]
    return blocks


class _LayerImages:
    """Image pointers of one layer (None = let the kernel split the fp32 weights itself)."""
    __slots__ = ("p", "q", "u1a", "u1b", "u2", "u2T", "u1aT", "u1bT", "pT", "qT")

    def __init__(self, ptrs=None):
        for k, name in enumerate(self.__slots__):
            setattr(self, name, ptrs[k] if ptrs is not None and k < len(ptrs) else None)


_NO_IMAGES = _LayerImages()


LAYER_GRAD_SIZES = [H * 260, H, H * H, H, H * 257, H, H * H, H, 2 * H * 4, H * 4]    # dW1 db1 dW2 db2 dW3 db3 dW4 db4 dW1c dW3x


def _layer_forward(parts, Xs, lp, bnbuf, training, nxts, exch, st, prep=None, bn_sums=None, im=_NO_IMAGES):
    """One GNN_Layer_FS_2D (gnn_2d.py:53-69) on every part.  Xs[p] [n_own,256]: cols 0..127 hold the layer input
    h, cols 128..255 must be ZERO on entry and receive the mean message.  Output BN(h + update) -> nxts[p] =
    (pointer, leading dimension)."""
    W1, b1, W2, b2, W3, b3, W4, b4, gam, bet = lp
    w1c, w1cq, w3x, _ = prep if prep is not None else _layer_prep([W1], [W3])[0]
    # message_net_1 split per node (gnn_2d.py:61): z1_ij = P'[i] + Q'[j] with
    #   P' = h W1a^T + b1 + node4 W1c^T,   Q' = h W1b^T - node4[:, :3] W1c[:, :3]^T
    PQs = []
    for part, Xl in zip(parts, Xs):
        N = part.n_own
        f32 = dict(dtype=torch.float32, device=Xl.device)
        x, n4 = _ptr(Xl), _ptr(part.node4)
        PQ = torch.empty(part.n_src, 2 * H, **f32)
        node_gemm(x, 2 * H, _ptr(W1), 260, 1, _ptr(PQ), 2 * H, N, ext=(n4, _ptr(w1c)), bias=_ptr(b1), st=st, img0=im.p)
        node_gemm(x, 2 * H, _ptr(W1, H), 260, 1, _ptr(PQ, H), 2 * H, N, ext=(n4, _ptr(w1cq)), st=st, img0=im.q)
        PQs.append(PQ)
    if exch is not None:
        exch.forward(PQs)                                         # Q' rows of the halo nodes
    saved, bn_items = [], []
    for part, Xl, PQ, nxt in zip(parts, Xs, PQs, nxts):
        N, E, edges = part.n_own, part.edges.n_edges, part.edges
        f32 = dict(dtype=torch.float32, device=Xl.device)
        x = _ptr(Xl)
        mask2 = torch.empty(mask_words(E), dtype=torch.int32, device=Xl.device)
        _cabi.call("mmpde_edge_fwd", _ptr(PQ), _ptr(edges.src), _ptr(edges.dst), _ptr(edges.inv_deg), E,
                   _ptr(W2), _ptr(b2), _ptr(Xl, H), 2 * H, _ptr(mask2), st)
        # update_net_1/2 + residual                                 (gnn_2d.py:65-69)
        h3 = torch.empty(N, H, **f32)
        node_gemm(x, 2 * H, _ptr(W3), 257, 1, _ptr(h3), H, N, A1=_ptr(Xl, H), lda1=2 * H, W1=_ptr(W3, H), w1_ns=257, w1_ks=1,
                  ext=(_ptr(part.node4), _ptr(w3x)), bias=_ptr(b3), relu=1, st=st, img0=im.u1a, img1=im.u1b)
        r4 = torch.empty(N, H, **f32)
        node_gemm(_ptr(h3), H, _ptr(W4), H, 1, _ptr(r4), H, N, bias=_ptr(b4), relu=1, st=st, img0=im.u2)
        saved.append((PQ, mask2, h3, r4))
        bn_items.append((x, 2 * H, _ptr(r4), H, N, nxt[0], nxt[1]))
    bn = _bn_forward(bn_items, gam, bet, 0, training, *bnbuf, st, sums=bn_sums)
    return saved, bn


def _layer_backward(parts, Xs, lp, saved, bn, g_hs, g_node4s, exch, st, prep=None, flat=None, bn_spread=None, im=_NO_IMAGES,
                    pending=None, defer=False):
    """Backward of _layer_forward.  g_hs[p] [n_own,128] = dL/d(output).  Returns ([dL/dh_in per part], 10 param
    grads summed over the local parts); adds the layer's dL/du into g_node4s[p][:,0] when given.
    ``pending``: the deferred weight-gradient launch of the layer above, handed to this layer's BatchNorm backward;
    ``defer``: return this layer's own weight-gradient launch as a third value instead of running it at the end."""
    W1, b1, W2, b2, W3, b3, W4, b4, gam, bet = lp
    dev = W1.device
    f32 = dict(dtype=torch.float32, device=dev)
    wu = (prep if prep is not None else _layer_prep([W1], [W3])[0])[3]
    # BatchNorm backward: y = h + r4
    # (its second output g_z4 = g_y * (r4 > 0) is the ReLU backward of update_net_2, fused into the same pass)
    g_ys = [torch.empty(part.n_own, H, **f32) for part in parts]
    g_z4s = [torch.empty(part.n_own, H, **f32) for part in parts]
    items = [(_ptr(g_h), H, None, 0, _ptr(Xl), 2 * H, _ptr(sv[3]), H, part.n_own, _ptr(g_y), H, _ptr(g_z4), H)
             for part, Xl, sv, g_h, g_y, g_z4 in zip(parts, Xs, saved, g_hs, g_ys, g_z4s)]
    dgam, dbet = _bn_backward(items, 0, bn, gam, st, spread=bn_spread, between=pending)
    # all accumulators of this layer in ONE zeroed buffer (every block is a multiple of 4 floats: 16-byte aligned rows)
    fold_here = flat is None
    if fold_here:
        flat = torch.zeros(sum(LAYER_GRAD_SIZES), **f32)
    dW1, db1, dW2, db2, dW3, db3, dW4, db4, dW1c, dW3x = torch.split(flat, LAYER_GRAD_SIZES)
    dW1, dW2, dW3, dW4 = dW1.view(H, 260), dW2.view(H, H), dW3.view(H, 257), dW4.view(H, H)
    dW1c, dW3x = dW1c.view(2, H, 4), dW3x.view(H, 4)
    dPQs, wtasks, keep = [], [], []
    for part, Xl, sv, g_y, g_z4 in zip(parts, Xs, saved, g_ys, g_z4s):
        PQ, mask2, h3, r4 = sv
        N, E, edges = part.n_own, part.edges.n_edges, part.edges
        x, n4 = _ptr(Xl), _ptr(part.node4)
        # node MLP backward (update_net_2, update_net_1).  No stand-alone ReLU-backward passes: g_z4 came out of the
        # BatchNorm backward, g_z3 = (g_z4 W4) * (h3 > 0) is gated in the epilogue of its dgrad, and the bias
        # gradients are the ones-column of the weight-gradient contractions.
        # The five weight-gradient contractions of the layer are only queued here: they run as ONE grouped launch
        # at the end (nothing but the optimizer waits for them), so their operands stay alive until then.
        wtasks.append(wgrad_task(_ptr(g_z4), H, N, B=_ptr(h3), ldb=H, dW=_ptr(dW4), ldw=H, dbias=_ptr(db4)))
        g_z3 = torch.empty(N, H, **f32)
        node_gemm(_ptr(g_z4), H, _ptr(W4), 1, H, _ptr(g_z3), H, N, relu=2, R1=_ptr(h3), ldr1=H, st=st, img0=im.u2T)
        wtasks.append(wgrad_task(_ptr(g_z3), H, N, B=x, ldb=2 * H, dW=_ptr(dW3), ldw=257, Bext=n4, dWext=_ptr(dW3x), dbias=_ptr(db3)))
        wtasks.append(wgrad_task(_ptr(g_z3), H, N, B=_ptr(Xl, H), ldb=2 * H, dW=_ptr(dW3, H), ldw=257))
        # dL/dh_in so far: g_y (residual) + g_z3 W3[:, :128];  dL/d(mean message) = g_z3 W3[:, 128:256]
        node_gemm(_ptr(g_z3), H, _ptr(W3), 1, 257, _ptr(g_y), H, N, R1=_ptr(g_y), ldr1=H, st=st, img0=im.u1aT)
        g_agg = torch.empty(N, H, **f32)
        node_gemm(_ptr(g_z3), H, _ptr(W3, H), 1, 257, _ptr(g_agg), H, N, st=st, img0=im.u1bT)
        keep.append((g_z3, g_z4))
        # message passing backward
        dPQ = torch.zeros(part.n_src, 2 * H, **f32)
        _cabi.call("mmpde_edge_bwd", _ptr(PQ), _ptr(edges.src), _ptr(edges.dst), _ptr(edges.inv_deg), E,
                   _ptr(W2), _ptr(mask2), _ptr(g_agg), H, _ptr(dPQ), _ptr(dW2), _ptr(db2), st)
        dPQs.append(dPQ)
    if exch is not None:
        exch.backward(dPQs)                                       # dL/dQ' of halo rows -> added at their owners
    for idx, (part, Xl, dPQ, g_y) in enumerate(zip(parts, Xs, dPQs, g_ys)):
        N = part.n_own
        x, n4 = _ptr(Xl), _ptr(part.node4)
        # message_net_1 parameters from dP', dQ':  dW1a = dP'^T h, dW1b = dQ'^T h, dW1c = dP'^T node4 - dQ'^T node4[:, :3]
        wtasks.append(wgrad_task(_ptr(dPQ), 2 * H, N, B=x, ldb=2 * H, dW=_ptr(dW1), ldw=260, Bext=n4, dWext=_ptr(dW1c), dbias=_ptr(db1)))
        wtasks.append(wgrad_task(_ptr(dPQ, H), 2 * H, N, B=x, ldb=2 * H, dW=_ptr(dW1, H), ldw=260, Bext=n4, dWext=_ptr(dW1c, 4 * H)))
        g_node4 = g_node4s[idx] if g_node4s is not None else None
        if g_node4 is not None:      # dL/du (column 0 of node4): dP' W1c[:,0] - dQ' W1c[:,0], one pass over dPQ
            _cabi.call("mmpde_rows_dot", _ptr(dPQ), 2 * H, 2 * H, _ptr(wu), _ptr(g_node4), 4, N, 1, st)
        # dL/dh_in += dP' W1a + dQ' W1b
        node_gemm(_ptr(dPQ), 2 * H, _ptr(W1), 1, 260, _ptr(g_y), H, N, A1=_ptr(dPQ, H), lda1=2 * H, W1=_ptr(W1, H), w1_ns=1,
                  w1_ks=260, R1=_ptr(g_y), ldr1=H, st=st, img0=im.pT, img1=im.qT)
    if defer:
        assert not fold_here
        alive = [keep, dPQs, g_z4s, list(Xs), list(saved)]          # operands of the queued contractions

        def run_wgrads():
            node_wgrad_grouped(wtasks, st)
            alive.clear()
        return g_ys, [dW1, db1, dW2, db2, dW3, db3, dW4, db4, dgam, dbet], run_wgrads
    node_wgrad_grouped(wtasks, st)
    del keep
    if fold_here:
        _fold_extension_grads(flat.view(1, -1))
        dgam, dbet = dgam.to(torch.float32), dbet.to(torch.float32)
    return g_ys, [dW1, db1, dW2, db2, dW3, db3, dW4, db4, dgam, dbet]


def _fold_extension_grads(flat2d):
    """flat2d [L, sum(LAYER_GRAD_SIZES)]: move the node-scalar extension gradients of every layer into the weight
    columns they belong to -- dW3[:, 256] = dW3x[:, 3];  dW1[:, 256:260] = dW1c[P'] - dW1c[Q'] (no v column on the
    source side) -- with one launch per term for all layers."""
    L = flat2d.shape[0]
    off = [0]
    for n in LAYER_GRAD_SIZES:
        off.append(off[-1] + n)
    dW1 = flat2d[:, off[0]:off[1]].view(L, H, 260)
    dW3 = flat2d[:, off[4]:off[5]].view(L, H, 257)
    dW1c = flat2d[:, off[8]:off[9]].view(L, 2, H, 4)
    dW3x = flat2d[:, off[9]:off[10]].view(L, H, 4)
    dW3[:, :, 2 * H] = dW3x[:, :, 3]
    dW1[:, :, 2 * H:2 * H + 4] = dW1c[:, 0]
    dW1[:, :, 2 * H:2 * H + 3] -= dW1c[:, 1, :, :3]


class LayerFn(torch.autograd.Function):
    """One stand-alone GNN_Layer_FS_2D.forward (gnn_2d.py:53-57): out[N,128] = BN(x + update(x, mean messages))."""

    @staticmethod
    def forward(ctx, x, node4, edges, training, bnbuf, *lp):
        with _on(x):
            _chk(x, name="x"); _chk(node4, name="node4")
            st = _stream()
            N = x.shape[0]
            Xl = torch.zeros(N, 2 * H, dtype=torch.float32, device=x.device)
            Xl[:, :H] = x
            out = torch.empty(N, H, dtype=torch.float32, device=x.device)
            parts = [GraphPart(node4, edges)]
            saved, bn = _layer_forward(parts, [Xl], lp, bnbuf, training, [(_ptr(out), H)], None, st)
        ctx.stuff = (parts, Xl, lp, saved, bn)
        return out

    @staticmethod
    def backward(ctx, g_out):
        parts, Xl, lp, saved, bn = ctx.stuff
        with _on(Xl):
            g_node4 = torch.zeros_like(parts[0].node4) if ctx.needs_input_grad[1] else None
            g_xs, grads = _layer_backward(parts, [Xl], lp, saved, bn, [g_out.contiguous()],
                                          [g_node4] if g_node4 is not None else None, None, _stream())
        return (g_xs[0], g_node4, None, None, None, *grads)


def _solver_forward(parts, L, training, scale, bn_buffers, params, exch, st, want_backward=True):
    """Encoder, L layers and decoder on every part.  Returns ([out [n_own] per part], saved state)."""
    dev = parts[0].node4.device
    f32 = dict(dtype=torch.float32, device=dev)
    We1, be1, g1, bt1, We2, be2, g2, bt2 = params[:N_ENC]
    dec = params[N_ENC + N_LAYER * L]
    # ---- encoder: Linear(4,128) BN ReLU Linear(128,128) BN        (gnn_2d.py:99-106,130-131)
    preps = _layer_prep([params[N_ENC + N_LAYER * l] for l in range(L)], [params[N_ENC + N_LAYER * l + 4] for l in range(L)]) if L else []
    bn_sums = bn_accumulators(2 + L, dev) if training else [None] * (2 + L)
    if training:                # num_batches_tracked of all 2 + L BatchNorms: one launch instead of one per BatchNorm pass
        counters = [b[2] for b in bn_buffers if b[2] is not None]
        if counters:
            torch._foreach_add_(counters, 1)
        bn_buffers = [(b[0], b[1], None) for b in bn_buffers]
    # every weight block of the pass as bf16 hi | lo operand images, one launch (the weights only change in the optimizer)
    per = 10 if want_backward else 5
    blocks = [(_ptr(We2), H, 1)] + ([(_ptr(We2), 1, H)] if want_backward else [])
    n_enc_img = len(blocks)
    for l in range(L):
        blocks += _layer_image_blocks(params[N_ENC + N_LAYER * l: N_ENC + N_LAYER * (l + 1)], want_backward)
    if USE_WEIGHT_IMAGES:
        img_ptrs, img_buf = weight_images(blocks, dev, st)
        images = [_LayerImages(img_ptrs[n_enc_img + per * l: n_enc_img + per * (l + 1)]) for l in range(L)]
        enc_images = (img_ptrs[0], img_ptrs[1] if want_backward else None)
    else:
        img_buf, images, enc_images = None, [_NO_IMAGES] * L, (None, None)
    e1s, e1ns, e2s = [], [], []
    for part in parts:
        e1 = torch.empty(part.n_own, H, **f32)
        _cabi.call("mmpde_node4_linear", _ptr(part.node4), _ptr(We1), _ptr(be1), _ptr(e1), H, part.n_own, st)
        e1s.append(e1)
        e1ns.append(torch.empty(part.n_own, H, **f32))
    bn1 = _bn_forward([(_ptr(e1), H, None, 0, part.n_own, _ptr(e1n), H) for part, e1, e1n in zip(parts, e1s, e1ns)],
                      g1, bt1, 1, training, *bn_buffers[0], st, sums=bn_sums[0])
    for part, e1n in zip(parts, e1ns):
        e2 = torch.empty(part.n_own, H, **f32)
        node_gemm(_ptr(e1n), H, _ptr(We2), H, 1, _ptr(e2), H, part.n_own, bias=_ptr(be2), st=st, img0=enc_images[0])
        e2s.append(e2)
    # X[l][p] = [h_l | agg_l]  ([n_own,256]); the last hidden state lives alone in hL
    X = [[torch.zeros(part.n_own, 2 * H, **f32) for part in parts] for _ in range(L)]
    hL = [torch.empty(part.n_own, H, **f32) for part in parts]

    def dest(l):
        return [(_ptr(t), 2 * H) for t in X[l]] if l < L else [(_ptr(t), H) for t in hL]

    bn2 = _bn_forward([(_ptr(e2), H, None, 0, part.n_own, d[0], d[1]) for part, e2, d in zip(parts, e2s, dest(0))],
                      g2, bt2, 0, training, *bn_buffers[1], st, sums=bn_sums[1])
    layers = []
    for l in range(L):
        lp = params[N_ENC + N_LAYER * l: N_ENC + N_LAYER * (l + 1)]
        layers.append(_layer_forward(parts, X[l], lp, bn_buffers[2 + l], training, dest(l + 1), exch, st,
                                     prep=preps[l], bn_sums=bn_sums[2 + l], im=images[l]))
    outs, dec_saved = _decoder_forward(parts, hL, dec, float(scale), st)
    return outs, dict(parts=parts, L=L, scale=float(scale), params=params, enc=(e1s, e1ns, e2s, bn1, bn2), X=X, hL=hL,
                      layers=layers, exch=exch, preps=preps, dec=dec_saved, images=images, enc_images=enc_images, img_buf=img_buf)


def _solver_backward(sv, g_outs, need_u, st):
    """Returns ([dL/dnode4 per part] or None, parameter grads summed over the local parts)."""
    parts, L, params, exch = sv["parts"], sv["L"], sv["params"], sv["exch"]
    dev = parts[0].node4.device
    f32 = dict(dtype=torch.float32, device=dev)
    We1, be1, g1, bt1, We2, be2, g2, bt2 = params[:N_ENC]
    dec = params[N_ENC + N_LAYER * L]
    e1s, e1ns, e2s, bn1, bn2 = sv["enc"]
    grads = [None] * len(params)
    g_node4s = [torch.zeros(part.n_own, 4, **f32) for part in parts] if need_u else None
    g_hs, g_dec = _decoder_backward(parts, sv["hL"], dec, sv["scale"], sv["dec"], g_outs, st)
    grads[N_ENC + N_LAYER * L] = g_dec
    flat_all = torch.zeros(max(L, 1), sum(LAYER_GRAD_SIZES), **f32)   # every layer's gradient accumulators, zeroed at once
    spread_all = bn_accumulators(2 + L, dev)
    # several ranks with the peer-memory exchange: a layer's weight-gradient launch is deferred into the BatchNorm backward
    # of the layer below, between the two halves of its cross-GPU exchange (_bn_backward)
    defer = SPLIT_BN_EXCHANGE and COMM.peer_args(COMM.branch) is not None and getattr(COMM, "world", 1) > 1
    pending = None
    for l in reversed(range(L)):
        base = N_ENC + N_LAYER * l
        saved, bn = sv["layers"][l]
        res = _layer_backward(parts, sv["X"][l], params[base:base + N_LAYER], saved, bn, g_hs, g_node4s, exch, st,
                              prep=sv["preps"][l], flat=flat_all[l], bn_spread=spread_all[2 + l], im=sv["images"][l],
                              pending=pending, defer=defer)
        g_hs, lg = res[0], res[1]
        pending = res[2] if defer else None
        grads[base:base + N_LAYER] = lg
    # ---- encoder backward
    g_e2s = [torch.empty(part.n_own, H, **f32) for part in parts]
    dg2, db2_ = _bn_backward([(_ptr(g_h), H, None, 0, _ptr(e2), H, None, 0, part.n_own, _ptr(g_e2), H)
                              for part, g_h, e2, g_e2 in zip(parts, g_hs, e2s, g_e2s)], 0, bn2, g2, st, spread=spread_all[1],
                             between=pending)
    if L:
        _fold_extension_grads(flat_all)
    dWe2, dbe2 = torch.zeros(H, H, **f32), torch.zeros(H, **f32)
    dWe1, dbe1 = torch.zeros(H, 4, **f32), torch.zeros(H, **f32)
    g_e1ns = []
    for part, g_e2, e1n in zip(parts, g_e2s, e1ns):
        node_wgrad(_ptr(g_e2), H, part.n_own, B=_ptr(e1n), ldb=H, dW=_ptr(dWe2), ldw=H, dbias=_ptr(dbe2), st=st)
        g_e1n = torch.empty(part.n_own, H, **f32)
        node_gemm(_ptr(g_e2), H, _ptr(We2), 1, H, _ptr(g_e1n), H, part.n_own, st=st, img0=sv["enc_images"][1])
        g_e1ns.append(g_e1n)
    g_e1s = g_e2s                                                     # reuse
    dg1, db1_ = _bn_backward([(_ptr(g_e1n), H, _ptr(e1n), H, _ptr(e1), H, None, 0, part.n_own, _ptr(g_e1), H)
                              for part, g_e1n, e1n, e1, g_e1 in zip(parts, g_e1ns, e1ns, e1s, g_e1s)], 1, bn1, g1, st,
                             spread=spread_all[0])
    we1_u = We1[:, 0].contiguous() if need_u else None
    for idx, (part, g_e1) in enumerate(zip(parts, g_e1s)):
        node_wgrad(_ptr(g_e1), H, part.n_own, Bext=_ptr(part.node4), dWext=_ptr(dWe1), dbias=_ptr(dbe1), st=st)
        if need_u:      # only the u column: positions/time feed the frozen mesh mover only
            _cabi.call("mmpde_rows_dot", _ptr(g_e1), H, H, _ptr(we1_u), _ptr(g_node4s[idx]), 4, part.n_own, 1, st)
    grads[:N_ENC] = [dWe1, dbe1, dg1, db1_, dWe2, dbe2, dg2, db2_]
    # BatchNorm parameter gradients come back as fp64 views: convert all of them with two launches
    bn_idx = [2, 3, 6, 7] + [N_ENC + N_LAYER * l + k for l in range(L) for k in (8, 9)]
    bn32 = torch.stack([grads[k] for k in bn_idx]).to(torch.float32)
    for row, k in enumerate(bn_idx):
        grads[k] = bn32[row]
    return g_node4s, grads


class SolverFn(torch.autograd.Function):
    """out[N,1] = MP_PDE_Solver_2D(node4, edges).  params = 8 encoder + 10 per layer + 1 flat decoder tensor.
    bn_buffers = [(running_mean, running_var, num_batches_tracked)] * (2 + L), updated in place when training."""

    @staticmethod
    def forward(ctx, node4, edges, n_layers, training, scale, bn_buffers, *params):
        with _on(node4):
            _chk(node4, name="node4")
            for i, p in enumerate(params):
                _chk(p, name=f"param{i}")
            # (grad mode is always off inside Function.forward: "a backward will follow" = some input requires grad)
            width = int(SOLVER_WIDTH) if any(ctx.needs_input_grad) else 0
            with persistent_ctas(width, node4.device):
                outs, ctx.sv = _solver_forward([GraphPart(node4, edges)], n_layers, training, scale, bn_buffers, params, None,
                                               _stream(), want_backward=any(ctx.needs_input_grad))
            if ctx.sv is not None:
                ctx.sv["width"] = width
        return outs[0].view(-1, 1)

    @staticmethod
    def backward(ctx, g_out):
        with _on(g_out), persistent_ctas(ctx.sv.get("width", 0), g_out.device):
            g_node4s, grads = _solver_backward(ctx.sv, [g_out], ctx.needs_input_grad[0], _stream())
        return (g_node4s[0] if g_node4s is not None else None, None, None, None, None, None, *grads)


class PartitionedSolverFn(torch.autograd.Function):
    """The same processor on a partitioned mesh: ``parts`` = [(edges, plan)] of the parts living in THIS process
    (one per rank in a multi-GPU run; all of them in the single-process emulation), ``exch`` moves the halo rows
    (partition.LocalExchange / dist.HaloExchange).  Tensor inputs: n_parts node4 tensors, then the parameters.
    Returns one [n_own,1] tensor per part."""

    @staticmethod
    def forward(ctx, parts, exch, n_layers, training, scale, bn_buffers, *tensors):
        n = len(parts)
        node4s, params = tensors[:n], tensors[n:]
        for t in node4s:
            _chk(t, name="node4")
        gps = [GraphPart(n4, edges, plan) for n4, (edges, plan) in zip(node4s, parts)]
        COMM.total_rows = float(parts[0][1].n_total)
        try:
            with _on(node4s[0]):
                outs, ctx.sv = _solver_forward(gps, n_layers, training, scale, bn_buffers, params, exch, _stream(),
                                               want_backward=any(ctx.needs_input_grad))
        finally:
            COMM.total_rows = None
        ctx.n = n
        return tuple(o.view(-1, 1) for o in outs)

    @staticmethod
    def backward(ctx, *g_outs):
        need_u = any(ctx.needs_input_grad[6:6 + ctx.n])
        COMM.total_rows = float(ctx.sv["parts"][0].plan.n_total)
        try:
            with _on(g_outs[0]):
                g_node4s, grads = _solver_backward(ctx.sv, list(g_outs), need_u, _stream())
        finally:
            COMM.total_rows = None
        g4 = g_node4s if g_node4s is not None else [None] * ctx.n
        return (None, None, None, None, None, None, *g4, *grads)


# ------------------------------------------------------------------------------------------------
# fused interpolation
# ------------------------------------------------------------------------------------------------
ITP_TENSOR_CORES = __import__("os").environ.get("MMPDE_ITP_TC", "1") != "0"     # 0: the direct fp32 kernels (csrc/itp.cu), for A/B checks


class InterpolateFn(torch.autograd.Function):
    """out[Q] = sum_k ItpNet(p_q)_k * src_val[idx[q,k]]   (data_creator_2d.py:77-83, interpolate.py:79-93) on the tensor
    cores (csrc/itp_tc.cu).  Gradients: flat ItpNet parameters and src_val; coordinates are treated as constants (they
    only lead to the frozen mesh mover).  The backward kernel redoes the forward, runs the
    two data-gradient contractions and leaves the operands of the three weight-gradient contractions over the query axis,
    which go out as ONE grouped tcgen05 weight-gradient launch."""

    @staticmethod
    def forward(ctx, src_val, src_xy, qry_xy, idx, flat_params):
        assert flat_params.numel() == ITP_NPARAM and idx.shape[1] == KN
        Q = qry_xy.shape[0]
        with _on(src_val):
            _chk(src_val, name="src_val"); _chk(src_xy, name="src_xy"); _chk(qry_xy, name="qry_xy")
            _chk(idx, torch.int32, "idx"); _chk(flat_params, name="flat_params")
            out = torch.empty(Q, dtype=torch.float32, device=src_val.device)
            _cabi.call("mmpde_itp_fwd_tc" if ITP_TENSOR_CORES else "mmpde_itp_fwd", _ptr(src_xy), _ptr(src_val), _ptr(qry_xy),
                       _ptr(idx), Q, _ptr(flat_params), _ptr(out), _stream())
        ctx.save_for_backward(src_val, src_xy, qry_xy, idx, flat_params)
        return out

    @staticmethod
    def backward(ctx, g_out):
        src_val, src_xy, qry_xy, idx, flat_params = ctx.saved_tensors
        g_out = g_out.contiguous()
        Q = qry_xy.shape[0]
        dev = src_val.device
        g_val = torch.zeros_like(src_val) if ctx.needs_input_grad[0] else None
        if not ITP_TENSOR_CORES:
            with _on(src_val):
                g_params = torch.zeros_like(flat_params)
                _cabi.call("mmpde_itp_bwd", _ptr(src_xy), _ptr(src_val), _ptr(qry_xy), _ptr(idx), Q, _ptr(flat_params),
                           _ptr(g_out), _ptr(g_params), _ptr(g_val), _stream())
            return g_val, None, None, None, g_params
        with _on(src_val):
            st = _stream()
            ws = torch.empty(4, max(Q, 1), H, dtype=torch.float32, device=dev)      # G1 G2 X1 X2
            G1, G2, X1, X2 = (_ptr(ws[k]) for k in range(4))
            _cabi.call("mmpde_itp_bwd_tc", _ptr(src_xy), _ptr(src_val), _ptr(qry_xy), _ptr(idx), Q, _ptr(flat_params),
                       _ptr(g_out), _ptr(g_val), G1, G2, X1, X2, st)
            acc = torch.zeros(3 * H * H + 2 * H, dtype=torch.float32, device=dev)
            T, db = acc[:3 * H * H].view(3, H, H), acc[3 * H * H:].view(2, H)
            node_wgrad_grouped([wgrad_task(G1, H, Q, B=X1, ldb=H, dW=_ptr(T[0]), ldw=H, dbias=_ptr(db[0])),
                                wgrad_task(G2, H, Q, B=X2, ldb=H, dW=_ptr(T[1]), ldw=H, dbias=_ptr(db[1])),
                                wgrad_task(G2, H, Q, B=X1, ldb=H, dW=_ptr(T[2]), ldw=H)], st)
            g_params = torch.cat((T[0][:, :62].reshape(-1), db[0], T[1][:64].reshape(-1), db[1][:64],
                                  T[2][64:64 + KN, 64:].reshape(-1), db[1][64:64 + KN]))
        return g_val, None, None, None, g_params


RESCUT_CHANNELS = (1, 4, 16, 4, 1)
RESCUT_NPARAM = 3425           # include/mmpde_b200.h: MMPDE_RESCUT_NPARAM
RESCUT_ACT_CHANNELS = 24
_RESCUT_SHAPES = [(4, 1, 5, 5), (4,), (16, 4, 5, 5), (16,), (4, 16, 5, 5), (4,), (1, 4, 5, 5), (1,)]


class ResCutFn(torch.autograd.Function):
    """ItpNet 'res_cut' (interpolate.py:54-63,95-97): four 5x5 convolutions 1 -> 4 -> 16 -> 4 -> 1 with tanh, on
    [B,1,H,W], as ONE tile-resident launch per direction (csrc/rescut.cu) instead of 12 cuDNN launches.  Gradients go to
    the eight convolution parameters; the input field carries none (it is the step's data)."""

    @staticmethod
    def forward(ctx, data, *params):
        assert [tuple(p.shape) for p in params] == _RESCUT_SHAPES
        B, _, Hh, Ww = data.shape
        with _on(data):
            data = _chk(data.contiguous(), name="data")
            flat = torch.cat([p.reshape(-1) for p in params])
            out = torch.empty(B, 1, Hh, Ww, dtype=torch.float32, device=data.device)
            need = any(ctx.needs_input_grad[1:])
            acts = torch.empty(B, RESCUT_ACT_CHANNELS, Hh, Ww, dtype=torch.float32, device=data.device) if need else None
            _cabi.call("mmpde_rescut_fwd", _ptr(data), B, Hh, Ww, _ptr(flat), _ptr(out), _ptr(acts), _stream())
        ctx.save_for_backward(data, flat, out, acts)
        return out

    @staticmethod
    def backward(ctx, g_out):
        data, flat, out, acts = ctx.saved_tensors
        B, _, Hh, Ww = data.shape
        with _on(data):
            g_out = _chk(g_out.contiguous(), name="g_out")
            tiles = B * ((Hh + 15) // 16) * ((Ww + 15) // 16)
            ws = torch.empty(max(tiles, 1) * RESCUT_NPARAM, dtype=torch.float32, device=data.device)
            g_flat = torch.empty(RESCUT_NPARAM, dtype=torch.float32, device=data.device)
            _cabi.call("mmpde_rescut_bwd", _ptr(data), B, Hh, Ww, _ptr(flat), _ptr(out), _ptr(acts), _ptr(g_out), _ptr(ws),
                       _ptr(g_flat), _stream())
        grads = [g.reshape(s) for g, s in zip(torch.split(g_flat, [int(torch.Size(s).numel()) for s in _RESCUT_SHAPES]), _RESCUT_SHAPES)]
        return (None, *grads)


def interpolate_direct(src_val, src_xy, qry_xy, idx, flat_params, g_out=None, want_g_val=True):
    """The direct fp32 form of the same operator (mmpde_itp_fwd / mmpde_itp_bwd, csrc/itp.cu): a second implementation
    the tests hold the tensor-core path against.  Returns out, or (out, g_params, g_val) when g_out is given."""
    Q = qry_xy.shape[0]
    with _on(src_val):
        out = torch.empty(Q, dtype=torch.float32, device=src_val.device)
        _cabi.call("mmpde_itp_fwd", _ptr(src_xy), _ptr(src_val), _ptr(qry_xy), _ptr(idx), Q, _ptr(flat_params), _ptr(out), _stream())
        if g_out is None:
            return out
        g_params = torch.zeros_like(flat_params)
        g_val = torch.zeros_like(src_val) if want_g_val else None
        _cabi.call("mmpde_itp_bwd", _ptr(src_xy), _ptr(src_val), _ptr(qry_xy), _ptr(idx), Q, _ptr(flat_params),
                   _ptr(g_out.contiguous()), _ptr(g_params), _ptr(g_val), _stream())
    return out, g_params, g_val
