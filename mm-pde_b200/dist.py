"""One-process-per-GPU scaling of the hot path over NCCL / NVLink (the reference has no distributed code;
SURVEY.md 8e).  Batch (trajectory) sharding: every rank runs mesh move, k-NN, interpolation and both
solvers on its own samples; the only couplings are
  (i)  BatchNorm batch statistics -> all-reduce of the fp64 [2,128] column sums per BN application
       (forward and backward), installed into ops.COMM;
  (ii) gradients -> ONE flat-bucket all-reduce per step (~1.28 M fp32 = 5 MB, latency-bound), averaged so
       that the result equals the global-batch mean-MSE gradient (mmpde.py:33-36).
On CPU (tests) the same code runs over gloo.
"""
import os

import torch
import torch.distributed as dist

from . import ops


BN_EXCHANGE_BYTES = 1024 + 4 * 16 * 256 * 8          # include/mmpde_b200.h: MMPDE_BN_EXCHANGE_BYTES


class PeerExchange:
    """Exchange buffers for mmpde_bn_exchange: one per rank in torch symmetric memory, so every rank can store into
    every peer's buffer over NVLink.  ``sum(spread)`` folds the local accumulator copies and returns the sums over all
    ranks with ONE kernel (no NCCL call; a 2 KB all-reduce is pure latency and a step does 32 of them in a row)."""

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise RuntimeError("mmpde_bn_exchange is built for at most 16 ranks")
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(BN_EXCHANGE_BYTES // 8, dtype=torch.float64, device=dev)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.peer_base = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        dist.barrier(group)                           # everybody's buffer is zeroed before anybody stores into it

    def sum(self, spread):
        out = torch.empty(2 * ops.H, dtype=torch.float64, device=spread.device)
        n_rep = spread.shape[0] if spread.dim() == 2 else 1
        ops._cabi.call("mmpde_bn_exchange", ops._ptr(spread), n_rep, ops._ptr(self.peer_base), self.rank, self.world,
                       ops._ptr(out), ops._stream())
        return out


class DistComm(ops._Comm):
    """Cross-rank sums of the BatchNorm statistics.  With peer memory every solver BRANCH of the step (the two solvers run
    concurrently on two streams, train_helper_2d._forward_gnn) gets its own exchange buffer and sequence: within a
    branch all ranks issue the same exchanges in the same order, while the two branches interleave differently on
    different GPUs -- one shared sequence would pair up the wrong exchanges.  (A spin-waiting exchange kernel is one
    small CTA; the persistent kernels of the other branch are ordinary grids without inter-CTA dependencies, so they make
    progress around it.)  Without peer memory the sums go through NCCL, whose collectives must be issued in one global
    order: then the branches are not overlapped (n_branches = 1)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.peer = None
        self.peers = []
        if self.world > 1 and torch.cuda.is_available() and os.environ.get("MMPDE_PEER_BN", "1") != "0":
            try:
                self.peers = [PeerExchange(group), PeerExchange(group)]
                self.peer = self.peers[0]
            except Exception as e:                    # no peer access / no symmetric memory: NCCL all-reduce instead
                self.peers, self.peer = [], None
                if dist.get_rank(group) == 0:
                    print(f"[mmpde_b200.dist] peer-memory BatchNorm exchange unavailable ({type(e).__name__}: {e}); "
                          "using NCCL all-reduce", flush=True)
        self.n_branches = len(self.peers) if (self.peers and os.environ.get("MMPDE_BRANCH_EXCHANGE", "1") != "0") else 1

    def reduce_bn_sums(self, spread, branch=0):
        """[n_rep, 256] local accumulator copies -> [256] sums over all ranks."""
        if self.peers and spread.is_cuda:
            return self.peers[branch if branch < self.n_branches else 0].sum(spread)
        return self.allreduce_(spread.sum(0) if spread.dim() == 2 else spread)

    def peer_args(self, branch=0):
        if not self.peers:
            return None                         # NCCL: fold / all-reduce / finalise as separate launches
        p = self.peers[branch if branch < self.n_branches else 0]
        return (ops._ptr(p.peer_base), p.rank, p.world)

    def allreduce_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def global_rows(self, n):
        if self.total_rows is not None:         # partitioned mesh: rows of the whole mesh
            return float(self.total_rows)
        return float(n) * self.world            # equal shards (the sharder below guarantees it)


def init_from_env(backend=None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    device = torch.device(f"cuda:{local}") if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": device} if use_cuda else {}
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world, **kw)
    if world > 1:
        ops.COMM = DistComm()
    return rank, world, device


def shutdown():
    ops.COMM = ops._Comm()
    if dist.is_initialized():
        dist.destroy_process_group()


def shard_batch(t, rank, world):
    """Equal contiguous shard of the leading (trajectory) axis; the global batch must divide evenly."""
    if t.shape[0] % world:
        raise ValueError(f"global batch {t.shape[0]} does not divide over {world} ranks")
    per = t.shape[0] // world
    return t[rank * per:(rank + 1) * per]


class GradBucket:
    """Flat fp32 bucket over all trainable parameters: one all-reduce per step."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.nbytes = n * 4

    def allreduce(self, average=True):
        """Gradients -> flat bucket -> ONE all-reduce -> back.  The two copies are multi-tensor launches
        (torch._foreach_copy_): one tiny kernel per parameter costs ~0.7 ms per step for the ~180 tensors here."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        have = [(p, v) for p, v in zip(self.params, self.views) if p.grad is not None]
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for _, v in have], [p.grad for p, _ in have])
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if average:
                self.flat.div_(world)
        if have:
            torch._foreach_copy_([p.grad for p, _ in have], [v for _, v in have])
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()


class HaloExchange:
    """Per-layer halo exchange of a graph-partitioned mesh, one part per rank (partition.py): the Q' half
    (columns 128..255) of the rows a peer needs is packed (mmpde_rows_gather), moved with ONE all-to-all-v over
    NCCL / NVLink, and lands in the contiguous halo rows of the receiver; the backward sends dL/dQ' of the halo
    rows home and adds it there (mmpde_rows_scatter_add)."""

    def __init__(self, plan, group=None):
        self.plan, self.group = plan, group
        self.send_idx = plan.send_idx.to(torch.int32).contiguous()
        self.n_send, self.n_halo, self.n_own = int(plan.send_idx.numel()), plan.n_halo, plan.n_own
        self.bytes_forward = self.n_halo * ops.H * 4

    def _pack(self, buf):                    # rows the peers need, Q' half -> contiguous [n_send,128]
        out = torch.empty(self.n_send, ops.H, dtype=torch.float32, device=buf.device)
        ops._cabi.call("mmpde_rows_gather", ops._ptr(buf, ops.H), 2 * ops.H, ops._ptr(self.send_idx), self.n_send, ops.H,
                       ops._ptr(out), ops._stream())
        return out

    def _unpack_add(self, recv, buf):        # rows that came home: add onto the owners' rows
        ops._cabi.call("mmpde_rows_scatter_add", ops._ptr(recv), ops._ptr(self.send_idx), self.n_send, ops.H,
                       ops._ptr(buf, ops.H), 2 * ops.H, ops._stream())

    def forward(self, bufs):
        (buf,) = bufs
        send = self._pack(buf)
        recv = torch.empty(self.n_halo, ops.H, dtype=torch.float32, device=buf.device)
        dist.all_to_all_single(recv, send, self.plan.recv_splits, self.plan.send_splits, group=self.group)
        buf[self.n_own:, ops.H:].copy_(recv)

    def backward(self, bufs):
        (buf,) = bufs
        send = buf[self.n_own:, ops.H:].contiguous()
        recv = torch.empty(self.n_send, ops.H, dtype=torch.float32, device=buf.device)
        dist.all_to_all_single(recv, send, self.plan.send_splits, self.plan.recv_splits, group=self.group)
        self._unpack_add(recv, buf)
