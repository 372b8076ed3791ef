"""Graph partitioning with a per-layer halo exchange for meshes that do not fit (or are too slow on) one GPU
(BASELINE.json config "Synthetic 1M-node unstructured mesh ... graph-partitioned with halo exchange").

The reference has no counterpart (single device); the contract is "partitioned result ==
unpartitioned result" for the processor of /root/reference/gnn_2d.py:119-141.

Owner-computes by TARGET node: rank p owns a spatially compact set of nodes and every edge whose target it
owns (the k-NN list of its own nodes), so the per-target mean needs no communication.  Sources owned by
another rank are the halo.  Because message_net_1 is split per node (z1_ij = P'[i] + Q'[j], edge features
folded into P'/Q', see csrc/edge_tc.cu), the ONLY per-layer traffic is the 128-float row Q'[j] of each halo
node forward and dL/dQ'[j] back; BatchNorm statistics go through the existing fp64 [2,128] all-reduce.

Local numbering of rank p: [0, n_own) owned nodes in ascending global id, then the halo grouped by owner rank
(ascending global id inside a group) so that what arrives from peer q lands in one contiguous slice.
Everything here is integer work on torch tensors (CPU or CUDA) and is testable without a GPU.
"""
from dataclasses import dataclass, field
from typing import List

import torch


def rcb_partition(pos, n_parts):
    """Recursive coordinate bisection: part id [N] (int64).  Splits the longer side of the bounding box at the
    weighted median so that parts get floor/ceil(N * share) nodes; any n_parts >= 1."""
    N = pos.shape[0]
    part = torch.zeros(N, dtype=torch.int64, device=pos.device)
    stack = [(torch.arange(N, device=pos.device), 0, n_parts)]
    while stack:
        ids, first, count = stack.pop()
        if count == 1:
            part[ids] = first
            continue
        left = count // 2
        p = pos[ids]
        ext = p.max(0).values - p.min(0).values
        axis = int(ext[1] > ext[0])
        # stable order by (coordinate, id): deterministic on every rank
        order = torch.argsort(p[:, axis], stable=True)
        n_left = (ids.numel() * left) // count
        stack.append((ids[order[:n_left]], first, left))
        stack.append((ids[order[n_left:]], first + left, count - left))
    return part


def morton_order(xy, bits=16):
    """Permutation that sorts points along a Z-order (Morton) curve of their bounding box.  Renumbering a large
    mesh this way keeps the neighbours of consecutive nodes close in memory, so the Q' row gathers of one edge
    tile hit L2 instead of HBM (at 1 M nodes the 512-byte rows of a row-major numbering are 0.5 MB apart).
    The graph is unchanged up to the relabelling; un-permute results with the inverse permutation."""
    lo, hi = xy.min(0).values, xy.max(0).values
    q = ((xy - lo) / (hi - lo).clamp_min(1e-30) * (2 ** bits - 1)).long().clamp_(0, 2 ** bits - 1)

    def spread(v):                                                # 16 bits -> every other bit of 32
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1)
    return torch.argsort(code, stable=True)


@dataclass
class PartPlan:
    rank: int
    n_parts: int
    n_total: int
    owned: torch.Tensor                 # [n_own] global ids, ascending
    halo: torch.Tensor                  # [n_halo] global ids, grouped by owner rank
    recv_splits: List[int]              # halo rows arriving from each rank (sum = n_halo)
    send_idx: torch.Tensor              # [n_send] LOCAL owned rows, grouped by destination rank
    send_splits: List[int]              # rows going to each rank
    src: torch.Tensor                   # [E_loc] int32 local source (owned or halo)
    dst: torch.Tensor                   # [E_loc] int32 local target (owned), sorted
    inv_deg: torch.Tensor               # [n_own] fp32
    extra: dict = field(default_factory=dict)

    @property
    def n_own(self):
        return int(self.owned.numel())

    @property
    def n_halo(self):
        return int(self.halo.numel())


def build_plans(part, edge_src, edge_dst, n_parts, ranks=None):
    """All ranks' plans (or those in ``ranks``) from the GLOBAL target-sorted edge list.  Deterministic: every
    rank computes the same table and keeps its own entry."""
    N = part.numel()
    dev = part.device
    edge_src, edge_dst = edge_src.long(), edge_dst.long()
    own_of_dst = part[edge_dst]
    owned, halos, halo_owner = [], [], []
    for p in range(n_parts):
        owned.append(torch.nonzero(part == p).flatten())
        e_src = edge_src[own_of_dst == p]
        remote = e_src[part[e_src] != p]
        h = torch.unique(remote)                                  # ascending global id
        o = part[h]
        order = torch.argsort(o, stable=True)                     # group by owner, ids stay ascending inside
        halos.append(h[order])
        halo_owner.append(o[order])
    plans = []
    for p in (range(n_parts) if ranks is None else ranks):
        g2l = torch.full((N,), -1, dtype=torch.int64, device=dev)
        n_own = owned[p].numel()
        g2l[owned[p]] = torch.arange(n_own, device=dev)
        g2l[halos[p]] = n_own + torch.arange(halos[p].numel(), device=dev)
        sel = own_of_dst == p
        src_l, dst_l = g2l[edge_src[sel]], g2l[edge_dst[sel]]
        assert bool((src_l >= 0).all()) and bool((dst_l >= 0).all())
        deg = torch.bincount(dst_l, minlength=n_own).clamp(min=1).to(torch.float32)
        recv_splits = [int((halo_owner[p] == q).sum()) for q in range(n_parts)]
        send, send_splits = [], []
        for q in range(n_parts):                                  # what rank q needs from me, in q's halo order
            need = halos[q][halo_owner[q] == p] if q != p else halos[q][:0]
            send.append(g2l[need])
            send_splits.append(int(need.numel()))
        send_idx = torch.cat(send) if send else torch.zeros(0, dtype=torch.int64, device=dev)
        assert bool((send_idx >= 0).all()) and bool((send_idx < n_own).all())
        plans.append(PartPlan(p, n_parts, N, owned[p], halos[p], recv_splits, send_idx, send_splits,
                              src_l.to(torch.int32).contiguous(), dst_l.to(torch.int32).contiguous(), 1.0 / deg))
    return plans


H = 128


class LocalExchange:
    """In-process stand-in for the NVLink all-to-all: all parts live in one process (single-GPU emulation of P
    ranks, CPU tests).  Operates on the per-part [n_own + n_halo, 256] buffers of one layer and only touches
    the Q' half (columns 128..255): ``forward`` fills every part's halo rows from the owners' rows,
    ``backward`` adds every part's halo rows back onto the owners' rows."""

    def __init__(self, plans):
        self.plans = plans
        self.bytes_forward = sum(p.n_halo for p in plans) * H * 4

    def _pairs(self):
        for p in self.plans:
            off = p.n_own
            for q, cnt in enumerate(p.recv_splits):
                if cnt:
                    pq = self.plans[q]
                    s0 = sum(pq.send_splits[:p.rank])
                    yield p.rank, off, cnt, q, pq.send_idx[s0:s0 + cnt]
                off += cnt

    def forward(self, bufs):
        for p, off, cnt, q, rows in self._pairs():
            if bufs[q].is_cuda:                                   # same pack kernel as the multi-GPU path
                from . import ops
                idx = rows.to(device=bufs[q].device, dtype=torch.int32).contiguous()
                tmp = torch.empty(cnt, H, dtype=torch.float32, device=bufs[q].device)
                ops._cabi.call("mmpde_rows_gather", ops._ptr(bufs[q], H), 2 * H, ops._ptr(idx), cnt, H, ops._ptr(tmp), ops._stream())
                bufs[p][off:off + cnt, H:].copy_(tmp)
            else:
                bufs[p][off:off + cnt, H:] = bufs[q][rows, H:]

    def backward(self, bufs):
        for p, off, cnt, q, rows in self._pairs():
            if bufs[q].is_cuda:
                from . import ops
                idx = rows.to(device=bufs[q].device, dtype=torch.int32).contiguous()
                tmp = bufs[p][off:off + cnt, H:].contiguous()
                ops._cabi.call("mmpde_rows_scatter_add", ops._ptr(tmp), ops._ptr(idx), cnt, H, ops._ptr(bufs[q], H), 2 * H, ops._stream())
            else:
                bufs[q][:, H:].index_add_(0, rows, bufs[p][off:off + cnt, H:])


class MeshPart:
    """What one rank holds of a partitioned graph: features / positions of its owned nodes, local edges, plan."""

    def __init__(self, x, pos, plan, edges):
        self.x, self.pos, self.plan, self.edges = x, pos, plan, edges


def split_graph(x, pos, edge_src, edge_dst, n_parts, part=None, ranks=None):
    """Partition a graph given by node features x [N,1], positions pos [N,3] = (t, x, y) and a target-sorted
    edge list into MeshParts (all parts, or only ``ranks``).  Returns (parts, plans)."""
    from . import ops
    if part is None:
        part = rcb_partition(pos[:, 1:3], n_parts)
    plans = build_plans(part, edge_src, edge_dst, n_parts, ranks)
    parts = [MeshPart(x[pl.owned].contiguous(), pos[pl.owned].contiguous(), pl,
                      ops.EdgeList(pl.src, pl.dst, pl.inv_deg, pl.n_own)) for pl in plans]
    return parts, plans
