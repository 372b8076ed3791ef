"""Graph partitioning + halo exchange plan (mm-pde_b200/partition.py): pure integer logic, checked on the CPU by
running all ranks' plans in one process ("simulating P ranks in one process") and, for the
torch.distributed all-to-all-v wiring, over world_size-2 gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmpde_b200 import partition as pt
from oracle import knn as oknn


def _graph(n, k, seed):
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(n)))
    g = np.stack(np.meshgrid(np.linspace(0, 1, side), np.linspace(0, 1, side), indexing="ij"), -1).reshape(-1, 2)[:n]
    pts = torch.from_numpy((g + rng.uniform(-0.3, 0.3, g.shape) / (side - 1)).astype(np.float32))
    ei = oknn.knn_graph(pts, k, torch.zeros(n, dtype=torch.long))
    return pts, ei[0], ei[1]


def _mean_agg(x, src, dst, n):
    out = torch.zeros(n, x.shape[1], dtype=x.dtype).index_add_(0, dst.long(), x[src.long()])
    deg = torch.bincount(dst.long(), minlength=n).clamp(min=1)
    return out / deg[:, None]


@pytest.mark.parametrize("n,k,P", [(400, 8, 2), (900, 35, 4), (333, 5, 3), (64, 3, 8)])
def test_partition_covers_graph_and_halo_exchange_is_exact(n, k, P):
    pts, src, dst = _graph(n, k, n + P)
    part = pt.rcb_partition(pts, P)
    sizes = torch.bincount(part, minlength=P)
    assert int(sizes.max() - sizes.min()) <= 1 + P                       # balanced bisection
    plans = pt.build_plans(part, src, dst, P)
    assert sum(p.n_own for p in plans) == n and sum(p.src.numel() for p in plans) == src.numel()
    assert torch.equal(torch.sort(torch.cat([p.owned for p in plans])).values, torch.arange(n))
    for p in plans:
        assert bool((part[p.halo] != p.rank).all()) and sum(p.recv_splits) == p.n_halo
        assert bool((p.dst[1:] >= p.dst[:-1]).all()) and int(p.dst.max()) < p.n_own       # target-sorted, owned targets
        assert all(plans[q].send_splits[p.rank] == p.recv_splits[q] for q in range(P))     # symmetric splits
    # forward exchange: every part sees the exact source rows -> identical per-target means
    torch.manual_seed(0)
    x = torch.randn(n, 128, dtype=torch.float64)
    ref = _mean_agg(x, src, dst, n)
    ex = pt.LocalExchange(plans)
    bufs = []
    for p in plans:
        b = torch.zeros(p.n_own + p.n_halo, 256, dtype=torch.float64)
        b[:p.n_own, 128:] = x[p.owned]
        bufs.append(b)
    ex.forward(bufs)
    for p, b in zip(plans, bufs):
        assert torch.equal(b[p.n_own:, 128:], x[p.halo])
        assert torch.allclose(_mean_agg(b[:, 128:], p.src, p.dst, p.n_own), ref[p.owned], rtol=0, atol=1e-12)
        assert float(b[:, :128].abs().max()) == 0.0                      # P' half untouched
    # backward exchange is the adjoint of the forward exchange: <F x, y> == <x, B y>
    ys = [torch.randn_like(b) for b in bufs]
    lhs = sum(float((b[p.n_own:, 128:] * y[p.n_own:, 128:]).sum()) for p, b, y in zip(plans, bufs, ys))
    back = [y.clone() for y in ys]
    for p, y in zip(plans, back):
        y[:p.n_own] = 0
    ex.backward(back)
    rhs = sum(float((x[p.owned] * y[:p.n_own, 128:]).sum()) for p, y in zip(plans, back))
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))


def test_single_part_has_no_halo():
    pts, src, dst = _graph(200, 6, 3)
    (plan,) = pt.build_plans(pt.rcb_partition(pts, 1), src, dst, 1)
    assert plan.n_halo == 0 and plan.n_own == 200 and plan.send_idx.numel() == 0
    assert torch.equal(plan.src.long(), src) and torch.equal(plan.dst.long(), dst)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mmpde_b200 import dist as mdist

    class CpuHalo(mdist.HaloExchange):           # host logic under test; the pack kernels are CUDA-only
        def _pack(self, buf):
            return buf[self.send_idx.long(), 128:].contiguous()

        def _unpack_add(self, recv, buf):
            buf[:, 128:].index_add_(0, self.send_idx.long(), recv)

    mdist.init_from_env(backend="gloo")
    pts, src, dst = _graph(500, 9, 11)
    part = pt.rcb_partition(pts, world)
    (plan,) = pt.build_plans(part, src, dst, world, ranks=[rank])
    x = torch.randn(500, 128, generator=torch.Generator().manual_seed(5))
    buf = torch.zeros(plan.n_own + plan.n_halo, 256)
    buf[:plan.n_own, 128:] = x[plan.owned]
    ex = CpuHalo(plan)
    ex.forward([buf])
    assert torch.equal(buf[plan.n_own:, 128:], x[plan.halo])
    # backward: every halo row goes home; the owner's row receives one copy per rank that holds it as halo
    g = torch.zeros_like(buf)
    g[plan.n_own:, 128:] = 1.0
    ex.backward([g])
    counts = torch.zeros(500)
    other = pt.build_plans(part, src, dst, world, ranks=[1 - rank])[0]
    counts[other.halo] += 1
    assert torch.equal(g[:plan.n_own, 128], counts[plan.owned])
    mdist.shutdown()
    q.put(rank)


def test_halo_exchange_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]


def test_morton_order_is_a_permutation_and_local():
    pts, _, _ = _graph(4096, 4, 1)
    perm = pt.morton_order(pts)
    assert torch.equal(torch.sort(perm).values, torch.arange(4096))
    # consecutive points along the curve are close: mean step far below the row-major mean step
    step_m = (pts[perm][1:] - pts[perm][:-1]).norm(dim=1).mean()
    assert float(step_m) < 0.05
