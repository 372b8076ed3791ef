"""world_size-2 gloo tests of the multi-GPU host logic (flat gradient bucket, sync-BN reductions,
batch sharding).  CPU only."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mmpde_b200 import dist as mdist, ops
    r, w, dev = mdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and isinstance(ops.COMM, mdist.DistComm)
    # sync-BN coupling: fp64 column sums add up, row count scales with the world size
    sums = torch.full((256,), float(rank + 1), dtype=torch.float64)
    ops.COMM.allreduce_(sums)
    assert torch.all(sums == 3.0) and ops.COMM.global_rows(10) == 20.0
    # flat-bucket gradient all-reduce == mean of per-rank grads; params without grad get zeros
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    extra = torch.nn.Parameter(torch.zeros(2))
    x = torch.arange(8, dtype=torch.float32).reshape(2, 4) + rank
    lin(x).sum().backward()
    local = [p.grad.clone() for p in lin.parameters()]
    bucket = mdist.GradBucket(list(lin.parameters()) + [extra])
    bucket.allreduce()
    gathered = [torch.zeros_like(local[0]) for _ in range(world)]
    dist.all_gather(gathered, local[0])
    assert torch.allclose(lin.weight.grad, sum(gathered) / world)
    assert extra.grad is not None and torch.all(extra.grad == 0)
    # sharding
    full = torch.arange(12).reshape(6, 2)
    assert torch.equal(mdist.shard_batch(full, rank, world), full[rank * 3:(rank + 1) * 3])
    with pytest.raises(ValueError):
        mdist.shard_batch(torch.zeros(5, 1), rank, world)
    mdist.shutdown()
    q.put(rank)


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]
