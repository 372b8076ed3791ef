"""Run under torchrun with >= 2 GPUs (tests/test_gpu_multi.py does): mmpde_bn_exchange (NVLink peer-memory sum of the
BatchNorm column sums, one kernel) against an NCCL all-reduce -- eager, many times in a row (slot reuse), and replayed
from a CUDA graph (the sequence number lives on the device)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mmpde_b200 import dist as mdist, ops  # noqa: E402


def main():
    rank, world, dev = mdist.init_from_env()
    comm = ops.COMM
    assert isinstance(comm, mdist.DistComm) and comm.peer is not None, "peer-memory exchange was not set up"
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    # eager, 37 exchanges back to back with different data
    for it in range(37):
        spread = torch.randn(16, 256, generator=g, dtype=torch.float64).to(dev) * (1 + it)
        want = spread.sum(0)
        dist.all_reduce(want)
        got = comm.reduce_bn_sums(spread)
        torch.cuda.synchronize()
        assert torch.allclose(got, want, rtol=1e-13, atol=1e-9), (rank, it, float((got - want).abs().max()))
        gathered = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        assert all(torch.equal(gathered[0], t) for t in gathered), "ranks must hold identical bits"
    # replayed from a CUDA graph: static input, three exchanges per replay
    static = torch.zeros(16, 256, dtype=torch.float64, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        comm.reduce_bn_sums(static)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        a = comm.reduce_bn_sums(static)
        b = comm.reduce_bn_sums(static * 2)
        c = comm.reduce_bn_sums(a.unsqueeze(0) + b.unsqueeze(0))
    for it in range(9):
        x = torch.randn(16, 256, generator=g, dtype=torch.float64).to(dev)
        static.copy_(x)
        graph.replay()
        torch.cuda.synchronize()
        want = x.sum(0)
        dist.all_reduce(want)
        assert torch.allclose(a, want, rtol=1e-13, atol=1e-9) and torch.allclose(b, 2 * want, rtol=1e-13, atol=1e-9)
        assert torch.allclose(c, 3 * want * world, rtol=1e-13, atol=1e-8)
    del graph
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("PEER_BN_OK", flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
