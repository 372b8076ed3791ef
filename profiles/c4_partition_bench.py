"""BASELINE.json config 4: synthetic 1M-node unstructured mesh, 6 MP layers, hidden 128, k = 35, graph-partitioned
with one halo exchange per layer.  Launch:  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1
profiles/c4_partition_bench.py [--nodes 1000000] [--steps 3]   (G = 1 runs the unpartitioned processor).
Prints one JSON line: edge-updates/s (fwd+bwd) and the halo rows / bytes per layer.  Equality with the unpartitioned
processor is what tests/test_gpu_path.py::test_partitioned_solver_equals_whole_graph checks."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmpde_b200 import dist as mdist, ops, partition as pt  # noqa: E402
from mmpde_b200.gnn_2d import MP_PDE_Solver_2D  # noqa: E402
from mmpde_b200.PDEs import burgers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1000000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-reorder", dest="no_reorder", action="store_true", help="keep the row-major node numbering")
    a = ap.parse_args()
    rank, world, dev = mdist.init_from_env()
    side = int(round(a.nodes ** 0.5))
    n = side * side
    rng = np.random.default_rng(0)                         # identical mesh on every rank
    g = np.stack(np.meshgrid(np.linspace(0, 1, side), np.linspace(0, 1, side), indexing="ij"), -1).reshape(-1, 2)
    xy = torch.from_numpy((g + rng.uniform(-0.3, 0.3, g.shape) / (side - 1)).astype(np.float32)).to(dev)
    if not a.no_reorder:
        xy = xy[pt.morton_order(xy)].contiguous()          # Z-order node numbering: neighbour rows stay close in memory
    edges = ops.EdgeList.from_knn(ops.knn_indices_grid(xy, xy, 35, 0, True), has_pad=False)
    torch.manual_seed(0)
    u = torch.randn(n, 1, device=dev)
    pos = torch.cat((torch.full((n, 1), 7.0, device=dev), xy), 1)
    r = torch.randn(n, 1, device=dev)
    model = MP_PDE_Solver_2D(burgers()).to(dev)
    model.train()
    params = [p for p in model.parameters()]
    bucket = mdist.GradBucket(params) if world > 1 else None
    if world > 1:
        (part,), (plan,) = pt.split_graph(u, pos, edges.src, edges.dst, world, ranks=[rank])
        exch = mdist.HaloExchange(plan)
        halo_rows, own_rows, my_edges = plan.n_halo, plan.n_own, int(plan.src.numel())
        del edges
    else:
        class G:
            pass
        whole = G()
        whole.x, whole.pos, whole.edge_index, whole.batch, whole._edges = u, pos, None, None, edges
        halo_rows, own_rows, my_edges = 0, n, edges.n_edges

    def step():
        model.zero_grad(set_to_none=True)
        if world > 1:
            (out,) = model.forward_partitioned([part], exch)
            loss = (out * r[plan.owned]).sum() / n
        else:
            out = model(whole)
            loss = (out * r).sum() / n
        loss.backward()
        if bucket is not None:
            bucket.allreduce(average=False)             # every rank holds the partial sums of its own nodes / edges
        return out, loss

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out, loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    stats = torch.tensor([float(halo_rows), float(own_rows), float(my_edges)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
    else:
        gathered = [stats]
    if rank == 0:
        E = n * 35
        line = {"config": "C4 synthetic mesh, graph-partitioned, halo exchange per layer", "nodes": n, "edges": E, "layers": 6,
                "node_order": "row-major" if a.no_reorder else "morton", "n_gpus": world, "ms_per_step": float(ms), "edge_updates_per_s": E * 6 / (float(ms) * 1e-3),
                "halo_rows_per_rank": [int(s[0]) for s in gathered], "owned_rows_per_rank": [int(s[1]) for s in gathered],
                "halo_bytes_per_layer_per_direction_max": int(max(s[0] for s in gathered)) * 128 * 4,
                "loss": float(loss) if world == 1 else None}
        print(json.dumps(line))
    if world > 1:
        mdist.shutdown()


if __name__ == "__main__":
    main()
