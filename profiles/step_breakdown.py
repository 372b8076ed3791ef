"""Where one training step of the bench workload spends its GPU time: every C-ABI call of the step is bracketed with
CUDA events on its stream and summed per entry point (warm caches, real launch order -- unlike the ncu launch list,
which is cold-cache and serialised).  usage: python profiles/step_breakdown.py [steps]"""
import collections
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mmpde_b200 import _cabi, synthetic  # noqa: E402
from mmpde_b200.PDEs import burgers  # noqa: E402
from mmpde_b200.data_creator_2d import GraphCreator_FS_2D  # noqa: E402
from mmpde_b200.gnn_2d import MP_PDE_Solver_2D  # noqa: E402
from mmpde_b200.interpolate import ItpNet  # noqa: E402
from mmpde_b200.mmpde import criterion  # noqa: E402
from mmpde_b200.train_helper_2d import training_loop_branch  # noqa: E402
import mmpde_b200.ops as ops_mod  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    random.seed(0)
    pde = burgers()
    pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = bench.RES
    gc = GraphCreator_FS_2D(pde, bench.K_NEIGH, "knn", 1, bench.RES[0])
    model, model_b = MP_PDE_Solver_2D(pde).to(dev), MP_PDE_Solver_2D(pde).to(dev)
    net = ItpNet(bench.RES[1], bench.RES[2], [128, 64], [128, 64], [1, 4, 16, 4, 1]).to(dev)
    mover = synthetic.AnalyticMover().to(dev)
    opt = torch.optim.AdamW([{"params": model.parameters()}, {"params": model_b.parameters()}, {"params": net.parameters()}], lr=2e-3)
    fields = synthetic.burgers_fields(bench.BATCH, *bench.RES, seed=100).to(dev)

    def step():
        training_loop_branch(model, model_b, net, mover, [0], bench.BATCH, opt, None, [(fields, fields)], gc, criterion, dev)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    # host cost of queueing one step (no sync inside) next to the GPU time of the same steps: if the two are close the
    # step is launch-bound and the kernels' speed no longer shows
    import time
    h0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(steps):
        step()
    g1.record()
    host_ms = (time.perf_counter() - h0) * 1e3 / steps
    torch.cuda.synchronize()
    print(f"host queueing {host_ms:.2f} ms/step, GPU {g0.elapsed_time(g1) / steps:.2f} ms/step")
    events = []
    real = _cabi.call

    def timed(name, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = real(name, *a)
        e.record()
        events.append((name, s, e))
        return rc

    ops_mod._cabi.call = timed
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    ops_mod._cabi.call = real
    tot, cnt = collections.Counter(), collections.Counter()
    for name, s, e in events:
        tot[name] += s.elapsed_time(e)
        cnt[name] += 1
    wall = t0.elapsed_time(t1) / steps
    ours = sum(tot.values()) / steps
    print(f"step {wall:.2f} ms (with event overhead); C-ABI kernels {ours:.2f} ms; the rest = torch glue (mesh mover, res_cut convs, "
          f"AdamW, elementwise) + launch gaps")
    for name, v in tot.most_common():
        print(f"  {v / steps:8.3f} ms {100 * v / steps / wall:5.1f}%  x{cnt[name] // steps:4d}  avg {1e3 * v / cnt[name]:7.1f} us  {name}")


if __name__ == "__main__":
    main()
