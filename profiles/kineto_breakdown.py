"""GPU-time breakdown of the bench step as it really runs (replayed from the CUDA graph): kernel durations from CUPTI
activity records through torch.profiler, summed per kernel name.  Unlike step_breakdown.py (events around eager calls)
this carries no per-launch event overhead and sees the torch glue kernels too.
usage: python profiles/kineto_breakdown.py [steps] [--eager] [--ops] [--shapes] [--cylinder] [--analytic-mover]"""
import collections
import os
import random
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mmpde_b200 import synthetic  # noqa: E402
from mmpde_b200.PDEs import burgers  # noqa: E402
from mmpde_b200.data_creator_2d import GraphCreator_FS_2D  # noqa: E402
from mmpde_b200.gnn_2d import MP_PDE_Solver_2D  # noqa: E402
from mmpde_b200.interpolate import ItpNet  # noqa: E402
from mmpde_b200.mmpde import criterion  # noqa: E402
from mmpde_b200.train_helper_2d import StepGraph, training_loop_branch  # noqa: E402


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    steps = int(argv[0]) if argv else 5
    eager = "--eager" in sys.argv
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    random.seed(0)
    from mmpde_b200.mesh.dmm_model import DMM
    cyl = "--cylinder" in sys.argv
    if cyl:                                      # BASELINE.json configs[2], set up like bench.py's workload("cylinder")
        from mmpde_b200.PDEs import cy
        cloud = synthetic.cylinder_cloud(bench.CY_RES[1], seed=0)
        pde = cy(ori_grid=cloud, device=dev)
        pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = bench.CY_RES
        gc = GraphCreator_FS_2D(pde, bench.K_NEIGH, "knn", 1, bench.CY_RES[0])
        net = ItpNet(bench.CY_RES[1], None, [128, 64], [128, 64], [1, 4, 16, 4, 1]).to(dev)
    else:
        pde = burgers()
        pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = bench.RES
        gc = GraphCreator_FS_2D(pde, bench.K_NEIGH, "knn", 1, bench.RES[0])
        net = ItpNet(bench.RES[1], bench.RES[2], [128, 64], [128, 64], [1, 4, 16, 4, 1]).to(dev)
    model, model_b = MP_PDE_Solver_2D(pde).to(dev), MP_PDE_Solver_2D(pde).to(dev)
    torch.manual_seed(bench.MOVER_SEED)
    if "--analytic-mover" in sys.argv:
        mover = synthetic.AnalyticMover()
    elif cyl:
        mover = DMM(mode="graph", grid=cloud.to(dev), **bench.DMM_GRAPH)
    else:
        mover = DMM(s=bench.RES[1], mode="array", **bench.DMM_ARRAY)
    mover = mover.to(dev).eval()
    opt = torch.optim.AdamW([{"params": model.parameters()}, {"params": model_b.parameters()}, {"params": net.parameters()}],
                            lr=2e-3, capturable=True, fused=True)
    fields = (synthetic.cylinder_fields(bench.BATCH, cloud, bench.CY_RES[0], seed=100) if cyl
              else synthetic.burgers_fields(bench.BATCH, *bench.RES, seed=100)).to(dev)
    sg = None if eager else StepGraph()

    def step():
        training_loop_branch(model, model_b, net, mover, [0], bench.BATCH, opt, None, [(fields, fields)], gc, criterion, dev,
                             step_graph=sg)

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes="--shapes" in sys.argv,
                 with_stack="--shapes" in sys.argv) as prof:
        t0.record()
        for _ in range(steps):
            step()
        t1.record()
        torch.cuda.synchronize()
    wall = t0.elapsed_time(t1) / steps
    tot, cnt = collections.Counter(), collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name
            for cut in ("<", "("):
                name = name.split(cut)[0]
            tot[name] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            cnt[name] += 1
    busy = sum(tot.values()) / steps / 1e3
    print(f"{'eager' if eager else 'graph replay'}: step {wall:.2f} ms under the profiler; kernels+copies busy {busy:.2f} ms/step")
    for name, v in tot.most_common(60):
        print(f"  {v / steps / 1e3:8.3f} ms {100 * v / steps / 1e3 / wall:5.1f}%  x{cnt[name] / steps:6.1f}  avg {v / cnt[name]:8.1f} us  {name[:90]}")


    if "--ops" in sys.argv:          # eager only: which aten ops the small torch kernels come from
        print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=60))


    if "--shapes" in sys.argv:       # eager only: torch ops by input shape and Python call site (which glue is worth fusing)
        rows = [e for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6) if e.key.startswith("aten::") and e.self_device_time_total > 0]
        rows.sort(key=lambda e: -e.self_device_time_total)
        for e in rows[:45]:
            site = next((fr for fr in e.stack if "mm-pde_b200" in fr or "mmpde_b200" in fr), e.stack[0] if e.stack else "")
            print(f"  {e.self_device_time_total / steps:8.1f} us/step x{e.count / steps:5.1f}  {e.key:28s} {str(e.input_shapes)[:70]:70s} {site[-70:]}")


if __name__ == "__main__":
    main()
