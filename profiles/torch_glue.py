"""Non-library GPU time of one training step (torch kernels around the C-ABI calls) via torch.profiler."""
import os, random, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mmpde_b200 import synthetic
from mmpde_b200.PDEs import burgers
from mmpde_b200.data_creator_2d import GraphCreator_FS_2D
from mmpde_b200.gnn_2d import MP_PDE_Solver_2D
from mmpde_b200.interpolate import ItpNet
from mmpde_b200.mmpde import criterion
from mmpde_b200.train_helper_2d import training_loop_branch
dev = torch.device("cuda:0")
torch.manual_seed(0); random.seed(0)
pde = burgers(); pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = bench.RES
gc = GraphCreator_FS_2D(pde, bench.K_NEIGH, "knn", 1, bench.RES[0])
model, model_b = MP_PDE_Solver_2D(pde).to(dev), MP_PDE_Solver_2D(pde).to(dev)
net = ItpNet(bench.RES[1], bench.RES[2], [128, 64], [128, 64], [1, 4, 16, 4, 1]).to(dev)
mover = synthetic.AnalyticMover().to(dev)
opt = torch.optim.AdamW([{"params": model.parameters()}, {"params": model_b.parameters()}, {"params": net.parameters()}], lr=2e-3)
fields = synthetic.burgers_fields(bench.BATCH, *bench.RES, seed=100).to(dev)
step = lambda: training_loop_branch(model, model_b, net, mover, [0], bench.BATCH, opt, None, [(fields, fields)], gc, criterion, dev)
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(((e.device_time_total / 2e3, e.count // 2, e.key) for e in ev if e.device_time_total > 0 and "mmpde" not in e.key), reverse=True)
tot_k = sum(e.device_time_total for e in ev if e.device_type.name == "CUDA") / 2e3 if hasattr(ev[0], "device_type") else 0
print("top non-mmpde device time per step (ms, count, name):")
for ms, n, k in rows[:28]:
    print(f"  {ms:7.3f} x{n:4d}  {k[:110]}")
cpu = sorted(((e.self_cpu_time_total / 2e3, e.count // 2, e.key) for e in ev), reverse=True)[:12]
print("top self CPU time per step (ms):")
for ms, n, k in cpu:
    print(f"  {ms:7.3f} x{n:4d}  {k[:110]}")
