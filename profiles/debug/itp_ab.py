"""A/B of the tensor-core interpolation against the direct fp32 kernels inside one MM-mode training step (the g5 fixture's
12 x 12 setup): per-parameter gradients of the interpolation network and of both solvers, loss."""
import os, random, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mmpde_b200 import ops
from mmpde_b200.data_creator_2d import GraphCreator_FS_2D
from mmpde_b200.gnn_2d import MP_PDE_Solver_2D
from mmpde_b200.interpolate import ItpNet
from mmpde_b200.mmpde import criterion
from mmpde_b200.train_helper_2d import _forward_gnn, _sample_steps
from mmpde_b200.PDEs import burgers
from tests.golden.common import SmoothMover, fill_params, synth_fields

dev = torch.device("cuda:0")
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 12
pde = burgers()
pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = [31, nx, nx]
gc = GraphCreator_FS_2D(pde, 35, "knn", 1, 31)
model_a = fill_params(MP_PDE_Solver_2D(pde, time_window=1, hidden_layer=2), 51).to(dev)
model_b = fill_params(MP_PDE_Solver_2D(pde, time_window=1, hidden_layer=2), 52).to(dev)
net = fill_params(ItpNet(nx, nx, [128, 64], [128, 64], [1, 4, 16, 4, 1]), 53).to(dev)
fields = synth_fields(2, 31, nx, nx, seed=50)
mover = SmoothMover()
state = {k: v.clone() for m in (model_a, model_b) for k, v in m.state_dict().items()}
res = {}
for tc in (True, False):
    ops.ITP_TENSOR_CORES = tc
    for m in (model_a, model_b, net):
        m.train(); m.zero_grad(set_to_none=True)
    random.seed(55)
    steps = _sample_steps(gc, [0], 2)
    data, labels = gc.create_data(fields, steps)
    pred = _forward_gnn(model_a, model_b, net, mover, gc, data, labels, steps, dev)
    loss = criterion(pred, labels.to(dev).reshape(-1, 1))
    loss.backward()
    res[tc] = (float(loss), {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None},
               {f"{tag}.{n}": p.grad.detach().clone() for tag, m in (("a", model_a), ("b", model_b)) for n, p in m.named_parameters()})
print("loss tc / direct:", res[True][0], res[False][0])
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
for n in res[False][1]:
    print(f"  itp {n:22s} rel {rel(res[True][1][n], res[False][1][n]):.3e}   |g| {float(res[False][1][n].norm()):.3e}")
worst = max(((rel(res[True][2][n], res[False][2][n]), n) for n in res[False][2]))
print("  solvers: worst", worst)

# ---- the interpolation round trip of training_itp on identical weights: every intermediate, TC vs direct
from mmpde_b200.train_helper_2d import training_itp, training_loop_branch
from mmpde_b200._h2d import to_device
out = {}
for tc in (True, False):
    ops.ITP_TENSOR_CORES = tc
    random.seed(56)
    steps = _sample_steps(gc, [0], 2)
    data, labels = gc.create_data(fields, steps)
    with torch.no_grad():
        moved = gc.create_graph(net, data, labels, steps, dev, mover)
        back = gc.interpolate_pred(net, moved.x, moved, data, dev)
        loss = criterion(back, to_device(data, dev).reshape(-1, 1))
    out[tc] = (moved.x.clone(), moved.pos.clone(), back.clone(), float(loss))
print("round trip, same weights: loss tc / direct", out[True][3], out[False][3])
print("   moved.x rel", rel(out[True][0], out[False][0]), " pos equal", bool(torch.equal(out[True][1], out[False][1])), " back rel", rel(out[True][2], out[False][2]))

# ---- the fixture's flow: two AdamW steps of training_loop_branch, then training_itp, from the same initial weights
init = {id(m): {k: v.clone() for k, v in m.state_dict().items()} for m in (model_a, model_b, net)}
flow = {}
for tc in (True, False):
    ops.ITP_TENSOR_CORES = tc
    for m in (model_a, model_b, net):
        m.load_state_dict(init[id(m)]); m.train(); m.zero_grad(set_to_none=True)
    opt = torch.optim.AdamW([{"params": model_a.parameters()}, {"params": model_b.parameters()}, {"params": net.parameters()}], lr=2e-3)
    f4 = synth_fields(4, 31, nx, nx, seed=50)
    loader = [(f4[:2], f4[:2]), (f4[2:], f4[2:])]
    random.seed(55)
    tr = training_loop_branch(model_a, model_b, net, mover, [0], 2, opt, None, loader, gc, criterion, dev)
    w_after = {n: p.detach().clone() for n, p in net.named_parameters()}
    random.seed(56)
    it = training_itp(net, mover, [0], 2, opt, None, loader, gc, criterion, dev)
    flow[tc] = (tr.cpu(), it.cpu(), w_after)
print("flow: tr tc", flow[True][0].tolist(), " direct", flow[False][0].tolist())
print("flow: it tc", flow[True][1].tolist(), " direct", flow[False][1].tolist())
for n in flow[False][2]:
    d = (flow[True][2][n] - flow[False][2][n]).abs()
    print(f"   weights after tr {n:20s} max|dw| {float(d.max()):.2e}  frac moved the other way {float((d > 1e-3).float().mean()):.4f}")
