"""Full-size parity on the configurations BASELINE.json quotes the numbers on ("Next round" 1c):

  C1/C2  Burgers 48x48, batch 16, 6 layers, MM mode (moved mesh + interpolation both ways + both solvers): one
         training step of the CUDA path against the oracle on the host cores -- edge sets bit-exact, prediction,
         loss, EVERY parameter gradient and the BatchNorm buffers within 1e-3 relative L2;
  C3     flow around a cylinder, n = 2 521 nodes per sample, MM mode;
  C5     k-NN + interpolation at 1 M points against scipy's cKDTree (the oracle's C rule is O(n^2)) and the
         oracle's ItpNet;  the graph k-NN (fp32 rule) at 1 M with a tie-aware validity check.

The oracle runs for tens of seconds per case on the box's host cores; the sizes are the real ones on purpose."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-3          # north-star: relative L2 <= 1e-3


def _rel(a, b, atol=0.0):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = float((a - b).norm())
    return 0.0 if err <= atol else err / float(b.norm().clamp_min(1e-30))


def _grad_report(model, omodel, tag, floor=1e-5):
    """Relative L2 of every parameter gradient.  Gradients that are analytically ZERO (a bias in front of a BatchNorm:
    the mean subtraction cancels it) are rounding noise on both sides -- recognised by a reference norm below `floor`
    times the largest gradient norm of the module, and compared absolutely against that scale instead."""
    worst, rows = 0.0, []
    on = dict(omodel.named_parameters())
    scale = max(float(q.grad.norm()) for q in on.values() if q.grad is not None)
    for k, p in model.named_parameters():
        ref = on[k].grad
        if ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        nrm = float(ref.norm())
        if nrm < floor * scale:
            assert float((p.grad.cpu() - ref).norm()) < 10 * floor * scale, (tag, k, nrm)
            continue
        r = _rel(p.grad, ref)
        rows.append((r, k))
        worst = max(worst, r)
    rows.sort(reverse=True)
    print(f"[{tag}] worst gradient rel-L2 {worst:.3e}; top: " + ", ".join(f"{k}={r:.2e}" for r, k in rows[:5]))
    return worst, rows


class HostMover(torch.nn.Module):
    """Evaluates the wrapped mesh mover (and, through autograd, its gradient) on the HOST whatever device the inputs live
    on, so that the CUDA path and the CPU oracle see bit-identical moved meshes.  The mover is outside the hot path (plain
    PyTorch on both sides); evaluated in fp32 on two different devices its sin/cos/tanh differ in the last bit, which on a
    barely moved lattice (displacement ~0.2 h) reorders nearly equidistant neighbours -- two equally valid graphs, but no
    longer comparable edge by edge."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, u, grid, rf=False):
        return self.inner(u.cpu(), grid.cpu()).to(grid.device)


def _copy_bn(model):
    return {k: v.detach().clone().cpu() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}


def test_c2_burgers_mm_step_full_size_vs_oracle():
    """BASELINE.json configs[0]/[1] at full size: 31x48x48, B = 16, N = 36 864, E = 1 290 240, two 6-layer solvers."""
    from mmpde_b200 import synthetic
    from mmpde_b200.PDEs import burgers
    from mmpde_b200.data_creator_2d import GraphCreator_FS_2D
    from mmpde_b200.gnn_2d import MP_PDE_Solver_2D
    from mmpde_b200.interpolate import ItpNet
    from oracle import creator as ocreator, itp as oitp, pdes as opdes, processor as oproc
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda:0")
    res = [31, 48, 48]
    B = 16
    pde, opde = burgers(), opdes.burgers()
    for p in (pde, opde):
        p.grid_size = p.movingmesh_grid_size = p.ori_grid_size = res
    fields = synthetic.burgers_fields(B, seed=0)
    steps = [1 + (7 * i) % 30 for i in range(B)]
    mover = HostMover(synthetic.AnalyticMover())
    torch.manual_seed(0)
    omodel, omodel_b = oproc.MP_PDE_Solver_2D(opde), oproc.MP_PDE_Solver_2D(opde)
    onet = oitp.ItpNet(48, 48, [128, 64], [128, 64], [1, 4, 16, 4, 1])
    model, model_b = MP_PDE_Solver_2D(pde), MP_PDE_Solver_2D(pde)
    net = ItpNet(48, 48, [128, 64], [128, 64], [1, 4, 16, 4, 1])
    model.load_state_dict(omodel.state_dict()); model_b.load_state_dict(omodel_b.state_dict())
    net.load_state_dict(onet.state_dict())
    model, model_b, net = model.to(dev).train(), model_b.to(dev).train(), net.to(dev).train()
    omodel.train(); omodel_b.train(); onet.train()
    gc = GraphCreator_FS_2D(pde, 35, "knn", 1, 31)
    ogc = ocreator.GraphCreator_FS_2D(opde, 35, "knn", 1, 31, knn_backend="rule")

    def step(gcx, m, mb, it, device):
        data, labels = gcx.create_data(fields, steps)
        moved = gcx.create_graph(it, data, labels, steps, device, mover)
        uniform = gcx.create_graph(it, data, labels, steps, device, None)
        pred = gcx.interpolate_pred(it, mb(moved), moved, data, device) + m(uniform)
        loss = torch.nn.functional.mse_loss(pred, labels.to(device).reshape(-1, 1))
        loss.backward()
        return pred.detach(), loss.detach(), moved, uniform

    pred, loss, moved, uniform = step(gc, model, model_b, net, dev)
    torch.cuda.synchronize()
    opred, oloss, omoved, ouniform = step(ogc, omodel, omodel_b, onet, "cpu")
    # integer work: bit-exact edge sets on the moved mesh and on the grid
    assert moved.edge_index.shape == (2, B * 2304 * 35)
    assert torch.equal(moved.edge_index.cpu(), omoved.edge_index)
    assert torch.equal(uniform.edge_index.cpu(), ouniform.edge_index)
    assert torch.equal(moved.pos.cpu(), omoved.pos) and _rel(moved.x, omoved.x) < 1e-4
    r_pred = _rel(pred, opred)
    print(f"[C2 full] pred rel-L2 {r_pred:.3e}  loss {float(loss):.7f} vs {float(oloss):.7f}")
    assert r_pred < TOL
    assert abs(float(loss) - float(oloss)) <= 1e-4 * abs(float(oloss))
    for tag, m, om in (("model", model, omodel), ("model_b", model_b, omodel_b), ("itp", net, onet)):
        worst, _ = _grad_report(m, om, f"C2 full / {tag}")
        assert worst < TOL, (tag, worst)
    for m, om in ((model, omodel), (model_b, omodel_b)):
        ref = _copy_bn(om)
        for k, v in _copy_bn(m).items():
            assert _rel(v.float(), ref[k].float()) < 1e-4, k


def test_c3_cylinder_mm_step_full_size_vs_oracle():
    """BASELINE.json configs[2] shape: 2 521 unstructured nodes per sample, cylinder branch of the graph creator
    (moving_mesh_tri, no re-interpolation of the input field, mode-'2' interpolation back, MLP res_cut), 6 layers."""
    from mmpde_b200 import synthetic
    from mmpde_b200.PDEs import cy
    from mmpde_b200.data_creator_2d import GraphCreator_FS_2D
    from mmpde_b200.gnn_2d import MP_PDE_Solver_2D
    from mmpde_b200.interpolate import ItpNet
    from oracle import creator as ocreator, itp as oitp, pdes as opdes, processor as oproc
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda:0")
    n, B = 2521, 16
    cloud = synthetic.cylinder_cloud(n, seed=0)
    pde, opde = cy(ori_grid=cloud, device=dev), opdes.cy(ori_grid=cloud)
    for p in (pde, opde):
        p.grid_size = p.movingmesh_grid_size = p.ori_grid_size = [30, n]
    fields = synthetic.cylinder_fields(B, cloud, 30, seed=3)
    steps = [1 + (5 * i) % 28 for i in range(B)]
    mover = HostMover(synthetic.AnalyticMover())
    torch.manual_seed(1)
    omodel, omodel_b = oproc.MP_PDE_Solver_2D(opde), oproc.MP_PDE_Solver_2D(opde)
    onet = oitp.ItpNet(n, None, [128, 64], [128, 64], [1, 4, 16, 4, 1])
    model, model_b = MP_PDE_Solver_2D(pde), MP_PDE_Solver_2D(pde)
    net = ItpNet(n, None, [128, 64], [128, 64], [1, 4, 16, 4, 1])
    model.load_state_dict(omodel.state_dict()); model_b.load_state_dict(omodel_b.state_dict())
    net.load_state_dict(onet.state_dict())
    model, model_b, net = model.to(dev).train(), model_b.to(dev).train(), net.to(dev).train()
    omodel.train(); omodel_b.train(); onet.train()
    gc = GraphCreator_FS_2D(pde, 35, "knn", 1, 30)
    ogc = ocreator.GraphCreator_FS_2D(opde, 35, "knn", 1, 30, knn_backend="rule")

    def step(gcx, m, mb, it, device):
        data, labels = gcx.create_data(fields, steps)
        moved = gcx.create_graph(it, data, labels, steps, device, mover)
        uniform = gcx.create_graph(it, data, labels, steps, device, None)
        pred = gcx.interpolate_pred(it, mb(moved), moved, data, device) + m(uniform)
        loss = torch.nn.functional.mse_loss(pred, labels.to(device).reshape(-1, 1))
        loss.backward()
        return pred.detach(), loss.detach(), moved, uniform

    pred, loss, moved, uniform = step(gc, model, model_b, net, dev)
    opred, oloss, omoved, ouniform = step(ogc, omodel, omodel_b, onet, "cpu")
    assert torch.equal(moved.edge_index.cpu(), omoved.edge_index)
    assert torch.equal(uniform.edge_index.cpu(), ouniform.edge_index)
    r_pred = _rel(pred, opred)
    print(f"[C3 full] pred rel-L2 {r_pred:.3e}  loss {float(loss):.7f} vs {float(oloss):.7f}")
    assert r_pred < TOL
    assert abs(float(loss) - float(oloss)) <= 1e-4 * abs(float(oloss))
    for tag, m, om in (("model", model, omodel), ("model_b", model_b, omodel_b), ("itp", net, onet)):
        worst, _ = _grad_report(m, om, f"C3 full / {tag}")
        assert worst < TOL, (tag, worst)


def _lattices(side, seed=0):
    from mmpde_b200 import synthetic
    moved = synthetic.jittered_lattice(side, 0.3, seed=seed)
    g = np.stack(np.meshgrid(np.linspace(0, 1, side), np.linspace(0, 1, side), indexing="ij"), -1).reshape(-1, 2)
    return moved, torch.tensor(g, dtype=torch.float32)


def test_c5_interpolation_1m_vs_ckdtree_and_oracle():
    """BASELINE.json configs[4] at 1 M: sources = jittered 1000x1000 lattice (the moved mesh), queries = the exact
    lattice (mode '2').  Ordered 30-NN lists against scipy's cKDTree on the same fp32 coordinates (fp64 distances, the
    interpolation rule); interpolated values against the oracle's ItpNet on the cKDTree lists."""
    from scipy.spatial import cKDTree
    from mmpde_b200 import ops
    from mmpde_b200.interpolate import ItpNet
    from oracle import itp as oitp
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda:0")
    side = 1000
    src, qry = _lattices(side)
    P = src.shape[0]
    idx = ops.knn_indices_grid(src.to(dev), qry.to(dev), 30, 1, False)
    torch.cuda.synchronize()
    tree = cKDTree(src.double().numpy())
    d_ref, i_ref = tree.query(qry.double().numpy(), k=30, workers=-1)
    got = idx.cpu().numpy().astype(np.int64)
    same = (got == i_ref)
    bad_rows = np.nonzero(~same.all(1))[0]
    # any disagreement must be an exact fp64 distance tie (two sources equally far from the query), resolved by index
    for r in bad_rows[:200]:
        dg = ((src[got[r]].double() - qry[r].double()) ** 2).sum(1).numpy()
        dr = ((src[i_ref[r]].double() - qry[r].double()) ** 2).sum(1).numpy()
        assert np.array_equal(np.sort(dg), np.sort(dr)), r
        assert np.all(np.diff(dg) >= 0), r
    print(f"[C5 1M] interpolation 30-NN rows equal to cKDTree: {1 - len(bad_rows) / P:.6f} ({len(bad_rows)} tie rows)")
    assert len(bad_rows) <= P // 10000
    torch.manual_seed(5)
    onet = oitp.ItpNet(8, 8, [128, 64], [128, 64], [1, 4, 16, 4, 1])
    net = ItpNet(8, 8, [128, 64], [128, 64], [1, 4, 16, 4, 1])
    net.load_state_dict(onet.state_dict())
    net = net.to(dev)
    vals = torch.randn(P)
    out = ops.InterpolateFn.apply(vals.to(dev), src.to(dev).contiguous(), qry.to(dev).contiguous(), idx, net.flat_params("2"))
    with torch.no_grad():
        ii = torch.from_numpy(got)
        ref = torch.empty(P)
        for a in range(0, P, 100000):                       # the oracle's formulation, chunked: [chunk,30,2] neighbours
            b = min(a + 100000, P)
            w = onet(src[ii[a:b]][None], qry[a:b][None].unsqueeze(-2), "2")[0]
            ref[a:b] = (w * vals[ii[a:b]]).sum(-1)
    r = _rel(out, ref)
    print(f"[C5 1M] interpolated values rel-L2 vs oracle {r:.3e}")
    assert r < 1e-5


def test_c4_graph_knn_1m_tie_aware():
    """Graph k-NN (k = 35, fp32 rule) on the 1 M-node jittered lattice of config C4: every neighbour list must be a valid
    35-NN set of its node (cKDTree in fp64 as the witness): same set, or differing only in members whose distance equals
    the 35th distance up to fp32 rounding of d^2; distances ascending; no self loops."""
    from scipy.spatial import cKDTree
    from mmpde_b200 import ops
    dev = torch.device("cuda:0")
    side = 1000
    xy, _ = _lattices(side, seed=0)
    n = xy.shape[0]
    nbr = ops.knn_indices_grid(xy.to(dev), xy.to(dev), 35, 0, True).cpu().numpy().astype(np.int64)
    assert nbr.shape == (n, 35) and (nbr >= 0).all() and (nbr != np.arange(n)[:, None]).all()
    tree = cKDTree(xy.double().numpy())
    d_ref, i_ref = tree.query(xy.double().numpy(), k=36, workers=-1)
    i_ref, d_ref = i_ref[:, 1:], d_ref[:, 1:]                # drop self (distance 0, jittered points are distinct)
    x64 = xy.double().numpy()
    # ascending in the frozen fp32 rule d2 = fmaf(dy, dy, dx*dx): differences in fp32, the fma emulated in fp64
    x32 = xy.numpy()
    dx = x32[nbr, 0] - x32[:, None, 0]
    dy = x32[nbr, 1] - x32[:, None, 1]
    d2_rule = (dy.astype(np.float64) * dy.astype(np.float64) + (dx * dx).astype(np.float64)).astype(np.float32)
    assert (np.diff(d2_rule, axis=1) >= 0).all()
    same_set = (np.sort(nbr, 1) == np.sort(i_ref, 1)).all(1)
    bad = np.nonzero(~same_set)[0]
    kth = d_ref[:, -1]
    for r in bad:
        extra = np.setdiff1d(nbr[r], i_ref[r])
        missing = np.setdiff1d(i_ref[r], nbr[r])
        de = np.sqrt(((x64[extra] - x64[r]) ** 2).sum(-1))
        dm = np.sqrt(((x64[missing] - x64[r]) ** 2).sum(-1))
        assert np.all(np.abs(de ** 2 - kth[r] ** 2) <= 1e-6 * kth[r] ** 2 + 1e-12), r
        assert np.all(np.abs(dm ** 2 - kth[r] ** 2) <= 1e-6 * kth[r] ** 2 + 1e-12), r
    print(f"[C4 1M] graph 35-NN sets equal to cKDTree: {same_set.mean():.6f} ({len(bad)} rows differ by fp32 near-ties)")
    assert len(bad) <= n // 1000
