"""Generate the golden fixtures under tests/golden/ by running the REFERENCE'S OWN, UNMODIFIED
Python sources (/root/reference/{gnn_2d,interpolate,data_creator_2d,train_helper_2d,PDEs}.py and
mesh/dmm_model.py) in this container.

The reference cannot be imported as-is: torch_geometric / torch_cluster / IPython / h5py are not
installed.  This script therefore installs thin shims of exactly the
third-party calls the hot path makes (semantics recorded in SURVEY.md section 2.3) into
``sys.modules`` and then imports the reference files from where they lie.  Everything the reference
itself spells out -- concat order, sign conventions, layer wiring, the interpolation driver, the
step loops, sklearn's NearestNeighbors (the real library) -- runs as the reference wrote it.

Run (only in the build container, /root/reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs: tests/golden/*.pt (committed).  tests/test_oracle_golden.py checks the oracle against them.
"""
import inspect
import os
import random
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import knn as oracle_knn  # noqa: E402  (integer rule; the reference leaves it to torch_cluster)
from tests.golden.common import SmoothMover, fill_params, synth_fields as _fields  # noqa: E402


# ------------------------------------------------------------------------------------------------
# shims of the un-vendored third-party operators
# ------------------------------------------------------------------------------------------------
class _MessagePassing(nn.Module):
    """torch_geometric.nn.MessagePassing 2.0.3 as used by gnn_2d.py:36,55 (aggr='mean', node_dim=-2,
    flow source_to_target): gather *_i by edge_index[1], *_j by edge_index[0]; message(); scatter-mean
    onto edge_index[1] (count clamped to >= 1); update(aggregated, remaining kwargs)."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        assert aggr == "mean" and node_dim == -2 and flow == "source_to_target"

    def propagate(self, edge_index, size=None, **kwargs):
        lifted = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_i"):
                lifted[name] = kwargs[name[:-2]].index_select(0, edge_index[1])
            elif name.endswith("_j"):
                lifted[name] = kwargs[name[:-2]].index_select(0, edge_index[0])
            else:
                lifted[name] = kwargs[name]
        msg = self.message(**lifted)
        n = kwargs["x"].shape[0]
        total = torch.zeros(n, msg.shape[1], dtype=msg.dtype).scatter_add_(
            0, edge_index[1][:, None].expand_as(msg), msg)
        count = torch.zeros(n, dtype=msg.dtype).scatter_add_(0, edge_index[1], torch.ones(len(edge_index[1])))
        count[count < 1] = 1
        agg = total / count[:, None]
        extra = [p for p in list(inspect.signature(self.update).parameters)[1:] if p in kwargs]
        return self.update(agg, **{p: kwargs[p] for p in extra})


class _PyGBatchNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)


class _Data:
    def __init__(self, x=None, edge_index=None, **kw):
        self.x, self.edge_index = x, edge_index
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


def _install_shims():
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_data = types.ModuleType("torch_geometric.data")
    tg_nn.MessagePassing = _MessagePassing
    tg_nn.BatchNorm = _PyGBatchNorm
    for unused in ("global_mean_pool", "InstanceNorm", "avg_pool_x"):
        setattr(tg_nn, unused, None)
    tg_data.Data = _Data
    tg.nn, tg.data = tg_nn, tg_data
    tc = types.ModuleType("torch_cluster")
    tc.knn_graph = lambda x, k, batch=None, loop=False: oracle_knn.knn_graph(x, k, batch, loop)
    tc.radius_graph = lambda x, r, batch=None, loop=False, max_num_neighbors=32: \
        oracle_knn.radius_graph(x, float(r), batch, loop, max_num_neighbors)
    ip = types.ModuleType("IPython")
    ip.embed = lambda *a, **k: None
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn, "torch_geometric.data": tg_data,
                        "torch_cluster": tc, "IPython": ip, "h5py": types.ModuleType("h5py")})


def _import_reference():
    _install_shims()
    sys.path.insert(0, REF)
    import PDEs as ref_pdes
    import gnn_2d as ref_gnn
    import interpolate as ref_itp
    import data_creator_2d as ref_dc
    import train_helper_2d as ref_th
    # mesh/dmm_model.py hard-codes device="cuda" in DenseNet.__init__ (:27-28); those two tensors are
    # never used in forward, so dropping the device kwarg while the class is built changes nothing.
    real_tensor, real_arange = torch.tensor, torch.arange
    strip = lambda f: (lambda *a, **k: f(*a, **{kk: vv for kk, vv in k.items() if kk != "device"}))
    sys.path.insert(0, os.path.join(REF, "mesh"))
    import dmm_model as ref_dmm
    ref_dmm._strip = (strip(real_tensor), strip(real_arange), real_tensor, real_arange)
    return ref_pdes, ref_gnn, ref_itp, ref_dc, ref_th, ref_dmm


def _build_dmm(ref_dmm, **kw):
    st, sa, rt, ra = ref_dmm._strip
    torch.tensor, torch.arange = st, sa
    try:
        return ref_dmm.DMM(**kw)
    finally:
        torch.tensor, torch.arange = rt, ra


def _seed(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


def _sd(m):
    return {k: v.clone() for k, v in m.state_dict().items()}


def _bn_stats(m):
    return {k: v.clone() for k, v in m.state_dict().items() if "running_" in k or "num_batches" in k}


def _grads(m, full_prefixes):
    """per-parameter grad norms for all, full grads for the named prefixes (keeps fixtures small)."""
    norms = {k: p.grad.norm().clone() for k, p in m.named_parameters() if p.grad is not None}
    full = {k: p.grad.clone() for k, p in m.named_parameters()
            if p.grad is not None and any(k.startswith(pre) for pre in full_prefixes)}
    return {"norms": norms, "full": full}


def _checksum(m):
    return {k: torch.stack([v.double().sum(), v.double().abs().sum()]) for k, v in m.state_dict().items()
            if v.dtype.is_floating_point}


def main():
    ref_pdes, ref_gnn, ref_itp, ref_dc, ref_th, ref_dmm = _import_reference()
    crit = lambda x, y: torch.nn.MSELoss()(x, y)                    # mmpde.py:33-36

    # ---- G1: one processor layer, train-mode forward + backward --------------------------------
    _seed(1)
    layer = fill_params(ref_gnn.GNN_Layer_FS_2D(128, 128, 128, 1, 1), seed=11)
    n_per, B = 90, 2
    pos = torch.rand(B * n_per, 2)
    batch = torch.arange(B).repeat_interleave(n_per)
    ei = oracle_knn.knn_graph(pos, 35, batch)
    x = torch.randn(B * n_per, 128, requires_grad=True)
    u = torch.randn(B * n_per, 1, requires_grad=True)
    var = torch.rand(B, 1).repeat_interleave(n_per, 0)
    r = torch.randn(B * n_per, 128)
    out = layer(x, u, pos[:, 0:1], pos[:, 1:2], var, ei, batch)
    (out * r).sum().backward()
    torch.save({"seed": 11, "x": x.detach(), "u": u.detach(), "pos": pos, "var": var, "edge_index": ei, "r": r,
                "out": out.detach(), "gx": x.grad, "gu": u.grad,
                "gparams": {k: p.grad.clone() for k, p in layer.named_parameters()},
                "bn_after": _bn_stats(layer)}, os.path.join(HERE, "g1_layer.pt"))

    # ---- G2: full solver (6 layers) on a 12x12 Burgers-like grid --------------------------------
    _seed(2)
    pde = ref_pdes.burgers()
    res = [31, 12, 12]
    pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = res
    gc = ref_dc.GraphCreator_FS_2D(pde, neighbors=35, connect_edge="knn", time_window=1, t_resolution=31)
    model = fill_params(ref_gnn.MP_PDE_Solver_2D(pde, time_window=1), seed=22)
    fields = _fields(3, 31, 12, 12, seed=20)
    steps = [4, 17, 30]
    data, labels = gc.create_data(fields, steps)
    graph = gc.create_graph(None, data, labels, steps, "cpu", None)
    model.train()
    pred = model(graph)
    loss = crit(pred, labels.reshape(-1, 1))
    loss.backward()
    grads = _grads(model, ["embedding_mlp", "output_mlp", "gnn_layers.0.", "gnn_layers.5.norm", "gnn_layers.5.message_net_1"])
    bn1 = _bn_stats(model)
    model.eval()
    with torch.no_grad():
        pred_eval = model(graph)
    torch.save({"seed": 22, "res": res, "fields_seed": 20, "steps": steps, "data": data, "labels": labels,
                "graph_x": graph.x, "graph_y": graph.y, "graph_pos": graph.pos, "graph_batch": graph.batch,
                "edge_index": graph.edge_index, "pred_train": pred.detach(), "loss": loss.detach(),
                "grads": grads, "bn_after": bn1, "pred_eval": pred_eval}, os.path.join(HERE, "g2_solver.pt"))

    # ---- G3: ItpNet, all three modes ------------------------------------------------------------
    _seed(3)
    itp = fill_params(ref_itp.ItpNet(12, 12, [128, 64], [128, 64], [1, 4, 16, 4, 1]), seed=33)
    nb = torch.rand(2, 50, 30, 2)
    q = torch.rand(2, 50, 1, 2)
    img = torch.randn(2, 1, 12, 12)
    itp_cy = fill_params(ref_itp.ItpNet(77, None, [128, 64], [128, 64], [1, 4, 16, 4, 1]), seed=34)
    vec = torch.randn(2, 77)
    with torch.no_grad():
        torch.save({"seed": 33, "seed_cy": 34, "keys": sorted(itp.state_dict().keys()),
                    "keys_cy": sorted(itp_cy.state_dict().keys()),
                    "nb": nb, "q": q, "img": img, "w1": itp(nb, q, "1"), "w2": itp(nb, q, "2"),
                    "res": itp(None, None, "res_cut", img), "vec": vec,
                    "res_cy": itp_cy(None, None, "res_cut", vec)}, os.path.join(HERE, "g3_itpnet.pt"))

    # ---- G4: graph creator in MM mode (moved mesh, interpolation both ways), sklearn kNN ---------
    _seed(4)
    mover = SmoothMover()
    graph_m = gc.create_graph(itp, data, labels, steps, "cpu", mover)
    pred_m = torch.randn(graph_m.x.shape[0], 1)
    back = gc.interpolate_pred(itp, pred_m, graph_m, data, "cpu")
    torch.save({"itp_seed": 33, "data": data, "labels": labels, "steps": steps,
                "x": graph_m.x.detach(), "y": graph_m.y.detach(), "pos": graph_m.pos.detach(),
                "edge_index": graph_m.edge_index, "batch": graph_m.batch, "pred": pred_m,
                "pred_on_grid": back.detach()}, os.path.join(HERE, "g4_creator_mm.pt"))

    # ---- G5: MM-PDE training steps (2-layer solvers) + per-time-step test sweep -----------------
    _seed(5)
    model_a = fill_params(ref_gnn.MP_PDE_Solver_2D(pde, time_window=1, hidden_layer=2), seed=51)
    model_b = fill_params(ref_gnn.MP_PDE_Solver_2D(pde, time_window=1, hidden_layer=2), seed=52)
    itp5 = fill_params(ref_itp.ItpNet(12, 12, [128, 64], [128, 64], [1, 4, 16, 4, 1]), seed=53)
    opt = torch.optim.AdamW([{"params": model_a.parameters()}, {"params": model_b.parameters()},
                             {"params": itp5.parameters()}], lr=2e-3)
    fields5 = _fields(4, 31, 12, 12, seed=50)
    loader = [(fields5[:2], fields5[:2]), (fields5[2:], fields5[2:])]
    model_a.train(); model_b.train(); itp5.train()
    random.seed(55)
    tr_losses = ref_th.training_loop_branch(model_a, model_b, itp5, mover, [0], 2, opt, None, loader, gc, crit, "cpu")
    random.seed(56)
    itp_losses = ref_th.training_itp(itp5, mover, [0], 2, opt, None, loader, gc, crit, "cpu")
    model_a.eval(); model_b.eval(); itp5.eval()
    curve = [ref_th.test_timestep_losses(model_a, model_b, itp5, mover, [st], 2, loader, gc, crit, "cpu")
             for st in (1, 9, 30)]
    torch.save({"seeds": [51, 52, 53], "fields_seed": 50, "train_losses": tr_losses, "itp_losses": itp_losses,
                "curve_steps": [1, 9, 30], "curve": torch.stack(curve),
                "sum_a_after": _checksum(model_a), "sum_b_after": _checksum(model_b),
                "sum_itp_after": _checksum(itp5)}, os.path.join(HERE, "g5_mm_steps.pt"))

    # ---- G6: DMM, array and graph modes, and the moved mesh it yields ---------------------------
    _seed(6)
    dmm_a = fill_params(_build_dmm(ref_dmm, s=12, mode="array", branch_layer=7, trunk_layer=[2, 32, 512],
                                   out_layer=[1024, 512, 1]), seed=61)
    dmm_a.eval()
    u6 = fields[:, 0]
    mx, my = gc.moving_mesh(u6, dmm_a, 12, 12)
    pts = torch.rand(60, 2)
    dmm_g = fill_params(_build_dmm(ref_dmm, mode="graph", grid=pts, branch_layer=[4, 3], trunk_layer=[2, 16, 512],
                                   out_layer=[1024, 512, 1]), seed=62)
    dmm_g.eval()
    u6g = torch.randn(2, 60)
    xi = pts[None].repeat(2, 1, 1).reshape(-1, 2)
    with torch.no_grad():
        phi_g = dmm_g(u6g, xi)
    torch.save({"seed_array": 61, "seed_graph": 62, "keys_array": sorted(dmm_a.state_dict().keys()),
                "keys_graph": sorted(dmm_g.state_dict().keys()), "u": u6, "mesh_x": mx.detach(),
                "mesh_y": my.detach(), "pts": pts, "u_graph": u6g, "phi_graph": phi_g},
               os.path.join(HERE, "g6_dmm.pt"))

    # ---- G7: cylinder-style (2-D grid_size) creator + radius-graph connectivity -----------------
    _seed(7)
    n_cy = 77
    grid_cy = torch.rand(n_cy, 2)
    pde_cy = ref_pdes.cy(ori_grid=grid_cy)
    pde_cy.grid_size = pde_cy.movingmesh_grid_size = pde_cy.ori_grid_size = [30, n_cy]
    gc_cy = ref_dc.GraphCreator_FS_2D(pde_cy, neighbors=35, connect_edge="knn", time_window=1, t_resolution=30)
    f_cy = torch.randn(2, 30, n_cy)
    d_cy, l_cy = gc_cy.create_data(f_cy, [3, 11])
    g_cy = gc_cy.create_graph(itp_cy, d_cy, l_cy, [3, 11], "cpu", mover)
    p_cy = torch.randn(2 * n_cy, 1)
    b_cy = gc_cy.interpolate_pred(itp_cy, p_cy, g_cy, d_cy, "cpu")
    gc_r = ref_dc.GraphCreator_FS_2D(pde, neighbors=2, connect_edge="radius", time_window=1, t_resolution=31)
    g_r = gc_r.create_graph(None, data, labels, steps, "cpu", None)
    torch.save({"grid": grid_cy, "fields": f_cy, "itp_seed": 34, "x": g_cy.x.detach(),
                "pos": g_cy.pos.detach(), "edge_index": g_cy.edge_index, "pred": p_cy, "back": b_cy.detach(),
                "radius_edge_index": g_r.edge_index, "pde_dt": [pde.dt, pde_cy.dt]},
               os.path.join(HERE, "g7_cy_radius.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
