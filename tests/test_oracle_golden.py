"""Pins the oracle (oracle/) against fixtures produced by the reference's OWN unmodified sources
(tests/golden/make_golden.py).  CPU only; runs here and on the GPU box (no /root/reference needed)."""
import os
import random

import pytest
import torch

from oracle import creator, dmm, itp, loops, pdes, processor
from tests.golden.common import SmoothMover, fill_params, synth_fields


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _rel(a, b, atol=1e-6):
    """relative L2 error; differences below ``atol`` (analytically-zero grads, e.g. a bias feeding a
    BatchNorm) count as zero."""
    err = float((a - b).norm())
    return 0.0 if err < atol else err / float(b.norm().clamp_min(1e-30))


def _burgers12():
    pde = pdes.burgers()
    pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = [31, 12, 12]
    return pde


def test_g1_layer_forward_backward(golden_dir):
    g = _load(golden_dir, "g1_layer.pt")
    layer = fill_params(processor.GNN_Layer_FS_2D(128, 128, 128, 1, 1), g["seed"])
    x = g["x"].clone().requires_grad_(True)
    u = g["u"].clone().requires_grad_(True)
    out = layer(x, u, g["pos"][:, 0:1], g["pos"][:, 1:2], g["var"], g["edge_index"])
    (out * g["r"]).sum().backward()
    assert _rel(out.detach(), g["out"]) < 1e-5
    assert _rel(x.grad, g["gx"]) < 1e-4
    assert _rel(u.grad, g["gu"]) < 1e-4
    named = dict(layer.named_parameters())
    assert set(named) == set(g["gparams"])          # state-dict key parity with the reference
    for k, gr in g["gparams"].items():
        assert _rel(named[k].grad, gr) < 1e-4, k
    for k, v in g["bn_after"].items():
        assert torch.allclose(layer.state_dict()[k].float(), v.float(), rtol=1e-5, atol=1e-6), k


def test_g2_solver_graph_and_outputs(golden_dir):
    g = _load(golden_dir, "g2_solver.pt")
    pde = _burgers12()
    gc = creator.GraphCreator_FS_2D(pde, 35, "knn", 1, 31)
    fields = synth_fields(3, 31, 12, 12, seed=g["fields_seed"])
    data, labels = gc.create_data(fields, g["steps"])
    assert torch.equal(data, g["data"]) and torch.equal(labels, g["labels"])
    graph = gc.create_graph(None, data, labels, g["steps"], "cpu", None)
    assert torch.equal(graph.edge_index, g["edge_index"])
    assert torch.equal(graph.x, g["graph_x"]) and torch.equal(graph.y, g["graph_y"])
    assert torch.equal(graph.pos, g["graph_pos"]) and torch.equal(graph.batch, g["graph_batch"])
    model = fill_params(processor.MP_PDE_Solver_2D(pde, time_window=1), g["seed"])
    assert repr(model) == "GNN"
    model.train()
    pred = model(graph)
    loss = loops.criterion(pred, labels.reshape(-1, 1))
    loss.backward()
    assert _rel(pred.detach(), g["pred_train"]) < 1e-5
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-6 * max(1.0, abs(float(g["loss"])))
    named = dict(model.named_parameters())
    assert set(named) == set(g["grads"]["norms"])
    for k, nrm in g["grads"]["norms"].items():
        assert abs(float(named[k].grad.norm()) - float(nrm)) <= 2e-4 * float(nrm) + 1e-9, k
    for k, gr in g["grads"]["full"].items():
        assert _rel(named[k].grad, gr) < 2e-4, k
    for k, v in g["bn_after"].items():
        assert torch.allclose(model.state_dict()[k].float(), v.float(), rtol=1e-5, atol=1e-6), k
    model.eval()
    with torch.no_grad():
        assert _rel(model(graph), g["pred_eval"]) < 1e-5


def test_g3_itpnet_modes(golden_dir):
    g = _load(golden_dir, "g3_itpnet.pt")
    net = fill_params(itp.ItpNet(12, 12, [128, 64], [128, 64], [1, 4, 16, 4, 1]), g["seed"])
    assert sorted(net.state_dict().keys()) == g["keys"]
    with torch.no_grad():
        assert _rel(net(g["nb"], g["q"], "1"), g["w1"]) < 1e-5
        assert _rel(net(g["nb"], g["q"], "2"), g["w2"]) < 1e-5
        assert _rel(net(None, None, "res_cut", g["img"]), g["res"]) < 1e-5
    net_cy = fill_params(itp.ItpNet(77, None, [128, 64], [128, 64], [1, 4, 16, 4, 1]), g["seed_cy"])
    assert sorted(net_cy.state_dict().keys()) == g["keys_cy"]
    with torch.no_grad():
        assert _rel(net_cy(None, None, "res_cut", g["vec"]), g["res_cy"]) < 1e-5


@pytest.mark.parametrize("backend", ["sklearn", "rule"])
def test_g4_creator_moving_mesh(golden_dir, backend):
    g = _load(golden_dir, "g4_creator_mm.pt")
    pde = _burgers12()
    gc = creator.GraphCreator_FS_2D(pde, 35, "knn", 1, 31, knn_backend=backend)
    net = fill_params(itp.ItpNet(12, 12, [128, 64], [128, 64], [1, 4, 16, 4, 1]), g["itp_seed"])
    graph = gc.create_graph(net, g["data"], g["labels"], g["steps"], "cpu", SmoothMover())
    assert torch.equal(graph.edge_index, g["edge_index"])
    assert torch.equal(graph.batch, g["batch"])
    assert _rel(graph.pos.detach(), g["pos"]) < 1e-6
    assert _rel(graph.x.detach(), g["x"]) < 1e-5
    assert _rel(graph.y.detach(), g["y"]) < 1e-5
    back = gc.interpolate_pred(net, g["pred"], graph, g["data"], "cpu")
    assert _rel(back.detach(), g["pred_on_grid"]) < 1e-5


def test_g5_training_and_test_loops(golden_dir):
    g = _load(golden_dir, "g5_mm_steps.pt")
    pde = _burgers12()
    gc = creator.GraphCreator_FS_2D(pde, 35, "knn", 1, 31, knn_backend="rule")
    sa, sb, si = g["seeds"]
    model_a = fill_params(processor.MP_PDE_Solver_2D(pde, time_window=1, hidden_layer=2), sa)
    model_b = fill_params(processor.MP_PDE_Solver_2D(pde, time_window=1, hidden_layer=2), sb)
    net = fill_params(itp.ItpNet(12, 12, [128, 64], [128, 64], [1, 4, 16, 4, 1]), si)
    opt = torch.optim.AdamW([{"params": model_a.parameters()}, {"params": model_b.parameters()},
                             {"params": net.parameters()}], lr=2e-3)
    fields = synth_fields(4, 31, 12, 12, seed=g["fields_seed"])
    loader = [(fields[:2], fields[:2]), (fields[2:], fields[2:])]
    mover = SmoothMover()
    model_a.train(); model_b.train(); net.train()
    random.seed(55)
    tr = loops.training_loop_branch(model_a, model_b, net, mover, [0], 2, opt, None, loader, gc, loops.criterion)
    random.seed(56)
    it = loops.training_itp(net, mover, [0], 2, opt, None, loader, gc, loops.criterion)
    assert torch.allclose(tr, g["train_losses"], rtol=2e-4, atol=1e-7)
    assert torch.allclose(it, g["itp_losses"], rtol=2e-4, atol=1e-7)
    model_a.eval(); model_b.eval(); net.eval()
    curve = torch.stack([loops.test_timestep_losses(model_a, model_b, net, mover, [s], 2, loader, gc,
                                                    loops.criterion) for s in g["curve_steps"]])
    assert torch.allclose(curve, g["curve"], rtol=5e-4, atol=1e-7)
    for mod, key in ((model_a, "sum_a_after"), (model_b, "sum_b_after"), (net, "sum_itp_after")):
        sd = mod.state_dict()
        assert set(k for k, v in sd.items() if v.dtype.is_floating_point) == set(g[key])
        for k, cs in g[key].items():
            if k in ("embedding_mlp.0.bias", "embedding_mlp.3.bias"):
                continue   # bias feeding a BatchNorm: analytically zero grad, Adam amplifies round-off to +-lr
            got = torch.stack([sd[k].double().sum(), sd[k].double().abs().sum()])
            assert torch.allclose(got, cs, rtol=1e-3, atol=1e-4), k


def test_g6_dmm_both_modes(golden_dir):
    g = _load(golden_dir, "g6_dmm.pt")
    pde = _burgers12()
    gc = creator.GraphCreator_FS_2D(pde, 35, "knn", 1, 31)
    d_arr = fill_params(dmm.DMM(s=12, mode="array", branch_layer=7, trunk_layer=[2, 32, 512],
                                out_layer=[1024, 512, 1]), g["seed_array"]).eval()
    assert sorted(d_arr.state_dict().keys()) == g["keys_array"]
    mx, my = gc.moving_mesh(g["u"], d_arr, 12, 12)
    assert _rel(mx.detach(), g["mesh_x"]) < 1e-5 and _rel(my.detach(), g["mesh_y"]) < 1e-5
    d_gr = fill_params(dmm.DMM(mode="graph", grid=g["pts"], branch_layer=[4, 3], trunk_layer=[2, 16, 512],
                               out_layer=[1024, 512, 1]), g["seed_graph"]).eval()
    assert sorted(d_gr.state_dict().keys()) == g["keys_graph"]
    xi = g["pts"][None].repeat(2, 1, 1).reshape(-1, 2)
    with torch.no_grad():
        assert _rel(d_gr(g["u_graph"], xi), g["phi_graph"]) < 1e-5


def test_g7_cylinder_and_radius(golden_dir):
    g = _load(golden_dir, "g7_cy_radius.pt")
    pde_cy = pdes.cy(ori_grid=g["grid"])
    n = g["grid"].shape[0]
    pde_cy.grid_size = pde_cy.movingmesh_grid_size = pde_cy.ori_grid_size = [30, n]
    assert [pdes.burgers().dt, pde_cy.dt] == g["pde_dt"]
    gc = creator.GraphCreator_FS_2D(pde_cy, 35, "knn", 1, 30)
    net = fill_params(itp.ItpNet(n, None, [128, 64], [128, 64], [1, 4, 16, 4, 1]), g["itp_seed"])
    data, labels = gc.create_data(g["fields"], [3, 11])
    graph = gc.create_graph(net, data, labels, [3, 11], "cpu", SmoothMover())
    assert torch.equal(graph.edge_index, g["edge_index"])
    assert _rel(graph.pos.detach(), g["pos"]) < 1e-6 and torch.equal(graph.x, g["x"])
    back = gc.interpolate_pred(net, g["pred"], graph, data, "cpu")
    assert _rel(back.detach(), g["back"]) < 1e-5
    pde = _burgers12()
    gc_r = creator.GraphCreator_FS_2D(pde, 2, "radius", 1, 31)
    fields = synth_fields(3, 31, 12, 12, seed=20)
    d, l = gc_r.create_data(fields, [4, 17, 30])
    assert torch.equal(gc_r.create_graph(None, d, l, [4, 17, 30], "cpu", None).edge_index, g["radius_edge_index"])
