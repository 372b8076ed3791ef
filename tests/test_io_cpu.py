"""Dataset / checkpoint I/O and resume of the experiment driver (/root/reference/mmpde.py:163-173,
191-200, 292-310), and the analytic mesh-mover Jacobian -- all host-side logic, no GPU needed."""
import argparse
import os

import numpy as np
import torch

from mmpde_b200 import mmpde
from mmpde_b200.mesh.dmm_model import DMM


def _args(**kw):
    a = mmpde.build_parser().parse_args([])
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_dmm_analytic_displacement_equals_autograd():
    """DMM.displacement (forward-mode Jacobian of trunk + out_nn) == the reference's autograd.grad of phi
    (/root/reference/data_creator_2d.py:98-107), array mode, default initialisation."""
    torch.manual_seed(0)
    m = DMM(s=24, mode="array", branch_layer=7, trunk_layer=[2, 32, 512], out_layer=[1024, 512, 1]).eval()
    B, n = 3, 24 * 24
    u, xi = torch.randn(B, 24, 24), torch.rand(B * n, 2)
    g1, g2 = m.displacement(u, xi)
    x1, x2 = xi[:, 0:1].clone().requires_grad_(True), xi[:, 1:2].clone().requires_grad_(True)
    phi = m(u, torch.cat((x1, x2), -1))
    a1, a2 = torch.autograd.grad(phi, (x1, x2), grad_outputs=torch.ones_like(phi))
    assert torch.allclose(g1, a1, rtol=1e-4, atol=1e-7) and torch.allclose(g2, a2, rtol=1e-4, atol=1e-7)
    # the graph creator uses it (and falls back to autograd for movers without a Jacobian of their own)
    from mmpde_b200.PDEs import burgers
    from mmpde_b200.data_creator_2d import GraphCreator_FS_2D
    from tests.golden.common import SmoothMover
    pde = burgers()
    pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = [31, 24, 24]
    gc = GraphCreator_FS_2D(pde, 35, "knn", 1, 31)
    mx, my = gc.moving_mesh(u, m, 24, 24)
    gx = np.linspace(0, 1, 24)
    grid = torch.tensor(np.array(np.meshgrid(gx, gx)), dtype=torch.float).reshape(2, -1).t().repeat(B, 1)
    d1, d2 = m.displacement(u, grid)
    assert torch.allclose(mx, grid[:, 0:1] + d1, atol=1e-6) and torch.allclose(my, grid[:, 1:2] + d2, atol=1e-6)
    sx, _ = gc.moving_mesh(u, SmoothMover(), 24, 24)
    assert sx.shape == mx.shape and not sx.requires_grad


def test_dataset_files_are_read_like_the_reference(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    os.makedirs("mesh/data")
    rng = np.random.default_rng(0)
    raw = rng.standard_normal((3, 31, 192, 192)).astype(np.float32)
    np.save("mesh/data/burgers_192.npy", raw)
    pde, u = mmpde._load_data(_args(experiment="burgers", synthetic=False, base_resolution=[31, 48, 48]), "cpu")
    assert u.shape == (3, 31, 48, 48) and torch.equal(u, torch.tensor(raw)[:, :, ::4, ::4])          # mmpde.py:170-173
    assert f"{pde}" == "PDE"
    cyl = torch.randn(4, 40, 50, 5)
    torch.save(cyl, "mesh/data/cylinder_rot_tri")
    pde, u = mmpde._load_data(_args(experiment="cy", synthetic=False, base_resolution=[30, 50]), "cpu")
    assert torch.equal(pde.ori_grid, 2 * cyl[0, 0, :, :2]) and torch.equal(u, cyl[:, 10:, :, 2])      # mmpde.py:163-168
    # without the files the seeded synthetic stand-ins are used
    monkeypatch.chdir(tmp_path / "mesh")
    _, u = mmpde._load_data(_args(experiment="burgers", synthetic=False, base_resolution=[31, 16, 16], n_traj=5), "cpu")
    assert u.shape == (5, 31, 16, 16)


def test_dmm_checkpoint_is_loaded_like_the_reference(tmp_path, monkeypatch):
    """burgers_checkpoint / cy_checkpoint: a dict with the pickled trainer args and the DMM state dict
    (/root/reference/mmpde.py:191-200)."""
    from mmpde_b200.PDEs import burgers
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(3)
    src = DMM(s=16, mode="array", branch_layer=7, trunk_layer=[2, 32, 512], out_layer=[1024, 512, 1])
    trainer_args = argparse.Namespace(branch_layers=7, trunk_layers=[32, 512], out_layers=[1024, 512, 1])
    torch.save({"args": trainer_args, "model_state_dict": src.state_dict()}, "burgers_checkpoint")
    pde = burgers()
    pde.movingmesh_grid_size = [31, 16, 16]
    got = mmpde._mesh_mover(_args(experiment="burgers", synthetic=False), pde, "cpu")
    assert isinstance(got, DMM) and not got.training
    for k, v in src.state_dict().items():
        assert torch.equal(got.state_dict()[k], v), k
    os.remove("burgers_checkpoint")
    assert isinstance(mmpde._mesh_mover(_args(experiment="burgers", synthetic_mover="dmm"), pde, "cpu"), DMM)
    assert not isinstance(mmpde._mesh_mover(_args(experiment="burgers"), pde, "cpu"), DMM)


def test_checkpoint_save_and_resume_roundtrip(tmp_path):
    """save_checkpoint writes the reference's keys plus optimizer / scheduler / epoch; load_checkpoint restores a fresh
    set of modules to the same state (and tolerates a reference checkpoint without the extra keys)."""
    from mmpde_b200.PDEs import burgers
    from mmpde_b200.gnn_2d import MP_PDE_Solver_2D
    from mmpde_b200.interpolate import ItpNet
    pde = burgers()

    def build(seed):
        torch.manual_seed(seed)
        m, mb = MP_PDE_Solver_2D(pde, hidden_layer=1), MP_PDE_Solver_2D(pde, hidden_layer=1)
        it = ItpNet(8, 8, [128, 64], [128, 64], [1, 4, 16, 4, 1])
        opt = torch.optim.AdamW([{"params": m.parameters()}, {"params": mb.parameters()}, {"params": it.parameters()}], lr=2e-3)
        sch = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[0, 30, 50, 70], gamma=0.4)
        return m, mb, it, opt, sch

    m, mb, it, opt, sch = build(1)
    for p in list(m.parameters()) + list(mb.parameters()) + list(it.parameters()):
        p.grad = torch.randn_like(p)
    opt.step(); sch.step()
    path = str(tmp_path / "ck.pt")
    state = mmpde.save_checkpoint(path, _args(), 4, m, mb, it, None, opt, sch, [[1.0]], [[2.0]], [3.0])
    assert {"model_state_dict", "model_b_state_dict", "itp_model_state_dict", "args", "train_losses", "itp_losses",
            "test_timestep_losses"} <= set(state)                                           # mmpde.py:292-310
    m2, mb2, it2, opt2, sch2 = build(2)
    nxt, tl, il, te = mmpde.load_checkpoint(path, m2, mb2, it2, opt2, sch2)
    assert nxt == 5 and tl == [[1.0]] and il == [[2.0]] and te == [3.0]
    for a, b in ((m, m2), (mb, mb2), (it, it2)):
        for k, v in a.state_dict().items():
            assert torch.equal(v, b.state_dict()[k]), k
    assert opt2.param_groups[0]["lr"] == opt.param_groups[0]["lr"] and sch2.last_epoch == sch.last_epoch
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys() and all(torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"]) for k in s1)
    # a reference checkpoint has no optimizer / scheduler / epoch entries
    ref_like = {k: v for k, v in state.items() if k not in ("optimizer_state_dict", "scheduler_state_dict", "epoch")}
    torch.save(ref_like, path)
    m3, mb3, it3, opt3, sch3 = build(3)
    assert mmpde.load_checkpoint(path, m3, mb3, it3, opt3, sch3)[0] == 0
    assert all(torch.equal(v, m3.state_dict()[k]) for k, v in m.state_dict().items())
