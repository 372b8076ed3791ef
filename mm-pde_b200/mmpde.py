"""Experiment driver -- drop-in for /root/reference/mmpde.py (``criterion`` :33-36, ``train`` :38-100,
``test`` :102-151, ``main`` :154-319, CLI flags :322-374) running the hot path on sm_100a kernels.

Additions: ``--synthetic`` (seeded data / analytic or default-initialised mesh mover when the Google-Drive
datasets and DMM checkpoints are absent), multi-GPU batch sharding when launched under torchrun, and
``--max_steps`` to bound a run.  Everything else (flags, defaults, epoch structure, checkpoint keys,
MultiStepLR with milestone ``unrolling`` = 0) follows the reference.
"""
import argparse
import os
import random
from datetime import datetime

import numpy as np
import torch
from torch import optim
from torch.utils.data import DataLoader, TensorDataset

from . import dist as mdist
from . import synthetic
from .PDEs import PDE, burgers, cy
from .data_creator_2d import GraphCreator_FS_2D
from .gnn_2d import MP_PDE_Solver_2D
from .interpolate import ItpNet
from .mesh.dmm_model import DMM
from .train_helper_2d import StepGraph, test_timestep_losses, training_itp, training_loop_branch


def check_directory():
    for d in ("logs", "models"):
        os.makedirs(d, exist_ok=True)


def criterion(x, y):
    return torch.nn.functional.mse_loss(x, y)


def train(args, pde, epoch, model, model_b, itp_model, mesh_model, optimizer, optimizer2, loader, graph_creator,
          criterion, device="cpu", after_backward=None, step_graph=None):
    print(f"Starting epoch {epoch}...")
    model.train()
    if model_b is not None:
        model_b.train()
    max_unrolling = epoch if epoch <= args.unrolling else args.unrolling
    unrolling = list(range(max_unrolling + 1))
    passes = graph_creator.t_res if getattr(args, "max_passes", None) is None else min(args.max_passes, graph_creator.t_res)
    itp_losses = []
    if mesh_model is not None:
        itp_model.train()
        if epoch == 0:
            for i in range(passes):
                losses = training_itp(itp_model, mesh_model, unrolling, 128 * args.batch_size, optimizer, optimizer2,
                                      loader, graph_creator, criterion, device, after_backward)
                if i % args.print_interval == 0:
                    print(f"Training ItpNet Loss (progress: {i / graph_creator.t_res:.2f}): {torch.mean(losses)}")
            itp_losses.append(torch.mean(losses))
    train_losses = []
    for i in range(passes):
        losses = training_loop_branch(model, model_b, itp_model, mesh_model, unrolling, args.batch_size, optimizer,
                                      optimizer2, loader, graph_creator, criterion, device, after_backward,
                                      step_graph=step_graph)
        if i % args.print_interval == 0:
            print(f"Training Loss (progress: {i / graph_creator.t_res:.2f}): {torch.mean(losses)}")
        train_losses.append(torch.mean(losses))
    return train_losses, itp_losses


def test(args, pde, model, model_b, itp_model, mesh_model, loader, graph_creator, criterion, device="cpu",
         step_graph=None):
    model.eval()
    if model_b is not None:
        model_b.eval()
    if itp_model is not None:
        itp_model.eval()
    steps = list(range(graph_creator.tw, graph_creator.t_res - graph_creator.tw + 1))
    return test_timestep_losses(model=model, model_b=model_b, itp_model=itp_model, mesh_model=mesh_model, steps=steps,
                                batch_size=args.batch_size, loader=loader, graph_creator=graph_creator,
                                criterion=criterion, device=device, step_graph=step_graph)


def _load_data(args, device):
    res = args.base_resolution
    if args.experiment == "cy":
        path = "mesh/data/cylinder_rot_tri"
        if not args.synthetic and os.path.exists(path):
            data = torch.load(path)
            data[:, :, :, :2] *= 2
            grid, u = data[0, 0, :, :2], data[:, 10:, :, 2]
        else:
            grid = synthetic.cylinder_cloud(res[1], seed=args.seed)
            u = synthetic.cylinder_fields(args.n_traj, grid, res[0], seed=args.seed)
        return cy(ori_grid=grid, device=device), u
    if args.experiment == "burgers":
        path = "mesh/data/burgers_192.npy"
        if not args.synthetic and os.path.exists(path):
            u = torch.tensor(np.load(path), dtype=torch.float)[:, :, ::int(192 / res[1]), ::int(192 / res[2])]
        else:
            u = synthetic.burgers_fields(args.n_traj, res[0], res[1], res[2], seed=args.seed)
        return burgers(device=device), u
    raise Exception("Wrong experiment")


def _mesh_mover(args, pde, device):
    ckpt_path = "cy_checkpoint" if args.experiment == "cy" else "burgers_checkpoint"
    if not args.synthetic and os.path.exists(ckpt_path):
        ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
        a = ckpt["args"]
        if args.experiment == "cy":
            m = DMM(mode="graph", grid=pde.ori_grid.to(device), branch_layer=a.branch_layers,
                    trunk_layer=[2] + a.trunk_layers, out_layer=a.out_layers)
        else:
            m = DMM(s=pde.movingmesh_grid_size[-1], mode="array", branch_layer=a.branch_layers,
                    trunk_layer=[2] + a.trunk_layers, out_layer=a.out_layers)
        m.load_state_dict(ckpt["model_state_dict"])
        return m.to(device).eval()
    if getattr(args, "synthetic_mover", "analytic") == "dmm":
        # default-initialised DMM with the reference's constructor arguments (mmpde.py:199, README.md:31)
        torch.manual_seed(args.seed + 1)
        if args.experiment == "cy":
            m = DMM(mode="graph", grid=pde.ori_grid.to(device), branch_layer=[4, 3], trunk_layer=[2, 16, 512],
                    out_layer=[1024, 512, 1])
        else:
            m = DMM(s=pde.movingmesh_grid_size[-1], mode="array", branch_layer=7, trunk_layer=[2, 32, 512],
                    out_layer=[1024, 512, 1])
        return m.to(device).eval()
    return synthetic.AnalyticMover().to(device).eval()


def save_checkpoint(path, args, epoch, model, model_b, itp_model, mesh_model, optimizer, scheduler, train_losses, itp_losses,
                    test_losses):
    """The reference's checkpoint dictionary (mmpde.py:292-310: model / model_b / mesh_model / itp_model state dicts, args,
    loss histories) plus what a resumed run needs and the reference does not store: optimizer, scheduler, epoch."""
    state = {"model_state_dict": model.state_dict(), "args": args, "train_losses": train_losses, "itp_losses": itp_losses,
             "test_timestep_losses": test_losses, "epoch": epoch, "optimizer_state_dict": optimizer.state_dict(),
             "scheduler_state_dict": scheduler.state_dict() if scheduler is not None else None}
    if model_b is not None:
        state.update(model_b_state_dict=model_b.state_dict(), itp_model_state_dict=itp_model.state_dict())
        if mesh_model is not None:
            state["mesh_model_state_dict"] = mesh_model.state_dict()
    torch.save(state, path)
    return state


def load_checkpoint(path, model, model_b, itp_model, optimizer=None, scheduler=None, map_location="cpu"):
    """Restores what save_checkpoint wrote (also accepts the reference's own checkpoints, which lack the optimizer /
    scheduler / epoch entries).  Returns (next epoch, train_losses, itp_losses, test_losses)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(ckpt["model_state_dict"])
    if model_b is not None and "model_b_state_dict" in ckpt:
        model_b.load_state_dict(ckpt["model_b_state_dict"])
    if itp_model is not None and "itp_model_state_dict" in ckpt:
        itp_model.load_state_dict(ckpt["itp_model_state_dict"])
    if optimizer is not None and ckpt.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    if scheduler is not None and ckpt.get("scheduler_state_dict") is not None:
        scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    return (int(ckpt.get("epoch", -1)) + 1, ckpt.get("train_losses", []), ckpt.get("itp_losses", []),
            ckpt.get("test_timestep_losses", []))


def main(args):
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    random.seed(args.seed)
    rank, world, device = mdist.init_from_env()
    if args.device != "auto":
        device = torch.device(args.device)
        if device.type == "cuda":
            torch.cuda.set_device(device)      # the C ABI launches on the current device's current stream
    check_directory()
    pde, u = _load_data(args, device)
    split = int(0.8 * u.shape[0]) if u.shape[0] < 100 else 80
    u_train, u_test = u[:split], u[split:]
    pde.grid_size = pde.movingmesh_grid_size = pde.ori_grid_size = args.base_resolution

    if not args.moving_mesh:
        itp_model = mesh_model = None
    else:
        if args.experiment == "cy":
            itp_model = ItpNet(pde.ori_grid_size[1], None, args.itpnet_node1, args.itpnet_node2, args.res_cut_node).to(device)
        else:
            itp_model = ItpNet(pde.ori_grid_size[-2], pde.ori_grid_size[-1], args.itpnet_node1, args.itpnet_node2,
                               args.res_cut_node).to(device)
        mesh_model = _mesh_mover(args, pde, device)

    per_rank = args.batch_size
    if world > 1:
        # Batch sharding: every rank loads its own slice of each global batch.  All ranks must run the SAME number of
        # steps with the SAME per-rank batch sizes (sync-BatchNorm counts rows as n * world and every step is a sequence
        # of collectives), so the trajectories that do not divide evenly are dropped.
        u_train = mdist.shard_batch(u_train[:u_train.shape[0] // world * world], rank, world)
        u_test = mdist.shard_batch(u_test[:u_test.shape[0] // world * world], rank, world)
        if u_train.shape[0] == 0 or u_test.shape[0] == 0:
            raise ValueError(f"fewer trajectories than ranks ({world}): nothing to shard")
    train_loader = DataLoader(TensorDataset(u_train, u_train), batch_size=per_rank, shuffle=True, num_workers=0)
    test_loader = DataLoader(TensorDataset(u_test, u_test), batch_size=per_rank, shuffle=False, num_workers=0)

    stamp = datetime.now()
    save_path = (f"models/{args.model}_{pde}_{args.experiment}_mesh{args.moving_mesh}_xresolution"
                 f"{args.base_resolution[0]}-{args.base_resolution[1]}_n{args.neighbors}_{args.connect_edge}_tw"
                 f"{args.time_window}_unrolling{args.unrolling}_time{stamp.month}-{stamp.day}-{stamp.hour}-"
                 f"{stamp.minute}-{stamp.second}.pt")
    print(f"Training on dataset of {args.experiment}")
    print(device)

    graph_creator = GraphCreator_FS_2D(pde=pde, neighbors=args.neighbors, connect_edge=args.connect_edge,
                                       time_window=args.time_window, t_resolution=args.base_resolution[0]).to(device)
    if args.model != "GNN":
        raise Exception("Wrong model specified (BaseCNN is outside the B200 hot path)")
    model = MP_PDE_Solver_2D(pde=pde, time_window=graph_creator.tw, eq_variables={}).to(device)
    model_b = MP_PDE_Solver_2D(pde=pde, time_window=graph_creator.tw, eq_variables={}).to(device) if args.moving_mesh else None

    groups = [{"params": model.parameters()}]
    if mesh_model is not None:
        groups += [{"params": model_b.parameters()}, {"params": itp_model.parameters()}]
    n_params = sum(p.numel() for g in groups for p in g["params"] if p.requires_grad)
    print(f"Number of parameters: {n_params}")
    groups = [{"params": model.parameters()}]
    if mesh_model is not None:
        groups += [{"params": model_b.parameters()}, {"params": itp_model.parameters()}]
    # step_graph: record the step once, replay it afterwards (train_helper_2d.StepGraph).  The optimizer then keeps
    # its step counters on the device (capturable) and updates all tensors in one fused multi-tensor kernel instead of
    # ~800 single-tensor launches per step.
    use_graph = bool(getattr(args, "step_graph", True)) and str(device).startswith("cuda")
    step_graph = StepGraph() if use_graph else None
    optimizer = optim.AdamW(groups, lr=args.lr, capturable=use_graph, fused=use_graph or None)
    scheduler = optim.lr_scheduler.MultiStepLR(optimizer, milestones=[args.unrolling, 30, 50, 70], gamma=args.lr_decay)
    after_backward = None
    if world > 1:
        bucket = mdist.GradBucket([p for g in optimizer.param_groups for p in g["params"]])
        after_backward = bucket.allreduce

    train_losses, itp_losses, test_losses = [], [], []
    first_epoch = 0
    if getattr(args, "resume", None):
        first_epoch, train_losses, itp_losses, test_losses = load_checkpoint(args.resume, model, model_b, itp_model, optimizer,
                                                                             scheduler, map_location=device)
        print(f"Resumed from {args.resume} at epoch {first_epoch}")
    for epoch in range(first_epoch, args.num_epochs):
        print(f"Epoch {epoch}")
        tl, il = train(args, pde, epoch, model, model_b, itp_model, mesh_model, optimizer, None, train_loader,
                       graph_creator, criterion, device=device, after_backward=after_backward, step_graph=step_graph)
        train_losses.append(tl)
        itp_losses.append(il)
        print("Testing:")
        test_losses.append(test(args, pde, model, model_b, itp_model, mesh_model, test_loader, graph_creator,
                                criterion, device=device, step_graph=step_graph))
        scheduler.step()                       # before the save: a resumed run continues with the next epoch's rate
        if rank == 0:
            save_checkpoint(save_path, args, epoch, model, model_b, itp_model, mesh_model, optimizer, scheduler, train_losses,
                            itp_losses, test_losses)
            print(f"Saved model at {save_path}\n")
        if world > 1:
            torch.distributed.barrier()        # rank-0-only work above: keep the ranks' exchange sequences aligned
    if step_graph is not None:
        step_graph.release()
    return train_losses, test_losses


def _int_list(s):
    return [int(item) for item in s.split(",")]


def build_parser():
    p = argparse.ArgumentParser(description="Train a PDE solver")
    p.add_argument("--seed", default=1, type=int, help="random seed")
    p.add_argument("--device", type=str, default="auto", help="Used device")
    p.add_argument("--experiment", type=str, default="burgers", help="[burgers, cy]")
    p.add_argument("--model", type=str, default="GNN", help="[GNN]")
    p.add_argument("--moving_mesh", type=eval, default=True, help="Use moving mesh method")
    p.add_argument("--itpnet_node1", type=_int_list, default=[128, 64], help="nodes of ItpNet1")
    p.add_argument("--itpnet_node2", type=_int_list, default=[128, 64], help="nodes of ItpNet2")
    p.add_argument("--res_cut_node", type=_int_list, default=[1, 4, 16, 4, 1], help="nodes of residual cut network")
    p.add_argument("--hidden_channels", type=int, default=40)
    p.add_argument("--batch_size", type=int, default=6)
    p.add_argument("--num_epochs", type=int, default=80)
    p.add_argument("--lr", type=float, default=2e-3)
    p.add_argument("--lr_decay", type=float, default=0.4)
    p.add_argument("--base_resolution", type=_int_list, default=[31, 48, 48])
    p.add_argument("--neighbors", type=int, default=35)
    p.add_argument("--connect_edge", type=str, default="knn", help="[knn, radius]")
    p.add_argument("--time_window", type=int, default=1)
    p.add_argument("--unrolling", type=int, default=0)
    p.add_argument("--print_interval", type=int, default=2)
    p.add_argument("--log", type=eval, default=True)
    # additions
    p.add_argument("--synthetic", type=eval, default=True, help="seeded synthetic data / analytic mesh mover")
    p.add_argument("--n_traj", type=int, default=20, help="synthetic trajectories")
    p.add_argument("--step_graph", type=eval, default=True, help="replay the recorded step as a CUDA graph")
    p.add_argument("--max_passes", type=int, default=None, help="bound the passes per epoch (default t_res)")
    p.add_argument("--synthetic_mover", type=str, default="analytic", help="[analytic, dmm] mover when no DMM checkpoint exists")
    p.add_argument("--resume", type=str, default=None, help="checkpoint written by this script to continue from")
    return p


if __name__ == "__main__":
    a = build_parser().parse_args()
    print(a)
    main(a)
