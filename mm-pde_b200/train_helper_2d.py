"""Step loops -- drop-in for /root/reference/train_helper_2d.py (same function names and signatures:
``training_itp`` :9-62, ``training_loop_branch`` :65-134, ``test_timestep_losses`` :137-200).

The loops are host-side control flow only; everything they call (graph creation, both solvers, the
interpolation) runs on the sm_100a kernels.  ``after_backward`` is the one addition: the multi-GPU
driver passes the flat-bucket gradient all-reduce there (mmpde_b200.dist.allreduce_gradients).
"""
import random

import torch

from ._h2d import to_device


def _sample_steps(graph_creator, unrolling, batch_size):
    # random.choice / random.choices in this order, like the reference, so seeded runs pick the same steps
    unrolled = random.choice(unrolling)
    first = graph_creator.tw
    last = graph_creator.t_res - graph_creator.tw - graph_creator.tw * unrolled
    return random.choices(range(first, last + 1), k=batch_size)


def _is_gnn(model):
    return f"{model}" == "GNN"


def _forward_gnn(model, model_b, itp_model, mesh_model, graph_creator, data, labels, steps, device):
    """Both branches of the MM-PDE prediction (train_helper_2d.py:107-118): the branch solver on the moved
    mesh, interpolated back to the grid (+ residual net), plus the solver on the uniform grid."""
    uniform = graph_creator.create_graph(itp_model, data, labels, steps, device, None)
    if mesh_model is None:
        return model(uniform)
    moved = graph_creator.create_graph(itp_model, data, labels, steps, device, mesh_model)
    return graph_creator.interpolate_pred(itp_model, model_b(moved), moved, data, device) + model(uniform)


def _step_optimizers(optimizer, optimizer2):
    optimizer.step()
    if optimizer2 is not None:
        optimizer2.step()


def _zero(optimizer, optimizer2):
    optimizer.zero_grad()
    if optimizer2 is not None:
        optimizer2.zero_grad()


def training_itp(itp_model, mesh_model, unrolling, batch_size, optimizer, optimizer2, loader, graph_creator,
                 criterion, device="cpu", after_backward=None):
    """Interpolation round trip grid -> moved mesh -> grid, trained to reproduce its input."""
    history = []
    for (_, u_super) in loader:
        _zero(optimizer, optimizer2)
        steps = _sample_steps(graph_creator, unrolling, batch_size)
        data, labels = graph_creator.create_data(u_super, steps)
        moved = graph_creator.create_graph(itp_model, data, labels, steps, device, mesh_model)
        round_trip = graph_creator.interpolate_pred(itp_model, moved.x, moved, data, device)
        loss = criterion(round_trip, to_device(data, device).reshape(-1, 1))
        loss.backward()
        if after_backward is not None:
            after_backward()
        history.append(loss.detach() / 2)
        _step_optimizers(optimizer, optimizer2)
    return torch.stack(history)


def training_loop_branch(model, model_b, itp_model, mesh_model, unrolling, batch_size, optimizer, optimizer2,
                         loader, graph_creator, criterion, device="cpu", after_backward=None):
    """One pass over the loader with a random start step per trajectory."""
    history = []
    for (_, u_super) in loader:
        _zero(optimizer, optimizer2)
        steps = _sample_steps(graph_creator, unrolling, batch_size)
        data, labels = graph_creator.create_data(u_super, steps)
        if _is_gnn(model):
            pred = _forward_gnn(model, model_b, itp_model, mesh_model, graph_creator, data, labels, steps, device)
            loss = criterion(pred, to_device(labels, device).reshape(-1, 1))
        else:
            data, labels = to_device(data, device), to_device(labels, device)
            loss = criterion(model(data), labels.squeeze())
        loss.backward()
        if after_backward is not None:
            after_backward()
        history.append(loss.detach())
        _step_optimizers(optimizer, optimizer2)
    return torch.stack(history)


def test_timestep_losses(model, model_b, itp_model, mesh_model, steps, batch_size, loader, graph_creator,
                         criterion, device="cpu", return_curve=False):
    """Teacher-forced one-step error for every start step (the reference's "rollout" curve)."""
    curve = []
    for step in steps:
        if step != graph_creator.tw and step % graph_creator.tw != 0:
            continue
        per_batch = []
        for (_, u_super) in loader:
            data, labels = graph_creator.create_data(u_super, [step] * batch_size)
            with torch.no_grad():
                if _is_gnn(model):
                    pred = _forward_gnn(model, model_b, itp_model, mesh_model, graph_creator, data, labels,
                                        [step] * batch_size, device)
                    per_batch.append(criterion(pred, to_device(labels, device).reshape(-1, 1)))
                else:
                    data, labels = to_device(data, device), to_device(labels, device)
                    per_batch.append(criterion(model(data), labels.squeeze()))
        curve.append(torch.stack(per_batch).mean())
        if step % 2 == 1:
            print(f"Step {step}, time step loss {curve[-1]}")
    curve = torch.stack(curve)
    print(f"Mean Timestep Test Error: {curve.mean()}")
    return (curve.mean(), curve) if return_curve else curve.mean()
