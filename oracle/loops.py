"""ORACLE (test infrastructure).  Step loops, restating /root/reference/train_helper_2d.py:9-200
and the MSE criterion of /root/reference/mmpde.py:33-36."""
import random

import torch


def criterion(x, y):
    return torch.nn.functional.mse_loss(x, y)


def _start_steps(gc, unrolling, batch_size):
    extra = random.choice(unrolling)
    lo, hi = gc.tw, gc.t_res - gc.tw - gc.tw * extra + 1
    return random.choices(list(range(lo, hi)), k=batch_size)


def _predict(model, model_b, itp_model, mesh_model, gc, data, labels, steps, device):
    """train_helper_2d.py:107-118 / :174-183 (GNN branch)."""
    if mesh_model is not None:
        graph = gc.create_graph(itp_model, data, labels, steps, device, mesh_model)
    graph_uni = gc.create_graph(itp_model, data, labels, steps, device, None)
    if mesh_model is not None:
        return gc.interpolate_pred(itp_model, model_b(graph), graph, data, device) + model(graph_uni)
    return model(graph_uni)


def training_itp(itp_model, mesh_model, unrolling, batch_size, optimizer, optimizer2, loader,
                 graph_creator, criterion, device="cpu"):
    losses = []
    for (_, u_super) in loader:
        optimizer.zero_grad()
        if optimizer2 is not None:
            optimizer2.zero_grad()
        steps = _start_steps(graph_creator, unrolling, batch_size)
        data, labels = graph_creator.create_data(u_super, steps)
        graph = graph_creator.create_graph(itp_model, data, labels, steps, device, mesh_model)
        back = graph_creator.interpolate_pred(itp_model, graph.x, graph, data, device)
        loss = criterion(back, data.to(device).reshape(-1, 1))
        loss.backward()
        losses.append(loss.detach() / 2)                    # train_helper_2d.py:56
        optimizer.step()
        if optimizer2 is not None:
            optimizer2.step()
    return torch.stack(losses)


def training_loop_branch(model, model_b, itp_model, mesh_model, unrolling, batch_size, optimizer,
                         optimizer2, loader, graph_creator, criterion, device="cpu"):
    assert f"{model}" == "GNN"
    losses = []
    for (_, u_super) in loader:
        optimizer.zero_grad()
        if optimizer2 is not None:
            optimizer2.zero_grad()
        steps = _start_steps(graph_creator, unrolling, batch_size)
        data, labels = graph_creator.create_data(u_super, steps)
        pred = _predict(model, model_b, itp_model, mesh_model, graph_creator, data, labels, steps, device)
        loss = criterion(pred, labels.to(device).reshape(-1, 1))
        loss.backward()
        losses.append(loss.detach())
        optimizer.step()
        if optimizer2 is not None:
            optimizer2.step()
    return torch.stack(losses)


def test_timestep_losses(model, model_b, itp_model, mesh_model, steps, batch_size, loader,
                         graph_creator, criterion, device="cpu", return_curve=False, verbose=False):
    assert f"{model}" == "GNN"
    curve = []
    for step in steps:
        if step != graph_creator.tw and step % graph_creator.tw != 0:
            continue
        per_batch = []
        for (_, u_super) in loader:
            same = [step] * batch_size
            data, labels = graph_creator.create_data(u_super, same)
            if mesh_model is not None:          # the moved graph is built OUTSIDE no_grad (:175-177)
                graph = graph_creator.create_graph(itp_model, data, labels, same, device, mesh_model)
            graph_uni = graph_creator.create_graph(itp_model, data, labels, same, device, None)
            with torch.no_grad():
                if mesh_model is not None:
                    pred = graph_creator.interpolate_pred(itp_model, model_b(graph), graph, data, device) \
                        + model(graph_uni)
                else:
                    pred = model(graph_uni)
                per_batch.append(criterion(pred, labels.to(device).reshape(-1, 1)))
        curve.append(torch.mean(torch.stack(per_batch)))
        if verbose and step % 2 == 1:
            print(f"Step {step}, time step loss {curve[-1]}")
    curve = torch.stack(curve)
    return (torch.mean(curve), curve) if return_curve else torch.mean(curve)
