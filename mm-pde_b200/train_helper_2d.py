"""Step loops -- drop-in for /root/reference/train_helper_2d.py (same function names and signatures:
``training_itp`` :9-62, ``training_loop_branch`` :65-134, ``test_timestep_losses`` :137-200).

The loops are host-side control flow only; everything they call (graph creation, both solvers, the
interpolation) runs on the sm_100a kernels.  Two additions, both optional keyword arguments:
``after_backward`` -- the multi-GPU driver passes the flat-bucket gradient all-reduce there
(mmpde_b200.dist.GradBucket.allreduce); ``step_graph`` -- a StepGraph that records the device work of one
step (~800 launches: both solvers, k-NN, interpolation, backward, optimizer) into a CUDA graph and replays
it, because queueing those launches from Python costs about as long as the GPU needs to run them.
"""
import random

import torch

from . import _cabi, ops
from ._h2d import to_device


class StepGraph:
    """CUDA-graph record/replay of the per-batch device work of the step loops.

    The first ``eager_steps`` calls with a given signature (input shapes, train/eval mode, optimizer
    hyper-parameters) run eagerly -- they fill the static caches (uniform-grid edges, optimizer state, cuDNN
    plans) -- the next one is captured on static input buffers, later ones copy the step's inputs into those
    buffers and replay.  Shapes are static by construction (fixed k, fixed batch); a new shape or a changed
    learning rate simply records a new graph.  Training needs optimizers built with ``capturable=True`` (their
    step counter must live on the device)."""

    def __init__(self, eager_steps=2, max_graphs=3):
        self.eager_steps = eager_steps
        self.max_graphs = max_graphs  # recordings kept per loop (e.g. full batch + the loader's last short batch)
        self._seen = {}
        self._graphs = {}            # signature -> (graph, static inputs, static output, launches per replay)
        self.replays = 0

    @staticmethod
    def hyper(*optimizers):
        sig = []
        for opt in optimizers:
            if opt is None:
                continue
            for g in opt.param_groups:
                if not g.get("capturable", False):
                    raise ValueError("StepGraph needs optimizers built with capturable=True")
                sig.append(tuple((k, v) for k, v in sorted(g.items())
                                 if isinstance(v, (int, float, bool, str, tuple)) and k != "params"))
        return tuple(sig)

    def release(self):
        """Drop every recording (their memory pools and, with several ranks, the NCCL work captured in them).  Call
        it before the process group is destroyed: a graph that still holds collectives of a dead communicator can
        block the teardown."""
        import gc
        self._graphs.clear()
        self._seen.clear()
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def run(self, signature, body, inputs, prepare=None):
        """body(*device tensors) -> device tensor.  ``inputs``: tensors already on the device.  ``prepare`` runs
        before every eager call and once before the recording (zero_grad: a replay rewrites the gradients the
        recording allocated, it never accumulates)."""
        n = self._seen.get(signature, 0)
        self._seen[signature] = n + 1
        if n < self.eager_steps or not inputs[0].is_cuda:
            if prepare is not None:
                prepare()
            return body(*inputs)
        hit = self._graphs.get(signature)
        if hit is None:
            same_loop = [k for k in self._graphs if k[0] == signature[0]]
            for k in same_loop[:max(0, len(same_loop) - self.max_graphs + 1)]:
                del self._graphs[k]          # oldest recordings of this loop: free their memory pools
            static = [torch.empty_like(t) for t in inputs]
            for s_, t in zip(static, inputs):
                s_.copy_(t)
            if prepare is not None:
                prepare()
            graph = torch.cuda.CUDAGraph()
            l0 = _cabi.launches
            with torch.cuda.graph(graph):
                out = body(*static)
            hit = self._graphs[signature] = (graph, static, out, _cabi.launches - l0)
            _cabi.launches = l0
        graph, static, out, launches = hit
        for s_, t in zip(static, inputs):
            s_.copy_(t, non_blocking=True)
        graph.replay()
        _cabi.launches += launches
        self.replays += 1
        return out.clone()


def _sample_steps(graph_creator, unrolling, batch_size):
    # random.choice / random.choices in this order, like the reference, so seeded runs pick the same steps
    unrolled = random.choice(unrolling)
    first = graph_creator.tw
    last = graph_creator.t_res - graph_creator.tw - graph_creator.tw * unrolled
    return random.choices(range(first, last + 1), k=batch_size)


def _is_gnn(model):
    return f"{model}" == "GNN"


_side_streams = {}


def _side_stream(device):
    key = str(device)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def _overlap_solvers(device):
    """The two solvers of a step (uniform grid, moved mesh) are independent until their outputs are added.  Their big
    kernels are persistent one-CTA-per-SM kernels, so they cannot share an SM -- but issued on two streams (two parallel
    branches of the step's CUDA graph) the launch ramp and the tail of every kernel of one branch are filled by the
    other branch.  With several ranks each branch needs its own cross-GPU exchange sequence (dist.DistComm gives every
    branch its own peer-memory buffer); a COMM that cannot (NCCL fallback: collectives need ONE global order) keeps the
    solvers on one stream."""
    import os
    from . import ops
    if os.environ.get("MMPDE_OVERLAP_SOLVERS", "1") == "0" or not torch.cuda.is_available():
        return False
    return torch.device(device).type == "cuda" and (type(ops.COMM) is ops._Comm or ops.COMM.n_branches >= 2)


def _branch_ctas(device, mesh_model):
    """(uniform, moved): CTAs per persistent kernel of the two solvers of an overlapped TRAINING step, or (0, 0) = one per
    SM.  At full width the persistent kernels of the two branches take turns on the whole chip at kernel granularity; with
    the SMs divided between them the branches advance side by side.  One GPU: the same step time at 74 / 74 (10.02 vs
    10.01 ms).  Several GPUs: every rank then interleaves its branches the same way and the ranks stop drifting apart
    between the cross-GPU BatchNorm exchanges (exchange-carrying kernels 83 -> 31 us at two ranks, e2e 10.64 -> 9.99 ms).
    The moved-mesh branch also runs the mesh mover, the neighbour searches and the interpolation (about 1 ms with no
    counterpart on the uniform branch), so it gets the larger share and the two finish together.
    MMPDE_BRANCH_SMS = "u,m" or one number overrides (0 = full width)."""
    import os
    if mesh_model is None or not _overlap_solvers(device):
        return 0, 0
    env = os.environ.get("MMPDE_BRANCH_SMS")
    if env is not None:
        parts = [max(int(v), 0) for v in env.split(",")]
        return (parts[0], parts[-1])
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    return sms // 2, sms - sms // 2


def _forward_gnn(model, model_b, itp_model, mesh_model, graph_creator, data, labels, steps, device):
    """Both branches of the MM-PDE prediction (train_helper_2d.py:107-118): the branch solver on the moved
    mesh, interpolated back to the grid (+ residual net), plus the solver on the uniform grid."""
    uniform = graph_creator.create_graph(itp_model, data, labels, steps, device, None)
    if mesh_model is None:
        return model(uniform)
    if _overlap_solvers(device):
        cur, side = torch.cuda.current_stream(), _side_stream(device)
        side.wait_stream(cur)
        w_uniform, w_moved = _branch_ctas(device, mesh_model)
        with torch.cuda.stream(side):
            ops.COMM.branch = 1              # this solver's BatchNorm exchanges (forward AND backward) use sequence 1
            ops.SOLVER_WIDTH = w_uniform
            try:
                on_uniform = model(uniform)
            finally:
                ops.COMM.branch = 0
                ops.SOLVER_WIDTH = 0
        # (the longer moved-mesh branch on a high-priority stream was measured: 10.20 vs 10.13 ms per step -- not kept)
        moved = graph_creator.create_graph(itp_model, data, labels, steps, device, mesh_model)
        ops.SOLVER_WIDTH = w_moved
        try:
            on_moved = graph_creator.interpolate_pred(itp_model, model_b(moved), moved, data, device)
        finally:
            ops.SOLVER_WIDTH = 0
        cur.wait_stream(side)
        on_uniform.record_stream(cur)
        return on_moved + on_uniform
    moved = graph_creator.create_graph(itp_model, data, labels, steps, device, mesh_model)
    return graph_creator.interpolate_pred(itp_model, model_b(moved), moved, data, device) + model(uniform)


def _steps_tensor(steps, device):
    return to_device(torch.as_tensor(list(steps), dtype=torch.int64), device)


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
                         loader, graph_creator, criterion, device="cpu", after_backward=None, step_graph=None):
    """One pass over the loader with a random start step per trajectory."""
    history = []

    def gnn_step(data, labels, steps):
        pred = _forward_gnn(model, model_b, itp_model, mesh_model, graph_creator, data, labels, steps, device)
        loss = criterion(pred, to_device(labels, device).reshape(-1, 1))
        loss.backward()
        if after_backward is not None:
            after_backward()
        _step_optimizers(optimizer, optimizer2)
        return loss.detach()

    for (_, u_super) in loader:
        steps = _sample_steps(graph_creator, unrolling, batch_size)
        data, labels = graph_creator.create_data(u_super, steps)
        if _is_gnn(model) and step_graph is not None:
            data, labels = to_device(data, device), to_device(labels, device)
            sig = ("train", tuple(data.shape), model.training, model_b.training if model_b is not None else None,
                   mesh_model is None, StepGraph.hyper(optimizer, optimizer2))
            history.append(step_graph.run(sig, gnn_step, (data, labels, _steps_tensor(steps, device)),
                                          prepare=lambda: _zero(optimizer, optimizer2)))
            continue
        _zero(optimizer, optimizer2)
        if _is_gnn(model):
            history.append(gnn_step(data, labels, steps))
            continue
        data, labels = to_device(data, device), to_device(labels, device)
        loss = criterion(model(data), labels.squeeze())
        loss.backward()
        if after_backward is not None:
            after_backward()
        history.append(loss.detach())
        _step_optimizers(optimizer, optimizer2)
    return torch.stack(history)


def test_timestep_losses(model, model_b, itp_model, mesh_model, steps, batch_size, loader, graph_creator,
                         criterion, device="cpu", return_curve=False, step_graph=None):
    """Teacher-forced one-step error for every start step (the reference's "rollout" curve)."""
    curve = []

    def gnn_eval(data, labels, step_idx):
        # full width: eval-mode BatchNorm has no cross-GPU exchange to keep in step, and the forward-only pass spends a
        # larger share of its time where only one branch has work (3.73 ms per batch at full width, 3.90 at half)
        pred = _forward_gnn(model, model_b, itp_model, mesh_model, graph_creator, data, labels, step_idx, device)
        return criterion(pred, to_device(labels, device).reshape(-1, 1))

    for step in steps:
        if step != graph_creator.tw and step % graph_creator.tw != 0:
            continue
        per_batch = []
        for (_, u_super) in loader:
            data, labels = graph_creator.create_data(u_super, [step] * batch_size)
            with torch.no_grad():
                if _is_gnn(model) and step_graph is not None:
                    data, labels = to_device(data, device), to_device(labels, device)
                    sig = ("eval", tuple(data.shape), model.training, model_b.training if model_b is not None else None,
                           mesh_model is None)
                    per_batch.append(step_graph.run(sig, gnn_eval, (data, labels, _steps_tensor([step] * batch_size, device))))
                elif _is_gnn(model):
                    per_batch.append(gnn_eval(data, labels, [step] * batch_size))
                else:
                    data, labels = to_device(data, device), to_device(labels, device)
                    per_batch.append(criterion(model(data), labels.squeeze()))
        curve.append(torch.stack(per_batch).mean())
        if step % 2 == 1:
            print(f"Step {step}, time step loss {curve[-1]}")
    curve = torch.stack(curve)
    print(f"Mean Timestep Test Error: {curve.mean()}")
    return (curve.mean(), curve) if return_curve else curve.mean()
