"""MP-PDE processor on sm_100a kernels -- drop-in for /root/reference/gnn_2d.py.

Same classes, constructor signatures, sub-module names and state-dict keys as the reference
(``GNN_Layer_FS_2D`` gnn_2d.py:19-69, ``MP_PDE_Solver_2D`` :72-141), so reference checkpoints load
unchanged; the arithmetic runs in libmmpde_b200.so through ops.LayerFn / ops.SolverFn.  No PyG.
Supported configuration = the one the reference's decoder type-checks for: hidden 128, time_window 1,
no extra eq_variables (one "variables" column = time).
"""
import torch
from torch import nn

from . import ops


class BatchNorm(nn.Module):
    """Parameter holder with PyG's key layout (``norm.module.weight`` ...); applied inside the fused path."""

    def __init__(self, channels):
        super().__init__()
        self.module = nn.BatchNorm1d(channels)

    def buffers_tuple(self):
        m = self.module
        return (m.running_mean, m.running_var, m.num_batches_tracked)


def _node4(u, pos_x, pos_y, variables):
    return torch.cat((u, pos_x, pos_y, variables), dim=-1).to(torch.float32).contiguous()


def _edges_of(holder, edge_index, n_nodes):
    """Edge list prepared once per graph object (cached on it)."""
    cached = getattr(holder, "_edges", None) if holder is not None else None
    if cached is not None and cached.n_nodes == n_nodes:
        return cached
    edges = ops.EdgeList.from_edge_index(edge_index, n_nodes)
    if holder is not None:
        try:
            holder._edges = edges
        except AttributeError:
            pass
    return edges


class GNN_Layer_FS_2D(nn.Module):
    """Message passing layer: edge MLP on (x_i, x_j, u_i-u_j, pos_i-pos_j, variables_i) -> mean over
    incoming edges -> node MLP -> residual -> BatchNorm."""

    def __init__(self, in_features, out_features, hidden_features, time_window, n_variables):
        super().__init__()
        if not (in_features == out_features == hidden_features == ops.H and time_window == 1 and n_variables == 1):
            raise NotImplementedError("sm_100a kernels are built for hidden=128, time_window=1, n_variables=1")
        edge_in = 2 * in_features + time_window + 2 + n_variables
        node_in = in_features + hidden_features + n_variables
        self.message_net_1 = nn.Sequential(nn.Linear(edge_in, hidden_features), nn.ReLU())
        self.message_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features), nn.ReLU())
        self.update_net_1 = nn.Sequential(nn.Linear(node_in, hidden_features), nn.ReLU())
        self.update_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features), nn.ReLU())
        self.norm = BatchNorm(hidden_features)

    def kernel_params(self):
        return [self.message_net_1[0].weight, self.message_net_1[0].bias,
                self.message_net_2[0].weight, self.message_net_2[0].bias,
                self.update_net_1[0].weight, self.update_net_1[0].bias,
                self.update_net_2[0].weight, self.update_net_2[0].bias,
                self.norm.module.weight, self.norm.module.bias]

    def forward(self, x, u, pos_x, pos_y, variables, edge_index, batch=None, edges=None):
        if edges is None:
            edges = ops.EdgeList.from_edge_index(edge_index, x.shape[0])
        return ops.LayerFn.apply(x.contiguous(), _node4(u, pos_x, pos_y, variables), edges, self.training,
                                 self.norm.buffers_tuple(), *self.kernel_params())


class MP_PDE_Solver_2D(nn.Module):
    def __init__(self, pde, time_window=1, hidden_features=128, hidden_layer=6, eq_variables={}):
        super().__init__()
        if hidden_features != ops.H or time_window != 1 or len(eq_variables) != 0:
            raise NotImplementedError("sm_100a kernels are built for hidden=128, time_window=1, eq_variables={}")
        self.pde = pde
        self.out_features = time_window
        self.hidden_features = hidden_features
        self.hidden_layer = hidden_layer
        self.time_window = time_window
        self.eq_variables = eq_variables
        Hh = hidden_features
        self.gnn_layers = nn.ModuleList(
            GNN_Layer_FS_2D(Hh, Hh, Hh, time_window, len(eq_variables) + 1) for _ in range(hidden_layer))
        self.embedding_mlp = nn.Sequential(nn.Linear(time_window + 3 + len(eq_variables), Hh), nn.BatchNorm1d(Hh),
                                           nn.ReLU(), nn.Linear(Hh, Hh), nn.BatchNorm1d(Hh))
        self.output_mlp = nn.Sequential(nn.Conv1d(1, 4, 16, stride=3), nn.ReLU(), nn.Conv1d(4, 8, 12, stride=3),
                                        nn.ReLU(), nn.Conv1d(8, 1, 8, stride=2))

    def __repr__(self):
        return "GNN"                      # the step loops dispatch on this string (train_helper_2d.py:107)

    def _kernel_inputs(self):
        e = self.embedding_mlp
        params = [e[0].weight, e[0].bias, e[1].weight, e[1].bias, e[3].weight, e[3].bias, e[4].weight, e[4].bias]
        bufs = [(e[1].running_mean, e[1].running_var, e[1].num_batches_tracked),
                (e[4].running_mean, e[4].running_var, e[4].num_batches_tracked)]
        for layer in self.gnn_layers:
            params += layer.kernel_params()
            bufs.append(layer.norm.buffers_tuple())
        o = self.output_mlp
        params.append(torch.cat([t.reshape(-1) for t in (o[0].weight, o[0].bias, o[2].weight, o[2].bias,
                                                         o[4].weight, o[4].bias)]))
        return params, bufs

    def forward(self, data):
        u, pos = data.x, data.pos
        n = u.shape[0]
        node4 = _node4(u, pos[:, 1:2] / self.pde.Lx, pos[:, 2:3] / self.pde.Ly, pos[:, 0:1] / self.pde.tmax)
        edges = _edges_of(data, data.edge_index, n)
        params, bufs = self._kernel_inputs()
        scale = 0.1 * self.pde.dt          # cumsum(ones(1,tw) * dt * 0.1) with tw = 1 (gnn_2d.py:137-139)
        return ops.SolverFn.apply(node4, edges, self.hidden_layer, self.training, scale, bufs, *params)

    def forward_partitioned(self, parts, exch):
        """The same forward on a partitioned mesh (partition.split_graph): ``parts`` = the MeshParts living in
        this process (one per rank in a multi-GPU run, all of them in the single-process emulation), ``exch`` the
        halo exchange (dist.HaloExchange / partition.LocalExchange).  Returns one [n_own,1] tensor per part;
        concatenated in owner order they equal ``forward`` on the whole graph."""
        node4s = [_node4(p.x, p.pos[:, 1:2] / self.pde.Lx, p.pos[:, 2:3] / self.pde.Ly, p.pos[:, 0:1] / self.pde.tmax)
                  for p in parts]
        params, bufs = self._kernel_inputs()
        scale = 0.1 * self.pde.dt
        meta = [(p.edges, p.plan) for p in parts]
        return list(ops.PartitionedSolverFn.apply(meta, exch, self.hidden_layer, self.training, scale, bufs, *node4s, *params))
