"""ORACLE (test infrastructure).  Plain-torch CPU restatement of the MP-PDE processor.

Follows /root/reference/gnn_2d.py:19-69 (GNN_Layer_FS_2D) and :72-141 (MP_PDE_Solver_2D) in the
reference's own formulation (materialised [E,260] edge tensor, index_add mean), with the
un-vendored third-party semantics of SURVEY.md section 2.3:
  PyG MessagePassing.propagate : *_i = t[edge_index[1]], *_j = t[edge_index[0]]
  torch_scatter mean           : sum / clamp(count, min=1), zero for isolated nodes
  PyG BatchNorm(H)             : wrapper holding .module = nn.BatchNorm1d(H)
Module / state-dict names equal the reference's so checkpoints load unchanged.
"""
import torch
from torch import nn


class PyGBatchNorm(nn.Module):
    """torch_geometric.nn.BatchNorm 2.0.3: state-dict keys live under '.module.'"""

    def __init__(self, channels):
        super().__init__()
        self.module = nn.BatchNorm1d(channels)

    def forward(self, x):
        return self.module(x)


def scatter_mean(src, index, n):
    out = torch.zeros(n, src.shape[1], dtype=src.dtype, device=src.device).index_add_(0, index, src)
    cnt = torch.zeros(n, dtype=src.dtype, device=src.device).index_add_(
        0, index, torch.ones_like(index, dtype=src.dtype))
    return out / cnt.clamp(min=1).unsqueeze(1)


class GNN_Layer_FS_2D(nn.Module):
    def __init__(self, in_features, out_features, hidden_features, time_window, n_variables):
        super().__init__()
        e_in = 2 * in_features + time_window + 2 + n_variables          # gnn_2d.py:38
        self.message_net_1 = nn.Sequential(nn.Linear(e_in, hidden_features), nn.ReLU())
        self.message_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features), nn.ReLU())
        n_in = in_features + hidden_features + n_variables               # gnn_2d.py:44
        self.update_net_1 = nn.Sequential(nn.Linear(n_in, hidden_features), nn.ReLU())
        self.update_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features), nn.ReLU())
        self.norm = PyGBatchNorm(hidden_features)

    def message(self, x, u, pos_x, pos_y, variables, edge_index):
        j, i = edge_index[0], edge_index[1]
        feats = torch.cat((x[i], x[j], u[i] - u[j], pos_x[i] - pos_x[j], pos_y[i] - pos_y[j],
                           variables[i]), dim=-1)                         # gnn_2d.py:61
        return self.message_net_2(self.message_net_1(feats))

    def forward(self, x, u, pos_x, pos_y, variables, edge_index, batch=None):
        m = self.message(x, u, pos_x, pos_y, variables, edge_index)
        agg = scatter_mean(m, edge_index[1], x.shape[0])
        upd = self.update_net_2(self.update_net_1(torch.cat((x, agg, variables), dim=-1)))
        return self.norm(x + upd)                                         # gnn_2d.py:56,69


class MP_PDE_Solver_2D(nn.Module):
    def __init__(self, pde, time_window=1, hidden_features=128, hidden_layer=6, eq_variables={}):
        super().__init__()
        self.pde = pde
        self.out_features = time_window
        self.hidden_features = hidden_features
        self.hidden_layer = hidden_layer
        self.time_window = time_window
        self.eq_variables = eq_variables
        H = hidden_features
        self.gnn_layers = nn.ModuleList(
            GNN_Layer_FS_2D(H, H, H, time_window, len(eq_variables) + 1) for _ in range(hidden_layer))
        self.embedding_mlp = nn.Sequential(
            nn.Linear(time_window + 3 + len(eq_variables), H), nn.BatchNorm1d(H), nn.ReLU(),
            nn.Linear(H, H), nn.BatchNorm1d(H))
        self.output_mlp = nn.Sequential(
            nn.Conv1d(1, 4, 16, stride=3), nn.ReLU(),
            nn.Conv1d(4, 8, 12, stride=3), nn.ReLU(),
            nn.Conv1d(8, 1, 8, stride=2))

    def __repr__(self):
        return "GNN"

    def forward(self, data, return_hidden=False):
        u, pos = data.x, data.pos
        pos_x = pos[:, 1:2] / self.pde.Lx
        pos_y = pos[:, 2:3] / self.pde.Ly
        variables = pos[:, 0:1] / self.pde.tmax
        h = self.embedding_mlp(torch.cat((u, pos_x, pos_y, variables), -1))
        hidden = [h]
        for layer in self.gnn_layers:
            h = layer(h, u, pos_x, pos_y, variables, data.edge_index, data.batch)
            hidden.append(h)
        diff = self.output_mlp(h[:, None]).squeeze(1)
        dt = torch.cumsum(torch.ones(1, self.time_window, dtype=h.dtype) * self.pde.dt * 0.1, dim=1)
        out = dt.to(h.device) * diff
        return (out, hidden) if return_hidden else out
