"""ORACLE (test infrastructure).  DMM mesh mover, restating /root/reference/mesh/dmm_model.py:9-234
without PyG and without the hard-coded device="cuda" (:27-28), so it also runs on CPU.
State-dict keys equal the reference's (including the unused ``fc0`` of DenseNet, :29)."""
import torch
from torch import nn

from .knn import knn_graph
from .processor import PyGBatchNorm, scatter_mean


class DenseNet(nn.Module):
    def __init__(self, layers, width=32, normalize=False):
        super().__init__()
        assert not normalize and len(layers) >= 2
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(layers[:-1], layers[1:]))
        self.fc0 = nn.Linear(4, width)          # registered but unused (dmm_model.py:29)

    def forward(self, x):
        for lin in self.layers[:-1]:
            x = torch.tanh(lin(x))
        return self.layers[-1](x), x


class ConvNet(nn.Module):
    def __init__(self, s, layers):
        super().__init__()
        assert layers == 7
        self.layers = nn.ModuleList([nn.Conv2d(1, 8, 5, stride=2, padding=2), nn.Conv2d(8, 16, 5, padding=2),
                                     nn.Conv2d(16, 8, 5, padding=2), nn.Conv2d(8, 1, 5, stride=2, padding=2)])
        self.fc2 = nn.Linear(int(((s + 1) / 2 + 1) / 2) ** 2, 1024)
        self.fc3 = nn.Linear(1024, 512)

    def forward(self, x):
        first = torch.tanh(self.layers[0](x))
        x = torch.tanh(self.layers[1](first))
        x = torch.tanh(first + self.layers[2](x))       # skip from layer 0 (dmm_model.py:70-73)
        x = torch.tanh(self.layers[3](x))
        x = torch.tanh(self.fc2(torch.flatten(x, 1)))
        return self.fc3(x)


class _TanhLayer(nn.Module):
    """mesh/dmm_model.py:94-142 -- tanh message-passing layer without the 'variables' input."""

    def __init__(self, in_features, out_features, hidden_features):
        super().__init__()
        self.message_net_1 = nn.Sequential(nn.Linear(2 * in_features + 3, hidden_features), nn.Tanh())
        self.message_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features), nn.Tanh())
        self.update_net_1 = nn.Sequential(nn.Linear(in_features + hidden_features, hidden_features), nn.Tanh())
        self.update_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features), nn.Tanh())
        self.norm = PyGBatchNorm(hidden_features)

    def forward(self, x, u, pos_x, pos_y, edge_index, batch=None):
        j, i = edge_index[0], edge_index[1]
        m = torch.cat((x[i], x[j], u[i] - u[j], pos_x[i] - pos_x[j], pos_y[i] - pos_y[j]), dim=-1)
        m = self.message_net_2(self.message_net_1(m))
        agg = scatter_mean(m, i, x.shape[0])
        return self.norm(x + self.update_net_2(self.update_net_1(torch.cat((x, agg), dim=-1))))


class DMM(nn.Module):
    def __init__(self, branch_layer, trunk_layer, grid=None, out_layer=None, s=None, mode="array"):
        super().__init__()
        self.mode = mode
        self.ori_grid = grid
        if mode == "array":
            self.branch = ConvNet(s, branch_layer)
        else:
            H, L = branch_layer
            self.hidden_features, self.hidden_layer = H, L
            self.gnn_layers = nn.ModuleList(_TanhLayer(H, H, H) for _ in range(L))
            self.embedding_mlp = nn.Sequential(nn.Linear(3, H), nn.BatchNorm1d(H), nn.Tanh(),
                                               nn.Linear(H, H), nn.BatchNorm1d(H))
            self.decoding_mlp = DenseNet([H, 128, 1])
            self.output_mlp = nn.Sequential(nn.Linear(grid.shape[0], 512), nn.Tanh(), nn.Linear(512, 256),
                                            nn.Tanh(), nn.Linear(256, trunk_layer[-1]))
        self.trunk = DenseNet(trunk_layer)
        self.out_nn = DenseNet(out_layer)
        self._edge_cache = {}

    def _branch_graph(self, u):
        B, n = u.shape[0], self.ori_grid.shape[0]
        grid = self.ori_grid.to(u.device)
        pos = grid[None].repeat(B, 1, 1).reshape(-1, 2)
        if B not in self._edge_cache:                    # static topology (dmm_model.py:222-234)
            batch = torch.arange(B).repeat_interleave(n)
            self._edge_cache[B] = knn_graph(pos.cpu(), 35, batch).to(u.device)
        ei = self._edge_cache[B]
        x = u.reshape(-1, 1)
        px, py = pos[:, 0:1], pos[:, 1:2]
        h = self.embedding_mlp(torch.cat((x, px, py), -1))
        for layer in self.gnn_layers:
            h = layer(h, x, px, py, ei)
        h, _ = self.decoding_mlp(h)
        return self.output_mlp(h.reshape(B, 1, -1))      # [B,1,latent]

    def forward(self, u, grid, rf=False):
        per = grid.shape[0] // u.shape[0]
        if self.mode == "array":
            branch = self.branch(u.unsqueeze(1)).unsqueeze(1)
        else:
            branch = self._branch_graph(u)
        branch = branch.repeat(1, per, 1)
        trunk, _ = self.trunk(grid)
        out, hidden = self.out_nn(torch.cat((branch.reshape(-1, branch.shape[-1]),
                                             trunk.reshape(-1, branch.shape[-1])), dim=-1))
        if not rf:
            return out
        return out, hidden, torch.ones_like(hidden).type_as(trunk).reshape(-1, 1)
