"""ORACLE (test infrastructure).  ItpNet restatement, /root/reference/interpolate.py:5-98.
Two tanh MLPs (2*30+2 -> layers -> 30) giving un-normalised neighbour weights for modes '1'/'2',
an unused-but-registered third MLP (layers3), and the 'res_cut' residual net ('down')."""
import torch
from torch import nn


def _mlp_list(sizes):
    return nn.ModuleList(nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:]))


class ItpNet(nn.Module):
    def __init__(self, ori_nx, ori_ny, layers1, layers2, layers3, normalize=False):
        super().__init__()
        assert not normalize, "normalize=True is broken in the reference (interpolate.py:17-22,81-85)"
        self.n = 30
        self.layers = _mlp_list([2 * self.n + 2] + list(layers1) + [self.n])
        self.layers2 = _mlp_list([2 * self.n + 2] + list(layers2) + [self.n])
        n_grid = ori_nx * ori_ny if ori_ny is not None else ori_nx
        self.layers3 = _mlp_list([n_grid] + list(layers3) + [n_grid])
        if ori_ny is not None:
            mods = []
            for a, b in zip(layers3[:-1], layers3[1:]):
                mods += [nn.Conv2d(a, b, 5, padding=2), nn.Tanh()]
            self.down = nn.Sequential(*mods)
        else:
            self.down = nn.Sequential(nn.Linear(ori_nx, 2048), nn.Tanh(), nn.Linear(2048, 512), nn.Tanh(),
                                      nn.Linear(512, 2048), nn.Tanh(), nn.Linear(2048, ori_nx))

    @staticmethod
    def _run(stack, z):
        for li, lin in enumerate(stack):
            z = lin(z)
            if li != len(stack) - 1:
                z = torch.tanh(z)
        return z

    def forward(self, neighbors, query_points, mode, data=None):
        if mode in ("1", "2"):
            z = torch.cat((neighbors, query_points), dim=-2)
            z = z.reshape(neighbors.shape[0], neighbors.shape[1], -1)      # interpolate.py:80
            return self._run(self.layers if mode == "1" else self.layers2, z)
        if mode == "res_cut":
            return self.down(data)
        return data
