"""PDE descriptors consumed by the graph creator and the solver (mirror of /root/reference/PDEs.py:9-67).

Same attribute names and defaults; ``dt`` is fixed at construction from the default grid and is NOT
recomputed when the driver overwrites ``grid_size`` (reference quirk)."""
from torch import nn


class PDE(nn.Module):
    """Parameter-less base: carries domain constants only."""

    def __repr__(self):
        return "PDE"

    def _common(self, tmin, tmax, default_tmax, L, device):
        self.tmin = tmin if tmin is not None else 0
        self.tmax = tmax if tmax is not None else default_tmax
        self.Lx = self.Ly = L if L is not None else 1
        self.device = device


class burgers(PDE):
    DEFAULT_GRID = (31, 96, 96)

    def __init__(self, tmin=None, tmax=None, grid_size=None, L=None, flux_splitting=None, device="cpu"):
        super().__init__()
        self._common(tmin, tmax, 30, L, device)
        self.grid_size = grid_size if grid_size is not None else self.DEFAULT_GRID
        self.movingmesh_grid_size = self.DEFAULT_GRID
        self.ori_grid_size = self.DEFAULT_GRID
        self.dt = self.tmax / (self.grid_size[0] - 1)


class cy(PDE):
    DEFAULT_GRID = (30, 2521)

    def __init__(self, tmin=None, tmax=None, grid_size=None, ori_grid=None, L=None, flux_splitting=None,
                 device="cpu"):
        super().__init__()
        self._common(tmin, tmax, 2.9, L, device)
        grid = grid_size if grid_size is not None else self.DEFAULT_GRID
        self.grid_size = self.ori_grid_size = self.movingmesh_grid_size = grid
        self.ori_grid = ori_grid
        self.dt = self.tmax / (self.grid_size[0] - 1)
