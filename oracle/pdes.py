"""ORACLE (test infrastructure).  PDE constants, restating /root/reference/PDEs.py:20-67.
dt is frozen at construction from the DEFAULT grid (PDEs.py:40,66) and is not recomputed when the
driver overwrites grid_size (mmpde.py:179-181) -- SURVEY.md appendix C.2."""
import torch
from torch import nn


class _PDE(nn.Module):
    def __repr__(self):
        return "PDE"


class burgers(_PDE):
    def __init__(self, tmin=None, tmax=None, grid_size=None, L=None, flux_splitting=None, device="cpu"):
        super().__init__()
        self.tmin = 0 if tmin is None else tmin
        self.tmax = 30 if tmax is None else tmax
        self.Lx = self.Ly = 1 if L is None else L
        default = (31, 96, 96)
        self.grid_size = default if grid_size is None else grid_size
        self.movingmesh_grid_size = default
        self.ori_grid_size = default
        self.dt = self.tmax / (self.grid_size[0] - 1)
        self.device = device


class cy(_PDE):
    def __init__(self, tmin=None, tmax=None, grid_size=None, ori_grid=None, L=None,
                 flux_splitting=None, device="cpu"):
        super().__init__()
        self.tmin = 0 if tmin is None else tmin
        self.tmax = 2.9 if tmax is None else tmax
        self.Lx = self.Ly = 1 if L is None else L
        g = (30, 2521) if grid_size is None else grid_size
        self.grid_size = self.ori_grid_size = self.movingmesh_grid_size = g
        self.ori_grid = ori_grid
        self.dt = self.tmax / (self.grid_size[0] - 1)
        self.device = device
