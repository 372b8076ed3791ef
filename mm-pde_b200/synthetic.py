"""Seeded synthetic inputs of the shapes BASELINE.json names (the datasets and DMM checkpoints of the
reference live on Google Drive, /root/reference/README.md:19, and there is no network).
Used by bench.py, the tests and ``mmpde.py --synthetic``.  CPU tensors; callers move them."""
import math

import numpy as np
import torch


def burgers_fields(n_traj, n_t=31, nx=48, ny=48, seed=0):
    """[traj, n_t, nx, ny] smooth Burgers-like fields: 4 travelling Fourier modes, amplitude U(-1,1)."""
    g = torch.Generator().manual_seed(seed)
    xs = torch.linspace(0, 1, nx)[None, None, :, None]
    ys = torch.linspace(0, 1, ny)[None, None, None, :]
    ts = torch.arange(n_t, dtype=torch.float32)[None, :, None, None]
    out = torch.zeros(n_traj, n_t, nx, ny)
    for _ in range(4):
        amp = (torch.rand(n_traj, 1, 1, 1, generator=g) * 2 - 1)
        kx = torch.randint(1, 4, (n_traj, 1, 1, 1), generator=g).float()
        ky = torch.randint(1, 4, (n_traj, 1, 1, 1), generator=g).float()
        ph = torch.rand(n_traj, 2, 1, 1, generator=g) * 2 * math.pi
        out += amp * torch.sin(2 * math.pi * kx * xs + ph[:, 0:1] + 0.1 * ts) * torch.cos(2 * math.pi * ky * ys + ph[:, 1:2])
    return out


def cylinder_cloud(n=2521, seed=0):
    """[n,2] unstructured points in the unit square: jittered lattice with a disc (the cylinder) removed."""
    rng = np.random.default_rng(seed)
    side = int(math.ceil(math.sqrt(n * 1.12)))
    g = (np.stack(np.meshgrid(np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 2) + 0.5) / side
    g = g + rng.uniform(-0.3, 0.3, g.shape) / side
    keep = np.hypot(g[:, 0] - 0.3, g[:, 1] - 0.5) > 0.08
    g = g[keep]
    assert len(g) >= n, (len(g), n)
    g = g[rng.permutation(len(g))[:n]]
    return torch.tensor(g, dtype=torch.float32)


def cylinder_fields(n_traj, cloud, n_t=30, seed=0):
    g = torch.Generator().manual_seed(seed)
    x, y = cloud[:, 0][None, None], cloud[:, 1][None, None]
    ts = torch.arange(n_t, dtype=torch.float32)[None, :, None]
    out = torch.zeros(n_traj, n_t, cloud.shape[0])
    for _ in range(4):
        amp = torch.rand(n_traj, 1, 1, generator=g) * 2 - 1
        k = torch.randint(1, 4, (n_traj, 2, 1), generator=g).float()
        ph = torch.rand(n_traj, 2, 1, generator=g) * 2 * math.pi
        out += amp * torch.sin(2 * math.pi * k[:, 0:1] * x + ph[:, 0:1] + 0.2 * ts) * torch.cos(2 * math.pi * k[:, 1:2] * y + ph[:, 1:2])
    return out


def jittered_lattice(n_side, jitter=0.3, seed=0):
    """[n_side^2, 2] lattice on [0,1]^2 jittered by U(-jitter,jitter)*h: no exact distance ties (config C4/C5)."""
    rng = np.random.default_rng(seed)
    h = 1.0 / (n_side - 1)
    g = np.stack(np.meshgrid(np.linspace(0, 1, n_side), np.linspace(0, 1, n_side), indexing="ij"), -1).reshape(-1, 2)
    return torch.tensor(g + rng.uniform(-jitter, jitter, g.shape) * h, dtype=torch.float32)


class AnalyticMover(torch.nn.Module):
    """Stand-in for a trained DMM when no checkpoint exists: a smooth potential phi(u, xi) whose gradient is a
    u-dependent displacement of about ``amp`` (fraction of the domain), non-zero everywhere so that moved
    nodes are in general position.  Same call signature as DMM.forward (mesh/dmm_model.py:185)."""

    def __init__(self, amp=0.004):
        super().__init__()
        self.amp = amp

    def forward(self, u, grid, rf=False):
        per = grid.shape[0] // u.shape[0]
        s = u.reshape(u.shape[0], -1).mean(dim=1, keepdim=True).repeat(1, per).reshape(-1, 1)
        x, y = grid[:, 0:1], grid[:, 1:2]
        return self.amp * (1 + 0.5 * torch.tanh(s)) * (torch.sin(1.3 * x + 0.4) * torch.cos(0.9 * y + 0.2)
                                                        + 0.3 * torch.sin(2.1 * x * y + 0.7) + 0.11 * x + 0.07 * y)
