"""Shared by make_golden.py (generator, build container only) and the tests (everywhere).

``fill_params`` gives every parameter a value that depends only on (seed, parameter name, shape), so
fixtures need not store multi-MB state dicts: the reference module at generation time and the
oracle / product module at test time are filled by the same rule.  Values imitate torch's default
Linear/Conv init scale: U(-1/sqrt(fan_in), 1/sqrt(fan_in)); BatchNorm affine ~ 1 +- 0.1 / +-0.1."""
import math
import zlib

import numpy as np
import torch


def fill_params(module, seed):
    with torch.no_grad():
        for name, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
            if p.dim() >= 2:
                fan_in = int(np.prod(p.shape[1:]))
                bound = 1.0 / math.sqrt(fan_in)
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
            elif ".module.weight" in name or (name.endswith("weight") and p.dim() == 1):
                p.copy_(1.0 + 0.1 * (torch.rand(p.shape, generator=g) * 2 - 1))
            else:
                p.copy_(0.1 * (torch.rand(p.shape, generator=g) * 2 - 1))
    return module


def synth_fields(B, T, nx, ny, seed):
    """Burgers-like smooth fields [B,T,nx,ny]: 4 travelling Fourier modes per trajectory."""
    g = torch.Generator().manual_seed(seed)
    xs = torch.linspace(0, 1, nx)[:, None]
    ys = torch.linspace(0, 1, ny)[None, :]
    out = torch.zeros(B, T, nx, ny)
    for b in range(B):
        for _ in range(4):
            a = torch.rand(1, generator=g) * 2 - 1
            kx, ky = torch.randint(1, 4, (2,), generator=g)
            ph = torch.rand(2, generator=g) * 6.28
            for t in range(T):
                out[b, t] += a * torch.sin(2 * math.pi * kx * xs + ph[0] + 0.1 * t) \
                    * torch.cos(2 * math.pi * ky * ys + ph[1])
    return out


class SmoothMover(torch.nn.Module):
    """Analytic stand-in for a trained DMM: a potential phi(u, xi) whose gradient is a smooth,
    u-dependent displacement.  The displacement is deliberately non-zero and asymmetric everywhere
    (boundary included) so that no query is equidistant from two lattice nodes: exact distance ties
    are implementation-defined in sklearn's kd-tree and would make the
    fixture depend on its traversal order.  Same call signature as DMM.forward
    (/root/reference/mesh/dmm_model.py:185)."""

    def __init__(self, amp=0.02):
        super().__init__()
        self.amp = amp

    def forward(self, u, grid, rf=False):
        per = grid.shape[0] // u.shape[0]
        s = u.reshape(u.shape[0], -1).mean(dim=1, keepdim=True).repeat(1, per).reshape(-1, 1)
        x, y = grid[:, 0:1], grid[:, 1:2]
        return self.amp * (1 + 0.5 * torch.tanh(s)) * (torch.sin(1.3 * x + 0.4) * torch.cos(0.9 * y + 0.2)
                                                        + 0.3 * torch.sin(2.1 * x * y + 0.7) + 0.11 * x + 0.07 * y)
