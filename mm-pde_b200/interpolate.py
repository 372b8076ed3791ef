"""Learned interpolation network -- drop-in for /root/reference/interpolate.py (``ItpNet`` :5-98).

Parameters / state-dict keys as in the reference (``layers``, ``layers2``, the registered-but-unused
``layers3``, ``down``).  The hot path (GraphCreator_FS_2D.interpolate) never materialises the
[nu,Q,30,2] neighbour tensor: it hands ``flat_params(mode)`` to the fused sm_100a kernel
(ops.InterpolateFn).  ``forward`` keeps the reference signature for callers that pass explicit
neighbour tensors; ``res_cut`` (dense regular Conv2d / MLP, interpolate.py:54-74) stays in cuDNN/cuBLAS
as SURVEY.md T12 prescribes.
"""
import os

import torch
from torch import nn

FP32_RES_CUT = os.environ.get("MMPDE_FP32_RES_CUT", "0") == "1"
FUSED_RES_CUT = os.environ.get("MMPDE_FUSED_RES_CUT", "1") != "0"


def _linears(widths):
    return nn.ModuleList(nn.Linear(i, o) for i, o in zip(widths[:-1], widths[1:]))


class ItpNet(nn.Module):
    def __init__(self, ori_nx, ori_ny, layers1, layers2, layers3, normalize=False):
        super().__init__()
        if normalize:
            raise NotImplementedError("normalize=True is unusable in the reference (interpolate.py:81-85)")
        self.n = 30
        feat = 2 * self.n + 2
        self.layers = _linears([feat, *layers1, self.n])
        self.layers2 = _linears([feat, *layers2, self.n])
        cells = ori_nx * ori_ny if ori_ny is not None else ori_nx
        self.layers3 = _linears([cells, *layers3, cells])            # never used in forward, kept for checkpoints
        if ori_ny is not None:
            stack = []
            for cin, cout in zip(layers3[:-1], layers3[1:]):
                stack.extend((nn.Conv2d(cin, cout, 5, padding=2), nn.Tanh()))
            self.down = nn.Sequential(*stack)
        else:
            self.down = nn.Sequential(nn.Linear(ori_nx, 2048), nn.Tanh(), nn.Linear(2048, 512), nn.Tanh(),
                                      nn.Linear(512, 2048), nn.Tanh(), nn.Linear(2048, ori_nx))

    def _fused_res_cut(self, data):
        """The reference's regular-grid configuration (layers3 = [1,4,16,4,1], mmpde.py:343) on the GPU runs as one
        tile-resident fp32 kernel per direction (csrc/rescut.cu); anything else stays in cuDNN / cuBLAS."""
        convs = [m for m in self.down if isinstance(m, nn.Conv2d)]
        return (FUSED_RES_CUT and torch.is_tensor(data) and data.is_cuda and data.dim() == 4 and data.dtype == torch.float32
                and not data.requires_grad and len(convs) == 4 and len(self.down) == 8
                and tuple([convs[0].in_channels] + [c.out_channels for c in convs]) == (1, 4, 16, 4, 1))

    def _stack(self, mode):
        return self.layers if mode == "1" else self.layers2

    def flat_params(self, mode):
        """Wa ba Wb bb Wc bc flattened in the order the fused kernel expects (62->128->64->30 only)."""
        stack = self._stack(mode)
        shapes = [tuple(l.weight.shape) for l in stack]
        if shapes != [(128, 62), (64, 128), (30, 64)]:
            raise NotImplementedError(f"fused interpolation kernel is built for 62->128->64->30, got {shapes}")
        return torch.cat([t.reshape(-1) for l in stack for t in (l.weight, l.bias)])

    def forward(self, neighbors, query_points, mode, data=None):
        if mode in ("1", "2"):
            z = torch.cat((neighbors, query_points), dim=-2).flatten(start_dim=-2)
            stack = self._stack(mode)
            for li, lin in enumerate(stack):
                z = lin(z)
                if li + 1 < len(stack):
                    z = torch.tanh(z)
            return z
        if mode == "res_cut":
            if self._fused_res_cut(data):
                from . import ops
                return ops.ResCutFn.apply(data, *[t for m in self.down if isinstance(m, nn.Conv2d) for t in (m.weight, m.bias)])
            if FP32_RES_CUT and data.is_cuda and isinstance(self.down[0], nn.Conv2d):
                # cuDNN runs these 5x5 convolutions in TF32 by default (torch.backends.cudnn.allow_tf32, as on the
                # reference's own GPU path): ~4e-4 of the step's 1e-3 tolerance.  MMPDE_FP32_RES_CUT=1
                # forces fp32 (cuDNN's fp32 kernels for this shape cost ~0.25 ms per training step).
                with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                    return self.down(data)
            return self.down(data)
        return data
