"""Frozen DMM mesh mover -- PyG-free, device-agnostic counterpart of /root/reference/mesh/dmm_model.py.

Adjacent to the hot path: it supplies the moved mesh and runs as plain PyTorch
(cuDNN/cuBLAS), as the survey prescribes; it is not one of the hand-written kernels.  State-dict keys
match the reference (including DenseNet's unused ``fc0``, :29) so its checkpoints load unchanged; the
reference's hard-coded ``device="cuda"`` tensors (:27-28) are dropped because forward never reads them.
The graph-mode branch reuses this repo's CUDA k-NN for its static 35-NN graph (:222-234).

``DMM.displacement`` gives the mesh displacement grad_xi phi(u, xi) ANALYTICALLY (forward-mode Jacobian of trunk +
out_nn) instead of the reference's two ``autograd.grad(create_graph=True)`` calls
(/root/reference/data_creator_2d.py:106-107): the double-backward graph those build only feeds the frozen mover
(SURVEY.md 8f-1, appendix C.10), and an autograd call inside the step cannot be recorded into the step's CUDA graph.
"""
import torch
from torch import nn

from .. import ops


class DenseNet(nn.Module):
    def __init__(self, layers, width=32, normalize=False):
        super().__init__()
        if normalize or len(layers) < 2:
            raise NotImplementedError
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(layers[:-1], layers[1:]))
        self.width = width
        self.fc0 = nn.Linear(4, width)

    def forward(self, x):
        hidden = x
        for lin in list(self.layers)[:-1]:
            hidden = torch.tanh(lin(hidden))
        return self.layers[-1](hidden), hidden


class ConvNet(nn.Module):
    def __init__(self, s, layers):
        super().__init__()
        if layers != 7:
            raise NotImplementedError("only the 7-layer branch of the reference is defined (:53-60)")
        self.layers = nn.ModuleList([nn.Conv2d(1, 8, 5, stride=2, padding=2), nn.Conv2d(8, 16, 5, padding=2),
                                     nn.Conv2d(16, 8, 5, padding=2), nn.Conv2d(8, 1, 5, stride=2, padding=2)])
        self.fc1 = None
        self.fc2 = nn.Linear(int(((s + 1) / 2 + 1) / 2) ** 2, 1024)
        self.fc3 = nn.Linear(1024, 512)

    def forward(self, x):
        stem = torch.tanh(self.layers[0](x))
        y = torch.tanh(self.layers[1](stem))
        y = torch.tanh(stem + self.layers[2](y))
        y = torch.tanh(self.layers[3](y))
        return self.fc3(torch.tanh(self.fc2(y.flatten(1))))


class _Norm(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.module = nn.BatchNorm1d(c)

    def forward(self, x):
        return self.module(x)


class GNN_Layer_FS_2D(nn.Module):
    """tanh message-passing layer of the graph-mode branch (:94-142); hidden width 4, so plain torch."""

    def __init__(self, in_features, out_features, hidden_features):
        super().__init__()
        self.message_net_1 = nn.Sequential(nn.Linear(2 * in_features + 3, hidden_features), nn.Tanh())
        self.message_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features), nn.Tanh())
        self.update_net_1 = nn.Sequential(nn.Linear(in_features + hidden_features, hidden_features), nn.Tanh())
        self.update_net_2 = nn.Sequential(nn.Linear(hidden_features, out_features), nn.Tanh())
        self.norm = _Norm(hidden_features)

    def packed(self):
        """W1 b1 W2 b2 W3 b3 W4 b4 flattened in the order mmpde_dmm_gnn_layer expects."""
        return torch.cat([t.reshape(-1) for seq in (self.message_net_1, self.message_net_2, self.update_net_1, self.update_net_2)
                          for t in (seq[0].weight, seq[0].bias)])

    def forward_fused(self, x, upos, row_ptr, src32):
        """The same layer as ONE pass over the edge list (csrc/dmm_gnn.cu); forward only -- the mover is frozen."""
        from .. import _cabi
        w = self.packed()
        assert w.numel() == 124, "the fused layer is built for hidden width 4"
        out = torch.empty_like(x)
        with ops._on(x):
            _cabi.call("mmpde_dmm_gnn_layer", ops._ptr(x), ops._ptr(upos), ops._ptr(row_ptr), ops._ptr(src32), x.shape[0],
                       ops._ptr(w), ops._ptr(out), ops._stream())
        return self.norm(out)

    def forward(self, x, u, pos_x, pos_y, src, dst, inv_deg):
        feats = torch.cat((x[dst], x[src], u[dst] - u[src], pos_x[dst] - pos_x[src], pos_y[dst] - pos_y[src]), -1)
        msg = self.message_net_2(self.message_net_1(feats))
        agg = torch.zeros_like(x).index_add_(0, dst, msg) * inv_deg[:, None]
        return self.norm(x + self.update_net_2(self.update_net_1(torch.cat((x, agg), -1))))


class DMM(nn.Module):
    def __init__(self, branch_layer, trunk_layer, grid=None, out_layer=None, s=None, mode="array"):
        super().__init__()
        self.mode = mode
        self.ori_grid = grid
        if mode == "array":
            self.branch = ConvNet(s, branch_layer)
        elif mode == "graph":
            self.hidden_features, self.hidden_layer = branch_layer[0], branch_layer[1]
            Hd = self.hidden_features
            self.gnn_layers = nn.ModuleList(GNN_Layer_FS_2D(Hd, Hd, Hd) for _ in range(self.hidden_layer))
            self.embedding_mlp = nn.Sequential(nn.Linear(3, Hd), nn.BatchNorm1d(Hd), nn.Tanh(), nn.Linear(Hd, Hd),
                                               nn.BatchNorm1d(Hd))
            self.decoding_mlp = DenseNet([Hd, 128, 1])
            self.output_mlp = nn.Sequential(nn.Linear(grid.shape[0], 512), nn.Tanh(), nn.Linear(512, 256), nn.Tanh(),
                                            nn.Linear(256, trunk_layer[-1]))
        else:
            raise ValueError(mode)
        self.trunk = DenseNet(trunk_layer)
        self.out_nn = DenseNet(out_layer)
        self._graph_cache = {}
        self.fused = True            # graph branch through csrc/dmm_gnn.cu when no autograd graph is being built

    def _static_graph(self, n_samples, device):
        key = (n_samples, str(device))
        if key not in self._graph_cache:
            n = self.ori_grid.shape[0]
            pts = self.ori_grid.to(device=device, dtype=torch.float32)[None].expand(n_samples, n, 2).reshape(-1, 2).contiguous()
            off = torch.arange(n_samples + 1, dtype=torch.int32, device=device) * n
            nbr = ops.knn_indices(pts, off, pts, off, 35, rule=0, exclude_self=True)
            e = ops.EdgeList.from_knn(nbr, has_pad=n - 1 < 35)
            deg = torch.bincount(e.dst.long(), minlength=n_samples * n)
            row_ptr = torch.cat((deg.new_zeros(1), deg.cumsum(0))).to(torch.int32)
            self._graph_cache[key] = (pts, e.src.long(), e.dst.long(), e.inv_deg, row_ptr, e.src.contiguous())
        return self._graph_cache[key]

    def _branch_graph(self, u):
        pts, src, dst, inv_deg, row_ptr, src32 = self._static_graph(u.shape[0], u.device)
        x = u.reshape(-1, 1)
        px, py = pts[:, 0:1], pts[:, 1:2]
        h = self.embedding_mlp(torch.cat((x, px, py), -1))
        # frozen mover (no autograd graph wanted), hidden width 4, on the GPU: each layer is one fused pass over the edges
        fused = self.fused and u.is_cuda and self.hidden_features == 4 and not torch.is_grad_enabled()
        if fused:
            upos = torch.cat((x, pts, torch.zeros_like(x)), -1).to(torch.float32).contiguous()
        for layer in self.gnn_layers:
            h = layer.forward_fused(h.contiguous(), upos, row_ptr, src32) if fused else layer(h, x, px, py, src, dst, inv_deg)
        h, _ = self.decoding_mlp(h)
        return self.output_mlp(h.reshape(u.shape[0], 1, -1))

    def forward(self, u, grid, rf=False):
        per_sample = grid.shape[0] // u.shape[0]
        latent = self.branch(u.unsqueeze(1)).unsqueeze(1) if self.mode == "array" else self._branch_graph(u)
        latent = latent.expand(-1, per_sample, -1).reshape(-1, latent.shape[-1])
        trunk, _ = self.trunk(grid)
        out, hidden = self.out_nn(torch.cat((latent, trunk), dim=-1))
        if rf:
            return out, hidden, torch.ones_like(hidden).type_as(trunk).reshape(-1, 1)
        return out

    def _latent(self, u):
        return self.branch(u.unsqueeze(1)).unsqueeze(1) if self.mode == "array" else self._branch_graph(u)

    @torch.no_grad()
    def displacement(self, u, grid):
        """(d phi / d xi_1, d phi / d xi_2), each [N,1], for grid = xi [N,2] (N = samples * nodes per sample).

        Only the two-layer tanh trunk and the two-layer tanh out_nn depend on xi:
            a = tanh(W1 xi + b1)                          [N,32]     trunk.layers[0]
            trunk = W2 a + b2                             [N,512]    trunk.layers[1]
            z = Wl latent + Wt trunk + bo ,  h = tanh(z)  [N,512]    out_nn.layers[0] = [Wl | Wt]
            phi = w h + c                                 [N,1]      out_nn.layers[1]
        so with M = Wt W2 ([512,32], weights only):  z = (Wl latent + Wt b2 + bo)[sample] + a M^T  and
            d phi / d xi_d = ((1 - h^2) * (((1 - a^2) * W1[:, d]) M^T)) . w
        i.e. one [3N,32] x [32,512] product for the value and both directional derivatives.  Deeper trunk / out stacks
        (not used by the reference's configurations, mmpde.py:199, README.md:31) fall back to autograd."""
        if len(self.trunk.layers) != 2 or len(self.out_nn.layers) != 2:
            return None
        per_sample = grid.shape[0] // u.shape[0]
        t1, t2 = self.trunk.layers
        o1, o2 = self.out_nn.layers
        n_lat = o1.weight.shape[1] - t2.weight.shape[0]
        Wl, Wt = o1.weight[:, :n_lat], o1.weight[:, n_lat:]
        latent = self._latent(u).reshape(u.shape[0], -1)                        # [B, n_lat]
        const = latent @ Wl.t() + (Wt @ t2.bias + o1.bias)                      # [B, 512]
        M = Wt @ t2.weight                                                      # [512, 32]
        K, J = t1.weight.shape[0], M.shape[0]
        if self.fused and grid.is_cuda and K <= 32 and J % 4 == 0 and J <= 1024 and grid.dtype == torch.float32:
            # one pass over the points (csrc/dmm_gnn.cu): the [3N,32] x [32,512] product and the [N,512] intermediates of
            # the tensor-op form below never exist
            from .. import _cabi
            xi = grid.contiguous()
            out = torch.empty(xi.shape[0], 2, dtype=torch.float32, device=xi.device)
            with ops._on(xi):
                _cabi.call("mmpde_dmm_displacement", ops._ptr(xi), ops._ptr(t1.weight.contiguous()), ops._ptr(t1.bias), K,
                           ops._ptr(M.contiguous()), ops._ptr(const.contiguous()), ops._ptr(o2.weight.reshape(-1).contiguous()), J,
                           xi.shape[0], per_sample, ops._ptr(out), ops._stream())
            return out[:, 0:1], out[:, 1:2]
        a = torch.tanh(grid @ t1.weight.t() + t1.bias)                          # [N, 32]
        da = 1.0 - a * a
        stacked = torch.cat((a, da * t1.weight[:, 0], da * t1.weight[:, 1])) @ M.t()     # [3N, 512]
        N = grid.shape[0]
        z = stacked[:N].reshape(u.shape[0], per_sample, -1) + const[:, None, :]
        h = torch.tanh(z).reshape(N, -1)
        gate = (1.0 - h * h) * o2.weight.reshape(1, -1)                         # [N, 512]
        return (gate * stacked[N:2 * N]).sum(1, keepdim=True), (gate * stacked[2 * N:]).sum(1, keepdim=True)
