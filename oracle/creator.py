"""ORACLE (test infrastructure).  GraphCreator_FS_2D restatement,
/root/reference/data_creator_2d.py:18-305, PyG-free.

``knn_backend='sklearn'`` runs the exact library the reference calls (:66,75-76);
``knn_backend='rule'`` runs the frozen (fp64 d2, index) rule of knn_oracle.c, which is what the
CUDA path must reproduce bit-exactly.  tests/test_oracle_knn.py checks the two agree away from ties.
"""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from . import knn as _knn


class Data:
    """Minimal stand-in for torch_geometric.data.Data (x, y, pos, batch, edge_index, .to())."""

    def __init__(self, x=None, edge_index=None, **kw):
        self.x, self.edge_index = x, edge_index
        self.y = self.pos = self.batch = None
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in vars(self).items():
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class GraphCreator_FS_2D(nn.Module):
    def __init__(self, pde, neighbors=2, connect_edge="knn", time_window=10, t_resolution=100,
                 knn_backend="rule"):
        super().__init__()
        assert isinstance(neighbors, int) and isinstance(time_window, int)
        self.pde, self.n, self.e, self.tw, self.t_res = pde, neighbors, connect_edge, time_window, t_resolution
        self.knn_backend = knn_backend

    # ---- data_creator_2d.py:46-85 -------------------------------------------------------------
    def _itp_indices(self, points, queries, k):
        if self.knn_backend == "sklearn":
            from sklearn.neighbors import NearestNeighbors
            nn_ = NearestNeighbors(n_neighbors=k)
            nn_.fit(points.detach().cpu().numpy())
            return torch.from_numpy(nn_.kneighbors(queries.detach().cpu().numpy())[1])
        idx, _ = _knn.knn_indices(points.detach().cpu().numpy(), queries.detach().cpu().numpy(), k, rule="f64")
        return torch.from_numpy(idx)

    def interpolate(self, itp_model, u, init_x, init_y, x, y, mode):
        nu = u.shape[0]
        pts = torch.cat((init_x, init_y), -1).reshape(nu, -1, 2)
        qry = torch.cat((x, y), -1).reshape(nu, -1, 2)
        nb_xy, nb_val = [], []
        for s in range(nu):
            idx = self._itp_indices(pts[s], qry[s], 30).to(u.device)
            nb_xy.append(pts[s][idx])
            nb_val.append(u[s].reshape(-1)[idx])
        nb_xy, nb_val = torch.stack(nb_xy), torch.stack(nb_val)
        w = itp_model(nb_xy, qry.unsqueeze(-2), mode)
        return torch.sum(w * nb_val, dim=-1).reshape(-1)

    # ---- data_creator_2d.py:88-137 ------------------------------------------------------------
    def _move(self, u, mesh_model, xi1, xi2):
        xi1.requires_grad = True
        xi2.requires_grad = True
        phi = mesh_model(u, torch.cat((xi1, xi2), dim=-1))
        ones = torch.ones_like(phi)
        g1 = torch.autograd.grad(phi, xi1, grad_outputs=ones, retain_graph=True, create_graph=True, allow_unused=True)[0]
        g2 = torch.autograd.grad(phi, xi2, grad_outputs=ones, retain_graph=True, create_graph=True, allow_unused=True)[0]
        return g1 + xi1, g2 + xi2

    def moving_mesh(self, u, mesh_model, n_grid_x, n_grid_y):
        gx = np.linspace(0, self.pde.Lx, n_grid_x)
        gy = np.linspace(0, self.pde.Ly, n_grid_y)
        grid = torch.tensor(np.array(np.meshgrid(gx, gy)), dtype=torch.float).reshape(2, -1).permute(1, 0).to(u.device)
        B = u.shape[0]
        xi1 = grid[:, [0]].unsqueeze(0).repeat(B, 1, 1).reshape(-1, 1)
        xi2 = grid[:, [1]].unsqueeze(0).repeat(B, 1, 1).reshape(-1, 1)
        mm = self.pde.movingmesh_grid_size
        if mm[-2] != n_grid_x or mm[-1] != n_grid_y:
            u = F.interpolate(u.reshape(-1, 1, u.shape[-2], u.shape[-1]), size=(mm[-2], mm[-1]),
                              mode="bilinear", align_corners=True).squeeze(1)
        return self._move(u, mesh_model, xi1, xi2)

    def moving_mesh_tri(self, u, mesh_model, grid_x, grid_y):
        return self._move(u, mesh_model, grid_x.reshape(-1, 1), grid_y.reshape(-1, 1))

    # ---- data_creator_2d.py:139-154 -----------------------------------------------------------
    def create_data(self, datapoints, steps):
        pairs = list(zip(datapoints, steps))
        if not pairs:
            return torch.Tensor(), torch.Tensor()
        data = torch.stack([dp[s - self.tw:s] for dp, s in pairs])
        labels = torch.stack([dp[s:s + self.tw] for dp, s in pairs])
        return data, labels

    # ---- data_creator_2d.py:157-267 -----------------------------------------------------------
    def create_graph(self, itp_model, data, labels, steps, device, mesh_model=None):
        data, labels = data.to(device), labels.to(device)
        pde = self.pde
        B = data.shape[0]
        if len(pde.grid_size) == 3:
            onx, ony = data.shape[-2], data.shape[-1]
            ogx, ogy = torch.meshgrid(torch.linspace(0, pde.Lx, onx).to(device),
                                      torch.linspace(0, pde.Ly, ony).to(device), indexing="ij")
            mm_nx, mm_ny = pde.movingmesh_grid_size[-2], pde.movingmesh_grid_size[-1]
            nt, nx, ny = pde.grid_size
            n = nx * ny
            xs = torch.linspace(0, pde.Lx, nx).to(device)
            ys = torch.linspace(0, pde.Ly, ny).to(device)
            radius = self.n * torch.sqrt((xs[1] - xs[0]) ** 2 + (ys[1] - ys[0]) ** 2) + 0.0001
            gx, gy = torch.meshgrid(xs, ys, indexing="ij")
            grid = torch.stack((gx, gy), 2).float().view(-1, 2)[None].repeat(B, 1, 1)
            if mesh_model is not None:
                coarse = data.reshape(-1, onx, ony)[:, ::int(onx / mm_nx), ::int(ony / mm_ny)]
                mesh_x, mesh_y = self.moving_mesh(coarse, mesh_model, nx, ny)
                mesh = torch.cat((mesh_x, mesh_y), dim=-1).reshape(-1, n, 2)
                src_x = ogx[None].repeat(B, 1, 1).reshape(-1, 1)
                src_y = ogy[None].repeat(B, 1, 1).reshape(-1, 1)
                data = self.interpolate(itp_model, data.reshape(-1, onx, ony), src_x, src_y,
                                        mesh_x, mesh_y, mode="1").reshape(-1, self.tw, nx, ny)
                labels = self.interpolate(itp_model, labels.reshape(-1, onx, ony), src_x, src_y,
                                          mesh_x, mesh_y, mode="1").reshape(-1, self.tw, nx, ny)
            else:
                mesh = grid
        else:
            n = pde.ori_grid_size[1]
            grid = pde.ori_grid[None].repeat(B, 1, 1).to(device)
            nt = pde.grid_size[0]
            side = int(np.sqrt(pde.grid_size[1]))
            xs = torch.linspace(0, pde.Lx, side).to(device)
            radius = self.n * torch.sqrt(2 * (xs[1] - xs[0]) ** 2) + 0.0001
            if mesh_model is not None:
                mesh_x, mesh_y = self.moving_mesh_tri(data.reshape(-1, n), mesh_model, grid[:, :, 0], grid[:, :, 1])
                mesh = torch.cat((mesh_x, mesh_y), dim=-1).reshape(-1, n, 2)
            else:
                mesh = grid
        t = torch.linspace(pde.tmin, pde.tmax, nt).to(device)
        B = min(B, len(steps))
        u_new = data[:B].reshape(B, self.tw, n).permute(0, 2, 1).reshape(B * n, self.tw)
        y_new = labels[:B].reshape(B, self.tw, n).permute(0, 2, 1).reshape(B * n, self.tw)
        x_new = mesh[:B].reshape(B * n, 2)
        t_new = t[torch.as_tensor(list(steps[:B]), device=device)].repeat_interleave(n)
        batch = torch.arange(B, device=device).repeat_interleave(n)
        if self.e == "radius":
            edge_index = _knn.radius_graph(x_new, float(radius), batch)
        else:
            edge_index = _knn.knn_graph(x_new, self.n, batch)
        g = Data(x=u_new, edge_index=edge_index.to(device))
        g.y, g.pos, g.batch = y_new, torch.cat((t_new[:, None], x_new), 1), batch
        return g.to(device)

    # ---- data_creator_2d.py:270-305 -----------------------------------------------------------
    def interpolate_pred(self, itp_model, pred, graph, data, device):
        data = data.to(device)
        pde = self.pde
        if len(pde.grid_size) == 3:
            onx, ony = pde.ori_grid_size[1], pde.ori_grid_size[2]
            ogx, ogy = torch.meshgrid(torch.linspace(0, pde.Lx, onx).to(device),
                                      torch.linspace(0, pde.Ly, ony).to(device), indexing="ij")
            nx, ny = pde.grid_size[1], pde.grid_size[2]
            nu = pred.shape[0] // (nx * ny)
            on_grid = self.interpolate(itp_model, pred.reshape(-1, nx, ny), graph.pos[:, [1]], graph.pos[:, [2]],
                                       ogx[None].repeat(nu, 1, 1).reshape(-1, 1),
                                       ogy[None].repeat(nu, 1, 1).reshape(-1, 1), mode="2").reshape(-1, 1, onx, ony)
            out = itp_model(None, None, mode="res_cut", data=data).reshape(-1, 1, onx, ony) + on_grid
        else:
            n = pde.ori_grid_size[1]
            nu = pred.shape[0] // n
            gx, gy = pde.ori_grid[:, 0].to(device), pde.ori_grid[:, 1].to(device)
            on_grid = self.interpolate(itp_model, pred.reshape(-1, n), graph.pos[:, [1]], graph.pos[:, [2]],
                                       gx[None].repeat(nu, 1).reshape(-1, 1), gy[None].repeat(nu, 1).reshape(-1, 1),
                                       mode="2").reshape(-1, n)
            out = itp_model(None, None, mode="res_cut", data=data.reshape(-1, n)).reshape(-1, n) + on_grid
        return out.reshape(-1, 1)
