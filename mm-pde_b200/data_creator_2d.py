"""Graph / data assembly and the interpolation driver -- drop-in for /root/reference/data_creator_2d.py
(``GraphCreator_FS_2D`` :18-305) with every third-party operator replaced by sm_100a kernels:

  torch_cluster.knn_graph / radius_graph (:258,:260)  -> ops.knn_indices(rule 0) / ops.radius_indices
  sklearn NearestNeighbors per sample on the CPU (:66,:75-76, two PCIe copies + a sync per sample)
                                                     -> ops.knn_indices(rule 1), all samples in one launch
  points[indices] / labels[indices] + ItpNet + sum (:77-83) -> ops.InterpolateFn (fused, no [nu,Q,30,2] tensor)
  per-sample torch.cat growth loops (:143-152,:236-254) -> batched views

Differences a caller can observe, both deliberate:
  * the moved mesh is returned detached: its gradient only reaches the frozen mesh mover;
  * the uniform-grid graph topology is cached (it is identical every step; the reference rebuilds it).
"""
import numpy as np
import torch

from ._h2d import to_device
import torch.nn.functional as F
from torch import nn

from . import ops


class Data:
    """Graph batch with the attribute surface the reference uses from torch_geometric.data.Data:
    x, y, pos, batch, edge_index, to().  ``edge_index`` ([2,E] int64, row 0 = source, row 1 = target) is
    materialised lazily from the int32 edge list the kernels consume."""

    def __init__(self, x=None, edge_index=None, edges=None, **kw):
        self.x = x
        self.y = self.pos = self.batch = None
        self._edges = edges
        self._edge_index = edge_index
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def edge_index(self):
        if self._edge_index is None and self._edges is not None:
            self._edge_index = self._edges.edge_index()
        return self._edge_index

    @edge_index.setter
    def edge_index(self, value):
        self._edge_index, self._edges = value, None

    def to(self, device):
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


def _offsets(n_samples, per, device):
    return torch.arange(n_samples + 1, dtype=torch.int32, device=device) * per


class GraphCreator_FS_2D(nn.Module):
    def __init__(self, pde, neighbors=2, connect_edge="knn", time_window=10, t_resolution=100):
        super().__init__()
        assert isinstance(neighbors, int)
        assert isinstance(time_window, int)
        self.pde = pde
        self.n = neighbors
        self.e = connect_edge
        self.tw = time_window
        self.t_res = t_resolution
        self._static_edges = {}
        self._mm_grids = {}
        self._ref_cache = {}

    def _bbox(self):
        """Box handed to the cell-binned k-NN (a hint: points outside are clamped, results stay exact)."""
        px, py = 0.02 * self.pde.Lx, 0.02 * self.pde.Ly
        return (-px, -py, self.pde.Lx + px, self.pde.Ly + py)

    def _ori_grid(self, device):
        """pde.ori_grid on the device, copied once (the reference re-uploads it on every call; under CUDA-graph
        capture a pageable host copy is not even allowed)."""
        src = self.pde.ori_grid
        hit = self._ref_cache.get(("ori_grid", str(device)))
        if hit is None or hit[0] is not src:
            hit = self._ref_cache[("ori_grid", str(device))] = (src, src.to(device))
        return hit[1]

    def _static(self, key, make):
        """Tensors that depend only on shapes and PDE constants (coordinate axes, the regular grid, the time axis, the
        sample index of every node): built once per key instead of by a handful of tiny launches in every step."""
        hit = self._ref_cache.get(key)
        if hit is None:
            hit = self._ref_cache[key] = make()
        return hit

    # ------------------------------------------------------------------ neighbour searches of one moved-mesh step
    def _ref_points(self, key, make):
        """The reference points (regular grid / original cloud, repeated per sample) never move: the tensor and its
        cell bins are built once per (shape, batch, device)."""
        hit = self._ref_cache.get(key)
        if hit is None:
            pts = make().to(torch.float32).contiguous()
            hit = self._ref_cache[key] = [pts, None]
        return hit

    def _moved_searches(self, itp_model, mesh_pts, B, n, ref, to_mesh):
        """All neighbour searches a step needs once the mesh has moved -- graph edges on the moved mesh (fp32 rule),
        interpolation reference -> mesh (``to_mesh``) and mesh -> reference (fp64 rule) -- in ONE launch: each search
        alone is latency-bound at ~10 % of the GPU's warp slots.  Returns None when the cell-binned path does not
        apply (small samples, radius graphs): the callers then search one by one as before."""
        ref_pts, P = ref[0], ref[0].shape[0] // B
        if self.e != "knn" or min(n, P) < ops.GRID_MIN_POINTS or not mesh_pts.is_cuda:
            return None
        dev = mesh_pts.device
        mesh_pts = mesh_pts.detach().to(torch.float32).contiguous()
        off_m, off_r = _offsets(B, n, dev), _offsets(B, P, dev)
        if ref[1] is None:
            ref[1] = ops.CellBins(ref_pts, off_r, self._bbox(), P)
        bins_m = ops.CellBins(mesh_pts, off_m, self._bbox(), n)
        searches = [(bins_m, mesh_pts, off_m, self.n, 0, True), (bins_m, ref_pts, off_r, itp_model.n, 1, False)]
        if to_mesh:
            searches.append((ref[1], mesh_pts, off_m, itp_model.n, 1, False))
        outs = ops.knn_grid_multi(searches)
        return {"graph": outs[0], "back": outs[1], "to_mesh": outs[2] if to_mesh else None}

    # ------------------------------------------------------------------ interpolation (:46-85)
    def interpolate(self, itp_model, u, init_x, init_y, x, y, mode, idx=None):
        """u (nu, ...) known at (init_x, init_y) [nu*P,1]; returns values at (x, y) [nu*Q,1], flattened.
        ``idx`` (optional, not in the reference signature): the ordered neighbour lists of an earlier call with the
        same source and query points (create_graph interpolates data and labels onto the same moved mesh), saved in
        ``self.last_itp_idx`` by every call."""
        nu = u.shape[0]
        src = torch.cat((init_x, init_y), dim=-1).detach().to(torch.float32).contiguous()
        qry = torch.cat((x, y), dim=-1).detach().to(torch.float32).contiguous()
        P, Q = src.shape[0] // nu, qry.shape[0] // nu
        dev = src.device
        if idx is None:
            idx = ops.knn_indices(src, _offsets(nu, P, dev), qry, _offsets(nu, Q, dev), itp_model.n, rule=1,
                                  exclude_self=False, bbox=self._bbox(), per_sample=P)
        self.last_itp_idx = idx
        vals = u.reshape(-1).to(torch.float32).contiguous()
        return ops.InterpolateFn.apply(vals, src, qry, idx, itp_model.flat_params(mode))

    # ------------------------------------------------------------------ mesh movement (:88-137)
    @staticmethod
    def _displace(u, mesh_model, xi1, xi2):
        # x = xi + grad_xi phi(u, xi)  (:98-113).  A mover that knows its own Jacobian (mesh.dmm_model.DMM.displacement:
        # analytic forward mode, no autograd graph, recordable into the step's CUDA graph) is asked for it; any other
        # callable goes through autograd like the reference.
        analytic = getattr(mesh_model, "displacement", None)
        if analytic is not None:
            g = analytic(u.detach(), torch.cat((xi1, xi2), dim=-1).detach())
            if g is not None:
                return (g[0] + xi1).detach(), (g[1] + xi2).detach()
        with torch.enable_grad():
            xi1 = xi1.detach().requires_grad_(True)
            xi2 = xi2.detach().requires_grad_(True)
            phi = mesh_model(u.detach(), torch.cat((xi1, xi2), dim=-1))
            g1, g2 = torch.autograd.grad(phi, (xi1, xi2), grad_outputs=torch.ones_like(phi), allow_unused=True)
        return (g1 + xi1).detach(), (g2 + xi2).detach()

    def moving_mesh(self, u, mesh_model, n_grid_x, n_grid_y):
        gx = np.linspace(0, self.pde.Lx, n_grid_x)
        gy = np.linspace(0, self.pde.Ly, n_grid_y)
        # x-fastest node order, as np.meshgrid gives it (:94-96)
        key = (n_grid_x, n_grid_y, float(self.pde.Lx), float(self.pde.Ly), str(u.device))
        grid = self._mm_grids.get(key)
        if grid is None:                      # static: built and moved to the device once, not every call
            grid = torch.tensor(np.array(np.meshgrid(gx, gy)), dtype=torch.float).reshape(2, -1).t().to(u.device)
            self._mm_grids[key] = grid
        nu = u.shape[0]
        xi1 = grid[:, 0:1].repeat(nu, 1)
        xi2 = grid[:, 1:2].repeat(nu, 1)
        mm = self.pde.movingmesh_grid_size
        if mm[-2] != n_grid_x or mm[-1] != n_grid_y:
            u = F.interpolate(u.reshape(-1, 1, u.shape[-2], u.shape[-1]), size=(mm[-2], mm[-1]), mode="bilinear",
                              align_corners=True).squeeze(1)
        return self._displace(u, mesh_model, xi1, xi2)

    def moving_mesh_tri(self, u, mesh_model, grid_x, grid_y):
        return self._displace(u, mesh_model, grid_x.reshape(-1, 1), grid_y.reshape(-1, 1))

    # ------------------------------------------------------------------ data slicing (:139-154)
    def create_data(self, datapoints, steps):
        pairs = list(zip(datapoints, steps))
        if len(pairs) == 0:
            return torch.Tensor(), torch.Tensor()
        idx = torch.as_tensor([s for _, s in pairs])
        dp = datapoints[:len(pairs)]
        win = torch.arange(self.tw)
        rows = torch.arange(len(pairs))[:, None]
        return dp[rows, (idx[:, None] - self.tw + win)], dp[rows, (idx[:, None] + win)]

    # ------------------------------------------------------------------ graph assembly (:157-267)
    def _edges(self, x_new, n_samples, per_sample, static_key=None, nbr=None):
        if static_key is not None and static_key in self._static_edges:
            return self._static_edges[static_key]
        if nbr is not None:                       # searched together with the interpolation lists (_moved_searches)
            return ops.EdgeList.from_knn(nbr, has_pad=per_sample - 1 < self.n)
        dev = x_new.device
        off = _offsets(n_samples, per_sample, dev)
        pts = x_new.detach().to(torch.float32).contiguous()
        if self.e == "radius":
            nbr = ops.radius_indices(pts, off, self._radius, 32)
            edges = ops.EdgeList.from_knn(nbr, has_pad=True)
        else:
            nbr = ops.knn_indices(pts, off, pts, off, self.n, rule=0, exclude_self=True, bbox=self._bbox(),
                                  per_sample=per_sample)
            edges = ops.EdgeList.from_knn(nbr, has_pad=per_sample - 1 < self.n)
        if static_key is not None:
            self._static_edges[static_key] = edges
        return edges

    def create_graph(self, itp_model, data, labels, steps, device, mesh_model=None):
        data, labels = to_device(data, device), to_device(labels, device)
        pde = self.pde
        B = data.shape[0]
        pre = None                                   # neighbour lists searched together once the mesh has moved
        if len(pde.grid_size) == 3:
            onx, ony = data.shape[-2], data.shape[-1]
            nt, nx, ny = pde.grid_size
            n = nx * ny
            if self.e == "radius":
                xs = torch.linspace(0, pde.Lx, nx, device=device)
                ys = torch.linspace(0, pde.Ly, ny, device=device)
                self._radius = float(self.n * torch.sqrt((xs[1] - xs[0]) ** 2 + (ys[1] - ys[0]) ** 2) + 0.0001)
            else:
                self._radius = None
            grid = self._static(("axes_grid", nx, ny, float(pde.Lx), float(pde.Ly), str(device)), lambda: torch.stack(
                torch.meshgrid(torch.linspace(0, pde.Lx, nx, device=device), torch.linspace(0, pde.Ly, ny, device=device),
                               indexing="ij"), dim=2).float().reshape(1, n, 2)).expand(B, n, 2)
            static_key = ("grid", B, nx, ny, str(device))
            if mesh_model is not None:
                mm_nx, mm_ny = pde.movingmesh_grid_size[-2], pde.movingmesh_grid_size[-1]
                coarse = data.reshape(-1, onx, ony)[:, ::int(onx / mm_nx), ::int(ony / mm_ny)]
                mesh_x, mesh_y = self.moving_mesh(coarse, mesh_model, nx, ny)
                mesh = torch.cat((mesh_x, mesh_y), dim=-1).reshape(-1, n, 2)
                ref = self._ref_points(("grid", B, onx, ony, str(device)), lambda: torch.stack(torch.meshgrid(
                    torch.linspace(0, pde.Lx, onx, device=device), torch.linspace(0, pde.Ly, ony, device=device),
                    indexing="ij"), dim=2).reshape(1, -1, 2).expand(B, -1, 2).reshape(-1, 2))
                og = ref[0]
                if self.tw == 1 and len(steps) >= B:
                    pre = self._moved_searches(itp_model, mesh.reshape(-1, 2), B, n, ref, to_mesh=True)
                data = self.interpolate(itp_model, data.reshape(-1, onx, ony), og[:, 0:1], og[:, 1:2], mesh_x, mesh_y, mode="1",
                                        idx=pre["to_mesh"] if pre else None).reshape(-1, self.tw, nx, ny)
                labels = self.interpolate(itp_model, labels.reshape(-1, onx, ony), og[:, 0:1], og[:, 1:2],
                                          mesh_x, mesh_y, mode="1", idx=self.last_itp_idx).reshape(-1, self.tw, nx, ny)
                static_key = None
            else:
                mesh = grid
        else:
            n = pde.ori_grid_size[1]
            nt = pde.grid_size[0]
            grid = self._ori_grid(device)[None].expand(B, n, 2)
            if self.e == "radius":
                side = int(np.sqrt(pde.grid_size[1]))
                hx = pde.Lx / (side - 1)
                self._radius = float(self.n * np.sqrt(2 * hx * hx) + 0.0001)
            static_key = ("cloud", B, n, str(device))
            if mesh_model is not None:
                mesh_x, mesh_y = self.moving_mesh_tri(data.reshape(-1, n), mesh_model,
                                                      grid[:, :, 0].contiguous(), grid[:, :, 1].contiguous())
                mesh = torch.cat((mesh_x, mesh_y), dim=-1).reshape(-1, n, 2)
                if len(steps) >= B:
                    ref = self._ref_points(("cloud", B, n, str(device)), lambda: grid.reshape(-1, 2))
                    pre = self._moved_searches(itp_model, mesh.reshape(-1, 2), B, n, ref, to_mesh=False)
                static_key = None
            else:
                mesh = grid
        t = self._static(("t_axis", float(pde.tmin), float(pde.tmax), nt, str(device)),
                         lambda: torch.linspace(pde.tmin, pde.tmax, nt, device=device))
        B = min(B, len(steps))
        u_new = data[:B].reshape(B, self.tw, n).permute(0, 2, 1).reshape(B * n, self.tw)
        y_new = labels[:B].reshape(B, self.tw, n).permute(0, 2, 1).reshape(B * n, self.tw)
        x_new = mesh[:B].reshape(B * n, 2)
        # ``steps`` may already be a device tensor (train_helper_2d.StepGraph: the step indices are a graph input)
        step_idx = steps[:B] if torch.is_tensor(steps) else to_device(torch.as_tensor(list(steps[:B])), device)
        t_new = t[step_idx].repeat_interleave(n)
        batch = self._static(("batch_index", B, n, str(device)), lambda: torch.arange(B, device=device).repeat_interleave(n))
        if static_key is not None:
            static_key = static_key + (B,)
        graph = Data(x=u_new, edges=self._edges(x_new, B, n, static_key, nbr=pre["graph"] if pre else None))
        graph._itp_back_idx = pre["back"] if pre else None      # mesh -> reference lists for interpolate_pred
        graph.y = y_new
        graph.pos = torch.cat((t_new[:, None], x_new), dim=1)
        graph.batch = batch
        return graph

    # ------------------------------------------------------------------ prediction back on the grid (:270-305)
    def interpolate_pred(self, itp_model, pred, graph, data, device):
        data = to_device(data, device)
        pde = self.pde
        if len(pde.grid_size) == 3:
            onx, ony = pde.ori_grid_size[1], pde.ori_grid_size[2]
            nx, ny = pde.grid_size[1], pde.grid_size[2]
            nu = pred.shape[0] // (nx * ny)
            og = self._static(("ori_axes_grid", nu, onx, ony, float(pde.Lx), float(pde.Ly), str(device)), lambda: torch.stack(
                torch.meshgrid(torch.linspace(0, pde.Lx, onx, device=device), torch.linspace(0, pde.Ly, ony, device=device),
                               indexing="ij"), dim=2).reshape(1, -1, 2).expand(nu, -1, 2).reshape(-1, 2))
            on_grid = self.interpolate(itp_model, pred.reshape(-1, nx, ny), graph.pos[:, 1:2], graph.pos[:, 2:3],
                                       og[:, 0:1], og[:, 1:2], mode="2",
                                       idx=getattr(graph, "_itp_back_idx", None)).reshape(-1, 1, onx, ony)
            out = itp_model(None, None, mode="res_cut", data=data).reshape(-1, 1, onx, ony) + on_grid
        else:
            n = pde.ori_grid_size[1]
            nu = pred.shape[0] // n
            g = self._ori_grid(device)[None].expand(nu, n, 2).reshape(-1, 2)
            on_grid = self.interpolate(itp_model, pred.reshape(-1, n), graph.pos[:, 1:2], graph.pos[:, 2:3],
                                       g[:, 0:1], g[:, 1:2], mode="2", idx=getattr(graph, "_itp_back_idx", None)).reshape(-1, n)
            out = itp_model(None, None, mode="res_cut", data=data.reshape(-1, n)).reshape(-1, n) + on_grid
        return out.reshape(-1, 1)
