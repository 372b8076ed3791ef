"""BASELINE.json config 5: interpolation-only sweep, 10k .. 4M nodes, moved <-> reference mesh.
For every size: ordered 30-NN search (fp64 rule = sklearn kd-tree arithmetic) + fused interpolation forward (+ backward)
on the GPU, timed with CUDA events; reported as queries/s, achieved HBM GB/s against the algorithmic 144 B/query
 and fp32 TFLOP/s against 36 096 FLOP/query, with the reference's own CPU path (sklearn
NearestNeighbors + torch MLP, data_creator_2d.py:66-83) timed beside it on a bounded sample.
usage: python profiles/c5_interp_sweep.py [--sizes 10000,40000,...] [--cpu-max 160000]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmpde_b200 import ops  # noqa: E402
from mmpde_b200.interpolate import ItpNet  # noqa: E402

BYTES_PER_QUERY = 30 * 4 + 8 + 4 + 12          # int32 idx, query xy, out, amortised source xy + value (P = Q)
FLOP_PER_QUERY = 36096


def mesh(n, seed):
    side = int(round(n ** 0.5))
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.linspace(0, 1, side), np.linspace(0, 1, side), indexing="ij"), -1).reshape(-1, 2)
    moved = (g + rng.uniform(-0.3, 0.3, g.shape) / (side - 1)).astype(np.float32)
    return moved, g.astype(np.float32)


def timeit(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="10000,40000,160000,640000,1000000,2000000,4000000")
    ap.add_argument("--cpu-max", type=int, default=160000)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    torch.manual_seed(0)
    net = ItpNet(48, 48, [128, 64], [128, 64], [1, 4, 16, 4, 1]).to(dev)
    out = []
    for n in [int(s) for s in a.sizes.split(",")]:
        moved_np, ref_np = mesh(n, n)
        n = moved_np.shape[0]
        moved, ref = torch.from_numpy(moved_np).to(dev), torch.from_numpy(ref_np).to(dev)
        vals = torch.randn(n, device=dev, requires_grad=True)
        row = {"nodes": n}
        for mode, (src, qry) in (("2 moved->reference", (moved, ref)), ("1 reference->moved", (ref, moved))):
            iters = 5 if n <= 1000000 else 2
            idx = ops.knn_indices_grid(src, qry, 30, 1, False)
            t_knn = timeit(lambda: ops.knn_indices_grid(src, qry, 30, 1, False), iters)
            flat = net.flat_params(mode[0])
            t_fwd = timeit(lambda: ops.InterpolateFn.apply(vals, src, qry, idx, flat), iters)

            def fb():
                vals.grad = None
                o = ops.InterpolateFn.apply(vals, src, qry, idx, flat)
                o.sum().backward()
            t_fb = timeit(fb, iters)
            row[mode] = {"knn_ms": t_knn, "knn_Mqueries_per_s": n / t_knn / 1e3,
                         "itp_fwd_ms": t_fwd, "itp_fwd_GBs": n * BYTES_PER_QUERY / t_fwd / 1e6,
                         "itp_fwd_frac_of_hbm_peak": n * BYTES_PER_QUERY / t_fwd / 1e6 / peaks["hbm_gbs"],
                         "itp_fwd_fp32_TFLOPs": n * FLOP_PER_QUERY / t_fwd / 1e9,
                         "itp_fwd_bwd_ms": t_fb, "interp_Mqueries_per_s_incl_knn": n / (t_knn + t_fwd) / 1e3}
        if n <= a.cpu_max:          # the reference's CPU path: sklearn kd-tree + gather + MLP + weighted sum
            from sklearn.neighbors import NearestNeighbors
            cpu_net = ItpNet(48, 48, [128, 64], [128, 64], [1, 4, 16, 4, 1])
            cpu_net.load_state_dict(net.state_dict())
            pts, q, lab = torch.from_numpy(moved_np), torch.from_numpy(ref_np), vals.detach().cpu()
            torch.set_num_threads(os.cpu_count() or 1)
            t0 = time.perf_counter()
            ind = NearestNeighbors(n_neighbors=30).fit(moved_np).kneighbors(ref_np, return_distance=False)
            t1 = time.perf_counter()
            with torch.no_grad():
                nb = pts[ind]
                w = cpu_net(nb[None], q[None, :, None, :], "2")
                res = (w * lab[ind][None]).sum(-1)
            t2 = time.perf_counter()
            row["cpu_reference_path"] = {"kdtree_ms": (t1 - t0) * 1e3, "mlp_ms": (t2 - t1) * 1e3, "cores": os.cpu_count(),
                                         "Mqueries_per_s": n / (t2 - t0) / 1e6}
            gpu = ops.InterpolateFn.apply(vals.detach(), moved, ref, ops.knn_indices_grid(moved, ref, 30, 1, False), net.flat_params("2"))
            row["gpu_vs_cpu_rel_l2"] = float((gpu.cpu() - res.reshape(-1)).norm() / res.norm())
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
