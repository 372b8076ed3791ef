"""k-NN at the bench shape: 16 samples x 2304 points, graph rule (fp32, k = 35, self excluded) and interpolation rule
(fp64, k = 30, moved queries against the regular grid).  usage: python profiles/knn_bench.py [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmpde_b200 import ops  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda:0")
    B, n = 16, 2304
    g = torch.linspace(0, 1, 48)
    grid = torch.stack(torch.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2).repeat(B, 1).to(dev).contiguous()
    torch.manual_seed(0)
    moved = (grid + 0.004 * torch.randn_like(grid)).contiguous()
    off = (torch.arange(B + 1, dtype=torch.int32) * n).to(dev)
    bbox = (-0.02, -0.02, 1.02, 1.02)
    cases = {"graph  fp32 k=35 (moved mesh on itself)": lambda: ops.knn_indices(moved, off, moved, off, 35, 0, True, bbox=bbox, per_sample=n),
             "interp fp64 k=30 (grid -> moved)": lambda: ops.knn_indices(grid, off, moved, off, 30, 1, False, bbox=bbox, per_sample=n),
             "interp fp64 k=30 (moved -> grid)": lambda: ops.knn_indices(moved, off, grid, off, 30, 1, False, bbox=bbox, per_sample=n)}
    for name, fn in cases.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:42s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us (grid build + search)")


if __name__ == "__main__":
    main()
