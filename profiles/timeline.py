"""Per-phase timeline of the edge kernels (CTA 0, first 48 tiles) from the clock stamps of the debug build
(`make -C mm-pde_b200/csrc timeline`).   usage: python profiles/timeline.py"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmpde_b200 import ops  # noqa: E402

TL_ITERS = 48


def main():
    lib = ctypes.CDLL(os.environ.get("MMPDE_TL_LIB", os.path.join(ROOT, "mm-pde_b200", "libmmpde_b200_tl.so")))
    P, L = ctypes.c_void_p, ctypes.c_int64
    lib.mmpde_edge_fwd.argtypes = [P, P, P, P, L, P, P, P, L, P, P]
    lib.mmpde_edge_bwd.argtypes = [P, P, P, P, L, P, P, P, L, P, P, P, P]
    lib.mmpde_debug_timeline.argtypes = [P]
    dev = torch.device("cuda:0")
    B, n, k = 16, 2304, 35
    N = B * n
    g = torch.linspace(0, 1, 48)
    pts = torch.stack(torch.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2).repeat(B, 1).to(dev).contiguous()
    off = (torch.arange(B + 1, dtype=torch.int32) * n).to(dev)
    nbr = ops.knn_indices(pts, off, pts, off, k, 0, True, bbox=(-0.02, -0.02, 1.02, 1.02), per_sample=n)
    edges = ops.EdgeList.from_knn(nbr, has_pad=False)
    E = edges.n_edges
    torch.manual_seed(0)
    PQ = torch.randn(N, 256, device=dev)
    w2, b2 = torch.randn(128, 128, device=dev) / 11, torch.randn(128, device=dev) * .1
    g_agg = torch.randn(N, 128, device=dev)
    agg = torch.zeros(N, 256, device=dev)
    mask = torch.zeros(ops.mask_words(E), dtype=torch.int32, device=dev)
    outs = [torch.zeros(N, 256, device=dev), torch.zeros(128, 128, device=dev), torch.zeros(128, device=dev)]
    buf = torch.zeros(4 * TL_ITERS * 8, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    pp = ops._ptr
    common = (pp(PQ), pp(edges.src), pp(edges.dst), pp(edges.inv_deg), E, pp(w2))

    def fwd():
        assert lib.mmpde_edge_fwd(*common, pp(b2), pp(agg, 128), 256, pp(mask), st) == 0

    def bwd():
        assert lib.mmpde_edge_bwd(*common, pp(mask), pp(g_agg), 128, pp(outs[0]), pp(outs[1]), pp(outs[2]), st) == 0

    bf = ["step start", "h_empty ok", "built+arrive", "rows 0-7 landed", "rows 0-7 built", "rows 8-15 landed"]
    names = {"fwd": (fwd, {0: bf, 1: bf, 2: ["h_full ok", "tm_empty ok", "issued"], 3: ["tm_full ok", "acc released", "done"]}),
             "bwd": (bwd, {0: ["step start", "hg_empty ok", "built+arrive", "st_full ok", "rows done"],
                           1: ["step start", "hg_empty ok", "built+arrive", "st_full ok", "rows done"],
                           2: ["hg_full ok", "d1_empty ok", "MMA-A issued", "MMA-B issued"],
                           3: ["d1_full ok", "st_empty ok", "ld done", "stage written"]})}
    if os.environ.get("MMPDE_EDGE_BWD") == "2":       # scatter in the epilogue: no staging tile, no row phase
        names["bwd"] = (bwd, {0: ["step start", "hg_empty ok", "built+arrive"], 1: ["step start", "hg_empty ok", "built+arrive"],
                              2: ["hg_full ok", "d1_empty ok", "MMA-A issued", "MMA-B issued"],
                              3: ["d1_full ok", "scatter done", "tile start", "d1 released"]})
    roles = ["builder w0", "builder w7", "MMA thread", "epilogue w0"]
    for kname, (fn, slots) in names.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        buf.zero_()
        assert lib.mmpde_debug_timeline(pp(buf)) == 0
        fn()
        torch.cuda.synchronize()
        assert lib.mmpde_debug_timeline(None) == 0
        t = buf.cpu().numpy().reshape(4, TL_ITERS, 8).astype(np.float64)
        t0 = t[t > 0].min()
        print(f"=== {kname}: clocks relative to the first stamp; steady state = tiles 8..40")
        lo, hi = 8, 40
        period = (t[2, hi, 0] - t[2, lo, 0]) / (hi - lo)
        print(f"    period per tile (MMA thread): {period:8.0f} clk")
        for r in range(4):
            s = slots[r]
            line = f"    {roles[r]:12s}"
            for j in range(len(s)):
                if s[j] == "-":
                    continue
                ref = t[r, lo:hi, 0]
                d = (t[r, lo:hi, j] - ref).mean()
                line += f"  {s[j]}: +{d:7.0f}"
            print(line)
        # cross-role offsets within a tile (relative to the builder's step start of the same tile)
        ref = t[0, lo:hi, 0]
        for r, j, label in ((0, 2, "builder w0 arrive"), (1, 2, "builder w7 arrive"), (2, 0, "MMA sees full"), (2, 2, "MMA issued(A)"),
                            (3, 0, "epi sees acc"), (3, 2, "epi released acc")):
            print(f"      {label:20s} at +{(t[r, lo:hi, j] - ref).mean():8.0f} clk after builder w0 started the tile")
        print("    first 6 tiles, MMA 'full ok' stamps:", [int(x - t0) for x in t[2, :6, 0]])
        if kname == "bwd":          # raw stamps of a few steady-state tiles, relative to builder w0's start of tile 8
            base = t[0, 8, 0]
            for i in range(8, 15):
                row = f"    tile {i:2d}:"
                for r in range(4):
                    row += f" | {roles[r]}: " + " ".join(f"{int(t[r, i, j] - base):6d}" if t[r, i, j] > 0 else "     -" for j in range(len(slots[r])))
                print(row)


if __name__ == "__main__":
    main()
