"""Times the hot kernels in isolation at config-1 size (N=36 864, E=1 290 240) with CUDA events.
usage: python profiles/edge_bench.py [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmpde_b200 import _cabi, ops  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
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
    st = ops._stream()
    agg = torch.zeros(N, 256, device=dev)
    mask = torch.zeros(ops.mask_words(E), dtype=torch.int32, device=dev)
    common = (ops._ptr(PQ), ops._ptr(edges.src), ops._ptr(edges.dst), ops._ptr(edges.inv_deg), E, ops._ptr(w2))
    fwd = lambda: _cabi.call("mmpde_edge_fwd", *common, ops._ptr(b2), ops._ptr(agg, 128), 256, ops._ptr(mask), st)
    outs = [torch.zeros(N, 256, device=dev), torch.zeros(128, 128, device=dev), torch.zeros(128, device=dev)]
    bwd = lambda: _cabi.call("mmpde_edge_bwd", *common, ops._ptr(mask), ops._ptr(g_agg), 128, ops._ptr(outs[0]),
                             ops._ptr(outs[1]), ops._ptr(outs[2]), st)
    X = torch.randn(N, 256, device=dev)
    W = torch.randn(128, 260, device=dev)
    C = torch.empty(N, 128, device=dev)
    dW = torch.zeros(128, 260, device=dev)
    gemm_nt = lambda: ops.gemm(ops._ptr(X), 256, 1, ops._ptr(W), 260, 1, ops._ptr(C), 128, N, 128, 128, st=st)
    gemm_tn = lambda: ops.gemm(ops._ptr(C), 128, 0, ops._ptr(X), 256, 0, ops._ptr(dW), 260, 128, 128, N, split_k=ops._split_for(N), st=st)
    knn = lambda: ops.knn_indices(pts, off, pts, off, k, 0, True, bbox=(-0.02, -0.02, 1.02, 1.02), per_sample=n)
    b1 = torch.randn(128, device=dev)
    n4 = torch.randn(N, 4, device=dev)
    wx = torch.randn(128, 4, device=dev)
    ng_lin = lambda: ops.node_gemm(ops._ptr(X), 256, ops._ptr(W), 260, 1, ops._ptr(C), 128, N, ext=(ops._ptr(n4), ops._ptr(wx)),
                                   bias=ops._ptr(b1), st=st)
    ng_k256 = lambda: ops.node_gemm(ops._ptr(X), 256, ops._ptr(W), 260, 1, ops._ptr(C), 128, N, A1=ops._ptr(X, 128), lda1=256,
                                    W1=ops._ptr(W, 128), w1_ns=260, w1_ks=1, bias=ops._ptr(b1), relu=1, st=st)
    ng_dgrad = lambda: ops.node_gemm(ops._ptr(X), 256, ops._ptr(W), 1, 260, ops._ptr(C), 128, N, R1=ops._ptr(C), ldr1=128, st=st)
    dWt = torch.zeros(128, 260, device=dev)
    dWx, dbt = torch.zeros(128, 4, device=dev), torch.zeros(128, device=dev)
    nw = lambda: ops.node_wgrad(ops._ptr(C), 128, N, B=ops._ptr(X), ldb=256, dW=ops._ptr(dWt), ldw=260, Bext=ops._ptr(n4),
                                dWext=ops._ptr(dWx), dbias=ops._ptr(dbt), st=st)
    print(f"node_gemm K=128+ext {timeit(ng_lin, iters):7.1f} us   K=256 relu {timeit(ng_k256, iters):7.1f} us   "
          f"dgrad+residual {timeit(ng_dgrad, iters):7.1f} us   node_wgrad {timeit(nw, iters):7.1f} us")
    t_f, t_b = timeit(fwd, iters), timeit(bwd, iters)
    flop = 99328 * E
    print(f"edge_fwd {t_f:8.1f} us  {flop / t_f / 1e6:7.1f} TFLOP/s algorithmic  ({3 * 32768 * E / t_f / 1e6:6.1f} executed bf16)")
    print(f"edge_bwd {t_b:8.1f} us  {2 * flop / t_b / 1e6:7.1f} TFLOP/s algorithmic  ({6 * 32768 * E / t_b / 1e6:6.1f} executed bf16)")
    print(f"gemm NT [N,128]x[128,128] {timeit(gemm_nt, iters):8.1f} us   gemm TN wgrad {timeit(gemm_tn, iters):8.1f} us   knn_graph {timeit(knn, iters):8.1f} us")


if __name__ == "__main__":
    main()
