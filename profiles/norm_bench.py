"""Stand-alone timing of the BatchNorm / ReLU-mask kernels at the bench shape ([36864,128] fp32), rotating over enough
buffer sets to miss L2 (cold) or reusing one (warm).  usage: python profiles/norm_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmpde_b200 import _cabi  # noqa: E402
from mmpde_b200.ops import _ptr, _stream  # noqa: E402

M, H = 36864, 128
dev = torch.device("cuda:0")


def bench(name, fn, sets, reps=200):
    for i in range(10):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    def mk():
        return dict(A=torch.randn(M, H, device=dev), B=torch.randn(M, H, device=dev), g=torch.randn(M, H, device=dev),
                    out=torch.randn(M, H, device=dev), gy=torch.empty(M, H, device=dev))
    cold = [mk() for _ in range(8)]
    warm = cold[:1]
    sums = torch.zeros(16, 2, 128, dtype=torch.float64, device=dev)     # MMPDE_BN_REPLICAS copies
    mr = torch.cat((torch.zeros(128, device=dev), torch.ones(128, device=dev))).contiguous()
    gamma, beta = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    colsum = torch.zeros(128, device=dev)
    st = _stream()
    kernels = {
        "bn_stats(A+B)": lambda t: _cabi.call("mmpde_bn_stats", _ptr(t["A"]), H, _ptr(t["B"]), H, M, _ptr(sums), st),
        "bn_stats(A)": lambda t: _cabi.call("mmpde_bn_stats", _ptr(t["A"]), H, None, 0, M, _ptr(sums), st),
        "bn_apply(A+B)": lambda t: _cabi.call("mmpde_bn_apply", _ptr(t["A"]), H, _ptr(t["B"]), H, M, _ptr(mr), _ptr(gamma),
                                              _ptr(beta), 0, _ptr(t["out"]), H, st),
        "bn_bwd_reduce(A+B)": lambda t: _cabi.call("mmpde_bn_bwd_reduce", _ptr(t["g"]), H, None, 0, 0, _ptr(t["A"]), H,
                                                   _ptr(t["B"]), H, M, _ptr(mr), _ptr(sums), st),
        "bn_bwd_apply(A+B)": lambda t: _cabi.call("mmpde_bn_bwd_apply", _ptr(t["g"]), H, None, 0, 0, _ptr(t["A"]), H,
                                                  _ptr(t["B"]), H, M, _ptr(mr), _ptr(gamma), _ptr(sums), float(M),
                                                  _ptr(t["gy"]), H, 0, None, 0, st),
        "bn_bwd_apply(A+B)+gated": lambda t: _cabi.call("mmpde_bn_bwd_apply", _ptr(t["g"]), H, None, 0, 0, _ptr(t["A"]), H,
                                                        _ptr(t["B"]), H, M, _ptr(mr), _ptr(gamma), _ptr(sums), float(M),
                                                        _ptr(t["gy"]), H, 0, _ptr(t["out"]), H, st),
        "relu_bwd(+colsum)": lambda t: _cabi.call("mmpde_relu_bwd", _ptr(t["g"]), H, _ptr(t["A"]), H, M, _ptr(t["gy"]), H,
                                                  _ptr(colsum), st),
    }
    traffic = {"bn_stats(A+B)": 2, "bn_stats(A)": 1, "bn_apply(A+B)": 3, "bn_bwd_reduce(A+B)": 3, "bn_bwd_apply(A+B)": 4, "bn_bwd_apply(A+B)+gated": 5,
               "relu_bwd(+colsum)": 3}
    print(f"MMPDE_REDUCE_CTAS_PER_SM={os.environ.get('MMPDE_REDUCE_CTAS_PER_SM', '(default)')}")
    for name, fn in kernels.items():
        c, w = bench(name, fn, cold), bench(name, fn, warm)
        gb = traffic[name] * M * H * 4 / 1e9
        print(f"  {name:22s} cold {c:6.1f} us ({gb / c * 1e6:6.0f} GB/s)   warm {w:6.1f} us ({gb / w * 1e6:6.0f} GB/s)")


if __name__ == "__main__":
    main()
