"""Phase stamps of one mmpde_node_gemm launch (CTA 0) from the debug build (make -C mm-pde_b200/csrc timeline)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmpde_b200 import ops, _cabi
lib = ctypes.CDLL(os.path.join(ROOT, "mm-pde_b200", "libmmpde_b200_tl.so"))
P, L, I = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
lib.mmpde_node_gemm.argtypes = _cabi.SIGNATURES["mmpde_node_gemm"]
lib.mmpde_debug_timeline_node.argtypes = [P]
dev = torch.device("cuda:0")
N = 36864
X, W, C = torch.randn(N, 256, device=dev), torch.randn(128, 260, device=dev), torch.empty(N, 128, device=dev)
b = torch.randn(128, device=dev)
buf = torch.zeros(4 * 48 * 8, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
pp = ops._ptr
def run():
    assert lib.mmpde_node_gemm(pp(X), 256, None, 0, pp(W), 260, 1, None, 0, 0, None, None, pp(b), 0, None, 0, None, 0, pp(C), 128, N, st) == 0
for _ in range(3): run()
torch.cuda.synchronize(); buf.zero_(); lib.mmpde_debug_timeline_node(pp(buf)); run(); torch.cuda.synchronize(); lib.mmpde_debug_timeline_node(None)
t = buf.cpu().numpy().reshape(4, 48, 8)
t0 = t[3, 0, 6]
rel = lambda x: int(x - t0) if x > 0 else None
print("kernel start -> (clks)  W in TMEM:", rel(t[3,0,0]), " end:", rel(t[3,0,7]), " epilogue before/after setmaxnreg:", rel(t[3,0,4]), rel(t[3,0,5]),
      " builder before/after setmaxnreg:", rel(t[0,0,4]), rel(t[0,0,5]))
for i in range(2):
    print(f" tile {i}: builder start {rel(t[0,i,0])} empty-ok {rel(t[0,i,1])} built {rel(t[0,i,2])} | mma tm_empty-ok {rel(t[2,i,0])} issued {rel(t[2,i,2])} | epi acc-ready {rel(t[3,i,1])} stored {rel(t[3,i,2])}")

# ---- the same contraction with the weight block as a pre-split image (TMA bulk copy + tcgen05.cp instead of the register split)
lib.mmpde_node_gemm_img.argtypes = _cabi.SIGNATURES["mmpde_node_gemm_img"]
imgs, keep = ops.weight_images([(pp(W), 260, 1)], dev)
def run_i():
    assert lib.mmpde_node_gemm_img(pp(X), 256, None, 0, imgs[0], None, None, None, pp(b), 0, None, 0, None, 0, pp(C), 128, N, st) == 0
for _ in range(3): run_i()
torch.cuda.synchronize(); buf.zero_(); lib.mmpde_debug_timeline_node(pp(buf)); run_i(); torch.cuda.synchronize(); lib.mmpde_debug_timeline_node(None)
t = buf.cpu().numpy().reshape(4, 48, 8)
t0 = t[3, 0, 6]
print("weight image: copies issued (MMA thread):", rel(t[2,0,4]), " end:", rel(t[3,0,7]))
for i in range(2):
    print(f" tile {i}: builder start {rel(t[0,i,0])} empty-ok {rel(t[0,i,1])} built {rel(t[0,i,2])} | mma tm_empty-ok {rel(t[2,i,0])} issued {rel(t[2,i,2])} | epi acc-ready {rel(t[3,i,1])} stored {rel(t[3,i,2])}")

# ---- stand-alone launch time (release build, CUDA events over 200 back-to-back launches; operands L2-resident)
def timed(fn, n=200):
    for _ in range(10): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
for M in (36864, 18432, 4608):
    t_reg = timed(lambda: ops.node_gemm(pp(X), 256, pp(W), 260, 1, pp(C), 128, M, bias=pp(b)))
    t_img = timed(lambda: ops.node_gemm(pp(X), 256, None, 0, 0, pp(C), 128, M, bias=pp(b), img0=imgs[0]))
    print(f"back-to-back launches, M = {M}: register split {t_reg:.2f} us, weight image {t_img:.2f} us")

# ---- dgrad + residual in place: W^T by strides, R1 = C
def run_d():
    assert lib.mmpde_node_gemm(pp(X), 256, None, 0, pp(W), 1, 260, None, 0, 0, None, None, None, 0, pp(C), 128, None, 0, pp(C), 128, N, st) == 0
for _ in range(3): run_d()
torch.cuda.synchronize(); buf.zero_(); lib.mmpde_debug_timeline_node(pp(buf)); run_d(); torch.cuda.synchronize(); lib.mmpde_debug_timeline_node(None)
t = buf.cpu().numpy().reshape(4, 48, 8)
t0 = t[3, 0, 6]
print("dgrad+residual: W in TMEM:", rel(t[3,0,0]), " end:", rel(t[3,0,7]))
for i in range(2):
    print(f" tile {i}: builder start {rel(t[0,i,0])} empty-ok {rel(t[0,i,1])} built {rel(t[0,i,2])} | mma tm_empty-ok {rel(t[2,i,0])} issued {rel(t[2,i,2])} | epi acc-ready {rel(t[3,i,1])} stored {rel(t[3,i,2])}")

# ---- wgrad
lib.mmpde_node_wgrad.argtypes = _cabi.SIGNATURES["mmpde_node_wgrad"]
n4 = torch.randn(N, 4, device=dev)
dW, dWx, db = torch.zeros(128, 260, device=dev), torch.zeros(128, 4, device=dev), torch.zeros(128, device=dev)
def runw():
    assert lib.mmpde_node_wgrad(pp(C), 128, pp(X), 256, pp(n4), pp(dW), 260, pp(dWx), 4, pp(db), N, st) == 0
for _ in range(3): runw()
torch.cuda.synchronize(); buf.zero_(); lib.mmpde_debug_timeline_node(pp(buf)); runw(); torch.cuda.synchronize(); lib.mmpde_debug_timeline_node(None)
t = buf.cpu().numpy().reshape(4, 48, 8)
t0 = t[3, 0, 6]
print("wgrad: kernel start -> all MMAs done:", rel(t[3,0,1]), " reductions done:", rel(t[3,0,2]), " end:", rel(t[3,0,7]))
for i in range(4):
    print(f" tile {i}: builder start {rel(t[0,i,0])} empty-ok {rel(t[0,i,1])} built {rel(t[0,i,2])} | mma full-ok {rel(t[2,i,0])} issued {rel(t[2,i,2])}")
