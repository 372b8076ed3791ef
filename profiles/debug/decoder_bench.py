"""decoder forward (with saved activations) stand-alone at C1 size; usage: python profiles/debug/decoder_bench.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mmpde_b200 import _cabi, ops
if os.environ.get("MMPDE_LIB"): _cabi.LIB_PATH = os.path.abspath(os.environ["MMPDE_LIB"])
dev = torch.device("cuda:0")
torch.manual_seed(0)
N = 36864
h = torch.randn(N, 128, device=dev)
dec = torch.randn(525, device=dev) * 0.2
out = torch.empty(N, device=dev); A1 = torch.empty(N, 256, device=dev); C2 = torch.empty(N, 128, device=dev)
fn = lambda: _cabi.call("mmpde_decoder_fwd_acts", ops._ptr(h), 128, N, ops._ptr(dec), 0.1, ops._ptr(out), ops._ptr(A1), ops._ptr(C2), ops._stream())
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): fn()
e1.record(); torch.cuda.synchronize()
print(f"mmpde_decoder_fwd_acts, {N} nodes: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch; checksum {float(out.double().sum()):.9e} {float(A1.double().sum()):.9e} {float(C2.double().sum()):.9e}")
