import sys, torch
sys.path.insert(0, '/root/repo')
from mmpde_b200.mesh.dmm_model import DMM
from mmpde_b200 import synthetic
dev = torch.device('cuda:0')
torch.manual_seed(11)
mover = DMM(s=48, mode="array", branch_layer=7, trunk_layer=[2, 32, 512], out_layer=[1024, 512, 1]).to(dev).eval()
g = torch.linspace(0, 1, 48)
grid = torch.stack(torch.meshgrid(g, g, indexing="xy"), -1).reshape(-1, 2)
u = synthetic.burgers_fields(16, 31, 48, 48, seed=1)[:, 9].to(dev)
xi = grid[None].expand(16, -1, -1).reshape(-1, 2).to(dev).contiguous()
from mmpde_b200 import _cabi, ops
t1, t2 = mover.trunk.layers
o1, o2 = mover.out_nn.layers
n_lat = o1.weight.shape[1] - t2.weight.shape[0]
Wl, Wt = o1.weight[:, :n_lat], o1.weight[:, n_lat:]
with torch.no_grad():
    const = (mover._latent(u).reshape(16, -1) @ Wl.t() + (Wt @ t2.bias + o1.bias)).contiguous()
    M = (Wt @ t2.weight).contiguous()
    w = o2.weight.reshape(-1).contiguous()
out = torch.empty(xi.shape[0], 2, device=dev)
W1 = t1.weight.detach().contiguous()
fn = lambda: _cabi.call("mmpde_dmm_displacement", ops._ptr(xi), ops._ptr(W1), ops._ptr(t1.bias), 32, ops._ptr(M), ops._ptr(const), ops._ptr(w), 512,
                        xi.shape[0], 2304, ops._ptr(out), ops._stream())
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): fn()
e1.record(); torch.cuda.synchronize()
print(f"mmpde_dmm_displacement, 36 864 points: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
