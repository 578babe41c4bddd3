"""Representative launches of the two roofline kernels for `ncu --set full` captures (run on the GPU box):
   conv_igemm_kernel: decoder 3x3 128->128 @512^2 and 512->512 @64^2 (2 planes), warp-generator 3x3x3 64->32 (3 planes)
   gs3_cl_kernel:     96ch 64^3 volume, jittered grid and fused affine lattice."""
import math, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from emoportraits_b200 import ops

dev = "cuda"
g = torch.Generator().manual_seed(0)

def conv(N, Cin, Cout, sp, k, planes):
    three = len(sp) == 3
    shape = (N,) + (tuple(sp) if three else (1,) + tuple(sp)) + (Cin,)
    x = torch.randn(shape, generator=g).to(dev)
    w = torch.randn((Cout, Cin) + (k,) * len(sp), generator=g) / math.sqrt(Cin * k ** len(sp))
    pw = ops.pack_conv_weight(w, planes=planes)
    a = ops.split_bf16(x, planes)
    res = torch.randn(shape[:-1] + (Cout,), generator=g).to(dev)
    st = ops.new_stats(N, 32, dev)
    for _ in range(3):
        ops.conv_igemm(a, pw, bias=torch.zeros(Cout, device=dev), residual=res, stats=st)
    torch.cuda.synchronize()

conv(1, 128, 128, (512, 512), 3, 2)
conv(1, 512, 512, (64, 64), 3, 2)
conv(1, 64, 32, (32, 64, 64), 3, 3)

C, S, D = 96, 64, 64
vol = torch.randn(1, D, S, S, C, generator=g).to(dev)
zs, ys = torch.linspace(-1, 1, D), torch.linspace(-1, 1, S)
w, v, u = torch.meshgrid(zs, ys, ys, indexing="ij")
grid = (torch.stack([u, v, w], -1)[None] + 0.1 * torch.randn(1, D, S, S, 3, generator=g)).contiguous().to(dev)
a = math.radians(30)
theta = torch.tensor([[[math.cos(a), -math.sin(a), 0, 0.2], [math.sin(a), math.cos(a), 0, 0.2], [0, 0, 1.0, 0.2]]]).to(dev)
for _ in range(3):
    ops.grid_sample3d(vol, grid=grid, in_layout="cl")
for _ in range(3):
    ops.grid_sample3d(vol, theta=theta, out_size=(D, S, S), in_layout="cl")
torch.cuda.synchronize()
print("done")
