"""Probe: how does tcgen05.mma accumulate into the fp32 TMEM accumulator?  (run on the GPU box)
bf16-exact operands (lo plane == 0, products exact in fp32) so the only error is the accumulation."""
import math, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
import torch.nn.functional as F
from emoportraits_b200 import ops

torch.manual_seed(0)
def cl(x): return x[:, :, None].permute(0, 2, 3, 4, 1).contiguous()
for signed in (False, True):
    for Cin in (64, 128, 256, 512):
        x = (torch.randint(0, 64, (1, Cin, 32, 32)).float() / 64 + (0.0 if signed else 1.0))
        w = (torch.randint(0, 64, (128, Cin, 3, 3)).float() / 64 + (0.0 if signed else 1.0)) / Cin
        if signed:
            x = x * (torch.randint(0, 2, x.shape).float() * 2 - 1)
            w = w * (torch.randint(0, 2, w.shape).float() * 2 - 1)
        x = x.bfloat16().float(); w = w.bfloat16().float()
        ref = F.conv2d(x.double(), w.double(), padding=1)
        out = ops.conv_igemm(ops.split_bf16(cl(x).cuda()), ops.pack_conv_weight(w))
        got = out.cpu()[:, 0].permute(0, 3, 1, 2).double()
        ref32 = F.conv2d(x, w, padding=1).double()
        e = (got - ref) / ref.abs().mean()
        e32 = (ref32 - ref) / ref.abs().mean()
        print(f"signed={signed} K={Cin*9:5d}: tcgen05 rel err mean {e.mean():+.3e} rms {e.pow(2).mean().sqrt():.3e} max {e.abs().max():.3e}"
              f" | torch-fp32-CPU mean {e32.mean():+.3e} rms {e32.pow(2).mean().sqrt():.3e}   (n_mma={Cin*9//16*3})")
