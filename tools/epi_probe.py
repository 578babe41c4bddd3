"""Store-warp epilogue (EMO_CONV_EPI=1) against the in-warp final phase (=0), per feature.  (GPU box, scratch tool;
run once per setting of EMO_CONV_EPI because the library reads it once.)"""
import math, os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from emoportraits_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(0)

def bench(Cin, Cout, sp, k, planes, use_res, use_stats, use_bias, reps=10):
    shape = (1, 1) + tuple(sp) + (Cin,)
    x = torch.randn(shape, generator=g).to(dev)
    w = torch.randn((Cout, Cin, k, k), generator=g) / math.sqrt(Cin * k * k)
    pw = ops.pack_conv_weight(w, planes=planes)
    a = ops.split_bf16(x, planes)
    out = torch.empty(shape[:-1] + (Cout,), device=dev)
    res = torch.randn(shape[:-1] + (Cout,), generator=g).to(dev) if use_res else None
    st = ops.new_stats(1, 32, dev) if use_stats else None
    bias = torch.zeros(Cout, device=dev) if use_bias else None
    f = lambda: ops.conv_igemm(a, pw, out=out, bias=bias, residual=res, stats=st)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            f()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1000

if len(sys.argv) > 1:
    for dbg in [0, 32, 128, 2, 2 + 128, 1 + 2, 1 + 2 + 128]:
        os.environ["EMO_CONV_DBG"] = str(dbg)
        print("dbg", dbg, " ".join(f"{bench(128, 128, (512, 512), 3, 2, r, s_, 1):7.1f}" for r, s_ in [(0, 0), (1, 1)]), flush=True)
    sys.exit(0)
print("EMO_CONV_EPI =", os.environ.get("EMO_CONV_EPI", "(default 1)"))
for (Cin, Cout, sp) in [(128, 128, (512, 512)), (512, 512, (64, 64)), (192, 192, (256, 256))]:
    row = []
    for name, r, s, b in [("plain", 0, 0, 0), ("+bias", 0, 0, 1), ("+res", 1, 0, 1), ("+stats", 0, 1, 1), ("+res+stats", 1, 1, 1)]:
        row.append(f"{name} {bench(Cin, Cout, sp, 3, 2, r, s, b):7.1f}")
    print(f"{Cin}->{Cout} {sp}: " + " | ".join(row), flush=True)
