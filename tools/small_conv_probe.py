"""Where does the time of the small-M convs go?  (GPU box)  Times CUDA-graph replays of 20 back-to-back launches."""
import math, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from emoportraits_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(0)
def bench(N, Cin, Cout, sp, k, planes, chunk):
    three = len(sp) == 3
    shape = (N,) + (tuple(sp) if three else (1,) + tuple(sp)) + (Cin,)
    x = torch.randn(shape, generator=g).to(dev)
    w = torch.randn((Cout, Cin) + (k,) * len(sp), generator=g) / math.sqrt(Cin * k ** len(sp))
    pw = ops.pack_conv_weight(w, planes=planes)
    a = ops.split_bf16(x, planes)
    out = torch.empty(shape[:-1] + (Cout,), device=dev)
    for _ in range(2):
        ops.conv_igemm(a, pw, out=out, acc_chunk_mmas=chunk)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20):
            ops.conv_igemm(a, pw, out=out, acc_chunk_mmas=chunk)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 100 * 1000
for (Cin, Cout, sp, k) in [(512, 512, (4, 4), 3), (256, 256, (8, 8), 3), (64, 64, (32, 32), 3), (512, 256, (8, 8, 8), 3), (64, 32, (32, 64, 64), 3), (128, 128, (512, 512), 3)]:
    for planes in (3, 2):
        row = []
        for chunk in (12, 24, 48, 96, 1000000):
            row.append(f"{bench(1, Cin, Cout, sp, k, planes, chunk):8.1f}")
        print(f"{Cin:4d}->{Cout:4d} {str(sp):14s} k{k} planes {planes}: us/launch for chunk 12/24/48/96/inf = " + " ".join(row), flush=True)
