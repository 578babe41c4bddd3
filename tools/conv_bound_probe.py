"""Which resource bounds the conv main loop?  (GPU box; runs on the instrumented build libemoport_dbg.so that
   `python -m emoportraits_b200.csrc.build --debug` writes next to the product library)

Times CUDA-graph replays of back-to-back launches with parts of the kernel switched off (EMO_CONV_DBG bits:
1 no TMA loads, 2 no MMAs, 4 no tile-epilogue global traffic, 8 no TMEM chunk reads).  Outputs are garbage in those
modes; only the time matters."""
import math, os, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("EMO_LIB", str(ROOT / "emoportraits_b200" / "csrc" / "libemoport_dbg.so"))
import torch
from emoportraits_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(0)

def bench(N, Cin, Cout, sp, k, planes, dbg, reps=10):
    os.environ["EMO_CONV_DBG"] = str(dbg)
    three = len(sp) == 3
    shape = (N,) + (tuple(sp) if three else (1,) + tuple(sp)) + (Cin,)
    x = torch.randn(shape, generator=g).to(dev)
    w = torch.randn((Cout, Cin) + (k,) * len(sp), generator=g) / math.sqrt(Cin * k ** len(sp))
    pw = ops.pack_conv_weight(w, planes=planes)
    a = ops.split_bf16(x, planes)
    out = torch.empty(shape[:-1] + (Cout,), device=dev)
    res = torch.randn(shape[:-1] + (Cout,), generator=g).to(dev)
    st = ops.new_stats(N, 32, dev)
    bias = torch.zeros(Cout, device=dev)
    for _ in range(2):
        ops.conv_igemm(a, pw, out=out, bias=bias, residual=res, stats=st)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            ops.conv_igemm(a, pw, out=out, bias=bias, residual=res, stats=st)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    os.environ["EMO_CONV_DBG"] = "0"
    return e0.elapsed_time(e1) / (5 * reps) * 1000

if len(sys.argv) > 1 and sys.argv[1] == "epi":
    for base, bname in [(0, "full kernel"), (3, "epilogue alone (no TMA, no MMA)")]:
        row = [f"{name} {bench(1, 128, 128, (512, 512), 3, 2, base | dbg):7.1f}" for dbg, name in
               [(0, "all"), (16, "-res"), (32, "-store"), (64, "-stats"), (48, "-res-store"), (80, "-res-stats"), (96, "-store-stats"), (112, "none")]]
        print(f"128->128 512^2 {bname}: " + " | ".join(row), flush=True)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "full":
    # the whole kernel only, per decoder / warp-generator shape (EMO_CONV_EPI picks the final-phase form)
    for (Cin, Cout, sp, k, planes) in [(128, 128, (512, 512), 3, 2), (192, 192, (256, 256), 3, 2), (320, 320, (128, 128), 3, 2), (512, 512, (64, 64), 3, 2),
                                       (1536, 512, (64, 64), 1, 2), (64, 64, (32, 32, 32), 3, "h2"), (64, 32, (32, 64, 64), 3, "h2")]:
        print(f"{Cin:4d}->{Cout:4d} {str(sp):14s} k{k} p{planes} EMO_CONV_EPI={os.environ.get('EMO_CONV_EPI', '(default)')}: {bench(1, Cin, Cout, sp, k, planes, 0):7.1f} us/launch", flush=True)
    sys.exit(0)
MODES = [(0, "full"), (1, "noTMA"), (2, "noMMA"), (4, "noEpiGmem"), (8, "noTmemLd"), (12, "noEpi"), (3, "noTMA+noMMA"),
         (6, "TMAonly+chunks"), (14, "TMAonly"), (13, "MMAonly"), (15, "skeleton")]
for (Cin, Cout, sp, k, planes) in [(128, 128, (512, 512), 3, 2), (512, 512, (64, 64), 3, 2), (320, 320, (128, 128), 3, 2),
                                   (192, 192, (256, 256), 3, 2), (64, 32, (32, 64, 64), 3, 3)]:
    row = [f"{name} {bench(1, Cin, Cout, sp, k, planes, dbg):7.1f}" for dbg, name in MODES]
    print(f"{Cin:4d}->{Cout:4d} {str(sp):14s} k{k} p{planes} us/launch: " + " | ".join(row), flush=True)
