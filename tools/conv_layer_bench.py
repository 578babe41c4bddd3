"""Per-layer device time of the driver frame's convolution shapes (GPU box): CUDA-graph replays of 10 back-to-back launches,
so neither host launch cost nor tensor-map encoding is in the number (bench.py's per-layer table is taken in eager mode and
carries both).  With EMO_CONV_EPI (0 in-warp final phase / 1 TMA epilogue) or EMO_CONV_YREUSE (0 per-tap main loop) set, the
instrumented build libemoport_dbg.so is loaded, which honours them (the product library selects per layer by itself); one
setting per process."""
import math, os, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
if os.environ.get("EMO_CONV_EPI") or os.environ.get("EMO_CONV_YREUSE"):
    os.environ.setdefault("EMO_LIB", str(ROOT / "emoportraits_b200" / "csrc" / "libemoport_dbg.so"))
import torch
from emoportraits_b200 import ops

dev = "cuda"
g = torch.Generator().manual_seed(0)
# (label, Cin, Cout, spatial, k, planes, residual: 0 none / 1 same resolution / 2 half resolution, sub-pixel up-conv, launches per frame)
LAYERS = [
    ("dec res 64^2 512->512 conv1", 512, 512, (64, 64), 3, 2, 0, False, 6),
    ("dec res 64^2 512->512 conv2+res", 512, 512, (64, 64), 3, 2, 1, False, 6),
    ("dec in 1x1 1536->512", 1536, 512, (64, 64), 1, 2, 0, False, 1),
    ("up 64->128 512->320 subpixel", 512, 320, (64, 64), 3, 2, 0, True, 1),
    ("128^2 320->320 conv2+res/2", 320, 320, (128, 128), 3, 2, 2, False, 1),
    ("128^2 320->320 conv1", 320, 320, (128, 128), 3, 2, 0, False, 1),
    ("128^2 320->320 conv2+res", 320, 320, (128, 128), 3, 2, 1, False, 1),
    ("up 128->256 320->192 subpixel", 320, 192, (128, 128), 3, 2, 0, True, 1),
    ("256^2 192->192 conv2+res/2", 192, 192, (256, 256), 3, 2, 2, False, 1),
    ("256^2 192->192 conv1", 192, 192, (256, 256), 3, 2, 0, False, 1),
    ("256^2 192->192 conv2+res", 192, 192, (256, 256), 3, 2, 1, False, 1),
    ("up 256->512 192->128 subpixel", 192, 128, (256, 256), 3, 2, 0, True, 1),
    ("512^2 128->128 conv2+res/2", 128, 128, (512, 512), 3, 2, 2, False, 1),
    ("512^2 128->128 conv1", 128, 128, (512, 512), 3, 2, 0, False, 1),
    ("512^2 128->128 conv2+res", 128, 128, (512, 512), 3, 2, 1, False, 1),
    ("warp 8^3 512->256", 512, 256, (8, 8, 8), 3, "h2", 0, False, 1),
    ("warp 16^3 256->128", 256, 128, (16, 16, 16), 3, "h2", 0, False, 1),
    ("warp 16^3 128->128+res", 128, 128, (16, 16, 16), 3, "h2", 1, False, 1),
    ("warp 32^3 128->64", 128, 64, (32, 32, 32), 3, "h2", 0, False, 1),
    ("warp 32^3 64->64+res", 64, 64, (32, 32, 32), 3, "h2", 1, False, 1),
    ("warp 32x64x64 64->32", 64, 32, (32, 64, 64), 3, "h2", 0, False, 1),
    ("warp 32x64x64 32->32+res", 32, 32, (32, 64, 64), 3, "h2", 1, False, 1),
    # ResNet-18 tails (two networks per frame): split-K convolution + fused finalize / GroupNorm / ReLU / plane split (residual = 3)
    ("resnet 32^2 64->64 conv+post", 64, 64, (32, 32), 3, "h2", 3, False, 8),
    ("resnet 16^2 128->128 conv+post", 128, 128, (16, 16), 3, "h2", 3, False, 6),
    ("resnet 8^2 256->256 conv+post", 256, 256, (8, 8), 3, "h2", 3, False, 6),
    ("resnet 4^2 512->512 conv+post", 512, 512, (4, 4), 3, "h2", 3, False, 6),
]


def bench(Cin, Cout, sp, k, planes, residual, upconv, reps=10):
    three = len(sp) == 3
    shape = (1,) + (tuple(sp) if three else (1,) + tuple(sp)) + (Cin,)
    x = torch.randn(shape, generator=g).to(dev)
    w = torch.randn((Cout, Cin) + (k,) * len(sp), generator=g) / math.sqrt(Cin * k ** len(sp))
    a = ops.split_bf16(x, planes)
    osp = tuple(2 * v for v in sp) if upconv else tuple(sp)
    oshape = (1,) + (osp if three else (1,) + osp) + (Cout,)
    pw = ops.pack_upconv_weight(w) if upconv else ops.pack_conv_weight(w, planes=planes)
    out = torch.empty(oshape, device=dev)
    res, rs = None, 0
    post = None
    if residual == 3:
        post = True
    elif residual == 1:
        res = torch.randn(oshape, generator=g).to(dev)
    elif residual == 2:
        res, rs = torch.randn((1, 1, osp[0] // 2, osp[1] // 2, Cout), generator=g).to(dev), 1
    ops.begin_pass(dev)
    st = ops.new_stats(1, 32, dev)
    bias = torch.zeros(Cout, device=dev)
    chunk = int(os.environ.get("EMO_ACC_CHUNK", "0"))  # MMAs per TMEM accumulation chunk (0 = the library's default: 48 / 24)
    if post:
        gn = dict(stats=st, count=math.prod(osp) * Cout / 32, gamma=torch.ones(Cout, device=dev), beta=torch.zeros(Cout, device=dev))
        run = lambda: ops.conv_igemm(a, pw, bias=bias, stats=st, acc_chunk_mmas=chunk, post=dict(gn=gn, act=ops.ACT_RELU, planes=planes))
    else:
        run = lambda: ops.conv_igemm(a, pw, out=out, bias=bias, residual=res, res_shift=rs, stats=st, upconv=upconv, acc_chunk_mmas=chunk)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            run()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1000


tot = 0.0
tag = f"EPI={os.environ.get('EMO_CONV_EPI', 'default')} YREUSE={os.environ.get('EMO_CONV_YREUSE', 'default')} CHUNK={os.environ.get('EMO_ACC_CHUNK', 'default')}"
for (label, Cin, Cout, sp, k, planes, residual, upconv, n) in LAYERS:
    us = bench(Cin, Cout, sp, k, planes, residual, upconv)
    fl = 2.0 * Cout * Cin * k ** len(sp) * math.prod(sp) * (4 if upconv else 1)
    tot += us * n
    print(f"{tag} | {label:34s} {us:8.1f} us x{n}  {fl / us / 1e6:7.1f} TFLOP/s algorithmic", flush=True)
print(f"{tag} | sum over the frame's launches of these shapes: {tot / 1000:.3f} ms")
