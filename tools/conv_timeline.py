"""Where does a conv launch spend its time?  (GPU box, instrumented build libemoport_dbg.so.)  One launch with
EMO_CONV_DBG=256 leaves per-CTA %globaltimer stamps of the kernel's phases (conv_igemm.cu: g_conv_stamps); printed here
relative to the earliest CTA entry, as min / median / max over the CTAs, next to the CUDA-event time of the launch."""
import ctypes as C, math, os, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("EMO_LIB", str(ROOT / "emoportraits_b200" / "csrc" / "libemoport_dbg.so"))
import numpy as np
import torch
from emoportraits_b200 import lib as L, ops

NAMES = ["entry", "prologue done", "producer starts", "producer done", "first operands landed", "last MMA issued",
         "last chunk consumed", "final-phase stores issued (TMA epilogue: staging tile free)", "statistics done (TMA epilogue: staging loop done)", "teardown barrier", "TMEM freed",
         "TMA epilogue: staging tile written", "TMA epilogue: stores issued", "TMA epilogue: statistics done", "TMA epilogue: staging tile read by the store"]
dev = "cuda"
g = torch.Generator().manual_seed(0)


def run(Cin, Cout, sp, k, planes, residual=True):
    three = len(sp) == 3
    shape = (1,) + (tuple(sp) if three else (1,) + tuple(sp)) + (Cin,)
    x = torch.randn(shape, generator=g).to(dev)
    w = torch.randn((Cout, Cin) + (k,) * len(sp), generator=g) / math.sqrt(Cin * k ** len(sp))
    pw = ops.pack_conv_weight(w, planes=planes)
    a = ops.split_bf16(x, planes)
    out = torch.empty(shape[:-1] + (Cout,), device=dev)
    res = torch.randn(shape[:-1] + (Cout,), generator=g).to(dev) if residual else None
    ops.begin_pass(dev)
    st = ops.new_stats(1, 32, dev)
    bias = torch.zeros(Cout, device=dev)
    os.environ["EMO_CONV_DBG"] = "0"
    for _ in range(3):
        ops.conv_igemm(a, pw, out=out, bias=bias, residual=res, stats=st)
    torch.cuda.synchronize()
    os.environ["EMO_CONV_DBG"] = "256"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv_igemm(a, pw, out=out, bias=bias, residual=res, stats=st)
    e1.record()
    torch.cuda.synchronize()
    os.environ["EMO_CONV_DBG"] = "0"
    buf = (C.c_ulonglong * (160 * 16))()
    fn = L.load().emo_debug_conv_stamps
    fn.restype = C.c_int
    assert fn(buf) == 0
    s = np.array(buf, dtype=np.uint64).reshape(160, 16).astype(np.float64)
    live = s[:, 0] > 0
    s = s[live]
    t0 = s[:, 0].min()
    print(f"--- {Cin}->{Cout} {sp} k{k} planes {planes} EMO_CONV_EPI={os.environ.get('EMO_CONV_EPI', 'default')}: {int(live.sum())} CTAs, "
          f"launch {e0.elapsed_time(e1) * 1000:.1f} us (events), last stamp {(s[:, :15].max() - t0) / 1000:.1f} us after the first entry")
    for k_, name in enumerate(NAMES):
        col = s[:, k_]
        col = col[col > 0]
        if len(col) == 0:
            continue
        r = (col - t0) / 1000.0
        print(f"    {name:28s} min {r.min():7.2f}  median {np.median(r):7.2f}  max {r.max():7.2f} us   ({len(col)} CTAs)")
    # zero the stamps for the next shape
    return


for epi in sys.argv[1:] or ["auto"]:
    if epi != "auto":
        os.environ["EMO_CONV_EPI"] = epi   # 0 in-warp final phase, 1 TMA epilogue; "auto" = the library's per-layer choice
    run(512, 512, (64, 64), 3, 2)
    run(128, 128, (512, 512), 3, 2)
    run(320, 320, (128, 128), 3, 2)
    run(64, 32, (32, 64, 64), 3, "h2", residual=False)
    break
