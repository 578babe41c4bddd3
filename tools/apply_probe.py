"""HBM roofline of the elementwise passes between the convolutions (GPU box): GroupNorm-apply + ReLU + plane split
(emo_apply) at the decoder's shapes, the fused image head (emo_gn_head) and the fp32 -> planes split, timed as CUDA-graph
replays of back-to-back launches on tensors larger than L2 where the layer is (smaller layers are L2-resident in the
model too).  Algorithmic bytes: 4 B read + 2 x 2 B written per element (x4 written with nearest-x2)."""
import json, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from emoportraits_b200 import ops

dev = "cuda"
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1000  # us


for (S, C, up, planes) in [(512, 128, 1, 2), (256, 192, 1, 2), (256, 192, 2, 2), (128, 320, 1, 2), (64, 512, 1, 2), (64, 512, 1, "h2"), (32, 64, 1, "h2")]:
    x = torch.randn(1, 1, S, S, C, device=dev)
    ops.begin_pass(dev)
    st = ops.gn_stats(x, 32)
    gn = dict(stats=st, count=x.numel() / 32, gamma=torch.ones(C, device=dev), beta=torch.zeros(C, device=dev))
    us = timed(lambda: ops.apply(x, gn=gn, act=ops.ACT_RELU, up=up, planes=planes))
    by = x.numel() * (4 + 4 * up * up)
    print(f"apply {S}^2 x {C} up{up} planes {planes}: {us:8.1f} us  {by / us / 1e3:7.0f} GB/s  {by / us / 1e3 / peak:.2f} of measured HBM peak ({by / 1e6:.0f} MB)")
x = torch.randn(1, 1, 512, 512, 128, device=dev)
ops.begin_pass(dev)
st = ops.gn_stats(x, 32)
gn = dict(stats=st, count=x.numel() / 32, gamma=torch.ones(128, device=dev), beta=torch.zeros(128, device=dev))
w, b = torch.randn(3, 128, device=dev), torch.zeros(3, device=dev)
us = timed(lambda: ops.gn_head(x, gn, w, b, act_out=ops.ACT_SIGMOID))
by = x.numel() * 4 + 3 * 512 * 512 * 4
print(f"gn_head 512^2 x 128 -> 3: {us:8.1f} us  {by / us / 1e3:7.0f} GB/s  {by / us / 1e3 / peak:.2f} of measured HBM peak")
us = timed(lambda: ops.split_bf16(x, 2))
by = x.numel() * 8
print(f"split 512^2 x 128: {us:8.1f} us  {by / us / 1e3:7.0f} GB/s  {by / us / 1e3 / peak:.2f} of measured HBM peak")
us = timed(lambda: ops.gn_stats(x, 32))
by = x.numel() * 4
print(f"gn_stats 512^2 x 128: {us:8.1f} us  {by / us / 1e3:7.0f} GB/s  {by / us / 1e3 / peak:.2f} of measured HBM peak")
