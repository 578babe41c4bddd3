"""scratch: which part of the in-flight mismatch is concurrency?"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
from emoportraits_b200.config import shipped_config
from emoportraits_b200.infer import Model, DriverPipeline
from oracle import frames as FR
S = 512
cfg = shipped_config(S)
model = Model(cfg, synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0), "cuda")
st = model.source_pass(FR.frame(S, 41, "smooth").cuda())
drv = [FR.frame(S, 50 + i, "smooth").cuda() for i in range(5)]
want = [model.driver_pass(st, d, mix=True)[0].clone() for d in drv]
want2 = [model.driver_pass(st, d, mix=True)[0].clone() for d in drv]
print("eager vs eager:", [f"{(a - b).abs().max().item():.2e}" for a, b in zip(want, want2)])
for depth in (1, 2):
    pipe = DriverPipeline(model, st, depth=depth, mix=True)
    for rep in range(2):
        outs = [torch.empty_like(drv[0]) for _ in drv]
        for d, o in zip(drv, outs):
            pipe.submit(d, dev_out=o)
        pipe.drain(); torch.cuda.synchronize()
        print(f"depth {depth} rep {rep}:", [f"{(a - b).abs().max().item():.2e}" for a, b in zip(want, outs)])
