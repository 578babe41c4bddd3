"""Run-to-run reproducibility of a driver frame (GPU box): two eager passes and the captured pipeline on the same inputs.
Prints the max-abs difference of the images (0 = bit-identical)."""
import pathlib, sys
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
from emoportraits_b200.config import shipped_config
from emoportraits_b200.infer import DriverPipeline, Model
from oracle import frames as FR

for size in (256, 512):
    cfg = shipped_config(size)
    model = Model(cfg, synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0), "cuda:0")
    for kind in ("smooth", "noise"):
        st = model.source_pass(FR.frame(size, 41, kind).cuda())
        st2 = model.source_pass(FR.frame(size, 41, kind).cuda())
        src_diff = (st.target_latent_volume - st2.target_latent_volume).abs().max().item()
        drv = [FR.frame(size, 50 + i, kind).cuda() for i in range(4)]
        a = [model.driver_pass(st, d, mix=True)[0].clone() for d in drv]
        b = [model.driver_pass(st, d, mix=True)[0].clone() for d in drv]
        pipe = DriverPipeline(model, st, depth=4, mix=True)
        outs = [torch.empty_like(a[0]) for _ in drv]
        for d, o in zip(drv, outs):
            pipe.submit(d, dev_out=o)
        pipe.drain()
        torch.cuda.synchronize()
        print(f"{size} {kind}: source pass twice {src_diff:.3e}; eager vs eager {max((x - y).abs().max().item() for x, y in zip(a, b)):.3e}; "
              f"pipeline vs eager {max((x - y).abs().max().item() for x, y in zip(outs, a)):.3e}")
