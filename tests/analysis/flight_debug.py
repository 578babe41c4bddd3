"""scratch: in-flight pipeline against the eager pass, several models / sizes in one process (module-order repro)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[2]))
import torch
from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
from emoportraits_b200.config import shipped_config
from emoportraits_b200.infer import Model, DriverPipeline
from oracle import frames as FR

def run(S, tag, host):
    cfg = shipped_config(S)
    model = Model(cfg, synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0), "cuda")
    st = model.source_pass(FR.frame(S, 41, "smooth").cuda())
    drv = [FR.frame(S, 50 + i, "smooth").cuda() for i in range(5)]
    want = [model.driver_pass(st, d, mix=True)[0].clone() for d in drv]
    pipe = DriverPipeline(model, st, depth=2, mix=True)
    for rep in range(2):
        if host:
            outs = [torch.empty((1, 3, S, S)).pin_memory() for _ in drv]
            for d, o in zip(drv, outs):
                pipe.submit(d, host_out=o)
        else:
            outs = [torch.empty_like(drv[0]) for _ in drv]
            for d, o in zip(drv, outs):
                pipe.submit(d, dev_out=o)
        pipe.drain(); torch.cuda.synchronize()
        print(f"{tag} S={S} host={host} rep {rep}:", [f"{(a.cpu() - b.cpu()).abs().max().item():.1e}" for a, b in zip(want, outs)], flush=True)
    want2 = [model.driver_pass(st, d, mix=True)[0].clone() for d in drv]
    print(f"{tag} S={S} eager again:", [f"{(a - b).abs().max().item():.1e}" for a, b in zip(want, want2)], flush=True)

run(256, "A", True)
run(256, "B", False)
run(512, "C", True)
run(512, "D", True)
