"""Parity cost of running one network at two bf16 planes instead of three (GPU box): image / uv_warp error against the
reference fixtures for each override of Model.PRECISION."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import test_model_gpu as T
from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
from emoportraits_b200.config import shipped_config
from emoportraits_b200.infer import Model

for size in (512, 256):
    cfg = shipped_config(size)
    sd, hsd = synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0)
    gold = torch.load(T.GOLD / f"va{size}_seed0.pt", weights_only=False)
    for name, pr in [("baseline", {}), ("warp=2", {"warp": 2}), ("expression=2", {"expression": 2}), ("head_pose=2", {"head_pose": 2})]:
        model = Model(cfg, sd, hsd, "cuda", precision=pr)
        for kind, inject in [("smooth", False), ("noise", True)]:
            errs = T._run_case(size, cfg, model, T._case(gold, kind), inject)
            img = max(v for k, v in errs.items() if k.endswith("img"))
            uv = max(v for k, v in errs.items() if k.endswith("uv_warp"))
            pe = max(v for k, v in errs.items() if k.endswith(".pose_embed"))
            print(f"## {size} {name:14s} {kind:6s}: img {img:.2e} uv_warp {uv:.2e} pose_embed {pe:.2e} xy_warp {errs['xy_warp']:.2e}", flush=True)
        del model
