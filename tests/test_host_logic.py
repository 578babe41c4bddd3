"""CPU test of the host-side network assembly (shapes, layouts, launch order) with the kernels skipped
(EMO_DRY_RUN=1 — see emoportraits_b200/lib.py).  Numerical parity lives in the -m gpu tests."""
import os
import subprocess
import sys
import pathlib

ROOT = pathlib.Path(__file__).resolve().parents[1]

SCRIPT = r"""
import torch, sys
sys.path.insert(0, %r)
from emoportraits_b200 import lib as L
assert L.DRY_RUN
from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
from emoportraits_b200.config import shipped_config
from emoportraits_b200.infer import Model
size = int(sys.argv[1])
cfg = shipped_config(size)
model = Model(cfg, synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0), "cpu")
x = torch.rand(1, 3, size, size)
st = model.source_pass(x)
assert st.target_latent_volume.shape == (1, cfg.D, cfg.S, cfg.S, cfg.C), st.target_latent_volume.shape
assert st.idt_embed.shape == (1, 512, 4, 4)
n0 = L.launch_count
taps = {}
img, deep_f, img_f, so = model.driver_pass(st, x, mix=True, taps=taps)
assert img.shape == (1, 3, size, size), img.shape
assert deep_f.shape == (1, 1, cfg.S, cfg.S, cfg.dec_channels[0]) and img_f.shape == (1, 1, size, size, cfg.dec_channels[-1])
assert taps["uv_warp"].shape == (1, cfg.D, cfg.S, cfg.S, 3)
print("launches per driver frame:", L.launch_count - n0)
"""


def test_dry_run_shapes_256_and_512():
    env = dict(os.environ, EMO_DRY_RUN="1")
    for size in (256, 512):
        r = subprocess.run([sys.executable, "-c", SCRIPT % str(ROOT), str(size)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "launches per driver frame" in r.stdout
