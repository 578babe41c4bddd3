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


SLOT_SCRIPT = r"""
import torch, sys
sys.path.insert(0, %r)
from emoportraits_b200 import ops
# one arena / workspace set per (device, slot); the device key does not depend on how the device is spelled
assert ops._dev_key("cpu") == ops._dev_key(torch.device("cpu")) == ops._dev_key(torch.zeros(1).device)
ops.begin_pass("cpu")
a0 = ops.new_stats(1, 32, torch.zeros(1).device)
prev = ops.set_slot(1)
assert prev == 0
ops.begin_pass(torch.device("cpu"))
a1 = ops.new_stats(1, 32, "cpu")
assert a0.data_ptr() != a1.data_ptr()
assert ops.set_slot(prev) == 1
ops.begin_pass("cpu")
b0 = ops.new_stats(1, 32, "cpu")
assert b0.data_ptr() == a0.data_ptr()          # begin_pass rewinds the slot's arena
assert ops.new_stats(1, 32, "cpu").data_ptr() != b0.data_ptr()
print("ok")
"""


def test_scratch_slots_and_device_key():
    env = dict(os.environ, EMO_DRY_RUN="1")
    r = subprocess.run([sys.executable, "-c", SLOT_SCRIPT % str(ROOT)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


OPTIONS_SCRIPT = r"""
import torch, sys
sys.path.insert(0, %r)
from emoportraits_b200 import lib as L
assert L.DRY_RUN
from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
from emoportraits_b200.config import shipped_config
from emoportraits_b200.infer import Model
size = 256
cfg = shipped_config(size)
model = Model(cfg, synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0), "cpu")
x = torch.rand(1, 3, size, size)
vol = torch.rand(1, cfg.C, cfg.D, cfg.S, cfg.S)
st = model.source_pass(x, mask=torch.rand(1, 1, size, size), c_source_latent_volume=vol, c_target_latent_volume=vol)
assert st.source_latent_volume.shape == (1, cfg.D, cfg.S, cfg.S, cfg.C)
assert torch.equal(st.target_latent_volume_1, vol.permute(0, 2, 3, 4, 1))      # the caller's volume, channels-last
try:
    model.source_pass(x, c_source_latent_volume=vol[:, :, :1])
    raise SystemExit("bad volume shape accepted")
except ValueError:
    pass
state = torch.zeros(3, 4)
n0 = L.launch_count
img, _, _, so = model.driver_pass(st, x, mix=True, mix_old=True, custom_srt=torch.rand(1, 9), smooth_state=state,
                                  smooth_momentum=0.5, smooth_init=True)
full = L.launch_count - n0
n0 = L.launch_count
img2, _, _, so2 = model.driver_pass(st, x, mix=True, custom_pose_embed=torch.rand(1, 128))
assert img.shape == img2.shape == (1, 3, size, size)
assert L.launch_count - n0 < full        # the expression encoder is skipped when its output is replaced
assert so2.target_pose_embed.shape == (1, 128)
# sub-pixel up-sampling convolutions: the first conv of every up block reads the low-resolution planes
ps = [b.c1_ps for b in model.decoder_nw.img if b.c1_ps is not None]
assert ps and ps[0].hi.shape[0] == 16 and ps[0].hi.shape[1] == ps[0].cout   # (the 160 -> 96 block of the 256^2 config keeps the plain form)
assert all(b.c2_pool is not None for b in model.local_encoder_nw.blocks)      # conv -> avgpool folded into 4x4 stride-2 convs
_, deep_f, img_f, _ = model.driver_pass(st, x, mix=True)
assert img_f.shape == (1, 1, size, size, cfg.dec_channels[-1])
print("ok")
"""


def test_dry_run_forward_options():
    env = dict(os.environ, EMO_DRY_RUN="1")
    r = subprocess.run([sys.executable, "-c", OPTIONS_SCRIPT % str(ROOT)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


H2_SCRIPT = r"""
import torch, sys
sys.path.insert(0, %r)
from emoportraits_b200 import lib as L, ops
assert L.DRY_RUN
from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
from emoportraits_b200.config import shipped_config
from emoportraits_b200.infer import Model
cfg = shipped_config(256)
m = Model(cfg, synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0), "cpu")
assert m.precision["warp"] == ops.H2 and m.precision["expression"] == ops.H2 and m.precision["decoder"] == 2, m.precision
x = torch.rand(1, 3, 256, 256)
st = m.source_pass(x)
img, _, _, _ = m.driver_pass(st, x, mix=True)
assert img.shape == (1, 3, 256, 256)
# load-time weight planes: fp16 of w * 256, reproducing w to ~2^-22
w = torch.randn(32, 64, 3, 3) * 0.05
pw = ops.pack_conv_weight(w, device="cpu", planes=ops.H2)
assert pw.f16 and pw.hi.dtype == torch.float16 and pw.scale == ops.F16_W_SCALE and pw.acc_chunk == ops.ACC_CHUNK_F16
rec = (pw.hi.float() + pw.lo.float()) / pw.scale                       # [taps][Cout_pad][Cin]
ref = w.permute(2, 3, 0, 1).reshape(9, 32, 64)
assert (rec - ref).abs().max().item() < 2 ** -21 * ref.abs().max().item()
a = ops.Split.empty((1, 1, 4, 4, 8), "cpu", ops.H2)
assert a.f16 and a.hi.dtype == torch.float16 and a.scale == ops.F16_ACT_SCALE and a.view(1, 1, 16, 1, 8).f16
try:
    Model(cfg, synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0), "cpu", precision=dict(decoder=ops.H2))
    raise SystemExit("decoder=h2 must be refused")
except ValueError as e:
    assert "decoder" in str(e)
print("ok")
"""


def test_dry_run_fp16_two_plane_networks():
    env = dict(os.environ, EMO_DRY_RUN="1")
    r = subprocess.run([sys.executable, "-c", H2_SCRIPT % str(ROOT)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


MASK_SCRIPT = r"""
import torch, sys
sys.path.insert(0, %r)
from emoportraits_b200 import lib as L
assert L.DRY_RUN
from emoportraits_b200.masks import FaceParsing, modnet_get_mask
from oracle.stub_nets import StubBiSeNet, StubMODNet
fp = FaceParsing(None, device="cpu", net=StubBiSeNet())
n0 = L.launch_count
m = fp.forward(torch.rand(2, 3, 300, 400))
assert len(m) == 4 and all(t.shape == (2, 1, 300, 400) and t.dtype == torch.int64 for t in m)
assert L.launch_count - n0 == 2                      # normalise + resize, resize + argmax + label sets
matte = modnet_get_mask(StubMODNet(), torch.rand(1, 3, 300, 400))
assert matte.shape == (1, 1, 300, 400)               # 300 x 400 -> 512 x 672 for the network -> back (infer.py:663-682)
assert L.launch_count - n0 == 4
print("ok")
"""


def test_dry_run_mask_processing_shapes():
    env = dict(os.environ, EMO_DRY_RUN="1")
    r = subprocess.run([sys.executable, "-c", MASK_SCRIPT % str(ROOT)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
