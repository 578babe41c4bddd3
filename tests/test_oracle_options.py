"""CPU tests: the oracle restatement of the NON-DEFAULT InferenceWrapper.forward arguments (notebooks/infer.py:355-357:
mix_old, mix=False, target_theta=False, smooth_pose, custome_target_pose_embed, custome_target_theta_embed, source_mask /
driver_mask, c_source_latent_volume, c_target_latent_volume) pinned against fixtures recorded from the unmodified
reference (`python -m oracle.make_golden options`, tests/golden/va256_options.pt)."""
import pathlib

import pytest
import torch

GOLD = pathlib.Path(__file__).parent / "golden"
SIZE = 256


@pytest.fixture(scope="module")
def ctx():
    from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
    from emoportraits_b200.config import shipped_config
    from oracle import frames as FR
    from oracle import restatement as R
    from oracle.make_golden import option_inputs

    gold = torch.load(GOLD / f"va{SIZE}_options.pt", weights_only=False)
    cfg = shipped_config(SIZE)
    sd, hsd = synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0)
    ocfg = R.config_from_state_dict(sd, SIZE)
    src = FR.frame(SIZE, gold["src_seed"], gold["kind"])
    drv = [FR.frame(SIZE, s, gold["kind"]) for s in gold["drv_seeds"]]
    X = option_inputs(SIZE, cfg)
    with torch.no_grad():
        st = R.source_pass(sd, hsd, src, ocfg)
    return dict(gold=gold["cases"], sd=sd, hsd=hsd, ocfg=ocfg, src=src, drv=drv, X=X, st=st, R=R)


def _check(case, img, taps, tol_img=1e-3):
    ref, stride = case["img"]
    err = (img.detach().float().reshape(-1)[::stride] - ref).abs().max().item()
    assert err < tol_img, f"image max-abs {err}"
    assert (taps["theta"] - case["pred_target_theta"][:, :3]).abs().max().item() < 1e-5
    assert (taps["pose_embed"] - case["target_pose_embed"]).abs().max().item() < 1e-4


@pytest.mark.parametrize("name,kw", [
    ("default", {}),
    ("mix_old", dict(mix_old=True)),
    ("no_mix", dict(mix=False)),
    ("target_theta_false", dict(target_theta=False)),
])
def test_pose_options(ctx, name, kw):
    c, R = ctx, ctx["R"]
    taps = {}
    with torch.no_grad():
        img = R.driver_pass(c["sd"], c["hsd"], c["st"], c["drv"][0], c["ocfg"], taps, **kw)
    _check(c["gold"][name], img, taps)


def test_options_change_the_result(ctx):
    """the fixtures are not vacuous: each option moves the image away from the default case"""
    base = ctx["gold"]["default"]["img"][0]
    for name, case in ctx["gold"].items():
        if name in ("default", "smooth_pose_0"):
            continue
        assert (case["img"][0] - base).abs().max().item() > 2e-3, name
    # smoothing starts at the first frame's pose: frame 0 equals the unsmoothed default (momentum 0.5 is exact in fp32)
    assert (ctx["gold"]["smooth_pose_0"]["img"][0] - base).abs().max().item() < 1e-6


def test_smooth_pose(ctx):
    c, R = ctx, ctx["R"]
    state = {"theta": None}
    with torch.no_grad():
        for i, d in enumerate(c["drv"]):
            taps = {}
            img = R.driver_pass(c["sd"], c["hsd"], c["st"], d, c["ocfg"], taps, smooth=state, pose_momentum=0.5)
            _check(c["gold"][f"smooth_pose_{i}"], img, taps)


def test_custom_embeddings(ctx):
    c, R, X = ctx, ctx["R"], ctx["X"]
    with torch.no_grad():
        taps = {}
        img = R.driver_pass(c["sd"], c["hsd"], c["st"], c["drv"][0], c["ocfg"], taps, custome_target_pose_embed=X["pose_embed"])
        _check(c["gold"]["custome_target_pose_embed"], img, taps)
        taps = {}
        img = R.driver_pass(c["sd"], c["hsd"], c["st"], c["drv"][0], c["ocfg"], taps, custome_target_theta_embed=X["theta_embed"])
        _check(c["gold"]["custome_target_theta_embed"], img, taps)


def test_source_mask_and_custom_volumes(ctx):
    c, R, X = ctx, ctx["R"], ctx["X"]
    with torch.no_grad():
        st = R.source_pass(c["sd"], c["hsd"], c["src"], c["ocfg"], src_mask=X["source_mask"])
        g = c["gold"]["source_mask"]
        assert (st["idt_embed"] - g["idt_embed"]).abs().max().item() < 1e-4
        assert (st["source_theta"] - g["pred_source_theta"]).abs().max().item() < 1e-5   # the regressor sees the unmasked image
        taps = {}
        _check(g, R.driver_pass(c["sd"], c["hsd"], st, c["drv"][0], c["ocfg"], taps), taps)
        for key in ("c_source_latent_volume", "c_target_latent_volume"):
            st = R.source_pass(c["sd"], c["hsd"], c["src"], c["ocfg"], **{key: X[key]})
            taps = {}
            _check(c["gold"][key], R.driver_pass(c["sd"], c["hsd"], st, c["drv"][0], c["ocfg"], taps), taps)
