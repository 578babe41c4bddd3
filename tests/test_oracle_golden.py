"""CPU tests: pin the oracle restatement against fixtures produced by the unmodified reference
(oracle/make_golden.py), and check the checkpoint-layout contract."""
import ast
import pathlib

import pytest
import torch

GOLD = pathlib.Path(__file__).parent / "golden"


def _manifest(path):
    out = {}
    for line in path.read_text().splitlines():
        k, shp = line.split(" ", 1)
        out[k] = tuple(ast.literal_eval(shp))
    return out


@pytest.mark.parametrize("size", [256, 512])
def test_layout_spec_matches_reference_manifest(size):
    from emoportraits_b200.checkpoint import head_pose_spec, state_dict_spec
    from emoportraits_b200.config import shipped_config

    ref = _manifest(GOLD / f"state_dict_manifest_{size}.txt")
    spec = {k: tuple(v) for k, v in state_dict_spec(shipped_config(size)).items()}
    assert spec == ref
    assert {k: tuple(v) for k, v in head_pose_spec().items()} == _manifest(GOLD / "head_pose_manifest.txt")


@pytest.mark.parametrize("size", [256, 512])
def test_args_txt_roundtrip(size):
    """the reference-written args.txt parses into the shipped hot-path configuration"""
    from emoportraits_b200.config import hot_path_config, parse_args, shipped_config

    cfg = hot_path_config(parse_args(GOLD / f"args_{size}.txt"))
    assert cfg == shipped_config(size)
    assert cfg.dec_channels == ([512, 320, 192, 128] if size == 512 else [256, 160, 96])
    assert cfg.warp_channels == [512, 256, 128, 64, 32]
    assert cfg.unet_channels == [96, 192, 384, 512]


def _cmp_sub(name, got, ref_pair, tol):
    ref, stride = ref_pair
    g = got.detach().float().reshape(-1)[::stride]
    err = (g - ref).abs().max().item()
    assert err < tol, f"{name}: max abs err {err} (ref max {ref.abs().max().item()})"


def _setup(size):
    from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
    from emoportraits_b200.config import shipped_config
    from oracle import restatement as R

    cfg = shipped_config(size)
    sd, hsd = synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0)
    return sd, hsd, R.config_from_state_dict(sd, size), R


@pytest.mark.parametrize("size,kind", [(256, "noise"), (256, "smooth")])
def test_restatement_matches_reference_golden(size, kind):
    """oracle/restatement.py == unmodified reference (notebooks/infer.py InferenceWrapper) on the same seeded
    checkpoint and frames, stage by stage."""
    from oracle import frames as FR

    gold = torch.load(GOLD / f"va{size}_seed0.pt", weights_only=False)
    case = next(c for c in gold["cases"] if c["kind"] == kind)
    sd, hsd, ocfg, R = _setup(size)
    with torch.no_grad():
        st = R.source_pass(sd, hsd, FR.frame(size, case["src_seed"], kind), ocfg)
        s = case["source"]
        assert (st["idt_embed"] - s["idt_embed"]).abs().max().item() < 1e-4
        assert (st["source_theta"] - s["pred_source_theta"]).abs().max().item() < 1e-5
        _cmp_sub("target_latent_volume", st["target_latent_volume"], s["target_latent_volume"], 2e-3)
        for fr in case["frames"]:
            taps = {}
            img = R.driver_pass(sd, hsd, st, FR.frame(size, fr["seed"], kind), ocfg, taps)
            assert (taps["theta"] - fr["pred_target_theta"][:, :3]).abs().max().item() < 1e-5
            assert (taps["pose_embed"] - fr["target_pose_embed"]).abs().max().item() < 1e-4
            _cmp_sub("uv_warp", taps["uv_warp"], fr["uv_warp"], 1e-4)
            _cmp_sub("logits", taps["logits"], fr["logits"], 2e-3)
            _cmp_sub("img", img, fr["img"], 1e-3)


@pytest.mark.parametrize("size", [256, 512])
def test_reference_noise_floor_of_white_noise_frames(size):
    """How reproducible is the REFERENCE's own arithmetic on the BASELINE white-noise frames?  Replace only its fp32 LU
    4x4 inverse (torch.inverse; infer.py:443, expression_embedder.py:168) by the exactly rounded inverse (fp64, rounded
    to fp32 - a <= 1-ulp change of the pose matrices) and the image moves by more than the 1e-3 parity budget; with
    smooth frames it does not.  This is why tests/test_model_gpu.py injects the reference's pose matrices for the
    white-noise case and runs the fully-on-device check on smooth frames."""
    from oracle import frames as FR

    sd, hsd, ocfg, R = _setup(size)

    def run(kind, exact_inverse):
        orig = torch.Tensor.inverse
        if exact_inverse:
            torch.Tensor.inverse = lambda self: orig(self.double()).float()
        try:
            with torch.no_grad():
                st = R.source_pass(sd, hsd, FR.frame(size, 0, kind), ocfg)
                return R.driver_pass(sd, hsd, st, FR.frame(size, 1, kind), ocfg)
        finally:
            torch.Tensor.inverse = orig

    d_noise = (run("noise", False) - run("noise", True)).abs().max().item()
    print(f"reference self-noise from a 1-ulp pose-matrix change @{size}^2: noise frames {d_noise:.2e}")
    assert d_noise > 3e-4, d_noise          # same order as the 1e-3 budget (measured 1.2e-3 @256^2 in the build container)
    if size == 256:                         # (the smooth-frame control runs once: the 512^2 oracle passes take ~15 s each)
        d_smooth = (run("smooth", False) - run("smooth", True)).abs().max().item()
        print(f"smooth frames {d_smooth:.2e}")
        assert d_smooth < 3e-4, d_smooth
