"""CPU tests: pin the oracle restatement against fixtures produced by the unmodified reference
(oracle/make_golden.py), and check the checkpoint-layout contract."""
import ast
import pathlib

import numpy as np
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


def frame(size, seed):
    a = (np.random.RandomState(seed).rand(size, size, 3) * 255).astype(np.uint8)
    return torch.from_numpy(a).permute(2, 0, 1)[None].float().div(255)


def _cmp_sub(name, got, ref_pair, tol):
    ref, stride = ref_pair
    g = got.detach().float().reshape(-1)[::stride]
    err = (g - ref).abs().max().item()
    assert err < tol, f"{name}: max abs err {err} (ref max {ref.abs().max().item()})"


@pytest.mark.parametrize("size", [256])
def test_restatement_matches_reference_golden(size):
    """oracle/restatement.py == unmodified reference (notebooks/infer.py InferenceWrapper) on the same seeded
    checkpoint and frames, stage by stage."""
    from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
    from emoportraits_b200.config import shipped_config
    from oracle import restatement as R

    gold = torch.load(GOLD / f"va{size}_seed0.pt", weights_only=False)
    cfg = shipped_config(size)
    sd, hsd = synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0)
    ocfg = R.config_from_state_dict(sd, size)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        st = R.source_pass(sd, hsd, frame(size, gold["src_seed"]), ocfg)
        s = gold["source"]
        assert (st["idt_embed"] - s["idt_embed"]).abs().max().item() < 1e-4
        assert (st["source_theta"] - s["pred_source_theta"]).abs().max().item() < 1e-5
        _cmp_sub("target_latent_volume", st["target_latent_volume"], s["target_latent_volume"], 2e-3)
        for fr in gold["frames"]:
            taps = {}
            img = R.driver_pass(sd, hsd, st, frame(size, fr["seed"]), ocfg, taps)
            assert (taps["theta"] - fr["pred_target_theta"][:, :3]).abs().max().item() < 1e-5
            assert (taps["pose_embed"] - fr["target_pose_embed"]).abs().max().item() < 1e-4
            _cmp_sub("uv_warp", taps["uv_warp"], fr["uv_warp"], 1e-4)
            assert (taps["logits"] - fr["logits"]).abs().max().item() < 2e-3
            assert (img - fr["img"]).abs().max().item() < 1e-3
