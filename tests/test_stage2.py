"""Stage-2 refinement path (BASELINE config 5): oracle pin on CPU, parity on the GPU."""
import ast
import pathlib

import pytest
import torch

from oracle import frames as FR

GOLD = pathlib.Path(__file__).parent / "golden"


def _sub_err(got, ref_pair):
    vals, stride = ref_pair
    return (got.detach().float().cpu().reshape(-1)[::stride] - vals).abs().max().item()


def _inputs(gold):
    return torch.cat([FR.frame(gold["input_size"], s, "smooth") for s in gold["seeds"]])


def test_stage2_layout_spec_matches_reference_manifest():
    from emoportraits_b200.stage2 import Stage2Config, state_dict_spec_s2

    ref = {}
    for line in (GOLD / "state_dict_manifest_s2_512.txt").read_text().splitlines():
        k, shp = line.split(" ", 1)
        ref[k] = tuple(ast.literal_eval(shp))
    assert {k: tuple(v) for k, v in state_dict_spec_s2(Stage2Config(output_size=512)).items()} == ref


def test_stage2_args_txt_parses_to_default_config():
    from emoportraits_b200.config import parse_args
    from emoportraits_b200.stage2 import Stage2Config, stage2_config

    assert stage2_config(parse_args(GOLD / "args_s2_512.txt")) == Stage2Config(output_size=512)


def test_stage2_restatement_matches_reference_golden():
    """oracle/restatement.py stage-2 functions == unmodified reference infer_s2.InferenceWrapper.forward"""
    from emoportraits_b200.stage2 import Stage2Config, synthetic_state_dict_s2
    from oracle import restatement as R

    gold = torch.load(GOLD / "s2_512_b1.pt", weights_only=False)
    sd = synthetic_state_dict_s2(Stage2Config(output_size=512), 0)
    with torch.no_grad():
        resized, add, ffhq = R.stage2_forward(sd, _inputs(gold), 512)
        vol = R.stage2_local_encoder(sd, resized)
    assert _sub_err(vol, gold["vol"]) < 1e-4
    assert _sub_err(add, gold["add"]) < 1e-4
    # fixture layout is (N,H,W,3) uint8 (np.asarray of the PIL images); ToPILImage quantises by mul(255).byte()
    assert _sub_err((ffhq * 255).floor().clamp(0, 255).permute(0, 2, 3, 1).contiguous(), gold["ffhq_uint8"]) <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["s2_512_b1.pt", "s2_1024_b4.pt"])
def test_stage2_gpu_matches_reference_golden(fixture):
    from emoportraits_b200.stage2 import Stage2Config, Stage2Model, synthetic_state_dict_s2

    gold = torch.load(GOLD / fixture, weights_only=False)
    cfg = Stage2Config(output_size=gold["output_size"])
    model = Stage2Model(cfg, synthetic_state_dict_s2(cfg, 0), "cuda")
    img = _inputs(gold).cuda()
    x4 = __import__("emoportraits_b200.ops", fromlist=["ops"]).resize_bilinear(img, (cfg.output_size, cfg.output_size))
    vol = model.local_encoder(x4)
    e_vol = _sub_err(vol[:, 0].permute(0, 3, 1, 2).contiguous(), gold["vol"])
    resized, add, ffhq = model.forward(img)
    e_add = _sub_err(add, gold["add"])
    e_img = _sub_err((ffhq * 255).floor().clamp(0, 255).permute(0, 2, 3, 1).contiguous(), gold["ffhq_uint8"])
    print(f"\n[stage-2 parity vs reference golden {fixture}] latent {e_vol:.2e} add {e_add:.2e} ffhq(uint8 steps) {e_img:.0f}")
    out = pathlib.Path("gpurun_out"); out.mkdir(exist_ok=True)
    (out / f"parity_{fixture}.txt").write_text(f"latent {e_vol:.3e}\nadd {e_add:.3e}\nffhq_uint8 {e_img:.0f}\n")
    assert e_vol < 1e-3 and e_add < 1e-3 and e_img <= 1.0


@pytest.mark.gpu
def test_stage2_wrapper_api(tmp_path):
    from emoportraits_b200.stage2 import InferenceWrapper, Stage2Config, synthetic_state_dict_s2

    cfg = Stage2Config(output_size=512)
    d = tmp_path / "logs_s2" / "exp" / "checkpoints"
    d.mkdir(parents=True)
    (tmp_path / "logs_s2" / "exp" / "args.txt").write_text((GOLD / "args_s2_512.txt").read_text())
    torch.save(synthetic_state_dict_s2(cfg, 0), d / "m.pth")
    w = InferenceWrapper(experiment_name="exp", model_file_name="m.pth", project_dir=str(tmp_path))
    gold = torch.load(GOLD / "s2_512_b1.pt", weights_only=False)
    pil, pil_resized, pil_ffhq, mask = w.forward(_inputs(gold).cuda())
    assert len(pil) == 1 and pil_resized[0].size == (512, 512) and pil_ffhq[0].size == (512, 512) and mask.shape == (1, 1, 256, 256)
    import numpy as np
    got = torch.from_numpy(np.asarray(pil_ffhq[0]).astype("float32"))[None]
    assert _sub_err(got, gold["ffhq_uint8"]) <= 1.0
