"""GPU parity of the non-default InferenceWrapper.forward arguments that reach the hot path (notebooks/infer.py:355-357:
mix_old, mix=False, target_theta=False, smooth_pose, custome_target_pose_embed, custome_target_theta_embed, source_mask /
driver_mask, c_source_latent_volume, c_target_latent_volume) against fixtures recorded from the UNMODIFIED reference
(tests/golden/va256_options.pt, `python -m oracle.make_golden options`).  The oracle restatement of the same options is
pinned to the same fixtures on the CPU (tests/test_oracle_options.py) and the device pose algebra source is checked on the
CPU by tests/test_pose_math_host.py."""
import pathlib

import pytest
import torch

from oracle import frames as FR

pytestmark = pytest.mark.gpu
GOLD = pathlib.Path(__file__).parent / "golden"
SIZE = 256
IMG_TOL = 1e-3      # BASELINE.json north_star: max-abs per pixel on the fp32 image
THETA_TOL = 1e-5


def _img_err(got, ref):
    vals, stride = ref
    return (got.detach().float().cpu().reshape(-1)[::stride] - vals).abs().max().item()


@pytest.fixture(scope="module")
def ctx():
    from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
    from emoportraits_b200.config import shipped_config
    from emoportraits_b200.infer import Model
    from oracle.make_golden import option_inputs

    cfg = shipped_config(SIZE)
    sd, hsd = synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0)
    model = Model(cfg, sd, hsd, "cuda")
    gold = torch.load(GOLD / f"va{SIZE}_options.pt", weights_only=False)
    src = FR.frame(SIZE, gold["src_seed"], gold["kind"]).cuda()
    drv = [FR.frame(SIZE, s, gold["kind"]).cuda() for s in gold["drv_seeds"]]
    return dict(cfg=cfg, sd=sd, hsd=hsd, model=model, gold=gold["cases"], src=src, drv=drv, X=option_inputs(SIZE, cfg),
                st=model.source_pass(src))


def _check(case, img, so, name):
    e_img = _img_err(img, case["img"])
    e_th = (so.pred_target_theta[:, :3].cpu() - case["pred_target_theta"][:, :3]).abs().max().item()
    e_pe = (so.target_pose_embed.cpu() - case["target_pose_embed"]).abs().max().item()
    print(f"\n[options parity vs reference @256] {name}: img {e_img:.2e} theta {e_th:.2e} pose_embed {e_pe:.2e}")
    assert e_th < THETA_TOL, (name, e_th)
    assert e_pe < 1e-4, (name, e_pe)
    assert e_img < IMG_TOL, (name, e_img)


@pytest.mark.parametrize("name,kw", [
    ("default", {}),
    ("mix_old", dict(mix_old=True)),
    ("no_mix", dict(mix=False)),
    ("target_theta_false", dict(target_theta=False)),
])
def test_pose_options(ctx, name, kw):
    kw = dict(dict(mix=True), **kw)
    img, _, _, so = ctx["model"].driver_pass(ctx["st"], ctx["drv"][0], **kw)
    _check(ctx["gold"][name], img, so, name)


def test_smooth_pose_state_carried_over_frames(ctx):
    state = torch.zeros((3, 4), device="cuda")
    for i, d in enumerate(ctx["drv"]):
        img, _, _, so = ctx["model"].driver_pass(ctx["st"], d, mix=True, smooth_state=state, smooth_momentum=0.5, smooth_init=(i == 0))
        _check(ctx["gold"][f"smooth_pose_{i}"], img, so, f"smooth_pose_{i}")
        assert torch.equal(state, so.pred_target_theta[0, :3])


def test_custom_embeddings(ctx):
    X = ctx["X"]
    img, _, _, so = ctx["model"].driver_pass(ctx["st"], ctx["drv"][0], mix=True, custom_pose_embed=X["pose_embed"])
    _check(ctx["gold"]["custome_target_pose_embed"], img, so, "custome_target_pose_embed")
    img, _, _, so = ctx["model"].driver_pass(ctx["st"], ctx["drv"][0], mix=True, custom_srt=torch.cat(X["theta_embed"], 1))
    _check(ctx["gold"]["custome_target_theta_embed"], img, so, "custome_target_theta_embed")


def test_source_mask_and_custom_volumes(ctx):
    m, X = ctx["model"], ctx["X"]
    g = ctx["gold"]["source_mask"]
    st = m.source_pass(ctx["src"], mask=X["source_mask"].cuda())
    assert (st.idt_embed.cpu() - g["idt_embed"]).abs().max().item() < 1e-4
    assert (st.pred_source_theta.cpu() - g["pred_source_theta"]).abs().max().item() < THETA_TOL   # regressor sees the unmasked image
    img, _, _, so = m.driver_pass(st, ctx["drv"][0], mix=True)
    _check(g, img, so, "source_mask")
    for key in ("c_source_latent_volume", "c_target_latent_volume"):
        st = m.source_pass(ctx["src"], **{key: X[key]})
        img, _, _, so = m.driver_pass(st, ctx["drv"][0], mix=True)
        _check(ctx["gold"][key], img, so, key)


def test_wrapper_forward_options(ctx, tmp_path):
    """the same options through the drop-in InferenceWrapper.forward (PIL in, (list[PIL], tensor) out)"""
    from emoportraits_b200.infer import InferenceWrapper

    gold, X = ctx["gold"], ctx["X"]
    exp = tmp_path / "logs" / "exp" / "checkpoints"
    exp.mkdir(parents=True)
    (tmp_path / "logs" / "exp" / "args.txt").write_text((GOLD / f"args_{SIZE}.txt").read_text())
    torch.save(ctx["sd"], exp / "000_model.pth")
    w = InferenceWrapper(experiment_name="exp", model_file_name="000_model.pth", project_dir=str(tmp_path), folder="logs",
                         print_params=False, head_pose_state_dict=ctx["hsd"])
    full = torch.load(GOLD / f"va{SIZE}_options.pt", weights_only=False)
    src = FR.pil(SIZE, full["src_seed"], full["kind"])
    drv = [FR.pil(SIZE, s, full["kind"]) for s in full["drv_seeds"]]
    base = dict(crop=False, mix=True, mix_old=False)
    _, img = w.forward(src, drv[0], crop=False, mix=True, mix_old=True)
    assert _img_err(img, gold["mix_old"]["img"]) < IMG_TOL
    w.forward(src, None, **base)
    for i, d in enumerate(drv):
        pil, img = w.forward(None, d, smooth_pose=True, reset_tracking=(i == 0), **base)
        assert _img_err(img, gold[f"smooth_pose_{i}"]["img"]) < IMG_TOL, i
        assert (w.pred_target_theta[:, :3].cpu() - gold[f"smooth_pose_{i}"]["pred_target_theta"][:, :3]).abs().max().item() < THETA_TOL
    assert w.theta.shape == (3, 4)
    _, img = w.forward(src, drv[0], custome_target_theta_embed=X["theta_embed"], **base)
    assert _img_err(img, gold["custome_target_theta_embed"]["img"]) < IMG_TOL
    _, img = w.forward(src, drv[0], custome_target_pose_embed=X["pose_embed"], **base)
    assert _img_err(img, gold["custome_target_pose_embed"]["img"]) < IMG_TOL
    _, img = w.forward(src, drv[0], source_mask=X["source_mask"], driver_mask=X["driver_mask"], **base)
    assert _img_err(img, gold["source_mask"]["img"]) < IMG_TOL
    assert w.source_img_mask.shape == (1, 1, SIZE, SIZE)
    _, img = w.forward(src, drv[0], c_target_latent_volume=X["c_target_latent_volume"], **base)
    assert _img_err(img, gold["c_target_latent_volume"]["img"]) < IMG_TOL


