"""GPU tests of everything that was written AFTER round 1's GPU budget was spent and has therefore never run on a B200.
They sort last (test_zz_*), carry xfail(strict=False) — an XPASS in the round-end log is the record of their first pass —
and a hard per-test timeout, so that a defect here cannot take the validated suite down with it.

  1. the non-default InferenceWrapper.forward arguments that reach the hot path (notebooks/infer.py:355-357: mix_old,
     mix=False, target_theta=False, smooth_pose, custome_target_pose_embed, custome_target_theta_embed, source_mask /
     driver_mask, c_source_latent_volume, c_target_latent_volume) against fixtures recorded from the UNMODIFIED reference
     (tests/golden/va256_options.pt, `python -m oracle.make_golden options`); the oracle restatement of the same options is
     pinned to the same fixtures on the CPU (tests/test_oracle_options.py) and the device pose algebra source is checked on
     the CPU by tests/test_pose_math_host.py;
  2. the sub-pixel up-sampling convolution (emo_conv_desc.upconv, EMO_UPCONV_PS=1);
  3. `conv -> avgpool` folded into one 4x4 stride-2 convolution (EMO_POOLCONV_FOLD=1);
  4. fp16 two-plane operands for the fp32-faithful networks (EMO_H2_NETS).
When they have passed once on the GPU, drop the xfail marker (and, for 2-4, flip the defaults if they are faster)."""
import pathlib

import pytest
import torch

from oracle import frames as FR

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first GPU run pending (round-1 GPU budget was spent before these were written)"),
              pytest.mark.timeout(600, method="thread")]   # never-run kernels: bound a hang instead of blocking the suite
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


# ------------------------------------------------------------------------------------------------------------------
# sub-pixel up-sampling convolution (emo_conv_desc.upconv; opt-in in the model via EMO_UPCONV_PS=1).  Same status as the
# tests above: written after the round-1 GPU budget was spent, first GPU run pending.  The weight folding and the
# kernel's index arithmetic are checked on the CPU by tests/test_upconv_fold.py.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Cin,Cout,S,residual", [(192, 128, 64, False), (320, 192, 32, True), (512, 320, 32, False),
                                                 (192, 128, 256, True)])
def test_upconv_subpixel_matches_conv_on_upsampled_planes(Cin, Cout, S, residual):
    """conv_igemm(low-res planes, folded weights, upconv=True) == conv_igemm(nearest-x2 planes, 3x3 weights): same operands
    up to the fp32 pre-summing of the folded taps, so agreement to ~2^-15 of the output scale; statistics likewise.
    S = 256 (-> 512^2, 2048 tiles) takes the store-warp (EPI = 1) variant of the kernel, the others EPI = 0."""
    import math

    from emoportraits_b200 import ops

    g = torch.Generator().manual_seed(Cin + S)
    x = torch.randn((1, 1, S, S, Cin), generator=g).cuda()
    w = torch.randn((Cout, Cin, 3, 3), generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, generator=g).cuda()
    res = torch.randn((1, 1, S, S, Cout), generator=g).cuda() if residual else None   # low-res skip, read with res_shift = 1
    ops.begin_pass("cuda")                      # zeroed statistics arena for new_stats()
    a_lo = ops.apply(x, act=ops.ACT_RELU, up=1)
    a_up = ops.apply(x, act=ops.ACT_RELU, up=2)
    st_ref, st_ps = ops.new_stats(1, 32, "cuda"), ops.new_stats(1, 32, "cuda")
    ref = ops.conv_igemm(a_up, ops.pack_conv_weight(w), bias=b, residual=res, res_shift=1 if residual else 0, stats=st_ref)
    out = ops.conv_igemm(a_lo, ops.pack_upconv_weight(w), bias=b, residual=res, res_shift=1 if residual else 0, stats=st_ps,
                         upconv=True)
    torch.cuda.synchronize()
    assert out.shape == ref.shape == (1, 1, 2 * S, 2 * S, Cout)
    scale = ref.abs().max().item()
    err = (out - ref).abs().max().item()
    print(f"\n[upconv sub-pixel vs upsampled planes] {Cin}->{Cout} @{S}^2: max-abs {err:.2e} (scale {scale:.2f})")
    assert err < scale * 2 ** -13
    assert ((st_ps - st_ref).abs() / st_ref.abs().clamp_min(1.0)).max().item() < 1e-5
    if S <= 64:   # and against torch fp32 on the CPU
        xr = torch.relu(x[0, 0].permute(2, 0, 1)[None].cpu())
        want = F_conv_up(xr, w, b.cpu(), res)
        assert (out[0, 0].permute(2, 0, 1)[None].cpu() - want).abs().max().item() < 2e-4 * max(1.0, scale)


def F_conv_up(xr, w, b, res):
    import torch.nn.functional as F

    y = F.conv2d(F.interpolate(xr.double(), scale_factor=2, mode="nearest"), w.double(), b.double(), padding=1)
    if res is not None:
        y = y + F.interpolate(res[0, 0].permute(2, 0, 1)[None].cpu().double(), scale_factor=2, mode="nearest")
    return y.float()


def test_model_with_subpixel_up_convolutions_matches_reference(ctx, monkeypatch):
    """whole driver pass with the three up-sampling convolutions of the image decoder in sub-pixel form vs the reference
    fixture (same 1e-3 bar as the default path)"""
    from emoportraits_b200 import nets
    from emoportraits_b200.infer import Model

    monkeypatch.setattr(nets, "SUBPIXEL_UP", True)
    model = Model(ctx["cfg"], ctx["sd"], ctx["hsd"], "cuda")
    assert any(b.c1_ps is not None for b in model.decoder_nw.img)
    st = model.source_pass(ctx["src"])
    img, _, _, so = model.driver_pass(st, ctx["drv"][0], mix=True)
    _check(ctx["gold"]["default"], img, so, "default+subpixel_up")
    base, _, _, _ = ctx["model"].driver_pass(ctx["st"], ctx["drv"][0], mix=True)
    print(f"[subpixel vs plain decoder] image max-abs {(img - base).abs().max().item():.2e}")


@pytest.mark.parametrize("fixture", ["s2_512_b1.pt", "s2_1024_b4.pt"])
def test_stage2_with_subpixel_up_convolutions_matches_reference(fixture, monkeypatch):
    """stage-2 refinement decoder (four nearest-x2 up blocks, batch 1 and 4) with the sub-pixel convolutions vs the
    reference fixtures; same bars as tests/test_stage2.py"""
    from emoportraits_b200 import nets
    from emoportraits_b200.stage2 import Stage2Config, Stage2Model, synthetic_state_dict_s2
    from test_stage2 import _inputs, _sub_err   # tests/ is on sys.path under pytest (rootdir conftest)

    monkeypatch.setattr(nets, "SUBPIXEL_UP", True)
    gold = torch.load(GOLD / fixture, weights_only=False)
    cfg = Stage2Config(output_size=gold["output_size"])
    model = Stage2Model(cfg, synthetic_state_dict_s2(cfg, 0), "cuda")
    assert sum(b.c1_ps is not None for b in model.up + model.feat) >= 3
    resized, add, ffhq = model.forward(_inputs(gold).cuda())
    e_add = _sub_err(add, gold["add"])
    e_img = _sub_err((ffhq * 255).floor().clamp(0, 255).permute(0, 2, 3, 1).contiguous(), gold["ffhq_uint8"])
    print(f"\n[stage-2 + sub-pixel up convs vs reference golden {fixture}] add {e_add:.2e} ffhq(uint8 steps) {e_img:.0f}")
    assert e_add < 1e-3 and e_img <= 1.0


# ------------------------------------------------------------------------------------------------------------------
# `3x3 conv -> 2x2 average pool` folded into one 4x4 stride-2 convolution (ops.fold_poolconv_weight; opt-in in the models
# via EMO_POOLCONV_FOLD=1).  No kernel change: the implicit-GEMM kernel is generic in the tap count.  First GPU run pending.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Cin,Cout,S,planes", [(128, 128, 64, 2), (256, 256, 128, 2), (128, 256, 32, 3)])
def test_poolconv_fold_matches_conv_then_avgpool(Cin, Cout, S, planes):
    import math

    import torch.nn.functional as F

    from emoportraits_b200 import ops

    g = torch.Generator().manual_seed(Cin + S + planes)
    x = torch.randn((1, Cin, S, S), generator=g)
    w = torch.randn((Cout, Cin, 3, 3), generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, generator=g)
    skip = torch.randn((1, Cout, S // 2, S // 2), generator=g)
    want = (F.avg_pool2d(F.conv2d(x.double(), w.double(), b.double(), padding=1), 2) + skip.double()).float()
    ops.begin_pass("cuda")
    a = ops.split_bf16(x.permute(0, 2, 3, 1)[:, None].contiguous().cuda(), planes)
    st = ops.new_stats(1, 32, "cuda")
    out = ops.conv_igemm(a, ops.pack_conv_weight(ops.fold_poolconv_weight(w), planes=planes), stride=(1, 2, 2), pad=(0, 1, 1),
                         bias=b.cuda(), residual=skip.permute(0, 2, 3, 1)[:, None].contiguous().cuda(), stats=st)
    torch.cuda.synchronize()
    got = out[:, 0].permute(0, 3, 1, 2).cpu()
    err = (got - want).abs().max().item()
    print(f"\n[4x4 stride-2 fold vs conv3x3 -> avgpool] {Cin}->{Cout} @{S}^2 planes {planes}: max-abs {err:.2e}")
    assert err < (2e-4 if planes == 2 else 2e-5) * max(1.0, want.abs().max().item())
    # statistics of the result, as the next GroupNorm needs them
    grp = want.view(1, 32, -1)
    ref_st = torch.stack([grp.double().sum(-1), (grp.double() ** 2).sum(-1)], -1)
    assert ((st.cpu() - ref_st).abs() / ref_st.abs().clamp_min(1.0)).max().item() < 1e-3


def test_models_with_poolconv_fold_match_reference(ctx, monkeypatch):
    """LocalEncoder (source pass) and the stage-2 encoder with the folded down-sampling convolutions vs the fixtures"""
    from emoportraits_b200 import nets
    from emoportraits_b200.infer import Model
    from emoportraits_b200.stage2 import Stage2Config, Stage2Model, synthetic_state_dict_s2
    from test_stage2 import _inputs, _sub_err

    monkeypatch.setattr(nets, "POOLCONV_FOLD", True)
    model = Model(ctx["cfg"], ctx["sd"], ctx["hsd"], "cuda")
    assert all(b.c2_pool is not None for b in model.local_encoder_nw.blocks)
    st = model.source_pass(ctx["src"])
    img, _, _, so = model.driver_pass(st, ctx["drv"][0], mix=True)
    _check(ctx["gold"]["default"], img, so, "default+poolconv_fold")
    gold = torch.load(GOLD / "s2_512_b1.pt", weights_only=False)
    cfg = Stage2Config(output_size=gold["output_size"])
    s2 = Stage2Model(cfg, synthetic_state_dict_s2(cfg, 0), "cuda")
    assert all(b.c2_pool is not None for b in s2.enc)
    resized, add, ffhq = s2.forward(_inputs(gold).cuda())
    e_add = _sub_err(add, gold["add"])
    e_img = _sub_err((ffhq * 255).floor().clamp(0, 255).permute(0, 2, 3, 1).contiguous(), gold["ffhq_uint8"])
    print(f"[stage-2 + folded encoder convs vs reference golden] add {e_add:.2e} ffhq(uint8 steps) {e_img:.0f}")
    assert e_add < 1e-3 and e_img <= 1.0


# ------------------------------------------------------------------------------------------------------------------
# fp16 two-plane operand mode ("h2", ops.H2; opt-in per network via EMO_H2_NETS / Model(precision=...)): three MMAs per
# product at fp32-level operand accuracy (tools/split_precision_emulation.py), meant to replace the six-MMA three-plane
# bf16 mode of the embedding / warp / source networks.  First GPU run pending.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Cin,Cout,sp,k", [(512, 512, (64, 64), 3), (96, 96, (16, 16), 3), (64, 32, (8, 16, 16), 3), (64, 128, (8, 8), 3),
                                           (256, 128, (8, 8, 8), 1)])
def test_conv_igemm_fp16_two_planes(Cin, Cout, sp, k):
    """same cases and bar as tests/test_ops_gpu.py::test_conv_igemm_three_planes: fp32-faithful against an fp64 reference"""
    import math

    import torch.nn.functional as F

    from emoportraits_b200 import ops

    g = torch.Generator().manual_seed(Cin + Cout)
    three_d = len(sp) == 3
    x = torch.randn(1, Cin, *sp, generator=g)
    w = torch.randn(Cout, Cin, *([k] * len(sp)), generator=g) / math.sqrt(Cin * k ** len(sp))
    b = torch.randn(Cout, generator=g)
    ref = (F.conv3d if three_d else F.conv2d)(x.double(), w.double(), b.double(), padding=k // 2)
    xc = x.permute(0, 2, 3, 4, 1).contiguous() if three_d else x.permute(0, 2, 3, 1)[:, None].contiguous()
    ops.begin_pass("cuda")
    a = ops.split_bf16(xc.cuda(), ops.H2)
    assert a.f16 and a.hi.dtype == torch.float16
    assert (a.float().cpu() - xc).abs().max().item() < 1e-5                # the planes reproduce the activations
    out = ops.conv_igemm(a, ops.pack_conv_weight(w, planes=ops.H2), bias=b.cuda())
    torch.cuda.synchronize()
    got = out.permute(0, 4, 1, 2, 3).cpu() if three_d else out[:, 0].permute(0, 3, 1, 2).cpu()
    err = (got.double() - ref).abs().max().item()
    print(f"\n[conv fp16 two planes vs fp64] {Cin}->{Cout} {sp} k{k}: max-abs {err:.2e} (scale {ref.abs().max().item():.2f})")
    assert err < 2e-5 * max(1.0, ref.abs().max().item())
    # GroupNorm + ReLU planes written by the apply pass in the same format
    st = ops.gn_stats(out, 32)
    gamma, beta = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)
    sp_h = ops.apply(out, gn=dict(stats=st, count=out.numel() / 32, gamma=gamma.cuda(), beta=beta.cuda()), act=ops.ACT_RELU, planes=ops.H2)
    want = torch.relu(F.group_norm(got, 32, gamma, beta, 1e-5))
    back = sp_h.float().permute(0, 4, 1, 2, 3).cpu() if three_d else sp_h.float()[:, 0].permute(0, 3, 1, 2).cpu()
    assert (back - want).abs().max().item() < 1e-4


def test_model_with_fp16_two_plane_networks_matches_reference(ctx):
    """the per-frame networks that run with three bf16 planes today (head pose, expression, warp generators) switched to
    fp16 two planes: same 1e-3 image bar against the reference fixture, and stage outputs close to the default path"""
    from emoportraits_b200 import ops
    from emoportraits_b200.infer import Model

    model = Model(ctx["cfg"], ctx["sd"], ctx["hsd"], "cuda",
                  precision=dict(head_pose=ops.H2, expression=ops.H2, warp=ops.H2, idt=ops.H2, local_encoder=ops.H2,
                                 volume_source=ops.H2, unet3d=ops.H2))
    st = model.source_pass(ctx["src"])
    img, _, _, so = model.driver_pass(st, ctx["drv"][0], mix=True)
    _check(ctx["gold"]["default"], img, so, "default+h2 networks")
    base, _, _, so0 = ctx["model"].driver_pass(ctx["st"], ctx["drv"][0], mix=True)
    print(f"[h2 vs three-plane networks] image max-abs {(img - base).abs().max().item():.2e} "
          f"pose_embed {(so.target_pose_embed - so0.target_pose_embed).abs().max().item():.2e}")


# ------------------------------------------------------------------------------------------------------------------
# grid_sample_3d at BASELINE's largest size and its edge cases (validated kernels, new tests)
# ------------------------------------------------------------------------------------------------------------------
def test_grid_sample3d_full_size_batch32_properties():
    """BASELINE configs[2] at its largest size (96ch x 64^3 volume, 64^3 lattice, batch 32: 25.8 GB in, 25.8 GB out), where a
    CPU reference would take minutes: size-independent properties instead, all bit-exact.
      * batch independence: sample n of the batched call == the single-sample call on (x[n], theta[n] / grid[n]);
      * linearity in the volume for a power-of-two factor: gs(0.5 x) == 0.5 gs(x);
      * zeros padding: a lattice entirely outside the volume samples exactly 0;
      * channel equivariance: permuting the channels of the volume permutes the channels of the result."""
    from emoportraits_b200 import ops
    from test_ops_gpu import _grid

    N, C, S = 32, 96, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((N, S, S, S, C), generator=g, device="cuda")
    ang = torch.linspace(-0.6, 0.6, N)
    th = torch.zeros(N, 3, 4)
    th[:, 0, 0], th[:, 0, 1], th[:, 1, 0], th[:, 1, 1], th[:, 2, 2] = ang.cos(), -ang.sin(), ang.sin(), ang.cos(), 0.9
    th[:, :, 3] = torch.tensor([0.2, -0.1, 0.05])
    th = th.cuda().contiguous()
    out = ops.grid_sample3d(x, theta=th, out_size=(S, S, S), in_layout="cl")
    assert out.shape == (N, S, S, S, C)
    for n in (0, 17, 31):
        one = ops.grid_sample3d(x[n:n + 1].contiguous(), theta=th[n:n + 1].contiguous(), out_size=(S, S, S), in_layout="cl")
        assert torch.equal(one[0], out[n]), n
    assert (out[31] != 0).float().mean().item() > 0.5            # a rotated, shifted lattice still lands mostly inside
    half = ops.grid_sample3d((x[:2] * 0.5).contiguous(), theta=th[:2].contiguous(), out_size=(S, S, S), in_layout="cl")
    assert torch.equal(half, out[:2] * 0.5)
    perm = torch.randperm(C, generator=torch.Generator().manual_seed(6)).cuda()
    pc = ops.grid_sample3d(x[:1, ..., perm].contiguous(), theta=th[:1].contiguous(), out_size=(S, S, S), in_layout="cl")
    assert torch.equal(pc, out[:1, ..., perm])
    far = th[:1].clone()
    far[:, :, 3] = 4.0
    assert ops.grid_sample3d(x[:1].contiguous(), theta=far.contiguous(), out_size=(S, S, S), in_layout="cl").abs().max().item() == 0.0
    del out, half, pc
    # explicit grid tensor (jittered identity lattice), batch 32
    grid = (_grid(1, S, S, S, 9).cuda() + 0.02 * torch.arange(N, device="cuda").view(N, 1, 1, 1, 1)).contiguous()
    outg = ops.grid_sample3d(x, grid=grid, in_layout="cl")
    for n in (0, 31):
        one = ops.grid_sample3d(x[n:n + 1].contiguous(), grid=grid[n:n + 1].contiguous(), in_layout="cl")
        assert torch.equal(one[0], outg[n]), n


def test_grid_sample3d_empty_and_invalid_inputs():
    """edge cases: an empty batch returns an empty tensor (as F.grid_sample does) without launching; malformed arguments
    are refused by the C-ABI with a message (EMO_ERR_INVALID -> RuntimeError), never launched."""
    from emoportraits_b200 import ops

    e = ops.grid_sample3d(torch.empty((0, 4, 4, 4, 8), device="cuda"), theta=torch.empty((0, 3, 4), device="cuda"),
                          out_size=(4, 4, 4), in_layout="cl")
    assert e.shape == (0, 4, 4, 4, 8)
    x = torch.randn((1, 4, 4, 4, 6), device="cuda")          # channels-last path needs C % 4 == 0
    with pytest.raises(RuntimeError, match="C % 4"):
        ops.grid_sample3d(x, theta=torch.eye(4, device="cuda")[None, :3].contiguous(), out_size=(4, 4, 4), in_layout="cl")


