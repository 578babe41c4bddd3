"""Stage-isolated GPU parity: every network of the path is fed the ORACLE's input for that stage and compared with the
oracle's output, so an error cannot hide behind (or be blamed on) an upstream stage.  CPU oracle run live, at 256^2 and at
512^2 (the shipped size: third LocalEncoder block, 512^2 decoder level)."""
import pathlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def frame(size, seed):
    from oracle import frames as FR
    return FR.frame(size, seed, "noise")


def cl(x):
    if x.dim() == 4:
        x = x[:, :, None]
    return x.permute(0, 2, 3, 4, 1).contiguous()


def uncl(x):
    return x.permute(0, 4, 1, 2, 3).contiguous()


@pytest.fixture(scope="module", params=[256, 512])
def ctx(request):
    from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
    from emoportraits_b200.config import shipped_config
    from emoportraits_b200.infer import Model
    from oracle import restatement as R

    size = request.param
    cfg = shipped_config(size)
    sd, hsd = synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0)
    model = Model(cfg, sd, hsd, "cuda")
    ocfg = R.config_from_state_dict(sd, size)
    src, drv = frame(size, 21), frame(size, 22)
    with torch.no_grad():
        staps, dtaps = {}, {}
        ost = R.source_pass(sd, hsd, src, ocfg, staps)
        oimg = R.driver_pass(sd, hsd, ost, drv, ocfg, dtaps)
    return dict(size=size, cfg=cfg, sd=sd, hsd=hsd, model=model, ocfg=ocfg, src=src, drv=drv, ost=ost, staps=staps,
                dtaps=dtaps, oimg=oimg, R=R, report={})


def _rec(ctx, name, got, ref, tol):
    err = (got.detach().float().cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    ctx["report"][name] = (err, scale)
    out = pathlib.Path("gpurun_out"); out.mkdir(exist_ok=True)
    with open(out / "stage_parity.txt", "a") as f:
        f.write(f"{ctx['size']} {name} err={err:.3e} scale={scale:.3e} rel={err / max(scale, 1e-30):.3e}\n")
    print(f"\n[stage @{ctx['size']}] {name}: max-abs err {err:.3e} (ref max {scale:.3e})")
    assert err < tol * max(scale, 1.0), (name, err, scale)


def test_head_pose(ctx):
    srt = ctx["model"].head_pose_regressor(ctx["drv"].cuda())
    _rec(ctx, "head_pose.srt", srt, ctx["dtaps"]["srt"], 2e-5)


def test_expression_embed(ctx):
    from emoportraits_b200 import ops
    R = ctx["R"]
    th = ctx["dtaps"]["theta"]
    align = R.align_theta_2d(th).contiguous().cuda()
    emb, aligned = ctx["model"].expression_embedder_nw(ctx["drv"].cuda(), align, want_aligned=True)
    _rec(ctx, "expression.aligned_face", aligned, ctx["dtaps"]["aligned_face"], 1e-5)
    _rec(ctx, "expression.pose_embed", emb, ctx["dtaps"]["pose_embed"], 2e-5)


def test_idt_embed(ctx):
    e = ctx["model"].idt_embedder_nw(ctx["src"].cuda())
    _rec(ctx, "idt_embed", e, ctx["ost"]["idt_embed"], 2e-5)


def test_predict_embed_and_warp_generator(ctx):
    m = ctx["model"]
    E = m.predict_embed(ctx["dtaps"]["pose_embed"].cuda().contiguous(), ctx["ost"]["idt_embed"].cuda().contiguous())
    _rec(ctx, "predict_embed.orig", E, ctx["dtaps"]["embed"], 1e-5)
    warp = m.uv_generator_nw(ctx["dtaps"]["embed"].cuda().contiguous())
    _rec(ctx, "uv_generator.warp", warp, ctx["dtaps"]["uv_warp"], 2e-5)


def test_warps(ctx):
    from emoportraits_b200 import ops
    cfg = ctx["cfg"]
    vol = cl(ctx["ost"]["target_latent_volume"]).cuda()
    v = ops.grid_sample3d(vol, grid=ctx["dtaps"]["uv_warp"].cuda().contiguous(), in_layout="cl")
    out = ops.grid_sample3d(v, theta=ctx["dtaps"]["theta"].cuda().contiguous(), out_size=(cfg.D, cfg.S, cfg.S), in_layout="cl")
    _rec(ctx, "grid_sample_pair", uncl(out.cpu()), ctx["dtaps"]["aligned_volume"], 2e-5)


def test_decoder(ctx):
    from emoportraits_b200 import ops
    cfg = ctx["cfg"]
    av = ctx["dtaps"]["aligned_volume"]  # (1,C,D,S,S) -> (h,w,d,c)
    hwdc = av[0].permute(2, 3, 1, 0).contiguous().view(1, 1, cfg.S, cfg.S, cfg.D * cfg.C).cuda()
    logits, feat, _ = ctx["model"].decoder_nw(ops.split_bf16(hwdc), want_logits=True)
    _rec(ctx, "decoder.feat", uncl(feat.cpu())[:, :, 0], ctx["dtaps"]["dec_feat"], 1e-4)
    _rec(ctx, "decoder.logits", logits, ctx["dtaps"]["logits"], 1e-3)


def test_local_encoder_and_volume_nets(ctx):
    cfg, m, R = ctx["cfg"], ctx["model"], ctx["R"]
    vol = m.local_encoder_nw(ctx["src"].cuda())
    ref = ctx["staps"]["latents"].view(1, cfg.C, cfg.D, cfg.S, cfg.S)
    _rec(ctx, "local_encoder.latents", uncl(vol.cpu()), ref, 5e-5)
    v2 = m.volume_source_nw(cl(ref).cuda())
    _rec(ctx, "volume_source", uncl(v2.cpu()), ctx["staps"]["vol_source"], 5e-5)
    u = m.volume_process_nw(cl(ctx["staps"]["vol_warped"]).cuda())
    _rec(ctx, "unet3d", uncl(u.cpu()), ctx["ost"]["target_latent_volume"], 1e-4)
