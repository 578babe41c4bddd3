"""Checkpoint layout contract of the reference (SURVEY.md §5 "Checkpoint / resume") and weight pre-folding.

  * `state_dict_spec(cfg)` — every key/shape `models/stage_1/volumetric_avatar/va.py` Model.state_dict() holds for the
    shipped configuration, derived structurally (so a unit test can diff it against the committed manifest that
    oracle/make_golden.py extracted from the real reference).
  * `fold_conv` — static weight transforms applied once at load instead of every forward:
        spectral norm (eval): W = W_orig / (u . (W_mat v))            utils/spectral_norm.py:131-168,180
        weight standardisation: (w - mean) / (std_unbiased + 1e-5)     networks/volumetric_avatar/utils.py:894-915
  * `synthetic_state_dict` — seeded, well-conditioned random weights of that exact layout (no checkpoint is
    downloadable; the reference's own random init is numerically chaotic: sigma = u.W.v of random u, v is ~0, so
    activations explode and two fp32 evaluations of the same net differ by 1e-2 in the image).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

from .config import HotPathConfig


# ------------------------------------------------------------------------------------------------------------------
# structural spec
# ------------------------------------------------------------------------------------------------------------------
def _sn(spec, p, co, ci, k, bias=False):
    if bias:
        spec[p + ".bias"] = (co,)
    spec[p + ".weight_orig"] = (co, ci) + tuple(k)
    spec[p + ".weight_u"] = (co,)
    spec[p + ".weight_v"] = (ci * math.prod(k),)


def _ws(spec, p, co, ci, k):
    spec[p + ".weight"] = (co, ci) + tuple(k)
    spec[p + ".bias"] = (co,)


def _gn(spec, p, c):
    spec[p + ".weight"] = (c,)
    spec[p + ".bias"] = (c,)


def _resblock(spec, p, cin, cout, dims, ws_first, skip):
    """utils.py:661-788 after apply_sp_to_nets + apply_ws_to_nets (positional WS rule utils.py:1071-1090)."""
    k3, k1 = (3,) * dims, (1,) * dims
    _gn(spec, p + ".block_feats.0", cin)
    if ws_first:
        _ws(spec, p + ".block_feats.2", cout, cin, k3)
    else:
        _sn(spec, p + ".block_feats.2", cout, cin, k3)
    _gn(spec, p + ".block_feats.3", cout)
    _sn(spec, p + ".block.0", cout, cout, k3)
    if skip:
        _sn(spec, p + ".skip.0", cout, cin, k1)


def _resnet(spec, p, layers, bottleneck, gn, fc_out, fc_conv):
    """torchvision resnet (+ replace_bn_to_gn, SN, WS as the reference applies them) / plain BN resnet."""
    def conv(q, co, ci, k, ws=False):
        if not gn:
            spec[q + ".weight"] = (co, ci, k, k)
        elif ws:
            _ws(spec, q, co, ci, (k, k))
        else:
            _sn(spec, q, co, ci, (k, k))

    def norm(q, c):
        _gn(spec, q, c)
        if not gn:
            spec[q + ".running_mean"] = (c,)
            spec[q + ".running_var"] = (c,)
            spec[q + ".num_batches_tracked"] = ()

    conv(p + ".conv1", 64, 3, 7)
    norm(p + ".bn1", 64)
    inpl = 64
    exp = 4 if bottleneck else 1
    for li, (planes, nblk) in enumerate(zip([64, 128, 256, 512], layers), 1):
        for bi in range(nblk):
            q = f"{p}.layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            if bottleneck:
                conv(q + ".conv1", planes, inpl, 1); norm(q + ".bn1", planes)
                conv(q + ".conv2", planes, planes, 3, ws=True); norm(q + ".bn2", planes)
                conv(q + ".conv3", planes * 4, planes, 1, ws=True); norm(q + ".bn3", planes * 4)
            else:
                conv(q + ".conv1", planes, inpl, 3); norm(q + ".bn1", planes)
                conv(q + ".conv2", planes, planes, 3, ws=True); norm(q + ".bn2", planes)
            if stride != 1 or inpl != planes * exp:
                conv(q + ".downsample.0", planes * exp, inpl, 1)
                norm(q + ".downsample.1", planes * exp)
            inpl = planes * exp
    if fc_conv:
        _sn(spec, p + ".fc", fc_out, 512 * exp, (1, 1))
    else:
        spec[p + ".fc.weight"] = (fc_out, 512 * exp)
        spec[p + ".fc.bias"] = (fc_out,)


def _warp_generator(spec, p, cfg: HotPathConfig):
    ch = cfg.warp_channels
    spec[p + ".identity_grid"] = (1, 3, cfg.D, cfg.S, cfg.S)
    _sn(spec, p + ".first_conv", ch[0] * cfg.embed_size, cfg.gen_max_channels, (1, 1))
    for i in range(len(ch) - 1):
        _resblock(spec, f"{p}.blocks_3d.{i}", ch[i], ch[i + 1], 3, ws_first=True, skip=True)
    _gn(spec, p + ".pre_head.0", ch[-1])
    _sn(spec, p + ".head.0.0", 3, ch[-1], (3, 3, 3), bias=True)
    j = 0
    for i in range(len(ch) - 1):
        for c in (ch[i], ch[i + 1]):
            spec[f"{p}.projector.u.{j}"] = (c, cfg.gen_max_channels)
            spec[f"{p}.projector.v.{j}"] = (cfg.embed_size ** 2, 2)
            j += 1


def state_dict_spec(cfg: HotPathConfig) -> "OrderedDict[str, tuple]":
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    # --- local encoder (local_encoder.py:48-111)
    p = "local_encoder_nw"
    ec = cfg.enc_channels
    s = cfg.image_size
    _sn(spec, f"{p}.from_rgb_{s}px", ec[0], 3, (7, 7), bias=True)
    for i in range(len(ec) - 1):
        _resblock(spec, f"{p}.enc_{i}_block={s}px", ec[i], ec[i + 1], 2, ws_first=True, skip=True if ec[i] != ec[i + 1] else False)
        s //= 2
    _gn(spec, p + ".finale_layers.0", ec[-1])
    _ws(spec, p + ".finale_layers.2", cfg.C * cfg.D, ec[-1], (1, 1))
    # --- identity embedder (identity_embedder.py:12-57): resnet50, GN, fc 1x1 conv
    spec["idt_embedder_nw.mean"] = (1, 3, 1, 1)
    spec["idt_embedder_nw.std"] = (1, 3, 1, 1)
    _resnet(spec, "idt_embedder_nw.net", [3, 4, 6, 3], True, True, cfg.idt_channels, True)
    # --- expression embedder (expression_embedder.py:58-131, 343-440): resnet18, GN
    q = "expression_embedder_nw"
    g = cfg.exp_image_size // 2
    spec[q + ".identity_grid"] = (1, g * g, 3)
    spec[q + ".identity_grid_512"] = (1, 512 * 512, 3)
    spec[q + ".aligned_keypoints"] = (1, 68, 3)
    spec[q + ".net_face.mean"] = (1, 3, 1, 1)
    spec[q + ".net_face.std"] = (1, 3, 1, 1)
    _resnet(spec, q + ".net_face.net", [2, 2, 2, 2], False, True, cfg.expr_channels, True)
    _sn(spec, q + ".net_face.pose_head", cfg.expr_channels, cfg.expr_channels * 16, ())
    # --- embed heads (va.py:172-181)
    spec["pose_unsqueeze_nw.weight"] = (cfg.gen_max_channels * cfg.embed_size ** 2, cfg.expr_channels)
    _sn(spec, "warp_embed_head_orig_nw", cfg.gen_max_channels, cfg.gen_max_channels, (1, 1))
    # --- warp generators (warp_generator_resnet.py:38-123)
    _warp_generator(spec, "xy_generator_nw", cfg)
    _warp_generator(spec, "uv_generator_nw", cfg)
    # --- volume_source_nw (vpn_resblocks.py / resblocks_3d.py)
    for i in range(cfg.source_volume_num_blocks):
        _resblock(spec, f"volume_source_nw.net.net.{i}", cfg.C, cfg.C, 3, ws_first=False, skip=False)
    # --- Unet3D (unet_3d.py:44-194)
    p = "volume_process_nw"
    uc = cfg.unet_channels
    nb = len(uc) - 1
    spec[p + ".input_tensor"] = (1, uc[-1], cfg.dummy_input_size, cfg.dummy_input_size, cfg.dummy_input_size)
    for i in range(nb):
        _resblock(spec, f"{p}.blocks_3d_down.{i}", uc[i], uc[i + 1], 3, ws_first=False, skip=uc[i] != uc[i + 1])
    for j, i in enumerate(range(nb - 1, -1, -1)):
        cin, cout = uc[i + 1], uc[i]
        _resblock(spec, f"{p}.blocks_3d_up.{j}", cin, cout, 3, ws_first=False, skip=cin != cout)
        _resblock(spec, f"{p}.skip_blocks_3d_up.{j}", cin, cin, 3, ws_first=False, skip=False)
    _gn(spec, p + ".head.0", uc[0])
    _sn(spec, p + ".head.2", uc[0], uc[0], (1, 1, 1), bias=True)
    # --- decoder (decoder.py:52-150, 241-358)
    p = "decoder_nw"
    dc = cfg.dec_channels
    _sn(spec, p + ".res_decoder.0", dc[0], cfg.C * cfg.D, (1, 1))
    for i in range(cfg.dec_num_blocks):
        _resblock(spec, f"{p}.res_decoder.{i + 1}", dc[0], dc[0], 2, ws_first=True, skip=False)
    j = 0
    for r in range(len(dc) - 1):
        cin = dc[r]
        for k in range(cfg.im_dec_lrs):
            _resblock(spec, f"{p}.img_decoder.dec_img_blocks.{j}", cin, dc[r + 1], 2, ws_first=True, skip=cin != dc[r + 1])
            cin = dc[r + 1]
            j += 1
    _gn(spec, p + ".img_decoder.dec_img_head.0", dc[-1])
    _ws(spec, p + ".img_decoder.dec_img_head.2", 3, dc[-1], (1, 1))
    # --- Model.bn (va.py:67, unused at inference)
    spec["bn.running_mean"] = (512,)
    spec["bn.running_var"] = (512,)
    spec["bn.num_batches_tracked"] = ()
    return spec


def head_pose_spec() -> "OrderedDict[str, tuple]":
    """torchvision resnet18(num_classes=9) with BatchNorm: the separate head_pose_regressor.pth
    (networks/volumetric_avatar/head_pose_regressor.py:14-15)."""
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    _resnet(spec, "X", [2, 2, 2, 2], False, False, 9, False)
    return OrderedDict((k[2:], v) for k, v in spec.items())


# ------------------------------------------------------------------------------------------------------------------
# synthetic, well-conditioned weights
# ------------------------------------------------------------------------------------------------------------------
def _gen(seed, key):
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))


SANE_POSE_BIAS = [1, 1, 1, .15, -.1, .05, .03, -.02, .01]


def _fill(spec, seed, small_gamma=()):
    sd = OrderedDict()
    for k, shp in spec.items():
        g = _gen(seed, k)
        leaf = k.rsplit(".", 1)[-1]
        if leaf in ("weight_orig",) or (leaf == "weight" and len(shp) >= 2):
            fan_in = math.prod(shp[1:])
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        elif leaf == "weight":  # norm gamma
            scale = 1.0
            for pat, s in small_gamma:
                if pat in k:
                    scale = s
            sd[k] = (1.0 + 0.1 * torch.randn(shp, generator=g)) * scale
        elif leaf == "bias":
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        elif leaf == "running_mean":
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        elif leaf == "running_var":
            sd[k] = 1.0 + 0.1 * torch.rand(shp, generator=g)
        elif leaf == "num_batches_tracked":
            sd[k] = torch.tensor(0, dtype=torch.long)
        elif leaf in ("weight_u", "weight_v"):
            sd[k] = torch.zeros(shp)  # filled below by power iteration
        elif leaf == "input_tensor":
            sd[k] = torch.randn(shp, generator=g)
        elif leaf == "mean":
            sd[k] = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
        elif leaf == "std":
            sd[k] = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
        elif ".projector.u." in k:
            a = math.sqrt(3 / shp[1])
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * a
        elif ".projector.v." in k:
            a = math.sqrt(3 / shp[0])
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * a
        elif leaf == "identity_grid" and len(shp) == 5:
            gs, gz = torch.linspace(-1, 1, shp[3]), torch.linspace(-1, 1, shp[2])
            w, v, u = torch.meshgrid(gz, gs, gs, indexing="ij")
            sd[k] = torch.stack([u, v, w], 0)[None]
        elif leaf in ("identity_grid", "identity_grid_512"):
            n = int(round(math.sqrt(shp[1])))
            gr = torch.linspace(-1, 1, n)
            v, u = torch.meshgrid(gr, gr, indexing="ij")
            sd[k] = torch.stack([u, v, torch.ones_like(u)], dim=2).view(1, -1, 3)
        elif leaf == "aligned_keypoints":
            sd[k] = torch.zeros(shp)
        else:
            raise KeyError(f"no synthetic rule for {k}")
    # spectral-norm vectors: 3 power iterations so sigma ~ ||W||_2 (what training converges to)
    for k in list(sd.keys()):
        if k.endswith(".weight_orig"):
            p = k[: -len(".weight_orig")]
            W = sd[k].reshape(sd[k].shape[0], -1)
            u = torch.randn(W.shape[0], generator=_gen(seed, p + ".u"))
            u = u / u.norm()
            for _ in range(3):
                v = torch.mv(W.t(), u); v = v / (v.norm() + 1e-12)
                u = torch.mv(W, v); u = u / (u.norm() + 1e-12)
            sd[p + ".weight_u"], sd[p + ".weight_v"] = u, v
    return sd


def synthetic_state_dict(cfg: HotPathConfig, seed: int = 0):
    # WS convs have unit-variance weights => gain sqrt(fan_in); the norm gamma in front of the two WS heads is made
    # small so that the latent volume is O(1) and the image logits do not saturate the sigmoid.
    return _fill(state_dict_spec(cfg), seed, small_gamma=(("finale_layers.0", 0.05), ("dec_img_head.0", 0.1)))


def synthetic_head_pose_state_dict(seed: int = 0):
    sd = _fill(head_pose_spec(), seed + 77)
    sd["fc.weight"] = sd["fc.weight"] * 0.01
    sd["fc.bias"] = torch.tensor(SANE_POSE_BIAS, dtype=torch.float32)
    return sd


# ------------------------------------------------------------------------------------------------------------------
# folding
# ------------------------------------------------------------------------------------------------------------------
def fold_conv(sd, p: str, ws: bool = False):
    """Effective (weight, bias) of the conv/linear stored under prefix `p` (fp32, CPU)."""
    if p + ".weight_orig" in sd:
        w = sd[p + ".weight_orig"].detach().float().cpu()
        u, v = sd[p + ".weight_u"].detach().float().cpu(), sd[p + ".weight_v"].detach().float().cpu()
        sigma = torch.dot(u, torch.mv(w.reshape(w.shape[0], -1), v))
        w = w / sigma
    elif p + ".weight" in sd:
        w = sd[p + ".weight"].detach().float().cpu()
        if ws:
            wm = w
            for d in range(1, w.dim()):
                wm = wm.mean(dim=d, keepdim=True)
            w = w - wm
            std = w.reshape(w.shape[0], -1).std(dim=1).reshape(-1, *([1] * (w.dim() - 1))) + 1e-5
            w = w / std
    else:
        raise KeyError(f"checkpoint has neither {p}.weight_orig nor {p}.weight")
    b = sd.get(p + ".bias")
    return w, (b.detach().float().cpu() if b is not None else None)
