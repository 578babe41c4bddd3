"""TEST INFRASTRUCTURE ONLY (oracle).  CPU restatement, in plain PyTorch fp32 functional ops, of the reference's
volumetric-avatar inference hot path.  Only tests/ (including the analysis scripts under tests/analysis/),
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this module; the product path
(emoportraits_b200/) never does.

Every function cites the reference file:line it restates.  The restatement is PINNED against outputs of the
unmodified reference run in the build container (oracle/make_golden.py -> tests/golden/*.pt; checked by
tests/test_oracle_golden.py; the non-default forward() arguments by tests/test_oracle_options.py against
tests/golden/va256_options.pt), because the reference itself has no tests or golden vectors.

The arithmetic itself lives in third-party PyTorch (reference pins pytorch=1.13.1, environment.yml:173-174; here torch
2.11 CPU, semantics of conv/group_norm/grid_sample/interpolate unchanged): F.conv2d/3d, F.group_norm, F.grid_sample,
F.interpolate, F.avg_pool*, torch.inverse.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------------------------
# config (the subset of experiments/args.txt + va_arguments.py defaults that shapes the hot path)
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    image_size: int = 512
    C: int = 96            # latent_volume_channels
    D: int = 16            # latent_volume_depth
    S: int = 64            # latent_volume_size
    embed_size: int = 4    # gen_embed_size
    max_channels: int = 512
    dummy_input_size: int = 8  # gen_dummy_input_size (Unet3D seed, WarpGenerator input_size is 4 = gen_embed_size?)
    warp_input_size: int = 4
    exp_grid: int = 128    # expression_embedder.py:88  image_size(=exp_image_size 256) // 2
    idt_image_size: int = 256


def config_from_state_dict(sd, image_size: int) -> OracleConfig:
    C_D = sd["local_encoder_nw.finale_layers.2.weight"].shape[0]
    idg = sd["uv_generator_nw.identity_grid"]
    D, S = idg.shape[2], idg.shape[3]
    return OracleConfig(image_size=image_size, C=C_D // D, D=D, S=S)


# ------------------------------------------------------------------------------------------------------------------
# weight transforms (static at inference)
# ------------------------------------------------------------------------------------------------------------------
def sn_weight(sd, p):
    """utils/spectral_norm.py:131-168 in eval mode (:180 no power iteration): W = W_orig / (u . (W_mat v))."""
    w = sd[p + ".weight_orig"].float()
    u, v = sd[p + ".weight_u"].float(), sd[p + ".weight_v"].float()
    sigma = torch.dot(u, torch.mv(w.reshape(w.shape[0], -1), v))
    return w / sigma


def ws_weight(w):
    """networks/volumetric_avatar/utils.py:894-902 / :909-915: per-out-channel mean, UNBIASED std + 1e-5."""
    w = w.float()
    dims = tuple(range(1, w.dim()))
    wm = w
    for d in dims:  # sequential means exactly as the reference chains .mean(dim=..)
        wm = wm.mean(dim=d, keepdim=True)
    w = w - wm
    std = w.reshape(w.shape[0], -1).std(dim=1).reshape(-1, *([1] * (w.dim() - 1))) + 1e-5
    return w / std


def conv_params(sd, p, ws=False):
    """Resolve a conv's effective (weight, bias): SN if `weight_orig` is stored, WS when the positional rule of
    utils.py:1071-1090 replaced it (caller knows), else plain."""
    if p + ".weight_orig" in sd:
        w = sn_weight(sd, p)
    else:
        w = sd[p + ".weight"].float()
        if ws:
            w = ws_weight(w)
    b = sd.get(p + ".bias")
    return w, (b.float() if b is not None else None)


# ------------------------------------------------------------------------------------------------------------------
# pose algebra
# ------------------------------------------------------------------------------------------------------------------
def get_transform_matrix(scale, rotation, translation):
    """utils/point_transforms.py:187-240."""
    b = scale.shape[0]
    eye = torch.eye(4, dtype=scale.dtype)[None].repeat_interleave(b, dim=0)
    S_ = eye.clone()
    S_[:, 0, 0], S_[:, 1, 1], S_[:, 2, 2] = scale[:, 0], scale[:, 1], scale[:, 2]
    R = eye.clone()
    rotation = rotation.clamp(-math.pi / 2, math.pi)
    yaw, pitch, roll = rotation[:, 0], rotation[:, 1], rotation[:, 2]
    cy, sy, cp, sp, cr, sr = yaw.cos(), yaw.sin(), pitch.cos(), pitch.sin(), roll.cos(), roll.sin()
    R[:, 0, 0] = cy * cp
    R[:, 0, 1] = cy * sp * sr - sy * cr
    R[:, 0, 2] = cy * sp * cr + sy * sr
    R[:, 1, 0] = sy * cp
    R[:, 1, 1] = sy * sp * sr + cy * cr
    R[:, 1, 2] = sy * sp * cr - cy * sr
    R[:, 2, 0] = -sp
    R[:, 2, 1] = cp * sr
    R[:, 2, 2] = cp * cr
    T = eye.clone()
    T[:, 0, 3], T[:, 1, 3], T[:, 2, 3] = translation[:, 0], translation[:, 1], translation[:, 2]
    return S_ @ R @ T


def get_mixing_theta(source_theta, target_theta, mix_old=False):
    """notebooks/infer.py:686-736, B = T = 1 per call.  mix_old=False: stretch_s * mean(stretch_t)/mean(stretch_s) @
    rotation_t @ translation_t (:729); mix_old=True: translation_t @ rotation_t @ stretch_s (:727).  Returns (T,3,4) fp32."""
    from scipy import linalg

    source_theta = source_theta[:, :3, :]
    target_theta = target_theta[:, :3, :]
    B = source_theta.shape[0]
    Tn = target_theta.shape[0] // B
    s_ = np.stack([np.eye(4) for _ in range(B)])
    t_ = np.stack([np.eye(4) for _ in range(B * Tn)])
    target_theta = target_theta.view(B, Tn, 3, 4).roll(1, dims=0).view(B * Tn, 3, 4)
    s_[:, :3, :] = source_theta.numpy()
    t_[:, :3, :] = target_theta.numpy()
    tt = np.stack([np.eye(4) for _ in range(B * Tn)])
    tt[:, :3, 3] = t_[:, :3, 3]
    sl = s_.copy(); sl[:, :3, 3] = 0
    tl = t_.copy(); tl[:, :3, 3] = 0
    out = []
    for b in range(B):
        s_rot, s_str = linalg.polar(sl[b])
        for t in range(Tn):
            t_rot, t_str = linalg.polar(tl[b * Tn + t])
            if mix_old:
                out.append(tt[b * Tn + t] @ t_rot @ s_str)
            else:
                out.append(s_str * t_str.mean() / s_str.mean() @ t_rot @ tt[b * Tn + t])
    return torch.from_numpy(np.stack(out))[:, :3].float()


def smooth_theta(state, theta34, momentum):
    """notebooks/infer.py:571-581 (smooth_pose=True): exponential smoothing of the driver pose over successive frames.
    state: dict holding 'theta' ((3,4) or None, the wrapper's self.theta); theta34 (b,3,4).  fp32, torch op order."""
    if state.get("theta") is None:
        state["theta"] = theta34[0].clone()
    out = []
    for i in range(theta34.shape[0]):
        state["theta"] = theta34[i] * momentum + state["theta"] * (1 - momentum)
        out.append(state["theta"].clone())
    return torch.stack(out)


def align_theta_2d(theta34):
    """expression_embedder.py:161-203 (use_smart_scale False): 2-D alignment theta from a (N,3,4) pose theta."""
    n = theta34.shape[0]
    eye = torch.zeros(n, 1, 4); eye[:, :, 3] = 1
    th = torch.cat([theta34[:, :3].float(), eye], dim=1)
    inv2d = th.inverse()[:, :, [0, 1, 3]][:, [0, 1, 3]]
    scale = torch.zeros_like(inv2d)
    scale[:, [0, 1], [0, 1]] = 0.5
    scale[:, 2, 2] = 1
    return torch.bmm(inv2d, scale)[:, :2]


def identity_grid_3d(D, S):
    """va.py:101-105."""
    gs, gz = torch.linspace(-1, 1, S), torch.linspace(-1, 1, D)
    w, v, u = torch.meshgrid(gz, gs, gs, indexing="ij")
    return torch.stack([u, v, w, torch.ones_like(u)], dim=3).view(1, -1, 4)


def rotation_warp(theta34, D, S):
    """infer.py:441-444 / :583-588: grid.bmm(theta[:, :3]^T).view(-1, D, S, S, 3)."""
    g = identity_grid_3d(D, S).repeat(theta34.shape[0], 1, 1)
    return g.bmm(theta34[:, :3].transpose(1, 2)).view(-1, D, S, S, 3)


def grid_sample(x, grid):
    """va.py:261-265: trilinear, zeros, align_corners=False."""
    return F.grid_sample(x.float(), grid.float(), padding_mode="zeros", align_corners=False)


# ------------------------------------------------------------------------------------------------------------------
# blocks
# ------------------------------------------------------------------------------------------------------------------
def group_norm(x, sd, p, ada=None):
    """nn.GroupNorm(32, C) utils.py:953,957; AdaptiveGroupNorm utils.py:302-325 when `ada` = (dw, db) per sample:
    y = (GN(x)*w + b) * (w + dw) + (b + db)  (assign_adaptive_norm_params :994-995)."""
    w, b = sd[p + ".weight"].float(), sd[p + ".bias"].float()
    y = F.group_norm(x, 32, w, b, 1e-5)
    if ada is not None:
        dw, db = ada
        aw, ab = w[None] + dw, b[None] + db
        shp = (x.shape[0], -1) + (1,) * (x.dim() - 2)
        y = y * aw.view(shp) + ab.view(shp)
    return y


def res_block(x, sd, p, dims=2, up=None, down=None, ws_first=True, ada=None, taps=None):
    """utils.py:661-788.  block_feats = [norm, relu, conv, norm, relu]; block = [conv (, pool)];
    skip = [1x1 conv] (, pool).  `up`: nearest scale factor applied first (2-D decoder); `down`: avgpool kernel."""
    conv = F.conv2d if dims == 2 else F.conv3d
    pool = F.avg_pool2d if dims == 2 else F.avg_pool3d
    inp = x
    if up:
        x = F.interpolate(x, scale_factor=up, mode="nearest")
    w1, b1 = conv_params(sd, p + ".block_feats.2", ws=ws_first)
    w2, b2 = conv_params(sd, p + ".block.0")
    h = F.relu(group_norm(x, sd, p + ".block_feats.0", ada[0] if ada else None))
    h = conv(h, w1, b1, padding=1)
    h = F.relu(group_norm(h, sd, p + ".block_feats.3", ada[1] if ada else None))
    if taps is not None:
        taps[p + ".feats"] = h
    h = conv(h, w2, b2, padding=1)
    if down:
        h = pool(h, down)
    s = inp
    if up:
        s = F.interpolate(s, scale_factor=up, mode="nearest")
    if (p + ".skip.0.weight_orig") in sd or (p + ".skip.0.weight") in sd:
        ws_, bs_ = conv_params(sd, p + ".skip.0")
        s = conv(s, ws_, bs_)
    if down:
        s = pool(s, down)
    return h + s


# ------------------------------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------------------------------
def local_encoder(sd, img, taps=None):
    """local_encoder.py:113-125 (ctor :48-111)."""
    p = "local_encoder_nw"
    s = img.shape[2]
    w, b = conv_params(sd, f"{p}.from_rgb_{s}px")
    x = F.conv2d(img, w, b, padding=3)
    i = 0
    while f"{p}.enc_{i}_block={s}px.block.0.weight_orig" in sd:
        x = res_block(x, sd, f"{p}.enc_{i}_block={s}px", down=2)
        s //= 2
        i += 1
    x = F.relu(group_norm(x, sd, f"{p}.finale_layers.0"))
    w, b = conv_params(sd, f"{p}.finale_layers.2", ws=True)
    return F.conv2d(x, w, b)


def _bn_or_gn(x, sd, p):
    """ResNet norm: eval-mode BatchNorm (running stats stored) or the GN that replace_bn_to_gn put there (utils.py:1020)."""
    if p + ".running_mean" in sd:
        return F.batch_norm(x, sd[p + ".running_mean"].float(), sd[p + ".running_var"].float(), sd[p + ".weight"].float(),
                            sd[p + ".bias"].float(), False, 0.0, 1e-5)
    return F.group_norm(x, 32, sd[p + ".weight"].float(), sd[p + ".bias"].float(), 1e-5)


def resnet_trunk(sd, p, x, gn: bool):
    """torchvision resnet18/50 forward up to layer4 (BasicBlock / Bottleneck), with the reference's SN/WS conv rewrites:
    with GN, convs that follow a GroupNorm sibling are weight-standardised (utils.py:1071-1090)."""
    w, b = conv_params(sd, p + ".conv1")
    x = F.conv2d(x, w, b, stride=2, padding=3)
    x = F.relu(_bn_or_gn(x, sd, p + ".bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li in range(1, 5):
        bi = 0
        while f"{p}.layer{li}.{bi}.bn1.weight" in sd:
            q = f"{p}.layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            bottleneck = f"{q}.bn3.weight" in sd
            idt = x
            if bottleneck:
                w, b = conv_params(sd, q + ".conv1")
                h = F.relu(_bn_or_gn(F.conv2d(x, w, b), sd, q + ".bn1"))
                w, b = conv_params(sd, q + ".conv2", ws=gn)
                h = F.relu(_bn_or_gn(F.conv2d(h, w, b, stride=stride, padding=1), sd, q + ".bn2"))
                w, b = conv_params(sd, q + ".conv3", ws=gn)
                h = _bn_or_gn(F.conv2d(h, w, b), sd, q + ".bn3")
            else:
                w, b = conv_params(sd, q + ".conv1")
                h = F.relu(_bn_or_gn(F.conv2d(x, w, b, stride=stride, padding=1), sd, q + ".bn1"))
                w, b = conv_params(sd, q + ".conv2", ws=gn)
                h = _bn_or_gn(F.conv2d(h, w, b, padding=1), sd, q + ".bn2")
            if f"{q}.downsample.1.weight" in sd:
                w, b = conv_params(sd, q + ".downsample.0")
                idt = _bn_or_gn(F.conv2d(x, w, b, stride=stride), sd, q + ".downsample.1")
            x = F.relu(h + idt)
            bi += 1
    return x


IMAGENET_MEAN = torch.tensor([0.485, 0.456, 0.406])[None, :, None, None]
IMAGENET_STD = torch.tensor([0.229, 0.224, 0.225])[None, :, None, None]


def idt_embed(sd, img, cfg: OracleConfig):
    """identity_embedder.py:80-89, 59-73: bilinear -> 256, normalise, ResNet-50 (GN), fc 1x1, adaptive avgpool 4x4."""
    p = "idt_embedder_nw"
    x = F.interpolate(img, size=(cfg.idt_image_size, cfg.idt_image_size), mode="bilinear")
    x = (x - IMAGENET_MEAN) / IMAGENET_STD
    x = resnet_trunk(sd, p + ".net", x, gn=True)
    w, b = conv_params(sd, p + ".net.fc")
    x = F.conv2d(x, w, b)
    return F.adaptive_avg_pool2d(x, cfg.embed_size)


def head_pose(hsd, img):
    """head_pose_regressor.py:22-31: bilinear -> 128, torchvision resnet18(num_classes=9) with eval BN -> (n, 9)
    = (scale xyz, yaw pitch roll, translation xyz)."""
    x = img
    if x.shape[2] != 128 or x.shape[3] != 128:
        x = F.interpolate(x, size=(128, 128), mode="bilinear")
    sd = {"r." + k: v for k, v in hsd.items()}
    f = resnet_trunk(sd, "r", x, gn=False).mean((2, 3))
    return F.linear(f, sd["r.fc.weight"].float(), sd["r.fc.bias"].float())


def expression_embed(sd, imgs, theta34, cfg: OracleConfig, taps=None):
    """expression_embedder.py:132-253 (estimate_kp_by_net=True, use_seg=False, eval) + ResNetWrapper :441-478.
    imgs (n,3,H,W), theta34 (n,3,4) -> pose_embed (n,128), aligned (n,3,128,128)."""
    p = "expression_embedder_nw"
    n = imgs.shape[0]
    th2d = align_theta_2d(theta34)
    g = torch.linspace(-1, 1, cfg.exp_grid)
    v, u = torch.meshgrid(g, g, indexing="ij")
    idg = torch.stack([u, v, torch.ones_like(u)], dim=2).view(1, -1, 3).repeat_interleave(n, dim=0)
    warp = idg.bmm(th2d.transpose(1, 2)).view(n, cfg.exp_grid, cfg.exp_grid, 2)
    aligned = F.grid_sample(imgs.float(), warp.float(), align_corners=False)
    x = (aligned - IMAGENET_MEAN) / IMAGENET_STD
    x = resnet_trunk(sd, p + ".net_face.net", x, gn=True)
    w, b = conv_params(sd, p + ".net_face.net.fc")
    x = F.conv2d(x, w, b)
    x = F.adaptive_avg_pool2d(x, 4).flatten(1)
    w, b = conv_params(sd, p + ".net_face.pose_head")
    if taps is not None:
        taps["aligned"] = aligned
    return F.linear(x, w, b), aligned


def predict_embed(sd, pose_embed, idt, cfg: OracleConfig):
    """va.py:813-885 with gen_pred_mixing True, cat_em False: 'orig' = conv1x1((unsqueeze(pose) + idt) * 0.5)."""
    b = pose_embed.shape[0]
    e = F.linear(pose_embed, sd["pose_unsqueeze_nw.weight"].float()).view(b, -1, cfg.embed_size, cfg.embed_size)
    w, _ = conv_params(sd, "warp_embed_head_orig_nw")
    o = F.conv2d((e + idt.repeat_interleave(b, dim=0)) * 0.5, w)
    return o.view(b, o.shape[1], cfg.embed_size ** 2)


def warp_generator(sd, p, embed, cfg: OracleConfig, taps=None):
    """warp_generator_resnet.py:125-181.  embed (b,512,16) -> warp (b,D,S,S,3), deltas (b,3,D,S,S)."""
    b = embed.shape[0]
    nblk = 0
    while f"{p}.blocks_3d.{nblk}.block.0.weight_orig" in sd:
        nblk += 1
    # ProjectorNorm utils.py:1140-1151: one (u, v) pair per AdaptiveGroupNorm in module order (2 per block)
    params = []
    i = 0
    while f"{p}.projector.u.{i}" in sd:
        u, v = sd[f"{p}.projector.u.{i}"].float(), sd[f"{p}.projector.v.{i}"].float()
        prm = u[None].matmul(embed).matmul(v[None])
        params.append((prm[..., 0], prm[..., 1]))
        i += 1
    isz = cfg.warp_input_size
    w, _ = conv_params(sd, p + ".first_conv")
    x = F.conv2d(embed.view(b, -1, isz, isz), w).view(b, -1, isz, isz, isz)
    size = [isz, isz, isz]
    ndr = int(math.log(cfg.S // isz, 2))
    for i in range(1, nblk + 1):
        size[1] *= 2; size[2] *= 2
        depth_new = min(cfg.D * 2 ** (ndr - i), size[1]) if i < ndr else cfg.D
        up_depth = depth_new > size[0]
        down_depth = depth_new < size[0]
        size[0] = depth_new
        x = F.interpolate(x, scale_factor=2 if up_depth else (1, 2, 2), mode="trilinear")
        x = res_block(x, sd, f"{p}.blocks_3d.{i - 1}", dims=3, ws_first=True, ada=(params[2 * (i - 1)], params[2 * (i - 1) + 1]))
        if down_depth:
            x = F.avg_pool3d(x, (2, 1, 1), (2, 1, 1))
        if taps is not None:
            taps[f"{p}.blk{i - 1}"] = x
    x = F.relu(group_norm(x.float(), sd, p + ".pre_head.0"))
    w, bb = conv_params(sd, p + ".head.0.0")
    deltas = torch.tanh(F.conv3d(x, w, bb, padding=1))
    warp = (sd[p + ".identity_grid"].float() + deltas).permute(0, 2, 3, 4, 1)
    return warp, deltas


def volume_source(sd, vol):
    """vpn_resblocks.py:22-49 -> resblocks_3d.py:9-62: ResBlock3d x num_blocks, gn_3d, all convs SN."""
    p = "volume_source_nw.net.net"
    i = 0
    while f"{p}.{i}.block.0.weight_orig" in sd:
        vol = res_block(vol, sd, f"{p}.{i}", dims=3, ws_first=False)
        i += 1
    return vol


def unet3d(sd, vol, cfg: OracleConfig, taps=None):
    """unet_3d.py:196-290 with gn_3d norms (projector has no adaptive layers, embed_dict ignored)."""
    p = "volume_process_nw"
    nb = 0
    while f"{p}.blocks_3d_down.{nb}.block.0.weight_orig" in sd:
        nb += 1
    x = vol
    feats = []
    size = [cfg.D, vol.shape[-1], vol.shape[-1]]
    for i in range(nb):
        kind = "none"
        if i < nb - 1:
            size[1] //= 2; size[2] //= 2
            depth_new = min(size[0] * 2, size[1])
            kind = "up" if depth_new > size[0] else ("down" if depth_new < size[0] else "none")
            size[0] = depth_new
            if kind == "up":
                x = F.interpolate(x, scale_factor=(2, 1, 1), mode="trilinear")
        x = res_block(x, sd, f"{p}.blocks_3d_down.{i}", dims=3, ws_first=False)
        feats.append(x)
        if i < nb - 1:
            x = F.avg_pool3d(x, 2, 2) if kind == "down" else F.avg_pool3d(x, (1, 2, 2), (1, 2, 2))
    feats = feats[::-1]
    x = sd[p + ".input_tensor"].float().repeat_interleave(vol.shape[0], dim=0)
    size = [x.shape[2], x.shape[3], x.shape[4]]
    for i in range(1, nb + 1):
        size[1] *= 2; size[2] *= 2
        depth_new = min(cfg.D * 2 ** (nb - i), size[1])
        kind = "up" if depth_new > size[0] else ("down" if depth_new < size[0] else "none")
        size[0] = depth_new
        x = F.interpolate(x, scale_factor=2 if kind == "up" else (1, 2, 2), mode="trilinear")
        skip = res_block(feats[i - 1], sd, f"{p}.skip_blocks_3d_up.{i - 1}", dims=3, ws_first=False)
        x = res_block(x + skip, sd, f"{p}.blocks_3d_up.{i - 1}", dims=3, ws_first=False)
        if kind == "down":
            x = F.avg_pool3d(x, (2, 1, 1), (2, 1, 1))
        if taps is not None:
            taps[f"{p}.up{i - 1}"] = x
    x = F.relu(group_norm(x, sd, p + ".head.0"))
    w, b = conv_params(sd, p + ".head.2")
    return F.conv3d(x, w, b)


def decoder(sd, feat2d, taps=None):
    """decoder.py:152-238 + ImageDecoder :398-410 (dec_use_adanorm False -> embed_dict ignored).
    Returns img (sigmoid), pre-sigmoid logits, res_decoder features, img features."""
    p = "decoder_nw"
    w, _ = conv_params(sd, p + ".res_decoder.0")
    x = F.conv2d(feat2d, w)
    i = 1
    while f"{p}.res_decoder.{i}.block.0.weight_orig" in sd:
        x = res_block(x, sd, f"{p}.res_decoder.{i}")
        if taps is not None:
            taps[f"dec.res{i}"] = x
        i += 1
    feat = x
    i = 0
    q = p + ".img_decoder.dec_img_blocks"
    while f"{q}.{i}.block.0.weight_orig" in sd:
        has_skip = f"{q}.{i}.skip.0.weight_orig" in sd
        # im_dec_num_lrs_per_resolution = 2: even blocks upsample (stride 2 nearest) and change width (decoder.py:287-300)
        x = res_block(x, sd, f"{q}.{i}", up=2 if has_skip else None)
        if taps is not None:
            taps[f"dec.img{i}"] = x
        i += 1
    h = F.relu(group_norm(x.float(), sd, p + ".img_decoder.dec_img_head.0"))
    w, b = conv_params(sd, p + ".img_decoder.dec_img_head.2", ws=True)
    logits = F.conv2d(h, w, b)
    return torch.sigmoid(logits), logits, feat, x


# ------------------------------------------------------------------------------------------------------------------
# the two passes of InferenceWrapper.forward (crop=False; the external mask networks are ones, as in oracle/ref_harness.py)
# ------------------------------------------------------------------------------------------------------------------
def source_pass(sd, hsd, src_img, cfg: OracleConfig, taps=None, src_mask=None, c_source_latent_volume=None,
                c_target_latent_volume=None):
    """notebooks/infer.py:374-507.  src_img (1,3,H,W) in [0,1] already at image_size.  src_mask (1,1,H,W) or None (ones):
    the identity embedder and the local encoder see src_img * mask (:425-426); the head-pose regressor (:430) and the
    expression embedder (use_seg=False, expression_embedder.py:134-137) see the image itself.  c_*_latent_volume replace
    the volumes exactly where the reference substitutes them (:491, :500)."""
    st = {}
    masked = src_img if src_mask is None else src_img * src_mask
    st["idt_embed"] = idt_embed(sd, masked, cfg)
    latents = local_encoder(sd, masked)
    srt = head_pose(hsd, src_img)
    theta_s = get_transform_matrix(srt[:, :3], srt[:, 3:6], srt[:, 6:9])
    st["source_theta"] = theta_s
    inv = theta_s.float().inverse()
    rot_warp = rotation_warp(inv, cfg.D, cfg.S)
    pose_embed, _ = expression_embed(sd, src_img, theta_s[:, :3], cfg)
    embed = predict_embed(sd, pose_embed, st["idt_embed"], cfg)
    xy_warp, _ = warp_generator(sd, "xy_generator_nw", embed, cfg)
    vol = latents.view(1, cfg.C, cfg.D, cfg.S, cfg.S)
    vol = volume_source(sd, vol)
    if c_source_latent_volume is not None:
        vol = c_source_latent_volume
    warped = grid_sample(grid_sample(vol, rot_warp), xy_warp)
    if c_target_latent_volume is not None:
        warped = c_target_latent_volume
    st["target_latent_volume"] = unet3d(sd, warped, cfg)
    if taps is not None:
        taps.update(latents=latents, source_pose_embed=pose_embed, source_embed=embed, xy_warp=xy_warp, vol_source=vol,
                    vol_warped=warped, srt_source=srt)
    return st


def driver_pass(sd, hsd, st, drv_img, cfg: OracleConfig, taps=None, mix=True, mix_old=False, target_theta=True,
                smooth=None, pose_momentum=0.5, custome_target_pose_embed=None, custome_target_theta_embed=None):
    """notebooks/infer.py:511-644.  drv_img (b,3,H,W).  Returns img (b,3,H,W) fp32 (before the clamp/PIL step).
    smooth: None, or the dict that carries the wrapper's self.theta between calls (smooth_pose=True, :571-581);
    custome_target_theta_embed = (scale, rotation, translation) replaces the regressed pose (:566-567);
    custome_target_pose_embed replaces the expression embedding (:602-603); target_theta=False rotates the volume by the
    SOURCE pose (:587-588)."""
    srt = head_pose(hsd, drv_img)
    theta_d = get_transform_matrix(srt[:, :3], srt[:, 3:6], srt[:, 6:9])
    if custome_target_theta_embed is not None:
        theta_d = get_transform_matrix(*custome_target_theta_embed)
    if mix:
        th34 = get_mixing_theta(st["source_theta"], theta_d, mix_old)
    else:
        th34 = theta_d[:, :3]
    if smooth is not None:
        th34 = smooth_theta(smooth, th34, pose_momentum)
    rot_warp = rotation_warp(th34 if target_theta else st["source_theta"][:, :3], cfg.D, cfg.S)
    pose_embed, aligned = expression_embed(sd, drv_img, th34, cfg)
    if custome_target_pose_embed is not None:
        pose_embed = custome_target_pose_embed
    embed = predict_embed(sd, pose_embed, st["idt_embed"], cfg)
    uv_warp, deltas = warp_generator(sd, "uv_generator_nw", embed, cfg, taps)
    vol = grid_sample(grid_sample(st["target_latent_volume"], uv_warp), rot_warp)
    feat2d = vol.view(vol.shape[0], cfg.C * cfg.D, cfg.S, cfg.S)
    img, logits, feat, img_f = decoder(sd, feat2d, taps)
    if taps is not None:
        taps.update(srt=srt, theta=th34, pose_embed=pose_embed, embed=embed, uv_warp=uv_warp, aligned_volume=vol,
                    logits=logits, dec_feat=feat, aligned_face=aligned)
    return img


# ------------------------------------------------------------------------------------------------------------------
# stage 2 (notebooks/infer_s2.py:351-387; local_encoder_old.py:25-117; decoder_s2_old.py:18-217, 346-475)
# ------------------------------------------------------------------------------------------------------------------
def _bn2d(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"].float(), sd[p + ".running_var"].float(), sd[p + ".weight"].float(),
                        sd[p + ".bias"].float(), False, 0.0, 1e-5)


def res_block_bn(x, sd, p, up=None, down=None):
    """utils.py:661-788 with eval-mode BatchNorm2d norms (stage-2 default norm_layer_type 'bn'), all convs SN."""
    inp = x
    if up:
        x = F.interpolate(x, scale_factor=up, mode="nearest")
    w1, b1 = conv_params(sd, p + ".block_feats.2")
    w2, b2 = conv_params(sd, p + ".block.0")
    h = F.relu(_bn2d(x, sd, p + ".block_feats.0"))
    h = F.conv2d(h, w1, b1, padding=1)
    h = F.relu(_bn2d(h, sd, p + ".block_feats.3"))
    h = F.conv2d(h, w2, b2, padding=1)
    if down:
        h = F.avg_pool2d(h, down)
    s = inp
    if up:
        s = F.interpolate(s, scale_factor=up, mode="nearest")
    if (p + ".skip.0.weight_orig") in sd:
        ws_, bs_ = conv_params(sd, p + ".skip.0")
        s = F.conv2d(s, ws_, bs_)
    if down:
        s = F.avg_pool2d(s, down)
    return h + s


def stage2_local_encoder(sd, img):
    """local_encoder_old.py:105-117."""
    p = "local_encoder"
    s = img.shape[2]
    w, b = conv_params(sd, f"{p}.from_rgb_{s}px")
    x = F.conv2d(img, w, b, padding=3)
    i = 0
    while f"{p}.enc_{i}_block={s}px.block.0.weight_orig" in sd:
        x = res_block_bn(x, sd, f"{p}.enc_{i}_block={s}px", down=2)
        s //= 2
        i += 1
    x = F.relu(_bn2d(x, sd, p + ".finale_layers.0"))
    w, b = conv_params(sd, p + ".finale_layers.2")
    return F.conv2d(x, w, b)


def stage2_decoder(sd, vol):
    """decoder_s2_old.py:130-217 + ImageDecoder_stage2.forward :461-475 (gen_use_adanorm False)."""
    p = "decoder"
    w, _ = conv_params(sd, p + ".res_decoder.0")
    x = F.conv2d(vol, w)
    i = 1
    while f"{p}.res_decoder.{i}.block.0.weight_orig" in sd:
        x = res_block_bn(x, sd, f"{p}.res_decoder.{i}")
        i += 1
    i = 0
    while f"{p}.img_decoder.dec_img_blocks.{i}.block.0.weight_orig" in sd:
        x = res_block_bn(x, sd, f"{p}.img_decoder.dec_img_blocks.{i}", up=2)
        i += 1
    for i in range(4):
        x = res_block_bn(x, sd, f"{p}.img_decoder.dec_img_feat_blocks.{i}", up=2 if i == 0 else None)
    h = F.relu(_bn2d(x.float(), sd, p + ".img_decoder.dec_img_head.0"))
    w, b = conv_params(sd, p + ".img_decoder.dec_img_head.2")
    return torch.tanh(F.conv2d(h, w, b))


def stage2_forward(sd, img, output_size):
    """infer_s2.py:351-376 with all masks == 1 (MODNet / BiSeNet out of scope).  Returns (resized, add, ffhq)."""
    resized = F.interpolate(img, mode="bilinear", size=(output_size, output_size), align_corners=False)
    vol = stage2_local_encoder(sd, resized)
    add = stage2_decoder(sd, vol)
    ffhq = (resized + add).clamp(min=0, max=1)
    return resized, add, ffhq


# ------------------------------------------------------------------------------------------------------------------
# mask pre/post-processing around the external mask networks.  `net` / `modnet` are callables with the
# signatures the reference uses; pinned against the unmodified reference code by oracle/make_golden_masks.py ->
# tests/golden/masks.pt (tests/test_masks.py)
# ------------------------------------------------------------------------------------------------------------------
PARSING_MEAN = (0.485, 0.456, 0.406)   # face_parcing.py:31
PARSING_STD = (0.229, 0.224, 0.225)    # face_parcing.py:32


def parsing_label_sets(mask_type=None):
    """face_parcing.py:34-53: the label lists of (mask, face_body, mask_body, mask_cloth)."""
    if mask_type is None:
        return ([1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 7, 8, 9, 14, 17, 18], [1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 7, 8, 9, 17, 18], [18], [16])
    mask = []
    for key, labs in (("face", [1, 2, 3, 4, 5, 6, 10, 11, 12, 13]), ("ears", [7, 8, 9]), ("neck", [14, 15]), ("hair", [17]), ("hat", [18]),
                      ("cloth", [16])):
        if key in mask_type:
            mask += labs
    return (mask, None, None, None)  # face_labels / body_labels / cloth_labels only exist for mask_type None (:40-42)


def face_parsing_forward(net, x, mask_type=None):
    """face_parcing.py:55-81 (FaceParsing.forward).  x (N,3,h,w) in [0,1] -> four int64 (N,1,h,w) masks."""
    h, w = x.shape[2:]
    mean = torch.tensor(PARSING_MEAN, dtype=x.dtype)[None, :, None, None]
    std = torch.tensor(PARSING_STD, dtype=x.dtype)[None, :, None, None]
    x = (x - mean) / std                                                   # :57
    x = F.interpolate(x, size=(512, 512), mode="bilinear")                 # :58
    y = net(x)[0]                                                          # :59
    y = F.interpolate(y, size=(h, w), mode="bilinear")                     # :60
    labels = y.argmax(1, keepdim=True)                                     # :62
    outs = []
    for labs in parsing_label_sets(mask_type):                             # :64-79
        m = torch.zeros_like(labels)
        for i in labs or []:
            m += labels == i
        outs.append(m)
    return tuple(outs), y, labels


def modnet_get_mask(modnet, img):
    """notebooks/infer.py:649-684 (InferenceWrapper.get_mask).  img (N,3,h,w) in [0,1] -> matte (N,1,h,w)."""
    im = (img - 0.5) / 0.5                                                 # :651-657 Normalize((0.5,)*3, (0.5,)*3)
    ref_size = 512
    im_b, im_c, im_h, im_w = im.shape
    if max(im_h, im_w) < ref_size or min(im_h, im_w) > ref_size:           # :663-669
        if im_w >= im_h:
            im_rh = ref_size
            im_rw = int(im_w / im_h * ref_size)
        else:
            im_rw = ref_size
            im_rh = int(im_h / im_w * ref_size)
    else:
        im_rh, im_rw = im_h, im_w
    im_rw = im_rw - im_rw % 32                                             # :674-675
    im_rh = im_rh - im_rh % 32
    im = F.interpolate(im, size=(im_rh, im_rw), mode="area")               # :676
    _, _, matte = modnet(im, True)                                         # :679
    return F.interpolate(matte, size=(im_h, im_w), mode="area")            # :682
