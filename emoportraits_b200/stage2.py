"""Stage-2 refinement path: drop-in for `notebooks/infer_s2.py:53-387`
InferenceWrapper.forward — `LocalEncoderOld` (`networks/volumetric_avatar/local_encoder_old.py:25-117`) followed by
`Decoder_stage2` / `ImageDecoder_stage2` (`decoder_s2_old.py:18-217, 346-475`), built as in
`models/stage_2/base/volumetric_avatar_two.py:338-444`, all on the same sm_100a kernels as stage 1.

Default stage-2 configuration (argparse defaults of volumetric_avatar_two.py — the shipped stage-2 args.txt is a
download): norm 'bn' (eval-mode BatchNorm == per-channel affine, folded at load), spectral norm on every conv, no weight
standardisation, encoder 3->128->256->512->512 from output_size_s2 down to 64^2, 1x1 -> 64*16 channels; decoder 1x1
1024->512, 8 ResBlocks @64^2, nearest-x2 ResBlocks 512->256->128->128 up to output_size_s2, feature blocks 128->128->64->32,
BN + ReLU + 1x1 32->3 + Tanh.  The output is an additive residual on the bilinearly resized input.
"""
from __future__ import annotations

import math
import pathlib
from collections import OrderedDict
from dataclasses import dataclass
from typing import Optional

import torch

from . import nets, ops
from .checkpoint import _fill, _sn, fold_conv
from .config import parse_args


@dataclass
class Stage2Config:
    output_size: int = 512        # output_size_s2
    S: int = 64                   # gen_latent_texture_size2
    depth: int = 16               # gen_latent_texture_depth
    C: int = 64                   # gen_latent_texture_channels2
    gen_num_channels: int = 32
    gen_max_channels: int = 512
    enc_channel_mult: float = 4.0  # enc_channel_mult_stage2
    dec_channel_mult: float = 4.0  # dec_channel_mult_stage2
    dec_num_blocks: int = 8        # dec_num_blocks_stage2
    dec_max_channels: int = 512    # dec_max_channels2

    @property
    def enc_channels(self):
        c = int(self.gen_num_channels * self.enc_channel_mult)
        out = [c]
        for _ in range(int(math.log2(self.output_size // self.S))):
            c = min(c * 2, self.gen_max_channels)
            out.append(c)
        return out

    @property
    def dec_width(self):
        nup = int(math.log2(self.output_size // self.S))
        return min(int(self.gen_num_channels * self.dec_channel_mult * 2 ** nup), self.dec_max_channels)

    @property
    def up_channels(self):
        """ImageDecoder_stage2.dec_img_blocks widths (decoder_s2_old.py:374-391): num_up_blocks - 1 halvings"""
        nup = int(math.log2(self.output_size // self.S))
        c = self.dec_width
        out = [c]
        for _ in range(nup - 1):
            c = max(c // 2, self.gen_num_channels)
            out.append(c)
        return out


def stage2_config(args) -> Stage2Config:
    g = lambda k, dflt: getattr(args, k, dflt)
    bad = []
    if g("norm_layer_type", "bn") != "bn":
        bad.append(f"norm_layer_type={g('norm_layer_type', 'bn')} (stage 2 implements the default 'bn')")
    if g("use_ws", False):
        bad.append("use_ws=True")
    if not g("use_sn", True):
        bad.append("use_sn=False")
    for k in ("gen_use_adaconv", "dec_pred_seg", "dec_pred_conf"):
        if g(k, False):
            bad.append(f"{k}=True")
    if bad:
        raise NotImplementedError("stage 2: unsupported configuration: " + "; ".join(bad))
    return Stage2Config(output_size=int(g("output_size_s2", 512)), S=int(g("gen_latent_texture_size2", 64)),
                        depth=int(g("gen_latent_texture_depth", 16)), C=int(g("gen_latent_texture_channels2", 64)),
                        gen_num_channels=int(g("gen_num_channels", 32)), gen_max_channels=int(g("gen_max_channels", 512)),
                        enc_channel_mult=float(g("enc_channel_mult_stage2", 4.0)),
                        dec_channel_mult=float(g("dec_channel_mult_stage2", 4.0)),
                        dec_num_blocks=int(g("dec_num_blocks_stage2", 8)), dec_max_channels=int(g("dec_max_channels2", 512)))


# ------------------------------------------------------------------------------------------------------------------
# checkpoint layout
# ------------------------------------------------------------------------------------------------------------------
def _bn(spec, p, c):
    spec[p + ".weight"] = (c,)
    spec[p + ".bias"] = (c,)
    spec[p + ".running_mean"] = (c,)
    spec[p + ".running_var"] = (c,)
    spec[p + ".num_batches_tracked"] = ()


def _resblock_bn(spec, p, cin, cout, skip):
    _bn(spec, p + ".block_feats.0", cin)
    _sn(spec, p + ".block_feats.2", cout, cin, (3, 3))
    _bn(spec, p + ".block_feats.3", cout)
    _sn(spec, p + ".block.0", cout, cout, (3, 3))
    if skip:
        _sn(spec, p + ".skip.0", cout, cin, (1, 1))


def state_dict_spec_s2(cfg: Stage2Config) -> "OrderedDict[str, tuple]":
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    p = "local_encoder"
    ec = cfg.enc_channels
    s = cfg.output_size
    _sn(spec, f"{p}.from_rgb_{s}px", ec[0], 3, (7, 7), bias=True)
    for i in range(len(ec) - 1):
        _resblock_bn(spec, f"{p}.enc_{i}_block={s}px", ec[i], ec[i + 1], skip=ec[i] != ec[i + 1])
        s //= 2
    _bn(spec, p + ".finale_layers.0", ec[-1])
    _sn(spec, p + ".finale_layers.2", cfg.C * cfg.depth, ec[-1], (1, 1), bias=True)
    p = "decoder"
    w = cfg.dec_width
    _sn(spec, p + ".res_decoder.0", w, cfg.C * cfg.depth, (1, 1))
    for i in range(cfg.dec_num_blocks):
        _resblock_bn(spec, f"{p}.res_decoder.{i + 1}", w, w, skip=False)
    uc = cfg.up_channels
    for i in range(len(uc) - 1):
        _resblock_bn(spec, f"{p}.img_decoder.dec_img_blocks.{i}", uc[i], uc[i + 1], skip=uc[i] != uc[i + 1])
    fc = [uc[-1], 128, 128, 64, 32]  # decoder_s2_old.py:396-423: up block -> 128, then feature blocks 128, 64, 32
    for i in range(4):
        _resblock_bn(spec, f"{p}.img_decoder.dec_img_feat_blocks.{i}", fc[i], fc[i + 1], skip=fc[i] != fc[i + 1])
    _bn(spec, p + ".img_decoder.dec_img_head.0", 32)
    _sn(spec, p + ".img_decoder.dec_img_head.2", 3, 32, (1, 1), bias=True)
    spec["pose_unsqueeze.weight"] = (512 * 16, 512)  # volumetric_avatar_two.py (unused by the inference path)
    return spec


def synthetic_state_dict_s2(cfg: Stage2Config, seed: int = 0):
    """Seeded, well-conditioned weights in the stage-2 checkpoint layout (small gamma in front of the tanh head so the
    additive residual is O(0.1))."""
    return _fill(state_dict_spec_s2(cfg), seed + 1000, small_gamma=(("dec_img_head.0", 0.2),))


# ------------------------------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------------------------------
class _BN:
    """eval-mode BatchNorm2d folded to y = x*A + B (per channel)"""

    def __init__(self, sd, p, dev):
        g, b = sd[p + ".weight"].float(), sd[p + ".bias"].float()
        m, v = sd[p + ".running_mean"].float(), sd[p + ".running_var"].float()
        A = g / torch.sqrt(v + 1e-5)
        self.A = A[None].to(dev).contiguous()
        self.B = (b - m * A)[None].to(dev).contiguous()


class ResBlockBN:
    """utils.py:661-788 with norm_layer_type 'bn' in eval mode: no statistics are needed, the norm is an affine."""

    def __init__(self, sd, p, dev, planes=2, subpixel_up=False, pool_fold=False):
        self.planes = planes
        self.n1 = _BN(sd, p + ".block_feats.0", dev)
        self.c1 = nets.ConvW(sd, p + ".block_feats.2", dev, planes=planes)
        self.c1_ps = None  # sub-pixel form of `nearest x2 -> bn -> relu -> 3x3 conv` (see nets.ResBlock)
        if subpixel_up and planes == 2:
            w, _ = fold_conv(sd, p + ".block_feats.2")
            if tuple(w.shape[2:]) == (3, 3) and w.shape[0] % 32 == 0 and w.shape[1] % 64 == 0:
                self.c1_ps = ops.pack_upconv_weight(w, device=dev)
        self.n2 = _BN(sd, p + ".block_feats.3", dev)
        self.c2 = nets.ConvW(sd, p + ".block.0", dev, planes=planes)
        self.skip = nets.ConvW(sd, p + ".skip.0", dev, planes=planes) if (p + ".skip.0.weight_orig") in sd else None
        self.c2_pool = None  # `conv -> avgpool (1,2,2)` as one 4x4 stride-2 conv (see nets.ResBlock)
        if pool_fold:
            w, _ = fold_conv(sd, p + ".block.0")
            if tuple(w.shape[2:]) == (3, 3):
                self.c2_pool = ops.pack_conv_weight(ops.fold_poolconv_weight(w), device=dev, planes=planes)

    def __call__(self, x, up=1, down=None):
        P = self.planes
        if up == 2 and self.c1_ps is not None and x.shape[1] == 1 and (x.shape[0] * x.shape[2] * x.shape[3]) % 256 == 0:
            a = ops.apply(x, self.n1.A, self.n1.B, per_sample=False, act=ops.ACT_RELU, up=1, planes=P)
            y = ops.conv_igemm(a, self.c1_ps, bias=self.c1.b, upconv=True)
        else:
            a = ops.apply(x, self.n1.A, self.n1.B, per_sample=False, act=ops.ACT_RELU, up=up, planes=P)
            y = ops.conv_igemm(a, self.c1.w, bias=self.c1.b)
        b = ops.apply(y, self.n2.A, self.n2.B, per_sample=False, act=ops.ACT_RELU, planes=P)
        s = ops.avgpool(x, down) if down else x
        if self.skip is not None:
            s = ops.conv_igemm(ops.split_bf16(s, P), self.skip.w, bias=self.skip.b)
        if down and self.c2_pool is not None and tuple(down) == (1, 2, 2) and b.shape[2] % 2 == 0 and b.shape[3] % 2 == 0:
            return ops.conv_igemm(b, self.c2_pool, stride=(1, 2, 2), pad=(0, 1, 1), bias=self.c2.b, residual=s)
        if down:
            full = ops.conv_igemm(b, self.c2.w, bias=self.c2.b)
            return ops.avgpool(full, down, add=s)
        return ops.conv_igemm(b, self.c2.w, bias=self.c2.b, residual=s, res_shift=1 if up == 2 else 0)


class Stage2Model:
    """`model_two.local_encoder` + `model_two.decoder` of notebooks/infer_s2.py:371-372."""

    def __init__(self, cfg: Stage2Config, state_dict, device="cuda", planes: int = 2):
        self.cfg = cfg
        dev = torch.device(device)
        self.device = dev
        sd = state_dict
        self.planes = planes
        p = "local_encoder"
        s = cfg.output_size
        w, b = fold_conv(sd, f"{p}.from_rgb_{s}px")
        wc = torch.zeros(7, 7, 4, w.shape[0])
        wc[:, :, :3] = w.permute(2, 3, 1, 0)
        self.stem_w, self.stem_b = wc.to(dev).contiguous(), b.to(dev).contiguous()
        self.enc = []
        for i in range(len(cfg.enc_channels) - 1):
            self.enc.append(ResBlockBN(sd, f"{p}.enc_{i}_block={s}px", dev, planes, pool_fold=True))
            s //= 2
        self.fin_norm = _BN(sd, p + ".finale_layers.0", dev)
        self.fin = nets.ConvW(sd, p + ".finale_layers.2", dev, planes=planes)
        p = "decoder"
        self.inp = nets.ConvW(sd, p + ".res_decoder.0", dev, planes=planes)
        self.res = [ResBlockBN(sd, f"{p}.res_decoder.{i + 1}", dev, planes) for i in range(cfg.dec_num_blocks)]
        ps = True
        self.up = [ResBlockBN(sd, f"{p}.img_decoder.dec_img_blocks.{i}", dev, planes, subpixel_up=ps) for i in range(len(cfg.up_channels) - 1)]
        self.feat = [ResBlockBN(sd, f"{p}.img_decoder.dec_img_feat_blocks.{i}", dev, planes, subpixel_up=ps and i == 0) for i in range(4)]
        self.head_norm = _BN(sd, p + ".img_decoder.dec_img_head.0", dev)
        self.head = nets.ConvW(sd, p + ".img_decoder.dec_img_head.2", dev, planes=planes)

    @torch.no_grad()
    def local_encoder(self, x4):
        """x4 (N,1,S,S,4) channels-last (masked, resized) image -> latent (N,1,64,64,C*depth)"""
        x = ops.conv_direct(x4, self.stem_w, 1, 3, self.stem_b)
        for blk in self.enc:
            x = blk(x, down=(1, 2, 2))
        a = ops.apply(x, self.fin_norm.A, self.fin_norm.B, per_sample=False, act=ops.ACT_RELU, planes=self.planes)
        return ops.conv_igemm(a, self.fin.w, bias=self.fin.b)

    @torch.no_grad()
    def decoder(self, vol):
        """latent channels-last (N,1,64,64,C*depth) -> additive residual (N,3,S,S) NCHW in (-1, 1) (Tanh head)"""
        P = self.planes
        x = ops.conv_igemm(ops.split_bf16(vol, P), self.inp.w)
        for blk in self.res:
            x = blk(x)
        for blk in self.up:
            x = blk(x, up=2)
        x = self.feat[0](x, up=2)
        for blk in self.feat[1:]:
            x = blk(x)
        a = ops.apply(x, self.head_norm.A, self.head_norm.B, per_sample=False, act=ops.ACT_RELU, planes=P)
        out = ops.conv_igemm(a, self.head.w, bias=self.head.b, act=ops.ACT_TANH, out_nchw=True)
        return out[:, :, 0]

    @torch.no_grad()
    def forward(self, img: torch.Tensor, mask: Optional[torch.Tensor] = None, face_mask: Optional[torch.Tensor] = None):
        """notebooks/infer_s2.py:351-376.  img (N,3,h,w) in [0,1] on device.  Masks default to ones (MODNet / BiSeNet are
        out of scope).  Returns (resized, add, ffhq) NCHW fp32."""
        S = self.cfg.output_size
        img = img.contiguous().float()
        N = img.shape[0]
        # bilinear resize (align_corners=False) both as the NCHW tensor the residual is added to and as the encoder input
        x4 = ops.resize_bilinear(img, (S, S))                                   # (N,1,S,S,4) channels-last, 4th channel 0
        resized = x4[:, 0, :, :, :3].permute(0, 3, 1, 2).contiguous()
        if mask is not None:
            x4 = (x4 * mask.to(self.device).float().permute(0, 2, 3, 1)[:, None]).contiguous()
        add = self.decoder(self.local_encoder(x4))
        fm = (mask if mask is not None else 1.0)
        if face_mask is not None:
            fm = fm * face_mask.to(self.device).float()
        add = add * fm if not isinstance(fm, float) else add
        ffhq = (resized + add).clamp_(min=0, max=1)
        return resized, add, ffhq


class InferenceWrapper(torch.nn.Module):
    """Drop-in for notebooks/infer_s2.py:53 InferenceWrapper (stage 2): same ctor arguments, `forward(img)` returns
    (pil, pil_resized, pil_ffhq, mask).  Reads `project_dir/logs_s2/<experiment_name>/{args.txt, checkpoints/<file>}`."""

    def __init__(self, experiment_name, which_epoch='latest', model_file_name='', use_gpu=True, num_gpus=1,
                 fixed_bounding_box=False, project_dir='./', torch_home='', debug=False, print_model=False,
                 args_overwrite={}, pose_momentum=0.5, experiment_name_s1=None, model_file_name_s1=None, cloth=False,
                 state_dict=None):
        super().__init__()
        if not use_gpu or not torch.cuda.is_available():
            raise RuntimeError("emoportraits_b200 stage-2 InferenceWrapper needs a CUDA device; there is no CPU path")
        self.cloth = cloth
        args_path = pathlib.Path(project_dir) / 'logs_s2' / experiment_name / 'args.txt'
        self.args = parse_args(args_path)
        self.args.project_dir = project_dir
        for k, v in (args_overwrite or {}).items():
            setattr(self.args, k, v)
        self.cfg = stage2_config(self.args)
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.model_checkpoint_s2 = pathlib.Path(project_dir) / 'logs_s2' / experiment_name / 'checkpoints' / model_file_name
        if state_dict is None:
            state_dict = torch.load(self.model_checkpoint_s2, map_location='cpu')
        self.model_two = Stage2Model(self.cfg, state_dict, self.device)

    @torch.no_grad()
    def forward(self, img, cloth=False, mask=None, face_mask=None):
        from PIL import Image

        img = img.to(self.device).float()
        if self.cloth:
            face_mask = None  # infer_s2.py:365-368: with cloth=True the refinement is not restricted to the face region
        resized, add, ffhq = self.model_two.forward(img, mask, face_mask)
        to_pil = lambda t: [Image.fromarray(h) for h in ops.image_to_u8(t.detach().contiguous().float()).cpu().numpy()]
        m = mask if mask is not None else torch.ones_like(img[:, :1])
        return to_pil(img), to_pil(resized), to_pil(ffhq), m.detach().cpu().clamp(0, 1)
