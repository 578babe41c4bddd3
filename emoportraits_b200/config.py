"""`args.txt` reader and the hot-path hyper-parameters derived from it.

Mirrors the reference's config contract (SURVEY.md §5 "Config / flag system"):
  - `args.txt` is one `k: v` per line as written by train.py:80-83 and re-parsed with type sniffing by
    utils/args.py:34-65 (digits -> int, float-looking -> float, 'True'/'False' -> bool, everything else str —
    including the *string* 'None');
  - values missing from the file fall back to the defaults of models/stage_1/volumetric_avatar/va_arguments.py:19-355.
Only the keys that shape the inference hot path are kept; configurations the B200 path does not implement
raise NotImplementedError instead of silently computing something else.
"""
from __future__ import annotations

import math
import pathlib
from dataclasses import dataclass, fields
from types import SimpleNamespace


def parse_args_line(line: str):
    """utils/args.py:34-52."""
    parts = line.rstrip("\n").split(": ")
    if len(parts) > 2:
        parts = [parts[0], ": ".join(parts[1:])]
    if len(parts) != 2:
        return None
    k, v = parts

    def isfloat(s):
        try:
            float(s)
            return True
        except ValueError:
            return False

    if v.isdigit():
        v = int(v)
    elif isfloat(v):
        v = float(v)
    elif v == "True":
        v = True
    elif v == "False":
        v = False
    return k, v


def parse_args(path) -> SimpleNamespace:
    """utils/args.py:54-65 (without argparse: the file alone defines the namespace)."""
    ns = SimpleNamespace()
    for line in pathlib.Path(path).read_text().splitlines():
        kv = parse_args_line(line)
        if kv:
            setattr(ns, kv[0], kv[1])
    return ns


# defaults: va_arguments.py:19-355 / datasets/voxceleb2hq_pairs.py:530-562 (only hot-path keys)
_DEFAULTS = dict(
    image_size=256, latent_volume_channels=64, latent_volume_depth=16, latent_volume_size=64,
    gen_num_channels=32, gen_max_channels=512, gen_max_channels_unet3d=512, enc_channel_mult=2.0,
    dec_max_channels=512, dec_channel_mult=2.0, dec_num_blocks=8, im_dec_num_lrs_per_resolution=1,
    im_dec_ch_div_factor=2.0, warp_channel_mult=1.0, warp_output_size=64, gen_embed_size=4, gen_dummy_input_size=4,
    lpe_output_channels_expression=512, lpe_output_size=4, idt_output_channels=512, idt_output_size=4,
    idt_image_size=256, exp_image_size=256, source_volume_num_blocks=0, pred_volume_num_blocks=0,
    norm_layer_type="bn", use_sn=True, use_ws=False, gen_use_adanorm=False, gen_use_adaconv=False,
    dec_use_adanorm=False, dec_use_adaconv=False, use_back=False, volume_rendering=False, use_tensor=False,
    unet_first=False, no_channel_increase_3d_source=True, grid_sample_padding_mode="zeros", warp_norm_grad=False,
    estimate_head_pose_from_keypoints=True, cat_em=False, gen_pred_mixing=True, dec_bigger=False, dec_pred_seg=False,
    lpe_face_backbone="resnet18", idt_backbone="resnet50", gen_activation_type="relu",
    gen_upsampling_type="trilinear", gen_downsampling_type="avgpool", enc_block_type="res", warp_block_type="res",
    dec_up_block_type="res", tex_use_skip_resblock=True, tex_pred_rgb=False, local_encoder_input_size=3,
    use_smart_scale=False, expr_custom_w=False, lpe_final_pooling_type="avg", num_source_frames=1,
    dec_use_sg3_img_dec=False,
)


@dataclass
class HotPathConfig:
    image_size: int
    C: int          # latent_volume_channels
    D: int          # latent_volume_depth
    S: int          # latent_volume_size
    gen_num_channels: int
    gen_max_channels: int
    unet_max_channels: int
    enc_channel_mult: float
    dec_max_channels: int
    dec_channel_mult: float
    dec_num_blocks: int
    im_dec_lrs: int
    im_dec_div: float
    warp_channel_mult: float
    warp_output_size: int
    embed_size: int
    dummy_input_size: int
    expr_channels: int
    idt_channels: int
    idt_image_size: int
    exp_image_size: int
    source_volume_num_blocks: int

    # ---- derived shapes (same arithmetic as the reference constructors) ----
    @property
    def enc_channels(self):
        """local_encoder.py:62,76-79: stem width then doubling per down block, capped."""
        c = int(self.gen_num_channels * self.enc_channel_mult)
        out = [c]
        for _ in range(int(math.log2(self.image_size // self.S))):
            c = min(c * 2, self.gen_max_channels)
            out.append(c)
        return out

    @property
    def dec_channels(self):
        """decoder.py:57-59, 283-287: trunk width, then per-resolution widths of ImageDecoder."""
        nup = int(math.log2(self.image_size // self.S))
        c = min(int(self.gen_num_channels * self.dec_channel_mult * 2 ** nup), self.dec_max_channels)
        out = [c]
        for _ in range(nup):
            c = max(int(c / self.im_dec_div / 32) * 32, self.gen_num_channels)
            out.append(c)
        return out

    @property
    def warp_channels(self):
        """warp_generator_resnet.py:58-72."""
        nb = int(math.log2(self.warp_output_size // self.embed_size))
        f = lambda i: (min(int(self.gen_num_channels * self.warp_channel_mult * 2 ** i), self.gen_max_channels)) // 32 * 32
        return [f(nb)] + [f(i) for i in range(nb - 1, -1, -1)]

    @property
    def unet_channels(self):
        """unet_3d.py:66-76: widths of the down path (C, 2C, 4C, ... capped)."""
        nb = int(math.log2(self.S // self.dummy_input_size))
        c = self.C
        out = [c]
        for _ in range(nb):
            c = min(c * 2, self.unet_max_channels)
            out.append(c)
        return out


def hot_path_config(args) -> HotPathConfig:
    g = lambda k: getattr(args, k, _DEFAULTS[k])
    unsupported = []
    if g("norm_layer_type") != "gn":
        unsupported.append(f"norm_layer_type={g('norm_layer_type')} (only 'gn', the shipped config)")
    for k in ["use_back", "volume_rendering", "use_tensor", "unet_first", "gen_use_adanorm", "gen_use_adaconv",
              "dec_use_adanorm", "dec_use_adaconv", "warp_norm_grad", "cat_em", "dec_bigger", "dec_pred_seg",
              "tex_pred_rgb", "use_smart_scale", "expr_custom_w", "dec_use_sg3_img_dec"]:
        if g(k):
            unsupported.append(f"{k}=True")
    for k, want in [("use_sn", True), ("use_ws", True), ("estimate_head_pose_from_keypoints", True),
                    ("gen_pred_mixing", True), ("no_channel_increase_3d_source", True), ("tex_use_skip_resblock", True),
                    ("grid_sample_padding_mode", "zeros"), ("lpe_face_backbone", "resnet18"), ("idt_backbone", "resnet50"),
                    ("gen_activation_type", "relu"), ("gen_upsampling_type", "trilinear"),
                    ("gen_downsampling_type", "avgpool"), ("enc_block_type", "res"), ("warp_block_type", "res"),
                    ("dec_up_block_type", "res"), ("local_encoder_input_size", 3), ("lpe_final_pooling_type", "avg"),
                    ("num_source_frames", 1), ("pred_volume_num_blocks", 0), ("lpe_output_size", 4), ("idt_output_size", 4)]:
        if g(k) != want:
            unsupported.append(f"{k}={g(k)} (need {want})")
    if unsupported:
        raise NotImplementedError("emoportraits_b200 implements the shipped inference configuration only; unsupported: "
                                  + "; ".join(unsupported))
    C = getattr(args, "latent_volume_channels", None) or _DEFAULTS["latent_volume_channels"]
    cfg = HotPathConfig(
        image_size=int(g("image_size")), C=int(C), D=int(g("latent_volume_depth")), S=int(g("latent_volume_size")),
        gen_num_channels=int(g("gen_num_channels")), gen_max_channels=int(g("gen_max_channels")),
        unet_max_channels=int(g("gen_max_channels_unet3d")), enc_channel_mult=float(g("enc_channel_mult")),
        dec_max_channels=int(g("dec_max_channels")), dec_channel_mult=float(g("dec_channel_mult")),
        dec_num_blocks=int(g("dec_num_blocks")), im_dec_lrs=int(g("im_dec_num_lrs_per_resolution")),
        im_dec_div=float(g("im_dec_ch_div_factor")), warp_channel_mult=float(g("warp_channel_mult")),
        warp_output_size=int(g("warp_output_size")), embed_size=int(g("gen_embed_size")),
        dummy_input_size=int(g("gen_dummy_input_size")), expr_channels=int(g("lpe_output_channels_expression")),
        idt_channels=int(g("idt_output_channels")), idt_image_size=int(g("idt_image_size")),
        exp_image_size=int(g("exp_image_size")), source_volume_num_blocks=int(g("source_volume_num_blocks")))
    if cfg.warp_output_size != cfg.S:
        raise NotImplementedError("warp_output_size != latent_volume_size (resize_warp) is not implemented")
    if cfg.embed_size != 4 or cfg.idt_channels != cfg.gen_max_channels:
        raise NotImplementedError("gen_embed_size != 4 or idt_output_channels != gen_max_channels")
    return cfg


def shipped_config(image_size: int = 512) -> HotPathConfig:
    """The one configuration the reference ships (experiments/args.txt:1 over the defaults)."""
    ns = SimpleNamespace(
        image_size=image_size, latent_volume_channels=96, latent_volume_depth=16, latent_volume_size=64,
        gen_num_channels=32, gen_max_channels=512, gen_max_channels_unet3d=512, enc_channel_mult=4.0, dec_max_channels=512,
        dec_channel_mult=2.0, dec_num_blocks=6, im_dec_num_lrs_per_resolution=2, im_dec_ch_div_factor=1.5,
        gen_dummy_input_size=8, lpe_output_channels_expression=128, source_volume_num_blocks=3, norm_layer_type="gn",
        use_ws=True)
    return hot_path_config(ns)


def config_fields():
    return [f.name for f in fields(HotPathConfig)]
