"""Drop-in for the reference's stage-1 inference API: `notebooks/infer.py:62-647` InferenceWrapper.

Same constructor and `forward` signature, same `args.txt` / checkpoint layout (`project_dir/folder/experiment_name/
{args.txt, checkpoints/<model_file_name>}`), same cached source state on `self`, same return value
`(list[PIL.Image], img Tensor (B,3,H,W) fp32 on device)`.  The hot path underneath is emoportraits_b200's sm_100a
kernels; there is no torch/CPU fallback — without a B200 and the built libemoport.so this module raises.

Out of scope here: the external face detector (`crop=True` raises) and the WEIGHTS / model
code of the external mask networks (MODNet, BiSeNet: separate checkouts, not in the reference tree).  The processing the
reference does around them is here (`masks.py`, csrc/masks.cu): plug the networks in with `wrapper.face_idt =
masks.FaceParsing(None, net=...)` / `wrapper.modnet = ...`; without them the source mask is `source_mask` or all ones (what the
stubbed reference oracle uses too) and `modnet_mask=True` raises.
"""
from __future__ import annotations

import pathlib
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import nets, ops
from .config import HotPathConfig, hot_path_config, parse_args


class Model:
    """Inference-only counterpart of models/stage_1/volumetric_avatar/va.py:39 Model (attribute names kept)."""

    # Operand planes per network (ops.conv_igemm): 2 = two bf16 planes, 3 MMAs per product, ~2^-16 relative (decoder: isolated
    # stage error 2e-4 on logits of magnitude 4); ops.H2 = two fp16 planes of (value x power of two), 3 MMAs per product,
    # ~2^-22 relative, i.e. fp32-faithful — for the embedding / warp networks, whose output the sampler differentiates (an
    # error of 1e-4 in the embedding moves the image by 1e-2), and for the source-pass networks.  Three bf16 planes (6 MMAs,
    # the round-1 setting of those networks) give the same accuracy at twice the MMAs (tests/test_conv_modes_gpu.py) and
    # remain available through `precision=`.  The decoder's input planes are written by grid_sample_3d in bf16.
    PRECISION = dict(head_pose=ops.H2, expression=ops.H2, idt=ops.H2, warp=ops.H2, local_encoder=ops.H2, volume_source=ops.H2,
                     unet3d=ops.H2, decoder=2)

    def __init__(self, cfg: HotPathConfig, state_dict, head_pose_state_dict, device="cuda", precision: Optional[dict] = None):
        self.cfg = cfg
        dev = torch.device(device)
        sd = state_dict
        pr = dict(self.PRECISION)
        pr.update(precision or {})
        if pr["decoder"] == ops.H2:
            raise ValueError("the decoder's input planes are written by grid_sample_3d in bf16; it runs with 2 or 3 bf16 planes")
        self.precision = pr
        self.local_encoder_nw = nets.LocalEncoder(sd, cfg, dev, planes=pr["local_encoder"])
        self.idt_embedder_nw = nets.IdtEmbed(sd, cfg, dev, planes=pr["idt"])
        self.expression_embedder_nw = nets.ExpressionEmbed(sd, cfg, dev, planes=pr["expression"])
        self.predict_embed = nets.PredictEmbed(sd, cfg, dev)
        self.xy_generator_nw = nets.WarpGenerator(sd, "xy_generator_nw", cfg, dev, planes=pr["warp"])
        self.uv_generator_nw = nets.WarpGenerator(sd, "uv_generator_nw", cfg, dev, planes=pr["warp"])
        self.volume_source_nw = nets.VolumeSource(sd, cfg, dev, planes=pr["volume_source"]) if cfg.source_volume_num_blocks > 0 else None
        self.volume_process_nw = nets.Unet3D(sd, cfg, dev, planes=pr["unet3d"])
        self.decoder_nw = nets.Decoder(sd, cfg, dev, planes=pr["decoder"])
        self.head_pose_regressor = nets.HeadPoseRegressor(head_pose_state_dict, dev, planes=pr["head_pose"])
        self.device = dev

    # ---- notebooks/infer.py:374-507 ----
    @torch.no_grad()
    def source_pass(self, src: torch.Tensor, taps: Optional[dict] = None, pose_override=None,
                    mask: Optional[torch.Tensor] = None, c_source_latent_volume: Optional[torch.Tensor] = None,
                    c_target_latent_volume: Optional[torch.Tensor] = None):
        """src (1,3,H,W) fp32 in [0,1] on device.  Returns the cached source state.
        mask (1,1,H,W) or None (= ones): the identity embedder and the local encoder see src * mask (infer.py:425-426);
        the head-pose regressor (:430) and the expression embedder (called with use_seg=False, :455) see src itself.
        c_source_latent_volume / c_target_latent_volume (1,C,D,S,S) replace the volumes where the reference substitutes
        them (:491, :500).
        pose_override = (theta (1,4,4), warp (1,3,4), align2d (1,2,3)) replaces the on-device pose algebra (tests only:
        lets a parity test inject the reference's own fp32 matrices, see tests/test_model_gpu.py)."""
        cfg = self.cfg
        src = src.contiguous().float()
        masked = src if mask is None else (src * mask.to(self.device).float()).contiguous()
        ops.begin_pass(self.device)
        st = SimpleNamespace()
        st.idt_embed = self.idt_embedder_nw(masked)                    # (1,512,4,4) NCHW
        vol = self.local_encoder_nw(masked)                            # (1,D,S,S,C)
        srt = self.head_pose_regressor(src)
        st.pred_source_theta, inv_warp, align = ops.pose_theta(srt, invert_warp=True)
        if pose_override is not None:
            st.pred_source_theta, inv_warp, align = [t.to(self.device).float().contiguous() for t in pose_override]
        st.source_theta_dev = st.pred_source_theta[0].contiguous()
        pose_embed, _ = self.expression_embedder_nw(src, align)
        st.pred_source_pose_embed = pose_embed
        E = self.predict_embed(pose_embed, st.idt_embed)
        xy_warp = self.xy_generator_nw(E)                              # (1,D,S,S,3)
        if self.volume_source_nw is not None:
            vol = self.volume_source_nw(vol)
        if c_source_latent_volume is not None:
            vol = self._volume_cl(c_source_latent_volume)
        st.source_latent_volume = vol
        st.source_rotation_warp_theta = inv_warp
        st.source_xy_warp_resize = xy_warp
        v = ops.grid_sample3d(vol, theta=inv_warp, out_size=(cfg.D, cfg.S, cfg.S), in_layout="cl")
        v = ops.grid_sample3d(v, grid=xy_warp, in_layout="cl")
        if c_target_latent_volume is not None:
            v = self._volume_cl(c_target_latent_volume)
        st.target_latent_volume_1 = v
        st.target_latent_volume = self.volume_process_nw(v)           # (1,D,S,S,C) channels-last
        if taps is not None:
            taps.update(srt_source=srt, source_pose_embed=pose_embed, source_embed=E, xy_warp=xy_warp)
        return st

    def _volume_cl(self, vol: torch.Tensor) -> torch.Tensor:
        """caller-supplied latent volume in the reference's (1,C,D,S,S) layout -> channels-last (1,D,S,S,C) on device"""
        cfg = self.cfg
        if tuple(vol.shape) != (1, cfg.C, cfg.D, cfg.S, cfg.S):
            raise ValueError(f"latent volume must be (1,{cfg.C},{cfg.D},{cfg.S},{cfg.S}), got {tuple(vol.shape)}")
        return vol.to(self.device).float().permute(0, 2, 3, 4, 1).contiguous()

    # ---- notebooks/infer.py:511-644 ----
    @torch.no_grad()
    def driver_pass(self, st, drv: torch.Tensor, mix: bool = True, target_theta: bool = True, taps: Optional[dict] = None,
                    want_logits: bool = False, pose_override=None, mix_old: bool = False,
                    custom_srt: Optional[torch.Tensor] = None, custom_pose_embed: Optional[torch.Tensor] = None,
                    smooth_state: Optional[torch.Tensor] = None, smooth_momentum: float = 0.5, smooth_init: bool = False):
        """drv (1,3,H,W) fp32 on device -> (img (1,3,H,W), feat_2d, img_feat, state).
        mix / mix_old: get_mixing_theta (infer.py:569-570, 686-736).  custom_srt (1,9) = the caller's
        custome_target_theta_embed (scale, rotation, translation) replacing the regressed pose (:566-567).
        custom_pose_embed (1,128) replaces the expression embedding (:602-603).  smooth_state (3,4) on device: the
        wrapper's self.theta, smoothed in place with smooth_momentum (smooth_pose=True, :571-581)."""
        cfg = self.cfg
        drv = drv.contiguous().float()
        ops.begin_pass(self.device)
        srt = self.head_pose_regressor(drv)
        srt_used = srt if custom_srt is None else custom_srt.to(self.device).float().reshape(1, 9).contiguous()
        theta, warp, align = ops.pose_theta(srt_used, source_theta=st.source_theta_dev if mix else None, mix=mix,
                                            mix_old=mix_old, smooth_state=smooth_state, smooth_momentum=smooth_momentum,
                                            smooth_init=smooth_init)
        if pose_override is not None:
            theta, warp, align = [t.to(self.device).float().contiguous() for t in pose_override]
        if not target_theta:
            warp = st.pred_source_theta[:, :3].contiguous()
        if custom_pose_embed is not None and taps is None:
            pose_embed, aligned = None, None      # the expression encoder's output would be discarded (infer.py:602-603)
        else:
            pose_embed, aligned = self.expression_embedder_nw(drv, align, want_aligned=taps is not None)
        if custom_pose_embed is not None:
            pose_embed = custom_pose_embed.to(self.device).float().reshape(1, -1).contiguous()
        E = self.predict_embed(pose_embed, st.idt_embed)
        uv_warp = self.uv_generator_nw(E)
        v = ops.grid_sample3d(st.target_latent_volume, grid=uv_warp, in_layout="cl")
        feat = ops.grid_sample3d(v, theta=warp, out_size=(cfg.D, cfg.S, cfg.S), in_layout="cl", out_layout="hwdc",
                                 want_f32=taps is not None, want_split=True, planes=self.decoder_nw.planes)
        if taps is not None:
            aligned_vol, feat = feat
            taps.update(srt=srt, theta=warp, pose_embed=pose_embed, embed=E, uv_warp=uv_warp, aligned_face=aligned,
                        aligned_volume_hwdc=aligned_vol)
        feat2d = feat.view(1, 1, cfg.S, cfg.S, cfg.D * cfg.C)
        img, deep_f, img_f = self.decoder_nw(feat2d, want_logits=want_logits)
        st_out = SimpleNamespace(pred_target_theta=theta, target_pose_embed=pose_embed, srt=srt)
        return img, deep_f, img_f, st_out


    def make_driver_graph(self, st, mix: bool = True, target_theta: bool = True, slot: int = 0, mix_old: bool = False,
                          smooth_state: Optional[torch.Tensor] = None, smooth_momentum: float = 0.5):
        """Capture one driver frame (all ~240 kernel launches) into a CUDA graph: removes the Python/ctypes launch
        overhead from the per-frame loop.  Returns replay(drv (1,3,H,W) on device) -> img (1,3,H,W) (static buffer).
        `slot` selects the scratch set (ops.set_slot) the captured frame uses; graphs that may replay concurrently on
        different streams need different slots.  smooth_state (3,4): the captured frame also smooths its pose into this
        tensor in place (smooth_pose=True, infer.py:571-581; the state must already be seeded by one eager frame) - frames of
        such a graph are a recurrence and must replay one after the other on one stream."""
        prev = ops.set_slot(slot)
        try:
            return self._make_driver_graph(st, mix, target_theta, mix_old, smooth_state, smooth_momentum)
        finally:
            ops.set_slot(prev)

    def _make_driver_graph(self, st, mix, target_theta, mix_old=False, smooth_state=None, smooth_momentum=0.5):
        s = self.cfg.image_size
        static_in = torch.zeros((1, 3, s, s), dtype=torch.float32, device=self.device)
        kw = dict(mix=mix, target_theta=target_theta, mix_old=mix_old)
        if smooth_state is not None:
            kw.update(smooth_state=smooth_state, smooth_momentum=smooth_momentum, smooth_init=False)
            saved = smooth_state.clone()  # the warm-up and capture passes smooth a dummy frame into the state: put it back after
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.driver_pass(st, static_in, **kw)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.driver_pass(st, static_in, **kw)
            static_out = out[0]
        if smooth_state is not None:
            smooth_state.copy_(saved)

        def replay(drv: torch.Tensor) -> torch.Tensor:
            static_in.copy_(drv, non_blocking=True)
            graph.replay()
            return static_out

        replay.graph = graph
        replay.static_in, replay.static_out, replay.static_state = static_in, static_out, out[3]
        return replay


class DriverPipeline:
    """Keeps `depth` driver frames in flight on one GPU.  Driver frames are independent given the cached source state
    (notebooks/infer.py:511-644 reads only self.* source tensors), and roughly a quarter of a frame is a chain of ~120
    tiny launches (head-pose and expression encoders at 4^2..56^2) that leaves most SMs idle; replaying the captured
    frames of consecutive inputs on alternating streams lets one frame's big decoder kernels fill those holes.
    Frame i's result is produced in order on stream i % depth; nothing is batched and every frame runs the same
    kernels as the single-stream path."""

    def __init__(self, model: "Model", st, depth: int = 2, mix: bool = True, target_theta: bool = True, mix_old: bool = False):
        self.model, self.depth, self.n = model, depth, 0
        self.st = st  # the captured graphs hold raw pointers into this state's tensors: keep it alive as long as they are
        self.slots = []
        for k in range(depth):
            run = model.make_driver_graph(st, mix=mix, target_theta=target_theta, slot=k, mix_old=mix_old)
            self.slots.append(SimpleNamespace(run=run, stream=torch.cuda.Stream(device=model.device),
                                              done=torch.cuda.Event(), busy=False))

    def submit(self, drv: torch.Tensor, host_out: Optional[torch.Tensor] = None, dev_out: Optional[torch.Tensor] = None):
        """Queue one frame (device tensor, or pinned host tensor -> asynchronous H2D on the frame's stream).  If `host_out`
        (pinned) / `dev_out` is given the image is copied into it on the same stream.  Returns a ticket: ticket.done is an
        event recorded when this frame (and its copy-out) has finished (ticket.slot.done is the same event, i.e. always the
        slot's latest frame); slot.run.static_out is valid until the slot's next submit.  A slot's frames are ordered by
        its stream, so a caller may queue further frames on a busy slot — it only has to keep the host_out / dev_out
        buffers of unfinished frames apart."""
        sl = self.slots[self.n % self.depth]
        self.n += 1
        if not drv.is_contiguous():
            drv = drv.contiguous()  # a strided host tensor would otherwise take torch's staged, synchronous copy path
        sl.stream.wait_stream(torch.cuda.current_stream(self.model.device))
        with torch.cuda.stream(sl.stream):
            sl.run.static_in.copy_(drv, non_blocking=True)
            sl.run.graph.replay()
            if host_out is not None:
                host_out.copy_(sl.run.static_out, non_blocking=True)
            if dev_out is not None:
                dev_out.copy_(sl.run.static_out, non_blocking=True)
            sl.done = torch.cuda.Event()
            sl.done.record()
        sl.busy = True
        return SimpleNamespace(done=sl.done, slot=sl, index=self.n - 1)

    def drain(self):
        cur = torch.cuda.current_stream(self.model.device)
        for sl in self.slots:
            cur.wait_stream(sl.stream)
            sl.busy = False


class InferenceWrapper(torch.nn.Module):
    def __init__(self, experiment_name, which_epoch='latest', model_file_name='', use_gpu=True, num_gpus=1,
                 fixed_bounding_box=False, project_dir='./', folder='mp_logs', model_='va',
                 torch_home='', debug=False, print_model=False, print_params=True, args_overwrite={}, state_dict=None,
                 pose_momentum=0.5, rank=0, args_path=None, head_pose_state_dict=None):
        super().__init__()
        if not use_gpu or not torch.cuda.is_available():
            raise RuntimeError("emoportraits_b200.InferenceWrapper needs a CUDA device (B200); there is no CPU path")
        if model_ != 'va':
            raise NotImplementedError(f"model_={model_!r}: only the stage-1 'va' model is implemented")
        self.use_gpu, self.debug, self.num_gpus, self.rank = use_gpu, debug, num_gpus, rank
        args_path = pathlib.Path(project_dir) / folder / experiment_name / 'args.txt' if args_path is None else args_path
        self.args = parse_args(args_path)
        self.args.project_dir = project_dir
        for k, v in (args_overwrite or {}).items():
            setattr(self.args, k, v)
        self.cfg = hot_path_config(self.args)
        self.device = torch.device('cuda', torch.cuda.current_device())

        self.model_checkpoint = pathlib.Path(project_dir) / folder / experiment_name / 'checkpoints' / model_file_name
        if state_dict is None:
            state_dict = torch.load(self.model_checkpoint, map_location='cpu')
        self.model_dict = state_dict
        if head_pose_state_dict is None:
            hp = getattr(self.args, 'head_pose_regressor_path', None)
            if not hp or not pathlib.Path(str(hp)).exists():
                raise FileNotFoundError(f"head pose regressor checkpoint not found ({hp}); pass head_pose_state_dict=")
            head_pose_state_dict = torch.load(hp, map_location='cpu')
        self.model = Model(self.cfg, state_dict, head_pose_state_dict, self.device)
        if rank == 0 and print_params:
            n = sum(v.numel() for v in state_dict.values())
            print(f'Number of parameters/buffers in checkpoint: {n}')

        # face tracking / smoothing state kept for API parity (notebooks/infer.py:159-177)
        self.fixed_bounding_box = fixed_bounding_box
        self.momentum = 0.01
        self.center = None
        self.size = None
        self.pose_momentum = pose_momentum
        self.theta = None
        self.norm_momentum = 0.1
        self.delta_yaw = None
        self.delta_pitch = None
        self.resize_warp = False
        self.use_seg = getattr(self.args, 'use_seg', True)
        self._state = None
        self._pipeline = self._pipeline_key = None
        self._smooth_graph = None
        # external mask networks (separate checkouts in the reference: repos/face_par_off behind model.face_idt, infer.py:410;
        # repos/MODNet, infer.py:140-149): None until the caller plugs them in (masks.FaceParsing(None, net=...), any MODNet module)
        self.face_idt = None
        self.modnet = None

    def get_mask(self, img):
        """notebooks/infer.py:649-684: MODNet matte of img (N,3,h,w) in [0,1]; normalisation and both 'area' resizes on the device."""
        from .masks import modnet_get_mask

        if self.modnet is None:
            raise NotImplementedError("get_mask needs the external MODNet network: set wrapper.modnet = <module>")
        return modnet_get_mask(self.modnet, img.to(self.device))

    def _driver_pipeline(self, mix, target_theta, mix_old):
        key = (bool(mix), bool(target_theta), bool(mix_old))
        if self._pipeline is None or self._pipeline.st is not self._state or self._pipeline_key != key:
            self._pipeline, self._pipeline_key = DriverPipeline(self.model, self._state, depth=4, mix=mix,
                                                                target_theta=target_theta, mix_old=mix_old), key
        return self._pipeline

    def _forward_list(self, frames, mix, target_theta, mix_old):
        """A list of PIL / numpy driver frames: every frame is converted, uploaded and queued on its own (four captured frames in
        flight), its uint8 image is produced and copied back on the frame's stream, so the host-side conversions of frame i + 1
        overlap the device work of frame i; the PIL images are wrapped as their copies land.  Same results as one call per frame."""
        from PIL import Image

        pipe = self._driver_pipeline(mix, target_theta, mix_old)
        n, s = len(frames), self.cfg.image_size
        img = torch.empty((n, 3, s, s), dtype=torch.float32, device=self.device)
        u8 = torch.empty((n, s, s, 3), dtype=torch.uint8, device=self.device)
        host = torch.empty((n, s, s, 3), dtype=torch.uint8).pin_memory()
        keep, done = [], []
        for i, f in enumerate(frames):
            t = self._prep(f)
            keep.append(t)  # read by the slot's stream later: must not return to the allocator (of the current stream) before
            sl = pipe.submit(t, dev_out=img[i:i + 1]).slot
            with torch.cuda.stream(sl.stream):
                ops.image_to_u8(img[i:i + 1], out=u8[i:i + 1])  # clamp(0, 1) + ToPILImage (mul(255).byte()) on the device
                host[i].copy_(u8[i], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            done.append(ev)
            if len(keep) > 2 * pipe.depth and done[i - 2 * pipe.depth].query():
                keep[i - 2 * pipe.depth] = None
        pred_target_img = []
        for i in range(n):
            done[i].synchronize()
            pred_target_img.append(Image.fromarray(host[i].numpy()))
        pipe.drain()
        so = sl.run.static_state
        self.pred_target_theta = so.pred_target_theta.clone()
        self.pred_target_srt = (so.srt[:, :3].clone(), so.srt[:, 3:6].clone(), so.srt[:, 6:9].clone())
        self.target_pose_embed = so.target_pose_embed.clone()
        return pred_target_img, img

    # -- notebooks/infer.py:229-243
    def convert_to_tensor(self, image):
        """PIL image(s) / numpy uint8 (H,W,C) -> fp32 (N,C,H,W) in [0,1] on the device.  uint8 pixels cross the bus as bytes
        (0.79 MB per 512^2 frame instead of 3.1 MB) and become u / 255 in emo_u8_to_image (= transforms.ToTensor)."""
        if isinstance(image, torch.Tensor):
            t = image if image.dim() == 4 else image[None]
            return t.to(self.device).float()
        items = image if isinstance(image, list) else [image]
        arrs = []
        for img in items:
            a = np.asarray(img)
            arrs.append(a[:, :, None] if a.ndim == 2 else a)
        if all(a.dtype == np.uint8 for a in arrs):
            u8 = torch.from_numpy(np.ascontiguousarray(np.stack(arrs))).to(self.device, non_blocking=True)
            return ops.u8_to_image(u8)
        return torch.stack([torch.from_numpy(np.ascontiguousarray(a)).permute(2, 0, 1).float() for a in arrs]).to(self.device)

    def _prep(self, image):
        t = self.convert_to_tensor(image)[:, :3].contiguous()
        s = self.cfg.image_size
        if t.shape[-2:] != (s, s):
            t = ops.resize_bicubic(t, (s, s))  # pre-processing resize, infer.py:401-402 / :554-555
        return t.contiguous().float()

    def forward(self, source_image=None, driver_image=None, source_mask=None, source_mask_add=0, driver_mask=None,
                crop=True, reset_tracking=False, smooth_pose=False, hard_normalize=False, soft_normalize=False,
                delta_yaw=None, delta_pitch=None, cloth=False, thetas_pass='', theta_n=0, target_theta=True, mix=False,
                mix_old=True, c_source_latent_volume=None, c_target_latent_volume=None, custome_target_pose_embed=None,
                custome_target_theta_embed=None, no_grad_infer=True, modnet_mask=False):
        if crop:
            raise NotImplementedError("crop=True needs the external mediapipe face detector (out of scope); pass crop=False "
                                      "with pre-cropped images, as notebooks/E_emo_infer_video.ipynb does")
        if modnet_mask and self.modnet is None:
            raise NotImplementedError("modnet_mask=True needs the external MODNet network: set wrapper.modnet = <module> "
                                      "(repos/MODNet is a separate checkout, not part of the reference tree)")
        # hard_normalize, soft_normalize, cloth, thetas_pass, theta_n are accepted and unused, exactly as in the reference
        # (infer.py:355-357 declares them; nothing in :358-647 reads them).  driver_mask only travels in the reference's
        # data_dict: the expression embedder is called with use_seg=False (:597) and never multiplies by it.
        self.no_grad_infer = no_grad_infer
        self.target_theta, self.mix, self.mix_old = target_theta, mix, mix_old
        if reset_tracking:
            self.center = self.size = self.theta = self.delta_yaw = self.delta_pitch = None
        if delta_yaw is not None:
            self.delta_yaw = delta_yaw
        if delta_pitch is not None:
            self.delta_pitch = delta_pitch

        if source_image is not None:
            src = self._prep(source_image)
            if src.shape[0] != 1:
                raise ValueError("one source image per call (the reference reshapes the latents with view(1, c, d, s, s), infer.py:483)")
            self.source_image = self.source_image_crop = src
            # infer.py:408-426.  face_idt (the external BiSeNet behind masks.FaceParsing) is optional here: without it the face
            # mask is all ones, which is what the reference computes for a frame the parser labels 'face' everywhere
            face_mask = None
            if self.face_idt is not None:
                face_mask = (self.face_idt.forward(src)[0] > 0.6).float()             # :410-411 (trashhold 0.6)
            modnet_matte = self.get_mask(src) if modnet_mask else None               # :413, before the face mask is applied
            if face_mask is not None:
                src = (src * face_mask).float().contiguous()                          # :417
            mask = source_mask.to(self.device).float() if source_mask is not None else face_mask   # :421
            if modnet_mask:
                mask = modnet_matte                                                   # :422
            if mask is not None and source_mask_add:
                mask = mask.clamp_(max=1, min=0)                                      # :423-424
            self.source_img_mask = mask if mask is not None else torch.ones_like(src[:, :1])
            self.source_img = self.source_img_crop_m = src
            st = self.model.source_pass(src, mask=mask, c_source_latent_volume=c_source_latent_volume,
                                        c_target_latent_volume=c_target_latent_volume)
            self._state = st
            self._pipeline = None  # captured for the previous identity
            # cached attributes of the reference wrapper (infer.py:405-507)
            self.idt_embed = st.idt_embed
            self.pred_source_theta = st.pred_source_theta
            self.pred_source_pose_embed = st.pred_source_pose_embed
            self.source_latent_volume = st.source_latent_volume.permute(0, 4, 1, 2, 3)       # NCDHW views
            self.source_xy_warp_resize = st.source_xy_warp_resize
            self.target_latent_volume_1 = st.target_latent_volume_1.permute(0, 4, 1, 2, 3)
            self.target_latent_volume = st.target_latent_volume.permute(0, 4, 1, 2, 3)

        if driver_image is None:
            return None
        if self._state is None:
            raise RuntimeError("forward(driver_image=...) called before a source image was given")
        custom_srt = None
        if custome_target_theta_embed is not None:   # (scale, rotation, translation), each (1,3): infer.py:566-567
            custom_srt = torch.cat([torch.as_tensor(t).float().reshape(1, 3) for t in custome_target_theta_embed], 1)
        per_frame = smooth_pose or custom_srt is not None or custome_target_pose_embed is not None
        if not per_frame and isinstance(driver_image, list) and len(driver_image) > 1 and not isinstance(driver_image[0], torch.Tensor):
            return self._forward_list(driver_image, mix, target_theta, mix_old)
        drv = self._prep(driver_image)
        if not per_frame:
            # captured driver frames (CUDA graphs), four in flight when a list of frames is given (DriverPipeline); same
            # kernels as the eager pass below
            self._driver_pipeline(mix, target_theta, mix_old)
            img = torch.empty_like(drv)
            for i in range(drv.shape[0]):
                sl = self._pipeline.submit(drv[i:i + 1], dev_out=img[i:i + 1]).slot
            self._pipeline.drain()
            so = sl.run.static_state
            self.pred_target_theta = so.pred_target_theta.clone()
            self.pred_target_srt = (so.srt[:, :3].clone(), so.srt[:, 3:6].clone(), so.srt[:, 6:9].clone())
            self.target_pose_embed = so.target_pose_embed.clone()
        elif smooth_pose and custom_srt is None and custome_target_pose_embed is None:
            # pose smoothing is a recurrence over the frames (infer.py:571-581): the first frame seeds self.theta eagerly, every
            # further frame replays a captured graph that smooths into the same tensor, one after the other on this stream
            imgs = []
            for i in range(drv.shape[0]):
                if self.theta is None:
                    self.theta = torch.zeros((3, 4), dtype=torch.float32, device=self.device)
                    im, _, _, so = self.model.driver_pass(self._state, drv[i:i + 1].contiguous(), mix=mix, target_theta=target_theta,
                                                          mix_old=mix_old, smooth_state=self.theta, smooth_momentum=self.pose_momentum,
                                                          smooth_init=True)
                    self._smooth_graph = None
                else:
                    key = (bool(mix), bool(target_theta), bool(mix_old), float(self.pose_momentum))
                    g = self._smooth_graph
                    if g is None or g.st is not self._state or g.theta is not self.theta or g.key != key:
                        run = self.model.make_driver_graph(self._state, mix=mix, target_theta=target_theta, mix_old=mix_old,
                                                           smooth_state=self.theta, smooth_momentum=self.pose_momentum)
                        g = self._smooth_graph = SimpleNamespace(run=run, st=self._state, theta=self.theta, key=key)
                    im, so = g.run(drv[i:i + 1]).clone(), g.run.static_state
                imgs.append(im)
            img = imgs[0] if len(imgs) == 1 else torch.cat(imgs)
            self.pred_target_theta = so.pred_target_theta.clone()
            self.pred_target_srt = (so.srt[:, :3].clone(), so.srt[:, 3:6].clone(), so.srt[:, 6:9].clone())
            self.target_pose_embed = so.target_pose_embed.clone()
        else:
            imgs = []
            for i in range(drv.shape[0]):
                kw = {}
                if smooth_pose:
                    # self.theta is the smoothed (3,4) pose carried from frame to frame (infer.py:571-581); the device
                    # kernel seeds it with the first frame's pose and updates it in place
                    init = self.theta is None
                    if init:
                        self.theta = torch.zeros((3, 4), dtype=torch.float32, device=self.device)
                    kw = dict(smooth_state=self.theta, smooth_momentum=self.pose_momentum, smooth_init=init)
                im, deep_f, img_f, so = self.model.driver_pass(self._state, drv[i:i + 1].contiguous(), mix=mix,
                                                               target_theta=target_theta, mix_old=mix_old, custom_srt=custom_srt,
                                                               custom_pose_embed=custome_target_pose_embed, **kw)
                imgs.append(im)
            img = imgs[0] if len(imgs) == 1 else torch.cat(imgs)
            self.pred_target_theta = so.pred_target_theta
            self.pred_target_srt = (so.srt[:, :3], so.srt[:, 3:6], so.srt[:, 6:9])
            self.target_pose_embed = so.target_pose_embed
        from PIL import Image

        host = ops.image_to_u8(img.detach().contiguous()).cpu().numpy()  # clamp(0, 1) + ToPILImage (mul(255).byte()) on the device
        pred_target_img = [Image.fromarray(h) for h in host]
        return pred_target_img, img
