"""GPU parity of the full hot path (through emoportraits_b200.infer) against
  (a) golden fixtures produced by the UNMODIFIED reference (tests/golden, oracle/make_golden.py), and
  (b) the CPU oracle restatement run live on fresh inputs.

Tolerance (BASELINE.json north_star): 1e-3 max-abs per pixel on the fp32 image; stage taps are checked as well so the
sigmoid cannot hide an error.

White-noise frames (the BASELINE-specified synthetic input) make the REFERENCE itself ill-conditioned: its fp32 LU
4x4 inverse (torch.inverse, notebooks/infer.py:443 / expression_embedder.py:168) differs from the exactly rounded
inverse by 1 ulp, and that alone moves the reference's own image by 1.2e-3 (tests/test_oracle_golden.py::
test_reference_noise_floor_of_white_noise_frames measures it on CPU).  So three checks:
  * noise frames, reference pose matrices injected (bit-identical 4x4 algebra)      -> image within 1e-3
  * smooth (image-like) frames, everything on device                                -> image within 1e-3
  * noise frames, everything on device                                              -> within 5e-3 (reference-noise-limited)
"""
import pathlib

import pytest
import torch

from oracle import frames as FR

pytestmark = pytest.mark.gpu
GOLD = pathlib.Path(__file__).parent / "golden"
IMG_TOL = 1e-3
# white-noise frames, pose algebra on the device: bounded by the REFERENCE's own sensitivity to a 1-ulp change of its fp32 LU
# inverse (measured per size by tests/test_oracle_golden.py::test_reference_noise_floor_of_white_noise_frames; see module doc)
# measured here (round 2): 1.6e-3 at 256^2, 3.8e-3 at 512^2; the reference's own floor from that one-ulp change: 1.2e-3 / see the
# CPU test's printout for 512^2.  Bound = measured maximum x 1.2 + the 1e-4 run-to-run spread.
NOISE_DEVICE_TOL = {256: 2.0e-3, 512: 4.6e-3}


def _sub_err(got, ref):
    if isinstance(ref, torch.Tensor):
        return (got.detach().float().cpu() - ref).abs().max().item()
    vals, stride = ref
    g = got.detach().float().cpu().reshape(-1)[::stride]
    return (g - vals).abs().max().item()


@pytest.fixture(scope="module", params=[256, 512])
def setup(request):
    size = request.param
    from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
    from emoportraits_b200.config import shipped_config
    from emoportraits_b200.infer import Model

    cfg = shipped_config(size)
    sd, hsd = synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0)
    model = Model(cfg, sd, hsd, "cuda")
    gold = torch.load(GOLD / f"va{size}_seed0.pt", weights_only=False)
    return size, cfg, model, gold


def _scale(ref):
    return float((ref if isinstance(ref, torch.Tensor) else ref[0]).abs().max())


# Stage taps, asserted at BOTH sizes so that the sigmoid cannot hide an error: bound = TAP_TOL[key] * max(1, max|reference tap|).
# The numbers are the stage-isolated bounds of tests/test_stage_parity_gpu.py widened for the chain upstream of each tap
# (an embedding error e moves the warp by ~10 e and the sampled volumes by ~100 e; target_latent_volume_1 is the raw encoder
# volume sampled through xy_warp, whose 5e-5 error in normalised coordinates is 1.6e-3 voxel of a volume that changes by its
# own magnitude from voxel to voxel: measured 2e-4 .. 8e-4 of the volume's maximum).  Measured values are printed, written to
# gpurun_out/parity_<case>.txt and committed per round under profiles/.
TAP_TOL = {"idt_embed": 5e-5, "source_theta": 2.5e-7, "source_pose_embed": 5e-5, "xy_warp": 1e-4, "source_latent_volume": 1e-4,
           "target_latent_volume_1": 1.5e-3, "target_latent_volume": 5e-4, "theta": 2.5e-7, "pose_embed": 5e-5, "uv_warp": 1e-4,
           "aligned_feat2d": 5e-4, "dec_feat": 5e-4, "logits": 1.5e-3, "img": IMG_TOL}
# theta: the device pose algebra is within two fp32 ulps of the reference's (CUDA sinf/cosf vs the host libm); on white-noise
# frames the reference amplifies ONE such ulp to 1.2e-3 / 1.5e-3 in the image (tests/test_oracle_golden.py), which is the whole
# white-noise exception: with the reference's matrices injected the same frames are within 1e-3.
# logits: pre-sigmoid, relative to their maximum (|logit| <= 4.3: an error of 1.5e-3 x 4.3 at sigmoid'(0) = 1/4 is 1.6e-3 in the
# image, so the image bound is the sharper one near 0 and this one guards the saturated range)
# taps that sit downstream of the reference's fp32 LU inverse of the pose matrix (the 1-ulp-fragile step, module docstring):
# not asserted in the all-on-device white-noise case
POSE_NOISE_LIMITED = ("source_pose_embed", "xy_warp", "target_latent_volume_1", "target_latent_volume", "pose_embed", "uv_warp",
                      "aligned_feat2d", "dec_feat", "logits", "img")


def _assert_taps(errs, scales, skip=()):
    bad = []
    for k, v in errs.items():
        name = k.split(".")[-1]
        if name in skip:
            continue
        bound = TAP_TOL[name] * (1.0 if name == "img" else max(1.0, scales[k]))
        if not v < bound:
            bad.append((k, v, bound))
    assert not bad, bad


def _run_case(size, cfg, model, case, inject_pose: bool):
    kind = case["kind"]
    s = case["source"]
    po = (s["pred_source_theta"], s["inv_warp"], s["align2d"]) if inject_pose else None
    st = model.source_pass(FR.frame(size, case["src_seed"], kind).cuda(), pose_override=po)
    errs, scales = {}, {}

    def rec(key, got, ref):
        errs[key], scales[key] = _sub_err(got, ref), _scale(ref)

    rec("idt_embed", st.idt_embed, s["idt_embed"])
    rec("source_theta", st.pred_source_theta, s["pred_source_theta"])
    rec("source_pose_embed", st.pred_source_pose_embed, s["pred_source_pose_embed"])
    rec("xy_warp", st.source_xy_warp_resize, s["xy_warp"])
    ncdhw = lambda t: t.permute(0, 4, 1, 2, 3).contiguous()
    rec("source_latent_volume", ncdhw(st.source_latent_volume), s["source_latent_volume"])
    rec("target_latent_volume_1", ncdhw(st.target_latent_volume_1), s["target_latent_volume_1"])
    rec("target_latent_volume", ncdhw(st.target_latent_volume), s["target_latent_volume"])
    for fr in case["frames"]:
        drv = FR.frame(size, fr["seed"], kind).cuda()
        th = fr["pred_target_theta"]
        th4 = torch.cat([th[:, :3], torch.tensor([[[0., 0., 0., 1.]]])], 1)
        po = (th4, th[:, :3], fr["align2d"]) if inject_pose else None
        taps = {}
        logits, _, _, so = model.driver_pass(st, drv, mix=True, taps=taps, want_logits=True, pose_override=po)
        img, deep_f, img_f, so = model.driver_pass(st, drv, mix=True, pose_override=po)
        k = f"f{fr['seed']}."
        rec(k + "theta", so.pred_target_theta[:, :3], th[:, :3])
        rec(k + "pose_embed", so.target_pose_embed, fr["target_pose_embed"])
        rec(k + "uv_warp", taps["uv_warp"], fr["uv_warp"])
        # reference (b, c*D+d, h, w)  <->  ours (h, w, d, c)
        av = taps["aligned_volume_hwdc"].view(1, cfg.S, cfg.S, cfg.D, cfg.C).permute(0, 4, 3, 1, 2).reshape(1, cfg.C * cfg.D, cfg.S, cfg.S)
        rec(k + "aligned_feat2d", av.contiguous(), fr["aligned_feat2d"])
        rec(k + "dec_feat", deep_f[:, 0].permute(0, 3, 1, 2).contiguous(), fr["dec_feat"])
        rec(k + "logits", logits, fr["logits"])
        rec(k + "img", img, fr["img"])
    tag = f"{size}_{kind}_{'refpose' if inject_pose else 'device'}"
    print(f"\n[parity vs reference golden @ {tag}] " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    out = pathlib.Path("gpurun_out"); out.mkdir(exist_ok=True)
    (out / f"parity_{tag}.txt").write_text("\n".join(f"{k} err {v:.3e} ref_max {scales[k]:.3e}" for k, v in errs.items()) + "\n")
    return errs, scales


def _case(gold, kind):
    return next(c for c in gold["cases"] if c["kind"] == kind)


def test_noise_frames_with_reference_pose_matrices(setup):
    size, cfg, model, gold = setup
    errs, scales = _run_case(size, cfg, model, _case(gold, "noise"), inject_pose=True)
    _assert_taps(errs, scales)


def test_smooth_frames_all_on_device(setup):
    size, cfg, model, gold = setup
    errs, scales = _run_case(size, cfg, model, _case(gold, "smooth"), inject_pose=False)
    _assert_taps(errs, scales)


def test_noise_frames_all_on_device_reference_noise_limited(setup):
    size, cfg, model, gold = setup
    errs, scales = _run_case(size, cfg, model, _case(gold, "noise"), inject_pose=False)
    _assert_taps(errs, scales, skip=POSE_NOISE_LIMITED)   # everything upstream of the fragile inverse holds its bound
    for k, v in errs.items():
        if k.endswith("img"):
            assert v < NOISE_DEVICE_TOL[size], (k, v)


def test_driver_pass_matches_cpu_oracle_on_fresh_inputs(setup):
    """same check against the oracle restatement run live on the CPU, on inputs NOT in the fixtures (smooth frames)"""
    size, cfg, model, gold = setup
    if size != 256:
        pytest.skip("CPU oracle at 512 is covered by the fixtures; keep the GPU suite short")
    from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
    from oracle import restatement as R

    sd, hsd = synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0)
    ocfg = R.config_from_state_dict(sd, size)
    src, drv = FR.frame(size, 31, "smooth"), FR.frame(size, 32, "smooth")
    with torch.no_grad():
        ost = R.source_pass(sd, hsd, src, ocfg)
        oimg = R.driver_pass(sd, hsd, ost, drv, ocfg)
    st = model.source_pass(src.cuda())
    img, _, _, _ = model.driver_pass(st, drv.cuda(), mix=True)
    err = (img.cpu() - oimg).abs().max().item()
    print(f"\n[parity vs CPU oracle @ {size}, fresh smooth frames] img max-abs err = {err:.3e}")
    assert err < IMG_TOL


def test_frames_in_flight_match_sequential(setup):
    """DriverPipeline (CUDA graphs of consecutive frames replayed on alternating streams, separate scratch slots) returns,
    frame by frame, what the plain one-after-the-other driver pass returns - bit for bit: every GroupNorm statistic is summed in
    a fixed order or in fp64 (sums of fp32 terms are exact there), so two runs of a frame are identical (tools/repro_check.py:
    0.0 at both sizes, eager vs eager and four frames in flight vs eager).  The last step, fp64 atomics across CTAs, can in
    principle move the final bit of a double and flip one fp32 coefficient (~1e-9 per coefficient), which the warp network would
    amplify to ~1e-4: the hard bound stays at 5e-4 (a scratch-sharing bug between in-flight frames shows up as 1e-2+) and at
    least four of the five frames must be bit-identical."""
    size, cfg, model, gold = setup
    from emoportraits_b200.infer import DriverPipeline

    st = model.source_pass(FR.frame(size, 41, "smooth").cuda())
    drv = [FR.frame(size, 50 + i, "smooth").cuda() for i in range(5)]
    want = [model.driver_pass(st, d, mix=True)[0].clone() for d in drv]
    pipe = DriverPipeline(model, st, depth=2, mix=True)
    hosts = [torch.empty((1, 3, size, size)).pin_memory() for _ in drv]
    for d, h in zip(drv, hosts):
        pipe.submit(d, host_out=h)
    pipe.drain()
    torch.cuda.synchronize()
    errs = [(h - w.cpu()).abs().max().item() for h, w in zip(hosts, want)]
    if max(errs) >= 5e-4:  # diagnostics: is it the pipeline or the eager pass, and does it persist?
        again = [torch.empty((1, 3, size, size)).pin_memory() for _ in drv]
        for d, h in zip(drv, again):
            pipe.submit(d, host_out=h)
        pipe.drain()
        torch.cuda.synchronize()
        errs2 = [(h - w.cpu()).abs().max().item() for h, w in zip(again, want)]
        want2 = [model.driver_pass(st, d, mix=True)[0] for d in drv]
        errs3 = [(a - b).abs().max().item() for a, b in zip(want, want2)]
        errs4 = [(h - w.cpu()).abs().max().item() for h, w in zip(hosts, want2)]
        raise AssertionError(f"pipeline vs eager {errs}; second pipeline run {errs2}; eager vs eager {errs3}; first pipeline vs second eager {errs4}")
    assert sum(e == 0.0 for e in errs) >= len(errs) - 1, errs
    assert (want[0] - want[1]).abs().max().item() > 1e-2  # the frames do differ


def test_inference_wrapper_api(setup, tmp_path):
    """drop-in API: same ctor/forward signature and return types as notebooks/infer.py InferenceWrapper"""
    size, cfg, model, gold = setup
    if size != 256:
        pytest.skip("API test runs once")
    from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
    from emoportraits_b200.infer import InferenceWrapper

    case = _case(gold, "smooth")
    exp = tmp_path / "logs" / "exp" / "checkpoints"
    exp.mkdir(parents=True)
    (tmp_path / "logs" / "exp" / "args.txt").write_text((GOLD / f"args_{size}.txt").read_text())
    torch.save(synthetic_state_dict(cfg, 0), exp / "000_model.pth")
    w = InferenceWrapper(experiment_name="exp", model_file_name="000_model.pth", project_dir=str(tmp_path), folder="logs",
                         print_params=False, head_pose_state_dict=synthetic_head_pose_state_dict(0))
    src, drv = FR.pil(size, case["src_seed"], "smooth"), FR.pil(size, case["frames"][0]["seed"], "smooth")
    pil, img = w.forward(src, drv, crop=False, mix=True, mix_old=False)
    assert isinstance(pil, list) and pil[0].size == (size, size) and img.shape == (1, 3, size, size) and img.is_cuda
    assert _sub_err(img, case["frames"][0]["img"]) < IMG_TOL
    pil2, img2 = w.forward(None, drv, crop=False, mix=True, mix_old=False)
    assert (img - img2).abs().max().item() < 5e-4  # normally 0.0: see test_frames_in_flight_match_sequential
    # a list of driver frames goes through the captured four-in-flight pipeline; same images, in order
    drv_b = FR.pil(size, case["frames"][0]["seed"] + 7, "smooth")
    pil3, img3 = w.forward(None, [drv, drv_b, drv], crop=False, mix=True, mix_old=False)
    assert len(pil3) == 3 and img3.shape == (3, 3, size, size)
    assert (img3[0:1] - img).abs().max().item() < 5e-4 and (img3[2:3] - img).abs().max().item() < 5e-4
    assert (img3[1:2] - img).abs().max().item() > 1e-2
    assert w.forward(src, None, crop=False) is None
    with pytest.raises(NotImplementedError):
        w.forward(src, drv)  # crop=True needs the external face detector
