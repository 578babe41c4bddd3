"""Caller-side video loop (emoportraits_b200/video.py = notebooks/E_emo_infer_video.ipynb cells 40-51) on synthetic frames:
chunked, pipelined driver frames == one-at-a-time wrapper calls; compositing; side-by-side video written and read back."""
import pathlib

import pytest
import torch

from oracle import frames as FR

pytestmark = pytest.mark.gpu
GOLD = pathlib.Path(__file__).parent / "golden"
SIZE = 256


def test_drive_image_with_video(tmp_path):
    from emoportraits_b200 import video
    from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
    from emoportraits_b200.config import shipped_config
    from emoportraits_b200.infer import InferenceWrapper

    cfg = shipped_config(SIZE)
    w = InferenceWrapper(experiment_name="x", project_dir=str(tmp_path), args_path=GOLD / f"args_{SIZE}.txt", state_dict=synthetic_state_dict(cfg, 0),
                         head_pose_state_dict=synthetic_head_pose_state_dict(0), print_params=False)
    src = FR.pil(SIZE, 41, "smooth")
    drv = [FR.pil(SIZE, 50 + i, "smooth") for i in range(5)]
    outs, crops, fr = video.drive_image_with_video(w, src, drv, chunk=3, crop_box=(32, 32, 224, 224))
    assert len(outs) == len(crops) == 5 and outs[0].size == crops[0].size == (SIZE, SIZE)
    kw = dict(crop=False, mix=True, mix_old=False)
    one = [w.forward(None, d, **kw)[0][0] for d in drv]
    import numpy as np
    for a, b in zip(outs, one):          # uint8 images: at most one quantisation step apart (run-to-run statistics noise)
        assert np.abs(np.asarray(a).astype(int) - np.asarray(b).astype(int)).max() <= 1
    # compositing over a background with a caller-supplied matte (the notebook's face-parsing / MODNet mask is an input here)
    bg = torch.rand(3, SIZE, SIZE)
    mask_fn = lambda img: torch.full_like(img[:, :1], 0.9)
    comp, _, _ = video.drive_image_with_video(w, src, drv[:2], bg=bg, mask_fn=mask_fn, chunk=2)
    _, img = w.forward(None, drv[:2], **kw)
    want = (0.9 ** 8) * img.cpu() + (1 - 0.9 ** 8) * bg[None]
    got = torch.stack([torch.from_numpy(np.asarray(c)).permute(2, 0, 1) for c in comp]).float() / 255
    assert (got - want.clamp(0, 1)).abs().max().item() < 2.5 / 255
    cv2 = pytest.importorskip("cv2")
    path = tmp_path / "out.mp4"
    video.make_video(src, drv, outs, str(path), fps=25.0, size=SIZE)
    back = video.get_video_frames_as_images(str(path), size=SIZE)
    assert len(back) == 5 and back[0].size == (SIZE, SIZE)
