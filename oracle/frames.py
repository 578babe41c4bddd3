"""TEST INFRASTRUCTURE ONLY.  Deterministic synthetic input frames shared by the golden generator and the tests.

noise_frame : uint8 = RandomState(seed).rand(H, W, 3) * 255   (BASELINE.md §4 — white noise; hypersensitive to sub-pixel
              sampling positions: a 1-ulp change of the reference's own fp32 4x4 inverse moves its image by 1.2e-3)
smooth_frame: the same generator at 16 x 16, upsampled by pixel replication and box-blurred (pure numpy, float64, then
              quantised) — image-like spectrum, gradients ~10x smaller.
"""
import numpy as np
import torch


def noise_array(size: int, seed: int) -> np.ndarray:
    return (np.random.RandomState(seed).rand(size, size, 3) * 255).astype(np.uint8)


def smooth_array(size: int, seed: int) -> np.ndarray:
    base = np.random.RandomState(seed).rand(16, 16, 3)
    rep = size // 16
    a = np.repeat(np.repeat(base, rep, axis=0), rep, axis=1)  # (size, size, 3) float64
    k = rep  # box blur of one cell width, twice (separable), edge-replicated
    for _ in range(2):
        for axis in (0, 1):
            pad = [(0, 0)] * 3
            pad[axis] = (k // 2, k - 1 - k // 2)
            ap = np.pad(a, pad, mode="edge")
            c = np.cumsum(ap, axis=axis)
            c = np.concatenate([np.zeros_like(np.take(c, [0], axis=axis)), c], axis=axis)
            hi = np.take(c, np.arange(k, k + size), axis=axis)
            lo = np.take(c, np.arange(0, size), axis=axis)
            a = (hi - lo) / k
    return np.clip(np.round(a * 255), 0, 255).astype(np.uint8)


def to_tensor(a: np.ndarray) -> torch.Tensor:
    """uint8 HWC -> float (1,3,H,W) in [0,1], contiguous (what ToTensor does, notebooks/infer.py:229-243)"""
    return torch.from_numpy(a).permute(2, 0, 1)[None].float().div(255).contiguous()


def frame(size: int, seed: int, kind: str = "noise") -> torch.Tensor:
    return to_tensor(noise_array(size, seed) if kind == "noise" else smooth_array(size, seed))


def pil(size: int, seed: int, kind: str = "noise"):
    from PIL import Image

    return Image.fromarray(noise_array(size, seed) if kind == "noise" else smooth_array(size, seed))
