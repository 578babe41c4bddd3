"""TEST INFRASTRUCTURE ONLY (oracle).  Seeded stand-ins for the EXTERNAL mask networks (BiSeNet face parsing:
repos/face_par_off, MODNet matting: repos/MODNet - separate checkouts that are not part of the reference tree).  They only
have the call signatures the reference uses (networks/volumetric_avatar/face_parcing.py:59 `self.net(x)[0]`,
notebooks/infer.py:679 `_, _, matte = self.modnet(im, True)`) and produce spatially varied, deterministic outputs, so that the
pre/post-processing AROUND the networks (the part that lives in the reference tree) can be pinned: oracle/make_golden_masks.py
runs the unmodified reference code with these networks plugged in, tests/ run the device path with the same networks."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class StubBiSeNet(nn.Module):
    """19-class logits at the input resolution from a fixed smooth random field + a weak dependence on the image."""

    def __init__(self, n_classes: int = 19, *a, **k):
        super().__init__()
        g = torch.Generator().manual_seed(1234)
        self.n_classes = n_classes
        self.register_buffer("field", torch.randn(1, n_classes, 12, 12, generator=g))
        self.register_buffer("w", torch.randn(n_classes, 3, 1, 1, generator=g) * 0.05)

    def forward(self, x):
        h, w = x.shape[2:]
        base = F.interpolate(self.field.to(x), size=(h, w), mode="bicubic", align_corners=False)
        return (base + F.conv2d(x, self.w.to(x)),)


class StubMODNet(nn.Module):
    """matte in (0, 1) at the input resolution: sigmoid of a smooth field + a weak dependence on the image."""

    def __init__(self, *a, **k):
        super().__init__()
        g = torch.Generator().manual_seed(4321)
        self.register_buffer("field", torch.randn(1, 1, 9, 9, generator=g) * 2)
        self.register_buffer("w", torch.randn(1, 3, 1, 1, generator=g) * 0.2)

    def forward(self, x, inference=True):
        h, w = x.shape[2:]
        base = F.interpolate(self.field.to(x), size=(h, w), mode="bicubic", align_corners=False)
        return None, None, torch.sigmoid(base + F.conv2d(x, self.w.to(x)))
