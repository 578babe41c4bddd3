"""CPU what-if: how much do the operand-split schemes of the convolution kernel move the IMAGE when they are applied to
the per-frame embedding / warp networks (head-pose regressor, expression encoder, predict_embed, uv warp generator)?

    python tests/analysis/h2_network_emulation.py [size=256]

The oracle restatement (oracle/restatement.py, torch fp32 on the CPU) is run once as is and once per scheme with every
F.conv2d / F.conv3d of those networks replaced by an emulation of the kernel's arithmetic on OPERANDS: inputs and weights
split into planes (bf16 x2, bf16 x3, fp16 x2 with the power-of-two scales of ops.H2), the kernel's products summed in
fp64 (so accumulation is ideal: this isolates the operand error; the real kernel adds ~1e-6 of truncating accumulation).
The decoder and the samplers stay exact.  Printed: max-abs change of pose embedding, uv warp and image.
This is analysis tooling for DESIGN.md section 7 (item 1c); it is not part of the product or of the test suite."""
import pathlib
import sys

import torch
import torch.nn.functional as F

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

T3 = [(0, 0), (0, 1), (1, 0)]
T6 = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]
SCHEMES = {
    "bf16 x2 (3 MMAs)": (torch.bfloat16, 2, T3, 1.0, 1.0),
    "bf16 x3 (6 MMAs)": (torch.bfloat16, 3, T6, 1.0, 1.0),
    "fp16 x2 scaled (3 MMAs, h2)": (torch.float16, 2, T3, 16.0, 256.0),
}


def planes(x, dt, n, scale):
    out, r = [], (x * scale).float()
    for _ in range(n):
        p = r.to(dt)
        out.append(p.double() / scale)
        r = r - p.float()
    return out


class Emulate:
    def __init__(self, scheme):
        self.dt, self.n, self.terms, self.sa, self.sw = SCHEMES[scheme]
        self.on = False

    def conv(self, fn):
        def wrapped(x, w, b=None, *a, **k):
            if not self.on or x.shape[1] < 32:   # RGB stems run as exact fp32 SIMT in the product too
                return fn(x, w, b, *a, **k)
            xp, wp = planes(x, self.dt, self.n, self.sa), planes(w, self.dt, self.n, self.sw)
            y = sum(fn(xp[i], wp[j], None, *a, **k) for i, j in self.terms)
            if b is not None:
                y = y + b.double().view(1, -1, *([1] * (y.dim() - 2)))
            return y.float()
        return wrapped


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    from emoportraits_b200.checkpoint import synthetic_head_pose_state_dict, synthetic_state_dict
    from emoportraits_b200.config import shipped_config
    from oracle import frames as FR
    from oracle import restatement as R

    cfg = shipped_config(size)
    sd, hsd = synthetic_state_dict(cfg, 0), synthetic_head_pose_state_dict(0)
    ocfg = R.config_from_state_dict(sd, size)
    src, drv = FR.frame(size, 10, "smooth"), FR.frame(size, 11, "smooth")
    conv2d, conv3d = F.conv2d, F.conv3d
    with torch.no_grad():
        st = R.source_pass(sd, hsd, src, ocfg)
        taps0 = {}
        img0 = R.driver_pass(sd, hsd, st, drv, ocfg, taps0)
        for name in SCHEMES:
            em = Emulate(name)
            F.conv2d, F.conv3d = em.conv(conv2d), em.conv(conv3d)
            orig_dec = R.decoder
            try:
                def dec(*a, **k):            # the decoder stays exact
                    em.on = False
                    try:
                        return orig_dec(*a, **k)
                    finally:
                        em.on = True
                R.decoder = dec
                em.on = True
                taps = {}
                img = R.driver_pass(sd, hsd, st, drv, ocfg, taps)
            finally:
                em.on = False
                F.conv2d, F.conv3d = conv2d, conv3d
                R.decoder = orig_dec
            print(f"{name:30s} pose_embed {(taps['pose_embed'] - taps0['pose_embed']).abs().max().item():.2e}  "
                  f"uv_warp {(taps['uv_warp'] - taps0['uv_warp']).abs().max().item():.2e}  "
                  f"image {(img - img0).abs().max().item():.2e}")


if __name__ == "__main__":
    main()
