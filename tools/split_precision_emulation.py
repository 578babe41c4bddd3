"""CPU emulation of the operand-split schemes of the convolution kernel (ideal fp64 accumulation: isolates the OPERAND error).

    python tools/split_precision_emulation.py

Prints max / rms relative error of a K = 1728 dot product (a 3x3x3 conv over 64 channels, post-GroupNorm+ReLU activations,
1/sqrt(K) weights) for
    bf16 x2  hi*hi + hi*lo + lo*hi                 3 MMAs   (decoder)
    bf16 x3  six products                          6 MMAs   (embedding / warp / source networks)
    fp16 x2  with power-of-two plane scales        3 MMAs   ("h2" mode, ops.H2: candidate replacement of bf16 x3)
Numbers quoted in DESIGN.md section 7 (next round, fp16 two-plane mode)."""
import math

import torch


def planes(x, dt, n, scale=1.0):
    out, r = [], (x * scale).float()
    for _ in range(n):
        p = r.to(dt)
        out.append(p.double() / scale)
        r = r - p.float()
    return out


def prod(ap, wp, terms):
    return sum(ap[i] @ wp[j].T for i, j in terms)


T3 = [(0, 0), (0, 1), (1, 0)]
T6 = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]


def main():
    torch.manual_seed(0)
    K, M, N = 1728, 512, 64
    for tag, amp in (("activations O(1)", 1.0), ("activations O(1e-3)", 1e-3)):
        a = (torch.relu(torch.randn(M, K, dtype=torch.float64)) * amp).float()
        w = (torch.randn(N, K, dtype=torch.float64) / math.sqrt(K)).float()
        exact = a.double() @ w.double().T

        def rel(y):
            return ((y - exact).abs().max() / exact.abs().max()).item(), ((y - exact).pow(2).mean().sqrt() / exact.pow(2).mean().sqrt()).item()

        print(tag)
        print("  bf16 x2 (3 MMAs)                max %.2e rms %.2e" % rel(prod(planes(a, torch.bfloat16, 2), planes(w, torch.bfloat16, 2), T3)))
        print("  bf16 x3 (6 MMAs)                max %.2e rms %.2e" % rel(prod(planes(a, torch.bfloat16, 3), planes(w, torch.bfloat16, 3), T6)))
        print("  fp16 x2 (3 MMAs) unscaled       max %.2e rms %.2e" % rel(prod(planes(a, torch.float16, 2), planes(w, torch.float16, 2), T3)))
        print("  fp16 x2 (3 MMAs) a x16, w x256  max %.2e rms %.2e" % rel(prod(planes(a, torch.float16, 2, 16.0), planes(w, torch.float16, 2, 256.0), T3)))


if __name__ == "__main__":
    main()
