"""Time single convolution shape classes of the UNet through the C ABI (development aid).
   python tools/bench_conv.py            # env CDDPM_CONV_PAIR / CDDPM_CONV_V2 / CDDPM_CONV_DEBUG select variants"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]
from cddpm import ops  # noqa: E402

SHAPES = [  # (B, H, W, cins, k, cout, residual)
    (32, 96, 96, [128], 3, 128, False),
    (32, 96, 96, [128], 3, 128, True),
    (32, 96, 96, [256], 3, 256, False),
    (32, 48, 48, [256], 3, 256, True),
    (32, 24, 24, [256], 3, 256, False),
    (32, 96, 96, [128, 128], 3, 128, False),
]


def main():
    global SHAPES
    if "--sweep" in sys.argv:  # batch sweep of one class: the intercept of time vs batch is the per-launch fixed cost
        SHAPES = [(b, 96, 96, [128], 3, 128, False) for b in (2, 4, 8, 16, 32, 64)] + \
                 [(b, 48, 48, [256], 3, 256, False) for b in (8, 16, 32, 64, 128)]
    if "--level24" in sys.argv:  # the 24 x 24 classes of a forward (CDDPM_CONV_IL=0/1: plain / image-interleaved tiles)
        SHAPES = [(b, 24, 24, cins, 3, 256, False) for b in (32, 64, 150) for cins in ([256], [256, 256])]
    dt = torch.float16
    tag = " ".join(f"{k}={os.environ[k]}" for k in ("CDDPM_CONV_PAIR", "CDDPM_CONV_V2", "CDDPM_CONV_DEBUG", "CDDPM_CONV_IL") if k in os.environ)
    print("variant:", tag or "default", flush=True)
    once = "--once" in sys.argv  # one launch per shape (for ncu captures)
    for B, H, W, cins, k, cout, res in SHAPES:
        srcs = [torch.randn(B, H, W, c, device="cuda", dtype=dt) for c in cins]
        w = torch.randn(cout, sum(cins), k, k, device="cuda") / (sum(cins) * k * k) ** 0.5
        wp = ops.pack_conv_weight(w, cins, dt)
        bias = torch.randn(cout, device="cuda")
        r = torch.randn(B, H, W, cout, device="cuda", dtype=dt) if res else None
        for _ in range(0 if once else 3):
            ops.conv_igemm(srcs, [k * k] * len(cins), wp, bias, r)
        torch.cuda.synchronize()
        n = 1 if once else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            ops.conv_igemm(srcs, [k * k] * len(cins), wp, bias, r)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        fl = 2.0 * B * H * W * cout * sum(cins) * k * k
        print(f"B{B} {H}x{W} {'+'.join(map(str, cins))}->{cout} k{k} res={int(res)}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
