"""Per-shape timing of the convolution backward kernels (development aid): weight gradient (conv_wgrad_kernel) and
data gradient (forward kernel over the transposed panel) for the UNet's shape classes at batch B.
   python tools/bench_wgrad.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]

from cddpm import ops  # noqa: E402

SHAPES = [  # (H, cins, cout, ksize)
    (96, [128], 128, 3), (96, [256], 256, 3), (96, [256, 128], 128, 3), (96, [128, 128], 128, 3),
    (48, [256], 256, 3), (48, [256, 256], 256, 3), (48, [128], 256, 3),
    (24, [256], 256, 3), (24, [256, 256], 256, 3), (24, [256], 768, 1), (48, [256, 256], 256, 1),
]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dt = torch.bfloat16
    for H, cins, cout, ks in SHAPES:
        srcs = [torch.randn(B, H, H, c, device="cuda").to(dt) for c in cins]
        dy = torch.randn(B, H, H, cout, device="cuda").to(dt)
        taps = [ks * ks] * len(cins)
        flops = 2.0 * B * H * H * cout * sum(cins) * ks * ks
        for _ in range(2):
            ops.conv_wgrad(srcs, taps, dy)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        dws = [torch.zeros(cout, sum(t * c for t, c in zip(taps, cins)), device="cuda") for _ in range(n)]
        torch.cuda.synchronize()
        from cddpm._lib import check, current_stream, fmt_of, int_array, lib, ptr, ptr_array
        e0.record()
        for i in range(n):
            check(lib().cddpm_conv_wgrad(len(srcs), ptr_array([ptr(s) for s in srcs]), int_array(cins), int_array(taps),
                                         int_array([0] * len(srcs)), ptr(dy), B, H, H, cout, ptr(dws[i]), fmt_of(dt),
                                         current_stream()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"wgrad {'+'.join(map(str, cins)):>8s}->{cout:<4d} k{ks} @{H:2d}x{H:<2d} B={B}: {ms * 1e3:8.1f} us  "
              f"{flops / ms / 1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
