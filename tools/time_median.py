"""Device-side timing of the 5x5x5 median (development aid): python tools/time_median.py [D H W]
CDDPM_MEDIAN_V2=0 selects the first-generation radix-descent kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]


def main():
    from cddpm._lib import check, current_stream, lib, ptr

    D, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (50, 96, 96)
    g = torch.Generator(device="cuda").manual_seed(0)
    # a residual-like volume: zero outside an ellipsoid "brain", positive inside
    zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, D, device="cuda"), torch.linspace(-1, 1, H, device="cuda"),
                                torch.linspace(-1, 1, W, device="cuda"), indexing="ij")
    mask = (zz ** 2 + (yy / 0.8) ** 2 + (xx / 0.7) ** 2) < 0.8
    vol = (torch.rand(D, H, W, device="cuda", generator=g) * mask).contiguous()
    out = torch.empty_like(vol)
    for _ in range(5):
        check(lib().cddpm_median3d(ptr(vol), ptr(out), H, W, D, 5, current_stream()), "median")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        check(lib().cddpm_median3d(ptr(vol), ptr(out), H, W, D, 5, current_stream()), "median")
    e1.record()
    torch.cuda.synchronize()
    print(f"median 5x5x5 over [{D},{H},{W}] (brain fraction {mask.float().mean().item():.2f}): "
          f"{e0.elapsed_time(e1) / n * 1000:.1f} us  V2={os.environ.get('CDDPM_MEDIAN_V2', '1')} "
          f"checksum {out.double().sum().item():.6f}", flush=True)


if __name__ == "__main__":
    main()
