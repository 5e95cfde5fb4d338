"""Quick device-side timing of the UNet engine forward at several batch sizes (development aid, not the bench).
   python tools/time_forward.py [B ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]

from cddpm.engine import UNetEngine  # noqa: E402
from oracle import unet_port  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402


def main():
    batches = [int(a) for a in sys.argv[1:]] or [1, 8, 32, 64]
    spec = unet_port.UNetSpec()
    dtype = torch.float16
    eng = UNetEngine(image_size=(96, 96), in_channels=1, model_channels=128, out_channels=1, num_res_blocks=3,
                     attention_resolutions=(3, 6, 12), channel_mult=(1, 2, 2), num_classes=128, dtype=dtype)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=1)
    eng.load_state_dict({k: v.cuda() for k, v in sd.items()})
    for B in batches:
        x = torch.randn(B, 1, 96, 96, device="cuda")
        t = torch.full((B,), 499, device="cuda", dtype=torch.long)
        c = torch.randn(B, 128, device="cuda")
        out = torch.empty_like(x)
        for _ in range(3):
            eng.forward(x, t, c, out)
        torch.cuda.synchronize()
        n = int(os.environ.get("TIME_FORWARD_ITERS", 10 if B >= 8 else 30))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.forward(x, t, c, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = eng.conv_flops_per_sample * B
        print(f"B={B:3d}: {ms:8.3f} ms/forward  {B / ms * 1e3:9.1f} slices/s  conv {fl / ms / 1e9:7.1f} TFLOP/s "
              f"({eng.conv_flops_per_sample / 1e9:.1f} GFLOP/slice, {eng.launches_per_forward} launches)", flush=True)


if __name__ == "__main__":
    main()
