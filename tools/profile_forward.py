"""One UNet forward between cudaProfilerStart/Stop for ncu (--profile-from-start off).
   python tools/profile_forward.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]

from cddpm.engine import UNetEngine  # noqa: E402
from oracle import unet_port  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    spec = unet_port.UNetSpec()
    eng = UNetEngine(image_size=(96, 96), in_channels=1, model_channels=128, out_channels=1, num_res_blocks=3,
                     attention_resolutions=(3, 6, 12), channel_mult=(1, 2, 2), num_classes=128, dtype=torch.float16)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=1)
    eng.load_state_dict({k: v.cuda() for k, v in sd.items()})
    x = torch.randn(B, 1, 96, 96, device="cuda")
    t = torch.full((B,), 499, device="cuda", dtype=torch.long)
    c = torch.randn(B, 128, device="cuda")
    out = torch.empty_like(x)
    for _ in range(2):
        eng.forward(x, t, c, out)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eng.forward(x, t, c, out)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("done", float(out.abs().mean()))


if __name__ == "__main__":
    main()
