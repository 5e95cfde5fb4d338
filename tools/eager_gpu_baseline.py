"""Eager-PyTorch-on-B200 reference point for one conditioned UNet forward (development aid, not the bench).

SURVEY.md §8(d) names "PyTorch eager on the GPU (fp16 autocast, cudnn.benchmark)" as the real bar next to the CPU
baseline.  /root/reference does not exist on the GPU box, so this times the oracle's PyTorch restatement of
UNetModel.forward (oracle/unet_port.py - the same op sequence: F.conv2d / F.group_norm / F.silu / softmax through
cuDNN and ATen) under torch.autocast(float16), channels-last weights, cudnn.benchmark on.
   python tools/eager_gpu_baseline.py [B ...]
   python tools/eager_gpu_baseline.py --train [B ...]    UNet forward + backward (bf16 autocast) + fused Adam"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

from oracle import unet_port  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402


def train(batches):
    spec = unet_port.UNetSpec()
    sd = {k: v.cuda().requires_grad_(True) for k, v in make_state_dict(unet_port.param_shapes(spec), seed=1).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=1e-4, fused=True)
    for B in batches:
        x = torch.randn(B, 1, 96, 96, device="cuda")
        t = torch.randint(0, 1000, (B,), device="cuda")
        c = torch.randn(B, 128, device="cuda")
        target = torch.randn(B, 1, 96, 96, device="cuda")

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = unet_port.unet_forward(sd, spec, x, t, c)
            loss = (out.float() - target).abs().mean()
            loss.backward()
            opt.step()

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        n = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"eager bf16-autocast UNet train step (fwd+bwd+Adam) B={B:3d}: {ms:8.3f} ms  {B / ms * 1e3:9.1f} slices/s",
              flush=True)


def main():
    if "--train" in sys.argv:
        torch.backends.cudnn.benchmark = True
        return train([int(a) for a in sys.argv[1:] if not a.startswith("--")] or [64])
    batches = [int(a) for a in sys.argv[1:]] or [32]
    torch.backends.cudnn.benchmark = True
    spec = unet_port.UNetSpec()
    sd = {k: v.cuda() for k, v in make_state_dict(unet_port.param_shapes(spec), seed=1).items()}
    for B in batches:
        x = torch.randn(B, 1, 96, 96, device="cuda")
        t = torch.full((B,), 499, device="cuda", dtype=torch.long)
        c = torch.randn(B, 128, device="cuda")
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            for _ in range(3):
                unet_port.unet_forward(sd, spec, x, t, c)
            torch.cuda.synchronize()
            n = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                unet_port.unet_forward(sd, spec, x, t, c)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"eager fp16-autocast B={B:3d}: {ms:8.3f} ms/forward  {B / ms * 1e3:9.1f} slices/s", flush=True)


if __name__ == "__main__":
    main()
