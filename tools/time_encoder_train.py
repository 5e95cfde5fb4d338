"""Device-side timing of the training-mode condition encoder (development aid): hand-written engine vs the library path.
   python tools/time_encoder_train.py [B] [mode ...]      modes: b200 tf32 bf16 fp32"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]


class Cfg(dict):
    __getattr__ = dict.get


def main():
    from cddpm.encoder import get_encoder

    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    modes = sys.argv[2:] or ["b200", "tf32"]
    x = torch.rand(B, 1, 96, 96, device="cuda")
    w = torch.randn(B, 128, device="cuda")
    for mode in modes:
        cfg = Cfg(imageDim=[192, 192, 100], rescaleFactor=2, backbone="Spark_Encoder_2D", version="resnet50", cond_dim=128,
                  encoder_train_dtype=mode, encoder_drop_path_rate=0.05)
        torch.manual_seed(0)
        enc, _ = get_encoder(cfg)
        with torch.no_grad():
            for n, p in enc.named_parameters():
                if n.endswith("bn3.weight"):
                    p.fill_(0.2)
        enc = enc.cuda().train()

        def step():
            for p in enc.parameters():
                p.grad = None
            out = enc(x)
            (out * w).sum().backward()

        for _ in range(4):
            step()
        torch.cuda.synchronize()
        n = int(os.environ.get("TIME_ITERS", "20"))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        print(f"encoder train step (fwd + bwd) B={B} mode={mode}: {e0.elapsed_time(e1) / n:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
