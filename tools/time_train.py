"""Device-side timing of the training step pieces (development aid): UNet engine forward / backward at batch B, and the
whole DDPM_2D.training_step + backward + Adam.   python tools/time_train.py [B] [--profile]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]

from cddpm.engine import UNetEngine  # noqa: E402
from oracle import unet_port  # noqa: E402
from oracle.weights import make_state_dict, synthetic_slices  # noqa: E402


def ev():
    return torch.cuda.Event(enable_timing=True)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    profile = "--profile" in sys.argv
    B = int(args[0]) if args else 64
    spec = unet_port.UNetSpec()
    eng = UNetEngine(image_size=(96, 96), in_channels=1, model_channels=128, out_channels=1, num_res_blocks=3,
                     attention_resolutions=(3, 6, 12), channel_mult=(1, 2, 2), num_classes=128, dtype=torch.bfloat16,
                     training=True)
    sd = make_state_dict(unet_port.param_shapes(spec), seed=1)
    eng.load_state_dict({k: v.cuda() for k, v in sd.items()})
    x = torch.randn(B, 1, 96, 96, device="cuda")
    t = torch.randint(0, 1000, (B,), device="cuda")
    c = torch.randn(B, 128, device="cuda")
    dout = torch.randn(B, 1, 96, 96, device="cuda") / (B * 9216)
    for _ in range(3):
        eng.forward(x, t, c)
        eng.backward(dout, want_dcond=True)
    torch.cuda.synchronize()
    if profile:
        torch.cuda.cudart().cudaProfilerStart()
        eng.forward(x, t, c)
        eng.backward(dout, want_dcond=True)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        print("profiled one forward + backward at B =", B)
        return
    n = 5
    e = [ev() for _ in range(3)]
    tf = tb = 0.0
    for _ in range(n):
        e[0].record()
        eng.forward(x, t, c)
        e[1].record()
        eng.backward(dout, want_dcond=True)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]) / n
        tb += e[1].elapsed_time(e[2]) / n
    ff = eng.conv_flops_per_sample * B
    fb = eng.bwd_flops_per_sample * B
    print(f"B={B}: UNet forward {tf:.3f} ms ({ff / tf / 1e9:.1f} TFLOP/s), backward {tb:.3f} ms "
          f"({fb / tb / 1e9:.1f} TFLOP/s over {eng.bwd_flops_per_sample / 1e9:.1f} GFLOP/slice), "
          f"fwd+bwd {tf + tb:.3f} ms = {B / (tf + tb) * 1e3:.1f} slices/s", flush=True)

    # whole training step through the LightningModule surface
    from cddpm.ddpm_2d import DDPM_2D

    class Cfg(dict):
        __getattr__ = dict.get

        def __setattr__(self, k, v):
            self[k] = v

    cfg = Cfg(imageDim=[192, 192, 100], rescaleFactor=2, unet_dim=128, dim_mults=[1, 2, 2], condition=True,
              backbone="Spark_Encoder_2D", version="resnet50", cond_dim=128, noisetype="simplex", test_timesteps=500,
              lr=1e-4, objective="pred_x0", pretrained_encoder=False, engine_dtype="bf16")
    del eng
    torch.cuda.empty_cache()
    m = DDPM_2D(cfg).cuda().train()
    opt = m.configure_optimizers()
    batch = {"vol": {"data": synthetic_slices(B, 96, seed=1).cuda().unsqueeze(-1)}}
    np.random.seed(0)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = m.training_step(batch, 0)["loss"]
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B}: DDPM_2D.training_step + backward + Adam {ms:.3f} ms/step = {B / ms * 1e3:.1f} train-slices/s", flush=True)

    # coarse wall-clock breakdown (a synchronize between stages, so the sum exceeds the pipelined step)
    import time

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, (time.perf_counter() - t0) * 1e3

    acc = {}
    for _ in range(n):
        opt.zero_grad(set_to_none=True)
        x = batch["vol"]["data"].squeeze(-1)
        feats, t_enc = timed(lambda: m(x))
        _, t_push = timed(lambda: m.diffusion.model.train_engine())
        (loss, _), t_fwd = timed(lambda: m.diffusion(x, cond=feats, noise=torch.zeros_like(x).half()))
        _, t_bwd = timed(lambda: loss.backward())
        _, t_opt = timed(lambda: opt.step())
        for k, v in (("encoder fwd (torch)", t_enc), ("param push", t_push), ("q_sample+UNet fwd+loss", t_fwd),
                     ("backward (UNet engine + encoder autograd)", t_bwd), ("Adam", t_opt)):
            acc[k] = acc.get(k, 0.0) + v / n
    print("   " + "; ".join(f"{k} {v:.2f} ms" for k, v in acc.items()), flush=True)


if __name__ == "__main__":
    main()
