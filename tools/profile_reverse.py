"""A short reverse loop (start_t steps) between cudaProfilerStart/Stop for ncu (development aid).
   python tools/profile_reverse.py [B] [start_t]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]
sys.argv = sys.argv[:1] + [a for a in sys.argv[1:]]
import bench  # noqa: E402
from cddpm.ddpm_2d import DDPM_2D  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T0 = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    torch.manual_seed(0)
    np.random.seed(0)
    model = DDPM_2D(bench.model_cfg(), prefix="p/").cuda().eval()
    x = torch.rand(B, 1, 96, 96, device="cuda")

    def step():
        with torch.no_grad():
            cond = model(x)
            return model.diffusion.sample(cond=cond, x_start=x * 2 - 1, start_t=T0, noise=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    print(f"B={B} T0={T0}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per loop", flush=True)
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
