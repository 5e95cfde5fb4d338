"""Two-rank check of the data-parallel training claims (run under torchrun on 2 GPUs):
  1. torch DistributedDataParallel (what Lightning's DDP strategy wraps the module in) around DDPM_2D: the UNet engine's
     autograd node feeds DDP's reducer like any other module - gradients are identical on both ranks afterwards;
  2. the plain-loop path cddpm.dist_train.sync_gradients gives the same averaged gradients as DDP.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "conditioned-diffusion-models-uad_b200")]
import bench  # noqa: E402
from cddpm.ddpm_2d import DDPM_2D  # noqa: E402
from cddpm.dist_train import sync_gradients  # noqa: E402
from oracle.weights import synthetic_slices  # noqa: E402


class TrainStep(torch.nn.Module):
    """forward() = training_step's loss, so DDP's forward hook sees the call."""

    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, batch):
        return self.m.training_step(batch, 0)["loss"]


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = bench.model_cfg()
    cfg["engine_dtype"] = "bf16"

    def build():
        torch.manual_seed(0)
        m = DDPM_2D(cfg).cuda().train()
        with torch.no_grad():
            for n, p in m.named_parameters():
                if p.dim() >= 2 and float(p.abs().sum()) == 0.0:
                    p.normal_(0, 0.02)
        return m

    batch = {"vol": {"data": synthetic_slices(8, 96, seed=10 + rank).cuda().unsqueeze(-1)}}  # different data per rank

    def grads_of(m):
        return torch.cat([p.grad.flatten().float() for p in m.parameters()])

    # 1. DDP
    os.environ["CDDPM_ENCODER_GRAPH"] = "0"  # DDP hooks + graph capture of the encoder do not mix in this check
    m1 = build()
    ddp = torch.nn.parallel.DistributedDataParallel(TrainStep(m1), device_ids=[local])
    torch.manual_seed(100)
    np.random.seed(100)
    ddp(batch).backward()
    g1 = grads_of(m1)
    other = [torch.empty_like(g1) for _ in range(2)]
    dist.all_gather(other, g1)
    same = float((other[0] - other[1]).abs().max())
    # 2. plain loop + sync_gradients
    m2 = build()
    torch.manual_seed(100)
    np.random.seed(100)
    m2.training_step(batch, 0)["loss"].backward()
    calls = sync_gradients(m2)
    g2 = grads_of(m2)
    rel = float((g1 - g2).norm() / g1.norm())
    if rank == 0:
        print(f"DDP: max |grad(rank0) - grad(rank1)| = {same:.3g}; sync_gradients ({calls} collectives) vs DDP rel-L2 = {rel:.3g}; "
              f"|g| = {float(g1.norm()):.4g}", flush=True)
    assert same == 0.0
    assert rel < 2e-2  # two independent bf16 backward passes with atomically ordered fp32 sums
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
