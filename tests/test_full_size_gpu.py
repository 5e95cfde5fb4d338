"""BASELINE.json's FULL sizes on the GPU, checked through size-independent properties (the oracle needs ~100 s per slice
for one 500-step reverse loop on host cores, so the point-wise comparisons live in the small-geometry tests):

* configs[1]: batch 32, encoder + full reverse loop from T0 = 500 (cond_DDPM.py:391-464) - finite, clipped to [-1, 1]
  (x0 is clamped every step, :417), batch-split consistent and batch-permutation equivariant (no cross-slice state:
  GroupNorm and attention are per image, the simplex field of a step is shared by the batch, generate_noise.py:12-14);
* UNet forward: slice i of a batch of 32 equals the same slice evaluated alone (different tile schedule, same maths);
* configs[2]: a 50-slice volume through test_step - the ensemble reconstruction of slice d does not depend on the
  other slices; thresholded volume after the component filter is a subset of the raw threshold mask.
Tolerances are stated per assertion; everything integer stays exact."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


class Cfg(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def _cfg(**over):
    c = Cfg(imageDim=[192, 192, 100], rescaleFactor=2, unet_dim=128, dim_mults=[1, 2, 2], condition=True,
            backbone="Spark_Encoder_2D", version="resnet50", cond_dim=128, noisetype="simplex", noise_ensemble=True,
            test_timesteps=500, lr=1e-4, resizedEvaluation=True, erodeBrainmask=True, medianFiltering=True,
            saveOutputImages=False, evalSeg=True, threshold="auto", spatial_transformer=False, pretrained_encoder=False,
            objective="pred_x0", force_num_eval_slices=False)
    c.update(over)
    return c


@pytest.fixture(scope="module")
def model():
    from src.models.DDPM_2D import DDPM_2D

    torch.manual_seed(21)
    m = DDPM_2D(_cfg(), prefix="t/")
    with torch.no_grad():  # the reference zero-initialises its output convolutions: give every tensor a value
        for _, p in m.named_parameters():
            if p.dim() >= 2 and float(p.abs().sum()) == 0.0:
                p.normal_(0.0, 1.0 / p[0].numel() ** 0.5)
    return m.cuda().eval()


def _slices(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 1, 96, 96, generator=g)
    yy, xx = torch.meshgrid(torch.arange(96.0), torch.arange(96.0), indexing="ij")
    brain = (((yy - 47.5) / 40) ** 2 + ((xx - 47.5) / 34) ** 2 <= 1).float()
    return (x * brain).cuda()


def test_unet_forward_is_batch_invariant(model):
    unet = model.diffusion.model
    x = _slices(32) * 2 - 1
    t = torch.full((32,), 499, device="cuda", dtype=torch.long)
    with torch.no_grad():
        cond = model(_slices(32))
        full = unet(x, t, cond).clone()
        for i in (0, 13, 31):
            one = unet(x[i:i + 1].contiguous(), t[:1], cond[i:i + 1].contiguous())
            err = (one - full[i:i + 1]).abs().max().item()
            assert err <= 2e-3, (i, err)  # fp16 activations; only the GroupNorm partial-sum order differs
        assert torch.isfinite(full).all()


def test_reverse_loop_full_size_properties(model):
    d = model.diffusion
    x = _slices(32)

    def run(xs, seed):
        np.random.seed(seed)
        with torch.no_grad():
            cond = model(xs)
            return d.sample(cond=cond, x_start=xs * 2 - 1, start_t=500, noise=True).clone()

    full = run(x, 5)
    assert full.shape == (32, 1, 96, 96) and torch.isfinite(full).all()
    assert full.min().item() >= -1.0 and full.max().item() <= 1.0  # x0 clamp of every step, cond_DDPM.py:417
    assert full.std().item() > 1e-3                                  # not collapsed
    # the first 8 slices alone, same noise stream: 500 steps of fp16 round-off differences stay small
    part = run(x[:8].contiguous(), 5)
    err = (part - full[:8]).abs().max().item()
    print(f"reverse loop B=32 vs B=8 split, T0=500: max-abs {err:.4g}")
    assert err <= 2e-2
    # permutation equivariance
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(1)).cuda()
    permuted = run(x[perm].contiguous(), 5)
    err = (permuted - full[perm]).abs().max().item()
    print(f"reverse loop permutation equivariance: max-abs {err:.4g}")
    assert err <= 2e-2
    # a different noise stream gives a different sample (the loop really consumes the noise)
    other = run(x, 6)
    assert (other - full).abs().max().item() > 1e-2


def test_volume_slices_are_independent_and_filter_is_a_subset(model):
    import bench
    from cddpm import eval_tail

    v = bench.synthetic_volume(3, 50)
    x = v["vol"].cuda().squeeze(0).permute(3, 0, 1, 2).contiguous()  # [50,1,96,96]
    with torch.no_grad():
        np.random.seed(9)
        reco, _, _ = model.reconstruct_slices(x)
        np.random.seed(9)
        reco_mid, _, _ = model.reconstruct_slices(x[20:30].contiguous())
    assert reco.shape == x.shape and torch.isfinite(reco).all()
    err = (reco_mid - reco[20:30]).abs().max().item()
    assert err <= 5e-3, err  # reconstruction tolerance of the path is 1e-2
    # anomaly map of the whole volume: the component filter only removes voxels, and removes only small components
    final = reco.squeeze(1).permute(1, 2, 0).unsqueeze(0).unsqueeze(0)
    vol, _ = eval_tail.residual_and_filter(final, v["vol"].cuda(), v["seg_orig"].cuda(), v["mask_orig"].cuda())
    thr = float(vol.diff.max().item()) * 0.25
    raw = (vol.diff > thr)
    kept = eval_tail.filter_3d_connected_components(raw)
    assert not (kept & ~raw).any()
    removed = (raw & ~kept)
    assert int(removed.sum()) <= int(raw.sum())
    if removed.any():  # every removed voxel has at most 6 set neighbours in its 3x3x3 window (component size <= 7)
        nb = torch.nn.functional.conv3d(raw[None, None].float(), torch.ones(1, 1, 3, 3, 3, device="cuda"), padding=1)[0, 0]
        assert nb[removed].max().item() <= 7
