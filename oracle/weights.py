"""ORACLE (test infrastructure, never imported by the product path).

Deterministic synthetic weights for parity tests.  The reference ships no checkpoints and its default init zeroes the
output convolutions (zero_module, util.py:174-180), which would make every parity test vacuous (a random-init UNet
outputs exactly 0).  So every tensor — including the zero-initialised ones — is drawn from a per-key seeded CPU
generator; the same function feeds the live reference (via load_state_dict, when generating tests/golden) and our
modules, so only inputs/outputs need to be committed, never weights.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import torch


def make_state_dict(shapes: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in shapes:
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.long)
        elif name.endswith("running_var"):
            sd[name] = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("running_mean"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[name] = gain * torch.randn(shape, generator=g) / fan_in ** 0.5
        elif name.endswith(".bias"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        else:  # norm scale
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
    return sd


def synthetic_slices(batch: int, size: int = 96, seed: int = 0) -> torch.Tensor:
    """[B,1,size,size] fp32 in [0,1]: uniform noise smoothed a little, masked by a centred ellipse (semi-axes
    40 x 34 px at 96) to mimic skull-stripped T2 slices."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 1, size, size, generator=g)
    x = torch.nn.functional.avg_pool2d(x, 3, stride=1, padding=1)
    yy, xx = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    c = (size - 1) / 2
    ell = (((yy - c) / (40 * size / 96)) ** 2 + ((xx - c) / (34 * size / 96)) ** 2) <= 1
    return x * ell[None, None].float()


def synthetic_volume(seed: int = 0, size: int = 96, depth: int = 50):
    """BraTS21-shaped synthetic case: returns dict of [1,1,H,W,D] float32 tensors
    vol (== vol_orig, resizedEvaluation), mask_orig (ellipsoid), seg_orig (1-3 spheres, r in [3,8]) and a plausible
    `reco` (the lesion-free image plus small noise) for tail-only tests."""
    g = torch.Generator().manual_seed(1000 + seed)
    H = W = size
    D = depth
    low = torch.rand(1, 1, H // 8, W // 8, max(D // 5, 2), generator=g)
    field = torch.nn.functional.interpolate(low, size=(H, W, D), mode="trilinear", align_corners=True)[0, 0]
    field = 0.2 + 0.6 * field
    yy, xx, zz = torch.meshgrid(torch.arange(H), torch.arange(W), torch.arange(D), indexing="ij")
    cy, cx, cz = (H - 1) / 2, (W - 1) / 2, (D - 1) / 2
    mask = (((yy - cy) / (0.42 * H)) ** 2 + ((xx - cx) / (0.36 * W)) ** 2 + ((zz - cz) / (0.48 * D)) ** 2) <= 1
    healthy = field * mask
    seg = torch.zeros(H, W, D, dtype=torch.bool)
    n_les = int(torch.randint(1, 4, (1,), generator=g))
    for _ in range(n_les):
        r = float(torch.randint(3, 9, (1,), generator=g))
        py = cy + float(torch.randn(1, generator=g)) * 0.12 * H
        px = cx + float(torch.randn(1, generator=g)) * 0.10 * W
        pz = cz + float(torch.randn(1, generator=g)) * 0.12 * D
        seg |= ((yy - py) ** 2 + (xx - px) ** 2 + (zz - pz) ** 2) <= r * r
    seg &= mask
    vol = (healthy + 0.3 * seg.float()).clamp(0, 1)
    reco = (healthy + 0.02 * torch.randn(H, W, D, generator=g) * mask).clamp(0, 1)
    u = lambda t: t.float()[None, None].contiguous()
    return {"vol": u(vol), "mask_orig": u(mask), "seg_orig": u(seg), "reco": u(reco)}


def synthetic_fullres_case(seed: int = 0, new_size=(160, 190, 160)):
    """A case for full-resolution evaluation (cfg.resizedEvaluation=False, utils_eval.py:24-25): the low-resolution
    synthetic volume's reconstruction stays [1,1,96,96,50]; vol_orig / mask_orig / seg_orig are its trilinear
    upsamplings to `new_size` (masks re-binarised), as a datamodule would hand them over."""
    v = synthetic_volume(seed)
    up = lambda t: torch.nn.functional.interpolate(t, size=tuple(new_size), mode="trilinear", align_corners=True)
    return {"reco": v["reco"], "vol_orig": up(v["vol"]).contiguous(), "mask_orig": (up(v["mask_orig"]) > 0.5).float(),
            "seg_orig": (up(v["seg_orig"]) > 0.5).float()}
