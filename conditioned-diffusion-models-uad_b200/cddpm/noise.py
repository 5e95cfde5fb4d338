"""gen_noise — drop-in for src.utils.generate_noise.gen_noise (generate_noise.py:8-15).

The reference builds an OpenSimplex-2D fractal field on the host with numba (float64), repeats it over the batch and
casts to float16; it then has to be copied to the GPU on every call (DDPM_2D.py:231, cond_DDPM.py:442).  Here only
the 256-entry permutation is made on the host (a 64-bit LCG seeded from numpy's global RNG, consuming the same two
draws per call as the reference); the field itself is evaluated on the GPU by cddpm_simplex_noise, bit-identical.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import CddpmError, check, current_stream, lib, ptr

_MUL = 6364136223846793005
_INC = 1442695040888963407
_MASK = (1 << 64) - 1


def _lcg(s: int) -> int:
    s = (s * _MUL + _INC) & _MASK
    return s - (1 << 64) if s >> 63 else s


def permutation(seed: int) -> bytes:
    """Permutation table of Simplex_CLASS.newSeed -> _init (generate_noise.py:60-63, :214-232)."""
    s = int(seed)
    for _ in range(3):
        s = _lcg(s)
    src = list(range(256))
    perm = [0] * 256
    for i in range(255, -1, -1):
        s = _lcg(s)
        r = (s + 31) % (i + 1)
        perm[i] = src[r]
        src[r] = src[i]
    return bytes(perm)


def _new_seed() -> int:
    seed = 0
    while not seed:
        seed = int(np.random.randint(-10000000000, 10000000000))
    return seed


def simplex_field(seed: int, shape, device="cuda", octaves=6, persistence=0.8, frequency=64):
    """[B,1,H,W] float16 CUDA tensor holding the same fractal field for every batch element."""
    B, C, H, W = (int(v) for v in shape)
    if C != 1:
        raise CddpmError("simplex noise is defined for single-channel images")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise CddpmError("gen_noise produces its field on the GPU; there is no CPU path")
    out = torch.empty(B, 1, H, W, dtype=torch.float16, device=dev)
    with torch.cuda.device(dev):
        check(lib().cddpm_simplex_noise(permutation(seed), ptr(out), None, B, H, W, int(octaves), float(persistence),
                                        float(frequency), current_stream()), "cddpm_simplex_noise")
    return out


def gen_noise(cfg, shape, device="cuda"):
    """gen_noise(cfg, shape): cfg.noisetype must be 'simplex'.  Returns float16 [B,1,H,W] on `device` (the reference
    returns a CPU tensor that every caller immediately moves to the GPU)."""
    noisetype = cfg.get("noisetype") if hasattr(cfg, "get") else getattr(cfg, "noisetype", None)
    if noisetype != "simplex":
        raise ValueError("Noise type not recognized")
    _new_seed()  # Simplex_CLASS.__init__ draws one seed ...
    seed = _new_seed()  # ... generate_simplex_noise draws the one that is used (generate_noise.py:25)
    return simplex_field(seed, shape, device=device)
