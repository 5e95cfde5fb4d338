"""ORACLE (test infrastructure, never imported by the product path).

numpy restatement of the reference's OpenSimplex-2D fractal noise producer:
  gen_noise                 src/utils/generate_noise.py:8-15
  generate_simplex_noise    :19-52   (octave=6, persistence=0.8, frequency=64; random_param branch is dead)
  Simplex_CLASS.newSeed     :60-63   (np.random.randint(-1e10, 1e10); TWO draws per gen_noise call: ctor + :25)
  rand_2d_octaves           :97-114
  _init (LCG permutation)   :214-232
  _noise2 / _extrapolate2   :235-239, :252-349
  _noise2a                  :352-358 (indexes i*y.size+j and reshapes (x.size,y.size): square images only)

Pinned against the live reference by tests/golden/simplex_*.npz (oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np

_STRETCH2 = -0.211324865405187
_SQUISH2 = 0.366025403784439
_NORM2 = 47
_GRAD2 = np.array([5, 2, 2, 5, -5, 2, -2, 5, 5, -2, 2, -5, -5, -2, -2, -5], dtype=np.int64)
_MUL = 6364136223846793005
_INC = 1442695040888963407


def _wrap64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def permutation_from_seed(seed: int) -> np.ndarray:
    """256-entry permutation drawn without replacement by a 64-bit LCG (generate_noise.py:214-232)."""
    perm = np.zeros(256, dtype=np.int64)
    pool = list(range(256))
    s = int(seed)
    for _ in range(3):
        s = _wrap64(s * _MUL + _INC)
    for i in range(255, -1, -1):
        s = _wrap64(s * _MUL + _INC)
        r = (s + 31) % (i + 1)
        perm[i] = pool[r]
        pool[r] = pool[i]
    return perm


def _grad_dot(perm, xs, ys, dx, dy):
    idx = perm[(perm[xs & 0xFF] + ys) & 0xFF] & 0x0E
    return _GRAD2[idx] * dx + _GRAD2[idx + 1] * dy


def noise2_grid(x: np.ndarray, y: np.ndarray, perm: np.ndarray) -> np.ndarray:
    """out[i, j] = OpenSimplex2(x[j], y[i]) in float64, evaluated with the reference's operation order."""
    X, Y = np.meshgrid(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))
    stretch = (X + Y) * _STRETCH2
    xs = X + stretch
    ys = Y + stretch
    xsb = np.floor(xs).astype(np.int64)
    ysb = np.floor(ys).astype(np.int64)
    squish = (xsb + ysb) * _SQUISH2
    xb = xsb + squish
    yb = ysb + squish
    xins = xs - xsb
    yins = ys - ysb
    in_sum = xins + yins
    dx0 = X - xb
    dy0 = Y - yb
    value = np.zeros_like(X)

    def contribute(dx, dy, gx, gy):
        attn = 2 - dx * dx - dy * dy
        a2 = attn * attn
        term = a2 * a2 * _grad_dot(perm, gx, gy, dx, dy)
        return np.where(attn > 0, term, 0.0)

    dx1 = dx0 - 1 - _SQUISH2
    dy1 = dy0 - 0 - _SQUISH2
    value = value + contribute(dx1, dy1, xsb + 1, ysb + 0)
    dx2 = dx0 - 0 - _SQUISH2
    dy2 = dy0 - 1 - _SQUISH2
    value = value + contribute(dx2, dy2, xsb + 0, ysb + 1)

    lower = in_sum <= 1
    # --- lower triangle (origin (0,0))
    zins_l = 1 - in_sum
    near0_l = (zins_l > xins) | (zins_l > yins)
    xgt = xins > yins
    # --- upper triangle (origin (1,1))
    zins_u = 2 - in_sum
    near0_u = (zins_u < xins) | (zins_u < yins)

    xsv = np.empty_like(xsb)
    ysv = np.empty_like(ysb)
    dxe = np.empty_like(dx0)
    dye = np.empty_like(dy0)

    m = lower & near0_l & xgt
    xsv[m], ysv[m], dxe[m], dye[m] = xsb[m] + 1, ysb[m] - 1, dx0[m] - 1, dy0[m] + 1
    m = lower & near0_l & ~xgt
    xsv[m], ysv[m], dxe[m], dye[m] = xsb[m] - 1, ysb[m] + 1, dx0[m] + 1, dy0[m] - 1
    m = lower & ~near0_l
    xsv[m], ysv[m] = xsb[m] + 1, ysb[m] + 1
    dxe[m], dye[m] = dx0[m] - 1 - 2 * _SQUISH2, dy0[m] - 1 - 2 * _SQUISH2
    m = ~lower & near0_u & xgt
    xsv[m], ysv[m] = xsb[m] + 2, ysb[m] + 0
    dxe[m], dye[m] = dx0[m] - 2 - 2 * _SQUISH2, dy0[m] + 0 - 2 * _SQUISH2
    m = ~lower & near0_u & ~xgt
    xsv[m], ysv[m] = xsb[m] + 0, ysb[m] + 2
    dxe[m], dye[m] = dx0[m] + 0 - 2 * _SQUISH2, dy0[m] - 2 - 2 * _SQUISH2
    m = ~lower & ~near0_u
    xsv[m], ysv[m], dxe[m], dye[m] = xsb[m], ysb[m], dx0[m], dy0[m]

    up = ~lower
    xsb0 = np.where(up, xsb + 1, xsb)
    ysb0 = np.where(up, ysb + 1, ysb)
    dx00 = np.where(up, dx0 - 1 - 2 * _SQUISH2, dx0)
    dy00 = np.where(up, dy0 - 1 - 2 * _SQUISH2, dy0)

    value = value + contribute(dx00, dy00, xsb0, ysb0)
    value = value + contribute(dxe, dye, xsv, ysv)
    return value / _NORM2


def fractal_field(shape_hw, perm, octaves=6, persistence=0.8, frequency=64.0) -> np.ndarray:
    """rand_2d_octaves (generate_noise.py:97-114): sum_o persistence^o * noise2(x/f_o, y/f_o), f_o = frequency/2^o."""
    h, w = shape_hw
    assert h == w, "the reference's _noise2a indexing is only correct for square images"
    y = np.arange(0, h)
    x = np.arange(0, w)
    out = np.zeros((h, w))
    amp = 1
    f = frequency
    for _ in range(octaves):
        out += amp * noise2_grid(x / f, y / f, perm)
        f /= 2
        amp *= persistence
    return out


def draw_seed() -> int:
    """One Simplex_CLASS.newSeed() draw from numpy's global RNG (generate_noise.py:60-63)."""
    s = 0
    while not s:
        s = int(np.random.randint(-10000000000, 10000000000))
    return s


def gen_noise_port(shape):
    """gen_noise(cfg, shape) for noisetype == 'simplex': one HxW field repeated over the batch, float16 torch tensor.
    Consumes two seeds from np.random exactly like the reference (constructor + generate_simplex_noise)."""
    import torch

    draw_seed()  # Simplex_CLASS.__init__
    seed = draw_seed()  # generate_simplex_noise -> newSeed()
    field = fractal_field((shape[2], shape[3]), permutation_from_seed(seed))
    t = torch.from_numpy(field).unsqueeze(0).repeat(shape[0], 1, 1, 1)
    return t.half()
