"""CPU tests of the preprocessing row: vol2slice against a golden of the LIVE reference class
(tests/golden/vol2slice.json, oracle/make_golden.py vol2slice), and the numpy / scipy oracle port's own invariants."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


class Cfg(dict):
    __getattr__ = dict.get


def _make_ds(n=6, depth=20):
    from cddpm.preprocess import Image

    g = torch.Generator().manual_seed(77)
    ds = []
    for i in range(n):
        vol = torch.rand(1, 8, 8, depth, generator=g)
        mask = torch.zeros(1, 8, 8, depth)
        mask[..., 3 + i % 3:15 - i % 2] = 1.0
        ds.append({"vol": Image(vol), "mask": Image(mask)})
    return ds


def test_vol2slice_matches_live_reference_golden():
    from cddpm.preprocess import vol2slice

    with open(os.path.join(GOLD, "vol2slice.json")) as f:
        gold = json.load(f)
    cases = {"random": dict(), "start": dict(slice=7), "window": dict(slice=5, seq_slices=6), "brain": dict(onlyBrain=True),
             "unique": dict(cfg=dict(unique_slice=True, batch_size=3))}
    for name, kw in cases.items():
        kw = dict(kw)
        cfg = Cfg(kw.pop("cfg", {}))
        torch.manual_seed(123)
        v2s = vol2slice(_make_ds(), cfg, **kw)
        for i in range(len(v2s)):
            s = v2s[i]
            ind = s["ind"] if isinstance(s["ind"], int) else int(s["ind"])
            want = gold[name][i]
            assert ind == want[0], (name, i, ind, want)
            assert list(s["vol"].data.shape) == want[3]
            assert float(s["vol"].data.double().sum()) == want[1] and float(s["mask"].data.double().sum()) == want[2]


def test_preprocess_port_invariants():
    from oracle import preprocess_port as pp

    g = np.random.default_rng(0)
    vol = g.random((21, 30, 17), dtype=np.float32)
    # CropOrPad: shapes, centre alignment, odd differences put the extra voxel at the start
    out = pp.crop_or_pad(vol, (24, 25, 17))
    assert out.shape == (24, 25, 17)
    assert np.array_equal(out[2:23, :, :], vol[:, 3:28, :])  # pad 3 -> (2, 1); crop 5 -> (3, 2)
    assert not out[:2].any() and not out[23:].any()
    # RescaleIntensity: range [0, 1], cut-offs are the masked percentiles
    mask = (g.random(vol.shape) > 0.3).astype(np.float32)
    r, cut = pp.rescale_intensity(vol * 7 + 2, mask, (0, 1), (1, 99))
    assert r.dtype == np.float32 and r.min() == 0.0 and r.max() == 1.0
    assert np.allclose(cut, np.percentile((vol * 7 + 2)[mask > 0], (1, 99)))
    assert np.array_equal(pp.rescale_intensity(vol, np.zeros_like(mask))[0], vol)  # empty mask: unchanged
    # Resample: factor 1 is the identity (a B-spline interpolates its samples); factor 2 halves the extents (ceil)
    same = pp.resample(vol, 1.0, True)
    assert same.shape == vol.shape and np.abs(same - vol).max() < 1e-5
    half = pp.resample(vol, 2.0, True)
    assert half.shape == (11, 15, 9)
    # odd extents: the last sample of an odd axis sits at N - 0.5, outside the buffer -> 0 (itk::ResampleImageFilter)
    assert not half[10].any() and half[:10, :, :8].any()
    lab = pp.resample((vol > 0.5).astype(np.float32), 2.0, False)
    assert set(np.unique(lab)) <= {0.0, 1.0}
    assert np.array_equal(lab[:10, :, :8], (vol > 0.5).astype(np.float32)[1:21:2, 1::2, 1:17:2][:10, :, :8])
    # a linear ramp is reproduced exactly by the cubic B-spline away from the mirrored borders
    ramp = np.broadcast_to(np.arange(40, dtype=np.float32)[:, None, None], (40, 8, 8)).copy()
    rs = pp.resample(ramp, 2.0, True)
    assert np.abs(rs[5:15, 0, 0] - (0.5 + 2 * np.arange(5, 15))).max() < 1e-4
