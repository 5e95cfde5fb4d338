"""ORACLE (test infrastructure, never imported by the product path).

numpy / scipy restatement of the once-per-volume transforms of the reference's datamodule
(src/datamodules/create_dataset.py:196-218, get_transform):

    tio.CropOrPad((h, w, d), padding_mode=0)
    tio.RescaleIntensity((0, 1), percentiles=(perc_low, perc_high), masking_method='mask')
    tio.Resample(rescaleFactor, image_interpolation='bspline')          (tio.LabelMap entries: nearest neighbour)

and of vol2slice (:143-193).  torchio==0.18.84 and SimpleITK==2.2.0 (the reference's requirements) are neither vendored
in /root/reference nor installed in this image, so these functions restate torchio's published source; everything they
delegate to NumPy (np.percentile, np.clip, the float32 arithmetic of RescaleIntensity.rescale) IS executed by the real
NumPy here, i.e. pinned.  PARITY UNPINNED for: the start/end split of CropOrPad, the sample grid of Resample, and the
B-spline itself (scipy.ndimage's exact mirror initialisation stands in for ITK's 1e-10-truncated one).
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi


def crop_or_pad(vol: np.ndarray, target, pad_value=0) -> np.ndarray:
    """tio.CropOrPad(target, padding_mode=pad_value) on an [H,W,D] array: pad first, then crop; for an odd difference
    the extra voxel goes to the START (CropOrPad._get_six_bounds_parameters: ini = ceil(n / 2), fin = floor(n / 2))."""
    src = np.array(vol.shape)
    tgt = np.array(target)
    diff = tgt - src
    pad = np.maximum(diff, 0)
    crop = -np.minimum(diff, 0)
    out = vol
    if pad.any():
        widths = [(int(np.ceil(p / 2)), int(np.floor(p / 2))) for p in pad]
        out = np.pad(out, widths, mode="constant", constant_values=pad_value)
    if crop.any():
        sl = []
        for c, n in zip(crop, out.shape):
            ini, fin = int(np.ceil(c / 2)), int(np.floor(c / 2))
            sl.append(slice(ini, n - fin))
        out = out[tuple(sl)]
    return np.ascontiguousarray(out)


def rescale_intensity(vol: np.ndarray, mask: np.ndarray, out_min_max=(0, 1), percentiles=(1, 99)):
    """tio.RescaleIntensity(out_min_max, percentiles, masking_method='mask').rescale on float32 [H,W,D] arrays.
    Returns (rescaled float32 array, cutoff float64[2])."""
    array = vol.astype(np.float32).copy()
    m = mask > 0
    if not m.any():
        return array, np.array([np.nan, np.nan])
    values = array[m]
    cutoff = np.percentile(values, percentiles)
    np.clip(array, *cutoff, out=array)
    in_min, in_max = array.min(), array.max()
    in_range = in_max - in_min
    if in_range == 0:
        return vol.astype(np.float32).copy(), cutoff
    array -= in_min
    array /= in_range
    out_range = out_min_max[1] - out_min_max[0]
    array *= out_range
    array += out_min_max[0]
    return array, cutoff


def resample_size(n: int, factor: float) -> int:
    return max(1, int(np.ceil(n / factor)))


def _sample_coords(n_in, n_out, f):
    return 0.5 * (f - 1.0) + f * np.arange(n_out, dtype=np.float64)


def resample(vol: np.ndarray, factor, bspline=True) -> np.ndarray:
    """tio.Resample(factor) for a unit-spacing [H,W,D] array.  Image: cubic B-spline (float64 coefficients, mirror
    boundary); label map: nearest neighbour with halves rounded up.  Samples outside [-0.5, N - 0.5) read 0."""
    f = (factor,) * 3 if np.isscalar(factor) else tuple(factor)
    shape_out = tuple(resample_size(n, fi) for n, fi in zip(vol.shape, f))
    coords = [_sample_coords(n, m, fi) for n, m, fi in zip(vol.shape, shape_out, f)]
    inside = [(c >= -0.5) & (c < n - 0.5) for c, n in zip(coords, vol.shape)]
    grid = np.meshgrid(*coords, indexing="ij")
    if bspline:
        coef = ndi.spline_filter(vol.astype(np.float64), order=3, mode="mirror", output=np.float64)
        out = ndi.map_coordinates(coef, grid, order=3, mode="mirror", prefilter=False)
    else:
        idx = [np.minimum(np.floor(c + 0.5).astype(np.int64), n - 1) for c, n in zip(coords, vol.shape)]
        out = vol[np.ix_(*idx)].astype(np.float64)
    ok = inside[0][:, None, None] & inside[1][None, :, None] & inside[2][None, None, :]
    return np.where(ok, out, 0.0).astype(np.float32)


def get_transform(vol, mask, cfg, seg=None):
    """get_transform(cfg) (create_dataset.py:196-218, unisotropic_sampling=True) applied to one case: returns the dict
    {vol, mask[, seg]} after CropOrPad -> RescaleIntensity -> Resample."""
    target = tuple(cfg.get("imageDim", (160, 192, 160)))
    f = cfg.get("rescaleFactor", 3.0)
    v = crop_or_pad(vol, target)
    m = crop_or_pad(mask, target)
    v, _ = rescale_intensity(v, m, (0, 1), (cfg.get("perc_low", 1), cfg.get("perc_high", 99)))
    out = {"vol": resample(v, f, True), "mask": resample(m, f, False)}
    if seg is not None:
        out["seg"] = resample(crop_or_pad(seg, target), f, False)
    return out
