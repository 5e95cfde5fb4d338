"""GPU parity of the preprocessing kernels (create_dataset.get_transform, :196-218) against the numpy /
scipy oracle port (oracle/preprocess_port.py).  CropOrPad, nearest-neighbour resampling and RescaleIntensity (NumPy's
float64 percentile, np.clip, float32 rescale) must be BIT-EXACT; the cubic B-spline is compared at 2e-6 (float64
coefficients on both sides, float32 result)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Cfg(dict):
    __getattr__ = dict.get


def _case(shape, seed):
    g = np.random.default_rng(seed)
    low = g.random((shape[0] // 6 + 2, shape[1] // 6 + 2, shape[2] // 5 + 2))
    import scipy.ndimage as ndi

    vol = ndi.zoom(low, [s / l for s, l in zip(shape, low.shape)], order=1)[:shape[0], :shape[1], :shape[2]]
    vol = (vol * 900 + 30 * g.random(shape)).astype(np.float32)
    yy, xx, zz = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    c = [(s - 1) / 2 for s in shape]
    mask = (((yy - c[0]) / (0.42 * shape[0])) ** 2 + ((xx - c[1]) / (0.38 * shape[1])) ** 2 +
            ((zz - c[2]) / (0.46 * shape[2])) ** 2 <= 1).astype(np.float32)
    return np.ascontiguousarray(vol * mask), mask


@pytest.mark.parametrize("src,tgt", [((20, 31, 17), (24, 25, 17)), ((33, 16, 9), (16, 33, 12)), ((8, 8, 8), (8, 8, 8))])
def test_crop_or_pad_bit_exact(src, tgt):
    from cddpm.preprocess import CropOrPad
    from oracle import preprocess_port as pp

    vol = np.random.default_rng(1).random(src, dtype=np.float32)
    out = CropOrPad(tgt, padding_mode=0)({"vol": torch.from_numpy(vol)[None].cuda()})["vol"]
    assert tuple(out.shape) == (1,) + tgt
    assert np.array_equal(out[0].cpu().numpy(), pp.crop_or_pad(vol, tgt))


@pytest.mark.parametrize("shape,perc", [((48, 40, 22), (1, 99)), ((30, 30, 30), (0, 100)), ((64, 48, 20), (2.5, 97.5))])
def test_rescale_intensity_bit_exact_vs_numpy(shape, perc):
    from cddpm.preprocess import RescaleIntensity
    from oracle import preprocess_port as pp

    vol, mask = _case(shape, 3)
    t = RescaleIntensity((0, 1), percentiles=perc, masking_method="mask")
    sub = t({"vol": torch.from_numpy(vol)[None].cuda(), "mask": torch.from_numpy(mask)[None].cuda()})
    ref, cut = pp.rescale_intensity(vol, mask, (0, 1), perc)
    got_cut = t.last_cutoffs["vol"].cpu().numpy()
    assert np.array_equal(got_cut, cut), (got_cut, cut)  # np.percentile in float64, bit for bit
    assert np.array_equal(sub["vol"][0].cpu().numpy(), ref)
    assert torch.equal(sub["mask"].cpu(), torch.from_numpy(mask)[None])  # label maps are not rescaled


def test_rescale_intensity_degenerate_cases():
    from cddpm.preprocess import RescaleIntensity

    vol, mask = _case((24, 24, 12), 5)
    v = torch.from_numpy(vol)[None].cuda()
    out = RescaleIntensity((0, 1), (1, 99), "mask")({"vol": v.clone(), "mask": torch.zeros_like(v)})["vol"]
    assert torch.equal(out, v)  # empty mask: torchio warns and returns the tensor unchanged
    const = torch.full_like(v, 3.0)
    out = RescaleIntensity((0, 1), (1, 99), "mask")({"vol": const.clone(), "mask": torch.ones_like(v)})["vol"]
    assert torch.equal(out, const)  # zero range: unchanged


@pytest.mark.parametrize("shape,factor", [((48, 40, 22), 2.0), ((31, 20, 15), 2.0), ((24, 24, 24), 3.0), ((16, 18, 10), 1.0)])
def test_resample_vs_oracle(shape, factor):
    from cddpm.preprocess import Resample
    from oracle import preprocess_port as pp

    vol, mask = _case(shape, 7)
    vol = (vol / 900).astype(np.float32)
    sub = Resample(factor, image_interpolation="bspline")({"vol": torch.from_numpy(vol)[None].cuda(),
                                                           "mask": torch.from_numpy(mask)[None].cuda()})
    ref_v, ref_m = pp.resample(vol, factor, True), pp.resample(mask, factor, False)
    assert tuple(sub["vol"].shape[1:]) == ref_v.shape
    err = np.abs(sub["vol"][0].cpu().numpy() - ref_v).max()
    print(f"bspline resample {shape} / {factor}: max-abs {err:.3g}")
    assert err <= 2e-6
    assert np.array_equal(sub["mask"][0].cpu().numpy(), ref_m)  # nearest neighbour: exact


def test_get_transform_full_size_vs_oracle():
    """The reference's pipeline at its own size: [1,192,192,100]-bound volume (here 180x200x96 before CropOrPad), cfg of
    configs/experiment/*: imageDim [192,192,100], rescaleFactor 2 -> 96x96x50."""
    from cddpm.preprocess import Image, get_transform
    from oracle import preprocess_port as pp

    cfg = Cfg(imageDim=[192, 192, 100], rescaleFactor=2, resizedEvaluation=True, perc_low=1, perc_high=99)
    vol, mask = _case((180, 200, 96), 11)
    sub = {"vol": Image(torch.from_numpy(vol)[None].cuda()), "mask": Image(torch.from_numpy(mask)[None].cuda()),
           "ID": "case0"}
    out = get_transform(cfg)(sub)
    ref = pp.get_transform(vol, mask, cfg)
    assert tuple(out["vol"].data.shape) == (1, 96, 96, 50) and out["ID"] == "case0"
    err = np.abs(out["vol"].data[0].cpu().numpy() - ref["vol"]).max()
    print(f"get_transform: vol max-abs {err:.3g}; range [{float(out['vol'].data.min()):.3f}, {float(out['vol'].data.max()):.3f}]")
    assert err <= 2e-6
    assert np.array_equal(out["mask"].data[0].cpu().numpy(), ref["mask"])
    # resizedEvaluation=False keeps the *_orig entries at full resolution (exclude list of tio.Resample)
    cfg2 = Cfg(cfg, resizedEvaluation=False)
    sub2 = {"vol": torch.from_numpy(vol)[None].cuda(), "mask": torch.from_numpy(mask)[None].cuda(),
            "vol_orig": torch.from_numpy(vol)[None].cuda(), "mask_orig": torch.from_numpy(mask)[None].cuda()}
    out2 = get_transform(cfg2)(sub2)
    assert tuple(out2["vol_orig"].shape) == (1, 192, 192, 100) and tuple(out2["vol"].shape) == (1, 96, 96, 50)
    assert float(out2["vol_orig"].min()) == 0.0 and float(out2["vol_orig"].max()) == 1.0


def test_preprocess_needs_cuda():
    from cddpm import CddpmError
    from cddpm.preprocess import CropOrPad

    with pytest.raises(CddpmError):
        CropOrPad((4, 4, 4))({"vol": torch.zeros(1, 4, 4, 4)})
