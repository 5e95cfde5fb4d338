"""GPU-resident volume preprocessing — the transforms of the reference's datamodule:
src/datamodules/create_dataset.py:196-218 `get_transform` (tio.CropOrPad -> tio.RescaleIntensity -> tio.Resample) and
:143-193 `vol2slice`.

The classes keep torchio's names and constructor arguments and act on a *subject*: a dict whose image entries are
CUDA tensors [C,H,W,D] (C = 1), the layout of `subject['vol'].data` in the reference.  Entries named in LABEL_KEYS are
label maps (tio.LabelMap in the reference: 'mask', 'seg', 'mask_orig', 'seg_orig'): RescaleIntensity skips them and
Resample interpolates them with nearest neighbour.  Everything voxel-sized runs in libcddpm_b200 (cddpm_crop_or_pad,
cddpm_rescale_intensity, cddpm_resample); there is no CPU path.

Pinned: the NumPy semantics torchio delegates to (np.percentile 'linear' in float64, np.clip, the float32 rescale) are
reproduced bit for bit (tests/test_preprocess_gpu.py against the real NumPy).  Parity unpinned (torchio / SimpleITK are
not installed): the start/end split of CropOrPad, Resample's sample grid and the B-spline initialisation
(scipy.ndimage stands in for ITK).  SimpleITK's CurvatureFlow denoising inside `sitk_reader` (:252-258) belongs to file
reading and is not part of this module.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, Optional, Sequence

import torch

from ._lib import CddpmError, check, current_stream, lib, ptr

LABEL_KEYS = ("mask", "seg", "mask_orig", "seg_orig")
DATA = "data"  # torchio.DATA


class Image(dict):
    """Stand-in for tio.ScalarImage / tio.LabelMap: a dict with a 'data' entry ([C,H,W,D] tensor) that also answers
    `.data`, which is how the reference reads and writes it (`subject['vol'].data`, `batch['vol'][tio.DATA]`)."""

    def __init__(self, tensor=None, **kw):
        super().__init__(**kw)
        if tensor is not None:
            self[DATA] = tensor

    @property
    def data(self):
        return self[DATA]

    @data.setter
    def data(self, value):
        self[DATA] = value

    @property
    def shape(self):
        return tuple(self[DATA].shape)


def _tensor_of(entry):
    if torch.is_tensor(entry):
        return entry
    if isinstance(entry, dict) and torch.is_tensor(entry.get(DATA)):
        return entry[DATA]
    return None


def _images(subject: Dict) -> Iterable[str]:
    out = []
    for k, v in subject.items():
        t = _tensor_of(v)
        if t is not None and t.dim() == 4:
            out.append(k)
    return out


def _get(subject: Dict, k: str) -> torch.Tensor:
    return _tensor_of(subject[k])


def _set(subject: Dict, k: str, value: torch.Tensor) -> None:
    if isinstance(subject[k], dict):
        subject[k][DATA] = value
    else:
        subject[k] = value


def _vol(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise CddpmError("preprocessing runs on CUDA tensors (there is no CPU path)")
    if t.dim() != 4 or t.shape[0] != 1:
        raise CddpmError(f"expected a [1,H,W,D] image tensor, got {tuple(t.shape)}")
    return t.float().contiguous()


class CropOrPad:
    """tio.CropOrPad(target_shape, padding_mode=0): centre crop / zero pad every image of the subject."""

    def __init__(self, target_shape: Sequence[int], padding_mode=0):
        if not isinstance(padding_mode, (int, float)):
            raise NotImplementedError("only constant padding (the reference passes padding_mode=0)")
        self.target_shape = tuple(int(v) for v in target_shape)
        self.padding_mode = float(padding_mode)

    def __call__(self, subject: Dict) -> Dict:
        h, w, d = self.target_shape
        for k in _images(subject):
            src = _vol(_get(subject, k))
            _, H, W, D = src.shape
            out = torch.empty(1, h, w, d, dtype=torch.float32, device=src.device)
            with torch.cuda.device(src.device):
                check(lib().cddpm_crop_or_pad(ptr(src), H, W, D, ptr(out), h, w, d, self.padding_mode, current_stream()),
                      "cddpm_crop_or_pad")
            _set(subject, k, out)
        return subject


class RescaleIntensity:
    """tio.RescaleIntensity(out_min_max, percentiles, masking_method=<name of a label map of the subject>)."""

    def __init__(self, out_min_max=(0, 1), percentiles=(0, 100), masking_method: Optional[str] = None):
        self.out_min, self.out_max = float(out_min_max[0]), float(out_min_max[1])
        self.percentiles = (float(percentiles[0]), float(percentiles[1]))
        self.masking_method = masking_method
        self.last_cutoffs: Dict[str, torch.Tensor] = {}

    def __call__(self, subject: Dict) -> Dict:
        for k in _images(subject):
            if k in LABEL_KEYS:
                continue
            vol = _vol(_get(subject, k)).clone()
            if self.masking_method is None:
                mask = torch.ones_like(vol)
            else:
                mask = _vol(_get(subject, self.masking_method))
                if mask.shape != vol.shape:
                    raise CddpmError(f"mask {tuple(mask.shape)} and image {tuple(vol.shape)} differ in shape")
            n = vol.numel()
            with torch.cuda.device(vol.device):
                nbytes = int(lib().cddpm_rescale_workspace_bytes(n))
                ws = torch.empty(nbytes, dtype=torch.uint8, device=vol.device)
                cut = torch.empty(2, dtype=torch.float64, device=vol.device)
                check(lib().cddpm_rescale_intensity(ptr(vol), ptr(mask), n, self.percentiles[0], self.percentiles[1],
                                                    self.out_min, self.out_max, ptr(ws), nbytes, ptr(cut),
                                                    current_stream()), "cddpm_rescale_intensity")
            self.last_cutoffs[k] = cut
            _set(subject, k, vol)
        return subject


class Resample:
    """tio.Resample(target, image_interpolation='bspline', exclude=[...]) for unit-spacing volumes: `target` is the new
    spacing (a number or three)."""

    def __init__(self, target=1, image_interpolation: str = "linear", exclude: Optional[Sequence[str]] = None):
        t = (target,) * 3 if isinstance(target, (int, float)) else tuple(target)
        self.target = tuple(float(v) for v in t)
        if image_interpolation not in ("bspline", "nearest"):
            raise NotImplementedError("image_interpolation: the reference uses 'bspline' (label maps: nearest)")
        self.image_interpolation = image_interpolation
        self.exclude = tuple(exclude or ())

    def __call__(self, subject: Dict) -> Dict:
        fy, fx, fz = self.target
        for k in _images(subject):
            if k in self.exclude:
                continue
            src = _vol(_get(subject, k))
            _, H, W, D = src.shape
            h, w, d = (int(lib().cddpm_resample_size(n, f)) for n, f in ((H, fy), (W, fx), (D, fz)))
            bspline = 1 if (k not in LABEL_KEYS and self.image_interpolation == "bspline") else 0
            out = torch.empty(1, h, w, d, dtype=torch.float32, device=src.device)
            with torch.cuda.device(src.device):
                nbytes = int(lib().cddpm_resample_workspace_bytes(H, W, D)) if bspline else 0
                ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=src.device)
                check(lib().cddpm_resample(ptr(src), H, W, D, fy, fx, fz, bspline, ptr(out), ptr(ws), nbytes,
                                           current_stream()), "cddpm_resample")
            _set(subject, k, out)
        return subject


class Compose:
    def __init__(self, transforms):
        self.transforms = list(transforms)

    def __call__(self, subject: Dict) -> Dict:
        for t in self.transforms:
            subject = t(subject)
        return subject


def get_transform(cfg) -> Compose:
    """create_dataset.get_transform (:196-218): the transforms applied once per volume before caching."""
    h, w, d = tuple(cfg.get("imageDim", (160, 192, 160)))
    exclude = ["vol_orig", "mask_orig", "seg_orig"] if not cfg.resizedEvaluation else None
    rescale = RescaleIntensity((0, 1), percentiles=(cfg.get("perc_low", 1), cfg.get("perc_high", 99)), masking_method="mask")
    resample = Resample(cfg.get("rescaleFactor", 3.0), image_interpolation="bspline", exclude=exclude)
    if cfg.get("unisotropic_sampling", True):
        return Compose([CropOrPad((h, w, d), padding_mode=0), rescale, resample])
    return Compose([rescale, resample])


class vol2slice(torch.utils.data.Dataset):
    """create_dataset.vol2slice (:143-193): one random axial slice of every volume (same constructor, same index rules,
    same torch.randint draws).  `ds[i]` must return a subject dict with 'vol' and 'mask' tensors [C,H,W,D]."""

    def __init__(self, ds, cfg, onlyBrain=False, slice=None, seq_slices=None):
        self.ds = ds
        self.onlyBrain = onlyBrain
        self.slice = slice
        self.seq_slices = seq_slices
        self.counter = 0
        self.ind = None
        self.cfg = cfg

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, index):
        subject = self.ds.__getitem__(index)
        vol, mask = _get(subject, "vol"), _get(subject, "mask")
        depth = vol.shape[-1]
        if self.onlyBrain:
            any_z = mask[0].reshape(-1, depth).ne(0).any(dim=0).cpu()
            start_ind = stop_ind = None
            for i in range(depth):
                if bool(any_z[i]) and start_ind is None:
                    start_ind = i
                if not bool(any_z[i]) and start_ind is not None:
                    stop_ind = i  # every later empty slice overwrites it, as in the reference's loop
            low, high = start_ind, stop_ind
        else:
            low, high = 0, depth
        if self.slice is not None:
            self.ind = self.slice
            if self.seq_slices is not None:
                low = self.ind
                high = self.ind + self.seq_slices
                self.ind = torch.randint(low, high, size=[1])
        else:
            if self.cfg.get("unique_slice", False):
                if self.counter % self.cfg.batch_size == 0 or self.ind is None:
                    self.ind = torch.randint(low, high, size=[1])
                self.counter = self.counter + 1
            else:
                self.ind = torch.randint(low, high, size=[1])
        subject["ind"] = self.ind
        ind = self.ind.to(vol.device) if torch.is_tensor(self.ind) else self.ind
        _set(subject, "vol", vol[..., ind])
        _set(subject, "mask", mask[..., ind])
        return subject
