"""Drop-in for the model-independent evaluation code of the reference, src/utils/utils_eval.py:
_test_step (:18-194), _test_end (:196-297), get_eval_dictionary (:324-445), apply_brainmask_volume (:454-460),
apply_3d_median_filter (:462-464), find_best_val (:508-539), dice (:540-545), compute_roc / compute_prc (:548-557),
filter_3d_connected_components (:489-503), tpr / fpr (:566-575).

Everything voxel-sized runs on the GPU (residual, eroded-mask multiply, 5x5x5 median, max, Dice counts for the
threshold bisection, thresholding, row statistics, sort-based AUC / AP): kernels cddpm_residual_erode,
cddpm_median3d, cddpm_max, cddpm_threshold_counts, cddpm_threshold_mask, cddpm_row_stats, cddpm_ranking_metrics, and
 the small-connected-component filter, the confusion counts of the filtered prediction and monai's
Hausdorff distance: cddpm_filter_small_components, cddpm_confusion_counts, cddpm_hausdorff.
The host keeps only scalar control flow (the bisection decisions, in numpy float32/float64 exactly like the reference).
PNG/wandb logging is out of scope.  The quirks of the reference are reproduced on purpose: per-"slice" loops run over image rows (axis 0), the
confusion-matrix names are permuted (:108), fpr() is FP/(FP+TP) (:572-575).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import CddpmError, VolView, check, current_stream, lib, ptr

_EVAL_KEYS = """IDs x reconstructions diffs diffs_volume Segmentation reconstructionTimes latentSpace Age AgeGroup
l1reconstructionErrors l1recoErrorAll l1recoErrorUnhealthy l1recoErrorHealthy l2recoErrorAll l2recoErrorUnhealthy
l2recoErrorHealthy l1reconstructionErrorMean l1reconstructionErrorStd l2reconstructionErrors l2reconstructionErrorMean
l2reconstructionErrorStd HausPerVol TPPerVol FPPerVol FNPerVol TNPerVol TPRPerVol FPRPerVol TPTotal FPTotal FNTotal
TNTotal TPRTotal FPRTotal PrecisionPerVol RecallPerVol PrecisionPerSlice RecallPerSlice lesionSizePerSlice
lesionSizePerVol Dice DiceScorePerSlice DiceScorePerVol BestDicePerVol BestThresholdPerVol AUCPerVol AUPRCPerVol
SpecificityPerVol AccuracyPerVol TPgradELBO FPgradELBO FNgradELBO TNgradELBO TPRgradELBO FPRgradELBO DicegradELBO
DiceScorePerVolgradELBO BestDicePerVolgradELBO BestThresholdPerVolgradELBO AUCPerVolgradELBO AUPRCPerVolgradELBO
KLD_to_learned_prior AUCAnomalyCombPerSlice AUPRCAnomalyCombPerSlice AnomalyScoreCombPerSlice AUCAnomalyKLDPerSlice
AUPRCAnomalyKLDPerSlice AnomalyScoreKLDPerSlice AUCAnomalyRecoPerSlice AUPRCAnomalyRecoPerSlice
AnomalyScoreRecoPerSlice AnomalyScoreRecoBinPerSlice AnomalyScoreAgePerSlice AUCAnomalyAgePerSlice
AUPRCAnomalyAgePerSlice labelPerSlice labelPerVol AnomalyScoreCombPerVol AnomalyScoreCombiPerVol
AnomalyScoreCombMeanPerVol AnomalyScoreRegPerVol AnomalyScoreRegMeanPerVol AnomalyScoreRecoPerVol
AnomalyScoreCombPriorPerVol AnomalyScoreCombiPriorPerVol AnomalyScoreAgePerVol AnomalyScoreRecoMeanPerVol
DiceScoreKLPerVol DiceScoreKLCombPerVol BestDiceKLCombPerVol BestDiceKLPerVol AUCKLCombPerVol AUPRCKLCombPerVol
AUCKLPerVol AUPRCKLPerVol TPKLCombPerVol FPKLCombPerVol TNKLCombPerVol FNKLCombPerVol TPRKLCombPerVol FPRKLCombPerVol
TPKLPerVol FPKLPerVol TNKLPerVol FNKLPerVol TPRKLPerVol FPRKLPerVol""".split()


def get_eval_dictionary():
    return {k: [] for k in _EVAL_KEYS}


# ------------------------------------------------------------------------------------------------ device helpers
def _as_cuda_f32(t, device=None) -> torch.Tensor:
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    if not torch.is_tensor(t):
        raise CddpmError("expected a tensor or ndarray")
    if not t.is_cuda:
        if device is None:
            raise CddpmError("the anomaly-scoring tail runs on CUDA tensors (there is no CPU path)")
        t = t.to(device)
    return t if t.dtype == torch.float32 else t.float()


def _view3(t: torch.Tensor) -> Tuple[VolView, Tuple[int, int, int], torch.Tensor]:
    """(view, (H, W, D), keep-alive) for a [.., H, W, D] tensor whose leading dims are singletons."""
    while t.dim() > 3 and t.shape[0] == 1:
        t = t[0]
    if t.dim() != 3:
        raise CddpmError(f"expected a [1,1,H,W,D] / [H,W,D] volume, got {tuple(t.shape)}")
    sy, sx, sd = t.stride()
    return VolView(t.data_ptr(), sy, sx, sd), tuple(t.shape), t


class _Volume:
    """Device state of one evaluated volume: filtered residual in [D,H,W] order plus the seg / mask views."""

    def __init__(self, diff_dhw: torch.Tensor, seg: torch.Tensor, mask: torch.Tensor, shape):
        self.diff = diff_dhw
        self.seg = seg
        self.mask = mask
        self.shape = shape  # (H, W, D)

    @property
    def seg_view(self):
        return _view3(self.seg)[0]

    @property
    def mask_view(self):
        return _view3(self.mask)[0]

    def diff_hwd(self) -> torch.Tensor:
        return self.diff.permute(1, 2, 0)


def trilinear_resize(volume, size) -> torch.Tensor:
    """F.interpolate(volume, size=size, mode="trilinear", align_corners=True) for a [1,1,H,W,D] (or [H,W,D]) CUDA
    volume, read in place through its strides (the UNet output arrives as a permuted [D,1,H,W] view); returns a
    contiguous [1,1,Ho,Wo,Do] fp32 tensor."""
    src = _as_cuda_f32(volume)
    view, (H, W, D), keep = _view3(src)
    Ho, Wo, Do = (int(v) for v in size)
    with torch.cuda.device(src.device):
        out = torch.empty(1, 1, Ho, Wo, Do, dtype=torch.float32, device=src.device)
        check(lib().cddpm_trilinear_resize(ctypes.byref(view), H, W, D, ptr(out), Ho, Wo, Do, current_stream()),
              "cddpm_trilinear_resize")
    del keep
    return out


def residual_and_filter(final_volume, data_orig, data_seg, data_mask, *, erode=True, median=True, kernelsize=5):
    """Steps (i)-(v) of _test_step: |orig - reco|, l1/l2 error sums, eroded-mask multiply, 3-D median.
    Returns (_Volume, sums[7] float64 numpy)."""
    orig = _as_cuda_f32(data_orig)
    dev = orig.device
    reco = _as_cuda_f32(final_volume, dev)
    seg = _as_cuda_f32(data_seg, dev)
    mask = _as_cuda_f32(data_mask, dev)
    vo, shape, orig = _view3(orig)
    vr, shape_r, reco = _view3(reco)
    vs, _, seg = _view3(seg)
    vm, _, mask = _view3(mask)
    if shape_r != shape or tuple(seg.shape) != shape or tuple(mask.shape) != shape:
        raise CddpmError(f"volume shapes differ: {shape} {shape_r} {tuple(seg.shape)} {tuple(mask.shape)}")
    H, W, D = shape
    with torch.cuda.device(dev):
        a = torch.empty(D, H, W, dtype=torch.float32, device=dev)
        sums = torch.empty(7, dtype=torch.float64, device=dev)
        check(lib().cddpm_residual_erode(ctypes.byref(vo), ctypes.byref(vr), ctypes.byref(vs), ctypes.byref(vm), H, W, D,
                                         W // 25, 1 if erode else 0, ptr(a), ptr(sums), current_stream()),
              "cddpm_residual_erode")
        if median:
            b = torch.empty_like(a)
            check(lib().cddpm_median3d(ptr(a), ptr(b), H, W, D, int(kernelsize), current_stream()), "cddpm_median3d")
            a = b
    return _Volume(a, seg, mask, shape), sums.cpu().numpy()


def apply_brainmask_volume(vol, mask_vol, erode=True, iterations=10):
    """[.. H, W, D] residual x eroded brain mask; like the reference, `erode`/`iterations` are ignored (W // 25 cross
    erosions are always applied) and the result is written back into `vol` when it is a tensor."""
    v = _as_cuda_f32(vol)
    m = _as_cuda_f32(mask_vol, v.device)
    vv, shape, v3 = _view3(v)
    vm, _, m3 = _view3(m)
    H, W, D = shape
    zero = torch.zeros_like(v3)
    vz, _, zero = _view3(zero)
    out = torch.empty(D, H, W, dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        # |v - 0| == v for the non-negative residuals this is applied to; negative inputs are handled by the sign fix
        check(lib().cddpm_residual_erode(ctypes.byref(vv), ctypes.byref(vz), None, ctypes.byref(vm), H, W, D, W // 25,
                                         1, ptr(out), None, current_stream()), "cddpm_residual_erode")
    res = out.permute(1, 2, 0)
    res = torch.where(v3 < 0, -res, res)
    if torch.is_tensor(vol) and vol.is_cuda:
        vol.squeeze().copy_(res)
        return vol
    return res


def apply_3d_median_filter(volume, kernelsize=5):
    """scipy.ndimage.median_filter(volume, (k,k,k)) (mode 'reflect') for a [H,W,D] CUDA tensor / ndarray."""
    is_np = isinstance(volume, np.ndarray)
    v = _as_cuda_f32(volume, "cuda" if is_np else None).squeeze()
    H, W, D = v.shape
    src = v.permute(2, 0, 1).contiguous()
    dst = torch.empty_like(src)
    with torch.cuda.device(v.device):
        check(lib().cddpm_median3d(ptr(src), ptr(dst), H, W, D, int(kernelsize), current_stream()), "cddpm_median3d")
    out = dst.permute(1, 2, 0)
    return out.cpu().numpy() if is_np else out


# ------------------------------------------------------------------------------------------------ threshold search
def dice(P, G):
    psum = np.sum(P.flatten())
    gsum = np.sum(G.flatten())
    pgsum = np.sum(np.multiply(P.flatten(), G.flatten()))
    with np.errstate(invalid="ignore", divide="ignore"):
        return (2 * pgsum) / (psum + gsum)


def _dice_counts(p, g, pg):
    with np.errstate(invalid="ignore", divide="ignore"):
        return (2 * np.int64(pg)) / (np.int64(p) + np.int64(g))


def _counts(vols: Sequence[_Volume], qs: Sequence[float], reduce_fn=None):
    """[#g, #p0, #pg0, #p1, #pg1] summed over `vols` (and over ranks when reduce_fn is given)."""
    dev = vols[0].diff.device
    with torch.cuda.device(dev):
        cnt = torch.zeros(1 + 2 * len(qs), dtype=torch.int64, device=dev)
        q = (ctypes.c_float * len(qs))(*[float(v) for v in qs])
        for v in vols:
            H, W, D = v.shape
            sv = v.seg_view
            check(lib().cddpm_threshold_counts(ptr(v.diff), ctypes.byref(sv), H, W, D, q, len(qs), ptr(cnt),
                                               current_stream()), "cddpm_threshold_counts")
    if reduce_fn is not None:
        reduce_fn(cnt)
    return cnt.cpu().numpy()


def _bisect(count_fn, val_range, max_steps):
    """find_best_val's control flow (utils_eval.py:508-539) with Dice evaluated from integer counts; the range
    arithmetic stays in numpy scalars so float32/float64 promotion matches the reference."""
    bottom, top = val_range
    max_val, max_point = 0, 0
    for _ in range(max_steps):
        if bottom == top:
            top = 1
        center = bottom + (top - bottom) * 0.5
        q_bottom = bottom + (top - bottom) * 0.25
        q_top = bottom + (top - bottom) * 0.75
        # the reference compares a float32 image against these scalars: the comparison happens in float32
        c = count_fn((np.float32(q_bottom), np.float32(q_top)))
        val_bottom = _dice_counts(c[1], c[0], c[2])
        val_top = _dice_counts(c[3], c[0], c[4])
        if val_bottom >= val_top:
            if val_bottom >= max_val:
                max_val, max_point = val_bottom, q_bottom
            top = center
        else:
            if val_top >= max_val:
                max_val, max_point = val_top, q_top
            bottom = center
    return max_val, max_point


def _device_max(vols: Sequence[_Volume], reduce_fn=None):
    dev = vols[0].diff.device
    best = None
    with torch.cuda.device(dev):
        out = torch.empty(1, dtype=torch.float32, device=dev)
        for v in vols:
            check(lib().cddpm_max(ptr(v.diff), v.diff.numel(), ptr(out), current_stream()), "cddpm_max")
            best = out.clone() if best is None else torch.maximum(best, out)
    if reduce_fn is not None:
        reduce_fn(best)
    return np.float32(best.item())


def find_best_val(x, y, val_range=(0, 1), max_steps=4, step=0, max_val=0, max_point=0):
    """Dice-optimal threshold by quartile bisection.  x: residual volume (CUDA tensor / ndarray, any shape), y: labels."""
    if step != 0 or max_val != 0 or max_point != 0:
        raise NotImplementedError("find_best_val is only ever entered at step 0 by the reference")
    xv = _as_cuda_f32(x, "cuda").reshape(1, 1, -1)
    yv = _as_cuda_f32(np.asarray(y, dtype=np.float32) if isinstance(y, np.ndarray) else y, xv.device).reshape(1, 1, -1)
    vol = _Volume(xv.reshape(-1, 1, 1).contiguous(), yv, yv, (1, 1, xv.numel()))
    return _bisect(lambda qs: _counts([vol], qs), val_range, max_steps)


# ------------------------------------------------------------------------------------------------ ranking metrics
def _binary_curve(scores: np.ndarray, labels: np.ndarray):
    order = np.argsort(scores, kind="mergesort")[::-1]
    s = scores[order]
    l = labels[order].astype(np.float64)
    idx = np.r_[np.where(np.diff(s))[0], l.size - 1]
    tps = np.cumsum(l)[idx]
    fps = 1 + idx - tps
    return fps, tps, s[idx]


def _roc_curve_host(scores, labels):
    """sklearn.metrics.roc_curve (drop_intermediate=True) restated for the small host-side uses."""
    fps, tps, thr = _binary_curve(np.asarray(scores, dtype=np.float64), np.asarray(labels).astype(int))
    if len(fps) > 2:
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        fps, tps, thr = fps[keep], tps[keep], thr[keep]
    tps = np.r_[0, tps]
    fps = np.r_[0, fps]
    thr = np.r_[np.inf, thr]
    with np.errstate(invalid="ignore", divide="ignore"):
        fpr = fps / fps[-1] if fps[-1] > 0 else np.repeat(np.nan, fps.shape)
        tpr_ = tps / tps[-1] if tps[-1] > 0 else np.repeat(np.nan, tps.shape)
    return fpr, tpr_, thr


def compute_roc(predictions, labels):
    """(auc, fpr, tpr, thresholds).  Large CUDA inputs use the device kernel for the AUC; the curve itself is only
    materialised on the host for small inputs (the reference never reads it for volumes)."""
    if torch.is_tensor(predictions) and predictions.is_cuda:
        auc_, _ = _ranking_device(predictions, labels)
        return auc_, None, None, None
    fpr, tpr_, thr = _roc_curve_host(np.asarray(predictions), np.asarray(labels))
    return float(np.trapezoid(tpr_, fpr)), fpr, tpr_, thr


def compute_prc(predictions, labels):
    if torch.is_tensor(predictions) and predictions.is_cuda:
        _, ap = _ranking_device(predictions, labels)
        return ap, None, None, None
    fps, tps, thr = _binary_curve(np.asarray(predictions, dtype=np.float64), np.asarray(labels).astype(int))
    with np.errstate(invalid="ignore", divide="ignore"):
        precision = tps / (tps + fps)
        recall = tps / tps[-1] if tps[-1] > 0 else np.ones_like(tps)
    precision = np.r_[precision[::-1], 1.0]
    recall = np.r_[recall[::-1], 0.0]
    return float(-np.sum(np.diff(recall) * precision[:-1])), precision, recall, thr[::-1]


def _ranking_device(x: torch.Tensor, labels) -> Tuple[float, float]:
    xv = x.float().reshape(-1, 1, 1).contiguous()
    lv = _as_cuda_f32(np.asarray(labels, dtype=np.float32) if isinstance(labels, np.ndarray) else labels, x.device)
    vol = _Volume(xv, lv.reshape(1, 1, -1), lv.reshape(1, 1, -1), (1, 1, xv.numel()))
    return _ranking_volume(vol)


def _ranking_volume(v: _Volume) -> Tuple[float, float]:
    H, W, D = v.shape
    n = H * W * D
    dev = v.diff.device
    with torch.cuda.device(dev):
        nbytes = int(lib().cddpm_ranking_workspace_bytes(n))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        res = torch.empty(2, dtype=torch.float64, device=dev)
        sv = v.seg_view
        check(lib().cddpm_ranking_metrics(ptr(v.diff), ctypes.byref(sv), H, W, D, ptr(ws), nbytes, ptr(res),
                                          current_stream()), "cddpm_ranking_metrics")
    r = res.cpu().numpy()
    return float(r[0]), float(r[1])


def _numpy_weak_scalars() -> bool:
    """True when NumPy keeps float32 scalars float32 in arithmetic with Python floats (NEP 50, NumPy >= 2) - the
    semantics the device-side bisection implements."""
    return (np.float32(3.0) * 0.5).dtype == np.float32


def _ranking_and_bisect(v: _Volume, max_steps: int = 10):
    """AUC, AUPRC (compute_roc / compute_prc) and find_best_val(val_range=(0, max), max_steps) of one volume with ONE
    host read: the bisection runs on the device over the ranking pass's sorted scores (cddpm_dice_bisect) - no pass over
    the volume per step, no host round trip per step.  Returns (AUC, AUPRC, bestDice, bestThresh)."""
    if not _numpy_weak_scalars() or os.environ.get("CDDPM_DEVICE_BISECT", "1") == "0":
        auc, auprc = _ranking_volume(v)
        top = _device_max([v])
        best, thr = _bisect(lambda qs: _counts([v], qs), (0, top), max_steps)
        return auc, auprc, best, thr
    H, W, D = v.shape
    n = H * W * D
    dev = v.diff.device
    with torch.cuda.device(dev):
        nbytes = int(lib().cddpm_ranking_workspace_bytes(n))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        res = torch.empty(5, dtype=torch.float64, device=dev)
        sv = v.seg_view
        check(lib().cddpm_ranking_metrics(ptr(v.diff), ctypes.byref(sv), H, W, D, ptr(ws), nbytes, ptr(res),
                                          current_stream()), "cddpm_ranking_metrics")
        check(lib().cddpm_dice_bisect(ptr(ws), n, int(max_steps), ptr(res[2:]), current_stream()), "cddpm_dice_bisect")
    r = res.cpu().numpy()
    # the reference returns the ints (0, 0) when no step improved on max_val = 0 (all-NaN Dice: empty prediction and
    # empty label); otherwise a float64 Dice and a float32 threshold - a Python float when max(x) == 0 sent the range
    # arithmetic through the Python ints (0, 1)
    if r[2] == 0 and r[3] == 0:
        return float(r[0]), float(r[1]), 0, 0
    thr = float(r[3]) if r[4] == 0 else np.float32(r[3])
    return float(r[0]), float(r[1]), r[2], thr


def tpr(P, G):
    tp = np.sum(np.multiply(P.flatten(), G.flatten()))
    fn = np.sum(np.multiply(np.invert(P.flatten()), G.flatten()))
    with np.errstate(invalid="ignore", divide="ignore"):
        return tp / (tp + fn)


def fpr(P, G):
    tp = np.sum(np.multiply(P.flatten(), G.flatten()))
    fp = np.sum(np.multiply(P.flatten(), np.invert(G.flatten())))
    with np.errstate(invalid="ignore", divide="ignore"):
        return fp / (fp + tp)


def _filter_small_components_device(mask_dhw: torch.Tensor, max_size: int = 7) -> torch.Tensor:
    """uint8 [D,H,W] on the GPU -> filtered uint8 [D,H,W] (cddpm_filter_small_components)."""
    D, H, W = mask_dhw.shape
    out = torch.empty_like(mask_dhw)
    with torch.cuda.device(mask_dhw.device):
        check(lib().cddpm_filter_small_components(ptr(mask_dhw), ptr(out), H, W, D, int(max_size), current_stream()),
              "cddpm_filter_small_components")
    return out


def filter_3d_connected_components(volume):
    """26-connected components whose hole-filled size is <= 7 are removed (utils_eval.py:489-503): scikit-image fills
    holes with a full 3x3x3 element, so for components this small `filled_area` is the voxel count and the filter is
    one GPU kernel in which every foreground voxel walks its own component up to the 8th voxel (csrc/tail_cc.cu).
    Accepts what the reference passes (bool numpy array or tensor, 3-D, or 4-D folded to [a*b, c, d]); returns the
    same kind on the same device."""
    is_np = isinstance(volume, np.ndarray)
    t = torch.as_tensor(volume)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise CddpmError("filter_3d_connected_components needs a CUDA device (no CPU fallback)")
        t = t.cuda()
    sz = None
    if t.ndim > 3:
        sz = t.shape
        t = t.reshape(sz[0] * sz[1], sz[2], sz[3])
    if t.ndim != 3:
        raise ValueError("filter_3d_connected_components expects a 3-D (or 4-D) volume")
    out = _filter_small_components_device((t != 0).to(torch.uint8).contiguous()).to(torch.bool)
    if sz is not None:
        out = out.reshape(sz)
    return out.cpu().numpy() if is_np else out


def _hausdorff_device(pred_dhw: torch.Tensor, v: "_Volume") -> float:
    """monai compute_hausdorff_distance(pred, seg>0) (utils_eval.py:134) from cddpm_hausdorff's integer results."""
    H, W, D = v.shape
    dev = pred_dhw.device
    with torch.cuda.device(dev):
        nbytes = int(lib().cddpm_hausdorff_workspace_bytes(H, W, D))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        res = torch.empty(4, dtype=torch.int64, device=dev)
        sv = v.seg_view
        check(lib().cddpm_hausdorff(ptr(pred_dhw), ctypes.byref(sv), H, W, D, ptr(ws), nbytes, ptr(res),
                                    current_stream()), "cddpm_hausdorff")
    return hausdorff_from_counts(res.cpu().numpy())


def hausdorff_from_counts(r) -> float:
    """[max d^2 pred->seg, max d^2 seg->pred, #surface(pred), #surface(seg)] -> monai's value: nan when both masks are
    empty, inf when exactly one is, else the float64 square root of the larger squared distance."""
    if r[2] == 0 and r[3] == 0:
        return float("nan")
    if r[2] == 0 or r[3] == 0:
        return float("inf")
    return float(np.sqrt(np.float64(max(int(r[0]), int(r[1])))))


def compute_hausdorff_distance(pred, seg) -> float:
    """Public form of the per-volume Hausdorff distance for [H,W,D] volumes (prediction > 0 vs seg > 0)."""
    if not torch.cuda.is_available():
        raise CddpmError("compute_hausdorff_distance needs a CUDA device (no CPU fallback)")
    p = _as_cuda_f32(pred, torch.device("cuda", torch.cuda.current_device()))
    g = _as_cuda_f32(seg, p.device)
    if p.ndim != 3 or g.shape != p.shape:
        raise ValueError("compute_hausdorff_distance expects two [H,W,D] volumes of the same shape")
    vol = _Volume(p.permute(2, 0, 1).contiguous(), g, g, tuple(p.shape))
    return _hausdorff_device((vol.diff != 0).to(torch.uint8).contiguous(), vol)


def _dist_sum(t: torch.Tensor):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def _dist_max(t: torch.Tensor):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)


# ---------------------------------------------------------------------------------------------- output image grids
class _ImageWriter:
    """PNG encoding and file writes off the critical path: `_test_step` only launches the grid kernel and an
    asynchronous device-to-host copy into pinned memory; ONE background thread waits for the copy's event, encodes the
    PNG (PIL) and writes it.  `flush()` (called by `_test_end`) waits for everything queued."""

    def __init__(self):
        import queue
        import threading

        self._q = queue.Queue()
        self._errors: List[BaseException] = []
        self._t = threading.Thread(target=self._run, name="cddpm-png-writer", daemon=True)
        self._t.start()

    def _run(self):
        from PIL import Image

        while True:
            event, host, path = self._q.get()
            try:
                event.synchronize()
                Image.fromarray(host.numpy()).save(path)
            except BaseException as e:  # surfaced by flush(): a lost image must not pass silently
                self._errors.append(e)
            finally:
                self._q.task_done()

    def submit(self, event, host, path):
        self._q.put((event, host, path))

    def flush(self):
        self._q.join()
        if self._errors:
            e = self._errors[0]
            self._errors.clear()
            raise CddpmError(f"writing an output image failed: {e!r}")


_image_writer: Optional[_ImageWriter] = None


def flush_image_writer():
    """Block until every image queued by log_images has been written (no-op when none was)."""
    if _image_writer is not None:
        _image_writer.flush()


def log_images(self, diff_volume, data_orig, data_seg, data_mask, final_volume, ID, diff_volume_KL=None, flow=None):
    """utils_eval.log_images (:586-628): for every 10th axial slice one image `grid/{ID}_{j}_Grid.png` under the working
    directory with four panels side by side - original, reconstruction (both 'gray', each normalised to its own range),
    difference ('inferno', 0 .. max of the whole difference volume + 0.01), segmentation ('gray') - each drawn as
    rot90(., 3).  The panels are composed on the device (`cddpm_compose_grid`), copied to pinned memory asynchronously and
    encoded / written by a background thread.  Differences from the reference, by necessity: matplotlib is not
    available, so the image is the four panels at native resolution (H x 4W pixels) without matplotlib's figure
    resampling and the inferno table is interpolated from its 8-class palette; the wandb upload (:625) is not done."""
    global _image_writer
    try:
        import PIL  # noqa: F401
    except ImportError as e:
        raise CddpmError("saveOutputImages needs Pillow to encode PNG files") from e
    if not torch.cuda.is_available():
        raise CddpmError("log_images needs a CUDA device (no CPU fallback)")
    dev = diff_volume.device if isinstance(diff_volume, torch.Tensor) and diff_volume.is_cuda else \
        torch.device("cuda", torch.cuda.current_device())
    d = _as_cuda_f32(diff_volume, dev).squeeze()
    o = _as_cuda_f32(data_orig, dev).squeeze()
    f = _as_cuda_f32(final_volume, dev).squeeze()
    g = _as_cuda_f32(data_seg, dev).squeeze()
    if d.ndim != 3 or o.shape != d.shape or f.shape != d.shape or g.shape != d.shape:
        raise ValueError("log_images expects [H,W,D] volumes of one shape")
    H, W, D = d.shape
    grid_dir = os.path.join(os.getcwd(), "grid")
    os.makedirs(grid_dir, exist_ok=True)
    if _image_writer is None:
        import atexit

        _image_writer = _ImageWriter()
        atexit.register(flush_image_writer)  # a run that never reaches _test_end still gets its files
    name = ID[0] if isinstance(ID, (list, tuple)) else ID
    with torch.cuda.device(dev):
        vmax = d.max() + 0.01
        for j in range(0, D, 10):
            panels = torch.stack([o[..., j], f[..., j], d[..., j], g[..., j]]).contiguous()
            ranges = torch.stack([panels.amin(dim=(1, 2)), panels.amax(dim=(1, 2))], dim=1).contiguous()
            ranges[2, 0] = 0.0
            ranges[2, 1] = vmax
            rgb = torch.empty(W, 4 * H, 3, dtype=torch.uint8, device=dev)
            check(lib().cddpm_compose_grid(ptr(panels), ptr(ranges), H, W, ptr(rgb), current_stream()),
                  "cddpm_compose_grid")
            if self.cfg.get("save_to_disc", True):
                host = torch.empty(rgb.shape, dtype=torch.uint8, pin_memory=True)
                host.copy_(rgb, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                _image_writer.submit(ev, host, os.path.join(grid_dir, "{}_{}_Grid.png".format(name, j)))


def _test_step(self, final_volume, data_orig, data_seg, data_mask, batch_idx, ID, label_vol):
    self.healthy_sets = ["IXI"]
    cfg = self.cfg
    if not cfg.resizedEvaluation:  # full-resolution evaluation (utils_eval.py:24-25)
        final_volume = trilinear_resize(final_volume, self.new_size)
    vol, sums = residual_and_filter(final_volume, data_orig, data_seg, data_mask, erode=bool(cfg["erodeBrainmask"]),
                                    median=bool(cfg["medianFiltering"]), kernelsize=cfg.get("kernelsize_median", 5))
    H, W, D = vol.shape
    n = H * W * D
    n_les = sums[6]
    ed = self.eval_dict
    with np.errstate(invalid="ignore", divide="ignore"):
        ed["l1recoErrorAll"].append(float(sums[0] / n))
        ed["l1recoErrorUnhealthy"].append(float(np.float64(sums[2]) / n_les))
        ed["l1recoErrorHealthy"].append(float(sums[4] / (n - n_les)))
        ed["l2recoErrorAll"].append(float(sums[1] / n))
        ed["l2recoErrorUnhealthy"].append(float(np.float64(sums[3]) / n_les))
        ed["l2recoErrorHealthy"].append(float(sums[5] / (n - n_les)))
    if cfg["saveOutputImages"]:  # utils_eval.py:72-74 (after the median filter, before the metrics)
        log_images(self, vol.diff.permute(1, 2, 0), data_orig, data_seg, data_mask, final_volume, ID)

    dev = vol.diff.device
    best_thresh = None
    if cfg.evalSeg and self.dataset[0] not in self.healthy_sets:
        AUC, AUPRC, bestDice, bestThresh = _ranking_and_bisect(vol, 10)
        if "test" in self.stage:
            bestThresh = self.threshold["total"]
        thr = bestThresh if cfg["threshold"] == "auto" else cfg["threshold"]
        with torch.cuda.device(dev):
            tmask = torch.empty(D, H, W, dtype=torch.uint8, device=dev)
            check(lib().cddpm_threshold_mask(ptr(vol.diff), n, float(np.float32(thr)), ptr(tmask), current_stream()),
                  "cddpm_threshold_mask")
        with torch.cuda.device(dev):
            if "node" not in self.dataset[0].lower():
                tmask = _filter_small_components_device(tmask)
            cc = torch.zeros(3, dtype=torch.int64, device=dev)
            sv = vol.seg_view
            check(lib().cddpm_confusion_counts(ptr(tmask), ctypes.byref(sv), H, W, D, ptr(cc), current_stream()),
                  "cddpm_confusion_counts")
        haus = _hausdorff_device(tmask, vol)
        c11, c10, c01 = (int(x) for x in cc.cpu().numpy())
        c00 = int(n - c11 - c10 - c01)
        diceScore = _dice_counts(c11 + c10, c11 + c01, c11)
        # confusion_matrix(pred, truth).ravel() unpacked as TP, FP, TN, FN by the reference (:108)
        TP, FP, TN, FN = c00, c01, c10, c11
        with np.errstate(invalid="ignore", divide="ignore"):
            TPR = np.int64(c11) / (np.int64(c11) + np.int64(c01))
            FPR = np.int64(c10) / (np.int64(c10) + np.int64(c11))
        ed["lesionSizePerVol"].append(c11 + c01)
        ed["DiceScorePerVol"].append(diceScore)
        ed["BestDicePerVol"].append(bestDice)
        ed["BestThresholdPerVol"].append(bestThresh)
        ed["AUCPerVol"].append(AUC)
        ed["AUPRCPerVol"].append(AUPRC)
        ed["TPPerVol"].append(TP)
        ed["FPPerVol"].append(FP)
        ed["TNPerVol"].append(TN)
        ed["FNPerVol"].append(FN)
        ed["TPRPerVol"].append(TPR)
        ed["FPRPerVol"].append(FPR)
        ed["IDs"].append(ID[0])
        ed["AccuracyPerVol"].append((c11 + c00) / n)
        ed["PrecisionPerVol"].append(c11 / (c11 + c10) if (c11 + c10) else 0.0)
        ed["RecallPerVol"].append(c11 / (c11 + c01) if (c11 + c01) else 0.0)
        ed["SpecificityPerVol"].append(TN / (TN + FP + 0.0000001))
        ed["HausPerVol"].append(haus)
        best_thresh = bestThresh

    # row ("slice") statistics in one pass: Dice / precision / recall per row with lesion, masked mean residual per row
    thr_rows = float(np.float32(best_thresh)) if best_thresh is not None else float("inf")
    with torch.cuda.device(dev):
        rows = torch.empty(H, 4, dtype=torch.int64, device=dev)
        rowsum = torch.empty(H, dtype=torch.float64, device=dev)
        sv, mv = vol.seg_view, vol.mask_view
        check(lib().cddpm_row_stats(ptr(vol.diff), ctypes.byref(sv), ctypes.byref(mv), H, W, D, thr_rows, ptr(rows),
                                    ptr(rowsum), current_stream()), "cddpm_row_stats")
    rows_h = rows.cpu().numpy()
    rowsum_h = rowsum.cpu().numpy()
    if best_thresh is not None:
        for y in range(H):
            p, g, pg = (int(v) for v in rows_h[y, :3])
            if g > 0:
                ed["DiceScorePerSlice"].append(_dice_counts(p, g, pg))
                ed["PrecisionPerSlice"].append(pg / p if p else 0.0)
                ed["RecallPerSlice"].append(pg / g)
                ed["lesionSizePerSlice"].append(g)

    if "val" in self.stage:
        if batch_idx == 0 or not isinstance(self.diffs_list, list):
            self.diffs_list, self.seg_list = [], []
        self.diffs_list.append(vol)
        self.seg_list.append(vol.seg)

    n_mask = int(rows_h[:, 3].sum())
    if cfg.get("use_postprocessed_score", True):
        score_vol = float(rowsum_h.sum() / n_mask) if n_mask else float("nan")
    # reconstruction-based anomaly score per image row (0.0 where the row holds no brain), utils_eval.py:158-183
    score_rows = [float(rowsum_h[y] / rows_h[y, 3]) if rows_h[y, 3] else 0.0 for y in range(H)]
    label = [1 if rows_h[y, 1] > 0 else 0 for y in range(H)]
    if self.dataset[0] not in self.healthy_sets:
        with _quiet():
            AUCs, _, _, _ = compute_roc(np.array(score_rows), np.array(label))
            AUPRCs, _, _, _ = compute_prc(np.array(score_rows), np.array(label))
        ed["AUCAnomalyRecoPerSlice"].append(AUCs)
        ed["AUPRCAnomalyRecoPerSlice"].append(AUPRCs)
        ed["labelPerSlice"].extend(label)
        ed["AnomalyScoreRecoPerSlice"].extend(score_rows)
    if cfg.get("use_postprocessed_score", True):
        for k in ("AnomalyScoreRecoPerVol", "AnomalyScoreCombPerVol", "AnomalyScoreCombiPerVol",
                  "AnomalyScoreCombPriorPerVol", "AnomalyScoreCombiPriorPerVol"):
            ed[k].append(score_vol)
    ed["labelPerVol"].append(label_vol.item() if hasattr(label_vol, "item") else label_vol)
    return vol


_MEAN_STD = [
    ("l1recoErrorAll", "l1recoErrorAll", True), ("l2recoErrorAll", "l2recoErrorAll", True),
    ("l1recoErrorHealthy", "l1recoErrorHealthy", True), ("l1recoErrorUnhealthy", "l1recoErrorUnhealthy", True),
    ("l2recoErrorHealthy", "l2recoErrorHealthy", True), ("l2recoErrorUnhealthy", "l2recoErrorUnhealthy", True),
    ("AUPRCPerVol", "AUPRCPerVol", True), ("AUCPerVol", "AUCPerVol", True), ("DicePerVol", "DiceScorePerVol", True),
    ("BestDicePerVol", "BestDicePerVol", False), ("BestThresholdPerVol", "BestThresholdPerVol", False),
    ("TPPerVol", "TPPerVol", True), ("FPPerVol", "FPPerVol", True), ("TNPerVol", "TNPerVol", True),
    ("FNPerVol", "FNPerVol", True), ("TPRPerVol", "TPRPerVol", True), ("FPRPerVol", "FPRPerVol", True),
]
_MEAN_STD_PLAIN = ["PrecisionPerVol", "RecallPerVol", "PrecisionPerSlice", "RecallPerSlice", "AccuracyPerVol",
                   "SpecificityPerVol"]


def _test_end(self):
    flush_image_writer()
    ed = self.eval_dict
    with np.errstate(invalid="ignore", divide="ignore"), _quiet():
        for out, src, nan_aware in _MEAN_STD:
            vals = ed[src]
            ed[out + "Mean"] = (np.nanmean if nan_aware else np.mean)(vals)
            ed[out + "Std"] = (np.nanstd if nan_aware else np.std)(vals)
        haus = np.array(ed["HausPerVol"], dtype=np.float64)
        haus = haus[np.isfinite(haus)]
        ed["HausPerVolMean"] = np.nanmean(haus)
        ed["HausPerVolStd"] = np.nanstd(haus)
        for k in _MEAN_STD_PLAIN:
            ed[k + "Mean"] = np.mean(ed[k])
            ed[k + "Std"] = np.std(ed[k])
    if "test" in self.stage:
        del self.threshold
    if "val" in self.stage:
        vols: List[_Volume] = list(self.diffs_list) if isinstance(self.diffs_list, list) else []
        if self.dataset[0] not in self.healthy_sets:
            # global Dice-optimal threshold over every validation voxel; counts are summed over volumes and, in a
            # sharded sweep, over ranks — the same decisions as the reference's concatenated arrays (:262-271)
            top = _device_max(vols, _dist_max)
            _, bestThresh = _bisect(lambda qs: _counts(vols, qs, _dist_sum), (0, top), 10)
            self.threshold["total"] = bestThresh
            if self.cfg.get("KLDBackprop", False):
                raise NotImplementedError("KLDBackprop thresholds belong to other model families of the reference")
        else:
            diffs = np.concatenate([v.diff.permute(1, 2, 0).reshape(-1).cpu().numpy() for v in vols])
            fpr_h, _, threshs = _roc_curve_host(diffs, np.zeros_like(diffs, dtype=int))
            self.threshholds_healthy = {
                "thresh_1p": threshs[np.argmax(fpr_h > 0.01)],
                "thresh_5p": threshs[np.argmax(fpr_h > 0.05)],
                "thresh_10p": threshs[np.argmax(fpr_h > 0.10)],
            }
            ed["t_1p"] = self.threshholds_healthy["thresh_1p"]
            ed["t_5p"] = self.threshholds_healthy["thresh_5p"]
            ed["t_10p"] = self.threshholds_healthy["thresh_10p"]


class _quiet:
    def __enter__(self):
        import warnings

        self._cm = warnings.catch_warnings()
        self._cm.__enter__()
        warnings.simplefilter("ignore")

    def __exit__(self, *a):
        return self._cm.__exit__(*a)
