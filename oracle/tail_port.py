"""ORACLE (test infrastructure, never imported by the product path).

numpy restatement of the anomaly-scoring tail of the reference (src/utils/utils_eval.py):
  residual |orig - reco|                       :29-33
  apply_brainmask_volume / apply_brainmask     :447-460   (per axial slice: cross erosion x (W//25), border 0, * diff)
  apply_3d_median_filter                       :462-464   (scipy median_filter size 5^3, mode 'reflect')
  find_best_val / dice                         :508-545
  compute_roc / compute_prc                    :548-557   (sklearn roc_curve+auc, average_precision_score)
  tpr / fpr                                    :566-575
  filter_3d_connected_components               :489-503   (skimage label + filled_area, restated; scikit-image absent)
  monai compute_hausdorff_distance             :134       (restated from monai's published source; monai absent)
  _test_step metric flow                       :18-194
Where the reference calls scipy / sklearn (both present in this image) the restatement is ALSO checked against
those libraries directly in tests/test_oracle_tail.py; against the live reference via tests/golden/tail_*.npz.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np


# ---------------------------------------------------------------------------------------------- stencils
def erode_cross(mask2d: np.ndarray, iterations: int) -> np.ndarray:
    """binary_erosion(mask, generate_binary_structure(2,1), iterations, border_value=0) == diamond of radius
    `iterations`; pixels outside the image count as background."""
    m = mask2d.astype(bool)
    done = 0
    while True:
        if iterations >= 1 and done == iterations:
            break
        p = np.pad(m, 1, mode="constant", constant_values=False)
        nxt = p[1:-1, 1:-1] & p[:-2, 1:-1] & p[2:, 1:-1] & p[1:-1, :-2] & p[1:-1, 2:]
        done += 1
        if iterations < 1 and np.array_equal(nxt, m):
            break  # scipy: iterations < 1 means "erode until nothing changes" (happens when W < 25)
        m = nxt
    return m


def apply_brainmask_volume(diff: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """diff, mask: [H, W, D].  Iterations = W // 25 (utils_eval.py:458: vol.squeeze().shape[1] // 25)."""
    it = diff.shape[1] // 25
    out = np.empty_like(diff)
    for s in range(diff.shape[2]):
        out[:, :, s] = erode_cross(mask[:, :, s] > 0, it) * diff[:, :, s]
    return out


def median_filter_3d(vol: np.ndarray, k: int = 5) -> np.ndarray:
    """scipy.ndimage.median_filter(vol, (k,k,k)) with mode='reflect' == symmetric padding by k//2 and, per voxel,
    the element of rank (k^3)//2 of the k^3 window."""
    r = k // 2
    p = np.pad(vol, r, mode="symmetric")
    win = np.lib.stride_tricks.sliding_window_view(p, (k, k, k)).reshape(*vol.shape, k * k * k)
    rank = (k * k * k) // 2
    return np.partition(win, rank, axis=-1)[..., rank].astype(vol.dtype)


# ---------------------------------------------------------------------------------------------- threshold search
def dice(P: np.ndarray, G: np.ndarray) -> float:
    psum = np.sum(P.flatten())
    gsum = np.sum(G.flatten())
    pg = np.sum(np.multiply(P.flatten(), G.flatten()))
    with np.errstate(invalid="ignore", divide="ignore"):
        return (2 * pg) / (psum + gsum)


def find_best_val(x: np.ndarray, y: np.ndarray, val_range=(0, 1), max_steps: int = 10) -> Tuple[float, float]:
    """Iterative form of the reference's recursive quartile bisection on Dice (utils_eval.py:508-539), keeping its
    `>=` tie-breaking (prefer the lower half; keep the later point on equal Dice) and the (lo, 1) degenerate range."""
    lo, hi = val_range
    best_val, best_pt = 0, 0
    for _ in range(max_steps):
        if lo == hi:
            hi = 1
        center = lo + (hi - lo) * 0.5
        q_lo = lo + (hi - lo) * 0.25
        q_hi = lo + (hi - lo) * 0.75
        v_lo = dice(x > q_lo, y)
        v_hi = dice(x > q_hi, y)
        if v_lo >= v_hi:
            if v_lo >= best_val:
                best_val, best_pt = v_lo, q_lo
            hi = center
        else:
            if v_hi >= best_val:
                best_val, best_pt = v_hi, q_hi
            lo = center
    return best_val, best_pt


# ---------------------------------------------------------------------------------------------- ranking metrics
def _cum_counts(scores: np.ndarray, labels: np.ndarray):
    """tps/fps at every distinct score, scores descending (sklearn _binary_clf_curve semantics)."""
    order = np.argsort(scores, kind="mergesort")[::-1]
    s = scores[order]
    l = labels[order].astype(np.float64)
    distinct = np.where(np.diff(s))[0]
    idx = np.r_[distinct, l.size - 1]
    tps = np.cumsum(l)[idx]
    fps = 1 + idx - tps
    return tps, fps


def roc_auc(scores: np.ndarray, labels: np.ndarray) -> float:
    tps, fps = _cum_counts(np.asarray(scores), np.asarray(labels).astype(int))
    tps = np.r_[0, tps]
    fps = np.r_[0, fps]
    with np.errstate(invalid="ignore", divide="ignore"):
        fpr = fps / fps[-1]
        tpr = tps / tps[-1]
    return float(np.trapezoid(tpr, fpr))


def average_precision(scores: np.ndarray, labels: np.ndarray) -> float:
    tps, fps = _cum_counts(np.asarray(scores), np.asarray(labels).astype(int))
    with np.errstate(invalid="ignore", divide="ignore"):
        precision = tps / (tps + fps)
        recall = tps / tps[-1]
    # sklearn: reversed curves with a final (P=1, R=0) point; AP = -sum(diff(recall) * precision[:-1])
    precision = np.r_[precision[::-1], 1.0]
    recall = np.r_[recall[::-1], 0.0]
    return float(-np.sum(np.diff(recall) * precision[:-1]))


def tpr(P, G):
    tp = np.sum(np.multiply(P.flatten(), G.flatten()))
    fn = np.sum(np.multiply(np.invert(P.flatten()), G.flatten()))
    with np.errstate(invalid="ignore", divide="ignore"):
        return tp / (tp + fn)


def fpr(P, G):
    """NB the reference's fpr() is FP / (FP + TP) (utils_eval.py:572-575) — reproduced as is."""
    tp = np.sum(np.multiply(P.flatten(), G.flatten()))
    fp = np.sum(np.multiply(P.flatten(), np.invert(G.flatten())))
    with np.errstate(invalid="ignore", divide="ignore"):
        return fp / (fp + tp)


def filter_small_components(vol: np.ndarray, max_size: int = 7) -> np.ndarray:
    """filter_3d_connected_components (utils_eval.py:489-503): skimage.measure.label(volume, connectivity=3), then every
    region whose regionprops `filled_area` <= 7 is zeroed.  scikit-image is NOT in this image (PARITY UNPINNED against
    it); restated from its published source: label(connectivity=ndim) = full 3x3x3 adjacency; RegionProperties.
    image_filled = scipy.ndimage.binary_fill_holes(region.image, structure=np.ones((3,) * ndim)) on the bounding-box
    crop, filled_area = its sum.  With the full structuring element the background leaks through every diagonal gap,
    so a hole needs a closed 26-voxel shell: for components this small filled_area equals area (asserted below)."""
    from scipy import ndimage

    vol = np.array(vol, dtype=bool)
    full = np.ones((3, 3, 3))
    lab, _ = ndimage.label(vol, structure=full)
    for i, sl in enumerate(ndimage.find_objects(lab)):
        if sl is None:
            continue
        region = lab[sl] == (i + 1)
        filled = int(ndimage.binary_fill_holes(region, structure=full).sum())
        if filled <= max_size:
            assert filled == int(region.sum())
            vol[sl][region] = False
    return vol


def mask_edges(mask: np.ndarray) -> np.ndarray:
    """monai.metrics.utils.get_mask_edges (monai 0.9, not installed here - PARITY UNPINNED, restated from the published
    source): edges = binary_erosion(mask) ^ mask with scipy's default cross structuring element and border_value 0.
    monai first crops both masks to the bounding box of their union; outside that box both are background, so the
    erosion sees the same zeros either way and the crop only shifts coordinates."""
    from scipy import ndimage

    mask = np.asarray(mask, dtype=bool)
    return ndimage.binary_erosion(mask) ^ mask


def _surface_distance(edges_pred: np.ndarray, edges_gt: np.ndarray) -> np.ndarray:
    """monai.metrics.utils.get_surface_distance(distance_metric='euclidean')."""
    from scipy import ndimage

    if not np.any(edges_gt):
        dis = np.inf * np.ones_like(edges_gt, dtype=np.float64)
    else:
        if not np.any(edges_pred):
            dis = np.inf * np.ones_like(edges_gt, dtype=np.float64)
            return np.asarray(dis[edges_gt])
        dis = ndimage.distance_transform_edt(~edges_gt)
    return np.asarray(dis[edges_pred])


def hausdorff_distance(pred: np.ndarray, gt: np.ndarray) -> float:
    """monai.metrics.compute_hausdorff_distance(pred[None, None], gt[None, None], include_background=False,
    distance_metric='euclidean', percentile=None, directed=False) for one binary volume pair (utils_eval.py:134):
    max of the two directed distances, each the max over one surface of the Euclidean distance to the other surface;
    nan when both masks are empty, inf when exactly one is."""
    ep, eg = mask_edges(pred), mask_edges(gt)

    def directed(a, b):
        sd = _surface_distance(a, b)
        return float("nan") if sd.shape == (0,) else float(sd.max())

    return max(directed(ep, eg), directed(eg, ep))


# ---------------------------------------------------------------------------------------------- _test_step flow
def volume_tail(reco: np.ndarray, orig: np.ndarray, seg: np.ndarray, mask: np.ndarray, *, stage: str,
                threshold_total: Optional[float] = None, erode: bool = True, median: bool = True,
                kernelsize_median: int = 5) -> Dict[str, object]:
    """Everything _test_step computes for one pathological-set volume with resizedEvaluation=True, evalSeg=True,
    threshold='auto'.  Inputs [H, W, D] float32; returns the per-volume scalars plus the intermediate volumes."""
    reco = reco.astype(np.float32)
    orig = orig.astype(np.float32)
    out: Dict[str, object] = {}
    diff = np.abs(orig - reco)
    d = reco - orig
    segb = seg > 0
    out["l1recoErrorAll"] = float(np.mean(np.abs(d), dtype=np.float32))
    out["l2recoErrorAll"] = float(np.mean(d * d, dtype=np.float32))
    with np.errstate(invalid="ignore", divide="ignore"):
        out["l1recoErrorUnhealthy"] = float(np.mean(np.abs(d[segb]), dtype=np.float32)) if segb.any() else float("nan")
        out["l1recoErrorHealthy"] = float(np.mean(np.abs(d[~segb]), dtype=np.float32))
        out["l2recoErrorUnhealthy"] = float(np.mean(d[segb] ** 2, dtype=np.float32)) if segb.any() else float("nan")
        out["l2recoErrorHealthy"] = float(np.mean(d[~segb] ** 2, dtype=np.float32))
    maskb = (mask > 0).astype(np.float32)
    if erode:
        diff = apply_brainmask_volume(diff, maskb)
    out["diff_masked"] = diff.copy()
    if median:
        diff = median_filter_3d(diff, kernelsize_median)
    out["diff_filtered"] = diff
    flat = diff.flatten()
    g = segb.flatten()
    out["AUC"] = roc_auc(flat, g)
    out["AUPRC"] = average_precision(flat, g)
    best_dice, best_thresh = find_best_val(flat, g, val_range=(0, np.max(diff)), max_steps=10)
    out["BestDice"], out["BestThreshold_own"] = best_dice, best_thresh
    if "test" in stage:
        best_thresh = threshold_total
    out["BestThreshold"] = best_thresh
    thr = diff > best_thresh
    thr = filter_small_components(thr)
    out["thresholded"] = thr
    out["Haus"] = hausdorff_distance(thr, segb)
    out["Dice"] = dice(thr, g)
    p = thr.flatten()
    # confusion_matrix(pred, truth).ravel() with the reference's (permuted) names (utils_eval.py:108)
    c00 = int(np.sum(~p & ~g))
    c01 = int(np.sum(~p & g))
    c10 = int(np.sum(p & ~g))
    c11 = int(np.sum(p & g))
    out["TP"], out["FP"], out["TN"], out["FN"] = c00, c01, c10, c11
    out["TPR"] = tpr(thr, g)
    out["FPR"] = fpr(thr, g)
    out["lesionSize"] = int(np.count_nonzero(g))
    with np.errstate(invalid="ignore", divide="ignore"):
        out["Accuracy"] = float(np.mean(p == g))
        out["Precision"] = c11 / (c11 + c10) if (c11 + c10) else 0.0
        out["Recall"] = c11 / (c11 + c01) if (c11 + c01) else 0.0
        out["Specificity"] = out["TN"] / (out["TN"] + out["FP"] + 0.0000001)
        inside = maskb > 0
        out["AnomalyScoreRecoPerVol"] = float(np.mean(diff[inside], dtype=np.float32)) if inside.any() else float("nan")
    # "per-slice" loops run over axis 0 (image rows), utils_eval.py:138-144, :160-165
    dice_rows, rows_scores, rows_labels = [], [], []
    for r in range(diff.shape[0]):
        gr = segb[r].flatten()
        if gr.any():
            dice_rows.append(dice((diff[r] > best_thresh), gr))
        m = maskb[r] > 0
        rows_scores.append(float(np.mean(diff[r][m], dtype=np.float32)) if m.any() else 0.0)
        rows_labels.append(1 if gr.any() else 0)
    out["DiceScorePerSlice"] = dice_rows
    out["AnomalyScoreRecoPerSlice"] = rows_scores
    out["labelPerSlice"] = rows_labels
    return out


# ---------------------------------------------------------------------------------------------- output image grid
_INFERNO8 = np.array([[0, 0, 4], [40, 11, 84], [101, 21, 110], [159, 42, 99], [212, 72, 66], [245, 125, 21],
                      [250, 193, 39], [252, 255, 164]], dtype=np.float32)


def compose_grid_port(panels: np.ndarray, ranges: np.ndarray) -> np.ndarray:
    """One row of log_images' figure (utils_eval.py:595-608) as an RGB array: panels [4,H,W] = original,
    reconstruction, difference, segmentation of one axial slice, ranges [4,2] = (vmin, vmax) of each panel's
    Normalize; every panel drawn as rot90(., 3), 'gray' except the difference ('inferno').  PARITY UNPINNED:
    matplotlib is absent from this image, so this restates imshow's documented pipeline - Normalize -> clip to [0, 1]
    -> colour table - at native resolution, with the inferno table interpolated linearly from matplotlib's 8-class
    palette instead of its 256 entries; it is not a pixel copy of the reference's resampled 1600 x 400 figure."""
    panels = np.asarray(panels, dtype=np.float32)
    ranges = np.asarray(ranges, dtype=np.float32)
    _, H, W = panels.shape
    out = np.zeros((W, 4 * H, 3), dtype=np.uint8)
    for k in range(4):
        img = np.rot90(panels[k], 3)  # [W, H]
        lo, hi = ranges[k]
        t = (img - lo) / (hi - lo) if hi > lo else np.zeros_like(img)
        t = np.clip(t, np.float32(0), np.float32(1)).astype(np.float32)
        if k == 2:
            u = t * np.float32(7)
            i = np.minimum(u.astype(np.int32), 6)
            fr = (u - i.astype(np.float32))[..., None]
            c = _INFERNO8[i] + fr * (_INFERNO8[i + 1] - _INFERNO8[i])
        else:
            c = np.repeat((t * np.float32(255))[..., None], 3, axis=2)
        out[:, k * H:(k + 1) * H, :] = np.rint(c).astype(np.uint8)
    return out
