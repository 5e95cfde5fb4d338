"""Stand-in for monai.metrics.compute_hausdorff_distance (utils_eval.py:134; monai is not installed here): the
restatement in oracle/tail_port.py (edges by erosion, scipy EDT, max of the two directed distances) for [B, C, ...]
one-hot inputs.  PARITY UNPINNED against real monai."""
import numpy as np
import torch


class metrics:
    @staticmethod
    def compute_hausdorff_distance(y_pred, y, include_background=False, distance_metric="euclidean", percentile=None,
                                   directed=False, **k):
        from oracle.tail_port import hausdorff_distance

        assert distance_metric == "euclidean" and percentile is None and not directed
        yp, yt = np.asarray(y_pred), np.asarray(y)
        if not include_background and yp.shape[1] > 1:
            yp, yt = yp[:, 1:], yt[:, 1:]
        out = np.empty(yp.shape[:2], dtype=np.float64)
        for b in range(yp.shape[0]):
            for c in range(yp.shape[1]):
                p_ = yp[b, c] if yp.dtype == bool else yp[b, c] == 1
                g_ = yt[b, c] if yt.dtype == bool else yt[b, c] == 1
                out[b, c] = hausdorff_distance(p_, g_)
        return torch.from_numpy(out)
