"""ORACLE (test infrastructure, never imported by the product path).

fp32 functional restatement of the condition encoder: ResNet-50 v1.5 (stride on the 3x3), 1 input channel,
fc -> cond_dim, eval mode (BatchNorm running statistics, no DropPath):
  SparK_2D_encoder.forward            src/models/modules/spark/Spark_2D.py:285-290
  patched ResNet.forward(pyramid=0)   src/models/modules/spark/resnet.py:13-46
  build_encoder                       src/models/modules/spark/models.py:89-109
The arithmetic itself lives in timm==0.6.7 (environment.yml:72), which is NOT vendored in the reference and not
installed here: PARITY UNPINNED against real timm; pinned against the torchvision-backed stand-in of oracle/stubs/timm
(same topology and parameter names) by tests/golden/encoder_*.npz.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

LAYERS = (3, 4, 6, 3)
WIDTHS = (64, 128, 256, 512)
EXPANSION = 4


def param_shapes(cond_dim: int = 128, in_chans: int = 1) -> List[Tuple[str, Tuple[int, ...]]]:
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def bn(p, c):
        out.extend([(p + ".weight", (c,)), (p + ".bias", (c,)), (p + ".running_mean", (c,)),
                    (p + ".running_var", (c,)), (p + ".num_batches_tracked", ())])

    out.append(("conv1.weight", (64, in_chans, 7, 7)))
    bn("bn1", 64)
    cin = 64
    for li, (n, w) in enumerate(zip(LAYERS, WIDTHS)):
        for bi in range(n):
            p = f"layer{li + 1}.{bi}"
            out.append((p + ".conv1.weight", (w, cin, 1, 1)))
            bn(p + ".bn1", w)
            out.append((p + ".conv2.weight", (w, w, 3, 3)))
            bn(p + ".bn2", w)
            out.append((p + ".conv3.weight", (w * EXPANSION, w, 1, 1)))
            bn(p + ".bn3", w * EXPANSION)
            if bi == 0:
                out.append((p + ".downsample.0.weight", (w * EXPANSION, cin, 1, 1)))
                bn(p + ".downsample.1", w * EXPANSION)
            cin = w * EXPANSION
    out.extend([("fc.weight", (cond_dim, 512 * EXPANSION)), ("fc.bias", (cond_dim,))])
    return out


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def resnet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """[B,1,H,W] -> [B,cond_dim]."""
    x = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], None, stride=2, padding=3), sd, "bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, n in enumerate(LAYERS):
        for bi in range(n):
            p = f"layer{li + 1}.{bi}"
            stride = 2 if (bi == 0 and li > 0) else 1
            idt = x
            h = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
            h = F.relu(_bn(F.conv2d(h, sd[p + ".conv2.weight"], None, stride=stride, padding=1), sd, p + ".bn2"))
            h = _bn(F.conv2d(h, sd[p + ".conv3.weight"]), sd, p + ".bn3")
            if bi == 0:
                idt = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride=stride), sd, p + ".downsample.1")
            x = F.relu(h + idt)
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)
    return F.linear(x, sd["fc.weight"], sd["fc.bias"])
