"""Condition encoder — drop-in for src.models.modules.DDPM_encoder.get_encoder (DDPM_encoder.py:6-29) and
src.models.modules.spark.Spark_2D.SparK_2D_encoder (spark/Spark_2D.py:268-290).

The reference builds `timm.create_model('resnet50', in_chans=1, num_classes=cond_dim, drop_path_rate=0.05)`
(spark/models.py:89-109) and calls its monkey-patched forward(x, pyramid=0) (spark/resnet.py:13-46).  timm is not
vendored in the reference; the modules below are parameter holders with timm's / torchvision's ResNet-50 key names
(conv1, bn1, layerN.i.{conv1,bn1,conv2,bn2,conv3,bn3,downsample.{0,1}}, fc — 320 state_dict entries), initialised like
timm (kaiming-normal convs, zero_init_last on bn3).  forward runs in the CUDA engine (cddpm_encoder_forward) with
BatchNorm folded from the running statistics, i.e. eval-mode arithmetic; there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import CddpmError, check, current_stream, lib, ptr

_LAYERS = {"resnet50": (3, 4, 6, 3)}
_WIDTHS = (64, 128, 256, 512)


def _no_eager(self, *a, **k):
    raise CddpmError("this module is a parameter holder; the encoder executes inside the CUDA engine")


class _Bottleneck(nn.Module):
    def __init__(self, cin, width, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, width * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(width * 4)
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(cin, width * 4, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(width * 4))
        else:
            self.downsample = None
        nn.init.zeros_(self.bn3.weight)  # timm zero_init_last

    forward = _no_eager


class ResNet(nn.Module):
    """ResNet-50 parameter holder + CUDA engine handle."""

    def __init__(self, name="resnet50", in_chans=1, num_classes=128, image_size=(96, 96), engine_dtype=torch.float16):
        super().__init__()
        if name not in _LAYERS:
            raise NotImplementedError(f"encoder backbone {name}: only resnet50 (the cDDPM configuration) is built")
        if in_chans != 1:
            raise NotImplementedError("the cDDPM encoder takes single-channel slices")
        self.num_classes = num_classes
        self.image_size = tuple(int(s) for s in image_size)
        self.engine_dtype = engine_dtype
        self.drop_rate = 0.0
        # timm's stochastic depth (the reference builds resnet50 with drop_path_rate=0.05, spark/models.py:50,92-109):
        # block i of the 16 bottlenecks drops its residual branch with probability rate * i / 15 in training mode
        self.drop_path_rate = 0.05
        self.conv1 = nn.Conv2d(in_chans, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for li, (n, w) in enumerate(zip(_LAYERS[name], _WIDTHS)):
            blocks = []
            for bi in range(n):
                blocks.append(_Bottleneck(cin, w, 2 if (bi == 0 and li > 0) else 1, bi == 0))
                cin = w * 4
            setattr(self, f"layer{li + 1}", nn.Sequential(*blocks))
        self.fc = nn.Linear(cin, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self._h: Optional[ctypes.c_void_p] = None
        self._versions = None
        self._items = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                lib().cddpm_encoder_destroy(h)
            except Exception:
                pass
        th = self.__dict__.get("_th")
        if th is not None and th.value:
            try:
                lib().cddpm_encoder_train_destroy(th)
            except Exception:
                pass

    def _apply(self, fn, *args, **kwargs):
        self._items = None
        self._versions = None
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode: bool = True):
        # The training forward updates running_mean / running_var inside CUDA-graph replays, which do not advance the
        # tensors' _version (the change signal _sync() uses): re-push everything after every mode change.
        self._versions = None
        return super().train(mode)

    def _engine_items(self) -> List[Tuple[str, torch.Tensor]]:
        if self._items is None:
            self._items = [(k, v) for k, v in self.state_dict(keep_vars=True).items()
                           if not k.endswith("num_batches_tracked")]
        return self._items

    def _sync(self):
        items = self._engine_items()
        dev = items[0][1].device
        if dev.type != "cuda":
            raise CddpmError("the encoder is on the CPU; the cDDPM engine has no CPU path — move the module to CUDA")
        if self._h is None:
            h = ctypes.c_void_p()
            check(lib().cddpm_encoder_create(self.image_size[0], self.image_size[1], self.num_classes,
                                             _lib.fmt_of(self.engine_dtype), ctypes.byref(h)), "cddpm_encoder_create")
            self._h = h
            n = lib().cddpm_encoder_param_count(h)
            name, numel = ctypes.c_char_p(), ctypes.c_int64()
            names = []
            for i in range(n):
                check(lib().cddpm_encoder_param_info(h, i, ctypes.byref(name), ctypes.byref(numel)))
                names.append(name.value.decode())
            if names != [k for k, _ in items]:
                raise CddpmError("encoder engine parameter list differs from the module's state_dict layout")
        if self._versions is None:
            self._versions = [None] * len(items)
        for i, (k, p) in enumerate(items):
            v = (p.data_ptr(), p._version)
            if self._versions[i] != v:
                t = p.detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.float().contiguous()
                check(lib().cddpm_encoder_set_param(self._h, k.encode(), ptr(t), t.numel(), current_stream()),
                      f"cddpm_encoder_set_param({k})")
                self._versions[i] = v

    def forward(self, x, pyramid=0):
        if pyramid != 0:
            raise NotImplementedError("pyramid features are only used by SparK pre-training (out of scope)")
        if not x.is_cuda:
            raise CddpmError("encoder forward needs CUDA tensors (there is no CPU path)")
        if self.training:
            # train() mode means batch statistics, running-statistics updates and DropPath whether or not a tape is
            # being recorded (nn.BatchNorm2d / timm DropPath in the reference, spark/resnet.py:13-46): under
            # torch.no_grad() the same engine forward runs and its autograd node simply never sees a backward
            return self._train_forward(x)
        self._sync()
        x = x.float().contiguous()
        B = x.shape[0]
        if tuple(x.shape[1:]) != (1, *self.image_size):
            raise CddpmError(f"expected [B,1,{self.image_size[0]},{self.image_size[1]}], got {tuple(x.shape)}")
        out = torch.empty(B, self.num_classes, dtype=torch.float32, device=x.device)
        check(lib().cddpm_encoder_forward(self._h, ptr(x), ptr(out), B, current_stream()), "cddpm_encoder_forward")
        return out


    def _train_forward(self, x):
        """Training-mode forward (batch-statistics BatchNorm, running-statistics updates, DropPath) with an autograd
        tape.  Default (`encoder_train_dtype: b200`): the hand-written engine - tcgen05 GEMMs for every convolution's
        forward, data gradient and weight gradient, BatchNorm forward / backward kernels (csrc/resnet_train.cu) - as ONE
        autograd node.  `tf32 | bf16 | fp32` select the library path below (torch autograd over cuDNN), kept as the
        A/B and parity reference."""
        import os

        mode = os.environ.get("CDDPM_ENCODER_TRAIN_DTYPE", getattr(self, "train_dtype", "b200"))
        if mode == "b200":
            return self._train_forward_engine(x)
        return self._train_forward_library(x)

    # ------------------------------------------------------------------ hand-written training engine
    def _train_engine(self):
        h = self.__dict__.get("_th")
        if h is None:
            h = ctypes.c_void_p()
            check(lib().cddpm_encoder_train_create(self.image_size[0], self.image_size[1], self.num_classes,
                                                   ctypes.byref(h)), "cddpm_encoder_train_create")
            n = lib().cddpm_encoder_train_entry_count(h)
            name, numel, isp, off = ctypes.c_char_p(), ctypes.c_int64(), ctypes.c_int(), ctypes.c_int64()
            meta = []
            for i in range(n):
                check(lib().cddpm_encoder_train_entry_info(h, i, ctypes.byref(name), ctypes.byref(numel), ctypes.byref(isp)))
                check(lib().cddpm_encoder_train_grad_offset(h, i, ctypes.byref(off)))
                meta.append((name.value.decode(), int(numel.value), bool(isp.value), int(off.value)))
            if [m[0] for m in meta] != [k for k, _ in self._engine_items()]:
                raise CddpmError("training-encoder entry list differs from the module's state_dict layout")
            self.__dict__["_th"] = h
            self.__dict__["_th_meta"] = meta
            self.__dict__["_th_blocks"] = int(lib().cddpm_encoder_train_num_blocks(h))
            self.__dict__["_th_total"] = int(lib().cddpm_encoder_train_grad_total(h))
        return h

    def _drop_scale(self, B, device):
        rate = float(getattr(self, "drop_path_rate", 0.0) or 0.0)
        if rate <= 0.0:
            return None
        n = self.__dict__["_th_blocks"]
        keep = 1.0 - rate * torch.arange(n, dtype=torch.float32, device=device) / max(1, n - 1)  # linear decay rule
        mask = torch.bernoulli(keep[:, None].expand(n, B).contiguous())
        return (mask / keep[:, None]).contiguous()

    def _train_forward_engine(self, x):
        self._train_engine()
        items = self._engine_items()
        for _, t in items:
            if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                raise CddpmError("the training encoder needs contiguous fp32 CUDA parameters and buffers")
        params = [t for (_, t), m in zip(items, self.__dict__["_th_meta"]) if m[2]]
        out = _EncoderTrainFunction.apply(self, x.detach().float().contiguous(), *params)
        with torch.no_grad():
            nbt = [b for n, b in self.named_buffers() if n.endswith("num_batches_tracked")]
            if nbt:
                torch._foreach_add_(nbt, 1)
        return out

    def _train_forward_library(self, x):
        """The same forward through torch autograd on library kernels (cuDNN convolutions; `encoder_train_dtype: tf32 |
        bf16 | fp32`) - the round-1 path, kept as the parity / A-B reference of the hand-written engine.
        The ~500 small launches of the forward and the backward are replayed as CUDA graphs
        (torch.cuda.make_graphed_callables, one pair per input shape; CDDPM_ENCODER_GRAPH=0 disables it).
        timm's DropPath(0.05) on the residual branches is restated from timm's published source (`drop_path_rate`
        attribute / cfg `encoder_drop_path_rate`; 0 disables it); the draws happen inside the replayed graph (torch's
        graph-safe Philox offsets), so every step sees fresh masks."""
        import os

        if os.environ.get("CDDPM_ENCODER_GRAPH", "1") == "0":
            return _encoder_train_eager(self, x)
        key = (tuple(x.shape), x.device.index, os.environ.get("CDDPM_ENCODER_TRAIN_DTYPE", getattr(self, "train_dtype", "b200")),
               float(getattr(self, "drop_path_rate", 0.0) or 0.0))
        graphs = self.__dict__.setdefault("_train_graphs", {})
        if key not in graphs:
            holder = _EncoderTrainModule(self)
            # capture runs warm-up iterations: keep them out of the BatchNorm running statistics
            saved = {k: v.clone() for k, v in self.state_dict().items() if "running_" in k or "num_batches" in k}
            try:
                graphs[key] = torch.cuda.make_graphed_callables(holder, (x.detach().float().clone(),))
            except Exception:  # capture is an optimisation of the library path, never a requirement
                graphs[key] = None
            with torch.no_grad():
                sd = self.state_dict()
                for k, v in saved.items():
                    sd[k].copy_(v)
        fn = graphs[key]
        if fn is None:
            return _encoder_train_eager(self, x)
        return fn(x.detach().float())


class _EncoderTrainFunction(torch.autograd.Function):
    """ResNet-50 training forward / backward in libcddpm_b200 as one autograd node (cddpm_encoder_train_forward /
    _backward).  The parameters are inputs of the node so that autograd routes their gradients: views of one flat fp32
    buffer in the reference's shapes."""

    @staticmethod
    def forward(ctx, net, x, *params):
        h = net._train_engine()
        items = net._engine_items()
        B = x.shape[0]
        if tuple(x.shape[1:]) != (1, *net.image_size):
            raise CddpmError(f"expected [B,1,{net.image_size[0]},{net.image_size[1]}], got {tuple(x.shape)}")
        table = (ctypes.c_void_p * len(items))(*[t.data_ptr() for _, t in items])
        drop = net._drop_scale(B, x.device)
        out = torch.empty(B, net.num_classes, dtype=torch.float32, device=x.device)
        check(lib().cddpm_encoder_train_forward(h, table, len(items), ptr(x), ptr(drop), ptr(out), B, current_stream()),
              "cddpm_encoder_train_forward")
        ctx.net = net
        ctx.B = B
        ctx.drop = drop  # the backward pass reads the same per-sample scales
        ctx.shapes = [tuple(p.shape) for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        net = ctx.net
        h = net._train_engine()
        flat = torch.empty(net.__dict__["_th_total"], dtype=torch.float32, device=dout.device)
        g = dout.detach().float().contiguous()
        check(lib().cddpm_encoder_train_backward(h, ptr(g), ptr(flat), ctx.B, current_stream()),
              "cddpm_encoder_train_backward")
        grads = []
        k = 0
        for (name, numel, is_param, off) in net.__dict__["_th_meta"]:
            if not is_param:
                continue
            grads.append(flat[off:off + numel].view(ctx.shapes[k]))
            k += 1
        return (None, None, *grads)


class _EncoderTrainModule(nn.Module):
    """nn.Module face of the eager training forward (make_graphed_callables collects parameters from modules)."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x):
        return _encoder_train_eager(self.net, x)


def _encoder_train_eager(self, x):
    import os

    import torch.nn.functional as F

    nhwc = os.environ.get("CDDPM_ENCODER_NHWC", "1") != "0"

    def cl(w):  # channels-last weights keep cuDNN on its NHWC tensor-core kernels (no layout transposes per layer)
        return w.contiguous(memory_format=torch.channels_last) if nhwc else w

    def bn(m, t):
        if m.num_batches_tracked is not None:
            m.num_batches_tracked.add_(1)
        return F.batch_norm(t, m.running_mean, m.running_var, m.weight, m.bias, True, m.momentum, m.eps)

    # Arithmetic of the training-mode encoder: "tf32" (default: fp32 storage, TF32 tensor-core convolutions - the
    # mantissa of the reference's fp16 autocast, and measured as fast as bf16 autocast at B=64: 41.56 vs 41.57 ms per
    # step), "bf16" (autocast) or "fp32" (50.6 ms).  Batch-statistics BatchNorm over a handful of samples (layer4 sees
    # 3x3 pixels per slice) amplifies operand rounding: at B=2 bf16 autocast moves the features by O(1) - on the CPU
    # as much as here - so small-batch parity runs use fp32.
    mode = os.environ.get("CDDPM_ENCODER_TRAIN_DTYPE", getattr(self, "train_dtype", "b200"))
    if mode == "b200":
        mode = "tf32"  # the library path asked for explicitly (CDDPM_ENCODER_GRAPH=0 etc.) under the default setting
    if mode == "bf16":
        ctx = torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False)
    else:
        ctx = torch.backends.cudnn.flags(enabled=True, benchmark=torch.backends.cudnn.benchmark, deterministic=False,
                                         allow_tf32=(mode == "tf32"))
    with ctx, torch.autocast("cuda", enabled=(mode == "bf16"), dtype=torch.bfloat16, cache_enabled=False):
        x = x.float()
        if nhwc:
            x = x.contiguous(memory_format=torch.channels_last)
        h = F.relu(bn(self.bn1, F.conv2d(x, cl(self.conv1.weight), None, 2, 3)))
        h = F.max_pool2d(h, 3, 2, 1)
        rate = float(getattr(self, "drop_path_rate", 0.0) or 0.0)
        n_blocks = sum(len(getattr(self, f"layer{li}")) for li in range(1, 5))
        bi = 0
        for li in range(1, 5):
            for blk in getattr(self, f"layer{li}"):
                idt = h
                o = F.relu(bn(blk.bn1, F.conv2d(h, cl(blk.conv1.weight))))
                o = F.relu(bn(blk.bn2, F.conv2d(o, cl(blk.conv2.weight), None, blk.conv2.stride, 1)))
                o = bn(blk.bn3, F.conv2d(o, cl(blk.conv3.weight)))
                # timm DropPath(block_dpr), block_dpr = rate * block_index / (blocks - 1) (timm/models/resnet.py make_blocks,
                # layers/drop.py drop_path: one Bernoulli(keep) draw per SAMPLE, scaled by 1 / keep), on the residual
                # branch before the shortcut is added.  timm is not installed here: restated, parity unpinned.
                dpr = rate * bi / max(1, n_blocks - 1)
                if dpr > 0.0:
                    keep = 1.0 - dpr
                    mask = torch.empty(o.shape[0], 1, 1, 1, dtype=o.dtype, device=o.device).bernoulli_(keep)
                    o = o * mask.div_(keep)
                bi += 1
                if blk.downsample is not None:
                    idt = bn(blk.downsample[1], F.conv2d(h, cl(blk.downsample[0].weight), None, blk.downsample[0].stride))
                h = F.relu(o + idt)
        h = h.mean((2, 3))
        out = F.linear(h, self.fc.weight, self.fc.bias)
    return out.float()


class SparK_2D_encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.pyramid = cfg.get("pyramid", 4)
        self.cfg = cfg
        size = (int(cfg.imageDim[0] / cfg.rescaleFactor), int(cfg.imageDim[1] / cfg.rescaleFactor))
        dtype = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16}.get(str(cfg.get("engine_dtype", "fp16")), torch.float16)
        self.encoder = ResNet(cfg.version, in_chans=1, num_classes=cfg.get("cond_dim", 128), image_size=size,
                              engine_dtype=dtype)
        # "b200" (default): the hand-written training engine; "tf32" | "bf16" | "fp32": torch autograd over cuDNN
        self.encoder.train_dtype = str(cfg.get("encoder_train_dtype", "b200"))
        self.encoder.drop_path_rate = float(cfg.get("encoder_drop_path_rate", 0.05))  # spark/models.py:50

    def forward(self, x):
        return self.encoder(x)


def get_encoder(cfg):
    """(encoder, out_features) — DDPM_encoder.get_encoder."""
    backbone = cfg.get("backbone", "resnet50")
    if "spark" in backbone.lower():
        encoder = SparK_2D_encoder(cfg)
    else:
        size = (int(cfg.imageDim[0] / cfg.rescaleFactor), int(cfg.imageDim[1] / cfg.rescaleFactor))
        encoder = ResNet(backbone, in_chans=1, num_classes=cfg.get("cond_dim", 256), image_size=size)
    return encoder, cfg.get("cond_dim", 256)
