"""Handle to the C++ UNet engine (cddpm_unet_* in include/cddpm_b200.h)."""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import CddpmError, UNetConfig, check, current_stream, lib, ptr


class UNetEngine:
    """Owns one cddpm_unet_t.  Parameters are pushed by reference state_dict key; forward runs the planned kernel
    sequence on the current CUDA stream.  No CPU path: everything here requires CUDA tensors."""

    def __init__(self, *, image_size: Tuple[int, int], in_channels: int, model_channels: int, out_channels: int,
                 num_res_blocks: int, attention_resolutions: Sequence[int], channel_mult: Sequence[int],
                 num_classes: Optional[int], num_head_channels: int = 64, dtype=torch.float16, training: bool = False):
        cfg = UNetConfig()
        cfg.image_h, cfg.image_w = int(image_size[0]), int(image_size[1])
        cfg.in_channels, cfg.model_channels, cfg.out_channels = in_channels, model_channels, out_channels
        cfg.num_res_blocks = num_res_blocks
        cfg.n_mult = len(channel_mult)
        for i, m in enumerate(channel_mult):
            cfg.channel_mult[i] = int(m)
        cfg.n_attn_res = len(attention_resolutions)
        for i, a in enumerate(attention_resolutions):
            cfg.attention_resolutions[i] = int(a)
        cfg.num_classes = int(num_classes) if num_classes else 0
        cfg.num_head_channels = num_head_channels
        cfg.fmt = _lib.fmt_of(dtype)
        self.dtype = dtype
        self.image_size = (cfg.image_h, cfg.image_w)
        self.num_classes = cfg.num_classes
        self._h = ctypes.c_void_p()
        check(lib().cddpm_unet_create(ctypes.byref(cfg), ctypes.byref(self._h)), "cddpm_unet_create")
        # training engines keep the intermediates backward() reads; inference engines fuse them away (cddpm_b200.h)
        self.training = bool(training)
        if self.training:
            check(lib().cddpm_unet_set_training(self._h, 1), "cddpm_unet_set_training")

    def set_training(self, on: bool) -> None:
        """Switch between a training engine (keeps the intermediates backward() reads) and an inference engine (may
        fuse them away).  Switching drops the current plan; a no-op when the mode does not change."""
        if bool(on) != self.training:
            check(lib().cddpm_unet_set_training(self._h, 1 if on else 0), "cddpm_unet_set_training")
            self.training = bool(on)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                lib().cddpm_unet_destroy(h)
            except Exception:
                pass
            self._h = ctypes.c_void_p()

    # ------------------------------------------------------------------ parameters
    def param_names(self) -> List[Tuple[str, int]]:
        n = lib().cddpm_unet_param_count(self._h)
        out = []
        name = ctypes.c_char_p()
        numel = ctypes.c_int64()
        for i in range(n):
            check(lib().cddpm_unet_param_info(self._h, i, ctypes.byref(name), ctypes.byref(numel)))
            out.append((name.value.decode(), int(numel.value)))
        return out

    def set_param(self, name: str, value: torch.Tensor) -> None:
        if not value.is_cuda:
            raise CddpmError(f"parameter {name} is on {value.device}; the cDDPM engine needs CUDA tensors")
        v = value.detach()
        if v.dtype != torch.float32 or not v.is_contiguous():
            v = v.float().contiguous()
        check(lib().cddpm_unet_set_param(self._h, name.encode(), ptr(v), v.numel(), current_stream()),
              f"cddpm_unet_set_param({name})")

    def set_params(self, values: Sequence[Optional[torch.Tensor]]) -> None:
        """Bulk push in param_names() order (None = unchanged): one C call instead of one per tensor."""
        keep = []
        ptrs = []
        for v in values:
            if v is None:
                ptrs.append(None)
                continue
            t = v.detach()
            if not t.is_cuda:
                raise CddpmError("the cDDPM engine needs CUDA tensors")
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
                keep.append(t)
            ptrs.append(t.data_ptr())
        check(lib().cddpm_unet_set_params(self._h, _lib.ptr_array(ptrs), len(ptrs), current_stream()),
              "cddpm_unet_set_params")

    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "") -> None:
        for name, _ in self.param_names():
            self.set_param(name, sd[prefix + name])

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, t: torch.Tensor, cond: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not x.is_cuda:
            raise CddpmError("UNet forward needs CUDA tensors (there is no CPU path)")
        B = x.shape[0]
        if tuple(x.shape[1:]) != (1, self.image_size[0], self.image_size[1]):
            raise CddpmError(f"expected input [B,1,{self.image_size[0]},{self.image_size[1]}], got {tuple(x.shape)}")
        x = x.float().contiguous()
        t = t.to(device=x.device, dtype=torch.int64).contiguous()
        if t.numel() != B:
            raise CddpmError("timesteps must have one entry per batch element")
        if self.num_classes:
            if cond is None:
                raise CddpmError("conditioned UNet called without cond")
            cond = cond.float().contiguous()
            if tuple(cond.shape) != (B, self.num_classes):
                raise CddpmError(f"cond must be [B,{self.num_classes}], got {tuple(cond.shape)}")
        else:
            cond = None
        if out is None:
            out = torch.empty_like(x)
        check(lib().cddpm_unet_forward(self._h, ptr(x), ptr(t), ptr(cond), ptr(out), B, current_stream()),
              "cddpm_unet_forward")
        self.forward_serial = getattr(self, "forward_serial", 0) + 1
        return out

    # ------------------------------------------------------------------ backward (training step)
    def grad_layout(self) -> Tuple[int, List[int]]:
        """(floats in the flat gradient buffer, offset of every parameter in param_names() order)."""
        if getattr(self, "_grad_layout", None) is None:
            total = int(lib().cddpm_unet_grad_total(self._h))
            offs = []
            off = ctypes.c_int64()
            for i in range(lib().cddpm_unet_param_count(self._h)):
                check(lib().cddpm_unet_grad_offset(self._h, i, ctypes.byref(off)), "cddpm_unet_grad_offset")
                offs.append(int(off.value))
            self._grad_layout = (total, offs)
        return self._grad_layout

    def backward(self, dout: torch.Tensor, want_dcond: bool = False):
        """Backward of the LAST forward: dout = dL/d out [B,1,H,W].  Returns (flat fp32 gradient buffer, dcond)."""
        if not dout.is_cuda:
            raise CddpmError("UNet backward needs CUDA tensors (there is no CPU path)")
        if self.dtype != torch.bfloat16:
            raise CddpmError("the training step runs on a bf16 engine")
        B = dout.shape[0]
        dout = dout.float().contiguous()
        total, _ = self.grad_layout()
        grads = torch.empty(total, dtype=torch.float32, device=dout.device)
        dcond = None
        if want_dcond and self.num_classes:
            dcond = torch.empty(B, self.num_classes, dtype=torch.float32, device=dout.device)
        check(lib().cddpm_unet_backward(self._h, ptr(dout), ptr(grads), ptr(dcond), B, current_stream()),
              "cddpm_unet_backward")
        return grads, dcond

    @property
    def bwd_flops_per_sample(self) -> int:
        return int(lib().cddpm_unet_bwd_flops(self._h))

    # ------------------------------------------------------------------ introspection (parity tests, bench)
    def tap(self, layer: str, batch: int) -> torch.Tensor:
        """NCHW fp32 copy of a layer output of the last forward."""
        p = ctypes.c_void_p()
        C, H, W = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(lib().cddpm_unet_tap(self._h, layer.encode(), ctypes.byref(p), ctypes.byref(C), ctypes.byref(H),
                                   ctypes.byref(W)), f"cddpm_unet_tap({layer})")
        n = batch * C.value * H.value * W.value
        buf = torch.empty(n, dtype=self.dtype, device="cuda")
        torch.cuda.current_stream().synchronize()
        _cudart_memcpy(buf.data_ptr(), p.value, n * 2)
        return buf.view(batch, H.value, W.value, C.value).permute(0, 3, 1, 2).float()

    def film(self, batch: int) -> torch.Tensor:
        p = ctypes.c_void_p()
        stride = ctypes.c_int()
        check(lib().cddpm_unet_film(self._h, ctypes.byref(p), ctypes.byref(stride)))
        buf = torch.empty(batch, stride.value, dtype=torch.float32, device="cuda")
        torch.cuda.current_stream().synchronize()
        _cudart_memcpy(buf.data_ptr(), p.value, batch * stride.value * 4)
        return buf

    def profile_arm(self) -> None:
        """Time every convolution launch of the next forward with CUDA events (bench roofline)."""
        check(lib().cddpm_unet_profile_arm(self._h), "cddpm_unet_profile_arm")

    def profile_read(self) -> Tuple[float, int]:
        ms, n = ctypes.c_double(), ctypes.c_int()
        check(lib().cddpm_unet_profile_read(self._h, ctypes.byref(ms), ctypes.byref(n)), "cddpm_unet_profile_read")
        return float(ms.value), int(n.value)

    @property
    def conv_flops_per_sample(self) -> int:
        return int(lib().cddpm_unet_conv_flops(self._h))

    @property
    def launches_per_forward(self) -> int:
        return int(lib().cddpm_unet_launches(self._h))


def _cudart_memcpy(dst: int, src: int, nbytes: int) -> None:
    """Device-to-device copy on the current stream (used only by the introspection helpers)."""
    check(lib().cddpm_memcpy_d2d(dst, src, nbytes, current_stream()), "cddpm_memcpy_d2d")
    torch.cuda.current_stream().synchronize()
