"""Python wrappers over the raw kernel entry points of libcddpm_b200 (one function per C symbol).

These exist for the parity tests and for host code that needs a single kernel; the UNet forward itself runs inside
the C++ engine (cddpm/engine.py -> cddpm_unet_*).  All tensors are CUDA tensors; activations are NHWC 16-bit.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, current_stream, fmt_of, int_array, lib, ptr, ptr_array


def pack_conv_weight(weight: torch.Tensor, splits: Sequence[int], dtype=torch.bfloat16) -> torch.Tensor:
    """OIHW fp32 weight -> packed [Cout, Ktot] 16-bit matrix; `splits` = input-channel count of each A source."""
    assert weight.is_cuda and weight.dtype == torch.float32 and weight.dim() == 4
    cout, cin, kh, kw = weight.shape
    assert kh == kw and sum(splits) == cin
    ktot = cin * kh * kw
    out = torch.empty(cout, ktot, device=weight.device, dtype=dtype)
    w = weight.contiguous()
    cin_off = koff = 0
    for c_s in splits:
        check(
            lib().cddpm_pack_conv_weight(ptr(w), cout, cin, kh, cin_off, c_s, ptr(out), ktot, koff, fmt_of(dtype),
                                         current_stream()),
            "cddpm_pack_conv_weight",
        )
        cin_off += c_s
        koff += c_s * kh * kw
    return out


def conv_igemm(
    srcs: Sequence[torch.Tensor],
    taps: Sequence[int],
    wpacked: torch.Tensor,
    bias: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    out_f32: bool = False,
) -> torch.Tensor:
    """sum_s conv(srcs[s]) + bias + residual; srcs are NHWC 16-bit, wpacked is [Cout, sum_s taps_s*C_s]."""
    B, H, W, _ = srcs[0].shape
    cout = wpacked.shape[0]
    dtype = srcs[0].dtype
    out = torch.empty(B, H, W, cout, device=srcs[0].device, dtype=torch.float32 if out_f32 else dtype)
    check(
        lib().cddpm_conv_igemm(
            len(srcs),
            ptr_array([ptr(s) for s in srcs]),
            int_array([s.shape[3] for s in srcs]),
            int_array(list(taps)),
            B, H, W, cout,
            ptr(wpacked), ptr(bias), ptr(residual), ptr(out),
            1 if out_f32 else 0, fmt_of(dtype), current_stream(),
        ),
        "cddpm_conv_igemm",
    )
    return out


def conv_igemm_stats(
    srcs: Sequence[torch.Tensor],
    taps: Sequence[int],
    wpacked: torch.Tensor,
    bias: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
):
    """conv_igemm with a 16-bit output plus the GroupNorm partial sums of the fp32 result: returns (out, stats) with
    stats [B, cout/4, 2] float64 = (sum, sum of squares) per image and 4-channel bucket (util.py:214-216 GroupNorm32
    over the output of a ResBlock convolution, OpenAI_Unet.py:284-338)."""
    B, H, W, _ = srcs[0].shape
    cout = wpacked.shape[0]
    dtype = srcs[0].dtype
    out = torch.empty(B, H, W, cout, device=srcs[0].device, dtype=dtype)
    stats = torch.zeros(B, cout // 4, 2, device=srcs[0].device, dtype=torch.float64)
    check(
        lib().cddpm_conv_igemm_stats(
            len(srcs),
            ptr_array([ptr(s) for s in srcs]),
            int_array([s.shape[3] for s in srcs]),
            int_array(list(taps)),
            B, H, W, cout,
            ptr(wpacked), ptr(bias), ptr(residual), ptr(out),
            fmt_of(dtype), ptr(stats), current_stream(),
        ),
        "cddpm_conv_igemm_stats",
    )
    return out, stats


def attention(qkv: torch.Tensor, channels: int) -> torch.Tensor:
    """QKVAttention over qkv [B, L, 3*C] (q | k | v, heads of 64 channels) -> [B, L, C] (OpenAI_Unet.py:457-476)."""
    B, L, c3 = qkv.shape
    assert c3 == 3 * channels
    out = torch.empty(B, L, channels, device=qkv.device, dtype=qkv.dtype)
    check(lib().cddpm_attention(ptr(qkv), ptr(out), B, L, channels, fmt_of(qkv.dtype), current_stream()),
          "cddpm_attention")
    return out


def pack_conv_weight_t(weight: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """OIHW fp32 weight -> panel [Cin, k*k*Cout] of the data-gradient convolution (flipped taps, channels exchanged)."""
    cout, cin, kh, kw = weight.shape
    out = torch.empty(cin, kh * kw * cout, device=weight.device, dtype=dtype)
    wc = weight.contiguous()
    check(lib().cddpm_pack_conv_weight_t(ptr(wc), cout, cin, kh, 0, cin, ptr(out), kh * kw * cout, 0,
                                         fmt_of(dtype), current_stream()), "cddpm_pack_conv_weight_t")
    return out


def conv_wgrad(srcs: Sequence[torch.Tensor], taps: Sequence[int], dy: torch.Tensor,
               skip: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Packed fp32 weight gradient [Cout, sum_s taps_s*C_s] of out = sum_s conv(srcs[s]) given dy (NHWC 16-bit)."""
    B, H, W, cout = dy.shape
    ktot = sum(t * s.shape[3] for t, s in zip(taps, srcs))
    dw = torch.zeros(cout, ktot, device=dy.device, dtype=torch.float32)
    check(lib().cddpm_conv_wgrad(len(srcs), ptr_array([ptr(s) for s in srcs]), int_array([s.shape[3] for s in srcs]),
                                 int_array(list(taps)), int_array(list(skip) if skip else [0] * len(srcs)), ptr(dy),
                                 B, H, W, cout, ptr(dw), fmt_of(dy.dtype), current_stream()), "cddpm_conv_wgrad")
    return dw


def unpack_conv_grad(dw: torch.Tensor, cout: int, cin: int, ksize: int, splits: Sequence[int]) -> torch.Tensor:
    """Packed gradient [Cout, Ktot] -> OIHW fp32 (inverse of pack_conv_weight)."""
    g = torch.zeros(cout, cin, ksize, ksize, device=dw.device, dtype=torch.float32)
    cin_off = koff = 0
    for c_s in splits:
        check(lib().cddpm_unpack_conv_grad(ptr(dw), cout, cin, ksize, cin_off, c_s, ptr(g), dw.shape[1], koff,
                                           current_stream()), "cddpm_unpack_conv_grad")
        cin_off += c_s
        koff += c_s * ksize * ksize
    return g
