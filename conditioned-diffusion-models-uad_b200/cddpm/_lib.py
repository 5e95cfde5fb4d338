"""ctypes binding of libcddpm_b200.so (the C ABI in include/cddpm_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, we raise.  torch is used only
to obtain device pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libcddpm_b200.so")

FMT_F16 = 0
FMT_BF16 = 1


class CddpmError(RuntimeError):
    """Raised when a libcddpm_b200 entry point returns a non-zero status."""


class UNetConfig(ctypes.Structure):
    """struct cddpm_unet_config (include/cddpm_b200.h)."""

    _fields_ = [
        ("image_h", ctypes.c_int), ("image_w", ctypes.c_int),
        ("in_channels", ctypes.c_int), ("model_channels", ctypes.c_int), ("out_channels", ctypes.c_int),
        ("num_res_blocks", ctypes.c_int),
        ("n_mult", ctypes.c_int), ("channel_mult", ctypes.c_int * 8),
        ("n_attn_res", ctypes.c_int), ("attention_resolutions", ctypes.c_int * 8),
        ("num_classes", ctypes.c_int), ("num_head_channels", ctypes.c_int), ("fmt", ctypes.c_int),
    ]


class VolView(ctypes.Structure):
    """struct cddpm_vol_view: a float32 volume addressed as (y, x, d) with element strides."""

    _fields_ = [("ptr", ctypes.c_void_p), ("sy", ctypes.c_int64), ("sx", ctypes.c_int64), ("sd", ctypes.c_int64)]


_lib: Optional[ctypes.CDLL] = None


def _signatures(c):
    """(restype, argtypes) of every symbol include/cddpm_b200.h declares, in header order."""
    vp, i32, i64, f32 = c.c_void_p, c.c_int, c.c_int64, c.c_float
    pvp, pi32 = c.POINTER(vp), c.POINTER(i32)
    pview = c.POINTER(VolView)
    return {
        "cddpm_last_error": (c.c_char_p, []),
        "cddpm_version": (c.c_char_p, []),
        "cddpm_memcpy_d2d": (i32, [vp, vp, i64, vp]),
        "cddpm_pack_conv_weight": (i32, [vp, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
        "cddpm_conv_igemm": (i32, [i32, pvp, pi32, pi32, i32, i32, i32, i32, vp, vp, vp, vp, i32, i32, vp]),
        "cddpm_conv_igemm_stats": (i32, [i32, pvp, pi32, pi32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp]),
        "cddpm_pack_conv_weight_t": (i32, [vp, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
        "cddpm_conv_wgrad": (i32, [i32, pvp, pi32, pi32, pi32, vp, i32, i32, i32, i32, vp, i32, vp]),
        "cddpm_unpack_conv_grad": (i32, [vp, i32, i32, i32, i32, i32, vp, i32, i32, vp]),
        "cddpm_gn_workspace_floats": (i64, [i32, i32]),
        "cddpm_groupnorm_film_silu": (
            i32, [vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, vp]),
        "cddpm_linear": (i32, [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
        "cddpm_timestep_embedding": (i32, [vp, vp, i32, i32, vp]),
        "cddpm_attention": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "cddpm_unet_create": (i32, [c.POINTER(UNetConfig), pvp]),
        "cddpm_unet_destroy": (None, [vp]),
        "cddpm_unet_param_count": (i32, [vp]),
        "cddpm_unet_param_info": (i32, [vp, i32, c.POINTER(c.c_char_p), c.POINTER(i64)]),
        "cddpm_unet_set_param": (i32, [vp, c.c_char_p, vp, i64, vp]),
        "cddpm_unet_forward": (i32, [vp, vp, vp, vp, vp, i32, vp]),
        "cddpm_unet_tap": (i32, [vp, c.c_char_p, pvp, pi32, pi32, pi32]),
        "cddpm_unet_film": (i32, [vp, pvp, pi32]),
        "cddpm_unet_conv_flops": (i64, [vp]),
        "cddpm_unet_launches": (i32, [vp]),
        "cddpm_unet_profile_arm": (i32, [vp]),
        "cddpm_unet_profile_read": (i32, [vp, c.POINTER(c.c_double), pi32]),
        "cddpm_unet_set_params": (i32, [vp, pvp, i32, vp]),
        "cddpm_unet_grad_total": (i64, [vp]),
        "cddpm_unet_grad_offset": (i32, [vp, i32, c.POINTER(i64)]),
        "cddpm_unet_backward": (i32, [vp, vp, vp, vp, i32, vp]),
        "cddpm_unet_set_training": (i32, [vp, i32]),
        "cddpm_unet_bwd_flops": (i64, [vp]),
        "cddpm_unet_bwd_launches": (i32, [vp]),
        "cddpm_adam_step": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, f32, f32, f32, f32, f32, f32, vp]),
        "cddpm_attention_bwd_scratch_bytes": (i64, [i32, i32, i32]),
        "cddpm_attention_bwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
        "cddpm_encoder_create": (i32, [i32, i32, i32, i32, pvp]),
        "cddpm_encoder_destroy": (None, [vp]),
        "cddpm_encoder_param_count": (i32, [vp]),
        "cddpm_encoder_param_info": (i32, [vp, i32, c.POINTER(c.c_char_p), c.POINTER(i64)]),
        "cddpm_encoder_set_param": (i32, [vp, c.c_char_p, vp, i64, vp]),
        "cddpm_encoder_forward": (i32, [vp, vp, vp, i32, vp]),
        "cddpm_encoder_train_create": (i32, [i32, i32, i32, pvp]),
        "cddpm_encoder_train_destroy": (None, [vp]),
        "cddpm_encoder_train_entry_count": (i32, [vp]),
        "cddpm_encoder_train_entry_info": (i32, [vp, i32, c.POINTER(c.c_char_p), c.POINTER(i64), pi32]),
        "cddpm_encoder_train_grad_total": (i64, [vp]),
        "cddpm_encoder_train_grad_offset": (i32, [vp, i32, c.POINTER(i64)]),
        "cddpm_encoder_train_num_blocks": (i32, [vp]),
        "cddpm_encoder_train_launches": (i32, [vp, i32]),
        "cddpm_encoder_train_forward": (i32, [vp, pvp, i32, vp, vp, vp, i32, vp]),
        "cddpm_encoder_train_backward": (i32, [vp, vp, vp, i32, vp]),
        "cddpm_flat_wgrad": (i32, [vp, vp, i32, i32, i32, vp, vp]),
        "cddpm_simplex_noise": (i32, [c.c_char_p, vp, vp, i32, i32, i32, i32, c.c_double, c.c_double, vp]),
        "cddpm_q_sample": (i32, [vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
        "cddpm_posterior_step": (i32, [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp]),
        "cddpm_ddim_step": (i32, [vp, vp, vp, i32, vp, f32, f32, f32, f32, f32, i32, i32, i32, i32, i32, vp]),
        "cddpm_loss_backward": (i32, [vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
        "cddpm_recon_finish": (
            i32, [vp, vp, vp, vp, i32, vp, f32, f32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
        "cddpm_residual_erode": (i32, [pview, pview, pview, pview, i32, i32, i32, i32, i32, vp, vp, vp]),
        "cddpm_trilinear_resize": (i32, [pview, i32, i32, i32, vp, i32, i32, i32, vp]),
        "cddpm_median3d": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "cddpm_compose_grid": (i32, [vp, vp, i32, i32, vp, vp]),
        "cddpm_max": (i32, [vp, i64, vp, vp]),
        "cddpm_threshold_counts": (i32, [vp, pview, i32, i32, i32, c.POINTER(f32), i32, vp, vp]),
        "cddpm_threshold_mask": (i32, [vp, i64, f32, vp, vp]),
        "cddpm_row_stats": (i32, [vp, pview, pview, i32, i32, i32, f32, vp, vp, vp]),
        "cddpm_ranking_workspace_bytes": (i64, [i64]),
        "cddpm_ranking_metrics": (i32, [vp, pview, i32, i32, i32, vp, i64, vp, vp]),
        "cddpm_dice_bisect": (i32, [vp, i64, i32, vp, vp]),
        "cddpm_crop_or_pad": (i32, [vp, i32, i32, i32, vp, i32, i32, i32, f32, vp]),
        "cddpm_rescale_workspace_bytes": (i64, [i64]),
        "cddpm_rescale_intensity": (i32, [vp, vp, i64, c.c_double, c.c_double, f32, f32, vp, i64, vp, vp]),
        "cddpm_resample_size": (i32, [i32, c.c_double]),
        "cddpm_resample_workspace_bytes": (i64, [i32, i32, i32]),
        "cddpm_resample": (i32, [vp, i32, i32, i32, c.c_double, c.c_double, c.c_double, i32, vp, vp, i64, vp]),
        "cddpm_filter_small_components": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "cddpm_confusion_counts": (i32, [vp, pview, i32, i32, i32, vp, vp]),
        "cddpm_hausdorff_workspace_bytes": (i64, [i32, i32, i32]),
        "cddpm_hausdorff": (i32, [vp, pview, i32, i32, i32, vp, i64, vp, vp]),
    }


def _declare(lib: ctypes.CDLL) -> None:
    for name, (res, args) in _signatures(ctypes).items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync: fail loudly
        fn.restype = res
        fn.argtypes = args


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CddpmError(
                f"{LIB_PATH} not found: build it with `python conditioned-diffusion-models-uad_b200/build.py` "
                "(there is no CPU or PyTorch fallback for this path)"
            )
        handle = ctypes.CDLL(LIB_PATH)
        _declare(handle)
        _lib = handle
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().cddpm_last_error()
        raise CddpmError(f"{what or 'libcddpm_b200'} failed ({status}): {msg.decode() if msg else ''}")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL).  The tensor must be contiguous and on CUDA."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CddpmError("libcddpm_b200 operates on CUDA tensors only (no CPU path)")
    if not t.is_contiguous():
        raise CddpmError("libcddpm_b200 requires contiguous tensors")
    return t.data_ptr()


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def fmt_of(dtype) -> int:
    import torch

    if dtype == torch.bfloat16:
        return FMT_BF16
    if dtype == torch.float16:
        return FMT_F16
    raise CddpmError(f"unsupported activation dtype {dtype}")


def int_array(values: Sequence[int]):
    return (ctypes.c_int * len(values))(*values)


def ptr_array(values: Sequence[Optional[int]]):
    return (ctypes.c_void_p * len(values))(*values)
