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


_lib: Optional[ctypes.CDLL] = None


def _signatures(c):
    """(restype, argtypes) of every symbol include/cddpm_b200.h declares, in header order."""
    vp, i32, i64, f32 = c.c_void_p, c.c_int, c.c_int64, c.c_float
    pvp, pi32 = c.POINTER(vp), c.POINTER(i32)
    return {
        "cddpm_last_error": (c.c_char_p, []),
        "cddpm_version": (c.c_char_p, []),
        "cddpm_pack_conv_weight": (i32, [vp, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
        "cddpm_conv_igemm": (i32, [i32, pvp, pi32, pi32, i32, i32, i32, i32, vp, vp, vp, vp, i32, i32, vp]),
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
