"""ctypes binding of libresize_b200.so (include/resize_b200.h) and the autograd front of the NHWC resize.
No fallback behind the call: a missing library raises; `usable()` only says whether the kernels cover a call."""
from __future__ import annotations

import ctypes
import os

import torch

from .build import RESIZE_LIB

SYMBOLS = ("resize_b200_version", "resize_b200_last_error", "resize_b200_supported", "resize_b200_forward",
           "resize_b200_backward")
_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
MODES = {"nearest": 0, "bilinear": 1}
_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(RESIZE_LIB):
        raise ImportError(f"{RESIZE_LIB} is missing: build it with `python -m yolo_dual_b200.build`")
    lib = ctypes.CDLL(RESIZE_LIB)
    for s in SYMBOLS:
        if not hasattr(lib, s):
            raise ImportError(f"{RESIZE_LIB} does not export {s}")
    vp, ip = ctypes.c_void_p, ctypes.c_int
    lib.resize_b200_version.restype = ip
    lib.resize_b200_last_error.restype = ctypes.c_char_p
    lib.resize_b200_supported.argtypes = [ip] * 7
    lib.resize_b200_forward.argtypes = [vp, vp] + [ip] * 8 + [vp]
    lib.resize_b200_backward.argtypes = [vp, vp] + [ip] * 8 + [vp]
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc:
        raise RuntimeError(f"{what}: {'argument' if rc < 0 else 'CUDA'} error {rc}: "
                           f"{load().resize_b200_last_error().decode('utf-8', 'replace')}")


def usable(x: torch.Tensor, size, mode: str) -> bool:
    """CUDA, NHWC-contiguous 4-D tensor of a supported dtype / channel count; nearest only by integer factors."""
    if os.environ.get("YOLO_DUAL_B200_RESIZE", "1") == "0":
        return False
    if not x.is_cuda or x.dim() != 4 or x.dtype not in _DTYPES or mode not in MODES:
        return False
    if not x.is_contiguous(memory_format=torch.channels_last):
        return False
    n, c, h, w = x.shape
    return bool(load().resize_b200_supported(_DTYPES[x.dtype], c, h, w, int(size[0]), int(size[1]), MODES[mode]))


class ResizeNHWC(torch.autograd.Function):
    """F.interpolate(x, size=(ho, wo), mode=...) ('nearest' by integer factors, 'bilinear' align_corners=False)."""

    @staticmethod
    def forward(ctx, x, ho, wo, mode):
        n, c, h, w = x.shape
        y = torch.empty((n, c, ho, wo), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        with torch.cuda.device_of(x):
            _check(load().resize_b200_forward(x.data_ptr(), y.data_ptr(), _DTYPES[x.dtype], n, h, w, c, ho, wo,
                                              MODES[mode], torch.cuda.current_stream().cuda_stream),
                   "resize_b200_forward")
        ctx.geom = (n, c, h, w, ho, wo, mode)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, c, h, w, ho, wo, mode = ctx.geom
        gy = gy.contiguous(memory_format=torch.channels_last)
        gx = torch.empty((n, c, h, w), dtype=gy.dtype, device=gy.device, memory_format=torch.channels_last)
        with torch.cuda.device_of(gy):
            _check(load().resize_b200_backward(gy.data_ptr(), gx.data_ptr(), _DTYPES[gy.dtype], n, h, w, c, ho, wo,
                                               MODES[mode], torch.cuda.current_stream().cuda_stream),
                   "resize_b200_backward")
        return gx, None, None, None


def resize(x: torch.Tensor, size, mode: str) -> torch.Tensor:
    return ResizeNHWC.apply(x, int(size[0]), int(size[1]), mode)
