"""ctypes binding of libdcnv3_b200.so (include/dcnv3_b200.h).

Fails loudly: if the library is missing or a symbol is absent the import of the
op raises — there is no Python, PyTorch or CPU fallback behind it.
"""
from __future__ import annotations

import ctypes
import os

from .build import LIB

F32, F16, BF16, F64 = 0, 1, 2, 3
ACC_OPMATH, ACC_STORAGE, ACC_TILE = 0, 1, 2
ENOTSUP = -7

SYMBOLS = (
    "dcnv3_b200_version",
    "dcnv3_b200_reload_knobs",
    "dcnv3_b200_last_error",
    "dcnv3_b200_output_size",
    "dcnv3_b200_forward",
    "dcnv3_b200_backward_workspace_bytes",
    "dcnv3_b200_backward",
    "dcnv3_b200_debug_indices",
    "dcnv3_b200_forward_packed",
    "dcnv3_b200_backward_packed",
)


class Geometry(ctypes.Structure):
    """struct dcnv3_b200_geometry"""
    _fields_ = [(n, ctypes.c_int) for n in (
        "N", "H", "W", "kernel_h", "kernel_w", "stride_h", "stride_w", "pad_h", "pad_w",
        "dilation_h", "dilation_w", "group", "group_channels")] + [("offset_scale", ctypes.c_float)]


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise ImportError(
            f"{LIB} is missing: build it with `python -m yolo_dual_b200.build` "
            "(or __graft_entry__.build()).  dcnv3_b200 has no fallback path.")
    lib = ctypes.CDLL(LIB)
    for s in SYMBOLS:
        if not hasattr(lib, s):
            raise ImportError(f"{LIB} does not export {s}")
    vp, ip, gp = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(Geometry)
    lib.dcnv3_b200_version.restype = ip
    lib.dcnv3_b200_last_error.restype = ctypes.c_char_p
    lib.dcnv3_b200_reload_knobs.restype = None
    lib.dcnv3_b200_reload_knobs.argtypes = []
    lib.dcnv3_b200_output_size.argtypes = [gp, ctypes.POINTER(ip), ctypes.POINTER(ip)]
    lib.dcnv3_b200_forward.argtypes = [vp, vp, vp, vp, ip, gp, ip, vp]
    lib.dcnv3_b200_backward_workspace_bytes.argtypes = [ip, gp, ip]
    lib.dcnv3_b200_backward_workspace_bytes.restype = ctypes.c_size_t
    lib.dcnv3_b200_backward.argtypes = [vp] * 8 + [ctypes.c_size_t, ip, gp, ip, ip, vp]
    lib.dcnv3_b200_debug_indices.argtypes = [vp, vp, vp, ip, gp, vp]
    lib.dcnv3_b200_forward_packed.argtypes = [vp, vp, vp, ip, gp, ip, vp]
    lib.dcnv3_b200_backward_packed.argtypes = [vp, vp, vp, vp, vp, ip, gp, ip, vp]
    _lib = lib
    return lib


def reload_knobs() -> None:
    """Re-read the DCNV3_B200_* environment knobs (the library caches them on first use).  Tests only."""
    load().dcnv3_b200_reload_knobs()


def last_error() -> str:
    return load().dcnv3_b200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = last_error()
    kind = "argument error" if rc < 0 else "CUDA error"
    raise RuntimeError(f"{what}: {kind} {rc}: {msg}")
