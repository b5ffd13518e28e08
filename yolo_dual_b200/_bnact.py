"""ctypes binding of libbnact_b200.so (include/bnact_b200.h) and the autograd front of the fused
training-mode BatchNorm2d + SiLU used by the `Conv` block.  No fallback behind the fused call: a missing
library raises; `usable()` only decides whether a given call is one the kernels are written for."""
from __future__ import annotations

import ctypes
import os

import torch
from torch import nn

from .build import BNACT_LIB

SYMBOLS = ("bnact_b200_version", "bnact_b200_last_error", "bnact_b200_supported", "bnact_b200_partial_floats",
           "bnact_b200_forward", "bnact_b200_forward_pitched", "bnact_b200_backward", "bnact_b200_backward_pitched", "bnact_b200_eval", "bnact_b200_eval_pitched")
_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(BNACT_LIB):
        raise ImportError(f"{BNACT_LIB} is missing: build it with `python -m yolo_dual_b200.build`")
    lib = ctypes.CDLL(BNACT_LIB)
    for s in SYMBOLS:
        if not hasattr(lib, s):
            raise ImportError(f"{BNACT_LIB} does not export {s}")
    vp, ip, i64, fl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    lib.bnact_b200_version.restype = ip
    lib.bnact_b200_last_error.restype = ctypes.c_char_p
    lib.bnact_b200_supported.argtypes = [ip, ip]
    lib.bnact_b200_partial_floats.argtypes = [ip, i64, ip]
    lib.bnact_b200_partial_floats.restype = ctypes.c_size_t
    lib.bnact_b200_forward.argtypes = [vp] * 8 + [ip, i64, ip, fl, fl, ip, vp]
    lib.bnact_b200_forward_pitched.argtypes = [vp] * 8 + [ip, i64, ip, fl, fl, ip, i64, vp]
    lib.bnact_b200_backward.argtypes = [vp] * 10 + [ip, i64, ip, ip, vp]
    lib.bnact_b200_eval.argtypes = [vp] * 6 + [ip, ip, i64, ip, fl, ip, vp]
    lib.bnact_b200_backward_pitched.argtypes = [vp] * 10 + [ip, i64, ip, ip, i64, vp]
    lib.bnact_b200_eval_pitched.argtypes = [vp] * 6 + [ip, ip, i64, ip, fl, ip, i64, vp]
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc:
        raise RuntimeError(f"{what}: {'argument' if rc < 0 else 'CUDA'} error {rc}: "
                           f"{load().bnact_b200_last_error().decode('utf-8', 'replace')}")


def enabled() -> bool:
    return os.environ.get("YOLO_DUAL_B200_FUSED_BN", "1") != "0"


def usable(y: torch.Tensor, bn: nn.Module, act: nn.Module) -> bool:
    """True when `act(bn(y))` is a call the fused kernels cover: CUDA, training-mode plain BatchNorm2d with affine
    parameters and a fixed momentum, SiLU or no activation, NHWC-contiguous y with a supported channel count."""
    if not (y.is_cuda and bn.training and enabled()):
        return False
    if type(bn) is not nn.BatchNorm2d or not bn.affine or bn.momentum is None:
        return False
    if not isinstance(act, (nn.SiLU, nn.Identity)):
        return False
    if y.dim() != 4 or y.dtype not in _DTYPES or y.numel() // y.size(1) < 2:
        return False
    if bn.weight.dtype != torch.float32 or not y.is_contiguous(memory_format=torch.channels_last):
        return False
    return bool(load().bnact_b200_supported(_DTYPES[y.dtype], y.size(1)))


class FusedBNAct(torch.autograd.Function):
    """z = act(batch_norm(x; batch statistics)) for NHWC-contiguous x; updates the running statistics in place."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, silu):
        n, c, h, w = x.shape
        m = n * h * w
        dt = _DTYPES[x.dtype]
        lib = load()
        z = torch.empty_like(x)                                   # keeps the NHWC strides
        save = torch.empty(4 * c, dtype=torch.float32, device=x.device)
        partial = torch.empty(lib.bnact_b200_partial_floats(dt, m, c), dtype=torch.float32, device=x.device)
        with torch.cuda.device_of(x):
            _check(lib.bnact_b200_forward(
                x.data_ptr(), z.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                running_mean.data_ptr() if running_mean is not None else None,
                running_var.data_ptr() if running_var is not None else None,
                save.data_ptr(), partial.data_ptr(), dt, m, c, float(eps), float(momentum), int(silu),
                torch.cuda.current_stream().cuda_stream), "bnact_b200_forward")
        ctx.save_for_backward(x, gamma, beta, save)
        ctx.silu = int(silu)
        return z

    @staticmethod
    def backward(ctx, gz):
        x, gamma, beta, save = ctx.saved_tensors
        n, c, h, w = x.shape
        m = n * h * w
        dt = _DTYPES[x.dtype]
        lib = load()
        # The gradient of a channel slice of a wider channels-last tensor (what torch.cat's backward hands to each input:
        # strides (H*W*Ct, 1, W*Ct, Ct)) is read in place at its row pitch; anything else is made NHWC-contiguous.
        if gz.dtype != x.dtype:
            gz = gz.to(x.dtype)
        pitch = slice_pitch(gz)
        if pitch is None:
            gz = gz.contiguous(memory_format=torch.channels_last)
            pitch = c
        dx = torch.empty_like(x)
        small = torch.empty(4 * c, dtype=torch.float32, device=x.device)   # dgamma, dbeta, coef[2]
        partial = torch.empty(lib.bnact_b200_partial_floats(dt, m, c), dtype=torch.float32, device=x.device)
        with torch.cuda.device_of(x):
            _check(lib.bnact_b200_backward_pitched(
                x.data_ptr(), gz.data_ptr(), dx.data_ptr(), gamma.data_ptr(), beta.data_ptr(), save.data_ptr(),
                small.data_ptr(), small[c:].data_ptr(), small[2 * c:].data_ptr(), partial.data_ptr(), dt, m, c,
                ctx.silu, pitch, torch.cuda.current_stream().cuda_stream), "bnact_b200_backward")
        return dx, small[:c], small[c:2 * c], None, None, None, None, None


class FusedBNActInto(torch.autograd.Function):
    """FusedBNAct whose result lands in channels [c0, c0 + C) of `buf`, a dense channels-last tensor that stands for the
    torch.cat of several such results: forward writes the slice in place and returns `buf` (marked dirty), backward reads
    ITS slice of buf's gradient at the row pitch and passes the whole gradient on to the previous writer.  Chaining
    `buf = FusedBNActInto.apply(y1, ..., buf, 0); buf = FusedBNActInto.apply(y2, ..., buf, C1)` builds cat((z1, z2), 1)
    with no copy in either direction."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, silu, buf, c0):
        n, c, h, w = x.shape
        m = n * h * w
        dt = _DTYPES[x.dtype]
        lib = load()
        z = buf[:, c0:c0 + c]
        pitch = slice_pitch(z)
        if pitch is None or z.shape != x.shape or buf.dtype != x.dtype or not buf.is_contiguous(memory_format=torch.channels_last):
            raise ValueError("buf must be a dense channels-last tensor of x's dtype with room for x's channels at c0")
        save = torch.empty(4 * c, dtype=torch.float32, device=x.device)
        partial = torch.empty(lib.bnact_b200_partial_floats(dt, m, c), dtype=torch.float32, device=x.device)
        with torch.cuda.device_of(x):
            _check(lib.bnact_b200_forward_pitched(
                x.data_ptr(), z.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                running_mean.data_ptr() if running_mean is not None else None,
                running_var.data_ptr() if running_var is not None else None,
                save.data_ptr(), partial.data_ptr(), dt, m, c, float(eps), float(momentum), int(silu), pitch,
                torch.cuda.current_stream().cuda_stream), "bnact_b200_forward")
        ctx.save_for_backward(x, gamma, beta, save)
        ctx.silu, ctx.c0 = int(silu), int(c0)
        ctx.mark_dirty(buf)
        return buf

    @staticmethod
    def backward(ctx, gbuf):
        x, gamma, beta, save = ctx.saved_tensors
        n, c, h, w = x.shape
        m = n * h * w
        dt = _DTYPES[x.dtype]
        lib = load()
        if gbuf.dtype != x.dtype:
            gbuf = gbuf.to(x.dtype)
        gz = gbuf[:, ctx.c0:ctx.c0 + c]
        pitch = slice_pitch(gz)
        if pitch is None:
            gz = gz.contiguous(memory_format=torch.channels_last)
            pitch = c
        dx = torch.empty_like(x)
        small = torch.empty(4 * c, dtype=torch.float32, device=x.device)
        partial = torch.empty(lib.bnact_b200_partial_floats(dt, m, c), dtype=torch.float32, device=x.device)
        with torch.cuda.device_of(x):
            _check(lib.bnact_b200_backward_pitched(
                x.data_ptr(), gz.data_ptr(), dx.data_ptr(), gamma.data_ptr(), beta.data_ptr(), save.data_ptr(),
                small.data_ptr(), small[c:].data_ptr(), small[2 * c:].data_ptr(), partial.data_ptr(), dt, m, c,
                ctx.silu, pitch, torch.cuda.current_stream().cuda_stream), "bnact_b200_backward")
        return dx, small[:c], small[c:2 * c], None, None, None, None, None, gbuf, None


def bn_act_into(y: torch.Tensor, bn: nn.BatchNorm2d, act: nn.Module, buf: torch.Tensor, c0: int) -> torch.Tensor:
    """act(bn(y)) written into channels [c0, c0 + C) of `buf`; returns `buf` (callers check `usable` first)."""
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    rm, rv = (bn.running_mean, bn.running_var) if bn.track_running_stats else (None, None)
    return FusedBNActInto.apply(y, bn.weight, bn.bias, rm, rv, bn.eps, bn.momentum, isinstance(act, nn.SiLU), buf, c0)


def bn_act(y: torch.Tensor, bn: nn.BatchNorm2d, act: nn.Module) -> torch.Tensor:
    """act(bn(y)) through the fused kernels (callers check `usable` first)."""
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    rm, rv = (bn.running_mean, bn.running_var) if bn.track_running_stats else (None, None)
    return FusedBNAct.apply(y, bn.weight, bn.bias, rm, rv, bn.eps, bn.momentum, isinstance(act, nn.SiLU))


def usable_eval(y: torch.Tensor, bn: nn.Module, act: nn.Module) -> bool:
    """True when `act(bn(y))` is an inference call the one-pass kernel covers: CUDA, no autograd, eval-mode plain
    BatchNorm2d with running statistics and affine parameters (float32 or y's own 16-bit dtype), SiLU or no
    activation, NHWC-contiguous y with a supported channel count."""
    if not (y.is_cuda and not bn.training and enabled()) or torch.is_grad_enabled():
        return False
    if type(bn) is not nn.BatchNorm2d or not bn.affine or bn.running_mean is None:
        return False
    if not isinstance(act, (nn.SiLU, nn.Identity)):
        return False
    if y.dim() != 4 or y.dtype not in _DTYPES or not y.is_contiguous(memory_format=torch.channels_last):
        return False
    pd = bn.weight.dtype
    if not (pd == bn.bias.dtype == bn.running_mean.dtype == bn.running_var.dtype) or pd not in (torch.float32, y.dtype):
        return False
    return bool(load().bnact_b200_supported(_DTYPES[y.dtype], y.size(1)))


def slice_pitch(t: torch.Tensor):
    """Row pitch (elements) of `t` if it is a channel slice of a channels-last tensor the kernels can address in place
    (strides (H*W*Ct, 1, W*Ct, Ct), 16-byte aligned rows), the channel count for a dense NHWC tensor, else None."""
    n, c, h, w = t.shape
    if t.is_contiguous(memory_format=torch.channels_last):
        return c
    st = t.stride()
    vec = 16 // t.element_size()
    if st[1] == 1 and st[3] >= c and st[2] == w * st[3] and st[0] == h * st[2] and st[3] % vec == 0 and t.data_ptr() % 16 == 0:
        return st[3]
    return None


def bn_act_eval(y: torch.Tensor, bn: nn.BatchNorm2d, act: nn.Module, out: torch.Tensor = None) -> torch.Tensor:
    """act(bn(y)) with running statistics, one pass (callers check `usable_eval` first).  `out`: write the result into
    this tensor — a dense NHWC tensor or a channel slice of a wider one (the destination of a would-be torch.cat)."""
    n, c, h, w = y.shape
    z = torch.empty_like(y) if out is None else out
    pitch = slice_pitch(z)
    if pitch is None or z.shape != y.shape or z.dtype != y.dtype:
        raise ValueError("out must be a dense NHWC tensor or a channel slice of one, same shape and dtype as the input")
    with torch.cuda.device_of(y):
        _check(load().bnact_b200_eval_pitched(y.data_ptr(), z.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr(),
                                              bn.running_mean.data_ptr(), bn.running_var.data_ptr(), _DTYPES[y.dtype],
                                              int(bn.weight.dtype != torch.float32), n * h * w, c, float(bn.eps),
                                              int(isinstance(act, nn.SiLU)), pitch, torch.cuda.current_stream().cuda_stream),
               "bnact_b200_eval")
    return z
