"""ctypes binding of libsegloss_b200.so (include/segloss_b200.h) and the autograd front of the fused loss.

No fallback behind it: a missing library or symbol raises on first use."""
from __future__ import annotations

import ctypes
import os

import torch

from .build import SEGLOSS_LIB

SYMBOLS = ("segloss_b200_version", "segloss_b200_last_error", "segloss_b200_forward", "segloss_b200_backward")
MAX_CLASSES = 16
_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SEGLOSS_LIB):
        raise ImportError(f"{SEGLOSS_LIB} is missing: build it with `python -m yolo_dual_b200.build`")
    lib = ctypes.CDLL(SEGLOSS_LIB)
    for s in SYMBOLS:
        if not hasattr(lib, s):
            raise ImportError(f"{SEGLOSS_LIB} does not export {s}")
    vp, ip = ctypes.c_void_p, ctypes.c_int
    lib.segloss_b200_version.restype = ip
    lib.segloss_b200_last_error.restype = ctypes.c_char_p
    lib.segloss_b200_forward.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip, ip, vp]
    lib.segloss_b200_backward.argtypes = [vp, vp, vp, vp, vp, ip, ip, ip, ip, ip, vp]
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc:
        raise RuntimeError(f"{what}: {'argument' if rc < 0 else 'CUDA'} error {rc}: "
                           f"{load().segloss_b200_last_error().decode('utf-8', 'replace')}")


class FusedSegLoss(torch.autograd.Function):
    """(pred [N,C,h,w] f32, target [N,h*s,w*s] i64, class_weights [C] f32, scale s) -> (total, ce, dice)
    with total = ce + 0.5 * dice as in SegmentationLoss.forward (seg_diceloss_yolov5.py:731-733).
    Only `total` is differentiable, and only with respect to `pred`."""
    EPS = 1e-6

    @staticmethod
    def forward(ctx, pred, target, class_weights, scale):
        if not pred.is_cuda:
            raise NotImplementedError("the fused segmentation loss has no CPU path")
        pred = pred.contiguous()
        target = target.contiguous()
        n, c, h, w = pred.shape
        if pred.dtype != torch.float32 or target.dtype != torch.int64 or class_weights.dtype != torch.float32:
            raise TypeError("pred and class_weights must be float32 and target int64")
        if tuple(target.shape) != (n, h * scale, w * scale):
            raise ValueError(f"target {tuple(target.shape)} does not match pred {tuple(pred.shape)} x scale {scale}")
        stats = torch.empty(n * 3 * c + 2, dtype=torch.float64, device=pred.device)
        with torch.cuda.device_of(pred):
            _check(load().segloss_b200_forward(pred.data_ptr(), target.data_ptr(), class_weights.data_ptr(),
                                               stats.data_ptr(), n, c, h, w, scale,
                                               torch.cuda.current_stream().cuda_stream), "segloss_b200_forward")
        ipo = stats[:n * 3 * c].view(n, 3, c)
        inter, psum, osum = ipo[:, 0], ipo[:, 1], ipo[:, 2]
        denom = psum + osum + FusedSegLoss.EPS
        dice = 1.0 - ((2.0 * inter + FusedSegLoss.EPS) / denom).mean()
        ce = stats[-2] / stats[-1]
        total = ce + 0.5 * dice
        # Labels outside [0, C) (a 255 "void" label, negatives) match no class in the kernels and would silently drop
        # out of every sum, where the reference's F.cross_entropy / scatter_ raise.  Raising needs a device->host sync
        # per step; instead the per-class label counts must add up to the number of labels or the loss is NaN (loud
        # in any training loop, no synchronisation).
        n_labels = float(n) * (h * scale) * (w * scale)
        total = torch.where(osum.sum() == n_labels, total, torch.full_like(total, float("nan")))
        ctx.save_for_backward(pred, target, class_weights, inter, denom, stats[-1])
        ctx.scale = scale
        ctx.mark_non_differentiable(ce, dice)
        return total.float(), ce.float(), dice.float()

    @staticmethod
    def backward(ctx, g_total, _g_ce, _g_dice):
        pred, target, cw, inter, denom, ce_den = ctx.saved_tensors
        n, c, h, w = pred.shape
        g = g_total.double()
        wd = cw.double().view(1, c)
        k = 0.5 * g / (n * c)                                    # d total / d dice[n,c] = -0.5 / (N C)
        a = -k * 2.0 * wd / denom                                # coefficient of [t = c]
        b = k * wd * (2.0 * inter + FusedSegLoss.EPS) / (denom * denom)
        coef = torch.cat([torch.stack([a, b], 1).reshape(-1), (g / ce_den).reshape(1)]).float()
        grad = torch.empty_like(pred)
        with torch.cuda.device_of(pred):
            _check(load().segloss_b200_backward(pred.data_ptr(), target.data_ptr(), cw.data_ptr(), coef.data_ptr(),
                                                grad.data_ptr(), n, c, h, w, ctx.scale,
                                                torch.cuda.current_stream().cuda_stream), "segloss_b200_backward")
        return grad, None, None, None
