"""DCNv3Function — autograd front of the B200-native DCNv3 core.

Mirrors the reference's ``DCNv3Function``
(/root/reference/models/ops_dcnv3/build/lib.linux-x86_64-cpython-38/functions/dcnv3_func.py:19-89):
same class name, same 15 positional arguments in the same order, same return
arity from ``backward`` (3 grads + 12 ``None``), same ONNX symbolic.  Underneath,
the pybind module ``DCNv3`` (dcnv3_func.py:16,39,54) is replaced by the C-ABI
library ``libdcnv3_b200.so`` (include/dcnv3_b200.h) called through ctypes.

Deliberate differences, all documented in DESIGN.md:
  * bf16 is accepted (the reference dispatches double/float/half only,
    src/cuda/dcnv3_cuda.cu:69,147);
  * ``im2col_step`` keeps its position and the reference's divisibility check
    (dcnv3_cuda.cu:46-49) but the whole batch is always one launch;
  * 16-bit gradients come back in the storage dtype; grad_input is summed in fp32 per tile of output
    pixels and across tiles in the storage dtype (``set_grad_accum('tile')``, the default), or in an fp32
    workspace with one rounding (``'opmath'``, what dcnv3_cuda.cu:126-133,168-170 does), or per contribution
    in the storage dtype (``'storage'``);
  * ``DCNv3SoftmaxFunction`` is the same op with the softmax over the sampling
    points fused in (``mask`` carries logits).

``dcnv3_core_pytorch`` (dcnv3_func.py:148-189) is NOT exported from the product
package: it is the oracle and lives under ``oracle/`` as test infrastructure.
There is no CPU path (reference: src/cpu/dcnv3_cpu.cpp:25,36 raise too).
"""
from __future__ import annotations

import ctypes
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import _lib

_DTYPES = {
    torch.float32: _lib.F32,
    torch.float16: _lib.F16,
    torch.bfloat16: _lib.BF16,
    torch.float64: _lib.F64,
}

_GRAD_ACCUM = {"opmath": _lib.ACC_OPMATH, "storage": _lib.ACC_STORAGE, "tile": _lib.ACC_TILE}
_grad_accum = _GRAD_ACCUM[os.environ.get("DCNV3_B200_GRAD_ACCUM", "tile")]


def set_grad_accum(mode: str) -> None:
    """How grad_input is accumulated for fp16/bf16 storage (include/dcnv3_b200.h).

    'tile'    (default) one kernel, no workspace: fp32 sums per 8x8 tile of output pixels on the SM, the tile's
              window leaves as packed 16-bit reductions -> at most four roundings per element (rtol 1e-2 bar);
              shapes the tile kernel does not take (group_channels != 16, not 3x3 s1 d1) run as 'opmath';
    'opmath'  fp32 workspace, rounded once — the reference's semantics (dcnv3_cuda.cu:126-133,168-170);
    'storage' packed 16-bit vector reductions per contribution straight into grad_input (~36 roundings)."""
    global _grad_accum
    _grad_accum = _GRAD_ACCUM[mode]


def get_grad_accum() -> str:
    return {v: k for k, v in _GRAD_ACCUM.items()}[_grad_accum]


def _geometry(input, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
              group, group_channels, offset_scale):
    N, H, W, _ = input.shape
    return _lib.Geometry(int(N), int(H), int(W), int(kernel_h), int(kernel_w), int(stride_h),
                         int(stride_w), int(pad_h), int(pad_w), int(dilation_h), int(dilation_w),
                         int(group), int(group_channels), float(offset_scale))


def _check_inputs(input, offset, mask, geo, im2col_step):
    """The reference's AT_ASSERTMs (dcnv3_cuda.cu:29-34,48-53) plus the shape algebra it left
    to luck; raises the same exception types."""
    for name, t in (("input", input), ("offset", offset), ("mask", mask)):
        if not t.is_cuda:
            # reference: AT_ERROR("Not implement on cpu"), src/cpu/dcnv3_cpu.cpp:25,36
            raise NotImplementedError(f"Not implement on cpu ({name} is on {t.device}); "
                                      "dcnv3_b200 has no CPU path")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
    if input.dim() != 4 or offset.dim() != 4 or mask.dim() != 4:
        raise RuntimeError("input, offset and mask must be 4-D channel-last tensors")
    if not (input.dtype == offset.dtype == mask.dtype):
        raise RuntimeError(f"input/offset/mask dtypes differ: {input.dtype}, {offset.dtype}, {mask.dtype}")
    if input.dtype not in _DTYPES:
        raise RuntimeError(f"unsupported dtype {input.dtype}")
    if not (input.device == offset.device == mask.device):
        raise RuntimeError("input, offset and mask must live on the same device")
    N, H, W, C = input.shape
    if C != geo.group * geo.group_channels:
        raise RuntimeError(f"Input channels and group times group channels wont match: "
                           f"({C} vs {geo.group * geo.group_channels}).")
    step = min(N, int(im2col_step)) if N else 1
    if step <= 0 or N % step != 0:
        raise RuntimeError(f"batch({N}) must divide im2col_step({step})")
    lib = _lib.load()
    ho, wo = ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.dcnv3_b200_output_size(ctypes.byref(geo), ctypes.byref(ho), ctypes.byref(wo)),
               "dcnv3_b200_output_size")
    P = geo.kernel_h * geo.kernel_w
    want_off = (N, ho.value, wo.value, geo.group * P * 2)
    want_mask = (N, ho.value, wo.value, geo.group * P)
    if tuple(offset.shape) != want_off:
        raise RuntimeError(f"offset shape {tuple(offset.shape)} != {want_off}")
    if tuple(mask.shape) != want_mask:
        raise RuntimeError(f"mask shape {tuple(mask.shape)} != {want_mask}")
    return ho.value, wo.value


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _forward(ctx, logits, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h,
             pad_w, dilation_h, dilation_w, group, group_channels, offset_scale, im2col_step):
    ctx.kernel_h, ctx.kernel_w = kernel_h, kernel_w
    ctx.stride_h, ctx.stride_w = stride_h, stride_w
    ctx.pad_h, ctx.pad_w = pad_h, pad_w
    ctx.dilation_h, ctx.dilation_w = dilation_h, dilation_w
    ctx.group, ctx.group_channels = group, group_channels
    ctx.offset_scale, ctx.im2col_step = offset_scale, im2col_step
    geo = _geometry(input, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                    dilation_w, group, group_channels, offset_scale)
    Ho, Wo = _check_inputs(input, offset, mask, geo, im2col_step)
    lib = _lib.load()
    with torch.cuda.device_of(input):
        output = torch.empty((input.shape[0], Ho, Wo, input.shape[3]), dtype=input.dtype,
                             device=input.device)
        rc = lib.dcnv3_b200_forward(input.data_ptr(), offset.data_ptr(), mask.data_ptr(),
                                    output.data_ptr(), _DTYPES[input.dtype], ctypes.byref(geo),
                                    int(logits), _stream(input.device))
    _lib.check(rc, "dcnv3_b200_forward")
    ctx.save_for_backward(input, offset, mask)
    return output


def _backward(ctx, logits, grad_output):
    input, offset, mask = ctx.saved_tensors
    grad_output = grad_output.contiguous()  # dcnv3_func.py:58
    if grad_output.dtype != input.dtype:
        grad_output = grad_output.to(input.dtype)
    geo = _geometry(input, ctx.kernel_h, ctx.kernel_w, ctx.stride_h, ctx.stride_w, ctx.pad_h,
                    ctx.pad_w, ctx.dilation_h, ctx.dilation_w, ctx.group, ctx.group_channels,
                    ctx.offset_scale)
    lib = _lib.load()
    dt = _DTYPES[input.dtype]
    with torch.cuda.device_of(input):
        grad_input = torch.empty_like(input)
        grad_offset = torch.empty_like(offset)
        grad_mask = torch.empty_like(mask)
        accum = _grad_accum
        if accum == _lib.ACC_TILE and ((input.data_ptr() | grad_output.data_ptr() | offset.data_ptr()) & 15 or mask.data_ptr() & 7):
            accum = _lib.ACC_OPMATH  # views at odd byte offsets: the tile kernel needs 16-byte channel vectors
        ws_bytes = lib.dcnv3_b200_backward_workspace_bytes(dt, ctypes.byref(geo), accum)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=input.device) if ws_bytes else None
        rc = lib.dcnv3_b200_backward(
            input.data_ptr(), offset.data_ptr(), mask.data_ptr(), grad_output.data_ptr(),
            grad_input.data_ptr(), grad_offset.data_ptr(), grad_mask.data_ptr(),
            ws.data_ptr() if ws is not None else None, ws_bytes, dt, ctypes.byref(geo),
            int(logits), accum, _stream(input.device))
    _lib.check(rc, "dcnv3_b200_backward")
    return (grad_input, grad_offset, grad_mask,
            None, None, None, None, None, None, None, None, None, None, None, None)


def _symbolic(g, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
              dilation_h, dilation_w, group, group_channels, offset_scale, im2col_step):
    """ONNX node of the reference (dcnv3_func.py:63-89): mmdeploy::TRTDCNv3."""
    return g.op(
        "mmdeploy::TRTDCNv3", input, offset, mask,
        kernel_h_i=int(kernel_h), kernel_w_i=int(kernel_w),
        stride_h_i=int(stride_h), stride_w_i=int(stride_w),
        pad_h_i=int(pad_h), pad_w_i=int(pad_w),
        dilation_h_i=int(dilation_h), dilation_w_i=int(dilation_w),
        group_i=int(group), group_channels_i=int(group_channels),
        offset_scale_f=float(offset_scale), im2col_step_i=int(im2col_step))


class DCNv3Function(Function):
    """``DCNv3Function.apply(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h,
    pad_w, dilation_h, dilation_w, group, group_channels, offset_scale, im2col_step)``"""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                dilation_h, dilation_w, group, group_channels, offset_scale, im2col_step):
        return _forward(ctx, False, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w,
                        pad_h, pad_w, dilation_h, dilation_w, group, group_channels, offset_scale,
                        im2col_step)

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        return _backward(ctx, False, grad_output)

    symbolic = staticmethod(_symbolic)


class DCNv3SoftmaxFunction(Function):
    """Same signature; ``mask`` holds the pre-softmax logits and the softmax over the
    kernel_h*kernel_w sampling points of each group (modules/dcnv3.py:122-123 in the reference)
    runs inside the kernels.  ``backward`` returns the gradient w.r.t. the logits."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                dilation_h, dilation_w, group, group_channels, offset_scale, im2col_step):
        return _forward(ctx, True, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w,
                        pad_h, pad_w, dilation_h, dilation_w, group, group_channels, offset_scale,
                        im2col_step)

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        return _backward(ctx, True, grad_output)


class DCNv3PackedFunction(Function):
    """``DCNv3PackedFunction.apply(input, heads, kernel_h, ..., offset_scale, im2col_step, mask_is_logits)``

    `heads` [N, Ho, Wo, 3*G*P]: per pixel the G*P*2 offsets followed by the G*P masks (or mask logits) — the output of
    ONE Linear(C, 3*G*P) holding the reference's two heads (LIB/modules/dcnv3.py:121-123) stacked.  The kernels read it
    with a pixel pitch (no split copies) and the backward writes ONE `grad_heads` tensor in the same layout (no cat), so
    the Linear's backward is a single pair of GEMMs.  Same arithmetic as DCNv3Function / DCNv3SoftmaxFunction with
    grad_accum 'tile'.  Shapes the staged-window kernels do not take (not 16-bit, group_channels != 16, not 3x3 s1 d1,
    group % 8, unaligned views) go through those two functions on split copies instead — same results."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, input, heads, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                group, group_channels, offset_scale, im2col_step, mask_is_logits):
        geo = _geometry(input, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                        group_channels, offset_scale)
        P = int(kernel_h) * int(kernel_w)
        n_off = group * P * 2
        if heads.dim() != 4 or heads.shape[-1] != 3 * group * P:
            raise RuntimeError(f"heads must be [N, Ho, Wo, {3 * group * P}], got {tuple(heads.shape)}")
        Ho, Wo = _check_inputs(input, _FakeShape(heads, n_off), _FakeShape(heads, group * P), geo, im2col_step)
        if tuple(heads.shape[:3]) != (input.shape[0], Ho, Wo):
            raise RuntimeError(f"heads shape {tuple(heads.shape)} does not match the output size ({Ho}, {Wo})")
        if not heads.is_contiguous():
            raise RuntimeError("heads tensor has to be contiguous")
        ctx.args = (kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                    group_channels, offset_scale, im2col_step)
        ctx.logits, ctx.n_off = bool(mask_is_logits), n_off
        lib = _lib.load()
        with torch.cuda.device_of(input):
            output = torch.empty((input.shape[0], Ho, Wo, input.shape[3]), dtype=input.dtype, device=input.device)
            rc = lib.dcnv3_b200_forward_packed(input.data_ptr(), heads.data_ptr(), output.data_ptr(),
                                               _DTYPES[input.dtype], ctypes.byref(geo), int(ctx.logits),
                                               _stream(input.device))
        ctx.packed = rc == 0
        if rc == _lib.ENOTSUP:  # split copies through the unpacked entry point (same kernels' generic siblings)
            off, msk = heads[..., :n_off].contiguous(), heads[..., n_off:].contiguous()
            with torch.cuda.device_of(input):
                rc = lib.dcnv3_b200_forward(input.data_ptr(), off.data_ptr(), msk.data_ptr(), output.data_ptr(),
                                            _DTYPES[input.dtype], ctypes.byref(geo), int(ctx.logits),
                                            _stream(input.device))
        _lib.check(rc, "dcnv3_b200_forward_packed")
        ctx.save_for_backward(input, heads)
        return output

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        input, heads = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        if grad_output.dtype != input.dtype:
            grad_output = grad_output.to(input.dtype)
        a = ctx.args
        geo = _geometry(input, *a[:11])
        lib = _lib.load()
        rc = _lib.ENOTSUP
        if ctx.packed and not ((input.data_ptr() | grad_output.data_ptr() | heads.data_ptr()) & 15):
            with torch.cuda.device_of(input):
                grad_input, grad_heads = torch.empty_like(input), torch.empty_like(heads)
                rc = lib.dcnv3_b200_backward_packed(input.data_ptr(), heads.data_ptr(), grad_output.data_ptr(),
                                                    grad_input.data_ptr(), grad_heads.data_ptr(),
                                                    _DTYPES[input.dtype], ctypes.byref(geo), int(ctx.logits),
                                                    _stream(input.device))
        if rc == _lib.ENOTSUP:
            class _C:  # the unpacked backward on split copies, gradients concatenated
                pass
            c = _C()
            (c.kernel_h, c.kernel_w, c.stride_h, c.stride_w, c.pad_h, c.pad_w, c.dilation_h, c.dilation_w, c.group,
             c.group_channels, c.offset_scale, c.im2col_step) = a
            c.saved_tensors = (input, heads[..., :ctx.n_off].contiguous(), heads[..., ctx.n_off:].contiguous())
            gi, go_, gm = _backward(c, ctx.logits, grad_output)[:3]
            return (gi, torch.cat((go_, gm), -1)) + (None,) * 13
        _lib.check(rc, "dcnv3_b200_backward_packed")
        return (grad_input, grad_heads) + (None,) * 13


class _FakeShape:
    """Stands in for the offset / mask slice of a packed heads tensor in `_check_inputs` (shape algebra only)."""

    def __init__(self, heads, width):
        self._h, self.shape = heads, tuple(heads.shape[:3]) + (width,)
        self.is_cuda, self.dtype, self.device = heads.is_cuda, heads.dtype, heads.device

    def is_contiguous(self):
        return True

    def dim(self):
        return 4


def dcnv3_debug_indices(offset, H, W, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                        dilation_h, dilation_w, group, offset_scale):
    """The integer contract of the kernels: (hw_low int32 [N,Ho,Wo,G,P,2], bounds uint8
    [N,Ho,Wo,G,P]) for a CUDA ``offset`` tensor.  See include/dcnv3_b200.h."""
    if not offset.is_cuda:
        raise NotImplementedError("Not implement on cpu")
    offset = offset.contiguous()
    N, Ho, Wo, _ = offset.shape
    P = kernel_h * kernel_w
    geo = _lib.Geometry(int(N), int(H), int(W), int(kernel_h), int(kernel_w), int(stride_h),
                        int(stride_w), int(pad_h), int(pad_w), int(dilation_h), int(dilation_w),
                        int(group), 1, float(offset_scale))
    lib = _lib.load()
    with torch.cuda.device_of(offset):
        hw = torch.empty((N, Ho, Wo, group, P, 2), dtype=torch.int32, device=offset.device)
        bd = torch.empty((N, Ho, Wo, group, P), dtype=torch.uint8, device=offset.device)
        rc = lib.dcnv3_b200_debug_indices(offset.data_ptr(), hw.data_ptr(), bd.data_ptr(),
                                          _DTYPES[offset.dtype], ctypes.byref(geo),
                                          _stream(offset.device))
    _lib.check(rc, "dcnv3_b200_debug_indices")
    return hw, bd
