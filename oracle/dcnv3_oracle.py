"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU oracles for the DCNv3 deformable-sampling core.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module, and only as the checker / the reported
CPU baseline.  Nothing under ``yolo_dual_b200/`` imports it.

Two oracles live here (paths relative to /root/reference/models/ops_dcnv3/):

``core_torch``
    A restatement of the reference's pure-PyTorch ``dcnv3_core_pytorch``
    (build/lib.linux-x86_64-cpython-38/functions/dcnv3_func.py:148-189, helpers
    :92-120 and :123-145): normalised sampling coordinates + ``F.grid_sample``.
    This is the float oracle north_star names and the CPU baseline bench.py
    times.  Its bilinear arithmetic lives in PyTorch (``grid_sample``,
    ``pad``) — a third-party dependency of the reference (requirements.txt:16
    pins ``torch>=1.7.0``; this image has 2.11.0).

``PixelOracle`` (C, ``oracle/dcnv3_oracle.c``)
    A restatement of the reference CUDA kernels' pixel-space arithmetic
    (src/cuda/dcnv3_im2col_cuda.cuh).  It defines the integer contract
    (h_low, w_low, bounds byte) and gives forward/backward in the reference
    CUDA's operation order, in float or double op-math.

Parity pin: both are checked against golden vectors generated from the
reference's own ``dcnv3_core_pytorch`` (tests/golden/make_golden.py, run where
/root/reference is mounted; fixtures committed under tests/golden/), by
tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdcnv3_oracle.so")


# --------------------------------------------------------------------------
# shape algebra (dcnv3_cuda.cu:40-45)
# --------------------------------------------------------------------------
def output_hw(H, W, kh, kw, sh, sw, ph, pw, dh, dw):
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    return Ho, Wo


# --------------------------------------------------------------------------
# float oracle: restatement of dcnv3_core_pytorch
# --------------------------------------------------------------------------
def _kernel_centres(size_padded, k, dil, stride, n_out):
    """dcnv3_func.py:97-115 — float32 centres of the output pixels in the padded
    frame, normalised by the padded extent.  ``linspace(c0, c0+(n-1)*stride, n)``
    there; every value is a small half-integer, exactly representable, so the
    arange form below is bit-identical."""
    c0 = (dil * (k - 1)) // 2 + 0.5
    pts = c0 + torch.arange(n_out, dtype=torch.float32) * float(stride)
    return pts / size_padded


def _tap_offsets(size_padded, k, dil):
    """dcnv3_func.py:126-140 — float32 tap positions -(dil*(k-1))//2 + t*dil,
    normalised by the padded extent."""
    first = -((dil * (k - 1)) // 2)
    taps = first + torch.arange(k, dtype=torch.float32) * float(dil)
    return taps / size_padded


def core_torch(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
               dilation_h, dilation_w, group, group_channels, offset_scale):
    """Float oracle with the reference's argument order (dcnv3_func.py:148-152).

    Differentiable through torch autograd (as the reference tests use it,
    test.py:110-115).  fp16/bf16 are not accepted: the reference raises a dtype
    error there too (its float32 grid meets a half input inside grid_sample);
    callers run it in fp32 on half-rounded values (see ``core_torch_lowp``).
    """
    # dcnv3_func.py:155-157: the pad list is (C, C, W, W, H, H) ordered, so the
    # reference pads W by pad_h and H by pad_w.  Restated as is.
    x = F.pad(input, [0, 0, pad_h, pad_h, pad_w, pad_w])
    N, Hp, Wp, C = x.shape
    _, Ho, Wo, _ = offset.shape
    G, gc, P = group, group_channels, kernel_h * kernel_w

    # dcnv3_func.py:92-120 (the reference recomputes the output extent from the
    # padded input; offset.shape already carries it).
    cy = _kernel_centres(Hp, kernel_h, dilation_h, stride_h, Ho)
    cx = _kernel_centres(Wp, kernel_w, dilation_w, stride_w, Wo)
    ref = torch.stack((cx.view(1, Wo).expand(Ho, Wo), cy.view(Ho, 1).expand(Ho, Wo)), -1)
    ref = ref.reshape(1, Ho, Wo, 1, 2)

    # dcnv3_func.py:123-145: meshgrid(x over kernel_w, y over kernel_h), 'ij':
    # point p = i_w * kernel_h + j_h, identical for every group.
    tx = _tap_offsets(Wp, kernel_w, dilation_w)
    ty = _tap_offsets(Hp, kernel_h, dilation_h)
    taps = torch.stack((tx.view(kernel_w, 1).expand(kernel_w, kernel_h),
                        ty.view(1, kernel_h).expand(kernel_w, kernel_h)), -1)
    taps = taps.reshape(1, P, 2).repeat(group, 1, 1).reshape(1, 1, 1, G * P, 2)

    # dcnv3_func.py:165-169
    norm = torch.tensor([Wp, Hp]).reshape(1, 1, 1, 2).repeat(1, 1, 1, G * P)
    loc = (ref + taps * offset_scale).repeat(N, 1, 1, 1, 1).flatten(3, 4) \
        + offset * offset_scale / norm

    # dcnv3_func.py:172-181
    grid = (2 * loc - 1).view(N, Ho * Wo, G, P, 2).transpose(1, 2).flatten(0, 1)
    x_g = x.view(N, Hp * Wp, G * gc).transpose(1, 2).reshape(N * G, gc, Hp, Wp)
    sampled = F.grid_sample(x_g, grid, mode="bilinear", padding_mode="zeros",
                            align_corners=False)  # [N*G, gc, Ho*Wo, P]

    # dcnv3_func.py:184-189
    m = mask.view(N, Ho * Wo, G, P).transpose(1, 2).reshape(N * G, 1, Ho * Wo, P)
    out = (sampled * m).sum(-1).view(N, G * gc, Ho * Wo)
    return out.transpose(1, 2).reshape(N, Ho, Wo, -1).contiguous()


def core_torch_fwd_bwd(input, offset, mask, grad_out, *args):
    """Forward + backward(grad_out) through ``core_torch``; returns
    (output, grad_input, grad_offset, grad_mask), all detached."""
    i = input.detach().clone().requires_grad_(True)
    o = offset.detach().clone().requires_grad_(True)
    m = mask.detach().clone().requires_grad_(True)
    out = core_torch(i, o, m, *args)
    out.backward(grad_out)
    return out.detach(), i.grad, o.grad, m.grad


def core_torch_lowp(input, offset, mask, grad_out, *args):
    """Oracle for fp16/bf16 storage: fp32 oracle on the half-rounded values."""
    f = lambda t: None if t is None else t.float()
    if grad_out is None:
        return core_torch(f(input), f(offset), f(mask), *args)
    return core_torch_fwd_bwd(f(input), f(offset), f(mask), f(grad_out), *args)


# --------------------------------------------------------------------------
# pixel-space oracle (C)
# --------------------------------------------------------------------------
def build(force=False):
    """Compile oracle/libdcnv3_oracle.so with the recipe in oracle/Makefile."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


class PixelOracle:
    """ctypes binding of oracle/dcnv3_oracle.c.  CPU tensors only; fp32 or fp64.
    fp16/bf16 callers pass ``tensor.float()`` (op-math of the reference is fp32
    for half storage, dcnv3_im2col_cuda.cuh:30)."""

    _geo = [ctypes.c_int] * 13  # N,H,W,G,gc,kh,kw,sh,sw,ph,pw,dh,dw

    def __init__(self):
        self.lib = ctypes.CDLL(build())
        p = ctypes.c_void_p
        for sfx in ("f32", "f64"):
            getattr(self.lib, f"dcnv3_oracle_forward_{sfx}").argtypes = \
                [p, p, p, p] + self._geo + [ctypes.c_float]
            getattr(self.lib, f"dcnv3_oracle_backward_{sfx}").argtypes = \
                [p] * 7 + self._geo + [ctypes.c_float]
            # indices: no gc
            getattr(self.lib, f"dcnv3_oracle_indices_{sfx}").argtypes = \
                [p, p, p] + [ctypes.c_int] * 12 + [ctypes.c_float]
        self.lib.dcnv3_oracle_num_threads.restype = ctypes.c_int

    @property
    def num_threads(self):
        return int(self.lib.dcnv3_oracle_num_threads())

    def set_threads(self, n):
        self.lib.dcnv3_oracle_set_threads(int(n))

    @staticmethod
    def _sfx(t):
        if t.dtype == torch.float32:
            return "f32"
        if t.dtype == torch.float64:
            return "f64"
        raise TypeError(f"PixelOracle takes fp32/fp64 CPU tensors, got {t.dtype}")

    @staticmethod
    def _chk(*ts):
        for t in ts:
            assert t.device.type == "cpu" and t.is_contiguous(), "CPU contiguous tensors only"

    def forward(self, input, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, G, gc, offset_scale):
        self._chk(input, offset, mask)
        N, H, W, C = input.shape
        assert C == G * gc
        Ho, Wo = output_hw(H, W, kh, kw, sh, sw, ph, pw, dh, dw)
        assert tuple(offset.shape) == (N, Ho, Wo, G * kh * kw * 2), offset.shape
        assert tuple(mask.shape) == (N, Ho, Wo, G * kh * kw), mask.shape
        out = torch.empty(N, Ho, Wo, C, dtype=input.dtype)
        rc = getattr(self.lib, f"dcnv3_oracle_forward_{self._sfx(input)}")(
            input.data_ptr(), offset.data_ptr(), mask.data_ptr(), out.data_ptr(),
            N, H, W, G, gc, kh, kw, sh, sw, ph, pw, dh, dw, float(offset_scale))
        if rc:
            raise RuntimeError(f"dcnv3_oracle_forward rc={rc}")
        return out

    def backward(self, input, offset, mask, grad_out, kh, kw, sh, sw, ph, pw, dh, dw, G, gc,
                 offset_scale):
        self._chk(input, offset, mask, grad_out)
        N, H, W, C = input.shape
        gin = torch.empty_like(input)
        goff = torch.empty_like(offset)
        gmask = torch.empty_like(mask)
        rc = getattr(self.lib, f"dcnv3_oracle_backward_{self._sfx(input)}")(
            input.data_ptr(), offset.data_ptr(), mask.data_ptr(), grad_out.data_ptr(),
            gin.data_ptr(), goff.data_ptr(), gmask.data_ptr(),
            N, H, W, G, gc, kh, kw, sh, sw, ph, pw, dh, dw, float(offset_scale))
        if rc:
            raise RuntimeError(f"dcnv3_oracle_backward rc={rc}")
        return gin, goff, gmask

    def indices(self, offset, H, W, kh, kw, sh, sw, ph, pw, dh, dw, G, offset_scale):
        """-> (hw_low int32 [N,Ho,Wo,G,P,2] as (h_low, w_low), bounds uint8 [N,Ho,Wo,G,P])."""
        self._chk(offset)
        N, Ho, Wo, _ = offset.shape
        P = kh * kw
        assert (Ho, Wo) == output_hw(H, W, kh, kw, sh, sw, ph, pw, dh, dw)
        hw = torch.empty(N, Ho, Wo, G, P, 2, dtype=torch.int32)
        bd = torch.empty(N, Ho, Wo, G, P, dtype=torch.uint8)
        rc = getattr(self.lib, f"dcnv3_oracle_indices_{self._sfx(offset)}")(
            offset.data_ptr(), hw.data_ptr(), bd.data_ptr(),
            N, H, W, G, kh, kw, sh, sw, ph, pw, dh, dw, float(offset_scale))
        if rc:
            raise RuntimeError(f"dcnv3_oracle_indices rc={rc}")
        return hw, bd


# --------------------------------------------------------------------------
# synthetic inputs shared by tests / bench (CPU generator, explicit seed)
# --------------------------------------------------------------------------
def make_inputs(N, H, W, G, gc, kh=3, kw=3, sh=1, sw=1, ph=1, pw=1, dh=1, dw=1, *,
                dist="unit", seed=0, dtype=torch.float32):
    """dist='ref': the reference test's distribution (test.py:35-39):
         input = rand*0.01, offset = rand*10, mask = (rand+1e-5) normalised over P.
       dist='unit': input = randn, offset = randn (sigma = 1 px), mask = softmax(randn).
       Returns (input, offset, mask, grad_out), CPU tensors of ``dtype``."""
    gen = torch.Generator().manual_seed(seed)
    Ho, Wo = output_hw(H, W, kh, kw, sh, sw, ph, pw, dh, dw)
    P, C = kh * kw, G * gc
    if dist == "ref":
        x = torch.rand(N, H, W, C, generator=gen) * 0.01
        off = torch.rand(N, Ho, Wo, G * P * 2, generator=gen) * 10
        m = torch.rand(N, Ho, Wo, G, P, generator=gen) + 1e-5
        m = (m / m.sum(-1, keepdim=True)).reshape(N, Ho, Wo, G * P)
    elif dist == "unit":
        x = torch.randn(N, H, W, C, generator=gen)
        off = torch.randn(N, Ho, Wo, G * P * 2, generator=gen)
        m = torch.softmax(torch.randn(N, Ho, Wo, G, P, generator=gen), -1).reshape(N, Ho, Wo, G * P)
    else:
        raise ValueError(dist)
    go = torch.randn(N, Ho, Wo, C, generator=gen)
    return tuple(t.to(dtype).contiguous() for t in (x, off, m, go))
