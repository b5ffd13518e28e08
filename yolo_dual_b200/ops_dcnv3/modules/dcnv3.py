"""DCNv3 layer on top of the B200 kernels.

Drop-in for the reference's `DCNv3` nn.Module
(/root/reference/models/ops_dcnv3/build/lib.linux-x86_64-cpython-38/modules/dcnv3.py:50-135): same
constructor arguments and defaults, same public attributes, and the same sub-module names
(`dw_conv.conv`, `dw_conv.bn`, `offset`, `mask`, `input_proj`, `output_proj`), so `state_dict`s and pickled
checkpoints interchange.  NHWC in, NHWC out.

Behaviour kept from the reference on purpose:
  * `dilation` is accepted and then forced to 1 (dcnv3.py:82 there);
  * `act_layer` / `norm_layer` are accepted and ignored — this fork puts a depthwise Conv+BN+SiLU in front of
    the offset / mask heads (dcnv3.py:88 there);
  * the offset and mask heads start at zero and the projections at Xavier-uniform (dcnv3.py:99-107 there):
    a fresh layer samples the regular grid with weight 1/K^2, i.e. it is a K x K average.

Added here: `fused_softmax=True` hands the mask *logits* to the kernels, which apply the softmax over the
K*K points of each group themselves, forward and backward (the reference calls F.softmax, dcnv3.py:122-123).
`packed_heads=True` (needs `fused_softmax`) computes the offset and mask heads with ONE GEMM — the two Linear layers keep
their parameters and names, their weights are stacked per call — and hands the [N, H, W, 3*G*K*K] result to the kernels
as it is: no split copies in the forward, one `grad_heads` tensor (no cat) in the backward.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn.functional as F
from torch import nn

from ..functions import DCNv3Function, DCNv3PackedFunction, DCNv3SoftmaxFunction
from .conv import Conv, autopad  # noqa: F401  (re-exported: the reference module exposes both names)

IM2COL_STEP = 256  # what the reference module always passes (dcnv3.py:133); a no-op here
MAX_FUSED_SOFTMAX_POINTS = 49  # kMaxSoftmaxP in csrc/dcnv3_kernels.cuh (forward and backward check the same bound)


def _is_power_of_2(n):
    if not isinstance(n, int) or n < 0:
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return n != 0 and n & (n - 1) == 0


class DCNv3(nn.Module):
    def __init__(self, channels=64, kernel_size=3, stride=1, pad=1, dilation=1, group=4, offset_scale=1.0,
                 act_layer='GELU', norm_layer='LN', fused_softmax=False, packed_heads=False):
        """
        channels       C of the NHWC input (= group * group_channels)
        kernel_size    K: K*K sampling points per group
        stride, pad    as in a convolution; pad defaults to 1, so only K = 3 keeps H x W
        dilation       ignored (see module docstring)
        group          G; group_channels = C // G — a multiple of 8 keeps the vector kernels
        offset_scale   scale applied to the learned offsets
        fused_softmax  softmax over the K*K points inside the CUDA kernels
        packed_heads   offset + mask heads as one GEMM whose output the kernels read in place (needs fused_softmax)
        """
        super().__init__()
        if channels % group:
            raise ValueError(f'channels must be divisible by group, but got {channels} and {group}')
        if fused_softmax and kernel_size * kernel_size > MAX_FUSED_SOFTMAX_POINTS:
            raise ValueError(f'fused_softmax supports at most {MAX_FUSED_SOFTMAX_POINTS} sampling points per group '
                             f'(kernel_size {kernel_size} has {kernel_size * kernel_size})')
        if not _is_power_of_2(channels // group):
            warnings.warn("You'd better set channels in DCNv3 to make the dimension of each attention head a "
                          "power of 2 which is more efficient in our CUDA implementation.")
        self.channels, self.group, self.group_channels = channels, group, channels // group
        self.kernel_size, self.stride, self.pad = kernel_size, stride, pad
        self.dilation = 1
        self.offset_scale = offset_scale
        self.fused_softmax = bool(fused_softmax)
        self.packed_heads = bool(packed_heads)
        if self.packed_heads and not self.fused_softmax:
            raise ValueError('packed_heads=True needs fused_softmax=True (an unfused softmax would have to rewrite the '
                             'mask part of the packed tensor)')

        points = group * kernel_size * kernel_size
        self.dw_conv = Conv(channels, channels, kernel_size, g=channels)
        self.offset = nn.Linear(channels, 2 * points)
        self.mask = nn.Linear(channels, points)
        self.input_proj = nn.Linear(channels, channels)
        self.output_proj = nn.Linear(channels, channels)
        self._reset_parameters()

    def _reset_parameters(self):
        for head in (self.offset, self.mask):
            nn.init.zeros_(head.weight)
            nn.init.zeros_(head.bias)
        for proj in (self.input_proj, self.output_proj):
            nn.init.xavier_uniform_(proj.weight)
            nn.init.zeros_(proj.bias)

    def _sampling_heads(self, feat: torch.Tensor, dtype: torch.dtype):
        """Offsets and per-group point weights from the depthwise features; both contiguous, in `dtype`."""
        n, h, w, _ = feat.shape
        offset = self.offset(feat).to(dtype)
        mask = self.mask(feat)
        if not self.fused_softmax:
            mask = F.softmax(mask.reshape(n, h, w, self.group, -1), -1).reshape(n, h, w, -1)
        return offset.contiguous(), mask.to(dtype).contiguous()

    def forward(self, input):
        """(N, H, W, C) -> (N, Ho, Wo, C)"""
        value = self.input_proj(input)
        # depthwise 3x3 works on NCHW views; with a channels-last model both permutes are free
        feat = self.dw_conv(input.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        k, s, p, d = self.kernel_size, self.stride, self.pad, self.dilation
        if self.packed_heads:  # one GEMM for both heads; the kernels split its output by address
            heads = F.linear(feat, torch.cat((self.offset.weight, self.mask.weight), 0),
                             torch.cat((self.offset.bias, self.mask.bias), 0))
            if heads.dtype != value.dtype:
                heads = heads.to(value.dtype)
            sampled = DCNv3PackedFunction.apply(value.contiguous(), heads.contiguous(), k, k, s, s, p, p, d, d, self.group,
                                                self.group_channels, self.offset_scale, IM2COL_STEP, True)
            return self.output_proj(sampled)
        offset, mask = self._sampling_heads(feat, value.dtype)
        core = DCNv3SoftmaxFunction if self.fused_softmax else DCNv3Function
        sampled = core.apply(value.contiguous(), offset, mask, k, k, s, s, p, p, d, d,
                             self.group, self.group_channels, self.offset_scale, IM2COL_STEP)
        return self.output_proj(sampled)
