"""DCNv3 nn.Module — same constructor, attribute names and forward contract as the reference
(/root/reference/models/ops_dcnv3/build/lib.linux-x86_64-cpython-38/modules/dcnv3.py:50-135), so
state_dicts and pickled checkpoints interchange: ``dw_conv.conv``, ``dw_conv.bn``, ``offset``,
``mask``, ``input_proj``, ``output_proj``.

Quirks of the reference kept on purpose:
  * ``self.dilation`` is 1 whatever the constructor was given (dcnv3.py:82);
  * ``act_layer`` / ``norm_layer`` are accepted and unused (dcnv3.py:54) — this fork uses a
    depthwise Conv+BN+SiLU in front of the offset/mask heads (dcnv3.py:88);
  * offset and mask heads start at zero (dcnv3.py:100-103): a fresh layer is a 3x3 average.

One addition: ``fused_softmax=True`` hands the mask logits to the kernels, which do the softmax
over the sampling points themselves (forward and backward) instead of ``F.softmax`` (dcnv3.py:122-123).
"""
from __future__ import annotations

import warnings

import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from ..functions import DCNv3Function, DCNv3SoftmaxFunction


def autopad(k, p=None, d=1):
    """'same' padding for kernel k, dilation d (dcnv3.py:17-23)."""
    if d > 1:
        k = d * (k - 1) + 1 if isinstance(k, int) else [d * (x - 1) + 1 for x in k]
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


class Conv(nn.Module):
    """Conv2d + BatchNorm2d + SiLU, YOLOv5 style (dcnv3.py:26-40)."""
    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))

    def forward_fuse(self, x):
        return self.act(self.conv(x))


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class DCNv3(nn.Module):
    def __init__(self, channels=64, kernel_size=3, stride=1, pad=1, dilation=1, group=4,
                 offset_scale=1.0, act_layer='GELU', norm_layer='LN', fused_softmax=False):
        """
        :param channels      C of the NHWC input
        :param kernel_size   K: K*K sampling points per group
        :param stride, pad   as in a convolution (pad defaults to 1: only K=3 keeps H, W)
        :param dilation      accepted, ignored (reference dcnv3.py:82)
        :param group         G; group_channels = C // G, best a multiple of 8 (one 16-byte vector
                             of fp16/bf16) and a power of two
        :param offset_scale  scale of the learned offsets
        :param fused_softmax softmax over the K*K points inside the CUDA kernels
        """
        super().__init__()
        if channels % group != 0:
            raise ValueError(f'channels must be divisible by group, but got {channels} and {group}')
        _d_per_group = channels // group
        if not _is_power_of_2(_d_per_group):
            warnings.warn(
                "You'd better set channels in DCNv3 to make the dimension of each attention head a "
                "power of 2 which is more efficient in our CUDA implementation.")

        self.offset_scale = offset_scale
        self.channels = channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.dilation = 1
        self.pad = pad
        self.group = group
        self.group_channels = channels // group
        self.fused_softmax = bool(fused_softmax)

        self.dw_conv = Conv(channels, channels, kernel_size, g=channels)
        self.offset = nn.Linear(channels, group * kernel_size * kernel_size * 2)
        self.mask = nn.Linear(channels, group * kernel_size * kernel_size)
        self.input_proj = nn.Linear(channels, channels)
        self.output_proj = nn.Linear(channels, channels)
        self._reset_parameters()

    def _reset_parameters(self):
        constant_(self.offset.weight.data, 0.)
        constant_(self.offset.bias.data, 0.)
        constant_(self.mask.weight.data, 0.)
        constant_(self.mask.bias.data, 0.)
        xavier_uniform_(self.input_proj.weight.data)
        constant_(self.input_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def forward(self, input):
        """
        :param input   (N, H, W, C)
        :return        (N, Ho, Wo, C)
        """
        N, H, W, _ = input.shape

        x = self.input_proj(input)
        dtype = x.dtype

        x1 = self.dw_conv(input.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        offset = self.offset(x1)
        mask = self.mask(x1)
        if self.fused_softmax:
            fn = DCNv3SoftmaxFunction
            mask = mask.type(dtype)
        else:
            fn = DCNv3Function
            mask = F.softmax(mask.reshape(N, H, W, self.group, -1), -1).reshape(N, H, W, -1).type(dtype)
        if offset.dtype != dtype:
            offset = offset.type(dtype)

        x = fn.apply(
            x, offset.contiguous(), mask.contiguous(),
            self.kernel_size, self.kernel_size,
            self.stride, self.stride,
            self.pad, self.pad,
            self.dilation, self.dilation,
            self.group, self.group_channels,
            self.offset_scale,
            256)
        return self.output_proj(x)
