"""YOLO glue around the DCNv3 module: the reference ships these only as a paste-in notes file
(/root/reference/models/ops_dcnv3/common and yolo.py:2-38); here they are an importable module
with the same class names, constructor arguments and forward semantics.

Additions, each labelled:
  * ``dcn_group``: the notes forward YOLO's conv-group argument ``g`` (default 1) as the DCNv3
    group count, i.e. one group of 128/256/512 channels, and that is the default here too
    (``dcn_group=None``): a default-built block has the reference's parameter shapes (offset head
    ``[18, C]``, mask head ``[9, C]``) and loads reference-trained weights.  BASELINE's shapes use
    group_channels = 16: ``dcn_group="gc16"`` picks ``channels // 16`` groups (what the seg / detection
    builders of this package pass by default), an integer sets the count (SURVEY §3.4 "Group count").
  * ``C2f_DCNV3``: a *derived* block — no DCNv3 variant of the YOLOv8 ``C2f_DCN``
    (unet-lite/yolo8-seg/seg_diceloss_yolov8.py:431-471) exists in the reference; this one keeps
    C2f's split/concat and swaps each inner block for ``DCNV3_YoLo``.
"""
from __future__ import annotations

import torch
from torch import nn

from .ops_dcnv3.modules.conv import Conv
from .ops_dcnv3.modules.dcnv3 import DCNv3


GC16 = "gc16"  # dcn_group value: group_channels = 16, the BASELINE shapes


def _pick_group(channels: int, g, dcn_group):
    """DCNv3 group count: the reference's `group=g` unless `dcn_group` says otherwise (module docstring)."""
    if dcn_group is None:
        return 1 if g is None else g
    if dcn_group == GC16:
        if g not in (None, 1):
            return g
        return max(channels // 16, 1) if channels % 16 == 0 else 1
    return int(dcn_group)


class DCNV3_YoLo(nn.Module):
    """1x1 Conv -> NHWC -> DCNv3 -> NCHW (common and yolo.py:2-13)."""

    def __init__(self, inc, ouc, k=1, s=1, p=None, g=1, d=1, act=True, dcn_group=None,
                 fused_softmax=False, packed_heads=False):
        super().__init__()
        self.conv = Conv(inc, ouc, k=1)
        self.dcnv3 = DCNv3(ouc, kernel_size=k, stride=s, group=_pick_group(ouc, g, dcn_group),
                           dilation=d, fused_softmax=fused_softmax, packed_heads=packed_heads)

    def forward(self, x):
        x = self.conv(x).permute(0, 2, 3, 1)   # NHWC view
        if not x.is_contiguous():              # a channels-last model never gets here: the view is already dense NHWC
            x = x.contiguous()                 # (NCHW-contiguous model: the reference's permute + copy)
        return self.dcnv3(x).permute(0, 3, 1, 2)   # NCHW view of the NHWC result = a channels-last tensor, no copy


class Bottleneck_DCNV3(nn.Module):
    """cv1 1x1 -> DCNV3_YoLo 3x3, residual when shapes allow (common and yolo.py:15-25)."""

    def __init__(self, c1, c2, shortcut=True, g=1, e=0.5, dcn_group=None, fused_softmax=False, packed_heads=False):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = DCNV3_YoLo(c_, c2, 3, 1, g=g, dcn_group=dcn_group, fused_softmax=fused_softmax,
                              packed_heads=packed_heads)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        return x + self.cv2(self.cv1(x)) if self.add else self.cv2(self.cv1(x))


class C3_DCNV3(nn.Module):
    """CSP bottleneck with 3 convolutions, DCNv3 inner blocks (common and yolo.py:27-38)."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5, dcn_group=None, fused_softmax=False, packed_heads=False):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck_DCNV3(c_, c_, shortcut, g, e=1.0, dcn_group=dcn_group,
                                                  fused_softmax=fused_softmax, packed_heads=packed_heads) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))


class C2f_DCNV3(nn.Module):
    """Derived block (see module docstring): YOLOv8 C2f with DCNv3 inner blocks."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5, dcn_group=None, fused_softmax=False, packed_heads=False):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck_DCNV3(self.c, self.c, shortcut, g, e=1.0, dcn_group=dcn_group,
                                                fused_softmax=fused_softmax, packed_heads=packed_heads) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))
