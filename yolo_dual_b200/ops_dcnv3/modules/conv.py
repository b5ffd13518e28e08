"""Conv2d + BatchNorm2d + SiLU building block (YOLOv5 style) used in front of DCNv3's offset / mask heads
and by the YOLO glue blocks.  Attribute names (`conv`, `bn`, `act`) are part of the checkpoint format of the
reference (modules/dcnv3.py:26-40 there defines an identical-looking local class), so they are kept."""
from __future__ import annotations

from torch import nn


def same_padding(kernel, padding=None, dilation=1):
    """Padding that keeps H x W for an odd kernel: (effective kernel) // 2 unless given."""
    if padding is not None:
        return padding
    eff = (lambda k: dilation * (k - 1) + 1) if dilation > 1 else (lambda k: k)
    return eff(kernel) // 2 if isinstance(kernel, int) else [eff(k) // 2 for k in kernel]


autopad = same_padding  # the name the reference uses


class Conv(nn.Module):
    """args: (ch_in, ch_out, kernel, stride, padding, groups, dilation, activation)"""
    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, same_padding(k, p, d), dilation=d, groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        if act is True:
            self.act = self.default_act
        else:
            self.act = act if isinstance(act, nn.Module) else nn.Identity()

    def forward(self, x, out=None):
        """`out` (inference only): a channel slice of a wider channels-last tensor to write the result into, so that the
        caller's torch.cat of this block's output costs nothing."""
        y = self.conv(x)
        if y.is_cuda:
            from ... import _bnact   # fused BatchNorm + SiLU (csrc/bnact_b200.cu) where it applies
            if self.training:        # batch statistics, with autograd
                if out is None and _bnact.usable(y, self.bn, self.act):
                    return _bnact.bn_act(y, self.bn, self.act)
            elif _bnact.usable_eval(y, self.bn, self.act):   # inference: running statistics, one pass
                if out is None or _bnact.slice_pitch(out) is not None:
                    return _bnact.bn_act_eval(y, self.bn, self.act, out)
        z = self.act(self.bn(y))
        if out is not None:
            out.copy_(z)
            return out
        return z

    def forward_into(self, x, buf, c0):
        """Training-time counterpart of `forward(x, out=...)`: the block's result becomes channels [c0, c0 + c2) of `buf`
        (a dense channels-last tensor standing for a torch.cat), with autograd; returns the tensor to use as `buf` from
        here on."""
        y = self.conv(x)
        if y.is_cuda and self.training:
            from ... import _bnact
            if _bnact.usable(y, self.bn, self.act) and buf.dtype == y.dtype:
                return _bnact.bn_act_into(y, self.bn, self.act, buf, c0)
        z = self.act(self.bn(y))           # anything the fused kernels do not take: autograd's own slice assignment
        buf = buf.clone() if buf.requires_grad and buf.is_leaf else buf
        buf[:, c0:c0 + z.shape[1]] = z.to(buf.dtype)
        return buf

    def forward_fuse(self, x):  # after conv+bn folding
        return self.act(self.conv(x))
