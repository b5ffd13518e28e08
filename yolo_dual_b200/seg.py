"""C3-DCN semantic-segmentation models, loss and a data-parallel training step — the callers of the
DCNv3 hot path, restated as an importable module.

The reference keeps these inside self-contained trainer scripts that cannot be imported here
(thop / IPython / matplotlib at import time, a CPU thop.profile in the constructor):
  unet-lite/yolo5-seg/seg_diceloss_yolov5.py   blocks :388-507, YOLOv5Seg :511-681, loss :693-750,
                                               optimizer / hot loop :970-980, :1073-1103
  unet-lite/yolo5-seg/yolov5_seg.yaml          layer table (restated below as data)
  unet-lite/yolo8-seg/seg_diceloss_yolov8.py   C2f :400-414, C2f_DCN :431-471
  unet-lite/yolo8-seg/yolov8_seg.yaml
  utils/torch_utils.py                         smart_optimizer :318-346, smart_DDP :55-63

Decisions (SURVEY §3.4 "Integration deltas"):
  * the yaml slots named C3_DCN / C2f_DCN are built as C3_DCNV3 / C2f_DCNV3 (``dcn='dcnv3'``,
    default) — the reference fills them with torchvision DeformConv2d, a different operator;
    ``dcn='none'`` builds plain C3 / C2f there (CPU-runnable: used by the CPU tests);
  * quirks of the reference builders are kept: the yaml repeat count is ignored (:544,557), the
    second yaml argument lands in ``n`` (``[512, False]`` -> n = 0, ``[128, True]`` -> n = 1),
    ``from`` indexes the growing list of all layer outputs, Concat resizes to its first input,
    the model ends in Softmax and the loss applies softmax again (:731-733);
  * no CPU forward in the constructor (the op has no CPU path; reference note
    ``models/ops_dcnv3/common and yolo.py:40-43``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from .blocks import C2f_DCNV3, C3_DCNV3
from .ops_dcnv3.modules.conv import Conv

# ---------------------------------------------------------------------------------------------
# layer tables: [from, repeat(ignored), module, args] — data restated from the reference yamls
# ---------------------------------------------------------------------------------------------
_HEAD_V5 = [
    [-1, 1, "Conv", [512, 1, 1]], [-1, 1, "nn.Upsample", [None, 2, "nearest"]], [6, 1, "Conv", [512, 1, 1]],
    [[-1, 1], 1, "Concat", [1]], [-1, 3, "C3", [512, False]],
    [-1, 1, "Conv", [256, 1, 1]], [-1, 1, "nn.Upsample", [None, 2, "nearest"]], [4, 1, "Conv", [256, 1, 1]],
    [[-1, 6], 1, "Concat", [1]], [-1, 3, "C3", [256, False]],
    [-1, 1, "Conv", [128, 1, 1]], [-1, 1, "nn.Upsample", [None, 2, "nearest"]], [2, 1, "Conv", [128, 1, 1]],
    [[-1, 11], 1, "Concat", [1]], [-1, 3, "C3", [128, False]],
    [-1, 1, "Conv", [64, 3, 1]], [-1, 1, "nn.Upsample", [None, 4, "nearest"]], [-1, 1, "Conv", [12, 1, 1]],
    [-1, 1, "nn.Softmax", [1]],
]
YOLOV5_SEG = {  # unet-lite/yolo5-seg/yolov5_seg.yaml:10-51
    "nc": 12,
    "backbone": [
        [-1, 1, "Conv", [64, 6, 2, 2]], [-1, 1, "Conv", [128, 3, 2]], [-1, 3, "C3", [128]],
        [-1, 1, "Conv", [256, 3, 2]], [-1, 6, "C3_DCN", [256]],
        [-1, 1, "Conv", [512, 3, 2]], [-1, 9, "C3_DCN", [512]],
        [-1, 1, "Conv", [1024, 3, 2]], [-1, 3, "C3_DCN", [1024]], [-1, 1, "SPPF", [1024, 5]],
    ],
    "head": _HEAD_V5,
}
YOLOV8_SEG = {  # unet-lite/yolo8-seg/yolov8_seg.yaml
    "nc": 12,
    "backbone": [
        [-1, 1, "Conv", [64, 3, 2, 1]], [-1, 1, "Conv", [128, 3, 2]], [-1, 3, "C2f", [128, True]],
        [-1, 1, "Conv", [256, 3, 2]], [-1, 6, "C2f_DCN", [256, True]],
        [-1, 1, "Conv", [512, 3, 2]], [-1, 6, "C2f_DCN", [512, True]],
        [-1, 1, "Conv", [1024, 3, 2]], [-1, 3, "C2f_DCN", [1024, True]], [-1, 1, "SPPF", [1024, 5]],
        [-1, 1, "Upsample", [None, 2, "nearest"]],
    ],
    "head": [
        [-1, 1, "Conv", [512, 1, 1]], [-1, 1, "Upsample", [None, 2, "nearest"]], [4, 1, "Conv", [512, 1, 1]],
        [[-1, -2], 1, "Concat", [1]], [-1, 3, "C3", [512, False]],
        [-1, 1, "Conv", [256, 1, 1]], [-1, 1, "Upsample", [None, 2, "nearest"]], [2, 1, "Conv", [256, 1, 1]],
        [[-1, -2], 1, "Concat", [1]], [-1, 3, "C3", [256, False]],
        [-1, 1, "Conv", [128, 1, 1]], [-1, 1, "Upsample", [None, 2, "nearest"]], [0, 1, "Conv", [128, 1, 1]],
        [[-1, -2], 1, "Concat", [1]], [-1, 3, "C3", [128, False]],
        [-1, 1, "Conv", [64, 3, 1]], [-1, 1, "Upsample", [None, 2, "nearest"]], [-1, 1, "Conv", [12, 1, 1]],
        [-1, 1, "nn.Softmax", [1]],
    ],
}
# CamVid class weights, unet-lite/yolo5-seg/weight.yaml
CAMVID_CLASS_WEIGHTS = [1.0, 2.0, 25.0, 2.0, 10.0, 3.0, 25.0, 10.0, 5.0, 15.0, 25.0, 1.0]


# ---------------------------------------------------------------------------------------------
# plain blocks of the seg scripts (outer residual when c1 == c2; no inner bottleneck residual)
# ---------------------------------------------------------------------------------------------
def _cat_free(x: torch.Tensor) -> bool:
    """Inference on a channels-last CUDA activation: `Conv(x, out=slice)` can fill a concatenated buffer in place."""
    return x.is_cuda and not torch.is_grad_enabled() and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)


def _cat_free_train(block: nn.Module, x: torch.Tensor) -> bool:
    """Training on a channels-last CUDA activation with the fused BatchNorm + SiLU in use."""
    from . import _bnact
    return (x.is_cuda and block.training and torch.is_grad_enabled() and x.dim() == 4 and _bnact.enabled()
            and x.is_contiguous(memory_format=torch.channels_last) and x.dtype in (torch.float32, torch.float16, torch.bfloat16))


class C3(nn.Module):
    """seg_diceloss_yolov5.py:415-428"""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Conv(c_, c_, 3, 1, g=g) for _ in range(int(n))))
        self.add = shortcut and c1 == c2

    def forward(self, x):
        if _cat_free(x):  # inference: both branches end in a Conv block, which writes straight into its half of the
            c_ = self.cv2.conv.out_channels   # concatenated tensor (no torch.cat copy)
            buf = torch.empty((x.shape[0], 2 * c_, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device,
                              memory_format=torch.channels_last)
            convs = [self.cv1] + list(self.m)
            t = x
            for conv in convs[:-1]:
                t = conv(t)
            convs[-1](t, out=buf[:, :c_])
            self.cv2(x, out=buf[:, c_:])
            y = self.cv3(buf)
        elif _cat_free_train(self, x):  # training: the same, with autograd (FusedBNActInto chains through the buffer)
            c_ = self.cv2.conv.out_channels
            dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
            buf = torch.empty((x.shape[0], 2 * c_, x.shape[2], x.shape[3]), dtype=dt, device=x.device,
                              memory_format=torch.channels_last)
            convs = [self.cv1] + list(self.m)
            t = x
            for conv in convs[:-1]:
                t = conv(t)
            buf = convs[-1].forward_into(t, buf, 0)
            buf = self.cv2.forward_into(x, buf, c_)
            y = self.cv3(buf)
        else:
            y = self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))
        return y + x if self.add else y


class C2f(nn.Module):
    """seg_diceloss_yolov8.py:400-414"""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        n = int(n)
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Conv(self.c, self.c, 3, 1, g=g) for _ in range(n))
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        out = self.cv2(torch.cat(y, 1))
        return out + x if self.add else out


class SPPF(nn.Module):
    """seg_diceloss_yolov5.py:468-481"""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        x = self.cv1(x)
        y1 = self.m(x)
        y2 = self.m(y1)
        return self.cv2(torch.cat([x, y1, y2, self.m(y2)], 1))


def _resize(x: torch.Tensor, size, mode: str) -> torch.Tensor:
    """F.interpolate(x, size, mode) ('nearest' | 'bilinear' with align_corners=False).

    NHWC-contiguous CUDA activations go through the repo's own kernels (csrc/resize_b200.cu: same index
    arithmetic as ATen, gather backward).  Otherwise torch — but in the tensor's own dtype: CUDA autocast runs the
    upsample_* ops in float32 whatever comes in, and everything downstream of one (torch.cat with bf16 neighbours,
    the copies in front of the next convolution) then moves 4-byte elements.  Same numbers either way: nearest
    replication is exact, and the bilinear kernels interpolate in float and round once — the rounding the next
    convolution's autocast cast would apply."""
    kw = {} if mode == "nearest" else {"align_corners": False}
    if x.is_cuda:
        from . import _resize as native
        if native.usable(x, size, mode):
            return native.resize(x, size, mode)
        if torch.is_autocast_enabled():
            with torch.autocast("cuda", enabled=False):
                return F.interpolate(x, size=tuple(size), mode=mode, **kw)
    return F.interpolate(x, size=tuple(size), mode=mode, **kw)


class Upsample(nn.Upsample):
    """nn.Upsample (the layer tables' 'nn.Upsample' / 'Upsample'): integer-factor nearest replication through
    `_resize`; any other configuration is nn.Upsample itself."""

    def forward(self, x):
        s = self.scale_factor
        if self.mode == "nearest" and s is not None and not isinstance(s, (tuple, list)) and float(s) == int(s):
            return _resize(x, (x.shape[2] * int(s), x.shape[3] * int(s)), "nearest")
        return super().forward(x)


class Concat(nn.Module):
    """Concatenate after resizing every input to the first one's H x W
    (seg_diceloss_yolov5.py:484-507)."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, xs: Sequence[torch.Tensor]):
        size = xs[0].shape[2:]
        xs = [x if x.shape[2:] == size else _resize(x, size, "bilinear") for x in xs]
        return torch.cat(xs, self.d)


class _WithOuterResidual(nn.Module):
    """The seg scripts add `+ x` around their C3_DCN / C2f_DCN when c1 == c2
    (seg_diceloss_yolov5.py:439,465); the paste-in C3_DCNV3 has its residual inside the bottleneck
    instead.  ``outer_residual=True`` adds the scripts' one on top; default is the paste-in as is."""

    def __init__(self, block: nn.Module, add: bool):
        super().__init__()
        self.block, self.add = block, add

    def forward(self, x):
        y = self.block(x)
        return y + x if self.add else y


# ---------------------------------------------------------------------------------------------
# model
# ---------------------------------------------------------------------------------------------
class SegModel(nn.Module):
    """YOLOv5Seg / YOLOv8Seg of the reference (seg_diceloss_yolov5.py:511-659) over a layer table."""

    def __init__(self, cfg: Dict = YOLOV5_SEG, num_classes: Optional[int] = None, dcn: str = "dcnv3",
                 dcn_group="gc16", fused_softmax: bool = False, outer_residual: bool = False, packed_heads: bool = False,
                 img_size: Sequence[int] = (640, 640), defer_upsample: bool = True):
        """defer_upsample: run the pointwise tail of the head (1x1 Conv + BN + SiLU, channel Softmax) BEFORE the last
        nearest Upsample instead of after it.  Every one of those ops commutes with pixel replication — the batch
        statistics of a replicated map are those of the map — so outputs and gradients are the reference's, while
        the 64-channel 640x640 activation (839 MB in bf16 at batch 16) and the BN / SiLU / Softmax passes over it are
        never materialised.  Only BatchNorm's unbiased-variance factor n/(n-1) in `running_var` sees the smaller n
        (relative 2e-6).  Layer indices, parameter names and `state_dict` are unchanged."""
        super().__init__()
        if dcn not in ("dcnv3", "none"):
            raise ValueError("dcn must be 'dcnv3' or 'none'")
        self.cfg = cfg
        self.num_classes = cfg["nc"] if num_classes is None else num_classes
        self.img_size = list(img_size)
        self.dcn, self.dcn_group = dcn, dcn_group
        self.fused_softmax, self.outer_residual, self.packed_heads = fused_softmax, outer_residual, packed_heads
        self.layers = nn.ModuleList()
        self.froms: List = []
        chs: List[int] = []
        table = list(cfg["backbone"]) + list(cfg["head"])
        self.n_backbone = len(cfg["backbone"])
        for i, (frm, _repeat, name, args) in enumerate(table):
            args = list(args)
            if name == "Conv" and i == len(table) - 2 and num_classes is not None:
                args[0] = num_classes
            if isinstance(frm, list):
                c1 = sum(chs[f] for f in frm)
            else:
                c1 = 3 if not chs else chs[frm]
            mod, c2 = self._make(name, c1, args)
            self.layers.append(mod)
            self.froms.append(frm)
            chs.append(c2)
        self._deferred = self._find_deferrable_upsample() if defer_upsample else None
        self._initialize_weights()

    def _find_deferrable_upsample(self) -> Optional[int]:
        """Index of the last nearest Upsample if everything after it is pointwise and reads only its predecessor."""
        n = len(self.layers)
        ups = [i for i, m in enumerate(self.layers) if isinstance(m, nn.Upsample)]
        if not ups:
            return None
        i = ups[-1]
        m = self.layers[i]
        if m.mode != "nearest" or float(m.scale_factor) != int(m.scale_factor) or i == n - 1:
            return None
        for j in range(i + 1, n):
            if self.froms[j] != -1:
                return None
            t = self.layers[j]
            if isinstance(t, Conv):
                c = t.conv
                if c.kernel_size != (1, 1) or c.stride != (1, 1) or c.padding != (0, 0):
                    return None
            elif not (isinstance(t, nn.Softmax) and t.dim == 1):
                return None
        for j, frm in enumerate(self.froms):  # nobody else may read the upsampled map
            for f in (frm if isinstance(frm, list) else [frm]):
                if f != -1 and (f if f >= 0 else j + f) >= i and j > i:
                    return None
        return i

    def _make(self, name, c1, args):
        if name == "Conv":
            return Conv(c1, *args), args[0]
        if name == "C3":
            return C3(c1, *args), args[0]
        if name == "C2f":
            return C2f(c1, *args), args[0]
        if name in ("C3_DCN", "C3_DCNV3", "C2f_DCN", "C2f_DCNV3"):
            c2 = args[0]
            n = int(args[1]) if len(args) > 1 else 1
            is_c3 = name.startswith("C3")
            if self.dcn == "none":
                return (C3 if is_c3 else C2f)(c1, c2, n), c2
            blk = (C3_DCNV3 if is_c3 else C2f_DCNV3)(c1, c2, n=max(n, 1), dcn_group=self.dcn_group,
                                                     fused_softmax=self.fused_softmax, packed_heads=self.packed_heads)
            return _WithOuterResidual(blk, self.outer_residual and c1 == c2), c2
        if name == "SPPF":
            return SPPF(c1, *args), args[0]
        if name in ("Upsample", "nn.Upsample"):
            scale = float(args[1]) if len(args) > 1 and args[1] is not None else 2.0
            mode = args[2] if len(args) > 2 else "nearest"
            return Upsample(scale_factor=scale, mode=mode), c1
        if name == "Concat":
            return Concat(*args), c1
        if name == "nn.Softmax":
            return nn.Softmax(dim=args[0] if args else 1), c1
        raise NotImplementedError(f"unknown module {name}")

    def _initialize_weights(self):
        """seg_diceloss_yolov5.py:661-670 (DCNv3's own reset, modules/dcnv3.py:99-107, already ran)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="leaky_relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x, lowres: bool = False):
        """lowres=True returns `(pred, scale)`: the output before the deferred nearest Upsample and its factor
        (scale 1 and the full-size output when nothing is deferred) for `SegmentationLoss(pred, target, scale)`."""
        outs: List[torch.Tensor] = []
        for i, (layer, frm) in enumerate(zip(self.layers, self.froms)):
            if i == self._deferred:
                pass  # replicated after the pointwise tail instead (see __init__)
            elif isinstance(frm, list):
                x = layer([outs[f] for f in frm])
            else:
                x = layer(x if not outs else outs[frm])
            outs.append(x)
        if self._deferred is not None:
            scale = int(self.layers[self._deferred].scale_factor)
            if lowres and [x.shape[2] * scale, x.shape[3] * scale] == self.img_size:
                return x, scale
            x = self.layers[self._deferred](x)
        if lowres:
            return (x, 1) if list(x.shape[2:]) == self.img_size else \
                (F.interpolate(x, size=self.img_size, mode="bilinear", align_corners=False), 1)
        if list(x.shape[2:]) != self.img_size:
            x = F.interpolate(x, size=self.img_size, mode="bilinear", align_corners=False)
        return x

    def dcn_sites(self):
        """[(name, DCNv3 module)] — the hot-path call sites of this model."""
        from .ops_dcnv3.modules import DCNv3
        return [(n, m) for n, m in self.named_modules() if isinstance(m, DCNv3)]


# ---------------------------------------------------------------------------------------------
# loss: CE(class weights) + 0.5 * weighted Dice on softmax(pred)   (seg_diceloss_yolov5.py:693-750)
# ---------------------------------------------------------------------------------------------
class SegmentationLoss(nn.Module):
    """`fused`: None = the fused CUDA kernels (yolo_dual_b200/csrc/segloss_b200.cu) whenever they apply (CUDA tensors,
    label_smoothing 0, at most 16 classes), True = require them, False = the reference's chain of torch ops.
    `scale`: `pred` is the model's output before its last nearest Upsample (`SegModel(x, lowres=True)`)."""
    accepts_lowres = True

    def __init__(self, num_classes: int = 12, label_smoothing: float = 0.0, class_weights=None,
                 fused: Optional[bool] = None):
        super().__init__()
        self.num_classes = num_classes
        w = torch.ones(num_classes) if class_weights is None else torch.as_tensor(class_weights, dtype=torch.float32)
        self.register_buffer("class_weights", w.float())
        self.label_smoothing = label_smoothing
        self.fused = fused

    def _use_fused(self, pred: torch.Tensor) -> bool:
        from ._segloss import MAX_CLASSES
        ok = pred.is_cuda and self.label_smoothing == 0.0 and pred.size(1) <= MAX_CLASSES
        if self.fused and not ok:
            raise RuntimeError("fused=True needs CUDA tensors, label_smoothing 0 and at most 16 classes")
        return ok and self.fused is not False

    def forward(self, pred: torch.Tensor, target: torch.Tensor, scale: int = 1):
        if pred.size(0) != target.size(0):
            raise ValueError(f"batch mismatch: {pred.size(0)} vs {target.size(0)}")
        full = (pred.shape[2] * scale, pred.shape[3] * scale)
        if self._use_fused(pred) and tuple(target.shape[1:]) == full:
            from ._segloss import FusedSegLoss
            total, ce, dice = FusedSegLoss.apply(pred.float(), target, self.class_weights, int(scale))
            return total, (total.detach(), ce, dice)
        if scale != 1:
            pred = F.interpolate(pred, scale_factor=float(scale), mode="nearest")
        if pred.shape[2:] != target.shape[1:]:
            target = F.interpolate(target.unsqueeze(1).float(), size=pred.shape[2:], mode="nearest").squeeze(1).long()
        pred = pred.float()
        ce = F.cross_entropy(pred, target, weight=self.class_weights, label_smoothing=self.label_smoothing)
        one_hot = torch.zeros_like(pred).scatter_(1, target.unsqueeze(1), 1.0)
        dice = self._dice(pred.softmax(1), one_hot)
        total = ce + 0.5 * dice
        return total, (total.detach(), ce.detach(), dice.detach())

    def _dice(self, prob, one_hot, eps: float = 1e-6):
        wp = prob * self.class_weights.view(1, -1, 1, 1)
        inter = (wp * one_hot).sum(dim=(2, 3))
        dice = (2.0 * inter + eps) / (wp.sum(dim=(2, 3)) + one_hot.sum(dim=(2, 3)) + eps)
        return 1.0 - dice.mean()


# ---------------------------------------------------------------------------------------------
# data-parallel step (batch shard; gradient all-reduce by DDP over NCCL / gloo)
# ---------------------------------------------------------------------------------------------
def smart_optimizer(model: nn.Module, lr=0.01, momentum=0.937, decay=5e-4):
    """SGD-nesterov with three parameter groups: weights (decay), norm weights, biases
    (utils/torch_utils.py:318-346; hyp defaults seg_diceloss_yolov5.py:851-862)."""
    g_w, g_bn, g_b = [], [], []
    norm = tuple(v for k, v in nn.__dict__.items() if "Norm" in k and isinstance(v, type))
    for m in model.modules():
        for pn, p in m.named_parameters(recurse=False):
            if not p.requires_grad:
                continue
            if pn == "bias":
                g_b.append(p)
            elif pn == "weight" and isinstance(m, norm):
                g_bn.append(p)
            else:
                g_w.append(p)
    opt = torch.optim.SGD(g_b, lr=lr, momentum=momentum, nesterov=True)
    opt.add_param_group({"params": g_w, "weight_decay": decay})
    opt.add_param_group({"params": g_bn, "weight_decay": 0.0})
    return opt


def shard_batch(n_global: int, rank: int, world: int):
    """Contiguous images of the global batch owned by `rank` (reference: batch_size // WORLD_SIZE,
    seg_diceloss_yolov5.py:1001).  Returns a slice; the shards tile [0, n_global) exactly."""
    if n_global % world:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world}")
    per = n_global // world
    return slice(rank * per, (rank + 1) * per)


def wrap_ddp(model: nn.Module, device=None, bucket_cap_mb: Optional[float] = None, grad_compress: Optional[str] = None,
             first_bucket_mb: Optional[float] = None, static_graph: bool = True):
    """smart_DDP (utils/torch_utils.py:55-63): DDP with static_graph when a process group is up.

    Additions (all off by default = the reference's plain DDP): `bucket_cap_mb` sizes the gradient buckets (DDP's 25 MB
    default puts this model's 21 M parameters into three ring all-reduces, the last of which cannot overlap with the
    backward), `first_bucket_mb` the first one (the gradients of the LAST layers: it fires while most of the
    backward is still to run), `grad_compress='bf16'` all-reduces the buckets in bfloat16 (torch's
    bf16_compress_hook: half the bytes; the sum of `world` bf16 values per element rounds once more than fp32)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    from torch.nn.parallel import DistributedDataParallel as DDP
    ids = [device.index] if device is not None and device.type == "cuda" else None
    kw = {}
    if bucket_cap_mb:
        kw["bucket_cap_mb"] = bucket_cap_mb
    if first_bucket_mb:
        dist._DEFAULT_FIRST_BUCKET_BYTES = int(first_bucket_mb * 1024 * 1024)  # read by DDP's constructor
    # static_graph (smart_DDP's choice) cannot be combined with a first iteration inside no_sync() (reducer.cpp asserts
    # expect_autograd_hooks_): callers that accumulate gradients pass static_graph=False (Trainer does)
    # (the layer tables compute layers whose outputs nobody reads: without static_graph DDP has to look for unused parameters)
    ddp = DDP(model, device_ids=ids, static_graph=static_graph, find_unused_parameters=not static_graph,
              gradient_as_bucket_view=True, **kw)  # grads live in the buckets
    if grad_compress == "bf16":
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        ddp.register_comm_hook(None, default_hooks.bf16_compress_hook)
    elif grad_compress not in (None, "none"):
        raise ValueError("grad_compress must be None or 'bf16'")
    return ddp


def forward_loss(model, criterion, imgs, labels, autocast_dtype=None):
    """model forward (optional autocast) + criterion.  When both sides can, the model stops before its deferred
    nearest Upsample and the loss reads the low-resolution map against the full-resolution labels."""
    lowres = getattr(criterion, "accepts_lowres", False) and \
        getattr(getattr(model, "module", model), "_deferred", None) is not None
    with torch.autocast(imgs.device.type, dtype=autocast_dtype, enabled=autocast_dtype is not None):
        pred = model(imgs, lowres=True) if lowres else model(imgs)
    return criterion(pred[0], labels, pred[1]) if lowres else criterion(pred, labels)


def train_step(model, criterion, optimizer, imgs, labels, autocast_dtype=None):
    """One optimizer step of the reference hot loop (seg_diceloss_yolov5.py:1073-1103) without
    logging: forward (optional autocast) -> CE+Dice -> backward (DDP all-reduces) -> SGD step."""
    loss, parts = forward_loss(model, criterion, imgs, labels, autocast_dtype)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return loss.detach(), parts


class GraphedTrainStep:
    """`train_step` as ONE CUDA graph: forward, CE + Dice, backward (DDP's bucketed NCCL all-reduces included) and the
    SGD update are captured once and replayed per step.  The eager step of these models is 850-1250 kernel launches;
    at 8 images per GPU (BASELINE configs[3] on 8 GPUs) the host cannot issue them as fast as the GPU retires them.
    Every kernel of this package is capturable (explicit stream, no host synchronisation, caller-visible workspaces).

    Rules kept from torch's CUDA-graph notes: `warmup` eager steps on a side stream first (cuDNN autotuning, DDP's
    first static-graph iteration, lazy workspaces; >= 11 under DDP), gradients set to None before the capture so that
    they live in the graph's private pool, a fixed learning rate while the graph is in use (re-capture after a
    scheduler step).  The warm-up steps are real optimizer steps on the batches `batches` yields."""

    def __init__(self, model, criterion, optimizer, batches, autocast_dtype=None, warmup: Optional[int] = None):
        ddp = isinstance(model, nn.parallel.DistributedDataParallel)
        warmup = (11 if ddp else 3) if warmup is None else warmup
        it = iter(batches)
        imgs, labels = next(it)
        self.imgs, self.labels = torch.empty_like(imgs), torch.empty_like(labels)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for k in range(warmup):
                if k:
                    imgs, labels = next(it)
                self.imgs.copy_(imgs)
                self.labels.copy_(labels)
                train_step(model, criterion, optimizer, self.imgs, self.labels, autocast_dtype)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss, parts = forward_loss(model, criterion, self.imgs, self.labels, autocast_dtype)
            loss.backward()
            optimizer.step()
        self.loss, self.parts = loss.detach(), parts

    def __call__(self, imgs, labels):
        self.imgs.copy_(imgs, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        return self.loss, self.parts
