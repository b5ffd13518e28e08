"""NHWC resize kernels (yolo_dual_b200/csrc/resize_b200.cu through its C-ABI) against F.interpolate on the same
inputs: nearest is bit-exact forward, bilinear within float32 rounding (1e-6) / one 16-bit ulp; backward within
summation-order tolerance (the gather adds up to 64 terms in float32, ATen accumulates atomically)."""
import pytest
import torch
import torch.nn.functional as F

from yolo_dual_b200 import _resize
from yolo_dual_b200.seg import Concat, Upsample

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref(x, size, mode):
    return F.interpolate(x, size=size, mode=mode, **({} if mode == "nearest" else {"align_corners": False}))


@pytest.mark.parametrize("shape,size,mode", [
    ((2, 64, 10, 12), (20, 24), "nearest"), ((1, 8, 5, 7), (20, 28), "nearest"), ((2, 32, 6, 6), (6, 6), "nearest"),
    ((2, 64, 10, 10), (40, 40), "bilinear"), ((2, 16, 9, 13), (18, 26), "bilinear"), ((1, 8, 32, 32), (4, 4), "bilinear"),
    ((2, 24, 7, 5), (10, 17), "bilinear"), ((1, 8, 1, 1), (3, 5), "bilinear"), ((2, 8, 12, 9), (5, 20), "bilinear"),
    ((16, 512, 40, 40), (160, 160), "bilinear"), ((16, 128, 160, 160), (320, 320), "nearest"),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_resize_matches_interpolate(shape, size, mode, dtype):
    g = torch.Generator().manual_seed(sum(shape) + size[0])
    x = torch.randn(shape, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last)
    assert _resize.usable(x, size, mode)
    go = torch.randn(shape[:2] + size, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last)
    xa = x.detach().float().clone().requires_grad_(True)
    ya = _ref(xa, size, mode)
    ya.backward(go.float())
    xb = x.detach().clone().requires_grad_(True)
    yb = _resize.resize(xb, size, mode)
    assert yb.dtype == dtype and yb.shape == ya.shape and yb.is_contiguous(memory_format=torch.channels_last)
    yb.backward(go)
    if mode == "nearest":
        assert torch.equal(yb.float(), ya)
    elif dtype == torch.float32:
        torch.testing.assert_close(yb, ya, rtol=1e-5, atol=1e-6)
    else:
        torch.testing.assert_close(yb.float(), ya, rtol=1e-2, atol=1e-2)
    if dtype == torch.float32:
        torch.testing.assert_close(xb.grad, xa.grad, rtol=1e-5, atol=1e-5)
    else:
        torch.testing.assert_close(xb.grad.float(), xa.grad, rtol=1e-2, atol=1e-2 * float(xa.grad.abs().max()))


def test_model_layers_use_the_kernels_and_keep_dtype_under_autocast():
    x = torch.randn(2, 16, 8, 8, device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        up = Upsample(scale_factor=2.0, mode="nearest")(x)
        cat = Concat(1)([up, x])
        nchw = Upsample(scale_factor=2.0, mode="nearest")(x.contiguous())       # not NHWC: torch, own dtype
    assert up.dtype == torch.bfloat16 and torch.equal(up, F.interpolate(x, scale_factor=2.0))
    assert cat.dtype == torch.bfloat16 and cat.shape == (2, 32, 16, 16)
    torch.testing.assert_close(cat[:, 16:].float(), F.interpolate(x.float(), size=(16, 16), mode="bilinear"),
                               rtol=1e-2, atol=1e-2)
    assert nchw.dtype == torch.bfloat16 and torch.equal(nchw, up)
    assert not _resize.usable(x, (12, 12), "nearest") and not _resize.usable(x[:, :12], (16, 16), "nearest")
