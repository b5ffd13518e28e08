"""BASELINE configs[4] at model level: C3-DCN seg inference in fp16 with the fused mask-softmax variant
(`DCNv3(fused_softmax=True)`: the mask *logits* go straight to the kernels, LIB/modules/dcnv3.py:122-123 is what gets
fused) against the unfused model with the same weights; and the channels-last model runs the DCNv3 sites without a
layout copy (`common and yolo.py:11-13` permutes + LIB/modules/dcnv3.py:119-120)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _pair(cfg, size, dtype):
    from yolo_dual_b200 import seg
    torch.manual_seed(0)
    a = seg.SegModel(cfg, dcn="dcnv3", fused_softmax=False, img_size=size).to(DEV)
    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():  # non-trivial sampling: fresh heads are zero (regular grid, uniform weights)
        for _, m in a.dcn_sites():
            m.offset.weight.copy_(torch.randn(m.offset.weight.shape, generator=gen) * 0.1)
            m.offset.bias.copy_(torch.randn(m.offset.bias.shape, generator=gen) * 0.5)
            m.mask.weight.copy_(torch.randn(m.mask.weight.shape, generator=gen) * 0.1)
            m.mask.bias.copy_(torch.randn(m.mask.bias.shape, generator=gen))
    b = seg.SegModel(cfg, dcn="dcnv3", fused_softmax=True, img_size=size).to(DEV)
    b.load_state_dict(a.state_dict())  # same parameter names whatever the softmax variant
    cast = lambda m: m.to(dtype).eval().to(memory_format=torch.channels_last)
    return cast(a), cast(b)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32], ids=["f16", "f32"])
@pytest.mark.parametrize("name", ["yolov5seg", "yolov8seg"])
def test_fused_softmax_model_matches_unfused(name, dtype):
    from yolo_dual_b200 import seg
    cfg = {"yolov5seg": seg.YOLOV5_SEG, "yolov8seg": seg.YOLOV8_SEG}[name]
    a, b = _pair(cfg, (256, 256), dtype)
    x = torch.randn(2, 3, 256, 256, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1)).to(dtype)
    x = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        ya, yb = a(x), b(x)
    assert ya.shape == yb.shape == (2, 12, 256, 256)
    # class probabilities: fp32 differs only by summation order inside the softmax; fp16 by the rounding of the
    # probabilities the unfused model writes to memory (north_star: fp16 rtol 1e-2)
    # (fp32: the two softmaxes differ in the last bit and ~30 layers of TF32 convolutions follow: 2.6e-5 measured)
    tol = dict(rtol=1e-2, atol=2e-3) if dtype == torch.float16 else dict(rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(yb.float(), ya.float(), **tol)
    assert float((ya.argmax(1) == yb.argmax(1)).float().mean()) > 0.999


def test_channels_last_model_feeds_dcnv3_without_layout_copies():
    """In a channels-last model the NCHW -> NHWC permute in front of DCNv3 is a view of the same memory and the module's
    output goes back as a view: DCNV3_YoLo must not copy on either side."""
    from yolo_dual_b200.blocks import DCNV3_YoLo
    blk = DCNV3_YoLo(64, 64, k=3, dcn_group="gc16").to(DEV).half().eval().to(memory_format=torch.channels_last)
    seen = {}
    blk.dcnv3.register_forward_hook(lambda m, inp, out: seen.update(inp=inp[0], out=out))
    x = torch.randn(2, 64, 16, 16, device=DEV).half().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y = blk(x)
    assert seen["inp"].is_contiguous() and seen["inp"].shape == (2, 16, 16, 64)      # NHWC, dense: no copy was needed
    assert y.data_ptr() == seen["out"].data_ptr() and y.is_contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32], ids=["f16", "f32"])
def test_cat_free_inference_blocks_match_the_cat_path(dtype):
    """Under no_grad the plain C3 blocks let their Conv branches write straight into the concatenated buffer
    (bnact_b200_eval_pitched); with autograd on they take torch.cat.  Same numbers.  (C3_DCNV3 keeps torch.cat: its DCNv3
    branch would need a strided copy that is slower than cat's own kernel.)"""
    from yolo_dual_b200 import seg
    from yolo_dual_b200.blocks import C3_DCNV3
    torch.manual_seed(0)
    for blk in (seg.C3(64, 64, n=2), seg.C3(64, 128, n=0), C3_DCNV3(64, 64, n=1, dcn_group="gc16")):
        blk = blk.to(DEV).to(dtype).eval().to(memory_format=torch.channels_last)
        for m in blk.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.5)
                m.running_var.uniform_(0.5, 1.5)
        x = torch.randn(2, 64, 24, 20, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            a = blk(x)                                   # cat-free path
        b = blk(x.clone().requires_grad_(True)).detach()  # autograd on: torch.cat path
        tol = dict(rtol=1e-2, atol=1e-2) if dtype == torch.float16 else dict(rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(a.float(), b.float(), **tol)
