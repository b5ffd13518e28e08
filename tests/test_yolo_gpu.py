"""yolov5n-DCNv3 detection model (yolo_dual_b200/yolo.py) on the GPU: one training forward/backward and the
inference decode, through the DCNv3 kernels at the model's three sites (C = 32 / 64 / 128, group_channels 16)."""
import pytest
import torch

from yolo_dual_b200.yolo import YOLOV5N_DCNV3, DetectionModel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("autocast", [None, torch.bfloat16])
def test_detection_model_trains_and_decodes(autocast):
    torch.manual_seed(0)
    m = DetectionModel(YOLOV5N_DCNV3, nc=4).to(DEV).to(memory_format=torch.channels_last).train()
    assert m.model[-1].stride.device.type == "cuda"
    x = torch.randn(2, 3, 128, 160, device=DEV).contiguous(memory_format=torch.channels_last)
    with torch.autocast("cuda", dtype=autocast, enabled=autocast is not None):
        raw = m(x)
    assert [tuple(r.shape) for r in raw] == [(2, 3, 16, 20, 9), (2, 3, 8, 10, 9), (2, 3, 4, 5, 9)]
    sum(r.float().square().mean() for r in raw).backward()
    for name, site in m.dcn_sites():
        for p in (site.input_proj.weight, site.output_proj.weight, site.offset.weight, site.mask.weight):
            assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0, name
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    m.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=autocast, enabled=autocast is not None):
        y, _ = m(x)
    assert y.shape == (2, 3 * (16 * 20 + 8 * 10 + 4 * 5), 9) and torch.isfinite(y.float()).all()
    assert float(y[..., 4:].min()) >= 0.0 and float(y[..., 4:].max()) <= 1.0
