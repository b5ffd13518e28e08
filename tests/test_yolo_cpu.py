"""Detection-model builder (yolo_dual_b200/yolo.py) against the facts the reference's parse_model / DetectionModel
produce for models/backbone/yolov5n-DCN.yaml (models/yolo.py:165-262, 296-390).  CPU only: the DCNv3 variants are built
(never run) here; `dcn="none"` is the stock yolov5n and runs."""
import math

import pytest
import torch
from torch import nn

from yolo_dual_b200.blocks import C3_DCNV3
from yolo_dual_b200.yolo import (C3, YOLOV5N_DCNV3, Concat, Detect, DetectionModel, build_layers, make_divisible)


def test_make_divisible():
    assert [make_divisible(x, 8) for x in (16, 17, 64 * 0.25, 1024 * 0.25, 1)] == [16, 24, 16, 256, 8]


def test_stock_yolov5n_structure_and_parameter_count():
    m = DetectionModel(YOLOV5N_DCNV3, dcn="none")
    # the published yolov5n summary: 1 872 157 parameters at nc = 80 (before conv+bn fusion)
    assert sum(p.numel() for p in m.parameters()) == 1872157
    assert m.save == [4, 6, 10, 14, 17, 20, 23]                       # parse_model's savelist for this table
    assert m.stride.tolist() == [8.0, 16.0, 32.0]                     # without any forward pass
    assert [m.channels[i] for i in (0, 1, 4, 6, 8, 9, 17, 20, 23)] == [16, 32, 64, 128, 256, 256, 64, 128, 256]
    assert [len(m.model[i].m) for i in (2, 4, 6, 8)] == [1, 2, 3, 1]   # depth_multiple 0.33 on 3, 6, 9, 3
    head = m.model[-1]
    assert isinstance(head, Detect) and head.nl == 3 and head.na == 3 and head.no == 85
    torch.testing.assert_close(head.anchors[0], torch.tensor([[10, 13], [16, 30], [33, 23]]) / 8.0)
    torch.testing.assert_close(head.anchors[2], torch.tensor([[116, 90], [156, 198], [373, 326]]) / 32.0)
    assert all(bn.eps == 1e-3 and bn.momentum == 0.03 for bn in m.modules() if isinstance(bn, nn.BatchNorm2d))
    assert list(m.state_dict())[0] == "model.0.conv.weight" and "model.24.anchors" in m.state_dict()


def test_detect_bias_priors():
    torch.manual_seed(0)
    m = DetectionModel(YOLOV5N_DCNV3, dcn="none")
    torch.manual_seed(0)
    layers, *_ = build_layers(YOLOV5N_DCNV3, [3], dcn="none")         # same initialisation, no priors applied
    for conv, raw, s in zip(m.model[-1].m, layers[-1].m, (8, 16, 32)):
        d = (conv.bias - raw.bias).view(3, -1)
        torch.testing.assert_close(d[:, 4], torch.full((3,), math.log(8 / (640 / s) ** 2)))
        torch.testing.assert_close(d[:, 5:], torch.full((3, 80), math.log(0.6 / (80 - 0.99999))))
        torch.testing.assert_close(d[:, :4], torch.zeros(3, 4))


def test_forward_shapes_and_box_decoding():
    torch.manual_seed(0)
    m = DetectionModel(YOLOV5N_DCNV3, dcn="none", nc=3)
    x = torch.randn(2, 3, 64, 96)
    m.train()
    raw = m(x)
    assert [tuple(r.shape) for r in raw] == [(2, 3, 8, 12, 8), (2, 3, 4, 6, 8), (2, 3, 2, 3, 8)]
    m.eval()
    with torch.no_grad():
        y, raw = m(x)
    assert y.shape == (2, 3 * (8 * 12 + 4 * 6 + 2 * 3), 8)
    head = m.model[-1]
    # the decoding formula of yolo.py:77-80 on level 1, cell (y=2, x=5), anchor 1
    r = raw[1][0, 1, 2, 5].sigmoid()
    want_xy = (r[:2] * 2 - 0.5 + torch.tensor([5.0, 2.0])) * 16
    want_wh = (r[2:4] * 2) ** 2 * head.anchors[1, 1] * 16
    got = y[0, 3 * 8 * 12 + 1 * 4 * 6 + 2 * 6 + 5]
    torch.testing.assert_close(got[:2], want_xy)
    torch.testing.assert_close(got[2:4], want_wh)
    torch.testing.assert_close(got[4:], r[4:])


def test_dcnv3_slots_and_aliases():
    m = DetectionModel(YOLOV5N_DCNV3, dcn="dcnv3")
    assert all(isinstance(m.model[i], C3_DCNV3) for i in (4, 6, 8)) and isinstance(m.model[2], C3)
    sites = m.dcn_sites()
    assert [n for n, _ in sites] == ["model.4.m.0.cv2.dcnv3", "model.4.m.1.cv2.dcnv3", "model.6.m.0.cv2.dcnv3",
                                    "model.6.m.1.cv2.dcnv3", "model.6.m.2.cv2.dcnv3", "model.8.m.0.cv2.dcnv3"]
    assert [(s.channels, s.group, s.group_channels) for _, s in sites] == \
        [(32, 2, 16)] * 2 + [(64, 4, 16)] * 3 + [(128, 8, 16)]
    # the reference yaml spells the slots C3_DCN: read as the DCNv3 block
    cfg = dict(YOLOV5N_DCNV3, backbone=[[f, n, "C3_DCN" if t == "C3_DCNV3" else t, a]
                                        for f, n, t, a in YOLOV5N_DCNV3["backbone"]])
    assert list(DetectionModel(cfg).state_dict()) == list(m.state_dict())
    with pytest.raises(NotImplementedError):
        build_layers(dict(YOLOV5N_DCNV3, head=[[-1, 1, "Focus", [64]]]), [3])
    with pytest.raises(ValueError):
        build_layers(dict(YOLOV5N_DCNV3, head=[[[-1, 4], 1, "Concat", [1]]]), [3], dcn="none")   # strides 32 vs 8
    assert isinstance(Concat(1)([torch.zeros(1, 2, 3, 3), torch.zeros(1, 5, 3, 3)]), torch.Tensor)


def test_repeated_plain_modules_are_independent_instances():
    cfg = dict(YOLOV5N_DCNV3, depth_multiple=1.0, backbone=[[-1, 2, "Conv", [64, 3, 1]]], head=[])
    layers, save, chans, strides = build_layers(cfg, [3], dcn="none")
    assert isinstance(layers[0], nn.Sequential) and len(layers[0]) == 2
    assert layers[0][0].conv.weight.data_ptr() != layers[0][1].conv.weight.data_ptr()
    assert not torch.equal(layers[0][0].conv.weight[:, :3], layers[0][1].conv.weight[:, :3])
