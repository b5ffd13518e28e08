"""CPU tests of the training-loop semantics restated from the reference's seg trainers
(yolo_dual_b200/trainer.py).  DCN slots are plain convs here (no CPU path for DCNv3)."""
import math
import warnings

import pytest
import torch

from yolo_dual_b200.seg import CAMVID_CLASS_WEIGHTS, YOLOV5_SEG, SegModel, SegmentationLoss
from yolo_dual_b200.trainer import HYP, ModelEMA, Trainer, lr_lambda


def tiny(dcn="none"):
    torch.manual_seed(0)
    return SegModel(YOLOV5_SEG, dcn=dcn, img_size=(32, 32))


def batch(n, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, 32, 32, generator=g), torch.randint(0, 12, (n, 32, 32), generator=g)


def test_lr_schedules_match_reference_formulas():
    lin, cos = lr_lambda(100, 0.2), lr_lambda(100, 0.2, cos=True)
    assert lin(0) == 1.0 and abs(lin(100) - 0.2) < 1e-12 and abs(lin(50) - 0.6) < 1e-12
    assert cos(0) == 1.0 and abs(cos(100) - 0.2) < 1e-12 and abs(cos(50) - 0.6) < 1e-12
    assert abs(cos(25) - (((1 - math.cos(math.pi / 4)) / 2) * (0.2 - 1) + 1)) < 1e-12


def test_ema_update_rule():
    m = torch.nn.Linear(3, 2)
    ema = ModelEMA(m, decay=0.9999, tau=2000)
    w0 = ema.ema.weight.clone()
    with torch.no_grad():
        m.weight.add_(1.0)
    ema.update(m)
    d = 0.9999 * (1 - math.exp(-1 / 2000))
    torch.testing.assert_close(ema.ema.weight, w0 * d + (1 - d) * m.weight.detach())
    assert ema.updates == 1 and not any(p.requires_grad for p in ema.ema.parameters())


@pytest.mark.parametrize("bs,acc", [(16, 4), (64, 1), (8, 8), (128, 1), (24, 3)])
def test_nominal_batch_accumulation(bs, acc):
    t = Trainer(tiny(), SegmentationLoss(12), batch_size=bs, epochs=10, ema=False)
    assert t.accumulate == acc
    wd = HYP["weight_decay"] * bs * acc / 64
    assert abs(t.optimizer.param_groups[1]["weight_decay"] - wd) < 1e-12
    assert t.optimizer.param_groups[0].get("weight_decay", 0) == 0  # biases never decay


def test_optimizer_steps_every_accumulate_micro_steps():
    t = Trainer(tiny(), SegmentationLoss(12, class_weights=CAMVID_CLASS_WEIGHTS), batch_size=32, epochs=10)
    assert t.accumulate == 2
    x, y = batch(2)
    w0 = [p.detach().clone() for p in t.raw_model.parameters()]
    _, parts, stepped = t.micro_step(x, y)
    assert not stepped and all(torch.equal(a, b.detach()) for a, b in zip(w0, t.raw_model.parameters()))
    g1 = [p.grad.clone() for p in t.raw_model.parameters() if p.grad is not None]
    _, _, stepped = t.micro_step(x, y)
    assert stepped and t.ema.updates == 1
    assert any(not torch.equal(a, b.detach()) for a, b in zip(w0, t.raw_model.parameters()))
    assert all(p.grad is None for p in t.raw_model.parameters())  # zero_grad after the step
    assert len(parts) == 3 and len(g1) > 0
    # the epoch's last batch always steps (seg_diceloss_yolov5.py:1095)
    _, _, stepped = t.micro_step(x, y, last_of_epoch=True)
    assert stepped


def test_accumulated_gradients_are_the_sum_of_micro_batch_gradients():
    crit = SegmentationLoss(12, class_weights=CAMVID_CLASS_WEIGHTS)
    t = Trainer(tiny().eval(), crit, batch_size=32, epochs=10, ema=False)  # eval: BN statistics fixed
    t.model.eval()
    (xa, ya), (xb, yb) = batch(2, 1), batch(2, 2)
    t.micro_step(xa, ya)
    # second half done by hand so the summed grads can be read before the optimizer clears them
    pred = t.model(xb)
    loss, _ = crit(pred, yb)
    loss.backward()
    got = {k: p.grad.clone() for k, p in t.raw_model.named_parameters() if p.grad is not None}
    ref = tiny().eval()
    la, _ = crit(ref(xa), ya)
    lb, _ = crit(ref(xb), yb)
    (la + lb).backward()
    for k, p in ref.named_parameters():
        if p.grad is not None:
            torch.testing.assert_close(got[k], p.grad, rtol=1e-5, atol=1e-7)


def test_checkpoint_layout_and_resume(tmp_path):
    crit = SegmentationLoss(12)
    t = Trainer(tiny(), crit, batch_size=64, epochs=10)
    x, y = batch(2)
    t.micro_step(x, y)
    t.end_epoch(fitness=0.4)
    t.micro_step(x, y)
    t.end_epoch(fitness=0.3)
    last, best = tmp_path / "last.pt", tmp_path / "best.pt"
    t.save(last, best, is_best=True)
    ck = torch.load(last, weights_only=False)
    assert set(ck) == {"model", "optimizer", "epoch", "best_fitness"}          # reference layout
    assert isinstance(ck["model"], SegModel) and ck["epoch"] == 1 and ck["best_fitness"] == 0.4  # index of the epoch just finished
    assert set(torch.load(best, weights_only=False)) == {"model"}
    t2 = Trainer(tiny(), crit, batch_size=64, epochs=10)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t2.resume(ck)
    assert t2.epoch == 2 and t2.best_fitness == 0.4   # two epochs trained -> continue with epoch index 2
    for (k, a), (_, b) in zip(ck["model"].state_dict().items(), t2.raw_model.state_dict().items()):
        assert torch.equal(a.float(), b.float()), k
    assert abs(t2.optimizer.param_groups[0]["lr"] - HYP["lr0"] * lr_lambda(10)(2)) < 1e-9
    assert t2.optimizer.state_dict()["state"].keys() == ck["optimizer"]["state"].keys()


def test_pickled_checkpoint_carries_the_dcnv3_class_paths(tmp_path):
    """The reference pickles whole modules, so class names/paths are part of the on-disk format
    (SURVEY §5): a C3-DCNV3 model must round-trip through torch.save / torch.load."""
    m = tiny(dcn="dcnv3")
    torch.save({"model": m}, tmp_path / "m.pt")
    back = torch.load(tmp_path / "m.pt", weights_only=False)["model"]
    from yolo_dual_b200.ops_dcnv3.modules import DCNv3
    sites = [x for x in back.modules() if isinstance(x, DCNv3)]
    assert len(sites) == 3 and type(sites[0]).__module__ == "yolo_dual_b200.ops_dcnv3.modules.dcnv3"
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), back.state_dict().values()))


def test_sync_bn_switch_converts_every_batchnorm():
    t = Trainer(tiny(dcn="dcnv3"), SegmentationLoss(12), batch_size=16, epochs=1, ema=False, sync_bn=True)
    kinds = {type(m).__name__ for m in t.raw_model.modules() if "BatchNorm" in type(m).__name__}
    assert kinds == {"SyncBatchNorm"}  # including the BN inside every DCNv3.dw_conv


def test_resume_warns_when_shapes_do_not_fit():
    """A checkpoint built with another DCNv3 group count must not lose its heads silently."""
    crit = SegmentationLoss(12)
    src = SegModel(YOLOV5_SEG, dcn="dcnv3", dcn_group=None, img_size=(32, 32))    # the reference's group = g = 1
    dst = Trainer(SegModel(YOLOV5_SEG, dcn="dcnv3", img_size=(32, 32)), crit, batch_size=64, epochs=2, ema=False)
    with pytest.warns(RuntimeWarning, match="NOT loaded"):
        dst.resume({"model": src, "optimizer": None, "epoch": 0})
    assert dst.dropped_keys and all(".offset." in k or ".mask." in k for k in dst.dropped_keys)


def test_default_block_has_the_reference_parameter_shapes():
    """common and yolo.py:9 passes group=g (default 1): offset head [18, C], mask head [9, C]."""
    from yolo_dual_b200.blocks import C3_DCNV3
    blk = C3_DCNV3(64, 64)
    d = blk.m[0].cv2.dcnv3
    assert d.group == 1 and tuple(d.offset.weight.shape) == (18, 32) and tuple(d.mask.weight.shape) == (9, 32)
    d16 = C3_DCNV3(64, 64, dcn_group="gc16").m[0].cv2.dcnv3
    assert d16.group == 2 and d16.group_channels == 16
