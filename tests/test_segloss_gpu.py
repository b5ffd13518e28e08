"""Fused CE + weighted-Dice loss (yolo_dual_b200/csrc/segloss_b200.cu, through its C-ABI) against the reference's
chain of torch ops (SegmentationLoss(fused=False), restating seg_diceloss_yolov5.py:712-750) on the same inputs.
fp32 bar: rtol 1e-5 on the loss values; gradients rtol 1e-4 / atol 1e-6 of the largest gradient (the unfused
chain itself sums 6.5 M fp32 terms per image in an unspecified order)."""
import pytest
import torch
import torch.nn.functional as F

from yolo_dual_b200.seg import CAMVID_CLASS_WEIGHTS, YOLOV5_SEG, SegModel, SegmentationLoss, forward_loss

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(n, c, h, w, scale, seed, weights):
    g = torch.Generator().manual_seed(seed)
    pred = torch.randn(n, c, h, w, generator=g).mul_(2.0).to(DEV)
    target = torch.randint(0, c, (n, h * scale, w * scale), generator=g).to(DEV)
    cw = torch.rand(c, generator=g).mul_(20.0).add_(0.5) if weights else None
    return pred, target, cw


@pytest.mark.parametrize("n,c,h,w,scale,weights", [
    (2, 12, 64, 64, 1, True), (3, 12, 37, 53, 1, True), (1, 5, 16, 16, 1, False), (2, 16, 40, 24, 1, True),
    (2, 12, 40, 40, 4, True), (2, 12, 33, 17, 2, True), (1, 1, 8, 8, 1, False), (16, 12, 160, 160, 4, True),
])
def test_fused_loss_matches_unfused(n, c, h, w, scale, weights):
    pred, target, cw = _case(n, c, h, w, scale, 7 * n + c + h, weights)
    ref = SegmentationLoss(c, class_weights=cw, fused=False).to(DEV)
    fus = SegmentationLoss(c, class_weights=cw, fused=True).to(DEV)
    pa = pred.clone().requires_grad_(True)
    pb = pred.clone().requires_grad_(True)
    la, (ta, cea, da) = ref(pa, target, scale)
    lb, (tb, ceb, db) = fus(pb, target, scale)
    for x, y in ((la, lb), (cea, ceb), (da, db)):
        torch.testing.assert_close(y.float(), x.float(), rtol=1e-5, atol=1e-6)
    up = torch.tensor(1.7, device=DEV)
    (la * up).backward()
    (lb * up).backward()
    torch.testing.assert_close(pb.grad, pa.grad, rtol=1e-4, atol=1e-6 * float(pa.grad.abs().max()))


def test_fused_loss_is_softmax_output_invariant_and_ignores_bad_labels():
    pred, target, cw = _case(2, 12, 32, 32, 1, 3, True)
    fus = SegmentationLoss(12, class_weights=cw, fused=True).to(DEV)
    l0, _ = fus(pred, target)
    l1, _ = fus(pred + 3.0, target)            # softmax is shift invariant
    torch.testing.assert_close(l0, l1, rtol=1e-5, atol=1e-6)
    # a uniform prediction on uniform labels: CE = log C, dice from the closed form
    flat = torch.zeros(1, 4, 8, 8, device=DEV)
    lab = torch.zeros(1, 8, 8, dtype=torch.long, device=DEV)
    _, (_, ce, dice) = SegmentationLoss(4, fused=True).to(DEV)(flat, lab)
    assert abs(float(ce) - 1.3862943611) < 1e-5
    want = 1.0 - ((2 * 16 + 1e-6) / (16 + 64 + 1e-6) + 3 * (1e-6 / (16 + 1e-6))) / 4
    assert abs(float(dice) - want) < 1e-5


def test_fused_loss_out_of_range_labels_poison_the_loss():
    """Labels outside [0, C) make the reference's cross_entropy / scatter_ raise; the fused path (no host sync)
    returns NaN instead of silently dropping those pixels."""
    pred, target, cw = _case(2, 12, 32, 32, 1, 3, True)
    fus = SegmentationLoss(12, class_weights=cw, fused=True).to(DEV)
    assert torch.isfinite(fus(pred, target)[0])
    bad = target.clone()
    bad[0, 3, 5] = 255
    assert torch.isnan(fus(pred, bad)[0])
    bad = target.clone()
    bad[1, 0, 0] = -1
    assert torch.isnan(fus(pred, bad)[0])


def test_fused_loss_argument_errors():
    fus = SegmentationLoss(12, fused=True)
    with pytest.raises(RuntimeError):
        fus(torch.zeros(1, 12, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))        # CPU tensors
    with pytest.raises(RuntimeError):
        SegmentationLoss(20, fused=True).to(DEV)(torch.zeros(1, 20, 4, 4, device=DEV),
                                                 torch.zeros(1, 4, 4, dtype=torch.long, device=DEV))


@pytest.fixture
def no_tf32():
    """Whole-model fp32 comparison: keep convolutions and matmuls in fp32 for the test, then restore."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_lowres_train_path_equals_full_resolution_path(no_tf32):
    """SegModel(x, lowres=True) + fused loss with scale == full-size output + the reference's unfused loss:
    same loss, same parameter gradients (bf16 autocast: compare in fp32 without autocast)."""
    torch.manual_seed(0)
    m = SegModel(YOLOV5_SEG, dcn="dcnv3", img_size=(128, 128)).to(DEV).train()
    x = torch.randn(2, 3, 128, 128, device=DEV)
    lab = torch.randint(0, 12, (2, 128, 128), device=DEV)
    fus = SegmentationLoss(12, class_weights=CAMVID_CLASS_WEIGHTS).to(DEV)
    ref = SegmentationLoss(12, class_weights=CAMVID_CLASS_WEIGHTS, fused=False).to(DEV)
    ref.accepts_lowres = False
    state = {k: v.clone() for k, v in m.state_dict().items()}
    la, _ = forward_loss(m, fus, x, lab)
    la.backward()
    ga = [p.grad.clone() if p.grad is not None else None for p in m.parameters()]
    m.zero_grad(set_to_none=True)
    m.load_state_dict(state)
    lb, _ = forward_loss(m, ref, x, lab)
    lb.backward()
    torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-6)
    for a, p in zip(ga, m.parameters()):
        assert (a is None) == (p.grad is None)
        if a is not None:
            torch.testing.assert_close(a, p.grad, rtol=1e-3, atol=2e-3 * float(p.grad.abs().max()) + 1e-9)
