"""Fused training-mode BatchNorm2d + SiLU (yolo_dual_b200/csrc/bnact_b200.cu through its C-ABI) against
torch's batch_norm + silu on the same inputs.  fp32: rtol 1e-4 / atol 1e-5 (different summation order);
16-bit: the fused path rounds once (after SiLU) where torch rounds after BN and after SiLU — compared with the
float32 computation on the same 16-bit inputs at rtol 1e-2 / atol 1e-2 of the largest value."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

from yolo_dual_b200 import _bnact
from yolo_dual_b200.ops_dcnv3.modules.conv import Conv

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference(x32, gamma, beta, rm, rv, eps, momentum, silu):
    y = F.batch_norm(x32, rm, rv, gamma, beta, True, momentum, eps)
    return F.silu(y) if silu else y


@pytest.mark.parametrize("shape", [(16, 64, 80, 80), (2, 1024, 5, 5), (3, 32, 7, 9), (4, 8, 3, 3), (2, 256, 40, 40),
                                   (1, 2048, 2, 1), (16, 128, 33, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("silu", [True, False])
def test_fused_bn_act_matches_torch(shape, dtype, silu):
    n, c, h, w = shape
    if not _bnact.load().bnact_b200_supported(_bnact._DTYPES[dtype], c):
        pytest.skip("channel count outside the fused kernels' range")
    g = torch.Generator().manual_seed(c + h)
    x = (torch.randn(shape, generator=g) * 1.5 + torch.randn(1, c, 1, 1, generator=g) * 3.0).to(DEV).to(dtype)
    x = x.contiguous(memory_format=torch.channels_last)
    gamma = (torch.rand(c, generator=g) + 0.5).to(DEV)
    beta = torch.randn(c, generator=g).to(DEV)
    go = torch.randn(shape, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last)
    rm_a, rv_a = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    rm_b, rv_b = rm_a.clone(), rv_a.clone()
    eps, mom = 1e-3, 0.03

    xa = x.detach().float().clone().requires_grad_(True)
    ga, ba = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    za = _reference(xa, ga, ba, rm_a, rv_a, eps, mom, silu)
    za.backward(go.float())

    xb = x.detach().clone().requires_grad_(True)
    gb, bb = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    zb = _bnact.FusedBNAct.apply(xb, gb, bb, rm_b, rv_b, eps, mom, silu)
    assert zb.dtype == dtype and zb.is_contiguous(memory_format=torch.channels_last)
    zb.backward(go)

    if dtype == torch.float32:
        tol = lambda ref: dict(rtol=1e-4, atol=1e-5 * max(1.0, float(ref.detach().abs().max())))
    else:
        tol = lambda ref: dict(rtol=1e-2, atol=1e-2 * float(ref.detach().abs().max()))
    torch.testing.assert_close(zb.float(), za, **tol(za))
    torch.testing.assert_close(xb.grad.float(), xa.grad, **tol(xa.grad))
    ptol = lambda ref: dict(rtol=2e-3 if dtype != torch.float32 else 1e-4, atol=2e-3 * float(ref.abs().max()) + 1e-6)
    torch.testing.assert_close(gb.grad, ga.grad, **ptol(ga.grad))
    torch.testing.assert_close(bb.grad, ba.grad, **ptol(ba.grad))
    torch.testing.assert_close(rm_b, rm_a, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rv_b, rv_a, rtol=1e-4, atol=1e-5)


def test_statistics_survive_a_large_mean():
    """|mean| >> std: the pivot keeps the variance (a plain E[x^2] - E[x]^2 in float32 would lose it)."""
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(8, 16, 32, 32, generator=g) * 0.01 + 300.0).to(DEV).contiguous(memory_format=torch.channels_last)
    gamma, beta = torch.ones(16, device=DEV), torch.zeros(16, device=DEV)
    rm, rv = torch.zeros(16, device=DEV), torch.ones(16, device=DEV)
    z = _bnact.FusedBNAct.apply(x, gamma, beta, rm, rv, 1e-5, 1.0, False)
    want = F.batch_norm(x.double(), None, None, gamma.double(), beta.double(), True, 0.0, 1e-5)
    torch.testing.assert_close(z.double(), want, rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(rv.double(), x.double().var(dim=(0, 2, 3)), rtol=1e-2, atol=0)


def test_conv_block_takes_the_fused_path_only_where_it_applies(monkeypatch):
    torch.manual_seed(0)
    m = Conv(16, 32, 3).to(DEV).to(memory_format=torch.channels_last).train()
    x = torch.randn(4, 16, 20, 20, device=DEV).contiguous(memory_format=torch.channels_last)
    calls = []
    real = _bnact.bn_act
    monkeypatch.setattr(_bnact, "bn_act", lambda *a: (calls.append(1), real(*a))[1])
    ref = Conv(16, 32, 3).to(DEV).to(memory_format=torch.channels_last).train()
    ref.load_state_dict(m.state_dict())
    monkeypatch.setenv("YOLO_DUAL_B200_FUSED_BN", "0")
    want = ref(x)
    assert not calls
    monkeypatch.setenv("YOLO_DUAL_B200_FUSED_BN", "1")
    got = m(x)
    assert calls == [1]
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    for k, v in m.state_dict().items():
        torch.testing.assert_close(v, ref.state_dict()[k], rtol=1e-4, atol=1e-6)
    calls.clear()
    m.eval()
    m(x)                                                       # eval: running statistics, torch path
    m.train()
    nchw = Conv(16, 32, 3).to(DEV).train()                     # NCHW weights and activations: torch path
    assert not nchw(x.contiguous()).is_contiguous(memory_format=torch.channels_last)
    Conv(16, 12, 1).to(DEV).to(memory_format=torch.channels_last).train()(x)   # 12 channels: not 2^k vectors
    assert not calls
    sync = nn.SyncBatchNorm.convert_sync_batchnorm(Conv(16, 32, 3)).to(DEV)
    assert not _bnact.usable(x.new_zeros(4, 32, 20, 20).contiguous(memory_format=torch.channels_last), sync.bn, sync.act)


@pytest.mark.parametrize("shape", [(4, 64, 40, 40), (2, 1024, 5, 5), (3, 32, 7, 9), (1, 8, 3, 3)])
@pytest.mark.parametrize("mode", ["f32", "half_model", "bf16_model", "bf16_autocast"])
@pytest.mark.parametrize("silu", [True, False])
def test_eval_bn_act_one_pass_matches_torch(shape, mode, silu):
    """Inference path (bnact_b200_eval): running statistics, one pass; parameters float32 or in the model's 16-bit dtype."""
    n, c, h, w = shape
    dt = {"f32": torch.float32, "half_model": torch.float16, "bf16_model": torch.bfloat16, "bf16_autocast": torch.bfloat16}[mode]
    if not _bnact.load().bnact_b200_supported(_bnact._DTYPES[dt], c):
        pytest.skip("channel count outside the fused kernels' range")
    g = torch.Generator().manual_seed(c + w)
    blk = Conv(c, c, 1, act=silu).to(DEV).eval()
    with torch.no_grad():
        blk.bn.weight.copy_(torch.rand(c, generator=g) + 0.5)
        blk.bn.bias.copy_(torch.randn(c, generator=g))
        blk.bn.running_mean.copy_(torch.randn(c, generator=g))
        blk.bn.running_var.copy_(torch.rand(c, generator=g) + 0.25)
    if mode in ("half_model", "bf16_model"):
        blk = blk.to(dt)
    blk = blk.to(memory_format=torch.channels_last)
    y = (torch.randn(shape, generator=g) * 2).to(DEV).to(dt).contiguous(memory_format=torch.channels_last)
    assert _bnact.usable_eval(y, blk.bn, blk.act) is False          # autograd on: not the inference path
    with torch.no_grad():
        assert _bnact.usable_eval(y, blk.bn, blk.act)
        got = _bnact.bn_act_eval(y, blk.bn, blk.act)
        p = [t.float() for t in (blk.bn.running_mean, blk.bn.running_var, blk.bn.weight, blk.bn.bias)]
        want = F.batch_norm(y.float(), p[0], p[1], p[2], p[3], False, 0.0, blk.bn.eps)
        want = F.silu(want) if silu else want
    assert got.dtype == dt and got.is_contiguous(memory_format=torch.channels_last)
    tol = dict(rtol=1e-5, atol=1e-5) if dt == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(got.float(), want, **tol)


def test_eval_conv_block_uses_the_one_pass_kernel_and_matches_the_torch_ops(monkeypatch):
    blk = Conv(32, 64, 3).to(DEV).half().eval().to(memory_format=torch.channels_last)
    x = torch.randn(2, 32, 20, 20, device=DEV).half().contiguous(memory_format=torch.channels_last)
    calls = []
    orig = _bnact.bn_act_eval
    monkeypatch.setattr(_bnact, "bn_act_eval", lambda *a: (calls.append(1), orig(*a))[1])
    with torch.no_grad():
        a = blk(x)
        monkeypatch.setenv("YOLO_DUAL_B200_FUSED_BN", "0")
        b = blk(x)
    assert calls == [1]
    torch.testing.assert_close(a.float(), b.float(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_backward_reads_the_gradient_of_a_cat_slice_in_place(dtype, monkeypatch):
    """torch.cat's backward hands each input a channel-slice VIEW of the concatenated gradient; the fused backward reads
    it at its row pitch (bnact_b200_backward_pitched) instead of copying it.  Same gradients as the torch ops."""
    torch.manual_seed(0)
    a, b = Conv(16, 32, 1).to(DEV), Conv(16, 64, 3).to(DEV)
    for m in (a, b):
        m.to(memory_format=torch.channels_last).train()
    x = torch.randn(4, 16, 12, 10, device=DEV).contiguous(memory_format=torch.channels_last)
    wgt = torch.randn(4, 96, 12, 10, device=DEV).contiguous(memory_format=torch.channels_last)
    seen = []
    lib = _bnact.load()
    orig = lib.bnact_b200_backward_pitched

    class Spy:
        def __call__(self, *args):
            seen.append((args[12], args[14]))   # (C, gz_pitch)
            return orig(*args)
    monkeypatch.setattr(lib, "bnact_b200_backward_pitched", Spy())

    def run():
        for m in (a, b):
            m.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=dtype, enabled=dtype != torch.float32):
            y = torch.cat((a(xi), b(xi)), 1)
        (y.float() * wgt).sum().backward()
        return [xi.grad.clone()] + [p.grad.clone() for m in (a, b) for p in m.parameters()]

    got = run()
    assert sorted(seen) == [(32, 96), (64, 96)], seen    # both slices were read at the cat's pitch: no copy
    monkeypatch.setenv("YOLO_DUAL_B200_FUSED_BN", "0")
    want = run()
    tol = dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
    for g, w_ in zip(got, want):
        s = max(1.0, float(w_.abs().max()))
        torch.testing.assert_close(g / s, w_ / s, **tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [0, 2])
def test_training_c3_without_cat_matches_the_cat_path(dtype, n, monkeypatch):
    """seg.C3 in training: both branches' Conv blocks write into one buffer (FusedBNActInto chained through it, no
    torch.cat, the backward reads the slices of the buffer's gradient in place) — same outputs, gradients and running
    statistics as the torch.cat path on the torch ops."""
    from yolo_dual_b200 import seg
    torch.manual_seed(0)
    blk = seg.C3(32, 64, n=n).to(DEV).to(memory_format=torch.channels_last).train()
    ref = seg.C3(32, 64, n=n).to(DEV).to(memory_format=torch.channels_last).train()
    ref.load_state_dict(blk.state_dict())
    x = torch.randn(4, 32, 16, 12, device=DEV).contiguous(memory_format=torch.channels_last)
    wgt = torch.randn(4, 64, 16, 12, device=DEV)
    used = []
    orig = _bnact.bn_act_into
    monkeypatch.setattr(_bnact, "bn_act_into", lambda *a: (used.append(a[4]), orig(*a))[1])

    def run(m):
        xi = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=dtype, enabled=dtype != torch.float32):
            y = m(xi)
        (y.float() * wgt).sum().backward()
        return y.detach().float(), xi.grad, [p.grad for p in m.parameters()], [b.clone() for b in m.buffers()]

    ya, gxa, gpa, bufa = run(blk)
    assert used == [0, 32], used                      # two slice writers, channel offsets 0 and c_
    monkeypatch.setenv("YOLO_DUAL_B200_FUSED_BN", "0")
    yb, gxb, gpb, bufb = run(ref)
    # (fp32: TF32 convolutions sit between the two BatchNorm implementations' last-bit differences: 2e-4 measured)
    tol = dict(rtol=1e-3, atol=1e-3) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)

    def close(a_, b_):
        s_ = max(1.0, float(b_.float().abs().max()))
        torch.testing.assert_close(a_.float() / s_, b_.float() / s_, **tol)
    close(ya, yb)
    close(gxa, gxb)
    for a_, b_ in zip(gpa, gpb):
        close(a_, b_)
    for a_, b_ in zip(bufa, bufb):
        close(a_, b_)
