"""Packed sampling heads (include/dcnv3_b200.h dcnv3_b200_*_packed, DCNv3PackedFunction, DCNv3(packed_heads=True)):
ONE tensor [N, Ho, Wo, 3*G*P] = offsets | mask logits, read by the kernels with a pixel pitch.  Must be the same
function as the unpacked op on the split tensors — forward bit-exact (same kernel, same arithmetic, only addresses
differ), backward within the storage rounding of the 'tile' accumulation (reduction order varies run to run)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _tile_mode(monkeypatch):
    from yolo_dual_b200.ops_dcnv3.functions import get_grad_accum, set_grad_accum
    monkeypatch.delenv("DCNV3_B200_FWD", raising=False)
    monkeypatch.delenv("DCNV3_B200_BWD", raising=False)
    prev = get_grad_accum()
    set_grad_accum("tile")
    yield
    set_grad_accum(prev)


def _inputs(N, H, W, G, gc, dtype, sigma=1.0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(N, H, W, G * gc, device=DEV, generator=g).to(dtype)
    off = (torch.randn(N, H, W, G * 18, device=DEV, generator=g) * sigma).to(dtype)
    m = (torch.randn(N, H, W, G * 9, device=DEV, generator=g) * 2).to(dtype)
    go = torch.randn(N, H, W, G * gc, device=DEV, generator=g).to(dtype)
    return x, off, m, go


CASES = {  # (N, H, W, G, gc), takes the packed kernels?
    "P3_like": ((2, 80, 80, 8, 16), True),
    "partial_tiles_G16": ((1, 21, 19, 16, 16), True),
    "P5_like_G32": ((2, 20, 20, 32, 16), True),
    "G4_falls_back": ((1, 12, 12, 4, 16), False),       # pixel pitch 216 B is not a multiple of 16
    "gc32_falls_back": ((1, 9, 9, 8, 32), False),
}


@pytest.mark.parametrize("logits", [True, False], ids=["logits", "mask"])
@pytest.mark.parametrize("sigma", [1.0, 4.0], ids=["s1", "s4"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", list(CASES))
def test_packed_equals_unpacked(case, dtype, sigma, logits):
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function, DCNv3PackedFunction, DCNv3SoftmaxFunction
    (N, H, W, G, gc), _ = CASES[case]
    x, off, m, go = _inputs(N, H, W, G, gc, dtype, sigma)
    if not logits:
        m = torch.softmax(m.float().view(N, H, W, G, 9), -1).view(N, H, W, G * 9).to(dtype)
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0, 256)
    a = [t.clone().requires_grad_(True) for t in (x, off, m)]
    ya = (DCNv3SoftmaxFunction if logits else DCNv3Function).apply(*a, *args)
    ya.backward(go)
    xb = x.clone().requires_grad_(True)
    heads = torch.cat((off, m), -1).contiguous().requires_grad_(True)
    yb = DCNv3PackedFunction.apply(xb, heads, *args, logits)
    yb.backward(go)
    torch.cuda.synchronize()
    if 10 * H * W >= 6 * 64 * ((H + 7) // 8) * ((W + 7) // 8):
        assert torch.equal(ya, yb)      # both sides ran the staged-window forward: same arithmetic, other addresses
    else:  # the unpacked call takes the vector kernel on maps that fill < 60 % of their 8x8 tiles (fp32 corner weights)
        torch.testing.assert_close(yb.float(), ya.float(), rtol=1e-2, atol=1e-2)
    n_off = G * 18
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    scale = lambda t: max(1.0, float(t.float().abs().max()))
    # grad_offset / grad_mask: one writer per element, same arithmetic
    torch.testing.assert_close(heads.grad[..., :n_off].float(), a[1].grad.float(), rtol=0, atol=0)
    torch.testing.assert_close(heads.grad[..., n_off:].float(), a[2].grad.float(), rtol=0, atol=0)
    assert heads.grad.is_contiguous() and heads.grad.shape == heads.shape
    s = scale(a[0].grad)
    torch.testing.assert_close(xb.grad.float() / s, a[0].grad.float() / s, rtol=0, atol=8 * eps)


def test_packed_entry_points_refuse_what_they_do_not_take():
    import ctypes
    from yolo_dual_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(1, 12, 12, 64, device=DEV, dtype=torch.float16)
    heads = torch.zeros(1, 12, 12, 3 * 4 * 9, device=DEV, dtype=torch.float16)
    out = torch.empty_like(x)
    geo = _lib.Geometry(1, 12, 12, 3, 3, 1, 1, 1, 1, 1, 1, 4, 16, 1.0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # forward: G = 4 is fine (4-byte offset words), backward needs a 16-byte pixel pitch
    assert lib.dcnv3_b200_forward_packed(x.data_ptr(), heads.data_ptr(), out.data_ptr(), _lib.F16, ctypes.byref(geo), 1, st) == 0
    gi, gh = torch.empty_like(x), torch.empty_like(heads)
    rc = lib.dcnv3_b200_backward_packed(x.data_ptr(), heads.data_ptr(), out.data_ptr(), gi.data_ptr(), gh.data_ptr(),
                                        _lib.F16, ctypes.byref(geo), 1, st)
    assert rc == _lib.ENOTSUP and "packed" in _lib.last_error()
    x32 = x.float()
    assert lib.dcnv3_b200_forward_packed(x32.data_ptr(), heads.data_ptr(), out.data_ptr(), _lib.F32, ctypes.byref(geo), 1, st) == _lib.ENOTSUP
    torch.cuda.synchronize()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_module_packed_heads_matches_two_linears(dtype):
    """Same parameters, same state_dict keys; one GEMM + in-place split against two Linear layers + fused softmax."""
    from yolo_dual_b200.ops_dcnv3.modules import DCNv3
    torch.manual_seed(0)
    a = DCNv3(channels=128, group=8, fused_softmax=True).to(DEV)
    with torch.no_grad():
        for lin in (a.offset, a.mask):
            lin.weight.normal_(0, 0.05)
            lin.bias.normal_(0, 0.5)
    b = DCNv3(channels=128, group=8, fused_softmax=True, packed_heads=True).to(DEV)
    assert list(a.state_dict()) == list(b.state_dict())
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 24, 20, 128, device=DEV)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=dtype):
        ya, yb = a(xa), b(xb)
    go = torch.randn_like(ya)
    ya.backward(go)
    yb.backward(go)
    tol = dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(yb.float(), ya.float(), **tol)
    torch.testing.assert_close(xb.grad, xa.grad, **tol)
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        s = max(1e-6, float(pa.grad.abs().max()))
        torch.testing.assert_close(pb.grad / s, pa.grad / s, rtol=0, atol=3e-2, msg=lambda m_, n=n: f"{n}: {m_}")
    with pytest.raises(ValueError):
        DCNv3(channels=64, group=4, packed_heads=True)
