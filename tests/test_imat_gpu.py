"""GPU parity of the interpolation-matrix kernel family (csrc/dcnv3_imat.cuh) through the C-ABI.

The family is the default backward for 16-bit storage with group_channels = 16, 3x3 / stride 1 /
dilation 1 (the C3-DCN shapes); its forward is opt-in (DCNV3_B200_FWD=imat).  Checked against the CPU
pixel oracle (test infrastructure) on shapes that exercise: whole and partial 8x8 tiles, maps smaller
than one tile, pad 0, offset_scale != 1, G = 4 ... 32, offsets far outside the staged window (the
per-lane global-memory path: the reference test's own `rand*10` distribution), fused softmax.

Tolerance: bf16/fp16 rtol 1e-2, atol 2e-3 after scaling by max|ref| (north_star).  The products run on
the tensor cores with fp16 interpolation weights (2^-11 relative) and exact 16-bit activations, fp32
accumulation; grad_input additionally sums per-tile partial windows with fp32 reductions.  For bf16
storage grad_output enters that product scaled by a power of two per (tile, group) so that it fits
fp16 (`test_imat_bf16_gradient_range` drives the scaling through tiny, huge and mixed magnitudes).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _reference_accumulation():
    """This module checks the reference's own 16-bit semantics (fp32 accumulation of grad_input, one rounding:
    grad_accum 'opmath') and the explicit 'storage' opt-in; the default 'tile' mode has its own module, test_win_gpu.py."""
    from yolo_dual_b200.ops_dcnv3.functions import get_grad_accum, set_grad_accum
    prev = get_grad_accum()
    set_grad_accum("opmath")
    yield
    set_grad_accum(prev)

CASES = {
    "cfg1_G4": ((2, 80, 80, 4, 16), dict()),
    "partial_tiles_G8": ((2, 21, 19, 8, 16), dict()),
    "pad0": ((1, 12, 12, 4, 16), dict(pad=0)),
    "scale1.5": ((1, 17, 23, 4, 16), dict(scale=1.5)),
    "smaller_than_a_tile": ((1, 3, 5, 4, 16), dict()),
    "P5_like_G32": ((1, 20, 20, 32, 16), dict()),
    "pad2": ((1, 10, 14, 4, 16), dict(pad=2)),
}


def _run(fn, x, off, m, go, args, dtype):
    xs, os_, ms = (t.to(DEV, dtype).contiguous().requires_grad_(True) for t in (x, off, m))
    out = fn.apply(xs, os_, ms, *args, 256)
    out.backward(go.to(DEV, dtype))
    torch.cuda.synchronize()
    return [t.float().cpu() for t in (out.detach(), xs.grad, os_.grad, ms.grad)]


def _close(got, want, what, rtol=1e-2, atol=2e-3):
    scale = max(1.0, float(want.abs().max()))
    torch.testing.assert_close(got.double() / scale, want.double() / scale, rtol=rtol, atol=atol,
                               msg=lambda s: f"{what}: {s}")


@pytest.mark.parametrize("fwd", ["vec", "imat", "pts", "win"])
@pytest.mark.parametrize("dist", ["unit", "ref"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", list(CASES))
def test_imat_vs_pixel_oracle(case, dtype, dist, fwd, pixel_oracle, monkeypatch):
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    monkeypatch.setenv("DCNV3_B200_FWD", fwd)
    monkeypatch.setenv("DCNV3_B200_BWD", "imat")
    (N, H, W, G, gc), kw = CASES[case]
    pad, scale = kw.get("pad", 1), kw.get("scale", 1.0)
    args = (3, 3, 1, 1, pad, pad, 1, 1, G, gc, scale)
    x, off, m, go = make_inputs(N, H, W, G, gc, 3, 3, 1, 1, pad, pad, 1, 1, dist=dist, seed=5)
    xr, offr, mr, gor = (t.to(dtype).float() for t in (x, off, m, go))
    want = [pixel_oracle.forward(xr, offr, mr, *args)] + list(pixel_oracle.backward(xr, offr, mr, gor, *args))
    got = _run(DCNv3Function, x, off, m, go, args, dtype)
    for g_, w_, name in zip(got, want, ("output", "grad_input", "grad_offset", "grad_mask")):
        _close(g_, w_, name)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_imat_matches_vector_family(dtype, monkeypatch):
    """Same inputs through both families (P3-like, 2 images): they must agree to storage rounding."""
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    N, H, W, G, gc = 2, 80, 80, 8, 16
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, m, go = make_inputs(N, H, W, G, gc, dist="unit", seed=11)
    res = {}
    for fam in ("vec", "imat"):
        monkeypatch.setenv("DCNV3_B200_FWD", fam)
        monkeypatch.setenv("DCNV3_B200_BWD", fam)
        res[fam] = _run(DCNv3Function, x, off, m, go, args, dtype)
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -10
    for a, b, name in zip(res["imat"], res["vec"], ("output", "grad_input", "grad_offset", "grad_mask")):
        _close(a, b, name, rtol=2 * eps, atol=2 * eps)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_imat_fused_softmax(dtype, pixel_oracle, monkeypatch):
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3SoftmaxFunction
    monkeypatch.setenv("DCNV3_B200_FWD", "imat")
    monkeypatch.setenv("DCNV3_B200_BWD", "imat")
    N, H, W, G, gc = 2, 21, 19, 8, 16
    P = 9
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, _, go = make_inputs(N, H, W, G, gc, dist="unit", seed=4)
    logits = torch.randn(N, H, W, G * P, generator=torch.Generator().manual_seed(8)) * 2
    xr, offr, lr, gor = (t.to(dtype).float() for t in (x, off, logits, go))
    prob = torch.softmax(lr.view(N, H, W, G, P), -1)
    pm = prob.reshape(N, H, W, G * P).contiguous()
    want_out = pixel_oracle.forward(xr, offr, pm, *args)
    want_gi, want_go, gm = pixel_oracle.backward(xr, offr, pm, gor, *args)
    gmv = gm.view(N, H, W, G, P)
    want_gl = (prob * (gmv - (prob * gmv).sum(-1, keepdim=True))).reshape(N, H, W, G * P)
    got = _run(DCNv3SoftmaxFunction, x, off, logits, go, args, dtype)
    for g_, w_, name in zip(got, (want_out, want_gi, want_go, want_gl), ("output", "grad_input", "grad_offset", "grad_logits")):
        _close(g_, w_, name)


def test_imat_is_the_default_backward_and_adjoint_holds(monkeypatch):
    """At a BASELINE site (P4, bf16) with no knobs set: <go, f(x)> = <grad_input, x> (the op is linear
    in the input), which only holds if the tiled reductions neither drop nor double-count a cell."""
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    monkeypatch.delenv("DCNV3_B200_FWD", raising=False)
    monkeypatch.delenv("DCNV3_B200_BWD", raising=False)
    N, H, W, G, gc = 4, 40, 40, 16, 16
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, m, go = make_inputs(N, H, W, G, gc, dist="unit", seed=3)
    out, gi, _, gm = _run(DCNv3Function, x, off, m, go, args, torch.bfloat16)
    xb, gob, mb = (t.to(torch.bfloat16).double() for t in (x, go, m))
    lhs = float((gob * out.double()).sum())
    rhs = float((gi.double() * xb).sum())
    assert abs(lhs - rhs) <= 2e-3 * max(abs(lhs), abs(rhs), float(N * H * W)), (lhs, rhs)
    rhs_m = float((gm.double() * mb).sum())  # and <grad_mask, mask> = <go, f>
    assert abs(lhs - rhs_m) <= 2e-3 * max(abs(lhs), abs(rhs_m), float(N * H * W)), (lhs, rhs_m)


@pytest.mark.parametrize("mode", ["tiny", "huge", "rows_mixed", "rows_mixed_rev", "outlier_pixel"])
def test_imat_bf16_gradient_range(mode, pixel_oracle, monkeypatch):
    """bf16 grad_output spans far more than fp16's range; the imat backward rescales it by a power of two per
    (tile, group) and, when a later pass of a tile is > 2^14 larger than an earlier one, rescales its
    accumulators.  grad_input must stay within the bf16 bar relative to max|ref| in every regime, and
    grad_offset / grad_mask (which never see the scaling) as always."""
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    monkeypatch.setenv("DCNV3_B200_BWD", "imat")
    N, H, W, G, gc = 1, 16, 16, 4, 16
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, m, go = make_inputs(N, H, W, G, gc, 3, 3, 1, 1, 1, 1, 1, 1, dist="unit", seed=11)
    if mode == "tiny":
        go = go * 1e-30
    elif mode == "huge":
        go = go * 1e30
    elif mode == "rows_mixed":      # upper half of every tile (pass 0) tiny, lower half (pass 1) 2^40 larger
        rows = torch.arange(H) % 8 < 4
        go = go * torch.where(rows, 2.0 ** -30, 2.0 ** 10).view(1, H, 1, 1)
    elif mode == "rows_mixed_rev":  # the large half first: the small half underflows fp16 after scaling (by design)
        rows = torch.arange(H) % 8 < 4
        go = go * torch.where(rows, 2.0 ** 10, 2.0 ** -30).view(1, H, 1, 1)
    else:                           # one pixel 2^20 above its neighbours
        go = go * 2.0 ** -10
        go[0, 5, 6] *= 2.0 ** 20
    dtype = torch.bfloat16
    xr, offr, mr, gor = (t.to(dtype).float() for t in (x, off, m, go))
    want = list(pixel_oracle.backward(xr, offr, mr, gor, *args))
    got = _run(DCNv3Function, x, off, m, go, args, dtype)[1:]
    for g_, w_, name in zip(got, want, ("grad_input", "grad_offset", "grad_mask")):
        assert torch.isfinite(g_).all(), name
        scale = float(w_.abs().max())
        torch.testing.assert_close(g_.double() / scale, w_.double() / scale, rtol=1e-2, atol=2e-3,
                                   msg=lambda s: f"{mode} {name}: {s}")
    if mode == "rows_mixed":
        # the small half is 2^-40 of the tile maximum: below fp16 after scaling, but the cells only IT reaches
        # (window rows above the large half's reach) must not be garbage: absolute error stays below 2^-24 max
        gi, wi = got[0].double(), want[0].double()
        assert float((gi - wi).abs().max()) <= 2.0 ** -8 * float(wi.abs().max())


@pytest.mark.parametrize("fwd", ["win", "vec"])
def test_forward_non_finite_input_stays_local(fwd, pixel_oracle, monkeypatch):
    """A closed sampling point (outside the map) contributes nothing — it must not multiply some unrelated
    staged cell by a zero weight: with an Inf in the input that would turn 0 * Inf into NaN far away from the
    pixels that really sample it.  The staged-window forward parks closed points on zero cells."""
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    monkeypatch.setenv("DCNV3_B200_FWD", fwd)
    N, H, W, G, gc = 1, 24, 24, 4, 16
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, m, _ = make_inputs(N, H, W, G, gc, 3, 3, 1, 1, 1, 1, 1, 1, dist="unit", seed=21)
    x[0, 2, 2, :] = float("inf")      # = window cell (0, 0) of tile (1, 1)
    off = off.view(N, H, W, G, 9, 2).clone()
    off[..., 0, :] = 100.0            # point 0 of every (pixel, group) leaves the map: gate closed
    off = off.view(N, H, W, G * 18)
    dtype = torch.bfloat16
    xr, offr, mr = (t.to(dtype).float() for t in (x, off, m))
    want = pixel_oracle.forward(xr, offr, mr, *args)
    xs, os_, ms = (t.to(DEV, dtype).contiguous() for t in (x, off, m))
    got = DCNv3Function.apply(xs, os_, ms, *args, 256).float().cpu()
    fin_w, fin_g = torch.isfinite(want), torch.isfinite(got)
    assert fin_w.float().mean() > 0.9            # the Inf reaches only the pixels around (2, 2)
    assert torch.equal(fin_w, fin_g)
    torch.testing.assert_close(got[fin_g], want[fin_w], rtol=1e-2, atol=2e-2)
