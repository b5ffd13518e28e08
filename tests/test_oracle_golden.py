"""Pins the oracles (test infrastructure) to the reference.

* core_torch (restatement of dcnv3_core_pytorch) must reproduce the golden vectors that the
  reference's own function produced (tests/golden/make_golden.py) essentially bit for bit.
* PixelOracle (C restatement of the reference CUDA arithmetic) must agree with the same
  vectors within the north_star tolerance (fp32 rtol 1e-5 / atol 1e-4); it is not bit-equal
  because the reference oracle works in normalised float32 coordinates (SURVEY §7 hard part 2).
* Known-answer tests that pin point order, (x, y) interleave and the bounds gate.
"""
import pytest
import torch
import torch.nn.functional as F

from cases import CASES, BY_NAME, op_args
from util_golden import load_golden
from oracle.dcnv3_oracle import core_torch, core_torch_fwd_bwd, make_inputs, output_hw

IDS = [c["name"] for c in CASES]


def frac_margin(offset, c, eps):
    """True where a sampling coordinate sits within eps px of an integer: grad_offset is
    discontinuous there, and the normalised-coordinate oracle may floor() the other way."""
    N, Ho, Wo, _ = offset.shape
    G, P = c["G"], c["kh"] * c["kw"]
    o = offset.double().view(N, Ho, Wo, G, P, 2)
    i = torch.arange(c["kw"]).repeat_interleave(c["kh"]).double()
    j = torch.arange(c["kh"]).repeat(c["kw"]).double()
    s = c["offset_scale"]
    hw_, hh_ = (c["dw"] * (c["kw"] - 1)) // 2, (c["dh"] * (c["kh"] - 1)) // 2
    wo = torch.arange(Wo).double().view(1, 1, Wo, 1, 1)
    ho = torch.arange(Ho).double().view(1, Ho, 1, 1, 1)
    lw = (hw_ - c["pw"] + wo * c["sw"]) - hw_ * s + (i * c["dw"] + o[..., 0]) * s
    lh = (hh_ - c["ph"] + ho * c["sh"]) - hh_ * s + (j * c["dh"] + o[..., 1]) * s
    near = lambda t: (t - t.round()).abs() < eps
    return (near(lw) | near(lh))  # [N,Ho,Wo,G,P]


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_inputs_regenerate(c):
    """The stored inputs are what make_inputs(seed) draws today."""
    g = load_golden(c["name"])
    dt = getattr(torch, c["dtype"])
    x, off, m, go = make_inputs(c["N"], c["H"], c["W"], c["G"], c["gc"], c["kh"], c["kw"], c["sh"],
                                c["sw"], c["ph"], c["pw"], c["dh"], c["dw"], dist=c["dist"],
                                seed=c["seed"], dtype=torch.float32)
    for a, b in ((x, "input"), (off, "offset"), (m, "mask"), (go, "grad_out")):
        assert torch.equal(a.to(dt), g[b]), b


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_core_torch_matches_reference_golden(c):
    g = load_golden(c["name"])
    torch.set_num_threads(1)
    out, gi, go_, gm = core_torch_fwd_bwd(g["input"], g["offset"], g["mask"], g["grad_out"], *op_args(c))
    tight = dict(rtol=1e-6, atol=1e-7) if c["dtype"] == "float32" else dict(rtol=1e-12, atol=1e-13)
    torch.testing.assert_close(out, g["output"], **tight)
    torch.testing.assert_close(gi, g["grad_input"], **tight)
    torch.testing.assert_close(gm, g["grad_mask"], **tight)
    torch.testing.assert_close(go_, g["grad_offset"], **tight)


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_pixel_oracle_matches_reference_golden(c, pixel_oracle):
    """north_star tolerance: fp32 rtol 1e-5 / atol 1e-4 (same bar used for fp64 here: the
    reference oracle's float32 grid gives it a ~1e-5 px noise floor even in double)."""
    g = load_golden(c["name"])
    a = op_args(c)
    out = pixel_oracle.forward(g["input"], g["offset"], g["mask"], *a)
    gi, go_, gm = pixel_oracle.backward(g["input"], g["offset"], g["mask"], g["grad_out"], *a)
    tol = dict(rtol=1e-5, atol=1e-4)
    # unit-scale data: errors scale with |value|; normalise by max|ref| (SURVEY §7 hard part 2)
    def close(x, y, what):
        scale = max(1.0, float(y.abs().max()))
        torch.testing.assert_close(x / scale, y / scale, msg=lambda m: f"{what}: {m}", **tol)
    close(out, g["output"], "output")
    close(gi, g["grad_input"], "grad_input")
    close(gm, g["grad_mask"], "grad_mask")
    # grad_offset: budget the cell-border points (discontinuous there)
    near = frac_margin(g["offset"], c, 1e-4)
    N, Ho, Wo, _ = g["offset"].shape
    keep = (~near).unsqueeze(-1).expand(*near.shape, 2).reshape(N, Ho, Wo, -1)
    assert near.float().mean() < 0.01, "too many excluded points"
    scale = max(1.0, float(g["grad_offset"].abs().max()))
    torch.testing.assert_close((go_ * keep) / scale, (g["grad_offset"] * keep) / scale, **tol)


def test_kat_avgpool(pixel_oracle):
    """offset = 0, mask = 1/9 (what DCNv3._reset_parameters + softmax(0) give,
    modules/dcnv3.py:100-103) == avg_pool2d(3, 1, 1, count_include_pad=True)."""
    gen = torch.Generator().manual_seed(5)
    N, H, W, G, gc = 2, 11, 13, 2, 4
    x = torch.randn(N, H, W, G * gc, generator=gen)
    off = torch.zeros(N, H, W, G * 9 * 2)
    m = torch.full((N, H, W, G * 9), 1.0 / 9)
    want = F.avg_pool2d(x.permute(0, 3, 1, 2), 3, 1, 1, count_include_pad=True).permute(0, 2, 3, 1)
    a = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    torch.testing.assert_close(pixel_oracle.forward(x, off, m, *a), want.contiguous(), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(core_torch(x, off, m, *a), want.contiguous(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("p,ox,oy", [(0, 0, 0), (5, 1, -1), (8, -2, 1), (3, 0, 2)])
def test_kat_shift(pixel_oracle, p, ox, oy):
    """One-hot mask on point p + integer offsets: out[h,w] = in[h+dy, w+dx] with
    (dx, dy) = (p // Kh - 1 + ox, p % Kh - 1 + oy).  Pins p = i_w*Kh + j_h and (x, y) interleave."""
    gen = torch.Generator().manual_seed(6)
    N, H, W, G, gc = 1, 7, 9, 2, 3
    x = torch.randn(N, H, W, G * gc, generator=gen)
    off = torch.zeros(N, H, W, G, 9, 2)
    off[..., 0] = ox
    off[..., 1] = oy
    m = torch.zeros(N, H, W, G, 9)
    m[..., p] = 1.0
    dx, dy = p // 3 - 1 + ox, p % 3 - 1 + oy
    want = torch.zeros_like(x)
    for h in range(H):
        for w in range(W):
            if 0 <= h + dy < H and 0 <= w + dx < W:
                want[0, h, w] = x[0, h + dy, w + dx]
    a = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    got = pixel_oracle.forward(x, off.reshape(N, H, W, -1), m.reshape(N, H, W, -1), *a)
    assert torch.equal(got, want)
    got2 = core_torch(x, off.reshape(N, H, W, -1), m.reshape(N, H, W, -1), *a)
    torch.testing.assert_close(got2, want, rtol=1e-5, atol=1e-5)


def test_kat_bounds(pixel_oracle):
    """Points pushed to loc <= -1 or loc >= H/W: exact zeros in output and all three grads, and
    a zero bounds byte (gate closed) — dcnv3_im2col_cuda.cuh:262-263."""
    gen = torch.Generator().manual_seed(7)
    N, H, W, G, gc = 1, 6, 6, 1, 4
    x = torch.randn(N, H, W, G * gc, generator=gen)
    off = torch.full((N, H, W, G * 9 * 2), 100.0)
    off[..., ::4] = -100.0
    m = torch.full((N, H, W, G * 9), 1.0 / 9)
    go = torch.randn(N, H, W, G * gc, generator=gen)
    a = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    assert pixel_oracle.forward(x, off, m, *a).abs().max() == 0
    gi, go_, gm = pixel_oracle.backward(x, off, m, go, *a)
    assert gi.abs().max() == 0 and go_.abs().max() == 0 and gm.abs().max() == 0
    hw, bd = pixel_oracle.indices(off, H, W, 3, 3, 1, 1, 1, 1, 1, 1, G, 1.0)
    assert int(bd.max()) == 0 and int(hw.abs().max()) == 0


def test_indices_exact_on_grid(pixel_oracle):
    """Zero offsets, stride 1, pad 1: h_low = ho + j - 1, w_low = wo + i - 1; the low corner is
    valid iff it lies inside the map; gate closed only where loc == -1 exactly."""
    N, H, W, G = 1, 5, 4, 1
    off = torch.zeros(N, H, W, G * 18)
    hw, bd = pixel_oracle.indices(off, H, W, 3, 3, 1, 1, 1, 1, 1, 1, G, 1.0)
    for ho in range(H):
        for wo in range(W):
            for i in range(3):
                for j in range(3):
                    p = i * 3 + j
                    h, w = ho + j - 1, wo + i - 1
                    inside = h > -1 and w > -1 and h < H and w < W
                    b = int(bd[0, ho, wo, 0, p])
                    assert (b & 1) == int(inside)
                    if inside:
                        assert tuple(hw[0, ho, wo, 0, p].tolist()) == (h, w)
                        assert (b >> 1) & 1 == 1
                        assert (b >> 2) & 1 == int(w + 1 <= W - 1)
                        assert (b >> 3) & 1 == int(h + 1 <= H - 1)


def test_output_hw_matches_reference_formula():
    # dcnv3_cuda.cu:40-45
    assert output_hw(80, 80, 3, 3, 1, 1, 1, 1, 1, 1) == (80, 80)
    assert output_hw(11, 13, 3, 3, 2, 2, 1, 1, 1, 1) == (6, 7)
    assert output_hw(12, 10, 5, 5, 1, 1, 4, 4, 2, 2) == (12, 10)
