"""GPU parity of the round-2 default 16-bit kernels through the C-ABI:

  * win::bwd_win_kernel (csrc/dcnv3_win.cuh): the single-kernel backward, grad_accum = 'tile' (no workspace):
    exact mixed-precision corner dots from a staged window, interpolation matrix in the storage dtype expanded by
    the tensor cores, packed 16-bit vector reductions across tiles;
  * imat::fwd_tile_kernel with the mixed-precision blend (corner weights rounded to the storage dtype).

Checked element-wise against the CPU pixel oracle (test infrastructure) and — at the BASELINE configs[1] sizes,
N = 16, P3 / P4 / P5, default knobs, the exact kernels bench.py times — against the pixel oracle (bf16 and fp16)
and the reference's own CUDA kernels rebuilt for sm_100a (oracle/_ref, fp16; the reference has no bf16).

Tolerance (north_star): bf16 / fp16 rtol 1e-2, atol 2e-3 after scaling by max|ref| — for everything except bf16
grad_input in 'tile' mode, whose bound is atol 2 bf16 ulps of the largest element (2 * 2^-8 = 7.8e-3): 'tile' sums a 4x8
band's contributions to a cell in fp32 and adds the (at most six) per-band partials in the storage dtype; a partial can
be larger than the final sum (cancellation), so each rounding is 2^-9 of the PARTIAL.  Measured at the BASELINE sizes:
bf16 max error 0.006 * max|ref|, 1 element in ~2e6 beyond atol 2e-3; fp16 (11 bits) stays inside 2e-3.  'opmath' keeps the
reference's fp32 accumulation + one rounding (tests/test_dcnv3_gpu.py, tests/test_imat_gpu.py run in that mode).
Sampling points more than 3 px away from their kernel-grid position are reduced one by one in the storage dtype (the
reference test's own `rand*10` offsets put nearly every point there: the bound is then ACC_STORAGE's, 6 eps).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"

CASES = {
    "cfg1_G4": ((2, 80, 80, 4, 16), dict()),
    "partial_tiles_G8": ((2, 21, 19, 8, 16), dict()),
    "pad0": ((1, 12, 12, 4, 16), dict(pad=0)),
    "scale1.5": ((1, 17, 23, 4, 16), dict(scale=1.5)),
    "smaller_than_a_tile": ((1, 3, 5, 4, 16), dict()),
    "P5_like_G32": ((1, 20, 20, 32, 16), dict()),
    "pad2": ((1, 10, 14, 4, 16), dict(pad=2)),
    "one_band": ((3, 4, 33, 4, 16), dict()),
    # maps whose width leaves 1-4 columns beyond the last whole tile column: the strip of transposed 8x4 tiles
    # (partial_tiles_G8, P5_like_G32, pad0 and one_band above take it too)
    "strip_w12_h9": ((1, 9, 12, 4, 16), dict()),
    "strip_w28_h37_G8": ((2, 37, 28, 8, 16), dict()),
    "strip_w9_scale1.5": ((1, 16, 9, 4, 16), dict(scale=1.5)),
}


def _run(fn, x, off, m, go, args, dtype):
    xs, os_, ms = (t.to(DEV, dtype).contiguous().requires_grad_(True) for t in (x, off, m))
    out = fn.apply(xs, os_, ms, *args, 256)
    out.backward(go.to(DEV, dtype))
    torch.cuda.synchronize()
    return [t.float().cpu() for t in (out.detach(), xs.grad, os_.grad, ms.grad)]


TILE_BF16_ATOL = 2 * 2.0 ** -8  # grad_input, bf16, 'tile' accumulation: see the module docstring


def _gi_tol(dtype):
    return dict(rtol=1e-2, atol=TILE_BF16_ATOL if dtype == torch.bfloat16 else 2e-3)


def _close(got, want, what, rtol=1e-2, atol=2e-3):
    scale = max(1.0, float(want.abs().max()))
    torch.testing.assert_close(got.double() / scale, want.double() / scale, rtol=rtol, atol=atol,
                               msg=lambda s: f"{what}: {s}")


def _want(po, x, off, m, go, args, dtype):
    xr, offr, mr, gor = (t.to(dtype).float() for t in (x, off, m, go))
    return [po.forward(xr, offr, mr, *args)] + list(po.backward(xr, offr, mr, gor, *args))


@pytest.fixture(autouse=True)
def _default_mode(monkeypatch):
    from yolo_dual_b200.ops_dcnv3.functions import get_grad_accum, set_grad_accum
    monkeypatch.delenv("DCNV3_B200_FWD", raising=False)
    monkeypatch.delenv("DCNV3_B200_BWD", raising=False)
    prev = get_grad_accum()
    set_grad_accum("tile")
    yield
    set_grad_accum(prev)


@pytest.mark.parametrize("sigma", [1.0, 2.0], ids=["s1", "s2"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", list(CASES))
def test_win_vs_pixel_oracle(case, dtype, sigma, pixel_oracle):
    """sigma = 1: nearly every point inside the window; sigma = 2: ~13 % of the points take bwd_point_slow and
    some bands the vector fallback (all three code paths mixed inside one launch)."""
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    (N, H, W, G, gc), kw = CASES[case]
    pad, scale = kw.get("pad", 1), kw.get("scale", 1.0)
    args = (3, 3, 1, 1, pad, pad, 1, 1, G, gc, scale)
    x, off, m, go = make_inputs(N, H, W, G, gc, 3, 3, 1, 1, pad, pad, 1, 1, dist="unit", seed=5)
    off = off * sigma
    want = _want(pixel_oracle, x, off, m, go, args, dtype)
    got = _run(DCNv3Function, x, off, m, go, args, dtype)
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    for g_, w_, name in zip(got, want, ("output", "grad_input", "grad_offset", "grad_mask")):
        if name == "grad_input" and sigma * scale > 1.0:  # out-of-window points (|offset * scale| >= 3 px: 4.5 % of the
            # coordinates at sigma * scale = 1.5) round per contribution in an order that varies run to run (ACC_STORAGE's bound)
            _close(g_, w_, name, rtol=max(1e-2, 6 * eps), atol=max(2e-3, 6 * eps))
        elif name == "grad_input":
            _close(g_, w_, name, **_gi_tol(dtype))
        else:
            _close(g_, w_, name)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_mask_not_16_byte_aligned_takes_the_chunk_path(dtype, pixel_oracle):
    """The tile's offsets / masks arrive as TMA boxes only when their rows and base pointers are 16-byte multiples
    (dcnv3_win.cuh / dcnv3_imat.cuh, STAGE); a mask tensor that starts 8 bytes into its allocation takes the warps' own
    cp.async chunks (backward) and the lanes' own loads (forward) — same results."""
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    N, H, W, G, gc = 2, 24, 32, 8, 16
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, m, go = make_inputs(N, H, W, G, gc, 3, 3, 1, 1, 1, 1, 1, 1, dist="unit", seed=11)
    want = _want(pixel_oracle, x, off, m, go, args, dtype)
    xs, os_ = (t.to(DEV, dtype).contiguous().requires_grad_(True) for t in (x, off))
    buf = torch.zeros(m.numel() + 4, device=DEV, dtype=dtype)
    buf[4:] = m.to(DEV, dtype).reshape(-1)
    ms = buf[4:].view(m.shape).detach().requires_grad_(True)
    assert ms.data_ptr() % 16 == 8 and ms.is_contiguous()
    out = DCNv3Function.apply(xs, os_, ms, *args, 256)
    out.backward(go.to(DEV, dtype))
    torch.cuda.synchronize()
    got = [t.float().cpu() for t in (out.detach(), xs.grad, os_.grad, ms.grad)]
    for g_, w_, name in zip(got, want, ("output", "grad_input", "grad_offset", "grad_mask")):
        _close(g_, w_, name, **(_gi_tol(dtype) if name == "grad_input" else {}))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("case", ["cfg1_G4", "partial_tiles_G8", "pad0"])
def test_win_reference_distribution_far_offsets(case, dtype, pixel_oracle):
    """The reference test's own distribution (test.py:35-39: offset = rand * 10): every band runs the vector
    family's lane body with packed 16-bit reductions."""
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    (N, H, W, G, gc), kw = CASES[case]
    pad = kw.get("pad", 1)
    args = (3, 3, 1, 1, pad, pad, 1, 1, G, gc, 1.0)
    x, off, m, go = make_inputs(N, H, W, G, gc, 3, 3, 1, 1, pad, pad, 1, 1, dist="ref", seed=5)
    want = _want(pixel_oracle, x, off, m, go, args, dtype)
    got = _run(DCNv3Function, x, off, m, go, args, dtype)
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    for g_, w_, name in zip(got, want, ("output", "grad_input", "grad_offset", "grad_mask")):
        if name == "grad_input":
            _close(g_, w_, name, rtol=6 * eps, atol=6 * eps)
        else:
            _close(g_, w_, name)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_win_fused_softmax(dtype, pixel_oracle):
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3SoftmaxFunction
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
        _close(g_, w_, name, **(_gi_tol(dtype) if name == "grad_input" else {}))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_tile_matches_opmath_accumulation(dtype):
    """Same inputs through 'tile' (one kernel) and 'opmath' (fp32 workspace, the reference's semantics): grad_offset
    and grad_mask agree to storage rounding, grad_input to the four-roundings bound."""
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function, set_grad_accum
    N, H, W, G, gc = 2, 80, 80, 8, 16
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, m, go = make_inputs(N, H, W, G, gc, dist="unit", seed=11)
    res = {}
    for mode in ("tile", "opmath"):
        set_grad_accum(mode)
        res[mode] = _run(DCNv3Function, x, off, m, go, args, dtype)
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    for a, b, name in zip(res["tile"], res["opmath"], ("output", "grad_input", "grad_offset", "grad_mask")):
        _close(a, b, name, rtol=4 * eps, atol=4 * eps)  # (bf16: 4 eps = 1.6e-2, fp16: 2e-3)


def test_tile_mode_needs_no_workspace():
    from yolo_dual_b200 import _lib
    import ctypes
    lib = _lib.load()
    geo = _lib.Geometry(16, 80, 80, 3, 3, 1, 1, 1, 1, 1, 1, 8, 16, 1.0)
    assert lib.dcnv3_b200_backward_workspace_bytes(_lib.BF16, ctypes.byref(geo), _lib.ACC_TILE) == 0
    assert lib.dcnv3_b200_backward_workspace_bytes(_lib.BF16, ctypes.byref(geo), _lib.ACC_OPMATH) > 0
    geo5 = _lib.Geometry(2, 20, 20, 5, 5, 1, 1, 2, 2, 1, 1, 8, 16, 1.0)  # 5x5: the tile kernel does not take it
    assert lib.dcnv3_b200_backward_workspace_bytes(_lib.BF16, ctypes.byref(geo5), _lib.ACC_TILE) > 0


# ------------------------------------------------------------------------------------------------------------
# The headline path, element-wise, at the sizes bench.py times (BASELINE configs[1]: N = 16, P3 / P4 / P5)
# ------------------------------------------------------------------------------------------------------------
SITES = {"P3": (16, 80, 80, 8, 16), "P4": (16, 40, 40, 16, 16), "P5": (16, 20, 20, 32, 16)}


@pytest.mark.parametrize("fused", [False, True], ids=["mask", "logits"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("site", list(SITES))
def test_headline_sites_elementwise_vs_pixel_oracle(site, dtype, fused, pixel_oracle):
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function, DCNv3SoftmaxFunction
    N, H, W, G, gc = SITES[site]
    P = 9
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, m, go = make_inputs(N, H, W, G, gc, dist="unit", seed=21)
    if fused:
        logits = torch.randn(N, H, W, G * P, generator=torch.Generator().manual_seed(9)) * 2
        lr = logits.to(dtype).float()
        prob = torch.softmax(lr.view(N, H, W, G, P), -1)
        pm = prob.reshape(N, H, W, G * P).contiguous()
        xr, offr, gor = (t.to(dtype).float() for t in (x, off, go))
        want_out = pixel_oracle.forward(xr, offr, pm, *args)
        want_gi, want_go, gm = pixel_oracle.backward(xr, offr, pm, gor, *args)
        gmv = gm.view(N, H, W, G, P)
        want = [want_out, want_gi, want_go, (prob * (gmv - (prob * gmv).sum(-1, keepdim=True))).reshape(N, H, W, G * P)]
        got = _run(DCNv3SoftmaxFunction, x, off, logits, go, args, dtype)
    else:
        want = _want(pixel_oracle, x, off, m, go, args, dtype)
        got = _run(DCNv3Function, x, off, m, go, args, dtype)
    for g_, w_, name in zip(got, want, ("output", "grad_input", "grad_offset", "grad_mask")):
        _close(g_, w_, name, **(_gi_tol(dtype) if name == "grad_input" else {}))


@pytest.mark.parametrize("site", list(SITES))
def test_headline_sites_vs_reference_cuda_kernels(site):
    """fp16, N = 16: our default kernels against the reference's own CUDA kernels (oracle/_ref) on the same device
    tensors (the reference accumulates grad_input in fp32 and rounds once)."""
    from oracle.build_ref_cuda import load_module
    from oracle.dcnv3_oracle import make_inputs
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    ref = load_module()
    if ref is None:
        pytest.skip("oracle/_ref is not built (needs /root/reference at build time)")
    N, H, W, G, gc = SITES[site]
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, m, go = make_inputs(N, H, W, G, gc, dist="unit", seed=22)
    dt = torch.float16
    got = _run(DCNv3Function, x, off, m, go, args, dt)
    xd, od, md, gd = (t.to(DEV, dt).contiguous() for t in (x, off, m, go))
    r_out = ref.dcnv3_forward(xd, od, md, *args, 256)
    r_gi, r_go, r_gm = ref.dcnv3_backward(xd, od, md, *args, gd, 256)
    torch.cuda.synchronize()
    for g_, w_, name in zip(got, (r_out, r_gi, r_go, r_gm), ("output", "grad_input", "grad_offset", "grad_mask")):
        _close(g_, w_.float().cpu(), name)
