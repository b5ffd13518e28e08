"""The reference-side binding shown in INTEGRATION.md §3 (tests/integration_shim/DCNv3.py)."""
import importlib
import os
import sys
import warnings

import pytest
import torch

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "integration_shim")
REF_LIB = "/root/reference/models/ops_dcnv3/build/lib.linux-x86_64-cpython-38"


def _shim():
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    sys.modules.pop("DCNv3", None)
    return importlib.import_module("DCNv3")


def test_shim_exports_the_pybind_entry_points():
    from yolo_dual_b200.build import build
    build()
    m = _shim()
    assert callable(m.dcnv3_forward) and callable(m.dcnv3_backward)
    import inspect
    assert list(inspect.signature(m.dcnv3_forward).parameters)[-1] == "im2col_step"
    assert list(inspect.signature(m.dcnv3_backward).parameters)[-2:] == ["grad_output", "im2col_step"]


@pytest.mark.skipif(not os.path.isdir(REF_LIB), reason="/root/reference is only mounted in the build container")
def test_reference_dcnv3_func_imports_unmodified_on_top_of_the_shim():
    """The reference's own functions/dcnv3_func.py does `import DCNv3` (:16); with the shim on sys.path it
    imports as is and its DCNv3Function resolves dcnv3_forward / dcnv3_backward in our library."""
    from yolo_dual_b200.build import build
    build()
    shim = _shim()
    sys.path.insert(0, REF_LIB)
    try:
        for k in [k for k in sys.modules if k == "functions" or k.startswith("functions.")]:
            sys.modules.pop(k)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = importlib.import_module("functions.dcnv3_func")
        assert ref.DCNv3 is shim
        assert hasattr(ref.DCNv3Function, "apply") and callable(ref.dcnv3_core_pytorch)
    finally:
        sys.path.remove(REF_LIB)
        for k in [k for k in sys.modules if k == "functions" or k.startswith("functions.")]:
            sys.modules.pop(k)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_shim_matches_oracle_on_gpu(dtype, pixel_oracle):
    from oracle.dcnv3_oracle import make_inputs
    m = _shim()
    N, H, W, G, gc = 2, 20, 24, 4, 16
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, msk, go = make_inputs(N, H, W, G, gc, dist="unit", seed=41)
    xs, os_, ms, gs = (t.cuda().to(dtype) for t in (x, off, msk, go))
    out = m.dcnv3_forward(xs, os_, ms, *args, 256)
    gi, goff, gm = m.dcnv3_backward(xs, os_, ms, *args, gs, 256)
    f = lambda t: t.to(dtype).float()
    want = pixel_oracle.forward(f(x), f(off), f(msk), *args)
    wgi, wgo, wgm = pixel_oracle.backward(f(x), f(off), f(msk), f(go), *args)
    tol = dict(rtol=1e-5, atol=1e-4) if dtype == torch.float32 else dict(rtol=1e-2, atol=2e-2)
    for a, b in ((out, want), (gi, wgi), (goff, wgo), (gm, wgm)):
        torch.testing.assert_close(a.float().cpu(), b, **tol)
