"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/dcnv3_b200.h declares, validates arguments before touching a device, and has no CPU
fallback (compute entry points return EDEVICE on a box without a GPU).  No kernel runs here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from yolo_dual_b200 import _lib
    from yolo_dual_b200.build import build
    build()  # compile if missing/stale (nvcc cross-compiles without a GPU)
    return _lib.load()


def _geo(**kw):
    from yolo_dual_b200._lib import Geometry
    d = dict(N=2, H=8, W=8, kernel_h=3, kernel_w=3, stride_h=1, stride_w=1, pad_h=1, pad_w=1,
             dilation_h=1, dilation_w=1, group=4, group_channels=16, offset_scale=1.0)
    d.update(kw)
    return Geometry(*[d[n] for n, _ in Geometry._fields_])


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "dcnv3_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dcnv3_b200_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 7
    from yolo_dual_b200 import _lib
    assert declared == set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_version_matches_header(lib):
    hdr = open(os.path.join(ROOT, "include", "dcnv3_b200.h")).read()
    v = int(re.search(r"#define DCNV3_B200_VERSION (\d+)", hdr).group(1))
    assert lib.dcnv3_b200_version() == v


def test_library_has_sm100a_code_only():
    from yolo_dual_b200.build import LIB
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("kw,HoWo", [
    (dict(H=80, W=80), (80, 80)),
    (dict(H=11, W=13, stride_h=2, stride_w=2), (6, 7)),
    (dict(H=12, W=10, kernel_h=5, kernel_w=5, pad_h=4, pad_w=4, dilation_h=2, dilation_w=2), (12, 10)),
    (dict(H=9, W=9, pad_h=0, pad_w=0), (7, 7)),
])
def test_output_size_matches_reference_formula(lib, kw, HoWo):
    from oracle.dcnv3_oracle import output_hw
    g = _geo(**kw)
    ho, wo = ctypes.c_int(), ctypes.c_int()
    assert lib.dcnv3_b200_output_size(ctypes.byref(g), ctypes.byref(ho), ctypes.byref(wo)) == 0
    assert (ho.value, wo.value) == HoWo
    assert HoWo == output_hw(g.H, g.W, g.kernel_h, g.kernel_w, g.stride_h, g.stride_w, g.pad_h,
                             g.pad_w, g.dilation_h, g.dilation_w)


@pytest.mark.parametrize("kw", [dict(group=0), dict(group_channels=-1), dict(kernel_h=0), dict(stride_w=0),
                                dict(pad_h=-1), dict(dilation_w=0), dict(H=0), dict(N=-1),
                                dict(H=2, W=2, kernel_h=7, kernel_w=7, pad_h=0, pad_w=0)])
def test_bad_geometry_is_rejected_with_a_message(lib, kw):
    g = _geo(**kw)
    ho, wo = ctypes.c_int(), ctypes.c_int()
    rc = lib.dcnv3_b200_output_size(ctypes.byref(g), ctypes.byref(ho), ctypes.byref(wo))
    assert rc == -1
    assert len(lib.dcnv3_b200_last_error()) > 0
    assert lib.dcnv3_b200_forward(1, 1, 1, 1, 0, ctypes.byref(g), 0, None) == -1


def test_argument_validation_order(lib):
    g = _geo()
    assert lib.dcnv3_b200_forward(1, 1, 1, 1, 99, ctypes.byref(g), 0, None) == -1      # dtype
    assert lib.dcnv3_b200_forward(1, 1, 1, 1, 0, ctypes.byref(g), 7, None) == -1       # flag
    assert lib.dcnv3_b200_forward(None, 1, 1, 1, 0, ctypes.byref(g), 0, None) == -2    # null
    assert lib.dcnv3_b200_backward(1, 1, 1, 1, 1, 1, 1, None, 0, 0, ctypes.byref(g), 0, 5, None) == -1
    # empty batch: nothing to do, no device needed
    assert lib.dcnv3_b200_forward(None, None, None, None, 0, ctypes.byref(_geo(N=0)), 0, None) == 0
    # fp32 accumulators (rounded up to 256 B) + 256 B for the kernel-family selector word
    assert lib.dcnv3_b200_backward_workspace_bytes(2, ctypes.byref(g), 0) == 2 * 8 * 8 * 64 * 4 + 256
    assert lib.dcnv3_b200_backward_workspace_bytes(2, ctypes.byref(g), 1) == 0
    assert lib.dcnv3_b200_backward_workspace_bytes(0, ctypes.byref(g), 0) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    """Without a device the compute entry points refuse (EDEVICE): nothing runs on the host."""
    g = _geo()
    buf = (ctypes.c_float * 16)()
    p = ctypes.addressof(buf)
    assert lib.dcnv3_b200_forward(p, p, p, p, 0, ctypes.byref(g), 0, None) == -6
    assert b"no CPU path" in lib.dcnv3_b200_last_error()
    assert lib.dcnv3_b200_debug_indices(p, p, p, 0, ctypes.byref(g), None) == -6


def test_python_front_refuses_cpu_tensors():
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function, DCNv3SoftmaxFunction, dcnv3_debug_indices
    x, off, m = torch.randn(1, 4, 4, 8), torch.zeros(1, 4, 4, 18), torch.zeros(1, 4, 4, 9)
    for f in (DCNv3Function, DCNv3SoftmaxFunction):
        with pytest.raises(NotImplementedError, match="Not implement on cpu"):
            f.apply(x, off, m, 3, 3, 1, 1, 1, 1, 1, 1, 1, 8, 1.0, 256)
    with pytest.raises(NotImplementedError):
        dcnv3_debug_indices(off, 4, 4, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1.0)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from yolo_dual_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no fallback"):
        _lib.load()


def test_module_api_surface_matches_reference():
    """Constructor defaults, attribute names and state_dict keys of the reference module
    (modules/dcnv3.py:51-54,88-96)."""
    import inspect
    from yolo_dual_b200.ops_dcnv3.modules import DCNv3
    sig = inspect.signature(DCNv3.__init__)
    names = list(sig.parameters)[1:10]
    assert names == ["channels", "kernel_size", "stride", "pad", "dilation", "group", "offset_scale",
                     "act_layer", "norm_layer"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["channels"], d["kernel_size"], d["stride"], d["pad"], d["dilation"], d["group"],
            d["offset_scale"]) == (64, 3, 1, 1, 1, 4, 1.0)
    m = DCNv3(channels=32, group=2, dilation=3)
    assert m.dilation == 1 and m.group_channels == 16
    keys = set(m.state_dict())
    for k in ("dw_conv.conv.weight", "dw_conv.bn.weight", "dw_conv.bn.running_mean", "offset.weight",
              "offset.bias", "mask.weight", "mask.bias", "input_proj.weight", "output_proj.bias"):
        assert k in keys
    assert m.offset.weight.shape == (2 * 9 * 2, 32) and m.mask.weight.shape == (2 * 9, 32)
    assert float(m.offset.weight.abs().sum() + m.mask.weight.abs().sum()) == 0.0
    with pytest.raises(ValueError):
        DCNv3(channels=30, group=4)


def test_function_signature_matches_reference():
    import inspect
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    want = ["ctx", "input", "offset", "mask", "kernel_h", "kernel_w", "stride_h", "stride_w", "pad_h",
            "pad_w", "dilation_h", "dilation_w", "group", "group_channels", "offset_scale", "im2col_step"]
    fwd = DCNv3Function.forward
    fwd = getattr(fwd, "__wrapped__", fwd)
    assert list(inspect.signature(fwd).parameters) == want
    assert list(inspect.signature(DCNv3Function.symbolic).parameters) == ["g"] + want[1:]


def test_segloss_library_exports_its_header():
    from yolo_dual_b200 import _segloss
    from yolo_dual_b200.build import build_segloss
    build_segloss()
    lib = _segloss.load()
    hdr = open(os.path.join(ROOT, "include", "segloss_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(segloss_b200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_segloss.SYMBOLS)
    v = int(re.search(r"#define SEGLOSS_B200_VERSION (\d+)", hdr).group(1))
    assert lib.segloss_b200_version() == v
    # argument validation happens before any device is touched; no CPU path behind it
    assert lib.segloss_b200_forward(None, None, None, None, 1, 12, 4, 4, 1, None) == -2
    assert lib.segloss_b200_forward(1, 1, 1, 1, 1, 17, 4, 4, 1, None) == -1
    assert lib.segloss_b200_forward(1, 1, 1, 1, 1, 12, 4, 4, 1, None) == -3
    assert b"no CPU path" in lib.segloss_b200_last_error()
    with pytest.raises(NotImplementedError):
        _segloss.FusedSegLoss.apply(torch.zeros(1, 12, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long),
                                    torch.ones(12), 1)


def test_bnact_library_exports_its_header():
    from yolo_dual_b200 import _bnact
    from yolo_dual_b200.build import build_bnact
    build_bnact()
    lib = _bnact.load()
    hdr = open(os.path.join(ROOT, "include", "bnact_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bnact_b200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_bnact.SYMBOLS)
    assert lib.bnact_b200_version() == int(re.search(r"#define BNACT_B200_VERSION (\d+)", hdr).group(1))
    assert [lib.bnact_b200_supported(2, c) for c in (8, 64, 1024, 2048, 4096, 12, 24, 0)] == [1, 1, 1, 1, 0, 0, 0, 0]
    assert [lib.bnact_b200_supported(0, c) for c in (4, 64, 1024, 2048, 6)] == [1, 1, 1, 0, 0]
    assert lib.bnact_b200_partial_floats(2, 16 * 80 * 80, 128) % (2 * 128) == 0
    assert lib.bnact_b200_forward(None, None, None, None, None, None, None, None, 2, 100, 64, 1e-3, 0.03, 1, None) == -2
    assert lib.bnact_b200_forward(1, 1, 1, 1, None, None, 1, 1, 2, 100, 12, 1e-3, 0.03, 1, None) == -1
    assert lib.bnact_b200_forward(1, 1, 1, 1, None, None, 1, 1, 2, 100, 64, 1e-3, 0.03, 1, None) == -3
    # on a CPU tensor the Conv block never reaches the fused call
    from yolo_dual_b200.ops_dcnv3.modules.conv import Conv
    m = Conv(8, 16, 1).train()
    assert not _bnact.usable(torch.zeros(2, 16, 4, 4), m.bn, m.act)
    assert m(torch.zeros(2, 8, 4, 4)).shape == (2, 16, 4, 4)


def test_resize_library_exports_its_header():
    from yolo_dual_b200 import _resize
    from yolo_dual_b200.build import build_resize
    build_resize()
    lib = _resize.load()
    hdr = open(os.path.join(ROOT, "include", "resize_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(resize_b200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_resize.SYMBOLS)
    assert lib.resize_b200_version() == int(re.search(r"#define RESIZE_B200_VERSION (\d+)", hdr).group(1))
    assert lib.resize_b200_supported(2, 64, 10, 10, 20, 20, 0) == 1
    assert lib.resize_b200_supported(2, 64, 10, 10, 25, 20, 0) == 0          # nearest: integer factors only
    assert lib.resize_b200_supported(2, 64, 10, 10, 25, 7, 1) == 1
    assert lib.resize_b200_supported(2, 12, 10, 10, 20, 20, 1) == 0          # 12 bf16 channels: not whole vectors
    assert lib.resize_b200_forward(None, None, 2, 1, 4, 4, 8, 8, 8, 0, None) == -2
    assert lib.resize_b200_forward(1, 1, 2, 1, 4, 4, 8, 8, 8, 0, None) == -3
    assert not _resize.usable(torch.zeros(1, 8, 4, 4), (8, 8), "nearest")   # CPU tensors stay with torch
