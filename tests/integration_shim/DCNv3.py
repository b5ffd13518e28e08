"""The reference-side binding of INTEGRATION.md §3, kept as a real file so the tests exercise it: a module
named `DCNv3` with the two entry points the reference's functions/dcnv3_func.py calls (:39, :54), backed
by libdcnv3_b200.so through ctypes.  Drop it next to the reference's ops_dcnv3 package (or anywhere on
sys.path) and the reference's own dcnv3_func.py / dcnv3.py / test.py run unmodified."""
import ctypes, torch

import os
_lib = ctypes.CDLL(os.environ.get("DCNV3_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..",
                                                                   "yolo_dual_b200", "csrc", "libdcnv3_b200.so"))

class _Geo(ctypes.Structure):              # struct dcnv3_b200_geometry (include/dcnv3_b200.h)
    _fields_ = [(n, ctypes.c_int) for n in (
        "N", "H", "W", "kernel_h", "kernel_w", "stride_h", "stride_w", "pad_h", "pad_w",
        "dilation_h", "dilation_w", "group", "group_channels")] + [("offset_scale", ctypes.c_float)]

_lib.dcnv3_b200_last_error.restype = ctypes.c_char_p
_lib.dcnv3_b200_backward_workspace_bytes.restype = ctypes.c_size_t
_vp, _gp = ctypes.c_void_p, ctypes.POINTER(_Geo)
_lib.dcnv3_b200_output_size.argtypes = [_gp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
_lib.dcnv3_b200_forward.argtypes = [_vp] * 4 + [ctypes.c_int, _gp, ctypes.c_int, _vp]
_lib.dcnv3_b200_backward_workspace_bytes.argtypes = [ctypes.c_int, _gp, ctypes.c_int]
_lib.dcnv3_b200_backward.argtypes = [_vp] * 8 + [ctypes.c_size_t, ctypes.c_int, _gp, ctypes.c_int, ctypes.c_int, _vp]
_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.float64: 3}

def _geo(x, kh, kw, sh, sw, ph, pw, dh, dw, g, gc, s):
    return _Geo(x.shape[0], x.shape[1], x.shape[2], kh, kw, sh, sw, ph, pw, dh, dw, g, gc, s)

def _ok(rc):
    if rc: raise RuntimeError(_lib.dcnv3_b200_last_error().decode())

def dcnv3_forward(input, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, group, group_channels,
                  offset_scale, im2col_step):
    geo = _geo(input, kh, kw, sh, sw, ph, pw, dh, dw, group, group_channels, offset_scale)
    ho, wo = ctypes.c_int(), ctypes.c_int()
    _ok(_lib.dcnv3_b200_output_size(ctypes.byref(geo), ctypes.byref(ho), ctypes.byref(wo)))
    with torch.cuda.device_of(input):
        out = input.new_empty((input.shape[0], ho.value, wo.value, input.shape[3]))
        _ok(_lib.dcnv3_b200_forward(input.data_ptr(), offset.data_ptr(), mask.data_ptr(), out.data_ptr(),
                                    _DT[input.dtype], ctypes.byref(geo), 0,
                                    torch.cuda.current_stream().cuda_stream))
    return out

def dcnv3_backward(input, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, group, group_channels,
                   offset_scale, grad_output, im2col_step):
    geo = _geo(input, kh, kw, sh, sw, ph, pw, dh, dw, group, group_channels, offset_scale)
    dt = _DT[input.dtype]
    with torch.cuda.device_of(input):
        gi, go, gm = torch.empty_like(input), torch.empty_like(offset), torch.empty_like(mask)
        n = _lib.dcnv3_b200_backward_workspace_bytes(dt, ctypes.byref(geo), 0)
        ws = torch.empty(max(n, 1), dtype=torch.uint8, device=input.device)
        _ok(_lib.dcnv3_b200_backward(input.data_ptr(), offset.data_ptr(), mask.data_ptr(),
                                     grad_output.data_ptr(), gi.data_ptr(), go.data_ptr(), gm.data_ptr(),
                                     ws.data_ptr(), n, dt, ctypes.byref(geo), 0, 0,
                                     torch.cuda.current_stream().cuda_stream))
    return gi, go, gm
