"""One site, default dispatch, a few backward calls (for ncu launch lists)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.dcnv3_oracle import make_inputs
from yolo_dual_b200 import _lib
lib = _lib.load()
N, H, W, G, gc = 16, 80, 80, 8, 16
geo = _lib.Geometry(N, H, W, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
x, off, m, go = (t.to("cuda", torch.bfloat16).contiguous() for t in make_inputs(N, H, W, G, gc, dist="unit", seed=0))
gi, goff, gm = (torch.empty_like(t) for t in (x, off, m))
wsb = lib.dcnv3_b200_backward_workspace_bytes(_lib.BF16, ctypes.byref(geo), _lib.ACC_OPMATH)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    rc = lib.dcnv3_b200_backward(x.data_ptr(), off.data_ptr(), m.data_ptr(), go.data_ptr(), gi.data_ptr(), goff.data_ptr(),
                                 gm.data_ptr(), ws.data_ptr(), wsb, _lib.BF16, ctypes.byref(geo), 0, _lib.ACC_OPMATH, st)
    assert rc == 0
torch.cuda.synchronize()
print("ok")
