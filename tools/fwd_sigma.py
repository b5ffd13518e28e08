"""Forward time at site P3 for several offset spreads and forward families.  python tools/fwd_sigma.py"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.dcnv3_oracle import make_inputs
from yolo_dual_b200 import _lib
lib = _lib.load()
N, H, W, G, gc = 16, 80, 80, 8, 16
geo = _lib.Geometry(N, H, W, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for label, dist, sc in (("s0", "unit", 0.0), ("s1", "unit", 1.0), ("s1.1", "unit", 1.1), ("s1.2", "unit", 1.2), ("s1.3", "unit", 1.3), ("s1.5", "unit", 1.5), ("s1.7", "unit", 1.7),
                        ("s2", "unit", 2.0), ("s3", "unit", 3.0), ("ref", "ref", 1.0)):
    sets = []
    for s in range(4):
        x, off, m, go = make_inputs(N, H, W, G, gc, dist=dist, seed=s)
        sets.append([t.to("cuda", torch.bfloat16).contiguous() for t in (x, off * sc, m)])
    out = torch.empty_like(sets[0][0])
    res = []
    for fam in ("vec", "win"):
        os.environ["DCNV3_B200_FWD"] = fam
        _lib.reload_knobs()
        def step(i):
            x, off, m = sets[i % 4]
            assert lib.dcnv3_b200_forward(x.data_ptr(), off.data_ptr(), m.data_ptr(), out.data_ptr(), _lib.BF16, ctypes.byref(geo), 0, st) == 0
        for i in range(10): step(i)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
        for i in range(200): step(i)
        e1.record(); torch.cuda.synchronize(); res.append(e0.elapsed_time(e1) / 200 * 1e3)
    print(f"{label:5s} vec {res[0]:7.1f} us   win {res[1]:7.1f} us")
