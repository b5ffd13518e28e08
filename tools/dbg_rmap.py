"""One small 16-bit backward through the C-ABI, compared with the pixel oracle (debugging the TMA-reduce flush).  python tools/dbg_rmap.py [fp16|bf16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.dcnv3_oracle import PixelOracle, make_inputs  # checker only
from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function, set_grad_accum
set_grad_accum("tile")
dtype = torch.float16 if (sys.argv[1:] or ["bf16"])[0] == "fp16" else torch.bfloat16
N, H, W, G, gc = 2, 24, 32, 4, 16
args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
x, off, m, go = make_inputs(N, H, W, G, gc, 3, 3, 1, 1, 1, 1, 1, 1, dist="unit", seed=1)
xs, os_, ms = (t.cuda().to(dtype).requires_grad_(True) for t in (x, off, m))
out = DCNv3Function.apply(xs, os_, ms, *args, 256)
out.backward(go.cuda().to(dtype))
torch.cuda.synchronize()
f = lambda t: t.to(dtype).float()
gi, _, _ = PixelOracle().backward(f(x), f(off), f(m), f(go), *args)
err = (xs.grad.float().cpu() - gi).abs().max().item() / gi.abs().max().item()
print("ok: grad_input max err / max|ref| =", err)
