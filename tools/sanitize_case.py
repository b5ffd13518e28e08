#!/usr/bin/env python
"""Small driver for compute-sanitizer runs (one tool per gpurun call, see B200_PROFILING.md):

    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py tile

Runs forward + backward of every kernel family on small shapes (vector 16/32 B per lane, generic,
fused softmax, privatised tile backward) and checks the results against the CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle.dcnv3_oracle import PixelOracle, make_inputs  # noqa: E402  (checker only)
from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function, DCNv3SoftmaxFunction  # noqa: E402


def run(shape, dtype, fn=DCNv3Function):
    from yolo_dual_b200 import _lib
    _lib.reload_knobs()  # the library caches its DCNV3_B200_* knobs; main() switches them between runs
    N, H, W, G, gc, k, s, pad = shape
    args = (k, k, s, s, pad, pad, 1, 1, G, gc, 1.0)
    x, off, m, go = make_inputs(N, H, W, G, gc, k, k, s, s, pad, pad, 1, 1, dist="unit", seed=1)
    xs, os_, ms = (t.cuda().to(dtype).requires_grad_(True) for t in (x, off, m))
    out = fn.apply(xs, os_, ms, *args, 256)
    out.backward(go.cuda().to(dtype))
    torch.cuda.synchronize()
    if fn is DCNv3Function:
        po = PixelOracle()
        f = lambda t: t.to(dtype).float()
        want = po.forward(f(x), f(off), f(m), *args)
        gi, _, _ = po.backward(f(x), f(off), f(m), f(go), *args)
        tol = dict(rtol=1e-5, atol=1e-4) if dtype == torch.float32 else dict(rtol=1e-2, atol=2e-2)
        torch.testing.assert_close(out.detach().float().cpu(), want, **tol)
        torch.testing.assert_close(xs.grad.float().cpu(), gi, **tol)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    shapes = [(2, 12, 10, 4, 16, 3, 1, 1), (1, 9, 11, 2, 16, 3, 2, 1), (1, 8, 8, 2, 30, 3, 1, 1), (1, 7, 7, 3, 8, 5, 1, 2)]
    if which in ("all", "vec"):
        for bpl in ("16", "32"):
            os.environ["DCNV3_B200_BPL"] = bpl
            for dt in (torch.float32, torch.bfloat16, torch.float16):
                for sh in shapes:
                    run(sh, dt)
                run(shapes[0], dt, DCNv3SoftmaxFunction)
        os.environ.pop("DCNV3_B200_BPL")
    if which in ("all", "tile"):
        os.environ["DCNV3_B200_BWD"] = "tile"
        for cfg in ("2,4", "0,1"):
            os.environ["DCNV3_B200_TILE"] = cfg
            for dt in (torch.float32, torch.bfloat16):
                run(shapes[0], dt)
                run(shapes[1], dt)
                run(shapes[0], dt, DCNv3SoftmaxFunction)
    print("sanitize_case ok:", which)


if __name__ == "__main__":
    main()
