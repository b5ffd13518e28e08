"""Quick GPU check of the interpolation-matrix kernel family against the vector family and the
CPU pixel oracle (test infrastructure), plus CUDA-event timings at the BASELINE sites.

    python tools/imat_check.py [parity] [bwd] [time]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle.dcnv3_oracle import PixelOracle, make_inputs  # noqa: E402
from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function, DCNv3SoftmaxFunction  # noqa: E402
from yolo_dual_b200 import _lib as _lib_mod  # noqa: E402

DEV = "cuda:0"
DIST = os.environ.get("IMAT_DIST", "unit")
OFFSCALE = float(os.environ.get("IMAT_OFFSCALE", "1.0"))


def run(fn, x, off, m, go, args, dtype, env):
    for k in ("DCNV3_B200_FWD", "DCNV3_B200_BWD"):
        os.environ.pop(k, None)
    os.environ.update(env)
    _lib_mod.reload_knobs()  # the library caches its knobs
    xs, os_, ms = (t.to(DEV, dtype).contiguous().requires_grad_(True) for t in (x, off, m))
    out = fn.apply(xs, os_, ms, *args, 256)
    res = [out.detach().float().cpu()]
    if go is not None:
        out.backward(go.to(DEV, dtype))
        res += [xs.grad.float().cpu(), os_.grad.float().cpu(), ms.grad.float().cpu()]
    torch.cuda.synchronize()
    return res


def err(a, b):
    s = max(1.0, float(b.abs().max()))
    return float((a - b).abs().max()) / s


CASES = {
    "cfg1": ((2, 80, 80, 4, 16), dict()),
    "odd": ((2, 21, 19, 8, 16), dict()),
    "pad0": ((1, 12, 12, 4, 16), dict(pad=0)),
    "scale1.5": ((1, 17, 23, 4, 16), dict(scale=1.5)),
    "tiny": ((1, 3, 5, 4, 16), dict()),
    "G32": ((1, 20, 20, 32, 16), dict()),
}


def parity(do_bwd):
    po = PixelOracle()
    bad = 0
    for name, ((N, H, W, G, gc), kw) in CASES.items():
        pad, scale = kw.get("pad", 1), kw.get("scale", 1.0)
        args = (3, 3, 1, 1, pad, pad, 1, 1, G, gc, scale)
        for dist in ("unit", "ref"):
            for dtype in (torch.bfloat16, torch.float16):
                x, off, m, go = make_inputs(N, H, W, G, gc, 3, 3, 1, 1, pad, pad, 1, 1, dist=dist, seed=5)
                xr, offr, mr, gor = (t.to(dtype).float() for t in (x, off, m, go))
                want = [po.forward(xr, offr, mr, *args)]
                if do_bwd:
                    want += list(po.backward(xr, offr, mr, gor, *args))
                env = {"DCNV3_B200_FWD": "imat"}
                if do_bwd:
                    env["DCNV3_B200_BWD"] = "imat"
                got = run(DCNv3Function, x, off, m, go if do_bwd else None, args, dtype, env)
                ref = run(DCNv3Function, x, off, m, go if do_bwd else None, args, dtype,
                          {"DCNV3_B200_FWD": "vec", "DCNV3_B200_BWD": "vec"})
                e_imat = [err(a, b) for a, b in zip(got, want)]
                e_vec = [err(a, b) for a, b in zip(ref, want)]
                ok = all(e < 4e-3 for e in e_imat)
                bad += not ok
                print(f"{name:9s} {dist:4s} {str(dtype)[6:]:8s} imat " + " ".join(f"{e:.2e}" for e in e_imat) +
                      "   vec " + " ".join(f"{e:.2e}" for e in e_vec) + ("" if ok else "   <-- FAIL"))
    # fused softmax
    (N, H, W, G, gc) = (2, 21, 19, 8, 16)
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    x, off, _, go = make_inputs(N, H, W, G, gc, dist="unit", seed=4)
    logits = torch.randn(N, H, W, G * 9, generator=torch.Generator().manual_seed(8)) * 2
    for dtype in (torch.bfloat16, torch.float16):
        env = {"DCNV3_B200_FWD": "imat"}
        if do_bwd:
            env["DCNV3_B200_BWD"] = "imat"
        got = run(DCNv3SoftmaxFunction, x, off, logits, go if do_bwd else None, args, dtype, env)
        ref = run(DCNv3SoftmaxFunction, x, off, logits, go if do_bwd else None, args, dtype,
                  {"DCNV3_B200_FWD": "vec", "DCNV3_B200_BWD": "vec"})
        e = [err(a, b) for a, b in zip(got, ref)]
        ok = all(v < 8e-3 for v in e)
        bad += not ok
        print(f"softmax   unit {str(dtype)[6:]:8s} imat-vs-vec " + " ".join(f"{v:.2e}" for v in e) + ("" if ok else "   <-- FAIL"))
    print("PARITY", "FAIL" if bad else "OK", bad)
    return bad


def timing(do_bwd):
    import ctypes
    from yolo_dual_b200 import _lib
    lib = _lib.load()
    sites = {"P3": (16, 80, 80, 8, 16), "P4": (16, 40, 40, 16, 16), "P5": (16, 20, 20, 32, 16)}
    for name, (N, H, W, G, gc) in sites.items():
        geo = _lib.Geometry(N, H, W, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
        sets = []
        for s in range(4):
            x, off, m, go = make_inputs(N, H, W, G, gc, dist=DIST, seed=s)
            if OFFSCALE != 1.0:
                off = off * OFFSCALE
            sets.append([t.to(DEV, torch.bfloat16).contiguous() for t in (x, off, m, go)])
        out = torch.empty_like(sets[0][0])
        gi, goff, gm = (torch.empty_like(sets[0][k]) for k in (0, 1, 2))
        wsb = lib.dcnv3_b200_backward_workspace_bytes(_lib.BF16, ctypes.byref(geo), _lib.ACC_OPMATH)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=DEV)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for fam in ("vec", "imat", "auto"):
            if fam == "auto":  # no knobs: the library's own choice (device-side selector for the backward)
                os.environ.pop("DCNV3_B200_FWD", None)
                os.environ.pop("DCNV3_B200_BWD", None)
            else:
                os.environ["DCNV3_B200_FWD"] = fam
                os.environ["DCNV3_B200_BWD"] = fam
            _lib_mod.reload_knobs()
            res = {}
            for what in (("fwd", "bwd") if do_bwd else ("fwd",)):
                def step(i):
                    x, off, m, go = sets[i % 4]
                    if what == "fwd":
                        rc = lib.dcnv3_b200_forward(x.data_ptr(), off.data_ptr(), m.data_ptr(), out.data_ptr(),
                                                    _lib.BF16, ctypes.byref(geo), 0, st)
                    else:
                        rc = lib.dcnv3_b200_backward(x.data_ptr(), off.data_ptr(), m.data_ptr(), go.data_ptr(),
                                                     gi.data_ptr(), goff.data_ptr(), gm.data_ptr(), ws.data_ptr(),
                                                     wsb, _lib.BF16, ctypes.byref(geo), 0, _lib.ACC_OPMATH, st)
                    assert rc == 0, rc
                for i in range(10):
                    step(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                for i in range(200):
                    step(i)
                e1.record()
                torch.cuda.synchronize()
                res[what] = e0.elapsed_time(e1) / 200 * 1e3
            print(f"{name} {fam:5s} " + " ".join(f"{k} {v:8.1f} us" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    do_bwd = "bwd" in a
    rc = 0
    if "parity" in a:
        rc = parity(do_bwd)
    if "time" in a:
        timing(do_bwd)
    sys.exit(1 if rc else 0)
