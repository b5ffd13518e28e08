#!/usr/bin/env python
"""BASELINE config #3: DCNv3 op shape sweep vs the HBM roofline on one B200.

    python tools/sweep.py > profiles/rNN_sweep.md

C in {64,128,256,512} x G in {4,8,16,32} x H=W in {160,80,40,20} x dtype in {fp32,fp16,bf16}, N=16,
3x3 s1 p1, unit-scale inputs (offset sigma 1 px), 10 warm-up + 50 timed launches per op with CUDA events,
inputs rotating over enough buffer sets to exceed L2.  Shapes whose group_channels is below one
16-byte vector run the generic kernels and are marked (generic).
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yolo_dual_b200 import _lib  # noqa: E402


def config5():
    """BASELINE config #5: high-res seg inference, 1280x1280, batch 32 fp16 sharded over 8 GPUs = 4 images per
    GPU; forward only at the three DCN sites (160^2x128, 80^2x256, 40^2x512, group_channels 16), with the mask
    softmax fused into the kernel (logits in) and, for comparison, with probabilities in."""
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    N, P, e = 4, 9, 2
    print("| site | N | HxW | C | G | mask input | fwd us | GB/s | % of HBM peak |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, HW, C in (("P3", 160, 128), ("P4", 80, 256), ("P5", 40, 512)):
        G, gc = C // 16, 16
        geo = _lib.Geometry(N, HW, HW, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
        by = e * N * HW * HW * (2 * C + 3 * G * P)
        nset = max(2, int(300e6 // by) + 1)
        g = torch.Generator(device=dev).manual_seed(HW)
        sets = []
        for _ in range(nset):
            x = torch.randn(N, HW, HW, C, device=dev, generator=g).half()
            off = torch.randn(N, HW, HW, G * P * 2, device=dev, generator=g).half()
            lg = torch.randn(N, HW, HW, G, P, device=dev, generator=g)
            sets.append((x, off, lg.reshape(N, HW, HW, G * P).half().contiguous(),
                         torch.softmax(lg, -1).reshape(N, HW, HW, G * P).half().contiguous(), torch.empty_like(x)))
        for logits in (1, 0):
            def fwd(k):
                x, off, lgt, pr, out = sets[k % nset]
                _lib.check(lib.dcnv3_b200_forward(x.data_ptr(), off.data_ptr(), (lgt if logits else pr).data_ptr(),
                                                  out.data_ptr(), _lib.F16, ctypes.byref(geo), logits, st), "fwd")
            for k in range(10):
                fwd(k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for k in range(100):
                fwd(k)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 100 * 1e3
            gb = by / us / 1e3
            print(f"| {name} | {N} | {HW}x{HW} | {C} | {G} | {'logits (fused softmax)' if logits else 'probabilities'} | "
                  f"{us:.1f} | {gb:.0f} | {100 * gb / peak:.1f} |", flush=True)


def main():
    if "--config5" in sys.argv:
        return config5()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    N, P = 16, 9
    print(f"| C | G | gc | HxW | dtype | path | fwd us | fwd GB/s | fwd % | bwd us | bwd GB/s | bwd % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for dtype, dt in ((torch.float32, _lib.F32), (torch.float16, _lib.F16), (torch.bfloat16, _lib.BF16)):
        e = torch.empty((), dtype=dtype).element_size()
        for HW in (160, 80, 40, 20):
            for C in (64, 128, 256, 512):
                for G in (4, 8, 16, 32):
                    gc = C // G
                    if HW == 160 and C >= 256:
                        continue  # > 2 GB of rotating buffers; not a shape of any BASELINE model
                    geo = _lib.Geometry(N, HW, HW, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
                    fwd_b = e * N * HW * HW * (2 * C + 3 * G * P)
                    bwd_b = e * N * HW * HW * (3 * C + 6 * G * P)
                    nset = max(2, min(6, int(300e6 // (fwd_b + bwd_b)) + 1))
                    g = torch.Generator(device=dev).manual_seed(C * 1000 + G * 10 + HW)
                    sets = []
                    for _ in range(nset):
                        x = torch.randn(N, HW, HW, C, device=dev, generator=g).to(dtype)
                        off = torch.randn(N, HW, HW, G * P * 2, device=dev, generator=g).to(dtype)
                        m = torch.softmax(torch.randn(N, HW, HW, G, P, device=dev, generator=g), -1).reshape(
                            N, HW, HW, G * P).to(dtype).contiguous()
                        go = torch.randn(N, HW, HW, C, device=dev, generator=g).to(dtype)
                        out, gi, goff, gm = (torch.empty_like(t) for t in (x, x, off, m))
                        wsb = lib.dcnv3_b200_backward_workspace_bytes(dt, ctypes.byref(geo), _lib.ACC_TILE)
                        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
                        sets.append((x, off, m, go, out, gi, goff, gm, ws, wsb))

                    def fwd(k):
                        x, off, m, go, out, gi, goff, gm, ws, wsb = sets[k % nset]
                        _lib.check(lib.dcnv3_b200_forward(x.data_ptr(), off.data_ptr(), m.data_ptr(), out.data_ptr(),
                                                          dt, ctypes.byref(geo), 0, st), "fwd")

                    def bwd(k):
                        x, off, m, go, out, gi, goff, gm, ws, wsb = sets[k % nset]
                        _lib.check(lib.dcnv3_b200_backward(x.data_ptr(), off.data_ptr(), m.data_ptr(), go.data_ptr(),
                                                           gi.data_ptr(), goff.data_ptr(), gm.data_ptr(), ws.data_ptr(),
                                                           wsb, dt, ctypes.byref(geo), 0, _lib.ACC_TILE, st), "bwd")

                    res = []
                    for f in (fwd, bwd):
                        for k in range(10):
                            f(k)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize()
                        e0.record()
                        for k in range(50):
                            f(k)
                        e1.record()
                        torch.cuda.synchronize()
                        res.append(e0.elapsed_time(e1) / 50 * 1e3)
                    vec = gc % (16 // e) == 0 and (gc // (16 // e)) & (gc // (16 // e) - 1) == 0 and gc // (16 // e) <= 32
                    name = {torch.float32: "fp32", torch.float16: "fp16", torch.bfloat16: "bf16"}[dtype]
                    fg, bg = fwd_b / res[0] / 1e3, bwd_b / res[1] / 1e3
                    win = e == 2 and gc == 16 and G % 4 == 0  # staged-window forward + window backward (grad_accum 'tile')
                    print(f"| {C} | {G} | {gc} | {HW}x{HW} | {name} | {'win' if win else 'vec' if vec else '(generic)'} | {res[0]:.1f} | "
                          f"{fg:.0f} | {100 * fg / peak:.1f} | {res[1]:.1f} | {bg:.0f} | {100 * bg / peak:.1f} |", flush=True)
                    del sets
                    torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
