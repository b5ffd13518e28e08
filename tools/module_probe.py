"""DCNv3 nn.Module forward + backward per C3-DCN site (batch 16, bf16 autocast), microseconds, CUDA events:
the reference's module structure (two Linear heads + F.softmax), fused softmax, fused softmax + packed heads
(one GEMM, kernels read its output in place).  VERDICT r1 item 8.   python tools/module_probe.py > profiles/r02_module_probe.md"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_dual_b200.ops_dcnv3.modules import DCNv3

dev = "cuda:0"
SITES = {"P3": (16, 80, 80, 128, 8), "P4": (16, 40, 40, 256, 16), "P5": (16, 20, 20, 512, 32)}
VARIANTS = {"two Linear + F.softmax (reference structure)": dict(),
            "fused softmax": dict(fused_softmax=True),
            "fused softmax + packed heads (one GEMM)": dict(fused_softmax=True, packed_heads=True)}
print("| site | variant | fwd+bwd us | kernels per call |\n|---|---|---|---|")
for s, (N, H, W, C, G) in SITES.items():
    for name, kw in VARIANTS.items():
        torch.manual_seed(0)
        m = DCNv3(channels=C, group=G, **kw).to(dev).train()
        with torch.no_grad():
            for lin in (m.offset, m.mask):
                lin.weight.normal_(0, 0.02); lin.bias.normal_(0, 0.5)
        x = torch.randn(N, H, W, C, device=dev, requires_grad=True)
        go = torch.randn(N, H, W, C, device=dev)
        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = m(x)
            y.backward(go)
            x.grad = None
            for p in m.parameters():
                p.grad = None
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        nk = sum(1 for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            step()
        e1.record(); torch.cuda.synchronize()
        print(f"| {s} | {name} | {e0.elapsed_time(e1) / 30 * 1e3:.0f} | {nk} |")
