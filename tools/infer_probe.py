"""Kernel-time breakdown of the configs[4] inference step (1280x1280, fp16, fused softmax, 4 images).  python tools/infer_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_dual_b200 import seg
dev = "cuda:0"
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = seg.SegModel(seg.YOLOV5_SEG, dcn="dcnv3", fused_softmax=True, img_size=(1280, 1280)).to(dev).half().eval().to(memory_format=torch.channels_last)
x = torch.randn(4, 3, 1280, 1280, device=dev).half().contiguous(memory_format=torch.channels_last)
def step():
    with torch.no_grad():
        return model(x).argmax(1).to(torch.uint8)
for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / 10:.2f} ms per step (4 images)")
from torch.profiler import profile, ProfilerActivity
from torch.autograd import DeviceType
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
tot = sum(e.self_device_time_total for e in ev) / 3 / 1e3
print(f"GPU kernel time per step {tot:.2f} ms; kernels per step {sum(e.count for e in ev) / 3:.0f}")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:22]:
    print(f"{e.self_device_time_total / 3 / 1e3:8.3f} ms  x{e.count / 3:6.1f}  {e.key[:140]}")
