"""Where does the C3-DCNV3 seg training step spend its time?  Wall per step with / without the per-step loss
read-back, and the sum of GPU kernel time from the torch profiler.  python tools/seg_probe.py [model]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yolo_dual_b200 import seg

import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    if rank:
        sys.stdout = open(os.devnull, "w")
name = sys.argv[1] if len(sys.argv) > 1 else "yolov5seg"
cfg = {"yolov5seg": seg.YOLOV5_SEG, "yolov8seg": seg.YOLOV8_SEG}[name]
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
model = seg.SegModel(cfg, dcn="dcnv3").to(dev).to(memory_format=torch.channels_last)
crit = seg.SegmentationLoss(12, class_weights=seg.CAMVID_CLASS_WEIGHTS).to(dev)
net = model
model = seg.wrap_ddp(model, dev)
opt = seg.smart_optimizer(model)
model.train()
imgs = torch.randn(16, 3, 640, 640, device=dev).contiguous(memory_format=torch.channels_last)
lab = torch.randint(0, 12, (16, 640, 640), device=dev)

def step(sync):
    loss, _ = seg.train_step(model, crit, opt, imgs, lab, autocast_dtype=torch.bfloat16)
    return float(loss) if sync else loss

for _ in range(5):
    step(True)

def phases():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        pred, scale = model(imgs, lowres=True)
    ev[1].record()
    loss, _ = crit(pred, lab, scale)
    ev[2].record()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    ev[3].record()
    opt.step()
    ev[4].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(4)], pred

ts = [phases()[0] for _ in range(5)][-1]
print("phases ms: model fwd %.2f, loss fwd %.2f, backward %.2f, optimizer %.2f" % tuple(ts))
pred = phases()[1]
print("pred", pred.dtype, tuple(pred.shape), pred.stride(), "deferred upsample:", net._deferred, "world", world)
with torch.autocast("cuda", dtype=torch.bfloat16):
    t = torch.randn(2, 8, 4, 4, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    print("autocast dtypes: nearest", torch.nn.Upsample(scale_factor=2.0)(t).dtype,
          "bilinear", torch.nn.functional.interpolate(t, size=(8, 8), mode="bilinear").dtype)
for sync in (True, False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        step(sync)
    torch.cuda.synchronize()
    print(f"sync={sync}: {(time.perf_counter() - t0) * 100:.2f} ms per step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step(True)
    torch.cuda.synchronize()
from torch.autograd import DeviceType
ev = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
tot = sum(e.self_device_time_total for e in ev) / 3 / 1e3
print(f"GPU kernel time per step {tot:.2f} ms; kernels per step {sum(e.count for e in ev) / 3:.0f}")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:40]:
    print(f"{e.self_device_time_total / 3 / 1e3:8.3f} ms  x{e.count / 3:6.1f}  {e.key[:150]}")
