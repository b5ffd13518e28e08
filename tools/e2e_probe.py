"""e2e pipeline probe: period per step for several pipeline depths, and the H2D / kernel / D2H spans of
one step (CUDA events on the three streams).  python tools/e2e_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from yolo_dual_b200.host import HostPipeline, HostSite

dev = torch.device("cuda:0")
wl = bench.Workload(dev, torch.bfloat16, ["P3", "P4", "P5"], "opmath", False)
sites = []
for b in wl.sets[0]:
    N, H, W, G, gc = bench.SITES[b.name]
    hs = HostSite(*(t.cpu().pin_memory() for t in (b.input, b.offset, b.mask, b.grad_out)),
                  args=(3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0))
    sites.append(hs.alloc_outputs(tuple(b.output.shape)))
for depth in (2, 3, 4, 2, 3):
    pipe = HostPipeline(dev, depth=depth)
    for _ in range(4):
        pipe.submit(sites)
    pipe.drain(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    host = 0.0
    for _ in range(40):
        h0 = time.perf_counter(); pipe.submit(sites); host += time.perf_counter() - h0
    pipe.drain(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / 40
    print(f"depth {depth}: {ms:.3f} ms per step (host time inside submit {host * 1e3 / 40:.3f} ms)")

# ---- spans of the three phases in steady state (depth 3): monkey-patch events around the phases
pipe = HostPipeline(dev, depth=3)
marks = []
orig_submit = pipe.submit
def timed_submit(s):
    a = torch.cuda.Event(enable_timing=True); a.record(pipe.s_in)
    k = orig_submit(s)
    slot = pipe.slots[k % pipe.depth]
    marks.append((a, slot["ev_in"], slot["ev_done"], slot["ev_out"]))
    return k
for sl in pipe.slots:
    for key in ("ev_in", "ev_done", "ev_out"):
        sl[key] = torch.cuda.Event(enable_timing=True)
# fresh events per step so that elapsed_time is well defined
def submit_fresh(s):
    k = pipe.step
    sl = pipe.slots[k % pipe.depth]
    if sl["busy"]:
        sl["ev_out"].synchronize()
    old = (sl["ev_in"], sl["ev_done"], sl["ev_out"])
    for key in ("ev_in", "ev_out"):
        sl[key] = torch.cuda.Event(enable_timing=True)
    # ev_done is waited on by s_in for the slot's previous use: keep ordering by waiting here
    pipe.s_in.wait_event(old[1])
    sl["ev_done"] = torch.cuda.Event(enable_timing=True)
    sl["busy"] = False
    return timed_submit(s)
for _ in range(12):
    submit_fresh(sites)
pipe.drain(); torch.cuda.synchronize()
base = marks[4][0]
for i in range(4, 10):
    a, e_in, e_done, e_out = marks[i]
    print(f"step {i}: h2d starts {base.elapsed_time(a):7.3f}  h2d {a.elapsed_time(e_in):6.3f} ms  "
          f"kernels end +{e_in.elapsed_time(e_done):6.3f}  d2h end +{e_done.elapsed_time(e_out):6.3f}")
