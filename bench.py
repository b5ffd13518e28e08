#!/usr/bin/env python
"""bench.py — DCNv3 hot-path benchmark (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): the three DCNv3 sites of the YOLOv5-style C3-DCN seg model
at 640x640, batch 16, bf16 — P3 (16,80,80,128,G=8), P4 (16,40,40,256,G=16), P5 (16,20,20,512,G=32),
3x3 kernel, stride 1, pad 1, group_channels 16.  A step = forward of P3,P4,P5 then backward of
P5,P4,P3 (training order) through the C-ABI.  Metric = algorithmic GB/s of the whole step
(BASELINE.md §3 byte count: every tensor once in its storage dtype).

One JSON line on stdout (rank 0).  Everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

SITES = {  # name: (N, H, W, G, gc)
    "P3": (16, 80, 80, 8, 16),
    "P4": (16, 40, 40, 16, 16),
    "P5": (16, 20, 20, 32, 16),
}
KGEO = dict(kh=3, kw=3, sh=1, sw=1, ph=1, pw=1, dh=1, dw=1, scale=1.0)
N_BUFFER_SETS = 4  # inputs rotate over 4 sets (~1.7 GB) so every step starts L2-cold
METRIC = "dcnv3_fwd_bwd_algorithmic_GBps"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def esize(dtype):
    return torch.empty((), dtype=dtype).element_size()


def algo_bytes(site, e, N=None):
    """(fwd, bwd) algorithmic bytes of one site — BASELINE.md §3 / SURVEY §8(d)."""
    n, H, W, G, gc = SITES[site]
    n = N if N is not None else n
    C, P = G * gc, KGEO["kh"] * KGEO["kw"]
    Ho, Wo = H, W  # stride 1, same pad
    fwd = e * n * (H * W * C + Ho * Wo * (C + 3 * G * P))
    bwd = e * n * (H * W * 2 * C + Ho * Wo * (C + 6 * G * P))
    return fwd, bwd


def ncu_traffic(op):
    """dram__bytes_read.sum + dram__bytes_write.sum of the op's main kernel, per launch, from the
    committed ncu capture (profiles/r01_ncu_traffic.json); None when that op was not captured."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    try:
        with open(p) as f:
            t = json.load(f).get(op)
        return None if t is None else t["dram_bytes_read"] + t["dram_bytes_write"]
    except OSError:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed regions)
# ----------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,utilization.gpu,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        rows = []
        for ln in out.splitlines():
            f = [s.strip() for s in ln.split(",")]
            if len(f) >= 9:
                try:
                    rows.append((float(f[0]), float(f[1]), float(f[2]), f[3], f[4:]))
                except ValueError:
                    pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [r for r in rows if r[2] >= 50] or rows
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in busy:
            for nm, v in zip(names, r[4][1:5]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(r[0] for r in busy), "sm_max_mhz": rows[0][1],
                "reasons": sorted(reasons), "samples": len(rows), "samples_under_load": len(busy)}


# ----------------------------------------------------------------------------------------------
# device-side workload, driven straight through the C-ABI
# ----------------------------------------------------------------------------------------------
class SiteBuffers:
    def __init__(self, name, dtype, dev, seed, fused_softmax):
        from yolo_dual_b200 import _lib
        N, H, W, G, gc = SITES[name]
        C, P = G * gc, 9
        g = torch.Generator(device=dev).manual_seed(seed)
        rn = lambda *s: torch.randn(*s, device=dev, generator=g)
        self.name = name
        self.input = rn(N, H, W, C).to(dtype)
        self.offset = rn(N, H, W, G * P * 2).to(dtype)          # sigma = 1 px
        logits = rn(N, H, W, G, P)
        self.mask = (logits if fused_softmax else torch.softmax(logits, -1)).reshape(N, H, W, G * P).to(dtype).contiguous()
        self.grad_out = rn(N, H, W, C).to(dtype)
        self.output = torch.empty_like(self.input)
        self.grad_input = torch.empty_like(self.input)
        self.grad_offset = torch.empty_like(self.offset)
        self.grad_mask = torch.empty_like(self.mask)
        self.geo = _lib.Geometry(N, H, W, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
        self.dt = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}[dtype]
        self.ws = None

    def alloc_ws(self, lib, accum):
        n = lib.dcnv3_b200_backward_workspace_bytes(self.dt, ctypes.byref(self.geo), accum)
        self.ws_bytes = n
        self.ws = torch.empty(max(n, 16), dtype=torch.uint8, device=self.input.device)


class Workload:
    def __init__(self, dev, dtype, sites, accum, fused_softmax):
        from yolo_dual_b200 import _lib
        self.lib = _lib.load()
        self._lib = _lib
        self.dev, self.dtype, self.sites = dev, dtype, sites
        self.accum = {"opmath": _lib.ACC_OPMATH, "storage": _lib.ACC_STORAGE, "tile": _lib.ACC_TILE}[accum]
        self.logits = int(fused_softmax)
        self.sets = []
        for r in range(N_BUFFER_SETS):
            bufs = [SiteBuffers(s, dtype, dev, 1000 * r + i, fused_softmax) for i, s in enumerate(sites)]
            for b in bufs:
                b.alloc_ws(self.lib, self.accum)
            self.sets.append(bufs)
        lowp = esize(dtype) == 2
        # my kernels per step and site: 1 forward + backward.  16-bit ACC_OPMATH backward at these shapes =
        # zero_select_kernel + bwd_imat_kernel + bwd_vec_kernel (returns at once unless the selector picked
        # it) + cast_ws_kernel; with a family forced by DCNV3_B200_BWD: backward kernel + cast (+ a driver memset).
        forced = os.environ.get("DCNV3_B200_BWD") in ("vec", "imat", "tile")
        bwd = (2 if forced else 4) if (lowp and accum == "opmath") else 1  # 'tile': bwd_win_kernel (+ a driver memset)
        self.launches_per_step = len(sites) * (1 + bwd)

    def fwd(self, b, st):
        rc = self.lib.dcnv3_b200_forward(b.input.data_ptr(), b.offset.data_ptr(), b.mask.data_ptr(),
                                         b.output.data_ptr(), b.dt, ctypes.byref(b.geo), self.logits, st)
        if rc:
            self._lib.check(rc, "dcnv3_b200_forward")

    def bwd(self, b, st):
        rc = self.lib.dcnv3_b200_backward(
            b.input.data_ptr(), b.offset.data_ptr(), b.mask.data_ptr(), b.grad_out.data_ptr(),
            b.grad_input.data_ptr(), b.grad_offset.data_ptr(), b.grad_mask.data_ptr(),
            b.ws.data_ptr(), b.ws_bytes, b.dt, ctypes.byref(b.geo), self.logits, self.accum, st)
        if rc:
            self._lib.check(rc, "dcnv3_b200_backward")

    def step(self, k, st):
        bufs = self.sets[k % N_BUFFER_SETS]
        for b in bufs:
            self.fwd(b, st)
        for b in reversed(bufs):
            self.bwd(b, st)


def time_steps(wl, steps, warmup, dist):
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for k in range(warmup):
        wl.step(k, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for k in range(steps):
        wl.step(warmup + k, st)
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=wl.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def time_ops(wl, steps, warmup):
    """Second timed region: CUDA-event pairs around every op of the step (same rotation, same
    order), for the per-op table and roofline.achieved of the dominant op."""
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    names = [f"fwd_{s}" for s in wl.sites] + [f"bwd_{s}" for s in reversed(wl.sites)]
    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in names]
           for _ in range(steps)]
    for k in range(warmup):
        wl.step(k, st)
    torch.cuda.synchronize()
    for k in range(steps):
        bufs = wl.sets[(warmup + k) % N_BUFFER_SETS]
        i = 0
        for b in bufs:
            evs[k][i][0].record(); wl.fwd(b, st); evs[k][i][1].record(); i += 1
        for b in reversed(bufs):
            evs[k][i][0].record(); wl.bwd(b, st); evs[k][i][1].record(); i += 1
    torch.cuda.synchronize()
    out = {}
    for i, nm in enumerate(names):
        ts = [evs[k][i][0].elapsed_time(evs[k][i][1]) * 1e3 for k in range(steps)]
        out[nm] = {"us_mean": sum(ts) / len(ts), "us_median": statistics.median(ts)}
    return out


def time_e2e(wl, steps, warmup, dist):
    """Same step through the public API with HOST buffers: yolo_dual_b200.host.HostPipeline drives
    DCNv3Function.apply + autograd; every step copies its inputs from pinned host memory and writes
    every result (output + three grads per site) back to pinned host memory.  Copies in, kernels and
    copies out run on three streams (PCIe is full duplex), three steps in flight."""
    from yolo_dual_b200.host import HostPipeline, HostSite, pack_sites
    from yolo_dual_b200.ops_dcnv3.functions import set_grad_accum
    set_grad_accum({0: "opmath", 1: "storage", 2: "tile"}[wl.accum])
    sites = []
    for b in wl.sets[0]:
        N, H, W, G, gc = SITES[b.name]
        hs = HostSite(*(t.cpu().pin_memory() for t in (b.input, b.offset, b.mask, b.grad_out)),
                      args=(3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0))
        sites.append(hs.alloc_outputs(tuple(b.output.shape)))
    sites = pack_sites(sites)  # one pinned arena per direction: one copy each way per step
    h2d = sum(s.h2d_bytes for s in sites)
    d2h = sum(s.d2h_bytes for s in sites)
    pipe = HostPipeline(wl.dev, depth=3, fused_softmax=bool(wl.logits))
    for _ in range(warmup):
        pipe.submit(sites)
    pipe.drain()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        pipe.submit(sites)
    pipe.drain()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        dist.barrier()
        t = torch.tensor([ms], device=wl.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, h2d, d2h


# ----------------------------------------------------------------------------------------------
# The caller of the hot path: C3-DCN seg training step (BASELINE configs[1]/[3]) — imgs/s.
# YOLOv5-style seg model with C3_DCNV3 in the three C3_DCN slots, 640x640, batch 16 per GPU,
# bf16 autocast, CE + 0.5 Dice, SGD-nesterov; DDP (NCCL) gradient all-reduce when world > 1.
# Every step copies its images/labels from pinned host memory and reads the loss back.
# ----------------------------------------------------------------------------------------------
def time_seg(dev, dist, world, steps, warmup, batch, model_name, channels_last=True):
    from yolo_dual_b200 import seg
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    cfg = {"yolov5seg": seg.YOLOV5_SEG, "yolov8seg": seg.YOLOV8_SEG}[model_name]
    model = seg.SegModel(cfg, dcn="dcnv3").to(dev)
    if channels_last:  # NHWC activations: the NCHW<->NHWC permutes around every DCNv3 become views
        model = model.to(memory_format=torch.channels_last)
    crit = seg.SegmentationLoss(12, class_weights=seg.CAMVID_CLASS_WEIGHTS).to(dev)
    ddp = seg.wrap_ddp(model, dev)
    opt = seg.smart_optimizer(ddp)
    ddp.train()
    g = torch.Generator().manual_seed(1 + (dist.get_rank() if dist is not None else 0))
    imgs_h = torch.randn(batch, 3, 640, 640, generator=g).pin_memory()
    lab_h = torch.randint(0, 12, (batch, 640, 640), generator=g).pin_memory()
    # the data-loader side of the loop: batch k+1 is uploaded on a copy stream while step k computes (two device
    # slots), and the loss of step k is read back into pinned memory and looked at one step later
    copy_st = torch.cuda.Stream(device=dev)
    slots = [(torch.empty(batch, 3, 640, 640, device=dev), torch.empty(batch, 640, 640, dtype=torch.int64, device=dev))
             for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    loss_h = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0, "last": float("nan")}

    def upload(k):
        with torch.cuda.stream(copy_st):
            copy_st.wait_event(freed[k % 2])
            slots[k % 2][0].copy_(imgs_h, non_blocking=True)
            slots[k % 2][1].copy_(lab_h, non_blocking=True)
            ready[k % 2].record(copy_st)

    for ev in freed:
        ev.record()
    upload(0)

    def one():
        k = state["k"]
        upload(k + 1)
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[k % 2])
        imgs, lab = slots[k % 2]
        if channels_last:
            imgs = imgs.contiguous(memory_format=torch.channels_last)
        loss, _ = seg.train_step(ddp, crit, opt, imgs, lab, autocast_dtype=torch.bfloat16)
        freed[k % 2].record(cur)
        loss_h[k % 2].copy_(loss, non_blocking=True)   # D2H: the step's result
        loss_ev[k % 2].record(cur)
        if k:
            loss_ev[(k - 1) % 2].synchronize()
            state["last"] = float(loss_h[(k - 1) % 2])
        state["k"] = k + 1
        return state["last"]

    for _ in range(warmup):
        last = one()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        last = one()
    e1.record()
    torch.cuda.synchronize()
    last = float(loss_h[(state["k"] - 1) % 2])
    ms = e0.elapsed_time(e1)
    if dist is not None:
        dist.barrier()
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    n_params = sum(p.numel() for p in model.parameters())
    return {"model": f"{model_name} with C3_DCNV3 at P3/P4/P5 (DCNv3 C=128/256/512, group_channels 16)",
            "imgs_per_s": batch * world * steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps,
            "batch_per_gpu": batch, "global_batch": batch * world, "image": "640x640", "autocast": "bf16",
            "memory_format": "channels_last" if channels_last else "contiguous",
            "optimizer": "SGD nesterov 3 groups", "params": n_params, "loss_last": last,
            "h2d_bytes_per_step": imgs_h.numel() * 4 + lab_h.numel() * 8, "d2h_bytes_per_step": 4,
            "input_pipeline": "pinned host batch uploaded every step on a copy stream, one step ahead; loss read back "
                              "into pinned memory every step and inspected one step later",
            "fused": "deferred last Upsample, fused CE+Dice loss, fused BN+SiLU, NHWC resize kernels (this repo's "
                     "segloss_b200 / bnact_b200 / resize_b200)",
            "data_parallel": f"DDP x{world} (NCCL all-reduce of {n_params * 4 / 1e6:.1f} MB fp32 grads)" if world > 1 else "single GPU"}


# ----------------------------------------------------------------------------------------------
# GPU baseline: the reference's OWN CUDA kernels rebuilt for sm_100a (oracle/_ref, test
# infrastructure, built by oracle/build_ref_cuda.py where /root/reference is mounted).  They have no
# bf16 path (dcnv3_cuda.cu:69,147), so the comparison runs in fp16 on both sides.
# ----------------------------------------------------------------------------------------------
def time_reference_cuda(dev, sites, steps, warmup):
    from oracle.build_ref_cuda import load_module
    ref = load_module()
    if ref is None:
        return None
    dt = torch.float16
    sets = []
    for r in range(N_BUFFER_SETS):
        sets.append([SiteBuffers(s, dt, dev, 1000 * r + i, False) for i, s in enumerate(sites)])

    def step(k):
        bufs = sets[k % N_BUFFER_SETS]
        for b in bufs:
            N, H, W, G, gc = SITES[b.name]
            ref.dcnv3_forward(b.input, b.offset, b.mask, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0, 256)
        for b in reversed(bufs):
            N, H, W, G, gc = SITES[b.name]
            ref.dcnv3_backward(b.input, b.offset, b.mask, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0, b.grad_out, 256)

    for k in range(warmup):
        step(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(steps):
        step(warmup + k)
    e1.record()
    torch.cuda.synchronize()
    ref_ms = e0.elapsed_time(e1) / steps
    del sets
    wl = Workload(dev, dt, sites, "opmath", False)
    ours_ms = time_steps(wl, steps, warmup, None) / steps
    return {"what": "the reference's own CUDA kernels (dcnv3_im2col_cuda.cuh) rebuilt for sm_100a, same step, "
                    "fp16 (the reference has no bf16), outputs allocated inside as it does",
            "dtype": "fp16", "steps": steps, "reference_ms_per_step": ref_ms, "ours_fp16_ms_per_step": ours_ms,
            "speedup": ref_ms / ours_ms}


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own CPU algorithm for this path is the pure-PyTorch
# dcnv3_core_pytorch; /root/reference is not on the GPU box, so its restatement in oracle/ is timed
# (kind "port").  This is the only place bench.py executes oracle/.
# ----------------------------------------------------------------------------------------------
def cpu_step(sites, n_img, tensors):
    from oracle.dcnv3_oracle import core_torch_fwd_bwd
    for s in sites:
        N, H, W, G, gc = SITES[s]
        x, off, m, go = tensors[s]
        core_torch_fwd_bwd(x[:n_img], off[:n_img], m[:n_img], go[:n_img], 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)


def cpu_tensors(sites):
    from oracle.dcnv3_oracle import make_inputs
    return {s: make_inputs(*SITES[s], dist="unit", seed=i) for i, s in enumerate(sites)}


def cpu_run(sites, e_workload, steps, warmup, budget_s):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tens = cpu_tensors(sites)
    n_img = SITES[sites[0]][0]
    t0 = time.perf_counter(); cpu_step(sites, 1, tens); t1 = time.perf_counter() - t0  # also a warm-up
    while n_img > 1 and t1 * n_img * (steps + warmup) > budget_s:
        n_img //= 2
    for _ in range(warmup):
        cpu_step(sites, n_img, tens)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); cpu_step(sites, n_img, tens); ts.append(time.perf_counter() - t0)
    total = sum(ts)
    by = sum(sum(algo_bytes(s, e_workload, N=n_img)) for s in sites)
    gbps = by * steps / total / 1e9
    sample = (f"{steps} steps of fwd+bwd over {'+'.join(sites)} on {n_img} of {SITES[sites[0]][0]} images, fp32 "
              f"torch restatement of dcnv3_core_pytorch, {torch.get_num_threads()} threads; GB/s counts the "
              f"workload's storage-dtype algorithmic bytes so the GPU/CPU ratio is a time ratio")
    return gbps, total / steps * 1e3, cores, sample, n_img


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner on init), so fd 1 is pointed at stderr for the whole run and the JSON goes to the saved fd."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    out_stream = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--sites", default="P3,P4,P5")
    ap.add_argument("--grad-accum", default="tile", choices=["tile", "opmath", "storage"])
    ap.add_argument("--fused-softmax", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-seg", action="store_true")
    ap.add_argument("--seg-steps", type=int, default=10)
    ap.add_argument("--seg-batch", type=int, default=16, help="images per GPU (weak scaling)")
    ap.add_argument("--seg-global-batch", type=int, default=0,
                    help="fixed global batch split over the GPUs (strong scaling, BASELINE configs[3]: 64)")
    ap.add_argument("--seg-model", default="yolov5seg", choices=["yolov5seg", "yolov8seg"])
    ap.add_argument("--seg-nchw", action="store_true", help="keep NCHW activations in the seg model")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    sites = a.sites.split(",")
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]
    e = esize(dtype)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    step_bytes = sum(sum(algo_bytes(s, e)) for s in sites)
    config = {"workload": "configs[1]: DCNv3 core fwd+bwd at the three C3-DCN sites of the 640x640 "
                          "YOLOv5-style seg model, batch 16 per GPU",
              "sites": {s: dict(zip(("N", "H", "W", "G", "gc"), SITES[s])) for s in sites},
              "kernel": "3x3 s1 p1 d1", "offset_scale": 1.0,
              "inputs": "input~N(0,1), offset~N(0,1) px, mask=softmax(N(0,1)) over 9 points",
              "grad_accum": a.grad_accum, "fused_softmax": bool(a.fused_softmax),
              "l2": f"inputs rotate over {N_BUFFER_SETS} buffer sets (> 126 MB L2), fwd P3,P4,P5 then bwd P5,P4,P3",
              "algorithmic_bytes_per_step": step_bytes, "parallelism": f"dp{world} (batch shard, no data-path collective)"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        gbps, ms, cores, sample, n_img = cpu_run(sites, e, a.steps, a.warmup, budget_s=150.0)
        print(file=out_stream, flush=True, *[json.dumps({
            "impl": "reference", "metric": METRIC, "value": gbps, "unit": "GB/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": gbps, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": gbps, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})])
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: dcnv3_b200 has no CPU path")
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    clocks = Clocks(local) if rank == 0 else None

    wl = Workload(dev, dtype, sites, a.grad_accum, a.fused_softmax)
    launches = wl.launches_per_step * a.steps
    ms = time_steps(wl, a.steps, a.warmup, dist)
    ops = time_ops(wl, min(a.steps, 100), a.warmup)
    e2e = None
    if not a.no_e2e:
        e2e_ms, h2d, d2h = time_e2e(wl, a.e2e_steps, 3, dist)
        e2e = {"value": step_bytes * world * a.e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / a.e2e_steps,
               "steps": a.e2e_steps,
               "api": "yolo_dual_b200.host.HostPipeline -> DCNv3Function.apply + autograd on pinned host tensors; "
                      "H2D of input/offset/mask/grad_out and D2H of output + 3 grads every step (one pinned arena per direction), copies and kernels "
                      "on three streams, three steps in flight"}
    seg_res = None
    if not a.no_seg:
        del wl
        torch.cuda.empty_cache()
        try:
            seg_batch = a.seg_global_batch // world if a.seg_global_batch else a.seg_batch
            seg_res = time_seg(dev, dist, world, a.seg_steps, 3, seg_batch, a.seg_model, not a.seg_nchw)
            seg_res["scaling"] = "strong (global batch fixed)" if a.seg_global_batch else "weak (batch per GPU fixed)"
        except Exception as ex:
            if world > 1:
                raise  # a rank must not leave a collective half-done
            seg_res = {"error": repr(ex)[:300]}
        wl = None
    ref_cuda = None
    if rank == 0 and not a.no_ref_cuda:
        try:
            ref_cuda = time_reference_cuda(dev, sites, 20, 3)
        except Exception as ex:  # the baseline must never take the bench down
            ref_cuda = {"error": repr(ex)[:300]}
    clk = clocks.stop() if clocks else None
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = load_peaks()
    value = step_bytes * world * a.steps / (ms * 1e-3) / 1e9
    table = {}
    for nm, t in ops.items():
        kind, s = nm.split("_")
        by = algo_bytes(s, e)[0 if kind == "fwd" else 1]
        t["algorithmic_bytes"] = by
        t["GBps"] = by / (t["us_mean"] * 1e-6) / 1e9
        t["frac_of_hbm"] = t["GBps"] / peak
        table[nm] = t
    dom = max(table, key=lambda k: table[k]["us_mean"])
    roofline = {"bound": "hbm", "kernel": dom + ((" (memset + bwd_vec_kernel + cast_ws_kernel)" if os.environ.get("DCNV3_B200_BWD") == "vec" else
                                        " (zero_select_kernel + bwd_imat_kernel [+ bwd_vec_kernel, skipped by the selector] + cast_ws_kernel)") if dom.startswith("bwd") and e == 2 and a.grad_accum == "opmath" else ""),
                "achieved": table[dom]["GBps"], "peak": peak, "unit": "GB/s", "frac": table[dom]["GBps"] / peak,
                "peak_source": peak_src, "traffic": ncu_traffic(dom),
                "step_frac": value / world / peak}
    out = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": a.dtype if a.dtype != "fp32" else "f32", "data": "synthetic",
           "config": config, "pct_hbm_peak": 100.0 * value / world / peak, "roofline": roofline,
           "ops": table, "seg_train": seg_res, "reference_cuda": ref_cuda, "e2e": e2e, "gpu_launches": launches, "clocks": clk}
    if not a.no_cpu_baseline:
        gbps, cms, cores, sample, _ = cpu_run(sites, e, 3, 1, budget_s=30.0)
        out["cpu_baseline"] = {"value": gbps, "unit": "GB/s", "cores": cores, "kind": "port",
                               "sample": sample, "ms_per_step_sample": cms}
    print(json.dumps(out), file=out_stream, flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
