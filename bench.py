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
    # BASELINE configs[4]: 1280x1280 inference, batch 32 over 8 GPUs = 4 images per GPU (--infer)
    "I3": (4, 160, 160, 8, 16),
    "I4": (4, 80, 80, 16, 16),
    "I5": (4, 40, 40, 32, 16),
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
    """dram__bytes_read.sum + dram__bytes_write.sum summed over EVERY kernel the op launches (16-bit backward:
    zero_fill_kernel + bwd_win_kernel), per call, from the committed ncu capture of this same command
    (profiles/r02_ncu_traffic.json, written by tools/ncu_traffic.py from profiles/r02_ncu_launches.csv);
    None when that op was not captured.  A constant of the capture, not something this run measured."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
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
        clip = float(os.environ.get("BENCH_OFFSET_CLIP", "0"))  # experiments only (profiles/r02_slow_path.md): |offset| <= clip px
        if clip > 0:
            self.offset = self.offset.clamp(-clip, clip)
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
        bwd = (2 if forced else 4) if (lowp and accum == "opmath") else (2 if lowp and accum == "tile" else 1)  # 'tile': zero_fill_kernel + bwd_win_kernel
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
    del pipe
    # The ceiling of this leg on THIS box at THIS rank count: the same bytes per step as two plain pinned copies, one
    # per direction on two streams (PCIe is full duplex), all ranks at once, no kernels.  What HostPipeline adds on top
    # of it is the pipeline's own cost; what the ranks lose against N x the single-rank figure is the host side of the
    # box (root complex / host memory shared by the ranks), not this library.
    a_in, a_out = sites[0]._arena[0], sites[0]._arena[1]
    d_in = torch.empty(a_in.numel(), dtype=torch.uint8, device=wl.dev)
    d_out = torch.empty(a_out.numel(), dtype=torch.uint8, device=wl.dev)
    s1, s2 = torch.cuda.Stream(wl.dev), torch.cuda.Stream(wl.dev)

    def raw(reps):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1):
                d_in.copy_(a_in, non_blocking=True)
            with torch.cuda.stream(s2):
                a_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / reps
    raw(2)
    raw_ms = raw(8)
    if dist is not None:
        t = torch.tensor([raw_ms], device=wl.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        raw_ms = float(t.item())
    return ms, h2d, d2h, raw_ms


# ----------------------------------------------------------------------------------------------
# The caller of the hot path: C3-DCN seg training step (BASELINE configs[1]/[3]) — imgs/s.
# YOLOv5-style seg model with C3_DCNV3 in the three C3_DCN slots, 640x640, batch 16 per GPU,
# bf16 autocast, CE + 0.5 Dice, SGD-nesterov; DDP (NCCL) gradient all-reduce when world > 1.
# Every step copies its images/labels from pinned host memory and reads the loss back.
# ----------------------------------------------------------------------------------------------
def pinned_like(t, channels_last):
    """Pinned host copy of an NCHW tensor; channels_last = stored HWC (what an image decoder produces), so that the
    upload into a channels_last device tensor is one plain DMA."""
    if not channels_last:
        return t.pin_memory()
    n, c, h, w = t.shape
    buf = torch.empty(n, h, w, c, dtype=t.dtype).pin_memory().permute(0, 3, 1, 2)
    buf.copy_(t)
    return buf


SEG_DCN_OPTS = {"fused_softmax": False, "packed_heads": False}  # --seg-fused-heads switches both on


def seg_model(dev, model_name, channels_last=True, fused_softmax=None):
    from yolo_dual_b200 import seg
    torch.manual_seed(0)  # same initial weights on every rank
    cfg = {"yolov5seg": seg.YOLOV5_SEG, "yolov8seg": seg.YOLOV8_SEG}[model_name]
    opts = dict(SEG_DCN_OPTS) if fused_softmax is None else {"fused_softmax": fused_softmax, "packed_heads": False}
    model = seg.SegModel(cfg, dcn="dcnv3", **opts).to(dev)
    if channels_last:  # NHWC activations: the NCHW<->NHWC permutes around every DCNv3 become views
        model = model.to(memory_format=torch.channels_last)
    return model


def time_seg(dev, dist, world, steps, warmup, batch, model_name, channels_last=True, ddp_opts=None, graph=False):
    from yolo_dual_b200 import seg
    torch.backends.cudnn.benchmark = True
    ddp_opts = ddp_opts or {}
    model = seg_model(dev, model_name, channels_last)
    crit = seg.SegmentationLoss(12, class_weights=seg.CAMVID_CLASS_WEIGHTS).to(dev)
    side = torch.cuda.Stream(device=dev)  # torch's CUDA-graph notes: build DDP on a side stream when it will be captured
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ddp = seg.wrap_ddp(model, dev, **ddp_opts)
    torch.cuda.current_stream().wait_stream(side)
    opt = seg.smart_optimizer(ddp)
    ddp.train()
    g = torch.Generator().manual_seed(1 + (dist.get_rank() if dist is not None else 0))
    imgs_h = pinned_like(torch.randn(batch, 3, 640, 640, generator=g), channels_last)
    lab_h = torch.randint(0, 12, (batch, 640, 640), generator=g).pin_memory()
    # the data-loader side of the loop: batch k+1 is uploaded on a copy stream while step k computes (two device
    # slots), and the loss of step k is read back into pinned memory and looked at one step later
    copy_st = torch.cuda.Stream(device=dev)
    slots = [(torch.empty(batch, 3, 640, 640, device=dev), torch.empty(batch, 640, 640, dtype=torch.int64, device=dev))
             for _ in range(2)]
    if channels_last:
        slots = [(a.contiguous(memory_format=torch.channels_last), b) for a, b in slots]
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

    graphed, graph_note = None, "off"
    if graph:
        try:
            def batches():
                while True:
                    yield slots[0]
            torch.cuda.current_stream().wait_event(ready[0])
            graphed = seg.GraphedTrainStep(ddp, crit, opt, batches(), autocast_dtype=torch.bfloat16)
            graph_note = "whole step (forward, loss, backward + DDP all-reduce, SGD) replayed as one CUDA graph"
        except Exception as ex:  # capture is an optimisation: report and run eagerly
            if world > 1:
                raise
            graphed, graph_note = None, "capture failed, eager: " + repr(ex)[:200]

    def one():
        k = state["k"]
        upload(k + 1)
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[k % 2])
        imgs, lab = slots[k % 2]
        if graphed is not None:
            loss, _ = graphed(imgs, lab)
        else:
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
    comp = ddp_opts.get("grad_compress") or "fp32"
    res = {"model": f"{model_name} with C3_DCNV3 at P3/P4/P5 (DCNv3 C=128/256/512, group_channels 16)",
           "imgs_per_s": batch * world * steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps,
           "batch_per_gpu": batch, "global_batch": batch * world, "image": "640x640", "autocast": "bf16",
           "memory_format": "channels_last" if channels_last else "contiguous",
           "optimizer": "SGD nesterov 3 groups", "params": n_params, "loss_last": last,
           "h2d_bytes_per_step": imgs_h.numel() * 4 + lab_h.numel() * 8, "d2h_bytes_per_step": 4,
           "input_pipeline": "pinned host batch uploaded every step on a copy stream, one step ahead; loss read back "
                             "into pinned memory every step and inspected one step later",
           "fused": "deferred last Upsample, fused CE+Dice loss, fused BN+SiLU, NHWC resize kernels (this repo's "
                    "segloss_b200 / bnact_b200 / resize_b200)",
           "cuda_graph": graph_note, "dcnv3_module": dict(SEG_DCN_OPTS),
           "data_parallel": (f"DDP x{world} (NCCL all-reduce of {n_params * (2 if comp == 'bf16' else 4) / 1e6:.1f} MB {comp} grads, "
                             f"bucket_cap_mb {ddp_opts.get('bucket_cap_mb') or 25}, first bucket {ddp_opts.get('first_bucket_mb') or 1} MB)")
           if world > 1 else "single GPU"}
    del graphed, ddp, opt, model, slots
    torch.cuda.empty_cache()
    return res


def dp_grad_check(dev, dist, world, model_name, batch, ddp_opts=None):
    """SURVEY §8(e)'s correctness test on the real thing: the gradients DDP leaves on every rank after ONE backward over
    the rank's shard (the CUDA DCNv3 op, bf16 autocast, fused loss, NCCL all-reduce) against rank 0's own
    recomputation over the concatenated global batch — shard by shard, gradients accumulated and divided by the
    world size, which is what data parallelism promises (the class-weighted CE normalises per shard, so it is the
    shard losses that are averaged, reference and here alike).  BatchNorm in eval mode: batch statistics are
    per-shard by design (no SyncBN, seg_diceloss_yolov5.py:991).  Returns max over parameters of
    max|g_ddp - g_ref| / max|g_ref|."""
    from yolo_dual_b200 import seg
    model = seg_model(dev, model_name)
    # non-zero DCNv3 heads: a fresh layer has zero offset / mask heads and would sample the regular grid only
    gen = torch.Generator(device="cpu").manual_seed(7)
    with torch.no_grad():
        for _, m in model.dcn_sites():
            m.offset.weight.copy_(torch.randn(m.offset.weight.shape, generator=gen) * 0.05)
            m.mask.weight.copy_(torch.randn(m.mask.weight.shape, generator=gen) * 0.05)
    crit = seg.SegmentationLoss(12, class_weights=seg.CAMVID_CLASS_WEIGHTS).to(dev)
    model.eval()
    ddp = seg.wrap_ddp(model, dev, **(ddp_opts or {}))
    rank = dist.get_rank()
    g = torch.Generator().manual_seed(100 + rank)
    imgs = torch.randn(batch, 3, 640, 640, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    lab = torch.randint(0, 12, (batch, 640, 640), generator=g).to(dev)
    loss, _ = seg.forward_loss(ddp, crit, imgs, lab, torch.bfloat16)
    loss.backward()
    torch.cuda.synchronize()
    got = {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}
    all_imgs = [torch.empty_like(imgs) for _ in range(world)]
    all_lab = [torch.empty_like(lab) for _ in range(world)]
    dist.all_gather(all_imgs, imgs.contiguous())
    dist.all_gather(all_lab, lab)
    out = None
    if rank == 0:
        model.zero_grad(set_to_none=True)
        for xi, yi in zip(all_imgs, all_lab):
            l, _ = seg.forward_loss(model, crit, xi.contiguous(memory_format=torch.channels_last), yi, torch.bfloat16)
            (l / world).backward()
        worst, name = 0.0, ""
        for n, p in model.named_parameters():
            if p.grad is None or n not in got:
                continue
            ref = p.grad.detach().float()
            rel = float((got[n] - ref).abs().max() / ref.abs().max().clamp_min(1e-20))
            if rel > worst:
                worst, name = rel, n
        out = {"dp_grad_max_rel": worst, "worst_param": name, "params_compared": len(got),
               "what": f"DDP x{world} gradients ({model_name}, {batch} images per rank, bf16 autocast, CUDA DCNv3 op, BN eval) vs rank 0's "
                       f"shard-by-shard recomputation on the gathered global batch"}
    dist.barrier()
    del ddp, model
    torch.cuda.empty_cache()
    return out


def trainer_smoke(dev, dist, world):
    """`Trainer` under NCCL: nominal-batch accumulation with DDP.no_sync on the non-boundary
    micro-steps, EMA, checkpoint + resume, with the real DCNv3 op.  Returns a short status dict (rank 0)."""
    import tempfile
    from yolo_dual_b200 import seg
    from yolo_dual_b200.trainer import Trainer
    model = seg_model(dev, "yolov5seg")
    crit = seg.SegmentationLoss(12, class_weights=seg.CAMVID_CLASS_WEIGHTS).to(dev)
    t = Trainer(model, crit, batch_size=16 * world, epochs=3, autocast_dtype=torch.bfloat16, device=dev)
    g = torch.Generator().manual_seed(5 + dist.get_rank())
    x = torch.randn(4, 3, 640, 640, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 12, (4, 640, 640), generator=g).to(dev)
    stepped = []
    for i in range(2 * t.accumulate):
        loss, _, st = t.micro_step(x, y)
        stepped.append(bool(st))
    t.end_epoch(fitness=0.5)
    ok_sync = True
    for p in model.parameters():  # after an optimizer step every rank must hold the same weights
        ref = p.detach().clone()
        dist.broadcast(ref, 0)
        ok_sync &= bool(torch.equal(ref, p.detach()))
    out = None
    if dist.get_rank() == 0:
        with tempfile.TemporaryDirectory() as d:
            t.save(os.path.join(d, "last.pt"), os.path.join(d, "best.pt"), is_best=True)
            ck = torch.load(os.path.join(d, "last.pt"), weights_only=False)
        out = {"accumulate": t.accumulate, "optimizer_steps": sum(stepped), "micro_steps": len(stepped),
               "ema_updates": t.ema.updates, "weights_identical_across_ranks": ok_sync, "ckpt_epoch": ck["epoch"],
               "loss_finite": bool(torch.isfinite(loss)), "world": world}
    dist.barrier()
    del t, model
    torch.cuda.empty_cache()
    return out


def time_infer(dev, dist, world, steps, warmup, batch):
    """BASELINE configs[4]: C3-DCN seg inference at 1280x1280, fp16, fused mask-softmax, `batch` images per GPU (batch
    32 sharded over 8 GPUs = 4).  Every step uploads its images from pinned host memory and reads the class map
    back; images/s over all ranks."""
    from yolo_dual_b200 import seg
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = seg.SegModel(seg.YOLOV5_SEG, dcn="dcnv3", fused_softmax=True, img_size=(1280, 1280)).to(dev)
    model = model.half().eval().to(memory_format=torch.channels_last)
    g = torch.Generator().manual_seed(3 + (dist.get_rank() if dist is not None else 0))
    imgs_h = pinned_like(torch.randn(batch, 3, 1280, 1280, generator=g).half(), True)
    out_h = torch.empty(batch, 1280, 1280, dtype=torch.uint8).pin_memory()
    x = torch.empty(batch, 3, 1280, 1280, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)

    def one():
        x.copy_(imgs_h, non_blocking=True)
        with torch.no_grad():
            prob = model(x)
            out_h.copy_(prob.argmax(1).to(torch.uint8), non_blocking=True)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        dist.barrier()
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    del model
    torch.cuda.empty_cache()
    return {"workload": "configs[4]: yolov5seg with C3_DCNV3, 1280x1280, fp16, fused mask-softmax (mask logits go straight to the "
                        "kernels), eval", "imgs_per_s": batch * world * steps / (ms * 1e-3), "ms_per_step": ms / steps,
            "batch_per_gpu": batch, "global_batch": batch * world, "steps": steps,
            "h2d_bytes_per_step": imgs_h.numel() * 2, "d2h_bytes_per_step": out_h.numel(),
            "classes_seen": int(out_h.max()) + 1}


# ----------------------------------------------------------------------------------------------
# GPU baseline: the reference's OWN CUDA kernels rebuilt for sm_100a (oracle/_ref, test
# infrastructure, built by oracle/build_ref_cuda.py where /root/reference is mounted).  They have no
# bf16 path (dcnv3_cuda.cu:69,147), so the comparison runs in fp16 on both sides.
# ----------------------------------------------------------------------------------------------
def time_reference_cuda(dev, sites, steps, warmup, accum="tile"):
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
    wl = Workload(dev, dt, sites, accum, False)
    ours_ms = time_steps(wl, steps, warmup, None) / steps
    return {"what": "the reference's own CUDA kernels (dcnv3_im2col_cuda.cuh) rebuilt for sm_100a, same step, "
                    "fp16 (the reference has no bf16), outputs allocated inside as it does",
            "dtype": "fp16", "steps": steps, "reference_ms_per_step": ref_ms, "ours_fp16_ms_per_step": ours_ms,
            "ours_grad_accum": accum,
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


CPU_SAMPLE_IMAGES = 2  # fixed: the CPU arm always runs the same 2 of the 16 images, so GPU/CPU ratios compare run to run


def cpu_run(sites, e_workload, steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tens = cpu_tensors(sites)
    n_img = min(CPU_SAMPLE_IMAGES, SITES[sites[0]][0])
    cpu_step(sites, 1, tens)  # page-in / thread-pool warm-up
    for _ in range(warmup):
        cpu_step(sites, n_img, tens)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); cpu_step(sites, n_img, tens); ts.append(time.perf_counter() - t0)
    total = sum(ts)
    by = sum(sum(algo_bytes(s, e_workload, N=n_img)) for s in sites)
    gbps = by * steps / total / 1e9
    sample = (f"{steps} steps of fwd+bwd over {'+'.join(sites)} on {n_img} of {SITES[sites[0]][0]} images (fixed sample), fp32 "
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
    ap.add_argument("--seg-fused-heads", action="store_true",
                    help="seg models with DCNv3(fused_softmax=True, packed_heads=True): softmax inside the kernels, both heads one GEMM")
    ap.add_argument("--no-seg-strong", action="store_true", help="skip the configs[3] line (yolov8seg, global batch 64)")
    ap.add_argument("--seg-graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the training step as one CUDA graph (auto: on)")
    ap.add_argument("--ddp-compress", default="none", choices=["bf16", "none"],
                    help="gradient all-reduce dtype (none = fp32, the reference's plain DDP; bf16 measured no faster at 2 GPUs)")
    ap.add_argument("--ddp-bucket-mb", type=float, default=0.0, help="DDP bucket_cap_mb (0: torch's 25)")
    ap.add_argument("--ddp-first-bucket-mb", type=float, default=0.0, help="DDP first bucket (0: torch's 1)")
    ap.add_argument("--no-dp-check", action="store_true")
    ap.add_argument("--no-infer", action="store_true", help="skip the configs[4] inference line")
    ap.add_argument("--infer-steps", type=int, default=10)
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.seg_fused_heads:
        SEG_DCN_OPTS.update(fused_softmax=True, packed_heads=True)
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
        # bounded: every step runs the same fixed 2-image sample (~0.36 s on 16 cores); the arm caps its own step count
        # so that any --steps / --warmup ends within a few minutes (the line reports the steps it ran)
        a.steps, a.warmup = min(a.steps, 600), min(a.warmup, 10)  # ~0.36 s per step on the box's 16 cores
        gbps, ms, cores, sample, n_img = cpu_run(sites, e, a.steps, a.warmup)
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
        e2e_ms, h2d, d2h, raw_ms = time_e2e(wl, a.e2e_steps, 3, dist)
        e2e = {"value": step_bytes * world * a.e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / a.e2e_steps,
               "steps": a.e2e_steps,
               "raw_duplex_copy_ms_per_step": raw_ms,
               "raw_duplex_copy": f"the same {h2d} + {d2h} bytes as two plain pinned copies (one per direction, two streams), "
                                  f"all {world} rank(s) at once, no kernels, max over ranks: the box's ceiling for this leg",
               "copy_bound_frac": raw_ms / (e2e_ms / a.e2e_steps),
               "api": "yolo_dual_b200.host.HostPipeline -> DCNv3Function.apply + autograd on pinned host tensors; "
                      "H2D of input/offset/mask/grad_out and D2H of output + 3 grads every step (one pinned arena per direction), copies and kernels "
                      "on three streams, three steps in flight"}
    seg_res, seg_detail, infer = None, None, None
    if not a.no_seg:
        del wl
        torch.cuda.empty_cache()
        ddp_opts = {"grad_compress": None if a.ddp_compress == "none" else a.ddp_compress,
                    "bucket_cap_mb": a.ddp_bucket_mb or None, "first_bucket_mb": a.ddp_first_bucket_mb or None}
        graph = a.seg_graph != "off"
        seg_res, seg_detail = {}, {}

        def short(r, scaling):
            return {"model": r["model"].split(" ")[0], "imgs_per_s": r["imgs_per_s"], "ms_per_step": r["ms_per_step"],
                    "batch_per_gpu": r["batch_per_gpu"], "global_batch": r["global_batch"], "scaling": scaling,
                    "cuda_graph": not r["cuda_graph"].startswith(("off", "capture failed"))}
        try:
            if world > 1 and not a.no_dp_check:
                chk = dp_grad_check(dev, dist, world, "yolov5seg", 4, ddp_opts)
                tr = trainer_smoke(dev, dist, world)
                if rank == 0:
                    seg_detail["dp_check"], seg_detail["trainer_nccl"] = chk, tr
                    seg_res["dp_grad_max_rel"] = chk["dp_grad_max_rel"]
                    seg_res["trainer_nccl"] = "ok" if tr["weights_identical_across_ranks"] and tr["loss_finite"] else "FAILED"
            seg_batch = a.seg_global_batch // world if a.seg_global_batch else a.seg_batch
            r = time_seg(dev, dist, world, a.seg_steps, 3, seg_batch, a.seg_model, not a.seg_nchw, ddp_opts, graph)
            sc = "strong (global batch fixed)" if a.seg_global_batch else "weak (batch per GPU fixed)"
            seg_detail["weak" if not a.seg_global_batch else "strong"] = r
            seg_res["weak" if not a.seg_global_batch else "strong"] = short(r, sc)
            if not a.no_seg_strong and not a.seg_global_batch and 64 % world == 0:
                # BASELINE configs[3]: YOLOv8-seg, global batch 64 split over the GPUs (strong scaling)
                r = time_seg(dev, dist, world, a.seg_steps, 3, 64 // world, "yolov8seg", not a.seg_nchw, ddp_opts, graph)
                seg_detail["strong"] = r
                seg_res["strong"] = short(r, "strong (global batch 64 fixed)")
        except Exception as ex:
            if world > 1:
                raise  # a rank must not leave a collective half-done
            seg_res = {"error": repr(ex)[:300]}
        wl = None
    if not a.no_infer:
        try:
            infer = time_infer(dev, dist, world, a.infer_steps, 3, 4)
            iw = Workload(dev, torch.float16, ["I3", "I4", "I5"], "tile", True)
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            per = {}
            for i, nm in enumerate(iw.sites):
                evs = []
                for k in range(3 + 20):
                    b = iw.sets[k % N_BUFFER_SETS][i]
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); iw.fwd(b, st); e1.record()
                    evs.append((e0, e1))
                torch.cuda.synchronize()
                us = statistics.mean(x.elapsed_time(y) for x, y in evs[3:]) * 1e3
                by = algo_bytes(nm, 2)[0]
                per[nm] = {"us": us, "GBps": by / us / 1e3, "shape": dict(zip(("N", "H", "W", "G", "gc"), SITES[nm]))}
            infer["dcnv3_fwd_sites_fp16_fused_softmax"] = per
            del iw
            torch.cuda.empty_cache()
        except Exception as ex:
            if world > 1:
                raise
            infer = {"error": repr(ex)[:300]}
    ref_cuda = None
    if rank == 0 and not a.no_ref_cuda:
        try:
            ref_cuda = time_reference_cuda(dev, sites, 20, 3, a.grad_accum)
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
    kernels = ""
    if dom.startswith("bwd") and e == 2:
        if a.grad_accum == "tile":
            kernels = " (zero_fill_kernel + bwd_win_kernel)"
        elif a.grad_accum == "opmath":
            kernels = (" (memset + bwd_vec_kernel + cast_ws_kernel)" if os.environ.get("DCNV3_B200_BWD") == "vec" else
                       " (zero_select_kernel + bwd_imat_kernel [+ bwd_vec_kernel, skipped by the selector] + cast_ws_kernel)")
    roofline = {"bound": "hbm", "kernel": dom + kernels,
                "achieved": table[dom]["GBps"], "peak": peak, "unit": "GB/s", "frac": table[dom]["GBps"] / peak,
                "peak_source": peak_src, "traffic": ncu_traffic(dom),
                "traffic_source": "profiles/r02_ncu_traffic.json: ncu dram__bytes_read.sum + dram__bytes_write.sum summed over "
                                  "every kernel of the op, per call (a committed capture of this command, not measured by this run)",
                "step_frac": value / world / peak}
    out = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": a.dtype if a.dtype != "fp32" else "f32", "data": "synthetic",
           "config": config, "ops": table, "reference_cuda": ref_cuda, "seg_detail": seg_detail,
           "pct_hbm_peak": 100.0 * value / world / peak, "roofline": roofline, "gpu_launches": launches, "clocks": clk}
    if not a.no_cpu_baseline:
        gbps, cms, cores, sample, _ = cpu_run(sites, e, 3, 1)
        out["cpu_baseline"] = {"value": gbps, "unit": "GB/s", "cores": cores, "kind": "port",
                               "sample": sample, "ms_per_step_sample": cms}
    # last keys (they survive a tail of the line): end-to-end, configs[4] inference, the data-parallel training step
    out["e2e"] = e2e
    out["infer"] = infer
    out["seg_train"] = seg_res
    print(json.dumps(out), file=out_stream, flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
