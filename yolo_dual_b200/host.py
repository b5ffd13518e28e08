"""Host-buffer front of the DCNv3 core: forward + backward for tensors that live in (pinned) host
memory, with the host->device copies, the kernels and the device->host copies overlapped on three
CUDA streams and double-buffered device staging.

The reference's op only takes CUDA tensors (src/cpu/dcnv3_cpu.cpp:25,36 throw); a caller with host data
does `.cuda()` / `.cpu()` around `DCNv3Function.apply`, which serialises PCIe in, compute and PCIe out.
`HostPipeline` is that same sequence — the compute still goes through `DCNv3Function.apply` and autograd —
arranged so that step k's results stream out while step k+1's inputs stream in (PCIe is full duplex).

    pipe = HostPipeline(device)
    t = pipe.submit(sites)      # sites: list of HostSite (pinned input/offset/mask/grad_out + output buffers)
    pipe.wait(t)                # that step's output / grad_input / grad_offset / grad_mask are in host memory

Steps in flight share nothing but the device staging slots; give each in-flight step its own HostSite
output buffers if its results must survive the next submit().
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import torch

from .ops_dcnv3.functions import DCNv3Function, DCNv3SoftmaxFunction


@dataclass
class HostSite:
    """One DCNv3 call site with host-resident tensors (pinned memory recommended).
    `args` = (kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
    group_channels, offset_scale) — the reference's argument order."""
    input: torch.Tensor
    offset: torch.Tensor
    mask: torch.Tensor
    grad_out: torch.Tensor
    args: Sequence
    output: torch.Tensor = None
    grad_input: torch.Tensor = None
    grad_offset: torch.Tensor = None
    grad_mask: torch.Tensor = None
    _dev: list = field(default_factory=list, repr=False)

    def alloc_outputs(self, out_shape):
        mk = lambda like, shape=None: torch.empty(shape or like.shape, dtype=like.dtype).pin_memory()
        self.output = mk(self.input, out_shape)
        self.grad_input, self.grad_offset, self.grad_mask = mk(self.input), mk(self.offset), mk(self.mask)
        return self

    @property
    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.input, self.offset, self.mask, self.grad_out))

    @property
    def d2h_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.output, self.grad_input, self.grad_offset, self.grad_mask))


class HostPipeline:
    def __init__(self, device, depth: int = 3, fused_softmax: bool = False):
        self.device = torch.device(device)
        self.depth = depth
        self.fn = DCNv3SoftmaxFunction if fused_softmax else DCNv3Function
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.slots = [dict(staging=None, ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event(),
                           ev_out=torch.cuda.Event(), busy=False) for _ in range(depth)]
        self.step = 0

    def _staging(self, slot, sites: List[HostSite]):
        if slot["staging"] is None:
            slot["staging"] = [[torch.empty(t.shape, dtype=t.dtype, device=self.device)
                                for t in (s.input, s.offset, s.mask, s.grad_out)] for s in sites]
        return slot["staging"]

    def submit(self, sites: List[HostSite]) -> int:
        """Enqueue one forward+backward step over `sites`; returns a ticket for wait()."""
        k = self.step
        self.step += 1
        slot = self.slots[k % self.depth]
        cur = torch.cuda.current_stream(self.device)
        if slot["busy"]:
            # the staging buffers of this slot were last read by the kernels of step k - depth, and the
            # host output buffers last written by its device->host copies
            self.s_in.wait_event(slot["ev_done"])
            slot["ev_out"].synchronize()
        staging = self._staging(slot, sites)
        # ---- host -> device on the copy-in stream
        with torch.cuda.stream(self.s_in), torch.no_grad():
            for s, dev in zip(sites, staging):
                for h, d in zip((s.input, s.offset, s.mask, s.grad_out), dev):
                    d.requires_grad_(False)
                    d.copy_(h, non_blocking=True)
        slot["ev_in"].record(self.s_in)
        # ---- kernels on the caller's stream, through the public autograd API
        cur.wait_event(slot["ev_in"])
        live = []
        for s, (x, off, m, go) in zip(sites, staging):
            x.grad = off.grad = m.grad = None
            x.requires_grad_(True); off.requires_grad_(True); m.requires_grad_(True)
            y = self.fn.apply(x, off, m, *s.args, 256)
            live.append((s, x, off, m, go, y))
        results = []
        for s, x, off, m, go, y in reversed(live):  # backward in reverse order, as a training step does
            y.backward(go)
            results.append((s, y.detach(), x.grad, off.grad, m.grad))
        slot["ev_done"].record(cur)
        # ---- device -> host on the copy-out stream
        self.s_out.wait_event(slot["ev_done"])
        with torch.cuda.stream(self.s_out):
            for s, y, gi, go_, gm in results:
                for d, h in ((y, s.output), (gi, s.grad_input), (go_, s.grad_offset), (gm, s.grad_mask)):
                    d.record_stream(self.s_out)
                    h.copy_(d, non_blocking=True)
        slot["ev_out"].record(self.s_out)
        slot["busy"] = True
        return k

    def wait(self, ticket: int) -> None:
        """Block until the results of step `ticket` are in the host output buffers."""
        self.slots[ticket % self.depth]["ev_out"].synchronize()

    def drain(self) -> None:
        for slot in self.slots:
            if slot["busy"]:
                slot["ev_out"].synchronize()
