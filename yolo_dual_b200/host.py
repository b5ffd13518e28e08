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

`pack_sites(sites)` re-homes the tensors of a list of sites into ONE pinned arena per direction.  submit() then
moves a step's inputs with one host->device copy and its results with one device->host copy (the results are
first gathered into a device arena, ~0.1 ms of device time at 169 MB): twelve separate copies per direction
cost ~10 % of the PCIe time of the 640x640 step (tools/pcie_probe.py).
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
    _arena: object = field(default=None, repr=False)   # set by pack_sites(): (in_arena, out_arena, in_offsets, out_offsets)

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


def _carve(arena: torch.Tensor, like: torch.Tensor, shape, offset: int) -> torch.Tensor:
    n = 1
    for d in shape:
        n *= d
    return arena[offset:offset + n * like.element_size()].view(like.dtype).view(shape)


def _aligned(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


def pack_sites(sites: List["HostSite"], out_shapes=None) -> List["HostSite"]:
    """New HostSites whose input tensors are views of one pinned arena and whose outputs are views of another
    (inputs are copied in; outputs are allocated).  `out_shapes[i]` defaults to the shape of `sites[i].output`."""
    ins, outs, n_in, n_out = [], [], 0, 0
    for i, s in enumerate(sites):
        oshape = tuple(out_shapes[i]) if out_shapes is not None else tuple(s.output.shape)
        io, oo = [], []
        for t in (s.input, s.offset, s.mask, s.grad_out):
            io.append(n_in); n_in += _aligned(t.numel() * t.element_size())
        for t, shape in ((s.input, oshape), (s.input, tuple(s.input.shape)), (s.offset, tuple(s.offset.shape)),
                         (s.mask, tuple(s.mask.shape))):
            oo.append((n_out, shape)); n_out += _aligned(t.element_size() * int(torch.Size(shape).numel()))
        ins.append(io); outs.append(oo)
    a_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    a_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    packed = []
    for s, io, oo in zip(sites, ins, outs):
        src = (s.input, s.offset, s.mask, s.grad_out)
        v_in = [_carve(a_in, t, tuple(t.shape), o) for t, o in zip(src, io)]
        for v, t in zip(v_in, src):
            v.copy_(t)
        like = (s.input, s.input, s.offset, s.mask)
        v_out = [_carve(a_out, t, shape, o) for t, (o, shape) in zip(like, oo)]
        ns = HostSite(*v_in, args=s.args, output=v_out[0], grad_input=v_out[1], grad_offset=v_out[2], grad_mask=v_out[3])
        ns._arena = (a_in, a_out, io, [o for o, _ in oo])
        packed.append(ns)
    return packed


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

    @staticmethod
    def _arena_of(sites: List[HostSite]):
        a = sites[0]._arena if sites else None
        if a is None or any(s._arena is None or s._arena[0] is not a[0] or s._arena[1] is not a[1] for s in sites):
            return None
        return a[0], a[1]

    def _staging(self, slot, sites: List[HostSite]):
        if slot["staging"] is None:
            arena = self._arena_of(sites)
            if arena is None:
                slot["staging"] = [[torch.empty(t.shape, dtype=t.dtype, device=self.device)
                                    for t in (s.input, s.offset, s.mask, s.grad_out)] for s in sites]
            else:  # one device arena per direction; the per-tensor staging buffers are views of it
                slot["dev_in"] = torch.empty(arena[0].numel(), dtype=torch.uint8, device=self.device)
                slot["dev_out"] = torch.empty(arena[1].numel(), dtype=torch.uint8, device=self.device)
                slot["staging"] = [[_carve(slot["dev_in"], t, tuple(t.shape), o)
                                    for t, o in zip((s.input, s.offset, s.mask, s.grad_out), s._arena[2])] for s in sites]
                slot["gather"] = [[_carve(slot["dev_out"], t, tuple(t.shape), o)
                                   for t, o in zip((s.output, s.grad_input, s.grad_offset, s.grad_mask), s._arena[3])]
                                  for s in sites]
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
        arena = self._arena_of(sites) if "dev_in" in slot else None
        with torch.cuda.stream(self.s_in), torch.no_grad():
            for dev in staging:
                for d in dev:
                    d.requires_grad_(False)
            if arena is not None:
                slot["dev_in"].copy_(arena[0], non_blocking=True)
            else:
                for s, dev in zip(sites, staging):
                    for h, d in zip((s.input, s.offset, s.mask, s.grad_out), dev):
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
        if arena is not None:  # gather the results into the device arena (device-to-device, caller's stream)
            by_site = {id(s): r for s, *r in results}
            with torch.no_grad():
                for s, views in zip(sites, slot["gather"]):
                    for v, d in zip(views, by_site[id(s)]):
                        v.copy_(d)
        slot["ev_done"].record(cur)
        # ---- device -> host on the copy-out stream
        self.s_out.wait_event(slot["ev_done"])
        with torch.cuda.stream(self.s_out):
            if arena is not None:
                arena[1].copy_(slot["dev_out"], non_blocking=True)
            else:
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
