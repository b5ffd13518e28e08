"""Raw PCIe ceiling for the e2e line: pinned host <-> device copies, one direction alone and both at once
(two streams), CUDA-event timed.  python tools/pcie_probe.py [MB]"""
import sys
import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 169
n = mb * 1000 * 1000
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


CH = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # copies per step and direction
def pieces(t):
    return list(t.chunk(CH))


def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                for a, b in zip(pieces(d_in), pieces(h_in)):
                    a.copy_(b, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                for a, b in zip(pieces(h_out), pieces(d_out)):
                    a.copy_(b, non_blocking=True)
    cur = torch.cuda.current_stream()
    cur.wait_stream(s1); cur.wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, a, b in (("h2d alone", 1, 0), ("d2h alone", 0, 1), ("both", 1, 1)):
    run(a, b, 2)
    ms = run(a, b)
    print(f"{name:10s} {ms:7.3f} ms per {mb} MB  -> {n / ms / 1e6:6.1f} GB/s per direction")
