"""Per-source-line digest of an .ncu-rep (stall samples, warp instructions, shared wavefronts), joined
to CUDA source lines through nvdisasm --print-line-info of the cubin inside the built library.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep <mangled-kernel-substring> [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "yolo_dual_b200", "csrc", "libdcnv3_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# address -> line for the kernel's section (device functions it calls live in the same section)
line_of = {}
insec = False; cur = None
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        insec = kname in l
        cur = None
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
base = int(body[0][ix["Address"]], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.defaultdict(collections.Counter)
T = 0
for r in body:
    a = int(r[ix["Address"]], 16) - base
    k = line_of.get(a)
    s = int(r[ix["# Samples"]] or 0)
    agg[k][0] += s; T += s
    agg[k][1] += int(r[ix["Instructions Executed"]] or 0)
    agg[k][2] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v: stalls[k][c[6:]] += v
srcs = {}
def text(k):
    if not k: return ""
    f = os.path.join(ROOT, "yolo_dual_b200", "csrc", k[0])
    if f not in srcs:
        srcs[f] = open(f).read().splitlines() if os.path.exists(f) else []
    L = srcs[f]
    return L[k[1] - 1].strip()[:90] if 0 < k[1] <= len(L) else ""
print("total samples", T)
for k, (s, n, w) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join(f"{a}:{b}" for a, b in stalls[k].most_common(3))
    print(f"{100*s/T:5.1f}% smp {s:6d} inst {n:9d} wf {w:9d}  {k[0] if k else '?'}:{k[1] if k else 0:4d} [{st}] {text(k)}")
if len(sys.argv) > 4:  # phase table: "name:lo-hi,name:lo-hi" over dcnv3_imat.cuh lines
    for spec in sys.argv[4].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        s = sum(v[0] for k, v in agg.items() if k and k[0] == "dcnv3_imat.cuh" and lo <= k[1] <= hi)
        n = sum(v[1] for k, v in agg.items() if k and k[0] == "dcnv3_imat.cuh" and lo <= k[1] <= hi)
        w = sum(v[2] for k, v in agg.items() if k and k[0] == "dcnv3_imat.cuh" and lo <= k[1] <= hi)
        print(f"phase {name:14s} {100*s/T:5.1f}% samples, {n:9d} inst, {w:9d} smem wf")
