"""Per-source-line digest of one kernel launch inside an .ncu-rep holding several (stall samples, warp instructions,
shared wavefronts), joined to CUDA source lines through nvdisasm --print-line-info of the cubin in the built library.
    python tools/ncu_lines2.py gpurun_out/x.ncu-rep <mangled-kernel-substring> <launch index in the report> [top] [sort: samples|instr]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kname, which = sys.argv[1], sys.argv[2], int(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sort = sys.argv[5] if len(sys.argv) > 5 else "samples"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("NCU_LINES_LIB") or os.path.join(ROOT, "yolo_dual_b200", "csrc", "libdcnv3_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
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
secs = []; cur = None
for ln in src.split("\n"):
    if ln.startswith('"Kernel Name"'):
        cur = [ln]; secs.append(cur)
    elif cur is not None:
        cur.append(ln)
sec = secs[which]
print(sec[0][:120])
rows = list(csv.reader(sec[1:]))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[1:] if len(r) >= len(hdr)]
base = int(body[0][ix["Address"]], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.defaultdict(collections.Counter)
T = 0; NI = 0
for r in body:
    a = int(r[ix["Address"]], 16) - base
    k = line_of.get(a)
    s = int(r[ix["# Samples"]] or 0)
    agg[k][0] += s; T += s
    n = int(r[ix["Instructions Executed"]] or 0)
    agg[k][1] += n; NI += n
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
    return L[k[1] - 1].strip()[:100] if 0 < k[1] <= len(L) else ""
print("total samples", T, "warp instructions", NI)
key = (lambda kv: -kv[1][0]) if sort == "samples" else (lambda kv: -kv[1][1])
for k, (s, n, w) in sorted(agg.items(), key=key)[:top]:
    st = ",".join(f"{a}:{b}" for a, b in stalls[k].most_common(3))
    print(f"{100*s/max(T,1):5.1f}% smp {100*n/max(NI,1):5.1f}% ins wf {w:8d} {str(k):34s} {text(k)}  [{st}]")
