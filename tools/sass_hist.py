"""Opcode histograms of the hot kernels from the built library (no GPU): cuobjdump -sass, grouped per kernel.
    python tools/sass_hist.py > profiles/r02_sass_histograms.md"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "yolo_dual_b200", "csrc", "libdcnv3_b200.so")
want = sys.argv[1:] or ["bwd_win_kernelI13__nv_bfloat16Lb0", "fwd_tile_kernelI13__nv_bfloat16Lb0", "zero_fill_kernel",
                        "bwd_imat_kernelI13__nv_bfloat16Lb0", "bwd_vec_kernelI13__nv_bfloat16fLi8ELi9ELb0", "fwd_vec_kernelI13__nv_bfloat16Li16ELi9ELb0"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        key = op if op.startswith(("LDS", "STS", "LDSM", "STSM", "LDG", "STG", "RED", "ATOM", "HMMA", "LDGSTS", "FHFMA", "UTC", "LDTM", "STTM", "UTMA", "BAR", "SHFL")) else op.split(".")[0]
        hist[cur][key] += 1
print("# SASS opcode histograms of the hot kernels (static instruction counts, `cuobjdump -sass` of the built library)\n")
print("Blackwell / Hopper-class mnemonics.  TMA does the bulk data movement of both kernels: the windows are box loads (`UTMALDG.4D`,")
print("completion on an mbarrier: `SYNCS.EXCH` / `SYNCS.ARRIVE.TRANS64`), so are the tile's offsets and masks (two more `UTMALDG.4D`; `LDGSTS` =")
print("cp.async only where their rows are not 16-byte multiples and in strip tiles); the backward's interpolation matrix is zeroed by a bulk copy")
print("(`UBLKCP`) and its `grad_input` window leaves as one TMA reduce-add per warp (`UTMAREDG.4D.ADD` + `UTMACMDFLUSH`; `REDG.E.ADD.BF16x8 / F16x8`")
print("vectors only for boxes that start at a negative coordinate, strip tiles and rare points).  Gathers are `LDS.128`, the multiplies `FHFMA`")
print("(mixed-precision FMA, sm_100) / `HMMA.16816` (mma.sync).  No `UTCHMMA` / `LDTM` (tcgen05 / TMEM): DESIGN.md par. 4.0 and 7-0 say where")
print("they would pay.\n")
for fn, h in hist.items():
    if not any(w in fn for w in want):
        continue
    n = sum(h.values())
    print(f"## `{fn}` — {n} instructions\n")
    notable = {k: v for k, v in h.items() if k.startswith(("UTMA", "SYNCS", "HMMA", "FHFMA", "RED", "LDGSTS", "LDSM", "STSM", "UTC", "LDTM", "UBLKCP"))}
    print("notable: " + ", ".join(f"`{k}` x{v}" for k, v in sorted(notable.items())) + "\n")
    print("| opcode | count | % |\n|---|---|---|")
    for op, c in h.most_common(32):
        print(f"| `{op}` | {c} | {100 * c / n:.1f} |")
    print()
