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
print("Blackwell / Hopper-class mnemonics: the windows of `bwd_win_kernel` and `fwd_tile_kernel` are TMA box loads (`UTMALDG.4D`, completion")
print("on an mbarrier: `SYNCS.EXCH` / `SYNCS.ARRIVE.TRANS64`); offsets / masks are staged with `LDGSTS` (cp.async); gathers are `LDS.128`, the")
print("multiplies `FHFMA` (mixed-precision FMA, sm_100) / `HMMA.16816` (mma.sync), the reductions `REDG.E.ADD.BF16x8 / F16x8` vectors.  No")
print("`UTCHMMA` / `LDTM` (tcgen05 / TMEM): DESIGN.md par. 4.0 and 7-0 say where they would pay.\n")
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
