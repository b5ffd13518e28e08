"""Digest an .ncu-rep here (no GPU): key raw metrics, opcode mix and hot SASS regions.
    python tools/ncu_digest.py gpurun_out/x.ncu-rep"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_global_red.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
for k in want:
    if k in d:
        print(f"{k:80s} {d[k][0]:>16s} {d[k][1]}")
for h in hdr:
    if "issue_stalled" in h and "per_issue_active" in h:
        print(f"{h:80s} {d[h][0]:>16s}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
tot = collections.Counter(); wf = collections.Counter(); N = 0
for r in body:
    t = r[ix["Source"]].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0] if not op.startswith(("LDS", "STS", "LDSM", "LDG", "STG", "HMMA", "RED", "ATOM")) else ".".join(op.split(".")[:3])
    n = int(r[ix["Instructions Executed"]] or 0)
    tot[op] += n; N += n
    wf[op] += int(r[ix["L1 Wavefronts Shared"]] or 0)
print("total warp instructions", N)
for op, n in tot.most_common(28):
    print(f"  {op:22s} {n:10d} {100*n/N:5.1f}%  smem wavefronts {wf[op]:9d}")
cur = None; start = 0; acc = 0; out = []
for i, r in enumerate(body):
    n = int(r[ix["Instructions Executed"]] or 0)
    if n != cur:
        if cur is not None: out.append((start, i - 1, cur, acc))
        cur = n; start = i; acc = 0
    acc += n
out.append((start, len(body) - 1, cur, acc))
print("hot SASS regions (>1.5%)")
for s, e, c, a in out:
    if a > N * 0.015:
        w = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in body[s:e + 1])
        smp = sum(int(r[ix["# Samples"]] or 0) for r in body[s:e + 1])
        print(f"  lines {s}-{e} ({e-s+1} instr) x{c} = {a} ({100*a/N:.1f}%) smem wf {w} samples {smp}")
