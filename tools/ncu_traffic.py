"""profiles/r02_ncu_traffic.json from an ncu launch list of the bench command (run here, no GPU):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"fwd_tile|zero_fill|bwd_win" -c 400 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-seg --no-ref-cuda --no-infer
    python tools/ncu_traffic.py gpurun_out/launches.csv profiles/r02_ncu_traffic.json

A step launches, in order: fwd_tile_kernel x3 (P3, P4, P5), then per backward (P5, P4, P3) zero_fill_kernel +
bwd_win_kernel.  The LAST complete step of the list is taken; an op's traffic is the sum over EVERY kernel it
launches."""
import collections, csv, io, json, sys

src, dst = sys.argv[1], sys.argv[2]
text = open(src).read()
text = text[text.index('"ID"'):]
launches = collections.OrderedDict()
for r in csv.DictReader(io.StringIO(text)):
    d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * (1e6 if r["Metric Unit"] == "Mbyte" else 1e3 if r["Metric Unit"] == "Kbyte" else 1)
seq = []
for i, d in launches.items():
    nm = d["name"]
    kind = "fwd" if "fwd_tile" in nm else "zero" if "zero_fill" in nm else "bwd" if "bwd_win" in nm else None
    if kind:
        seq.append((kind, d))
pattern = ["fwd"] * 3 + ["zero", "bwd"] * 3
last = None
for i in range(len(seq) - len(pattern), -1, -1):
    if [k for k, _ in seq[i:i + len(pattern)]] == pattern:
        last = seq[i:i + len(pattern)]
        break
if last is None:
    raise SystemExit("no complete step (fwd x3, (zero, bwd) x3) in the launch list")
ops = {"fwd_P3": [last[0]], "fwd_P4": [last[1]], "fwd_P5": [last[2]],
       "bwd_P5": last[3:5], "bwd_P4": last[5:7], "bwd_P3": last[7:9]}
out = {"source": f"{src}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
                 "python bench.py --steps 1 --warmup 3 (kernels only); the last complete step of the list; every kernel of an op summed"}
tot_t = sum(d["gpu__time_duration.sum"] for _, d in last)
for op, ks in ops.items():
    out[op] = {"dram_bytes_read": int(sum(d["dram__bytes_read.sum"] for _, d in ks)),
               "dram_bytes_write": int(sum(d["dram__bytes_write.sum"] for _, d in ks)),
               "ncu_us": sum(d["gpu__time_duration.sum"] for _, d in ks) / 1e3,
               "share_of_step": sum(d["gpu__time_duration.sum"] for _, d in ks) / tot_t,
               "kernels": [{"name": d["name"].split("(")[0].replace("void ", ""), "grid": d["grid"], "us": d["gpu__time_duration.sum"] / 1e3,
                            "dram_read": int(d["dram__bytes_read.sum"]), "dram_write": int(d["dram__bytes_write.sum"])} for _, d in ks]}
out["step_ncu_us"] = tot_t / 1e3
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "kernels"}) for k, v in out.items()}, indent=1))
