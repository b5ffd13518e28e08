set -x
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 300 --warmup 10 --no-seg --no-cpu-baseline --no-e2e --no-ref-cuda --no-infer > gpurun_out/r2n_quick.json 2>gpurun_out/r2n_quick.err; python - <<P
import json
d=json.load(open("gpurun_out/r2n_quick.json")); print("final lib", round(d["value"],1), round(d["ms_per_step"],4), {k:round(x["us_median"],1) for k,x in d["ops"].items()})
P
