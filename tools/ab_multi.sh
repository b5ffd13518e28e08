#!/bin/bash
# A/B/C... on ONE box: every tools/var_*.bin in turn as the library, two rounds, quick bench (ops table)
L=yolo_dual_b200/csrc/libdcnv3_b200.so
cp $L /tmp/base.so
for rep in 1 2; do for f in tools/var_*.bin; do
v=$(basename $f .bin); cp $f $L
python bench.py --steps ${STEPS:-200} --warmup 5 --no-seg --no-cpu-baseline --no-e2e --no-ref-cuda --no-infer $BENCH_ARGS > gpurun_out/ab_$v.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python - <<P
import json
d=json.load(open("gpurun_out/ab_$v.json")); print("$v", round(d["ms_per_step"],4), {k:round(x["us_median"],1) for k,x in d["ops"].items()})
P
done; done
cp /tmp/base.so $L
