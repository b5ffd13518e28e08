#!/bin/bash
# Slow-path experiment (profiles/r02_slow_path.md): parity of the variant library, A/B on one box, and the bound
# (offsets clipped to +-2.9 px: no point leaves the window / the pixel's reach)
L=yolo_dual_b200/csrc/libdcnv3_b200.so
cp $L /tmp/orig.so
for v in d_coop; do
  cp tools/var_$v.bin $L
  python -m pytest tests/test_win_gpu.py tests/test_dcnv3_gpu.py -m gpu -x -q 2>&1 | tail -2 | sed "s/^/$v: /"
done
cp /tmp/orig.so $L
STEPS=200 bash tools/ab_multi.sh
for v in a_base d_coop; do
cp tools/var_$v.bin $L
for clip in 2.9; do
BENCH_OFFSET_CLIP=$clip python bench.py --steps 200 --warmup 5 --no-seg --no-cpu-baseline --no-e2e --no-ref-cuda --no-infer > gpurun_out/clip_$clip.json 2> gpurun_out/clip.err || tail -3 gpurun_out/clip.err
python - <<P
import json
d=json.load(open("gpurun_out/clip_$clip.json")); print("$v clip=$clip", round(d["ms_per_step"],4), {k:round(x["us_median"],1) for k,x in d["ops"].items()})
P
done; done
cp /tmp/orig.so $L
