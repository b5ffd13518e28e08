L=yolo_dual_b200/csrc/libdcnv3_b200.so
cp $L /tmp/base.so
for rep in 1 2; do for v in base variant; do
if [ $v = variant ]; then cp tools/libvariant.bin $L; else cp /tmp/base.so $L; fi
python bench.py --no-seg --no-cpu-baseline --no-e2e --no-ref-cuda > gpurun_out/b_$v.json 2> gpurun_out/b.err
python - <<P
import json
d=json.load(open("gpurun_out/b_$v.json")); print("$v", d["ms_per_step"], {k:round(v["us_median"],1) for k,v in d["ops"].items()})
P
done; done
cp /tmp/base.so $L
