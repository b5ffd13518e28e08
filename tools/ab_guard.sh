for rep in 1 2; do for v in 1 2 4 8 32; do
DCNV3_B200_GUARD_PER_CTA=$v python bench.py --no-seg --no-cpu-baseline --no-e2e --no-ref-cuda > gpurun_out/b_$v.json 2> gpurun_out/b.err
python - <<P
import json
d=json.load(open("gpurun_out/b_$v.json")); print("per_cta $v", d["ms_per_step"], {k:round(v["us_median"],1) for k,v in d["ops"].items()})
P
done; done
for v in 1 2 4 8 32; do echo fallback per_cta $v; DCNV3_B200_GUARD_PER_CTA=$v IMAT_OFFSCALE=2 python tools/imat_check.py bwd time 2>&1 | grep "^P3 auto"; done
