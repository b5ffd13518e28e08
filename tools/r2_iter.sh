# quick kernel iteration: win-kernel parity + per-op times (+ optional ncu capture of bwd_win at P3: NCU=1)
set -x
timeout 300 python -m pytest tests/test_win_gpu.py tests/test_imat_gpu.py -x -q > gpurun_out/it_pytest.log 2>&1; tail -3 gpurun_out/it_pytest.log
for i in 1 2; do timeout 120 python -m pytest "tests/test_win_gpu.py::test_win_vs_pixel_oracle" -x -q -k "scale1.5" 2>&1 | tail -1; done
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-seg --no-ref-cuda --no-infer"
timeout 300 $B > gpurun_out/it_ops.json 2> gpurun_out/it_ops.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/it_ops.json'))
print('step_us', d['ms_per_step']*1e3, 'GBps', d['value'])
print({k: round(v['us_mean'],1) for k,v in d['ops'].items()})
PY
if [ -n "$NCU" ]; then
B2="python bench.py --steps 3 --warmup 3 --sites P3 --no-cpu-baseline --no-e2e --no-seg --no-ref-cuda --no-infer"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_win -s 4 -c 1 -o gpurun_out/it_bwd -f $B2 > gpurun_out/it_ncu.log 2>&1; tail -2 gpurun_out/it_ncu.log
fi
