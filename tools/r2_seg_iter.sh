timeout 300 python -m pytest tests/test_bnact_gpu.py tests/test_segloss_gpu.py tests/test_dcnv3_gpu.py -q -k "bn or seg or lowres or cat" 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 3 --no-e2e --no-ref-cuda --no-cpu-baseline --seg-steps 20 > gpurun_out/seg_iter.json 2> gpurun_out/seg_iter.err || tail -5 gpurun_out/seg_iter.err
python -c "
import json; d=json.load(open('gpurun_out/seg_iter.json')); print({k:(round(v['imgs_per_s'],1), round(v['ms_per_step'],2)) for k,v in d['seg_train'].items() if isinstance(v,dict)}, 'infer', round(d['infer']['imgs_per_s'],1), round(d['infer']['ms_per_step'],2))"
