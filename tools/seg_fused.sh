for extra in "" "--seg-fused-heads"; do
timeout 400 python bench.py --steps 20 --warmup 3 --no-e2e --no-ref-cuda --no-cpu-baseline --no-infer --seg-steps 20 $extra > gpurun_out/seg_fused.json 2> gpurun_out/seg_fused.err || tail -5 gpurun_out/seg_fused.err
python -c "
import json; d=json.load(open('gpurun_out/seg_fused.json')); print('$extra', {k:(round(v['imgs_per_s'],1), round(v['ms_per_step'],2)) for k,v in d['seg_train'].items() if isinstance(v,dict)})"
done
