# 1 GPU: the per-rank work of configs[3] at 8 GPUs (yolov8seg, 8 images) without any gradient exchange
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-ref-cuda --no-cpu-baseline --no-infer --seg-steps 20 --seg-model yolov8seg --seg-batch 8 --no-seg-strong > gpurun_out/seg_b8.json 2> gpurun_out/seg_b8.err
python -c "
import json; d=json.load(open('gpurun_out/seg_b8.json')); print(d['seg_train'])"
