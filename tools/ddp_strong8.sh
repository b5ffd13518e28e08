# configs[3] at N GPUs (strong scaling, global batch 64): gradient-exchange variants on ONE box
N=${N:-8}
run() {
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-ref-cuda --no-cpu-baseline --no-infer --no-dp-check --seg-steps 20 \
    --seg-model yolov8seg --seg-global-batch 64 "$@" > gpurun_out/strong_${N}_$name.json 2> gpurun_out/strong_${N}_$name.err || { echo "$name FAILED"; tail -5 gpurun_out/strong_${N}_$name.err; }
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/strong_${N}_$name.json")); s = d["seg_train"]
    print("$name", {k: (round(v["imgs_per_s"], 1), round(v["ms_per_step"], 2), v["cuda_graph"]) for k, v in s.items() if isinstance(v, dict)})
except Exception as e:
    print("$name: no line", e)
P
}
run fp32 X=1 --
run bf16 X=1 -- --ddp-compress bf16
run bf16_b100 X=1 -- --ddp-compress bf16 --ddp-bucket-mb 100
run fp32_b8 X=1 -- --ddp-bucket-mb 8
