# Data-parallel seg step, variants of the gradient exchange on ONE box (run with gpurun --gpus N):
#   N=${N:-2} bash tools/ddp_ab.sh
N=${N:-2}
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-ref-cuda --no-cpu-baseline --no-infer --seg-steps ${SEG_STEPS:-15} "$@" \
    > gpurun_out/ddp_${N}_$name.json 2> gpurun_out/ddp_${N}_$name.err || { echo "$name FAILED"; tail -5 gpurun_out/ddp_${N}_$name.err; }
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/ddp_${N}_$name.json")); s = d["seg_train"]
    print("$name", {k: (round(v["imgs_per_s"], 1), round(v["ms_per_step"], 2), v["cuda_graph"]) for k, v in s.items() if isinstance(v, dict)},
          {k: v for k, v in s.items() if not isinstance(v, dict)})
except Exception as e:
    print("$name: no line", e)
P
}
run graph_bf16 X=1 --
run eager_fp32 X=1 -- --seg-graph off --ddp-compress none --no-dp-check
run eager_bf16 X=1 -- --seg-graph off --no-dp-check
run graph_fp32 X=1 -- --ddp-compress none --no-dp-check
run graph_bf16_b100 X=1 -- --ddp-bucket-mb 100 --no-dp-check --no-seg-strong
run graph_bf16_b8 X=1 -- --ddp-bucket-mb 8 --no-dp-check --no-seg-strong
