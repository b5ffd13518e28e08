# one default data-parallel bench line (dp check + Trainer under NCCL + weak + strong + infer + e2e) on N GPUs
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus $N --steps 50 --warmup 5 --no-ref-cuda --no-cpu-baseline $EXTRA > gpurun_out/ddp_${N}_default.json 2> gpurun_out/ddp_${N}_default.err || tail -20 gpurun_out/ddp_${N}_default.err
tail -c 2500 gpurun_out/ddp_${N}_default.json
