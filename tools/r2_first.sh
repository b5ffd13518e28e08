# round-2 first GPU call: tests + bench (single GPU) + launch list with DRAM traffic
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -3 gpurun_out/r2a_pytest.log
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-seg --no-ref-cuda --no-infer"
$B > gpurun_out/r2a_ops.json 2> gpurun_out/r2a_ops.err; tail -c 1500 gpurun_out/r2a_ops.json
timeout 900 python bench.py --steps 100 --warmup 5 --no-e2e --no-ref-cuda --no-cpu-baseline > gpurun_out/r2a_seg.json 2> gpurun_out/r2a_seg.err; tail -c 1200 gpurun_out/r2a_seg.json; tail -5 gpurun_out/r2a_seg.err
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-seg --no-ref-cuda --no-infer"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fwd_tile|zero_fill|bwd_win" -c 400 --csv --log-file gpurun_out/r2a_launches.csv $B1 > gpurun_out/r2a_ncu_l.log 2>&1
tail -2 gpurun_out/r2a_launches.csv | cut -c1-300
