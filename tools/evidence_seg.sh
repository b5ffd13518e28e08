# Evidence for the seg-step kernels (bnact / resize / segloss): GPU tests, the default bench line, and an ncu launch
# list with DRAM bytes and throughput of those kernels over one profiled stretch of the yolov5seg training step.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python tools/seg_probe.py yolov5seg > gpurun_out/seg_probe_pre.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"stats_kernel|apply_kernel|bwd_reduce_kernel|finalize_kernel|fwd_kernel|bwd_kernel|segloss" \
    -s 1300 -c 330 --csv --log-file gpurun_out/seg_launches.csv python tools/seg_probe.py yolov5seg > gpurun_out/ncu_seg.log 2>&1
tail -c 400 gpurun_out/bench_full.json
