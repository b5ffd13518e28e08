# Round-end evidence: full bench line, reference arm, ncu launch list and full captures (bwd_imat, fwd_tile).
# Each ncu command runs only after the same command exited 0 without ncu.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-seg --no-ref-cuda"
$B > gpurun_out/pre.json 2> gpurun_out/pre.err && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vec_kernel|cast_ws|imat|zero_select|fwd_tile" -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_l.log 2>&1
B2="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-seg --no-ref-cuda"
$B2 > gpurun_out/pre2.json 2> gpurun_out/pre2.err && ncu --set full --clock-control none --import-source on -k regex:bwd_imat -s 8 -c 1 -o gpurun_out/imat_final -f $B2 > gpurun_out/ncu_f.log 2>&1
$B2 > gpurun_out/pre3.json 2> gpurun_out/pre3.err && ncu --set full --clock-control none --import-source on -k regex:fwd_tile -s 6 -c 1 -o gpurun_out/fwd_tile_final -f $B2 > gpurun_out/ncu_f2.log 2>&1
tail -c 300 gpurun_out/bench_full.json; tail -c 300 gpurun_out/bench_ref.json
