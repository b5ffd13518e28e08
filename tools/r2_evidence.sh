# Round-2 evidence: smoke, full GPU tests, the default bench line (timed), the reference arm, launch list with DRAM traffic,
# full ncu captures of the two hot kernels.  Each ncu command runs only after the same command exited 0 without ncu.
set -x
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log
S=$SECONDS; python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "default bench wall: $((SECONDS-S)) s"; tail -c 1500 gpurun_out/r2m_bench.json
S=$SECONDS; python bench.py --impl reference --steps 500 --warmup 10 > gpurun_out/r2m_ref.json 2> gpurun_out/r2m_ref.err; echo "reference arm wall: $((SECONDS-S)) s"; tail -c 600 gpurun_out/r2m_ref.json
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-seg --no-ref-cuda --no-infer"
$B1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fwd_tile|zero_fill|bwd_win" -c 400 --csv --log-file gpurun_out/r2m_launches.csv $B1 > gpurun_out/r2m_ncu_l.log 2>&1
B2="python bench.py --steps 3 --warmup 3 --sites P3 --no-cpu-baseline --no-e2e --no-seg --no-ref-cuda --no-infer"
$B2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bwd_win -s 4 -c 1 -o gpurun_out/r2m_bwd_win -f $B2 > gpurun_out/r2m_ncu_b.log 2>&1
$B2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fwd_tile -s 4 -c 1 -o gpurun_out/r2m_fwd_tile -f $B2 > gpurun_out/r2m_ncu_f.log 2>&1
ls -la gpurun_out/r2m_*
