# one full ncu capture of the backward kernel at P3 (after the same command exited 0 without ncu)
B2="python bench.py --steps 3 --warmup 3 --sites P3 --no-cpu-baseline --no-e2e --no-seg --no-ref-cuda --no-infer"
$B2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bwd_win -s 4 -c 1 -o gpurun_out/${1:-r2h}_bwd_win -f $B2 > gpurun_out/${1:-r2h}_ncu_b.log 2>&1
ls -la gpurun_out/${1:-r2h}_bwd_win.ncu-rep
