set -x
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
S=$SECONDS; python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "default bench wall: $((SECONDS-S)) s"; tail -c 900 gpurun_out/r2g_bench.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err; tail -c 300 gpurun_out/r2g_ref.json
