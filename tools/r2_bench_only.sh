set -x
S=$SECONDS; python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "default bench wall: $((SECONDS-S)) s rc=$?"; tail -c 1800 gpurun_out/r2j_bench.json
S=$SECONDS; python bench.py --impl reference --steps 500 --warmup 10 > gpurun_out/r2j_ref.json 2> gpurun_out/r2j_ref.err; echo "reference arm wall: $((SECONDS-S)) s"; tail -c 700 gpurun_out/r2j_ref.json
