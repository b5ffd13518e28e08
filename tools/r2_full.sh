set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; tail -4 gpurun_out/r2b_pytest.log
timeout 300 python tools/module_probe.py > gpurun_out/r2b_module_probe.md 2> gpurun_out/r2b_module_probe.err; cat gpurun_out/r2b_module_probe.md; tail -3 gpurun_out/r2b_module_probe.err
