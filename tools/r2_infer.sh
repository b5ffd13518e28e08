timeout 300 python -m pytest tests/test_bnact_gpu.py tests/test_infer_gpu.py -q 2>&1 | tail -3
timeout 300 python tools/infer_probe.py 2>&1 | grep -v -i warn | head -14
