L=yolo_dual_b200/csrc/libdcnv3_b200.so
cp $L /tmp/orig.so; cp tools/var_b_tstage.bin $L
CUDA_LAUNCH_BLOCKING=1 python -m pytest tests/test_win_gpu.py -m gpu -x -q 2>&1 | grep -E "^(FAILED|ERROR|tests/|E  )" | head -12
cp /tmp/orig.so $L
