#!/bin/bash
# parity of every variant library (win + dcnv3 tests), then A/B on one box (tools/ab_multi.sh)
L=yolo_dual_b200/csrc/libdcnv3_b200.so
cp $L /tmp/orig.so
for f in tools/var_*.bin; do
  v=$(basename $f .bin); [ "$v" = "${SKIP_TEST:-var_a}"* ] && continue
  case $v in var_a*) continue;; esac
  cp $f $L
  python -m pytest tests/test_win_gpu.py tests/test_dcnv3_gpu.py tests/test_imat_gpu.py tests/test_infer_gpu.py tests/test_packed_gpu.py -m gpu -x -q 2>&1 | tail -2 | sed "s/^/$v: /"
done
cp /tmp/orig.so $L
STEPS=${STEPS:-200} bash tools/ab_multi.sh
