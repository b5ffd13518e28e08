#!/bin/bash
# build a variant of libdcnv3_b200.so with extra nvcc flags into tools/var_<name>.bin (A/B runs on one box: tools/ab_multi.sh)
#   tools/build_variant.sh <name> [-DFLAG ...]
name=$1; shift
exec nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I include "$@" \
  -o tools/var_$name.bin yolo_dual_b200/csrc/dcnv3_b200.cu
