L=yolo_dual_b200/csrc/libdcnv3_b200.so
cp $L /tmp/orig.so; cp tools/var_j_dbg.bin $L
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/dbg_rmap.py 2>&1 | grep -v "^  File\|^    " | head -60
cp /tmp/orig.so $L
