#!/usr/bin/env bash
# developer helper: time build variants of the seg kernel (tools/_variants/*.so through LPBOX_LIB)
set -u
mkdir -p gpurun_out; : > gpurun_out/seg_var.log
for v in "$@"; do
  lib=${v%%:*}; T=${v##*:}; B=1024; [ "$T" = 256 ] && B=740
  echo "== $lib T=$T B=$B" | tee -a gpurun_out/seg_var.log
  LPBOX_LIB=$PWD/tools/_variants/$lib.so LPBOX_SEG_T=$T python tools/quick_bench_seg.py $B 375 500 80 2>&1 | tail -3 | grep "B=\|algorithmic" | tee -a gpurun_out/seg_var.log
done
