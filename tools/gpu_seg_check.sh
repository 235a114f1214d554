#!/usr/bin/env bash
# developer helper: segmentation parity tests + timing of the three launch shapes (run through gpurun)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_seg_parity_gpu.py -x -q -m gpu > gpurun_out/seg_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/seg_tests.log
tail -5 gpurun_out/seg_tests.log
: > gpurun_out/seg_qb.log
for T in 256 192 160; do
  echo "== LPBOX_SEG_T=$T 1024 images x ${1:-80} iterations" | tee -a gpurun_out/seg_qb.log
  LPBOX_SEG_T=$T python tools/quick_bench_seg.py 1024 375 500 ${1:-80} 2>&1 | tail -3 | tee -a gpurun_out/seg_qb.log
done
echo "== T=256 740 images" | tee -a gpurun_out/seg_qb.log
LPBOX_SEG_T=256 python tools/quick_bench_seg.py 740 375 500 ${1:-80} 2>&1 | tail -3 | tee -a gpurun_out/seg_qb.log
