#!/usr/bin/env bash
# developer helper: launch-order experiment (nnz-descending vs natural order) on the bench's own batch
set -u
mkdir -p gpurun_out; : > gpurun_out/sort_qb.log
for B in 10000 1250; do
  echo "== B=$B sorted" | tee -a gpurun_out/sort_qb.log
  python tools/quick_bench.py $B 20000 gen 2>&1 | tail -2 | tee -a gpurun_out/sort_qb.log
  echo "== B=$B natural order" | tee -a gpurun_out/sort_qb.log
  LPBOX_NO_SORT=1 python tools/quick_bench.py $B 20000 gen 2>&1 | tail -2 | tee -a gpurun_out/sort_qb.log
done
python -m pytest tests/test_lp_parity_gpu.py tests/test_lp_edge_cases_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/sort_qb.log
