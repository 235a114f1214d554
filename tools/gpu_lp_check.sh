#!/usr/bin/env bash
# developer helper: LP parity tests + quick timing probes on a GPU box (run through gpurun)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_lp_parity_gpu.py tests/test_lp_edge_cases_gpu.py tests/test_l2f_device_gpu.py -x -q -m gpu > gpurun_out/lp_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/lp_tests.log
tail -5 gpurun_out/lp_tests.log
python tools/quick_bench.py 2072 600 2>&1 | tail -3 | tee gpurun_out/qb.log
LPBOX_PLAIN_SLOTS=1 python tools/quick_bench.py 2072 600 2>&1 | tail -3 | tee -a gpurun_out/qb.log
