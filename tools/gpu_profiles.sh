#!/usr/bin/env bash
# developer helper: the ncu captures behind profiles/r02_* (run through gpurun, one GPU)
# usage: gpu_profiles.sh [lp] [seg]   (default: both)
set -u
mkdir -p gpurun_out
WHAT="${*:-lp seg}"
if [[ " $WHAT " == *" lp "* ]]; then
# 1. full-set capture of the window kernel on the 1036 x 60 probe (source view included)
ncu --set full --import-source on --clock-control none -k regex:lp_admm_window -c 1 -o gpurun_out/r02_lp_window -f python tools/quick_bench.py 1036 60 > gpurun_out/r02_ncu_window.log 2>&1
# 2. metrics-only pass over the real bench launch (10,000 instances to convergence): DRAM / L2 bytes, shared wavefronts, duration
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum \
    --clock-control none -k regex:lp_admm_window -c 1 --csv --log-file gpurun_out/r02_window_traffic.csv python tools/quick_bench.py 10000 20000 gen > gpurun_out/r02_ncu_traffic.log 2>&1
# 3. launch list of a whole (short) bench run
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --batch 2072 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
fi
if [[ " $WHAT " == *" seg "* ]]; then
# 4. segmentation kernel: full-set capture of a full wave (740 images x 30 iterations, 256-thread shape), and a metrics-only pass over
#    the bench's own launch (1024 images to convergence, 192-thread shape, 64 distinct images)
ncu --set full --import-source on --clock-control none -k regex:seg_admm -c 1 -o gpurun_out/r02_seg_kernel -f python tools/quick_bench_seg.py 740 375 500 30 > gpurun_out/r02_ncu_seg.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max,smsp__inst_executed.sum \
    --clock-control none -k regex:seg_admm -c 1 --csv --log-file gpurun_out/r02_seg_traffic.csv python tools/quick_bench_seg.py 1024 375 500 10000 64 > gpurun_out/r02_ncu_segtraffic.log 2>&1
fi
for f in window traffic launches seg segtraffic; do tail -n 2 gpurun_out/r02_ncu_$f.log; done
