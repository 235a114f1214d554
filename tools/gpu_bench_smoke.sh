#!/usr/bin/env bash
# developer helper: every bench configuration once at a small size (run through gpurun)
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout 600 python bench.py "$@" > gpurun_out/smoke_$name.json 2> gpurun_out/smoke_$name.err; echo "exit $?"; head -c 1500 gpurun_out/smoke_$name.json; echo; tail -3 gpurun_out/smoke_$name.err; }
run lp_plain --batch 1036 --steps 1 --warmup 1 --e2e-steps 1
run lp_fast --config lp_fast --batch 1036 --steps 1 --warmup 1 --e2e-steps 1
run lp_large --config lp_large --batch 296 --steps 1 --warmup 1 --e2e-steps 1
run lp_l2f --config lp_l2f --batch 1036 --steps 1 --warmup 1 --e2e-steps 1
run seg --config seg --batch 16 --steps 1 --warmup 1 --e2e-steps 1
run sa --config sa --batch 256 --steps 1 --warmup 1 --e2e-steps 1
run policy --config policy --batch 65536 --steps 1 --warmup 1 --e2e-steps 1
