#!/usr/bin/env bash
# developer helper: driver-format bench lines of every configuration at its full size -> gpurun_out/r02_bench_<config>.json
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name $*"; timeout 1500 python bench.py "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; echo "exit $?"; python - <<PY
import json
try:
    l = json.load(open("gpurun_out/r02_bench_$name.json"))
    print({k: l.get(k) for k in ("metric", "value", "unit", "ms_per_step")}, "e2e", l.get("e2e", {}).get("value"), "parity", l.get("parity"), "cpu", (l.get("cpu_baseline") or {}).get("value"), "roofline frac", (l.get("roofline") or {}).get("frac"))
    for k in ("quality", "fast_accuracy"):
        if k in l: print("  ", k, l[k])
except Exception as e:
    print("no line:", e)
PY
tail -2 gpurun_out/r02_bench_$name.err; }
for c in "$@"; do
  case $c in
    lp_plain) run lp_plain ;;
    lp_fast) run lp_fast --config lp_fast ;;
    lp_large) run lp_large --config lp_large ;;
    lp_l2f) run lp_l2f --config lp_l2f ;;
    seg) run seg --config seg --steps 2 --warmup 1 ;;
    sa) run sa --config sa ;;
    policy) run policy --config policy --steps 10 ;;
  esac
done
